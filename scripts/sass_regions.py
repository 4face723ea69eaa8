"""Static SASS of sweep_ws_kernel<AXIS, CLIP, SLAB, PX> attributed to source lines and summed over line ranges of sweep_ws_kernel.cuh
(needs -lineinfo): python scripts/sass_regions.py [AXIS=2] [CLIP=0] [SLAB=0] [PX=2] [top=30]"""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
axis, clip, slab, px, top = (int(a) for a in (sys.argv[1:] + ["2", "0", "0", "2", "30"][len(sys.argv) - 1:])[:5])
K = "sweep_ws_kernel.cuh"
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "tbraymarcherplugin_b200" / "build" / "sweep.cu.o")], cwd=tmp, check=True, capture_output=True)
    cubin = next(Path(tmp).glob("*.cubin"))
    text = subprocess.run(["nvdisasm", "--print-line-info", str(cubin)], capture_output=True, text=True, check=True).stdout
name = f".text._ZN4tbrm15sweep_ws_kernelILi{axis}ELb{clip}ELb{slab}ELi{px}E"
lines = text.splitlines()
begin = next(i for i, l in enumerate(lines) if l.startswith(name))
seq, cur = [], None
for l in lines[begin + 1:]:
    if l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        seq.append((int(m.group(1), 16), cur, m.group(3).split(".")[0]))
src = (ROOT / "tbraymarcherplugin_b200" / "csrc" / K).read_text().splitlines()
marks = [(n, s.strip()) for n, s in enumerate(src, 1) if "====" in s or s.strip().startswith("// ----") or "#pragma unroll 1" in s or "for (int n = 0; n < nblk" in s]
print(f"kernel total {len(seq)} instructions = {len(seq) * 16 / 1024:.1f} KB")
per = collections.Counter()
for a, c, o in seq:
    per[c] += 1
# by position in the instruction stream: runs of consecutive instructions, labelled by the source mark preceding their line
def region(c):
    if not c or c[0] != K:
        return "(inlined helpers: " + (c[0] if c else "?") + ")"
    lab = "prologue"
    for n, s in marks:
        if n <= c[1]:
            lab = f"{n}: {s[:70]}"
    return lab
reg = collections.Counter()
for a, c, o in seq:
    reg[region(c)] += 1
for k, v in sorted(reg.items(), key=lambda kv: -kv[1]):
    print(f"{v:5d}  {k}")
print("--- top lines")
for (c), n in per.most_common(top):
    print(f"{n:4d} {c} | {src[c[1] - 1].strip()[:100] if c and c[0] == K else ''}")
