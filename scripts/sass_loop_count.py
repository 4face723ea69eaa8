"""Static SASS of the per-slice loop of sweep_tma_kernel<AXIS, CLIP, SLAB> with the instructions attributed to source lines (needs -lineinfo, which
build.py passes): `python scripts/sass_loop_count.py [AXIS=2] [CLIP=0] [SLAB=0] [top=25] [PX=2] [L8=0] [TH=7]`. Round 1's breakdown is profiles/r1_sweep_sass_breakdown.txt."""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
axis, clip, slab, top, px, l8, th = (int(a) for a in (sys.argv[1:] + ["2", "0", "0", "25", "2", "0", "7"][len(sys.argv) - 1:])[:7])
K = "sweep_tma_kernel.cuh"
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "tbraymarcherplugin_b200" / "build" / "sweep.cu.o")], cwd=tmp, check=True, capture_output=True)
    cubin = next(Path(tmp).glob("*.cubin"))
    text = subprocess.run(["nvdisasm", "--print-line-info", str(cubin)], capture_output=True, text=True, check=True).stdout
name = f".text._ZN4tbrm16sweep_tma_kernelILi{axis}ELb{clip}ELb{slab}ELi{px}ELb{l8}ELi{th}E"  # <AXIS, CLIP, SLAB, PX, L8, TH>
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
first = next(n for n, s in enumerate(src, 1) if "for (int sl = 0; sl < kSB; ++sl)" in s) + 1
last = next(n for n, s in enumerate(src, 1) if "write the updated light brick back" in s) - 1
start = next(i for i, (a, c, o) in enumerate(seq) if c == (K, first))
end = max(i for i, (a, c, o) in enumerate(seq) if c and c[0] == K and first - 1 <= c[1] <= last) + 1
per, ops, tot = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
for a, c, o in seq[start:end]:
    per[c] += 1
    ops[c][o] += 1
    tot[o] += 1
print(f"kernel total {len(seq)} instructions; per-slice loop {end - start} ({hex(seq[start][0])} .. {hex(seq[end][0])})")
for (f, ln), c in sorted(per.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{c:4d} {f}:{ln} {dict(ops[(f, ln)].most_common(4))} | {src[ln - 1].strip()[:90] if f == K else ''}")
print(tot.most_common(30))
