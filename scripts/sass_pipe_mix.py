"""Static SASS of a kernel's loop split by execution pipe (ALU: integer add / logic / shift / compare / min-max / conversions, 16 lanes per SM
sub-partition; FMA: FP32 add / mul / fma and IMAD; the rest: memory, control, uniform datapath):
`python scripts/sass_pipe_mix.py <object> <kernel substring> <source file> <first line> <last line>` — the loop is the address range between
the first and the last instruction attributed to those source lines (needs -lineinfo)."""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
obj, kern, srcfile, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
ALU = {"IADD3", "LEA", "SHF", "FMNMX", "I2FP", "I2F", "FSETP", "MOV", "F2I", "VIMNMX", "ISETP", "FRND", "LOP3", "PLOP3", "SEL", "VIADD", "VIADDMNMX",
       "FSEL", "PRMT", "IABS", "P2R", "R2P", "CS2R", "F2FP", "BMSK", "SGXT", "VIMNMX3", "FCHK", "MUFU", "POPC", "FLO"}
FMA = {"FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "FFMA2", "HADD2", "HMUL2"}
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / obj)], cwd=tmp, check=True, capture_output=True)
    text = subprocess.run(["nvdisasm", "--print-line-info", str(next(Path(tmp).glob("*.cubin")))], capture_output=True, text=True, check=True).stdout
lines = text.splitlines()
begin = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
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
idx = [i for i, (a, c, o) in enumerate(seq) if c and c[0] == srcfile and lo <= c[1] <= hi]
start, end = idx[0], idx[-1] + 1
tot = collections.Counter(o for a, c, o in seq[start:end])
n = end - start
alu, fma = sum(v for k, v in tot.items() if k in ALU), sum(v for k, v in tot.items() if k in FMA)
print(f"{lines[begin][6:70]}...: loop {n} instructions ({hex(seq[start][0])} .. {hex(seq[end - 1][0])}): ALU pipe {alu}, FMA pipes {fma} (IMAD {tot['IMAD']}), other {n - alu - fma}")
print("  ", tot.most_common(26))
