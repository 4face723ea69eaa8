"""SASS evidence of the TMA / async-copy path per kernel of libtbrm.so: counts of UTMALDG (tensor TMA load), UTMASTG (tensor TMA store), UBLKCP
(1-D bulk copy), SYNCS (mbarrier), LDGSTS (cp.async), BAR, and the kernel's size.  python scripts/sass_tma_counts.py > profiles/r2_sass_tma_counts.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
objs = sorted((ROOT / "tbraymarcherplugin_b200" / "build").glob("*.cu.o"))
KEYS = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "BAR", "LDG", "STG", "LDS", "STS", "FFMA", "IMAD"]
print("# cuobjdump -sass of tbraymarcherplugin_b200/build/*.cu.o (sm_100a); instructions per kernel")
print(f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in KEYS))
for o in objs:
    out = subprocess.run(["cuobjdump", "-sass", str(o)], capture_output=True, text=True).stdout
    name, cnt, total = None, collections.Counter(), 0
    def flush():
        if name and total:
            short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            short = re.sub(r"\(.*", "", short).replace("void ", "").replace("tbrm::", "")
            print(f"{short[:78]:78s} {total:6d} " + " ".join(f"{cnt[k]:7d}" for k in KEYS))
    for l in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            flush()
            name, cnt, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            total += 1
            op = m.group(1)
            for k in KEYS:
                if op == k or (k in ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS") and op.startswith(k)):
                    cnt[k] += 1
    flush()
