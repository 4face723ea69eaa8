"""Writes profiles/traffic.json — DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu) of the two dominant kernels of
bench.py's step, measured on the commit it records. Run on the GPU box from the repo root:

    python scripts/ncu_traffic.py            (about a minute; bench.py reads the file for roofline.traffic)

The ncu pass replays kernels with cold caches and serialised launches: only the byte counts are used, never a time."""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CMD = [sys.executable, "bench.py", "--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--no-parity", "--no-cfg4", "--no-formats"]
NCU = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:sweep_tma_kernel|sweep_chain_kernel|occlusion_kernel|raymarch_fast", "-c", "40", "--csv"]
out = subprocess.run(NCU + CMD, cwd=ROOT, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):]))) if '"ID"' in out else []
if not rows:
    sys.exit("ncu produced no CSV:\n" + out[-2000:])
h = rows[0]
ik, im, iv, iu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
acc = {}
for r in rows[1:]:
    if len(r) <= iv:
        continue
    name = r[ik].split("(")[0].replace("void ", "").split("<")[0].strip()
    v = float(r[iv].replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[iu], 1)
    acc.setdefault(name, {}).setdefault(r[im], []).append(v * mult)
traffic = {}
for k, m in acc.items():
    rd, wr = m.get("dram__bytes_read.sum", []), m.get("dram__bytes_write.sum", [])
    n = min(len(rd), len(wr))
    if n:
        traffic[k] = int(sum(rd[:n]) / n + sum(wr[:n]) / n)
        traffic[k + ".launches"] = n
try:
    commit = subprocess.run(["git", "rev-parse", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip() or None
except OSError:
    commit = None
lib = ROOT / "tbraymarcherplugin_b200" / "libtbrm.so"
import hashlib
traffic["_meta"] = {"commit": commit, "libtbrm_sha256_16": hashlib.sha256(lib.read_bytes()).hexdigest()[:16] if lib.exists() else None,
                    "command": " ".join(NCU + CMD[1:]), "note": "bytes per launch = mean over the captured launches of dram__bytes_read.sum + dram__bytes_write.sum"}
(ROOT / "profiles" / "traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
print(json.dumps(traffic))
