"""Executed warp-instructions and stall samples of an .ncu-rep (one launch, --import-source on) summed over regions of sweep_ws_kernel.cuh.
usage: python scripts/ncu_regions.py rep.ncu-rep [units]"""
import csv, io, subprocess, sys, collections, re
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
K = "sweep_ws_kernel.cuh"
srcl = (ROOT / "tbraymarcherplugin_b200" / "csrc" / K).read_text().splitlines()
def find(s):
    return next(n for n, l in enumerate(srcl, 1) if s in l)
bounds = [
    (1, "helpers / prologue"),
    (find("ws_poll_cell(const unsigned long long* cell"), "poll slow paths"),
    (find("sweep_ws_kernel(const __grid_constant__"), "kernel prologue (all threads)"),
    (find("CONSUMERS: the propagation chain"), "consumer set-up"),
    (find("int store_pending = -1"), "consumer: block head (waits, prefetch)"),
    (find("(a) issue the halo loads"), "consumer slice: (a) halo loads, probe, T load"),
    (find("(c) the halo of slice k-1 must have arrived"), "consumer slice: (c) halo check / park, barrier, flag, TMA store"),
    (find("(d) propagate, forward, export"), "consumer slice: (d) propagate, forward, export"),
    (find("the T block is free again"), "consumer: block end (light brick update)"),
    (find("PRODUCERS: transmission factors"), "producer set-up"),
    (find("per-slice sampler constants of the block"), "producer: block head"),
    (find("pass 1 (cheap)"), "producer unit: pass 1 (empty-space test)"),
    (find("pass 2: walk the planes"), "producer unit: pass 2 (decode, interpolate, opacity)"),
]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, h2 = None, None
inst, samp = collections.Counter(), collections.Counter()
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 3 and r[0] == "Line No":
        h2 = r
        ie, isamp = h2.index("Instructions Executed"), h2.index("# Samples")
    elif cur and h2 and len(r) > ie and r[0].isdigit() and r[2] == "-":
        try:
            n, s = int(r[ie]), int(r[isamp])
        except ValueError:
            continue
        if cur == K:
            lab = [b for ln, b in bounds if ln <= int(r[0])][-1]
        else:
            lab = "inlined: " + cur
        inst[lab] += n
        samp[lab] += s
tot, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print(f"total warp-instructions {tot}" + (f" = {tot * 32 / units:.1f} thread-instr per unit" if units else ""))
for lab, n in inst.most_common():
    print(f"  {100 * n / tot:5.1f}% instr  {100 * samp[lab] / ts:5.1f}% stall samples  " + (f"{n * 32 / units:6.1f}/unit  " if units else "") + lab)
