"""Summarise an .ncu-rep (one kernel launch) into text: key metrics, stall reasons, hottest source lines.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [voxels_or_steps_per_launch] > profiles/xxx.txt"""
import csv, io, subprocess, sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit_row = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name"))
    for k in KEYS:
        if k in d:
            print(f"  {k:70s} {d[k]:>18s} {unit_row[hdr.index(k)]}")
    if units:
        inst = float(d["smsp__inst_executed.sum"].replace(",", ""))
        print(f"  warp-instructions per unit ({units:.0f} units/launch): {inst / units:.3f}  (thread-instr/unit ~ {inst * 32 / units:.1f})")
    st = sorted(((float(d[h].replace(",", "")) if d[h] else 0.0, h) for h in hdr if "issue_stalled" in h and "per_issue_active" in h), reverse=True)
    print("  stall reasons (warps per issue-active cycle):")
    for v, h in st[:8]:
        print(f"    {v:6.2f}  {h.split('stalled_')[1].replace('_per_issue_active.ratio', '')}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, h2, out = None, None, []
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
        if n:
            out.append((n, s, cur, r[0], r[1].strip()[:110]))
tot, ts = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print("  hottest source lines (share of executed warp-instructions, share of stall samples):")
for n, s, f, l, t in sorted(out, reverse=True)[:25]:
    print(f"    {100 * n / tot:5.1f}% {100 * s / ts:5.1f}%  {f}:{l}  {t}")
