#!/bin/bash
# Round 2, first gpurun call (about 8 minutes of box time): everything written after round 1's GPU budget was spent, validated and timed in one go.
#   gpurun --timeout 1500 -- 'bash scripts/r2_first_call.sh'
# Results land in gpurun_out/r2_*; DESIGN.md §6 "Code newer than every number above" lists what is being measured.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
for n in 256 512; do
  for px in 2 1 auto; do TBRM_SWEEP_PX=$px python scripts/time_sweep_ab.py $n >> gpurun_out/r2_ab.jsonl 2>> gpurun_out/r2_ab.err; done
  TBRM_RAYMARCH_ADDR64=1 python scripts/time_sweep_ab.py $n >> gpurun_out/r2_ab.jsonl 2>> gpurun_out/r2_ab.err
  TBRM_RAYMARCH_V2=1 python scripts/time_sweep_ab.py $n >> gpurun_out/r2_ab.jsonl 2>> gpurun_out/r2_ab.err
done
TBRM_SWEEP_PX=auto python -m pytest tests -q -m gpu -k "golden or parity or slab or fullsize or pixels" 2>&1 | tail -8 > gpurun_out/r2_tests_px_auto.txt
python scripts/time_f_rows.py 512 > gpurun_out/r2_f_rows.txt 2>&1
TBRM_MANDELBULB_TRIG=1 python scripts/time_f_rows.py 512 > gpurun_out/r2_f_rows_trig.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -3 gpurun_out/r2_tests.txt; cat gpurun_out/r2_smoke.txt | tail -1; cat gpurun_out/r2_bench_n1.json | cut -c1-400; cat gpurun_out/r2_ab.jsonl
