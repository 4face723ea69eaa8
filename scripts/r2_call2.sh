#!/bin/bash
# Round 2, GPU call 2: the warp-specialised sweep (sweep_ws_kernel) — parity on the GPU, A/B against the first generation, one ncu --set full
# capture, the bench line. Every step under its own timeout (an mbarrier protocol bug would spin).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_slab.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/c2_tests.txt
tail -3 gpurun_out/c2_tests.txt
for n in 256 512; do
  for gen in 1 2; do TBRM_SWEEP_GEN=$gen timeout 120 python scripts/time_sweep_ab.py $n >> gpurun_out/c2_ab.jsonl 2>> gpurun_out/c2_ab.err; done
  TBRM_SWEEP_GEN=2 TBRM_SWEEP_PX=2 timeout 120 python scripts/time_sweep_ab.py $n >> gpurun_out/c2_ab.jsonl 2>> gpurun_out/c2_ab.err
  TBRM_SWEEP_GEN=2 TBRM_SWEEP_PX=1 timeout 120 python scripts/time_sweep_ab.py $n >> gpurun_out/c2_ab.jsonl 2>> gpurun_out/c2_ab.err
done
python - <<'PY'
import json
for l in open('gpurun_out/c2_ab.jsonl'):
    d = json.loads(l); print(d['volume'], d['env'], 'reset', round(d['reset_2_lights']['ms_min'], 3), 'frame', round(d['frame']['ms_min'], 3))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_ws -s 4 -c 1 -o gpurun_out/c2_sweep_ws -f python scripts/prof_sweep.py 512 > gpurun_out/c2_ncu.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/c2_bench_n1.json 2> gpurun_out/c2_bench_n1.err
cut -c1-300 gpurun_out/c2_bench_n1.json
