#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_ref_fullsize.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/c3_tests.txt
tail -3 gpurun_out/c3_tests.txt
timeout 300 python -m pytest tests/test_gpu_slab.py -q -m gpu 2>&1 | tail -12 > gpurun_out/c3_tests_slab.txt
tail -4 gpurun_out/c3_tests_slab.txt
TBRM_TEST_SLAB_TIMEOUT_MS=1500 timeout 120 python scripts/debug_slab.py 256,256,256 8 scaled_rotated 2>&1 | tail -6
TBRM_SWEEP_PX=2 TBRM_TEST_SLAB_TIMEOUT_MS=1500 timeout 120 python scripts/debug_slab.py 256,256,256 8 scaled_rotated 2>&1 | tail -6
rm -f gpurun_out/c3_ab.jsonl
for n in 256 512; do
  for gen in 1 2; do TBRM_SWEEP_GEN=$gen timeout 120 python scripts/time_sweep_ab.py $n >> gpurun_out/c3_ab.jsonl 2>> gpurun_out/c3_ab.err; done
done
TBRM_SWEEP_GEN=2 TBRM_SWEEP_PX=2 timeout 120 python scripts/time_sweep_ab.py 256 >> gpurun_out/c3_ab.jsonl 2>> gpurun_out/c3_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/c3_ab.jsonl'):
    d = json.loads(l); print(d['volume'], d['env'], 'reset', round(d['reset_2_lights']['ms_min'], 3), 'frame', round(d['frame']['ms_min'], 3))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_ws -s 4 -c 1 -o gpurun_out/c3_sweep_ws -f python scripts/prof_sweep.py 512 > gpurun_out/c3_ncu.log 2>&1
