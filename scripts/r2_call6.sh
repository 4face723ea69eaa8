#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_ref_fullsize.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/c6_tests.txt
tail -3 gpurun_out/c6_tests.txt
timeout 300 python -m pytest tests/test_gpu_slab.py -q -m gpu 2>&1 | tail -12 > gpurun_out/c6_tests_slab.txt
tail -4 gpurun_out/c6_tests_slab.txt
rm -f gpurun_out/c6_ab.jsonl
for n in 256 512; do
  for gen in 1 2; do TBRM_SWEEP_GEN=$gen timeout 120 python scripts/time_sweep_ab.py $n >> gpurun_out/c6_ab.jsonl 2>> gpurun_out/c6_ab.err; done
done
TBRM_SWEEP_GEN=2 TBRM_SWEEP_PX=2 timeout 120 python scripts/time_sweep_ab.py 256 >> gpurun_out/c6_ab.jsonl 2>> gpurun_out/c6_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/c6_ab.jsonl'):
    d = json.loads(l); print(d['volume'], d['env'], 'reset', round(d['reset_2_lights']['ms_min'], 3), 'frame', round(d['frame']['ms_min'], 3))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"occlusion_kernel|sweep_chain_kernel" -s 8 -c 2 -o gpurun_out/c6_sweep_ws -f python scripts/prof_sweep.py 512 > gpurun_out/c6_ncu.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c6_launches.csv python scripts/prof_sweep.py 512 > /dev/null 2>&1
grep -E "occlusion|sweep_chain|fill" gpurun_out/c6_launches.csv | awk -F'","' '{print $5, $NF}' | tail -12
