#!/bin/bash
# tile rows of the first-generation sweep: 8 (round 1), 7, 6, automatic — cfg2 reset of two lights, torch-free timing
for th in 8 7 auto; do
  echo "== TBRM_SWEEP_TH=$th"
  TBRM_SWEEP_TH=$th timeout 120 python scripts/time_sweep_ab.py 512 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['reset_2_lights']))"
done
for n in 256 384 640; do
  for th in 8 auto; do
    echo "== N=$n TBRM_SWEEP_TH=$th"
    TBRM_SWEEP_TH=$th timeout 120 python scripts/time_sweep_ab.py $n 640 360 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['reset_2_lights']))"
  done
done
timeout 600 python -m pytest tests/test_zzz_gpu_more.py -x -q -k "one_and_two_pixels" 2>&1 | tail -3
