#!/bin/bash
# 7-row tiles in banded / Z-sharded passes: slab tests, 640^3 and the bench line with its 1024^3 sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -3
for th in 8 auto; do
  echo "== N=640 TBRM_SWEEP_TH=$th"
  TBRM_SWEEP_TH=$th timeout 120 python scripts/time_sweep_ab.py 640 640 360 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['reset_2_lights']))"
done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c46_bench_n1.json 2> gpurun_out/c46_bench_n1.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/c46_bench_n1.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['parity']); print(json.dumps(d['scale_cfg4'])[:600])"
