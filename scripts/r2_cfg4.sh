#!/bin/bash
# the north-star configuration (BASELINE.json configs[3]): 1024^3 / 2160p / 768 steps / 3 lights, whole step, on N GPUs
N=${1:-8}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_n$N.json 2> gpurun_out/bench_cfg4_n$N.err
fi
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_cfg4_n$N.json').read().strip().splitlines()[-1])
    print('cfg4 N', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e ms', round(d['e2e']['ms_per_step'], 3))
    print('stages', {k: (round(v['ms'], 3) if isinstance(v, dict) else round(v, 3)) for k, v in d['stages'].items() if k != 'ray_steps_per_frame'})
    print('parity', d['parity'])
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/bench_cfg4_n$N.err').read()[-1500:])
PY
