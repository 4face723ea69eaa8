#!/bin/bash
# multi-GPU: bit-parity tests of the sharded volume + the bench line (cfg2 step, streaming e2e, scale_cfg4) on N GPUs of one box
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu ${2:+-k "$2"} 2>&1 | tail -5 > gpurun_out/multi_tests_n$N.txt; tail -3 gpurun_out/multi_tests_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
    print('N', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e ms', round(d['e2e']['ms_per_step'], 3), 'serial', round(d['e2e']['serial_ms_per_step'], 3))
    print('stages', {k: (round(v['ms'], 3) if isinstance(v, dict) else round(v, 3)) for k, v in d['stages'].items() if k != 'ray_steps_per_frame'})
    print('parity', d['parity']); print('cfg4', {k: d['scale_cfg4'][k] for k in ('sweep_ms', 'Mvoxels_per_s', 'parity')})
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/bench_n$N.err').read()[-1500:])
PY
