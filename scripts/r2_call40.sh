#!/bin/bash
# G8 ChangeDirLight and 7-row G8 tiles through the TMA-staged sweep: parity tests, then the format timings of the bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zzz_gpu_more.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-cfg4 > gpurun_out/c40_bench.json 2> gpurun_out/c40_bench.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/c40_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['parity']); print(json.dumps(d['formats']))"
