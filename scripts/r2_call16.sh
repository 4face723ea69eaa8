#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/time_sweep_ab.py 512 | cut -c150-330
timeout 120 python scripts/time_sweep_ab.py 256 | cut -c150-330
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_slab.py tests/test_gpu_zz_materials.py tests/test_zzz_gpu_more.py -x -q -m gpu 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_tma_kernel -s 4 -c 1 -o gpurun_out/c16_sweep_tma -f python scripts/prof_sweep.py 512 > gpurun_out/c16_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:raymarch_fast -s 1 -c 1 -o gpurun_out/c16_raymarch -f python scripts/prof_raymarch.py > gpurun_out/c16_ncu2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c16_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-cfg4 > /dev/null 2>&1
ls -la gpurun_out/c16_*
