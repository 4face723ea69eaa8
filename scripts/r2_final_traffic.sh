#!/bin/bash
# after the last library change of round 2 (a debug getter; the kernels are unchanged): DRAM traffic of the shipped libtbrm.so, the new tile-row test
mkdir -p gpurun_out
timeout 200 python scripts/ncu_traffic.py > gpurun_out/h_traffic.log 2>&1; cp profiles/traffic.json gpurun_out/h_traffic.json; tail -1 gpurun_out/h_traffic.log | cut -c1-200
timeout 200 python -m pytest tests/test_zzz_gpu_more.py -x -q -m gpu -k "seven_row_tiles_and_stays or half_resolution" 2>&1 | tail -2
