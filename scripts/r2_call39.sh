#!/bin/bash
# N = 1: the whole GPU suite and the bench line after the 7-row tiles, the 128-thread march blocks and the half-resolution TMA path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/c39_tests.txt; tail -3 gpurun_out/c39_tests.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c39_bench_n1.json 2> gpurun_out/c39_bench_n1.err; cut -c1-200 gpurun_out/c39_bench_n1.json; tail -3 gpurun_out/c39_bench_n1.err
