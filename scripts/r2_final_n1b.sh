#!/bin/bash
# Last N = 1 pass of round 2 (after the march's block shape changed): traffic of the shipped library, the whole GPU suite, smoke(), the bench
# line, an ncu capture of the shipped march instantiation
mkdir -p gpurun_out
timeout 200 python scripts/ncu_traffic.py > gpurun_out/g_traffic.log 2>&1; cp profiles/traffic.json gpurun_out/g_traffic.json
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/g_tests.txt; tail -2 gpurun_out/g_tests.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/g_bench_n1.json 2> gpurun_out/g_bench_n1.err; cut -c1-200 gpurun_out/g_bench_n1.json; tail -3 gpurun_out/g_bench_n1.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:raymarch_fast -s 1 -c 1 -o gpurun_out/g_raymarch -f python scripts/prof_raymarch.py > gpurun_out/g_ncu2.log 2>&1
ls gpurun_out/g_*
