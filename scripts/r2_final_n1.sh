#!/bin/bash
# Final N = 1 pass of round 2: DRAM traffic of the shipped kernels first (the bench line reads profiles/traffic.json), then the whole GPU
# suite, smoke(), the bench line, ncu captures of the shipped sweep / march instantiations, the launch list, the sanitizer passes
mkdir -p gpurun_out
timeout 300 python scripts/ncu_traffic.py > gpurun_out/f_traffic.log 2>&1; tail -2 gpurun_out/f_traffic.log | cut -c1-300; cp profiles/traffic.json gpurun_out/f_traffic.json
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/f_tests.txt; tail -3 gpurun_out/f_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; cut -c1-200 gpurun_out/f_bench_n1.json; tail -3 gpurun_out/f_bench_n1.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; cut -c1-300 gpurun_out/f_bench_ref.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_tma_kernel -s 4 -c 1 -o gpurun_out/f_sweep_tma -f python scripts/prof_sweep.py 512 > gpurun_out/f_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:raymarch_fast -s 1 -c 1 -o gpurun_out/f_raymarch -f python scripts/prof_raymarch.py > gpurun_out/f_ncu2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-cfg4 --no-formats > /dev/null 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_target.py > gpurun_out/f_sanitizer_$tool.txt 2>&1; echo "$tool rc=$?"; tail -2 gpurun_out/f_sanitizer_$tool.txt | cut -c1-200
done
ls -la gpurun_out/f_*
