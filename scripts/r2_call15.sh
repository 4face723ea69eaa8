#!/bin/bash
# N = 1: the whole GPU suite, bench line, traffic, Mandelbulb mismatch, sanitizer passes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/c15_tests.txt; tail -3 gpurun_out/c15_tests.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c15_bench_n1.json 2> gpurun_out/c15_bench_n1.err; cut -c1-200 gpurun_out/c15_bench_n1.json; tail -3 gpurun_out/c15_bench_n1.err
timeout 300 python scripts/ncu_traffic.py > gpurun_out/c15_traffic.log 2>&1; tail -2 gpurun_out/c15_traffic.log | cut -c1-300
timeout 300 python scripts/mandelbulb_mismatch.py > gpurun_out/c15_mandelbulb_p8.json 2> gpurun_out/c15_mb.err; cat gpurun_out/c15_mandelbulb_p8.json
TBRM_MANDELBULB_TRIG=1 timeout 300 python scripts/mandelbulb_mismatch.py > gpurun_out/c15_mandelbulb_trig.json 2>> gpurun_out/c15_mb.err; cat gpurun_out/c15_mandelbulb_trig.json
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_target.py > gpurun_out/c15_sanitizer_$tool.txt 2>&1; echo "$tool rc=$?"; tail -3 gpurun_out/c15_sanitizer_$tool.txt
done
