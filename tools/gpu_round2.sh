#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py --tc 2>&1 | grep conv3x3 > gpurun_out/kernel_bench_tc.log; cat gpurun_out/kernel_bench_tc.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_adapt_tc.log 2>&1; tail -2 gpurun_out/bench_adapt_tc.log | cut -c1-900
timeout 600 python bench.py --steps 10 --warmup 3 --workload infer --no-cpu-baseline > gpurun_out/bench_infer_tc.log 2>&1; tail -2 gpurun_out/bench_infer_tc.log | cut -c1-400
