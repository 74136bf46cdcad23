#!/bin/bash
# first GPU contact: parity tests (no -x), smoke, kernel timings, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 600 python tools/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; tail -5 gpurun_out/kernel_bench.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_adapt.log 2>&1; tail -3 gpurun_out/bench_adapt.log
timeout 600 python bench.py --steps 5 --warmup 3 --workload infer --no-cpu-baseline > gpurun_out/bench_infer.log 2>&1; tail -3 gpurun_out/bench_infer.log
