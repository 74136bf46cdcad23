#!/bin/bash
# session re-entry check: parity tests, smoke, bench (both workloads), reference arm, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_adapt.log 2>&1; tail -1 gpurun_out/bench_adapt.log | cut -c1-2500
timeout 600 python bench.py --steps 20 --warmup 3 --workload infer --no-cpu-baseline > gpurun_out/bench_infer.log 2>&1; tail -1 gpurun_out/bench_infer.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4200 -c 1500 --csv --log-file gpurun_out/launches_r1b.csv \
    python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/launches_r1b.csv
timeout 300 python tools/profile_step.py > gpurun_out/profile_step.txt 2>&1; tail -5 gpurun_out/profile_step.txt
