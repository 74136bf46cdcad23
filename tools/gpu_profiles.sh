#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one adapted frame (eager launches, one pipeline: same kernels as the graphs),
# (2) full-section capture of the dominant kernel, (3) the bench lines themselves (not under a profiler)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --warmup 6 > gpurun_out/bench_adapt.json 2> gpurun_out/bench_adapt.err; tail -c 600 gpurun_out/bench_adapt.json
timeout 600 python bench.py --steps 30 --warmup 6 --workload infer --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
timeout 600 python bench.py --steps 30 --warmup 6 --pipelines 1 --no-cpu-baseline > gpurun_out/bench_adapt_p1.json 2> gpurun_out/bench_adapt_p1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4300 -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graphs --pipelines 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 3 -c 1 -o gpurun_out/prof_tc2 -f python tools/one_conv.py 5 176 320 64 64 3 > gpurun_out/ncu_tc2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mdcn_tc -s 2 -c 1 -o gpurun_out/prof_mdcn -f python tools/kernel_bench.py --tc --only-mdcn > gpurun_out/ncu_mdcn.log 2>&1
ls -la gpurun_out/*.ncu-rep
