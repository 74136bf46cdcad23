#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one adapted frame (eager launches, same kernels as the graph),
# (2) full-section capture of the dominant kernel
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4200 -c 1500 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/launches_r1.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 3 -c 1 -o gpurun_out/prof_tc2_r1 python tools/one_conv.py 5 176 320 64 64 3 > gpurun_out/ncu_tc2.log 2>&1
ls -la gpurun_out/*.ncu-rep
