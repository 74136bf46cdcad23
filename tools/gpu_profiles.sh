#!/bin/bash
# ncu evidence for profiles/: (1) the bench lines themselves (not under a profiler), (2) launch list of one adapted frame (eager
# launches, one pipeline: same kernels as the graphs), (3) full-section captures of the dominant kernel and of one kernel per
# family named by the north star (DCN fwd / bwd, TSA, weight gradient, fused update, activation backward)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 60 --warmup 12 > gpurun_out/bench_adapt.json 2> gpurun_out/bench_adapt.err; tail -c 300 gpurun_out/bench_adapt.json
timeout 600 python bench.py --steps 60 --warmup 6 --workload infer --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
timeout 600 python bench.py --steps 60 --warmup 6 --pipelines 1 --no-cpu-baseline > gpurun_out/bench_adapt_p1.json 2> gpurun_out/bench_adapt_p1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 1500 --csv --log-file gpurun_out/launches.csv \
    python tools/one_frame.py 3 > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 3 -c 1 -o gpurun_out/prof_tc2 -f python tools/one_conv.py 5 176 320 64 64 3 > gpurun_out/ncu_tc2.log 2>&1
# per-family captures from one eager frame: the LAST instance of each kernel in the frame is the full-resolution one (final forward),
# backward kernels only exist at the SLR resolution
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o gpurun_out/prof_$1 -f python tools/one_frame.py 1 > gpurun_out/ncu_$1.log 2>&1
}
cap mdcn_fwd '^mdcn_tcs_kernel$' 2 1          # staged kernel: only the full-resolution launches use it; #3 = L1 DCN of the final forward
cap tsa '^tsa_temporal_kernel$' 2 1           # 3 per frame: the third is the final forward
cap mdcn_bwd '^mdcn_bwd_data_kernel$' 3 1     # L1 DCN backward of step 1 (5x44x80)
cap wgrad '^conv_wgrad_tc_kernel$' 0 1
cap sgd '^sgd_kernel$' 0 1
cap actbwd '^act_bwd_kernel$' 0 1
ls -la gpurun_out/*.ncu-rep
