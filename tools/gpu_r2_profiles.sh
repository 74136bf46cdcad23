#!/bin/bash
# Round-2 evidence for profiles/ (code state at the end of the round): bench.py's own launch list under ncu (graph nodes), ncu --set
# full of the dominant kernels, SASS mnemonic evidence is produced on the build host (tools/sass_evidence.sh).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 16000 --csv --log-file gpurun_out/r2j_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-parity --no-roofline > gpurun_out/r2j_bench_under_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2j_bench_launches.csv
python tools/ncu_summary.py gpurun_out/r2j_bench_launches.csv > gpurun_out/r2j_bench_launches_summary.md 2>/dev/null; head -14 gpurun_out/r2j_bench_launches_summary.md
cap() {  # name regex skip count script...
  n=$1; rg=$2; sk=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$rg" -s $sk -c 1 -o gpurun_out/r2j_prof_$n -f "$@" > gpurun_out/r2j_ncu_$n.log 2>&1; tail -1 gpurun_out/r2j_ncu_$n.log
}
cap tc2_x3 conv_tc2 3 python tools/one_conv.py 5 176 320 64 64 3 --precision bf16x3
cap tc2_bf16 conv_tc2 3 python tools/one_conv.py 5 176 320 64 64 3 --precision bf16
cap mdcn_fwd mdcn_tcs 3 python tools/one_dcn.py 5 176 320 --offset-std 1.0
ls -la gpurun_out/r2j_*.ncu-rep | awk '{print $5, $9}'
