#!/bin/bash
# Round-2 fourth GPU call: ncu --set full of the two roofline kernels (traffic numbers for profiles/ncu_traffic.json, stall reasons of
# the restructured DCN gather), the DCN backward kernel, and the wgrad split-K policy sweep.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "mdcn_tensor_core_backward" 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mdcn_tcs -s 3 -c 1 -o gpurun_out/r2_prof_mdcn_fwd -f python tools/one_dcn.py > gpurun_out/r2_ncu_mdcn_fwd.log 2>&1; tail -2 gpurun_out/r2_ncu_mdcn_fwd.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mdcn_bwd_tc -s 1 -c 1 -o gpurun_out/r2_prof_mdcn_bwd -f python tools/one_dcn.py 5 44 80 --bwd > gpurun_out/r2_ncu_mdcn_bwd.log 2>&1; tail -2 gpurun_out/r2_ncu_mdcn_bwd.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 3 -c 1 -o gpurun_out/r2_prof_tc2 -f python tools/one_conv.py 5 176 320 64 64 3 > gpurun_out/r2_ncu_tc2.log 2>&1; tail -2 gpurun_out/r2_ncu_tc2.log
python tools/one_dcn.py 5 176 320 --bwd; python tools/one_dcn.py 5 176 320 --offset-std 0.3; python tools/one_dcn.py 5 176 320 --offset-std 3.0
for wg in 24 48 96; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline --wg-chunks $wg 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('wg-chunks $wg: value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
for p in 4 8; do
  timeout 300 python bench.py --steps 32 --warmup 8 --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline --pipelines $p 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('pipelines $p: value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
ls -la gpurun_out/*.ncu-rep | tail -4
