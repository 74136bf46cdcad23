#!/bin/bash
# Round-2 seventh GPU call: conv_tc2 ablations (which resource paces the kernel), dry-run epilogue warm-up, rotated weight loads.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "conv or tc" > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest.log | cut -c1-300
for prec in bf16x3 bf16; do
  for abl in 0 1 2 4 8 3 7 15; do
    echo -n "ablate $abl: "; DVSR_T2_ABLATE=$abl timeout 120 python tools/one_conv.py 5 176 320 64 64 3 --precision $prec 2>&1 | tail -1
  done
done
for prec in bf16x3 bf16; do
  echo "== conv_tc2 trace $prec 5x176x320"; timeout 120 python tools/one_conv.py 5 176 320 64 64 3 --trace --precision $prec 2>&1 | tail -10 | cut -c1-200
done
echo "== conv_tc2 trace bf16 5x44x80"; timeout 120 python tools/one_conv.py 5 44 80 64 64 3 --trace --precision bf16 2>&1 | tail -9 | cut -c1-60
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 3 -c 1 -o gpurun_out/r2g_prof_tc2_bf16 -f python tools/one_conv.py 5 176 320 64 64 3 --precision bf16 > gpurun_out/r2g_ncu_tc2_bf16.log 2>&1; tail -2 gpurun_out/r2g_ncu_tc2_bf16.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 3 -c 1 -o gpurun_out/r2g_prof_tc2_x3 -f python tools/one_conv.py 5 176 320 64 64 3 --precision bf16x3 > gpurun_out/r2g_ncu_tc2_x3.log 2>&1; tail -2 gpurun_out/r2g_ncu_tc2_x3.log
