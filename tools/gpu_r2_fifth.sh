#!/bin/bash
# Round-2 fifth GPU call: re-check the DCN backward test with failure messages, the whole GPU suite, and the in-kernel timeline of
# conv_tc2 in its three precisions (which stage bounds the single-product mode?).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "mdcn_tensor_core_backward" > gpurun_out/r2e_pytest_mdcn.log 2>&1
echo "pytest mdcn rc=$?"; grep -E "^E  |passed|failed" gpurun_out/r2e_pytest_mdcn.log | cut -c1-200 | head -20
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2e_pytest_gpu.log 2>&1
echo "pytest all rc=$?"; tail -5 gpurun_out/r2e_pytest_gpu.log | cut -c1-300
for prec in bf16x3 bf16 tf32; do
  echo "== conv_tc2 trace $prec"; timeout 120 python tools/one_conv.py 5 176 320 64 64 3 --trace --precision $prec 2>&1 | tail -14
done
