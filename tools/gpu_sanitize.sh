#!/bin/bash
# compute-sanitizer memcheck over one representative test of every tensor-core / gather kernel family (small shapes)
mkdir -p gpurun_out
K="tc_3x3_64 or tc_cat2 or tc_offmask216 or tc_shuffle or tc_4x4s2_p1 or tc_3x3s2_odd or tc_last_64to3_res or conv3d_rgb or mdcn_nhwc or legacy or test_upsample or pool or pad2d or tsa or pixel_loss or fused_updates"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_ops_gpu.py tests/test_degradation.py -m gpu -q -x -p no:cacheprovider -k "$K or cuda_degradation_matches" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|rc=|Invalid|Error" gpurun_out/sanitize_memcheck.log | head -20
