#!/bin/bash
# compute-sanitizer memcheck over one representative test of every kernel family touched in round 2 (small shapes)
mkdir -p gpurun_out
K="tc_3x3_64 or tc_cat2 or tc_offmask216 or tc_shuffle or tc_4x4s2_p1 or tc_last_64to3_res or mdcn_tensor_core_forward or mdcn_tensor_core_backward or mdcn_nhwc or sliced_peer or single_product or fused_updates or tsa"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "$K" > gpurun_out/r2_sanitize_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_sanitize_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|rc=|Invalid|Error" gpurun_out/r2_sanitize_memcheck.log | head -20
