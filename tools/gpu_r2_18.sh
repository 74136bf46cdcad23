#!/bin/bash
# staged DCN forward with offsets / masks staged in shared memory by TMA: parity, then same-box A/B against the previous kernel
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_ops_gpu.py tests/test_reference_cuda_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "mdcn or dcn or deform" > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2p_pytest.log | cut -c1-200
cp dynavsr_b200/libdvsr_b200.so /tmp/lib_default.so
for v in default o0_m4_s3 o1_m3_s3 o1_m4_s3 default o0_m4_s3; do
  if [ $v = default ]; then cp /tmp/lib_default.so dynavsr_b200/libdvsr_b200.so; else cp tools/variants/lib_$v.so dynavsr_b200/libdvsr_b200.so; fi
  for s in 1.0 1.5 3.0; do echo -n "$v offset std $s: "; timeout 60 python tools/one_dcn.py 5 176 320 --offset-std $s 2>&1 | tail -1; done
done
cp /tmp/lib_default.so dynavsr_b200/libdvsr_b200.so
