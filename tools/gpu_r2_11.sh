#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "mdcn" > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2l_pytest.log | cut -c1-300
for s in 0.3 1.0 1.5 3.0; do echo -n "dcn fwd offset std $s: "; timeout 60 python tools/one_dcn.py 5 176 320 --offset-std $s | tail -1; done
