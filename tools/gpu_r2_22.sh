#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_edvr_gpu.py tests/test_fullsize_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2r_pytest.log | cut -c1-200
for i in 1 2; do
timeout 300 python bench.py --steps 36 --warmup 6 --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
timeout 300 python bench.py --steps 36 --warmup 6 --pipelines 1 --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('p1 value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
