#!/bin/bash
# two CTAs per SM for conv_tc2 in the bf16 mode (DVSR_T2_OCC2=0 switches it off): parity, kernel times, frame rate
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_fullsize_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "single_product or conv or baseline_shapes" > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2n_pytest.log | cut -c1-200
for o in 0 1 0 1; do for shape in "5 176 320" "5 44 80" "1 44 80" "1 176 320"; do
  echo -n "occ2=$o conv $shape: "; DVSR_T2_OCC2=$o timeout 120 python tools/one_conv.py $shape 64 64 3 --precision bf16 2>&1 | tail -1
done; done
for o in 0 1 0 1; do
DVSR_T2_OCC2=$o timeout 300 python bench.py --steps 36 --warmup 6 --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('occ2=$o: value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
DVSR_T2_OCC2=1 timeout 300 python bench.py --steps 36 --warmup 6 --pipelines 1 --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('occ2=1 p1: value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
DVSR_T2_OCC2=0 timeout 300 python bench.py --steps 36 --warmup 6 --pipelines 1 --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('occ2=0 p1: value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
