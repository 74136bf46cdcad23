#!/bin/bash
# conv_tc2 epilogue width A/B: 4 warps x 64 columns vs 8 warps x 32 columns, per precision and launch size; parity of both.
mkdir -p gpurun_out
for e in 4 8; do
  DVSR_T2_EPI=$e timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "conv or tc" > gpurun_out/r2m_pytest_epi$e.log 2>&1
  echo "epi $e pytest rc=$?"; tail -1 gpurun_out/r2m_pytest_epi$e.log | cut -c1-200
done
for prec in bf16x3 bf16; do for e in 4 8; do
  for shape in "5 176 320" "1 176 320" "5 44 80" "1 44 80"; do
    echo -n "epi $e conv $shape: "; DVSR_T2_EPI=$e timeout 120 python tools/one_conv.py $shape 64 64 3 --precision $prec 2>&1 | tail -1
  done
done; done
for e in 4 8; do
DVSR_T2_EPI=$e timeout 300 python bench.py --steps 36 --warmup 6 --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('epi $e: value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
DVSR_T2_EPI=$e timeout 300 python bench.py --steps 36 --warmup 6 --workload infer --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('epi $e infer: value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
