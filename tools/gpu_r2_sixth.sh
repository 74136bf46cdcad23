#!/bin/bash
# Round-2 sixth GPU call: conflict-free staged DCN gather (unswizzled window, parity-ordered halves) and conv_tc2 with hi-only resident
# weights / deeper ring in the bf16 mode -- parity tests first, then timings and the finer epilogue trace.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_edvr_gpu.py tests/test_fullsize_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2f_pytest.log | cut -c1-300
for s in 1.0 0.3 3.0; do echo "offset std $s"; python tools/one_dcn.py 5 176 320 --offset-std $s; done
for prec in bf16x3 bf16; do
  echo "== conv_tc2 trace $prec 5x176x320"; timeout 120 python tools/one_conv.py 5 176 320 64 64 3 --trace --precision $prec 2>&1 | tail -13
done
echo "== conv_tc2 trace bf16 5x44x80"; timeout 120 python tools/one_conv.py 5 44 80 64 64 3 --trace --precision bf16 2>&1 | tail -13 | cut -c1-80
echo "== conv_tc2 trace bf16x3 1x44x80"; timeout 120 python tools/one_conv.py 1 44 80 64 64 3 --trace --precision bf16x3 2>&1 | tail -13 | cut -c1-80
timeout 300 python bench.py --steps 30 --warmup 5 --no-reference-cuda --no-cpu-baseline 2>gpurun_out/r2f_bench.err | tail -1 > gpurun_out/r2f_bench.json
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2f_bench.json').read())
    print('value %.2f e2e %.2f ms %.3f parity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['rel_l2']))
    print('   roofline', d['roofline']['frac'], d['roofline']['launch_us'], 'inner', d['roofline_inner']['launch_us'], 'dcn', d['roofline_dcn']['frac'], d['roofline_dcn']['launch_us'], 'launches', d['gpu_launches'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2f_bench.err').read()[-2000:])
PY
