#!/bin/bash
# Round-2 ninth GPU call: conv_tc2 with the k-steps of a tap alternating between accumulator column ranges.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_edvr_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2i_pytest.log | cut -c1-300
for prec in bf16x3 bf16 tf32; do
  for shape in "5 176 320" "5 44 80" "1 44 80"; do
    echo -n "conv $shape: "; timeout 120 python tools/one_conv.py $shape 64 64 3 --precision $prec 2>&1 | tail -1
  done
done
timeout 400 python bench.py --steps 30 --warmup 5 --no-reference-cuda --no-cpu-baseline --no-parity 2>gpurun_out/r2i_bench.err | tail -1 > gpurun_out/r2i_bench.json
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2i_bench.json').read())
    print('value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
    print('   roofline', d['roofline']['frac'], d['roofline']['launch_us'], 'inner', d['roofline_inner']['launch_us'], 'dcn', d['roofline_dcn']['frac'], d['roofline_dcn']['launch_us'], 'launches', d['gpu_launches'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2i_bench.err').read()[-2000:])
PY
