#!/bin/bash
# Round-2 eighth GPU call: parity of the coalesced epilogues (conv_tc2, conv_tc, DCN forward), graph-timed kernels, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log | cut -c1-300
for s in 1.0 3.0; do echo -n "dcn fwd offset std $s: "; python tools/one_dcn.py 5 176 320 --offset-std $s | tail -1; done
for prec in bf16x3 bf16; do
  for shape in "5 176 320" "5 44 80" "1 44 80" "1 176 320"; do
    echo -n "conv $shape: "; timeout 120 python tools/one_conv.py $shape 64 64 3 --precision $prec 2>&1 | tail -1
  done
done
timeout 400 python bench.py --steps 30 --warmup 5 --no-reference-cuda --no-cpu-baseline 2>gpurun_out/r2h_bench.err | tail -1 > gpurun_out/r2h_bench.json
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2h_bench.json').read())
    print('value %.2f e2e %.2f ms %.3f parity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['rel_l2']))
    print('   roofline', d['roofline']['frac'], d['roofline']['launch_us'], 'inner', d['roofline_inner']['launch_us'], 'dcn', d['roofline_dcn']['frac'], d['roofline_dcn']['launch_us'], 'launches', d['gpu_launches'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2h_bench.err').read()[-2000:])
PY
