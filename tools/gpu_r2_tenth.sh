#!/bin/bash
# Round-2 tenth GPU call: DCN forward with L1 prefetch of the next tap's out-of-window corners; small-Co weight-gradient kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log | cut -c1-300
for s in 0.3 1.0 1.5 3.0; do echo -n "dcn fwd offset std $s: "; python tools/one_dcn.py 5 176 320 --offset-std $s | tail -1; done
timeout 400 python bench.py --steps 36 --warmup 6 --no-reference-cuda --no-cpu-baseline --no-parity 2>gpurun_out/r2k_bench.err | tail -1 > gpurun_out/r2k_bench.json
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2k_bench.json').read())
    print('value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
    print('   roofline', d['roofline']['frac'], d['roofline']['launch_us'], 'inner', d['roofline_inner']['launch_us'], 'dcn', d['roofline_dcn']['frac'], d['roofline_dcn']['launch_us'], 'launches', d['gpu_launches'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2k_bench.err').read()[-2000:])
PY
