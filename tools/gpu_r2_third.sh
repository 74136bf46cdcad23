#!/bin/bash
# Round-2 third GPU call: the tcgen05 DCN backward kernel and the restructured staged DCN forward gather.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "mdcn" > gpurun_out/r2c_pytest_mdcn.log 2>&1
echo "pytest mdcn rc=$?" | tee -a gpurun_out/r2c_pytest_mdcn.log; tail -30 gpurun_out/r2c_pytest_mdcn.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2c_pytest_gpu.log 2>&1
echo "pytest all rc=$?" | tee -a gpurun_out/r2c_pytest_gpu.log; tail -30 gpurun_out/r2c_pytest_gpu.log
timeout 600 python tools/ref_cuda_bench.py > gpurun_out/r2c_ref_cuda_bench.md 2> gpurun_out/r2c_ref_cuda_bench.err; echo "ref_cuda_bench rc=$?"; head -16 gpurun_out/r2c_ref_cuda_bench.md; tail -3 gpurun_out/r2c_ref_cuda_bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-reference-cuda 2> gpurun_out/r2c_bench.err | tail -1 > gpurun_out/r2c_bench.json
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2c_bench.json').read())
    print('value %.2f e2e %.2f ms %.3f parity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['rel_l2']))
    print('   roofline', d['roofline']['frac'], d['roofline']['launch_us'], 'dcn', d['roofline_dcn']['frac'], d['roofline_dcn']['launch_us'], 'launches', d['gpu_launches'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2c_bench.err').read()[-2000:])
PY
