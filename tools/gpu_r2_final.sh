#!/bin/bash
# End-of-round evidence run: every GPU test, smoke(), the bench lines (default = what the driver runs, infer, single-frame latency),
# the kernel-to-beat table.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2z_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8 > gpurun_out/r2z_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2z_pytest_gpu.log; tail -14 gpurun_out/r2z_pytest_gpu.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2z_bench.json').read().strip().split('\n')[-1])
    print('value %.2f e2e %.2f ms %.3f parity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['rel_l2']))
    print('   roofline', d['roofline']['frac'], d['roofline']['launch_us'], 'inner', d['roofline_inner']['launch_us'], 'dcn', d['roofline_dcn']['frac'], d['roofline_dcn']['launch_us'], 'step', d['roofline_step']['frac'], 'launches', d['gpu_launches'])
    print('   cpu', d['cpu_baseline']['value'], 'refcuda', d['reference_cuda'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2z_bench.err').read()[-2000:])
PY
timeout 300 python bench.py --workload infer --no-cpu-baseline --no-reference-cuda 2> gpurun_out/r2z_bench_infer.err | tail -1 > gpurun_out/r2z_bench_infer.json; python -c "
import json; d=json.loads(open('gpurun_out/r2z_bench_infer.json').read()); print('infer value %.2f e2e %.2f' % (d['value'], d['e2e']['value']))"
timeout 300 python bench.py --pipelines 1 --no-cpu-baseline --no-reference-cuda --no-parity 2> gpurun_out/r2z_bench_p1.err | tail -1 > gpurun_out/r2z_bench_p1.json; python -c "
import json; d=json.loads(open('gpurun_out/r2z_bench_p1.json').read()); print('p1 value %.2f e2e %.2f' % (d['value'], d['e2e']['value']))"
timeout 300 python bench.py --workload meta --steps 10 --warmup 3 2> gpurun_out/r2z_meta.err | tail -1 > gpurun_out/r2z_meta.json; python -c "
import json; d=json.loads(open('gpurun_out/r2z_meta.json').read()); print('meta value %.2f ms %.2f launches %s' % (d['value'], d['ms_per_step'], d['gpu_launches']))"
timeout 600 python tools/ref_cuda_bench.py > gpurun_out/r2z_ref_cuda_bench.md 2> gpurun_out/r2z_ref_cuda_bench.err; echo "ref_cuda_bench rc=$?"; cat gpurun_out/r2z_ref_cuda_bench.md | head -30
