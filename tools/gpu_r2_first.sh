#!/bin/bash
# Round-2 first GPU call: every GPU test (incl. the 8 never-run ones, the full-size oracle parity tests and the reference-CUDA
# GPU-side oracle), the kernel-to-beat table, the bench in both inner precisions, bench.py's own launch list under ncu.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_r2_first.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt 2>&1
python -c "import os; print('cores', os.cpu_count())" >> gpurun_out/r2_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=15 > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_gpu.log; tail -40 gpurun_out/r2_pytest_gpu.log
timeout 600 python tools/ref_cuda_bench.py > gpurun_out/r2_ref_cuda_bench.md 2> gpurun_out/r2_ref_cuda_bench.err; echo "ref_cuda_bench rc=$?"; cat gpurun_out/r2_ref_cuda_bench.md; tail -3 gpurun_out/r2_ref_cuda_bench.err
for prec in bf16 bf16x3; do
  timeout 600 python bench.py --steps 30 --warmup 5 --inner-precision $prec 2> gpurun_out/r2_bench_$prec.err | tail -1 > gpurun_out/r2_bench_$prec.json
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_$prec.json').read())
    print('$prec', 'value %.2f e2e %.2f ms %.3f parity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']))
    print('   roofline', d['roofline']['frac'], d['roofline']['launch_us'], 'inner', d['roofline_inner'], 'dcn', d['roofline_dcn']['frac'], d['roofline_dcn']['launch_us'], 'step', d['roofline_step'])
    print('   cpu', d['cpu_baseline'], 'refcuda', d['reference_cuda'])
except Exception as e:
    print('$prec bench failed', e); print(open('gpurun_out/r2_bench_$prec.err').read()[-2000:])
PY
done
timeout 300 python bench.py --steps 30 --warmup 5 --workload infer --no-cpu-baseline 2> gpurun_out/r2_bench_infer.err | tail -1 > gpurun_out/r2_bench_infer.json; head -c 600 gpurun_out/r2_bench_infer.json; echo
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
# launch list of bench.py itself (graph replay): skip the warm-up/capture launches, list ~3 adapted frames
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 40000 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-parity --no-roofline > gpurun_out/r2_bench_under_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2_bench_launches.csv
python tools/ncu_summary.py gpurun_out/r2_bench_launches.csv > gpurun_out/r2_bench_launches_summary.md 2>/dev/null; head -30 gpurun_out/r2_bench_launches_summary.md
