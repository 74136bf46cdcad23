#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/bench_q.log 2>&1
tail -1 gpurun_out/bench_q.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('value %.2f e2e %.2f ms %.2f parity %s roof %.1f us' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['rel_l2'], d['roofline']['launch_us']))
"
timeout 600 python bench.py --steps 24 --warmup 4 --workload infer --no-cpu-baseline > gpurun_out/bench_qi.log 2>&1
tail -1 gpurun_out/bench_qi.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('infer value %.2f e2e %.2f ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
"
timeout 300 python tools/profile_step.py > gpurun_out/profile_step.txt 2>&1; grep -c . gpurun_out/profile_step.txt
