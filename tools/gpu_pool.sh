#!/bin/bash
# frames-in-flight pool: parity test + throughput at 1..4 pipelines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_edvr_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "pool or adaptation" > gpurun_out/pytest_pool.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_pool.log; tail -5 gpurun_out/pytest_pool.log
for P in 1 2 3 4; do
  timeout 600 python bench.py --steps 24 --warmup 4 --pipelines $P --no-cpu-baseline > gpurun_out/bench_p$P.log 2>&1
  tail -1 gpurun_out/bench_p$P.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('P=$P value %.2f e2e %.2f ms %.2f clocks %s parity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'], d['parity']['rel_l2']))
except Exception as e: print('P=$P failed', e)
"
done
tail -5 gpurun_out/bench_p3.log | cut -c1-600
