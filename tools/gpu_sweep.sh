#!/bin/bash
mkdir -p gpurun_out
for wg in 4 8 12 20 32; do
  timeout 600 python bench.py --steps 48 --warmup 12 --wg-chunks $wg --no-cpu-baseline > gpurun_out/bench_s.log 2>&1
  tail -1 gpurun_out/bench_s.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('wg_chunks=$wg value %.2f e2e %.2f ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
except Exception as e: print('failed', e)
"
done
