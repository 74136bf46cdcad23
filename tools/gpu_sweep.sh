#!/bin/bash
mkdir -p gpurun_out
for cfg in "6 2 37" "6 2 24" "8 2 37" "8 2 24" "8 1 37" "12 2 24"; do
  set -- $cfg
  timeout 600 python bench.py --steps 48 --warmup 12 --pipelines $1 --min-tiles $2 --cta-budget $3 --no-cpu-baseline > gpurun_out/bench_s.log 2>&1
  tail -1 gpurun_out/bench_s.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('P=$1 min_tiles=$2 budget=$3 value %.2f e2e %.2f ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
except Exception as e: print('P=$1 failed', e)
"
done
nvidia-smi --query-gpu=memory.used --format=csv
