#!/bin/bash
mkdir -p gpurun_out
for st in 0 1 2; do
  timeout 600 python bench.py --steps 48 --warmup 12 --inner-steps $st --no-cpu-baseline > gpurun_out/bench_s.log 2>&1
  tail -1 gpurun_out/bench_s.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('inner_steps=$st value %.2f ms %.2f' % (d['value'], d['ms_per_step']))
except Exception as e: print('failed', e)
"
done
timeout 600 python bench.py --steps 48 --warmup 12 --workload infer --pipelines 6 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('infer P=6 value %.2f ms %.2f' % (d['value'], d['ms_per_step']))"
for shape in "5 176 320 64 216 3" "5 176 320 64 64 3" "5 176 320 64 256 3" "1 704 1280 64 64 3" "5 44 80 64 64 3" "1 44 80 64 64 3"; do echo $shape; python tools/one_conv.py $shape 2>&1 | grep avg; done
