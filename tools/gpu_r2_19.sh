#!/bin/bash
for r in 8 32 8 32 64; do
DVSR_ACTBWD_ROWS=$r timeout 300 python bench.py --steps 36 --warmup 6 --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('act_bwd rows/thread $r: value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
