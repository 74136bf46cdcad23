#!/bin/bash
# Launch-policy sweep of the adaptation pool after the round-2 kernel changes (frames in flight x tiles per conv CTA x CTA budget).
run() {
  timeout 300 python bench.py --steps 36 --warmup 6 --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$*: value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
}
run
run --pipelines 8
run --pipelines 10
run --min-tiles 4
run --min-tiles 4 --pipelines 8
run --min-tiles 3 --pipelines 8 --cta-budget 30
run --cta-budget 50 --pipelines 6
run --cta-budget 30 --pipelines 8
run --wg-chunks 48 --pipelines 8
