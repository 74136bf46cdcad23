#!/bin/bash
# what the driver runs at round end: GPU tests, smoke(), bench (ours + reference arm)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); print('bench value %.2f e2e %.2f cpu %.3f launches %d clocks %s' % (d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks']))"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
