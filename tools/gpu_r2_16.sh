#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2o_pytest.log | cut -c1-200
for i in 1 2; do
timeout 300 python bench.py --steps 36 --warmup 6 --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:act_bwd -c 60 python tools/one_frame.py 1 2>&1 | grep -E "gpu__time" | awk '{s+=$3; n++} END{print "act_bwd avg us under ncu:", s/n, "n", n}'
