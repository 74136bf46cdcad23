#!/bin/bash
# meta.MetaPool: task lanes (config 4 with several tasks per rank per outer step)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_meta_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log | cut -c1-250
for cfg in "1 1" "4 1" "4 2" "4 4" "8 4" "8 8"; do set -- $cfg
  timeout 400 python bench.py --workload meta --steps 6 --warmup 3 --tasks-per-rank $1 --task-lanes $2 2>gpurun_out/r2q_meta_$1_$2.err | tail -1 > gpurun_out/r2q_meta_$1_$2.json
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2q_meta_$1_$2.json').read()); print('tasks/rank $1 lanes $2: %.2f tasks/s, %.2f ms/outer step, e2e %.2f, loss %.5f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['loss_q']))
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2q_meta_$1_$2.err').read()[-1500:])"
done
