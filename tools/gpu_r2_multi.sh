#!/bin/bash
# Multi-GPU evidence (usage: gpurun --gpus N --timeout 900 -- 'bash tools/gpu_r2_multi.sh N'): the metric workload (frame-sharded, no
# collective) and BASELINE config 4 (meta-training: one fused exchange + update per outer step over NVLink peer memory, and the NCCL
# fallback) at N ranks; at N = 2 also the 2-rank-vs-1-process numerical check of the exchange.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
[ -z "$SKIP_ADAPT" ] && timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 36 --warmup 6 --no-cpu-baseline --no-parity > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -1 gpurun_out/r2_bench_n$N.json | cut -c1-300; tail -2 gpurun_out/r2_bench_n$N.err
for ex in peer peer-all nccl; do
  timeout 600 $TR --master-port 29512 bench.py --gpus $N --workload meta --steps 10 --warmup 3 --exchange $ex > gpurun_out/r2_meta_n${N}_$ex.json 2> gpurun_out/r2_meta_n${N}_$ex.err
  tail -1 gpurun_out/r2_meta_n${N}_$ex.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('meta N=$N $ex: %.2f tasks/s, %.2f ms/outer step, e2e %.2f, exchange %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['exchange']))
except Exception as e:
    print('meta N=$N $ex failed', e)"
  tail -2 gpurun_out/r2_meta_n${N}_$ex.err | cut -c1-300
done
if [ "$N" = "2" ]; then
  timeout 600 $TR --master-port 29513 tools/meta_dist_check.py > gpurun_out/r2_meta_check_n2.log 2>&1; grep -v 'Warning\|warn' gpurun_out/r2_meta_check_n2.log | tail -8
fi
