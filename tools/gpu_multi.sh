#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 36 --warmup 6 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -1 gpurun_out/bench_n2.json | cut -c1-400; tail -3 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
tail -1 gpurun_out/bench_ref_n2.json | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/meta_dist_check.py > gpurun_out/meta_n2.log 2>&1; grep -v 'Warning\|warn' gpurun_out/meta_n2.log | tail -8
# EDVR-L meta-training outer steps (config 4 diagnostic), both exchange paths
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/meta_bench.py --steps 5 2>/dev/null | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/meta_bench.py --steps 5 --exchange nccl 2>/dev/null | tail -1
