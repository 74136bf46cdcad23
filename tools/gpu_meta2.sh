#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/meta_bench.py --steps 5 > gpurun_out/meta_bench_n2.log 2>&1; tail -1 gpurun_out/meta_bench_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/meta_bench.py --steps 5 --exchange nccl > gpurun_out/meta_bench_n2_nccl.log 2>&1; tail -1 gpurun_out/meta_bench_n2_nccl.log
