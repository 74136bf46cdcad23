#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/meta_dist_check.py > gpurun_out/meta_n2.log 2>&1; grep -v "Warning\|warn\|^$" gpurun_out/meta_n2.log | tail -16
