#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 36 --warmup 6 --no-cpu-baseline > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
tail -1 gpurun_out/bench_n4.json | cut -c1-300; tail -2 gpurun_out/bench_n4.err | cut -c1-200
