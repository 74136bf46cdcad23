#!/bin/bash
# Round-2 second GPU call: full GPU test suite after the SFDN / use_patch / use_real / meta-graph work, and the meta workload at N=1.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2b_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest_gpu.log; tail -40 gpurun_out/r2b_pytest_gpu.log
for g in "" "--no-graphs"; do
  timeout 600 python bench.py --workload meta --steps 10 --warmup 3 $g 2> gpurun_out/r2b_meta$g.err | tail -1 > "gpurun_out/r2b_meta$g.json"
  echo "meta $g rc=$?"; head -c 1800 "gpurun_out/r2b_meta$g.json"; echo; tail -5 "gpurun_out/r2b_meta$g.err"
done
timeout 600 python bench.py --workload meta --steps 10 --warmup 3 --meta-precision bf16x3 2> gpurun_out/r2b_meta_x3.err | tail -1 > gpurun_out/r2b_meta_x3.json; head -c 600 gpurun_out/r2b_meta_x3.json; echo
