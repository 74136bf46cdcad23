#!/bin/bash
# First GPU call of a new round: the checks that were written after the previous round's GPU minutes ran out
# ("First on a GPU next round").  BEFORE the call, in the build container:
#     git apply tools/patches/conv_tc2_single_product.diff && make -C dynavsr_b200/csrc
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_first_checks.sh'
mkdir -p gpurun_out
DVSR_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_models_gpu.py -x -q -m gpu -k wrapper_golden -p no:cacheprovider \
    > gpurun_out/first_wrapper_golden.log 2>&1; echo "wrapper golden rc=$?"; tail -3 gpurun_out/first_wrapper_golden.log
DVSR_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k single_product -p no:cacheprovider \
    > gpurun_out/first_single_product.log 2>&1; echo "single-product conv rc=$?"; tail -3 gpurun_out/first_single_product.log
# device-resident clip + SSIM (the gt_ready event path of driver.evaluate)
timeout 300 python - > gpurun_out/first_resident_ssim.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch.nn.functional as F
from util import gold
import test_driver_gpu as T
from dynavsr_b200 import adapt, clips, driver
from dynavsr_b200.degradation import Degradation
g = gold('driver_sgd2_l2.npz')
netG, netE, netF, base = T._nets(g)
gen = torch.Generator().manual_seed(5)
hr = F.interpolate(torch.rand(1, 3, 10, 14, generator=gen), size=(136, 200), mode='bicubic', align_corners=False).clamp(0, 1)
hr = torch.stack([hr[0, :, t:t + 128, t:t + 192] for t in range(6)])
store = clips.ResidentClips(5, 'new_info', 4, 'cuda').add_degraded('pan', hr, Degradation(21, 4, sigma=[1.6, 1.6]))
eng = adapt.InnerLoopAdapter(netG, netE, netF, steps=2, lr_alpha=1e-4, optimizer='SGD', criterion='l2', slr_weight=10.0, use_graphs=False)
with torch.cuda.stream(torch.cuda.Stream()):
    rows = driver.evaluate(eng, store, baseline_netG=base, compute_ssim=True)
for k, v in rows.items():
    print(k, ['%.4f' % x for x in v])
    assert all(x == x for x in v)
print('resident + ssim ok')
PY
echo "resident ssim rc=$?"; tail -3 gpurun_out/first_resident_ssim.log
# the precision policy end to end: default vs single-product inner steps (value, e2e, parity.rel_l2 must stay <= 1e-3)
for prec in bf16x3 bf16; do
  timeout 300 python bench.py --no-cpu-baseline --inner-precision $prec 2>gpurun_out/first_bench_$prec.err | tail -1 > gpurun_out/first_bench_$prec.json
  python -c "
import json; d=json.loads(open('gpurun_out/first_bench_$prec.json').read()); print('$prec', 'value %.2f e2e %.2f parity %.2e' % (d['value'], d['e2e']['value'], d['parity']['rel_l2']))"
done
# launch list of bench.py itself (the profiles/ list of round 1 came from tools/one_frame.py, the same kernels launched eagerly)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 2000 --csv --log-file gpurun_out/first_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/first_bench_under_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/first_bench_launches.csv > gpurun_out/first_bench_launches_summary.md 2>/dev/null; head -12 gpurun_out/first_bench_launches_summary.md
