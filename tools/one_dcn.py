"""tools/one_dcn.py [N H W] [--bwd] [--offset-std S] -- launch the modulated DCN forward (and backward) a few times at the EDVR L1
shape (target for `ncu -k regex:mdcn`).  Offsets ~ N(0, S^2) pixels (default 1.0: what the bench's roofline_dcn uses)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynavsr_b200 import ops  # noqa: E402

nums = [int(v) for v in sys.argv[1:] if v.isdigit()]
N, H, W = nums[:3] if len(nums) >= 3 else (5, 176, 320)
std = float(sys.argv[sys.argv.index('--offset-std') + 1]) if '--offset-std' in sys.argv else 1.0
ops.set_conv_backend(True)
g = torch.Generator().manual_seed(0)
x = torch.randn(N, H, W, 64, generator=g).cuda()
om = torch.cat([torch.randn(N, H, W, 144, generator=g) * std, torch.rand(N, H, W, 72, generator=g)], 3).cuda().contiguous()
w = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).cuda()
b = torch.zeros(64).cuda()


from bench import graph_time  # noqa: E402  (CUDA-graph replay of 20 launches, CUDA events)


def timed(fn, reps=20):
    return graph_time(fn, reps) * 1e6


def eager(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


STAGED = int(sys.argv[sys.argv.index('--staged') + 1]) if '--staged' in sys.argv else 0      # dvsr_policy.mdcn_staged: 0 auto, 1 direct, 2 staged
BUDGET = int(sys.argv[sys.argv.index('--cta-budget') + 1]) if '--cta-budget' in sys.argv else 0
with torch.no_grad(), ops.scope(ops.new_scope(ops.LaunchPolicy(cta_budget=BUDGET, mdcn_staged=STAGED))):
    print('fwd avg us', timed(lambda: ops.mdcn(x, om, w, b, 8, 1, 1, 1, ops.ACT_LRELU)))
if '--bwd' in sys.argv:
    xr, omr, wr = x.clone().requires_grad_(True), om.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y = ops.mdcn(xr, omr, wr, b, 8, 1, 1, 1)
    gy = torch.randn_like(y)
    print('bwd avg us (eager launches)', eager(lambda: torch.autograd.grad(y, [xr, omr, wr], gy, retain_graph=True)))
