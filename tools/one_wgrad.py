"""tools/one_wgrad.py [N H W] -- the tensor-core weight gradient of a 3x3 64->64 conv under ncu-style timing (CUDA events around the C call
are host-bound for short kernels: use `ncu --metrics gpu__time_duration.sum -k regex:conv_wgrad_tc`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynavsr_b200 import ops  # noqa: E402

nums = [int(v) for v in sys.argv[1:] if v.isdigit()]
N, H, W = nums[:3] if len(nums) >= 3 else (5, 44, 80)
ops.set_conv_backend(True)
pol = ops.LaunchPolicy(cta_budget=37, min_tiles=2, min_chunks=24) if '--pool-policy' in sys.argv else None
with ops.scope(ops.new_scope(pol)):
    x = torch.randn(N, H, W, 64, device='cuda')
    w = (torch.randn(64, 64, 3, 3, device='cuda') * 0.05).requires_grad_(True)
    b = torch.zeros(64, device='cuda')
    y = ops.conv(x, w, b, pad=1)
    gy = torch.randn_like(y)
    for _ in range(4):
        (gw,) = torch.autograd.grad(y, [w], gy, retain_graph=True)
    torch.cuda.synchronize()
print('done', float(gw.abs().sum()))
