"""tools/one_wgrad_small.py -- weight gradient of conv_last (64 -> 3, 3x3) at 1x176x320, graph-timed (the small-Co kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynavsr_b200 import ops  # noqa: E402

ops.set_conv_backend(True)
x = torch.randn(1, 176, 320, 64, device='cuda')
w = (torch.randn(3, 64, 3, 3, device='cuda') * 0.05).requires_grad_(True)
b = torch.zeros(3, device='cuda')
y = ops.conv(x, w, b, pad=1)
gy = torch.randn_like(y)
(gw,) = torch.autograd.grad(y, [w], gy, retain_graph=True)
ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2).double(), w.shape, gy.permute(0, 3, 1, 2).double(), padding=1)
print('rel err', float((gw.double() - ref).norm() / ref.norm()))
from dynavsr_b200 import _lib  # noqa: E402
for _ in range(3):
    torch.autograd.grad(y, [w], gy, retain_graph=True)
_lib.PROFILE['on'] = True
for _ in range(20):
    torch.autograd.grad(y, [w], gy, retain_graph=True)
print(_lib.profile_report())
