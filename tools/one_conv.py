"""tools/one_conv.py [N H W Ci Co k] -- launch one tensor-core conv a few times (target for `ncu -k regex:conv_tc`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynavsr_b200 import ops  # noqa: E402

a = [int(v) for v in sys.argv[1:7]] if len(sys.argv) >= 7 else [5, 176, 320, 64, 64, 3]
N, H, W, Ci, Co, k = a
PREC = sys.argv[sys.argv.index('--precision') + 1] if '--precision' in sys.argv else 'bf16x3'
ops.set_conv_backend(True, PREC)
x = torch.randn(N, H, W, Ci, device='cuda')
w = torch.randn(Co, Ci, k, k, device='cuda') * 0.05
b = torch.zeros(Co, device='cuda')
from bench import graph_time  # noqa: E402  (CUDA-graph replay of 20 launches, CUDA events: the kernel's time, not the Python launch path's)
with torch.no_grad():
    for _ in range(3):
        y = ops.conv(x, w, b, pad=k // 2, act=ops.ACT_LRELU)
print(PREC, 'graph-timed avg us %.2f' % (graph_time(lambda: ops.conv(x, w, b, pad=k // 2, act=ops.ACT_LRELU)) * 1e6))

if '--trace' in sys.argv:
    import ctypes
    from dynavsr_b200 import _lib
    tr = torch.zeros(12 * 64, dtype=torch.int64, device='cuda')
    _lib.lib().dvsr_conv_tc2_set_trace(ctypes.c_void_p(tr.data_ptr()))
    with torch.no_grad():
        ops.conv(x, w, b, pad=k // 2, act=ops.ACT_LRELU)
    torch.cuda.synchronize()
    _lib.lib().dvsr_conv_tc2_set_trace(None)
    t = tr.view(12, 64).cpu()
    t0 = int(t[0, 0])
    names = ['prod:slot free', 'round:tile landed', 'round:done', 'mma:ready seen', 'mma:issued', 'epi:acc full', 'epi:done', 'epi:tmem read']
    for ev in range(8):
        print('%-18s' % names[ev], ' '.join('%6d' % (int(v) - t0) for v in t[ev, :(30 if ev < 5 else 15)]))
if '--both' in sys.argv:
    for prec in ('tf32', 'bf16x3'):
        ops.set_conv_backend(True, prec)
        with torch.no_grad():
            for _ in range(3):
                ops.conv(x, w, b, pad=k // 2, act=ops.ACT_LRELU)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                y = ops.conv(x, w, b, pad=k // 2, act=ops.ACT_LRELU)
            e1.record()
            torch.cuda.synchronize()
            ops.set_conv_backend(False)
            ref = ops.conv(x, w, b, pad=k // 2, act=ops.ACT_LRELU)
        print(prec, 'avg us', e0.elapsed_time(e1) / 20 * 1e3, 'rel vs fp32', float((y.double() - ref.double()).norm() / ref.double().norm()))
