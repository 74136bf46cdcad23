"""tools/small_kernels.py -- latency of single C-ABI kernels at the inner-loop (SLR) resolution, each timed as 20 back-to-back
launches replayed from a CUDA graph (CUDA events; launch overhead excluded)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynavsr_b200 import _lib, ops  # noqa: E402
from dynavsr_b200._lib import ConvDesc, call  # noqa: E402

P, S = ops._ptr, ops._stream
ops.set_conv_backend(True)


def graph_time(fn, reps=20):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def wgrad_case(N, H, W, C, Co, k=3, stride=1, pad=1):
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    x = torch.randn(N, H, W, C, device='cuda')
    gy = torch.randn(N, Ho, Wo, Co, device='cuda')
    w = torch.randn(Co, C, k, k, device='cuda')
    gw = torch.zeros_like(w)
    wl = ops._layout(w, [C], False)
    d = ConvDesc()
    d.N, d.H, d.W, d.Ho, d.Wo = N, H, W, Ho, Wo
    d.KH, d.KW, d.stride, d.pad, d.dil = k, k, stride, pad, 1
    d.nseg = 1
    ops._fill_seg(d.seg[0], x)
    d.Co = Co
    return lambda: call('dvsr_conv_wgrad_tc', ctypes.byref(d), 0, P(gy), Co, P(gw), ctypes.byref(wl), S()), (x, gy, w, gw, wl, d)


def main():
    for shape in [(1, 44, 80, 64, 64), (5, 44, 80, 64, 64), (5, 22, 40, 64, 64), (5, 176, 320, 64, 64)]:
        for env in [{}]:
            for k in ('DVSR_WG_DEBUG', 'DVSR_WG_PER'):
                os.environ.pop(k, None)
            os.environ.update(env)
            fn, keep = wgrad_case(*shape)
            print('wgrad_tc %-22s %-45s %7.1f us' % (shape, env, graph_time(fn)), flush=True)
    for k in ('DVSR_WG_DEBUG', 'DVSR_WG_PER'):
        os.environ.pop(k, None)
    for (N, H, W) in [(1, 44, 80), (5, 44, 80), (5, 22, 40), (5, 11, 20), (1, 176, 320)]:
        x = torch.randn(N, H, W, 64, device='cuda')
        w = torch.randn(64, 64, 3, 3, device='cuda') * 0.05
        b = torch.zeros(64, device='cuda')
        with torch.no_grad():
            t = graph_time(lambda: ops.conv(x, w, b, act=ops.ACT_LRELU))
        print('conv_tc2 fprop %dx%dx%d 64->64: %7.1f us' % (N, H, W, t), flush=True)
        gy = torch.randn(N, H, W, 64, device='cuda'); y = torch.randn(N, H, W, 64, device='cuda'); gp = torch.empty_like(gy); gb = torch.zeros(64, device='cuda')
        t = graph_time(lambda: call('dvsr_act_bwd', P(gy), P(y), None, P(gp), P(gb), N * H * W, 64, 2, 0.1, 0, 0, H, W, S()))
        t2 = graph_time(lambda: call('dvsr_act_bwd', P(gy), P(y), None, P(gp), None, N * H * W, 64, 2, 0.1, 0, 0, H, W, S()))
        print('act_bwd %dx%dx%d C64: with bias grad %7.1f us, without %7.1f us' % (N, H, W, t, t2), flush=True)


if __name__ == '__main__':
    main()
