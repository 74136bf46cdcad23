"""tools/kernel_bench.py -- per-kernel timings (CUDA events) at EDVR-M shapes; prints one JSON line per case.
Used to find which kernel to optimise next and for the roofline sweep (BASELINE config 5)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynavsr_b200 import ops  # noqa: E402


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def main():
    tc = '--tc' in sys.argv
    ops.set_conv_backend(tc)
    out = []
    conv_shapes = [] if '--only-mdcn' in sys.argv else [(1, 176, 320), (5, 176, 320), (5, 88, 160), (5, 44, 80), (1, 704, 1280), (5, 44, 80)]
    for (N, H, W) in conv_shapes:
        for (Ci, Co) in [(64, 64), (128, 64), (64, 216), (64, 256)]:
            if H >= 704 and Co != 64:
                continue
            x = torch.randn(N, H, W, Ci, device='cuda')
            w = torch.randn(Co, Ci, 3, 3, device='cuda') * 0.05
            b = torch.zeros(Co, device='cuda')
            with torch.no_grad():
                t = timeit(lambda: ops.conv(x, w, b, act=ops.ACT_LRELU))
            fl = 18.0 * N * H * W * Ci * Co
            by = 4.0 * N * H * W * (Ci + Co) + 36.0 * Ci * Co
            out.append(dict(k='conv3x3', tc=tc, N=N, H=H, W=W, Ci=Ci, Co=Co, us=t * 1e6, tflops=fl / t / 1e12, gbs=by / t / 1e9))
            print(json.dumps(out[-1]), flush=True)
    # DCN forward / backward at the three pyramid levels (5 frames batched)
    for (N, H, W) in [(5, 176, 320), (5, 88, 160), (5, 44, 80), (5, 44, 80 // 1)]:
        x = torch.randn(N, H, W, 64, device='cuda', requires_grad=True)
        om = torch.cat([torch.randn(N, H, W, 144, device='cuda') * 2, torch.rand(N, H, W, 72, device='cuda')], 3).requires_grad_(True)
        w = (torch.randn(64, 64, 3, 3, device='cuda') * 0.05).requires_grad_(True)
        b = torch.zeros(64, device='cuda', requires_grad=True)
        with torch.no_grad():
            t = timeit(lambda: ops.mdcn(x, om, w, b, 8))
        by = 4.0 * N * H * W * (64 + 216 + 64) + 4 * 64 * 64 * 9
        fl = 2.0 * N * H * W * 64 * 64 * 9
        print(json.dumps(dict(k='mdcn_fwd', N=N, H=H, W=W, us=t * 1e6, gbs=by / t / 1e9, tflops=fl / t / 1e12)), flush=True)
        y = ops.mdcn(x, om, w, b, 8)
        gy = torch.randn_like(y)
        t = timeit(lambda: torch.autograd.grad(y, [x, om, w, b], gy, retain_graph=True), reps=5, warm=1)
        print(json.dumps(dict(k='mdcn_bwd_all', N=N, H=H, W=W, us=t * 1e6)), flush=True)
    # conv backward (dgrad + wgrad + act) at the inner-loop resolution
    for (N, H, W) in ([] if '--only-mdcn' in sys.argv else [(5, 44, 80), (1, 176, 320), (5, 176, 320)]):
        x = torch.randn(N, H, W, 64, device='cuda', requires_grad=True)
        w = (torch.randn(64, 64, 3, 3, device='cuda') * 0.05).requires_grad_(True)
        b = torch.zeros(64, device='cuda', requires_grad=True)
        y = ops.conv(x, w, b, act=ops.ACT_LRELU)
        gy = torch.randn_like(y)
        t = timeit(lambda: torch.autograd.grad(y, [x, w, b], gy, retain_graph=True), reps=10, warm=2)
        print(json.dumps(dict(k='conv3x3_bwd_all', N=N, H=H, W=W, us=t * 1e6, tflops=2 * 18.0 * N * H * W * 64 * 64 / t / 1e12)), flush=True)


if __name__ == '__main__':
    main()
