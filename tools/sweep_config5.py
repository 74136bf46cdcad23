"""tools/sweep_config5.py -- BASELINE config 5: throughput sweep of the 3x3 64->64 trunk conv (tcgen05, BF16x3) and the DCN forward
(offset gather + tcgen05 GEMM) at HW in {64^2, 128^2, 180x320, 256^2, 360^2} and batch in {1, 5, 32}.  One markdown row per case:
CUDA-event time of 10 launches replayed from a CUDA graph, algorithmic TFLOP/s and GB/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynavsr_b200 import ops  # noqa: E402

ops.set_conv_backend(True)


def gtime(fn, reps=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
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
    return a.elapsed_time(b) / reps * 1e-3


print('| kernel | N | H x W | us | TFLOP/s (algorithmic) | GB/s (algorithmic) |')
print('|---|---:|---|---:|---:|---:|')
w = torch.randn(64, 64, 3, 3, device='cuda') * 0.05
b = torch.zeros(64, device='cuda')
for (H, W) in [(64, 64), (128, 128), (180, 320), (256, 256), (360, 360)]:
    for N in (1, 5, 32):
        Hc, Wc = H // 4 * 4, W // 4 * 4
        x = torch.randn(N, Hc, Wc, 64, device='cuda')
        with torch.no_grad():
            t = gtime(lambda: ops.conv(x, w, b, act=ops.ACT_RELU))
        fl, by = 18.0 * N * Hc * Wc * 64 * 64, 4.0 * N * Hc * Wc * 128 + 36.0 * 64 * 64
        print('| conv3x3 64->64 | %d | %dx%d | %.1f | %.1f | %.0f |' % (N, Hc, Wc, t * 1e6, fl / t / 1e12, by / t / 1e9), flush=True)
        om = torch.cat([torch.randn(N, Hc, Wc, 144, device='cuda'), torch.rand(N, Hc, Wc, 72, device='cuda')], 3)
        with torch.no_grad():
            t = gtime(lambda: ops.mdcn(x, om, w, b, 8, 1, 1, 1, ops.ACT_LRELU))
        by = 4.0 * N * Hc * Wc * (64 + 216 + 64) + 4 * 64 * 64 * 9
        print('| DCN fwd 64->64 dg8 (offsets ~N(0,1) px) | %d | %dx%d | %.1f | %.1f | %.0f |' % (N, Hc, Wc, t * 1e6, fl / t / 1e12, by / t / 1e9), flush=True)
        del x, om
