"""tools/tc_check.py -- tcgen05 conv vs the exact-fp32 CUDA-core conv on the same inputs (prints rel errors)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynavsr_b200 import ops  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def run(tag, fn):
    ops.set_conv_backend(False)
    ref = fn()
    torch.cuda.synchronize()
    ops.set_conv_backend(True)
    out = fn()
    torch.cuda.synchronize()
    ops.set_conv_backend(False)
    if isinstance(ref, (tuple, list)):
        print(tag, ['%.2e' % rel(o, r) for o, r in zip(out, ref)], flush=True)
    else:
        print(tag, '%.2e' % rel(out, ref), 'max|ref|=%.3f' % float(ref.abs().max()), flush=True)


def main():
    torch.manual_seed(0)
    dev = 'cuda'
    for (N, H, W, Ci, Co, k) in [(1, 8, 16, 32, 64, 1), (1, 8, 16, 64, 64, 3), (2, 20, 40, 64, 64, 3), (1, 44, 80, 64, 216, 3),
                                 (1, 16, 32, 320, 64, 1), (1, 16, 16, 64, 256, 3), (5, 11, 20, 64, 64, 3)]:
        x = torch.randn(N, H, W, Ci, device=dev)
        w = torch.randn(Co, Ci, k, k, device=dev) * (2.0 / (Ci * k * k)) ** 0.5
        b = torch.randn(Co, device=dev) * 0.1
        with torch.no_grad():
            run('fwd N%d %dx%d %d->%d k%d' % (N, H, W, Ci, Co, k), lambda: ops.conv(x, w, b, pad=k // 2, act=ops.ACT_LRELU))
    # two segments + broadcast + residual
    x1, x2 = torch.randn(5, 16, 32, 64, device=dev), torch.randn(1, 16, 32, 64, device=dev)
    w = torch.randn(64, 128, 3, 3, device=dev) * 0.03
    b = torch.zeros(64, device=dev)
    res = torch.randn(5, 16, 32, 64, device=dev)
    with torch.no_grad():
        run('cat+broadcast+res', lambda: ops.conv([x1, ops.Seg(x2, T=5, Tsrc=1, t_fixed=0)], w, b, res=res))
    # shuffle + sigmoid split
    x = torch.randn(1, 16, 32, 64, device=dev)
    w = torch.randn(256, 64, 3, 3, device=dev) * 0.05
    with torch.no_grad():
        run('shuffle', lambda: ops.conv(x, w, None, act=ops.ACT_LRELU, shuffle=2))
    w = torch.randn(216, 64, 3, 3, device=dev) * 0.05
    b = torch.randn(216, device=dev)
    with torch.no_grad():
        run('sigmoid-split 216', lambda: ops.conv(x, w, b, act=ops.ACT_SIGMOID_SPLIT, sig_split=144))
    # backward through TC data gradients
    def fb():
        xx = x1.clone().requires_grad_(True)
        rr = x2.clone().requires_grad_(True)
        ww = (torch.ones(64, 128, 3, 3, device=dev) * w128).requires_grad_(True)
        y = ops.conv([xx, ops.Seg(rr, T=5, Tsrc=1, t_fixed=0)], ww, None, act=ops.ACT_LRELU)
        return torch.autograd.grad(y, [xx, rr, ww], gy)
    w128 = torch.randn(64, 128, 3, 3, device=dev) * 0.03
    gy = torch.randn(5, 16, 32, 64, device=dev)
    run('bwd (gx, gref, gw)', fb)
    # backward without an activation (no sign flips): data gradients (TC dgrad) and weight gradients (TC wgrad)
    for (N, H, W, Ci, Co, k, pad) in [(5, 16, 32, 64, 64, 3, 1), (2, 20, 24, 128, 64, 3, 1), (1, 16, 16, 64, 216, 3, 1),
                                      (1, 16, 24, 320, 64, 1, 0), (2, 18, 34, 64, 64, 3, 0), (1, 44, 80, 64, 256, 3, 1)]:
        xx0 = torch.randn(N, H, W, Ci, device=dev)
        ww0 = torch.randn(Co, Ci, k, k, device=dev) * 0.05
        Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
        gy0 = torch.randn(N, Ho, Wo, Co, device=dev)

        def fb2():
            xx = xx0.clone().requires_grad_(True)
            ww = ww0.clone().requires_grad_(True)
            bb = torch.zeros(Co, device=dev, requires_grad=True)
            y = ops.conv(xx, ww, bb, pad=pad)
            return torch.autograd.grad(y, [xx, ww, bb], gy0)
        run('bwd-noact N%d %dx%d %d->%d k%d (gx, gw, gb)' % (N, H, W, Ci, Co, k), fb2)
    # stride-2 convolutions: forward through TMA element strides, data gradient through 4 parity classes
    for (N, H, W, Ci, Co, k, pad) in [(2, 16, 24, 64, 64, 3, 1), (1, 18, 34, 64, 128, 4, 0), (2, 10, 18, 128, 64, 4, 0), (5, 44, 80, 64, 64, 3, 1)]:
        xx0 = torch.randn(N, H, W, Ci, device=dev)
        ww0 = torch.randn(Co, Ci, k, k, device=dev) * 0.05
        Ho, Wo = (H + 2 * pad - k) // 2 + 1, (W + 2 * pad - k) // 2 + 1
        gy0 = torch.randn(N, Ho, Wo, Co, device=dev)

        def fb3():
            xx = xx0.clone().requires_grad_(True)
            ww = ww0.clone().requires_grad_(True)
            y = ops.conv(xx, ww, None, stride=2, pad=pad)
            return (y,) + tuple(torch.autograd.grad(y, [xx, ww], gy0))
        run('stride2 N%d %dx%d %d->%d k%d (y, gx, gw)' % (N, H, W, Ci, Co, k), fb3)
    # modulated deformable conv forward on tensor cores
    for (N, H, W) in [(1, 11, 13), (5, 44, 80), (2, 90, 160)]:
        xd = torch.randn(N, H, W, 64, device=dev)
        om = torch.cat([torch.randn(N, H, W, 144, device=dev) * 3.0, torch.rand(N, H, W, 72, device=dev)], 3).contiguous()
        wd = torch.randn(64, 64, 3, 3, device=dev) * 0.05
        bd = torch.randn(64, device=dev) * 0.1
        with torch.no_grad():
            run('mdcn fwd N%d %dx%d' % (N, H, W), lambda: ops.mdcn(xd, om, wd, bd, 8, 1, 1, 1, ops.ACT_LRELU))
    # valid conv (pad 0) as in MFDN
    xp = torch.randn(2, 18, 34, 64, device=dev)
    w = torch.randn(64, 64, 3, 3, device=dev) * 0.05
    with torch.no_grad():
        run('valid pad0', lambda: ops.conv(xp, w, None, pad=0, act=ops.ACT_LRELU))
    print('tc_check done')


if __name__ == '__main__':
    main()      # network-level parity against the reference goldens lives in tests/test_edvr_gpu.py
