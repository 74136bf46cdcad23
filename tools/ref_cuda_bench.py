"""tools/ref_cuda_bench.py -- the kernel to beat: per-call device times of the reference's OWN
modulated-DCN CUDA kernels (oracle/_ref: deform_conv_cuda.cpp:486-679 + deform_conv_cuda_kernel.cu:569-866 compiled for
sm_100a, im2col + cuBLAS SGEMM) next to this library's kernels at the three EDVR sizes 5x64x{44x80, 88x160, 176x320}, plus
cuDNN's 3x3 64->64 convolution next to conv_tc2, plus whole-model numbers (reference EDVR forward / adapted frame in stock
eager PyTorch).  Writes a markdown table to stdout (gpurun_out/ref_cuda_bench.md -> profiles/r2_reference_cuda.md).

    python tools/ref_cuda_bench.py > gpurun_out/ref_cuda_bench.md
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dynavsr_b200 import ops  # noqa: E402


def ev_time(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ts = []
    for _ in range(reps):
        flush.zero_()                                  # L2 flush between timed iterations (256 MB > 126 MB L2)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    E, L, ext = bench.reference_cuda_modules()
    dc = sys.modules['models.archs.dcn.deform_conv']
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    hbm, tc_peak, _, src = bench.measured_peaks()
    print('# Reference CUDA path on this B200 vs dynavsr_b200 (tools/ref_cuda_bench.py)\n')
    print('GPU: %s, torch %s, peaks %s: HBM %.0f GB/s, bf16 %.0f TF/s.  Median of 20 CUDA-event timings, L2 flushed between '
          'iterations, one launch sequence per timing (the reference forward = memset + im2col + SGEMM + bias kernels).\n' % (
              torch.cuda.get_device_name(0), torch.__version__, src, hbm, tc_peak))
    print('## Modulated deformable conv 3x3, 64 -> 64, dg = 8 (deform_conv_cuda.cpp:486-679)\n')
    print('| size N x H x W | reference fwd us | ours fwd (tcgen05) us | ours fwd (exact fp32) us | speed-up fwd | reference bwd us | ours bwd (tc) us | ours bwd (exact) us | speed-up bwd | ours fwd GB/s (frac of HBM) |')
    print('|---|---|---|---|---|---|---|---|---|---|')
    for (N, H, W) in ((5, 44, 80), (5, 88, 160), (5, 176, 320)):
        g = torch.Generator().manual_seed(1)
        x = torch.randn(N, 64, H, W, generator=g).cuda()
        off = (torch.randn(N, 144, H, W, generator=g) * 1.5).cuda()
        m = torch.rand(N, 72, H, W, generator=g).cuda()
        w = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).cuda()
        b = torch.zeros(64).cuda()
        y = torch.empty(N, 64, H, W, device='cuda')
        gy = torch.randn(N, 64, H, W, device='cuda')
        e = x.new_empty(0)

        def ref_fwd():
            ext.modulated_deform_conv_cuda_forward(x, w, b, e, off, m, y, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 8, True)

        gx, gw, gb, go, gm = (torch.zeros_like(t) for t in (x, w, b, off, m))

        def ref_bwd():
            for t in (gx, gw, gb, go, gm):
                t.zero_()                                   # deform_conv.py:128-132 zeros_like per call
            ext.modulated_deform_conv_cuda_backward(x, w, b, e, off, m, e, gx, gw, gb, go, gm, gy, 3, 3, 1, 1, 1, 1, 1, 1, 1, 8, True)

        xs = x.permute(0, 2, 3, 1).contiguous()
        om = torch.cat([off.permute(0, 2, 3, 1), m.permute(0, 2, 3, 1)], 3).contiguous()
        gys = gy.permute(0, 2, 3, 1).contiguous()
        res = {}
        for tc in (True, False):
            ops.set_conv_backend(tc)
            with torch.no_grad():
                res[('f', tc)] = ev_time(lambda: ops.mdcn(xs, om, w, b, 8, 1, 1, 1))
            xr, omr, wr, br = xs.clone().requires_grad_(True), om.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
            yy = ops.mdcn(xr, omr, wr, br, 8, 1, 1, 1)

            def our_bwd():
                torch.autograd.grad(yy, [xr, omr, wr, br], gys, retain_graph=True)
            res[('b', tc)] = ev_time(our_bwd)
        ops.set_conv_backend(True)
        rf, rb = ev_time(ref_fwd), ev_time(ref_bwd)
        bytes_ = 4.0 * N * H * W * (64 + 216 + 64) + 4 * 64 * 64 * 9
        gbs = bytes_ / (res[('f', True)] * 1e-6) / 1e9
        print('| %dx%dx%d | %.1f | %.1f | %.1f | %.2fx | %.1f | %.1f | %.1f | %.2fx | %.0f (%.3f) |' % (
            N, H, W, rf, res[('f', True)], res[('f', False)], rf / res[('f', True)], rb, res[('b', True)], res[('b', False)],
            rb / res[('b', True)], gbs, gbs / hbm))
    print('\n(ours bwd = autograd.grad through ops.mdcn: act_bwd/bias grad + dvsr_mdcn_bwd_data + the deformable weight gradient, '
          'incl. the Python launch path; the reference bwd includes its five zero-fills)\n')
    print('## 3x3 conv 64 -> 64 (+bias, ReLU): cuDNN (what the reference calls) vs conv_tc2\n')
    print('| size | cuDNN fp32 NCHW us | cuDNN TF32 NCHW us | cuDNN fp32 channels_last us | conv_tc2 BF16x3 us | conv_tc2 bf16 single us | TFLOP/s BF16x3 (frac of bf16 peak) |')
    print('|---|---|---|---|---|---|---|')
    for (N, H, W) in ((5, 44, 80), (5, 88, 160), (5, 176, 320), (1, 704, 1280)):
        x = torch.randn(N, 64, H, W, device='cuda')
        conv = torch.nn.Conv2d(64, 64, 3, 1, 1).cuda()
        with torch.no_grad():
            t32 = ev_time(lambda: torch.relu_(conv(x)))
            torch.backends.cudnn.allow_tf32 = True
            ttf = ev_time(lambda: torch.relu_(conv(x)))
            torch.backends.cudnn.allow_tf32 = False
            xcl = x.to(memory_format=torch.channels_last)
            ccl = conv.to(memory_format=torch.channels_last)
            tcl = ev_time(lambda: torch.relu_(ccl(xcl)))
            xs = x.permute(0, 2, 3, 1).contiguous()
            t3 = ev_time(lambda: ops.conv(xs, conv.weight, conv.bias, act=ops.ACT_RELU))
            with ops.conv_precision('bf16'):
                t1 = ev_time(lambda: ops.conv(xs, conv.weight, conv.bias, act=ops.ACT_RELU))
        fl = 18.0 * N * H * W * 64 * 64
        print('| %dx%dx%d | %.1f | %.1f | %.1f | %.1f | %.1f | %.0f (%.3f) |' % (N, H, W, t32, ttf, tcl, t3, t1, fl / t3 / 1e6, fl / t3 / 1e6 / tc_peak))
    print('\n## Whole model, stock eager reference (unmodified modules + its own DCN kernels) vs this library\n')
    S = bench._synth()
    rnet = S.seed_parameters(E.EDVR(nf=64, nframes=5, groups=8, front_RBs=5, back_RBs=10), 1234).cuda().eval()
    clip = S.synth_clip(21, 176, 320).cuda()
    from dynavsr_b200.models.archs import EDVR_arch
    net = EDVR_arch.EDVR(nf=64, nframes=5, groups=8, front_RBs=5, back_RBs=10, scale=4)
    net.load_state_dict(rnet.state_dict(), strict=True)
    net = net.cuda()
    with torch.no_grad():
        tr = ev_time(lambda: rnet(clip), reps=10)
        to = ev_time(lambda: net(clip), reps=10)
        want, got = rnet(clip), net(clip)
    err = float((got.double() - want.double()).norm() / want.double().norm())
    print('| workload | reference-CUDA ms | ours ms | speed-up | rel. L2 difference of the outputs |')
    print('|---|---|---|---|---|')
    print('| EDVR-M 4x forward 5x3x176x320 -> 3x704x1280, one frame at a time, eager | %.2f | %.2f | %.2fx | %.2e |' % (tr / 1e3, to / 1e3, tr / to, err))


if __name__ == '__main__':
    main()
