"""tools/e2e_probe.py -- why is the end-to-end (host buffers) number below the resident-input number?  Builds the same AdaptationPool
as bench.py and times, for several variants of the user-facing call, the host's submission cost per frame (wall clock) and the
throughput (CUDA events around K frames)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import INNER, LR_H as H, LR_W as W, NFR, SCALE, synth_clip  # noqa: E402
from dynavsr_b200 import adapt, ops  # noqa: E402
from dynavsr_b200.models.archs import EDVR_arch, LRimg_estimator  # noqa: E402
from dynavsr_b200.synth import seed_parameters  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 6
K = 36
ops.set_conv_backend(True)
netG = seed_parameters(EDVR_arch.EDVR(nf=64, nframes=NFR, groups=8, front_RBs=5, back_RBs=10, scale=SCALE), 1234).cuda()
netE = seed_parameters(LRimg_estimator.DirectKernelEstimatorVideo(64, 3, SCALE), 77).cuda()
netF = seed_parameters(LRimg_estimator.DirectKernelEstimatorVideo(64, 3, SCALE), 78).cuda()
pool = adapt.AdaptationPool(netG, netE, netF, pipelines=P, use_graphs=True, inner_precision=('bf16', 'bf16'), **INNER)
clips_host = [synth_clip(100 + i, H, W).pin_memory() for i in range(4)]
frames_dev = [ops.to_nhwc(c.cuda().reshape(NFR, 3, H, W)) for c in clips_host]
pool.warm(frames_dev[0])


def run_variant(name, ring, h2d, d2h, layout, sync):
    R = max(1, ring) * P
    hr_host = [torch.empty(1, 3, SCALE * H, SCALE * W).pin_memory() for _ in range(R)]
    hr_host_nhwc = [torch.empty(1, SCALE * H, SCALE * W, 3).pin_memory() for _ in range(R)]
    done = [None] * R
    host_t = [0.0]

    def step(i):
        k = i % R
        if sync and done[k] is not None:
            done[k].synchronize()
        t0 = time.perf_counter()

        def work(e):
            if h2d:
                x = ops.to_nhwc(clips_host[i % 4].cuda(non_blocking=True).reshape(NFR, 3, H, W))
            else:
                x = frames_dev[i % 4]
            out = e.adapt_and_infer_nhwc(x)
            if d2h:
                if layout:
                    hr_host[k].copy_(ops.to_nchw(out), non_blocking=True)
                else:
                    hr_host_nhwc[k].copy_(out, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            return ev
        done[k] = pool.submit(work, pipeline=i % P)[1]
        host_t[0] += time.perf_counter() - t0

    for i in range(2 * P):
        step(i)
    pool.join()
    torch.cuda.synchronize()
    host_t[0] = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for i in range(K):
        step(2 * P + i)
    pool.join()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0
    ms = e0.elapsed_time(e1)
    print('%-58s %6.2f frames/s  (%.2f ms/frame device, %.2f ms/frame wall, host submission %.2f ms/frame)'
          % (name, K / ms * 1e3, ms / K, wall / K * 1e3, host_t[0] / K * 1e3), flush=True)


run_variant('resident inputs, no copies, no host sync', 1, False, False, False, False)
run_variant('resident inputs, no copies, host sync ring 1', 1, False, False, False, True)
run_variant('resident inputs, no copies, host sync ring 2', 2, False, False, False, True)
run_variant('H2D only, ring 2', 2, True, False, False, True)
run_variant('D2H only (NHWC, no layout pass), ring 2', 2, False, True, False, True)
run_variant('D2H only (to_nchw + copy), ring 2', 2, False, True, True, True)
run_variant('full e2e, ring 1 (round-1 bench)', 1, True, True, True, True)
run_variant('full e2e, ring 2', 2, True, True, True, True)
run_variant('full e2e, ring 3', 3, True, True, True, True)
run_variant('full e2e NHWC out, ring 2', 2, True, True, False, True)
