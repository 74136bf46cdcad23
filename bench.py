#!/usr/bin/env python
"""bench.py -- HR frames/s at 4x including the 2-step inner adaptation (BASELINE.json metric).

One "step" = one output HR frame of the DynaVSR test-time path (codes/test_dynavsr.py:197-283) on a
synthetic REDS4-shaped window: restore meta-weights, 2 x {MFDN(LR)->SLR, EDVR(SLR), L2 + 10*L1(SLR),
backward, fused SGD on EDVR+MFDN}, final EDVR(LR) -> HR.  The 5x3x180x320 LR window is cropped to
5x3x176x320 exactly as the reference loader does (video_test_dataset_int.py:185-189: SLR must be a
multiple of 4), so HR is 3x704x1280.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload adapt|infer]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'HR frames/sec at 4x (5x3x180x320 in) incl. 2-step inner adapt'
UNIT = 'frames/s'
LR_H, LR_W, NFR, SCALE = 176, 320, 5, 4        # 180 -> 176: reference crop rule
INNER = dict(steps=2, lr_alpha=1e-5, optimizer='SGD', criterion='l2', slr_weight=10.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=60)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='adapt', choices=['adapt', 'infer'])
    ap.add_argument('--no-graphs', action='store_true')
    ap.add_argument('--inner-steps', type=int, default=None, help='diagnostic only: override the 2 inner adaptation steps of the metric')
    ap.add_argument('--cta-budget', type=int, default=None, help='CTAs per launch of the persistent kernels (default: AdaptationPool policy)')
    ap.add_argument('--wg-chunks', type=int, default=None, help='wgrad split-K policy: min chunks per CTA')
    ap.add_argument('--min-tiles', type=int, default=None, help='conv_tc2 grid policy: tiles per CTA (default: AdaptationPool policy)')
    ap.add_argument('--pipelines', type=int, default=None, help='independent frames kept in flight per GPU (adapt.AdaptationPool)')
    ap.add_argument('--no-tc', action='store_true', help='exact-fp32 CUDA-core convolutions only')
    ap.add_argument('--inner-precision', default=None, choices=['bf16x3', 'bf16', 'tf32'],
                    help='operand precision of the tensor-core convs during the inner steps (default: same as the final forward, bf16x3)')
    ap.add_argument('--inner-backward-precision', default=None, choices=['bf16x3', 'bf16', 'tf32'],
                    help='... of their backward passes only (default: --inner-precision)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--height', type=int, default=LR_H)
    ap.add_argument('--width', type=int, default=LR_W)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
def synth_clip(seed, H, W, nfr=NFR):
    from dynavsr_b200.synth import synth_clip as f
    return f(seed, H, W, nfr)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe): NVML polled every 20 ms when
    nvidia_ml_py is importable, else one nvidia-smi query per ~0.3 s."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.reasons, self._stop_evt = index, [], [], set(), threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[index]) if vis and vis.split(',')[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n = self.nvml
        self.sm.append(int(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        self.mx.append(int(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(n, 'nvmlDeviceGetCurrentClocksEventReasons') \
            else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        for name, bit in (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4)):
            if r & bit:
                self.reasons.add(name)

    def _poll_smi(self):
        out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
        v = [t.strip() for t in out.split(',')] if out else []
        if len(v) >= 2 and v[0].isdigit():
            self.sm.append(int(v[0]))
            self.mx.append(int(v[1]) if v[1].isdigit() else 0)
            for i, name in enumerate(self.NAMES):
                if len(v) > 2 + i and v[2 + i].lower().startswith('active'):
                    self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._poll_nvml() if self.nvml else self._poll_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.02 if self.nvml else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(self.sm)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(self.mx) if self.mx else None,
                'reasons': [n for n in self.NAMES if n in self.reasons], 'samples': len(sm),
                'source': 'nvml' if self.nvml else 'nvidia-smi'}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1590.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------------------
def cpu_adapt_sample(steps, warmup, budget_s=150.0, full_hw=(LR_H, LR_W)):
    """The oracle port (reference algorithm, plain PyTorch CPU, all host threads) on a bounded sample of the same
    workload.  One warm-up on a 48x80 LR crop estimates the speed; if `steps + warmup` adapted frames at the full
    176x320 window fit in `budget_s` the full window is timed (no extrapolation), otherwise the crop is timed and
    frames/s are scaled by the pixel ratio (stated in `sample`)."""
    import torch
    from oracle import edvr_oracle as O
    from oracle import params as P
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sdG = P.make_params(P.edvr_param_shapes(), seed=1234)
    sdE = P.make_params(P.mfdn_param_shapes(), seed=77)
    sdF = P.make_params(P.mfdn_param_shapes(), seed=78)

    def run(clip, n):
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            O.adapt_and_infer(sdG, sdE, sdF, clip, **INNER)
            ts.append(time.perf_counter() - t0)
        return ts

    ch, cw = 48, 80
    crop = synth_clip(0, ch, cw)
    run(crop, 1)                                  # cold start (thread pools, oneDNN primitives)
    t_crop = min(run(crop, 2))
    ratio = (full_hw[0] * full_hw[1]) / float(ch * cw)
    if t_crop * ratio * (steps + warmup) <= budget_s:
        full = synth_clip(0, full_hw[0], full_hw[1])
        run(full, warmup)
        ts = run(full, steps)
        t = sum(ts) / len(ts)
        value, what = 1.0 / t, 'the full %dx%d LR window: %.2f s per adapted frame (%d timed, %d warm-up)' % (
            full_hw[0], full_hw[1], t, steps, warmup)
    else:
        ts = run(crop, max(steps, 2))
        t = sum(ts) / len(ts)
        value, what = 1.0 / (t * ratio), 'an LR crop %dx%d: %.2f s per adapted frame, scaled by the pixel ratio %.1f to %dx%d' % (
            ch, cw, t, ratio, full_hw[0], full_hw[1])
    return {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'oracle/edvr_oracle.adapt_and_infer (reference algorithm, torch CPU fp32, %d threads) on %s' % (
                torch.get_num_threads(), what), 'seconds_per_sample': t}


def run_reference(args, rank):
    if rank != 0:
        return
    base = cpu_adapt_sample(max(1, args.steps), max(0, args.warmup))
    line = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 / base['value'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'adapt2_sgd_l2+final_forward EDVR-M 4x + MFDN, REDS4-shaped 5x3x180x320 window cropped to %dx%d -> 3x%dx%d' % (
                LR_H, LR_W, SCALE * LR_H, SCALE * LR_W), 'inner': INNER, 'arm': 'reference algorithm (oracle port), host cores, bounded sample'},
            'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.inner_steps is not None:
        INNER['steps'] = args.inner_steps
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    from dynavsr_b200 import _lib, adapt, ops
    from dynavsr_b200.synth import seed_parameters
    from dynavsr_b200.models.archs import EDVR_arch, LRimg_estimator
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local))
    H, W = args.height, args.width
    use_tc = (not args.no_tc) and hasattr(_lib.lib(), 'dvsr_conv_tc_fprop')
    ops.set_conv_backend(use_tc)

    def build(seedG):
        netG = seed_parameters(EDVR_arch.EDVR(nf=64, nframes=NFR, groups=8, front_RBs=5, back_RBs=10, scale=SCALE), seedG)
        netE = seed_parameters(LRimg_estimator.DirectKernelEstimatorVideo(64, 3, SCALE), 77)
        netF = seed_parameters(LRimg_estimator.DirectKernelEstimatorVideo(64, 3, SCALE), 78)
        return netG.cuda(), netE.cuda(), netF.cuda()

    # frames in flight: 6 for the adaptation workload (latency-bound inner steps), 2 for plain inference (GPU-filling kernels)
    P = max(1, args.pipelines if args.pipelines is not None else (6 if args.workload == 'adapt' else 2))
    pool = adapt.AdaptationPool(*build(1234), pipelines=P, cta_budget=args.cta_budget, min_tiles_per_cta=args.min_tiles,
                                use_graphs=not args.no_graphs, inner_precision=(args.inner_precision, args.inner_backward_precision or args.inner_precision), **INNER)
    min_tiles, budget = pool.min_tiles_per_cta, pool.cta_budget
    if args.wg_chunks is not None:
        _lib.lib().dvsr_conv_wgrad_tc_set_min_chunks_per_cta(args.wg_chunks)
    eng = pool.engines[0]
    # distinct windows per step and per rank (clip sharding: frame i -> rank i % world, train_dynavsr.py:509)
    n_clips = 4
    clips_host = [synth_clip(100 + rank * n_clips + i, H, W).pin_memory() for i in range(n_clips)]
    frames_dev = [ops.to_nhwc(c.cuda().reshape(NFR, 3, H, W)) for c in clips_host]
    hr_host = [torch.empty(1, 3, SCALE * H, SCALE * W).pin_memory() for _ in range(P)]
    done = [None] * P
    if args.workload == 'adapt' and not args.no_graphs:
        pool.warm(frames_dev[0])        # capture every pipeline's CUDA graphs outside the timed regions

    def run(e, fr):
        return e.adapt_and_infer_nhwc(fr) if args.workload == 'adapt' else e.infer_nhwc(fr)

    def step_dev(i):
        # frame i goes to pipeline i % P (its own stream, parameter copy and graphs); inputs already resident in HBM
        pool.submit(lambda e: run(e, frames_dev[i % n_clips]))

    def step_e2e(i):
        # the user-facing call with HOST buffers: H2D of the pinned LR window, adapt + super-resolve, D2H of the HR frame,
        # all on the frame's pipeline stream; the host takes delivery of a pipeline's previous frame before reusing its
        # pinned output buffer
        k = i % P
        if done[k] is not None:
            done[k].synchronize()

        def work(e):
            x = clips_host[i % n_clips].cuda(non_blocking=True).reshape(NFR, 3, H, W)
            hr_host[k].copy_(ops.to_nchw(run(e, ops.to_nhwc(x))), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            return ev
        done[k] = pool.submit(work, pipeline=k)[1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        pool.join()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.COUNTER[0] = 0
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        pool.join()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _lib.COUNTER[0]
        if world > 1:
            t = torch.tensor([ms], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, launches

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, host_launches = timed(step_dev, args.steps, max(args.warmup, 3, P))
    clk = clocks.stop() if rank == 0 else None
    ms_e2e, _ = timed(step_e2e, args.steps, max(2, P))
    value = world * args.steps / (ms / 1000.0)
    e2e = world * args.steps / (ms_e2e / 1000.0)
    # kernels launched per step: counted while the step was captured / run eagerly
    per_step = eng.launches_per_step if (getattr(eng, 'launches_per_step', None) and args.workload == 'adapt') \
        else host_launches / max(1, args.steps)

    # ---- roofline of the dominant kernel: the 3x3 64->64 convolution (feature extraction / PCD / trunk shape,
    # N frames x 176x320), timed alone on the launching stream with CUDA events.  20 back-to-back launches are
    # replayed from a CUDA graph so that the number is the kernel's, not the Python launch path's.
    hbm_peak, tc_peak, peak_src = measured_peaks()
    x = torch.randn(NFR, H, W, 64, device='cuda')
    wgt = torch.randn(64, 64, 3, 3, device='cuda') * 0.05
    bia = torch.zeros(64, device='cuda')
    reps = 20
    _lib.lib().dvsr_set_cta_budget(148)                        # the kernel alone: whole GPU, one tile per CTA
    _lib.lib().dvsr_conv_tc2_set_min_tiles_per_cta(1)
    with torch.no_grad():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                ops.conv(x, wgt, bia, act=ops.ACT_RELU)
        torch.cuda.current_stream().wait_stream(side)
        gconv = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gconv):
            for _ in range(reps):
                ops.conv(x, wgt, bia, act=ops.ACT_RELU)
        gconv.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gconv.replay()
        b.record()
        torch.cuda.synchronize()
    t_conv = a.elapsed_time(b) / reps / 1000.0
    _lib.lib().dvsr_set_cta_budget(budget)
    _lib.lib().dvsr_conv_tc2_set_min_tiles_per_cta(min_tiles)
    flops = 18.0 * NFR * H * W * 64 * 64
    prec = ops._backend['precision'] if use_tc else 'fp32'
    roofline = {'kernel': 'conv3x3 64->64 fprop @%dx%dx%d (%s)' % (NFR, H, W, ('tcgen05 ' + prec) if use_tc else 'CUDA-core fp32'),
                'bound': 'tensor', 'achieved': flops / t_conv / 1e12, 'peak': tc_peak, 'unit': 'TFLOP/s',
                'frac': flops / t_conv / 1e12 / tc_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this very launch shape from the committed ncu --set full
                # capture profiles/r1_ncu_tc2.csv (72.3 MB read + 29.8 MB written; the rest of y stays in L2)
                'traffic': 102.1e6 if (use_tc and prec == 'bf16x3' and (H, W) == (LR_H, LR_W)) else None,
                'peak_source': peak_src + ': bf16 dense burst; algorithmic FLOPs = 18*N*H*W*Cin*Cout (the BF16x3 mode issues 3x '
                                          'that many tensor-core MACs and is bounded by the 128 B/clk shared-memory operand path: DESIGN.md section 3)',
                'algorithmic_bytes': 4.0 * NFR * H * W * 128 + 36.0 * 64 * 64, 'launch_us': t_conv * 1e6,
                'tensor_macs_issued_x': 3 if prec == 'bf16x3' else 1}

    # ---- parity of the timed configuration at full size: tcgen05 path vs this library's exact-fp32 CUDA-core path
    parity = None
    if rank == 0 and use_tc:
        fr = frames_dev[0]
        ops.set_conv_backend(False)
        eng_ref = adapt.InnerLoopAdapter(*build(1234), use_graphs=False, **INNER)
        ref = (eng_ref.adapt_and_infer_nhwc(fr) if args.workload == 'adapt' else eng_ref.infer_nhwc(fr)).detach()
        ops.set_conv_backend(True)
        out = (eng.adapt_and_infer_nhwc(fr) if args.workload == 'adapt' else eng.infer_nhwc(fr)).detach()
        err = float((out.double() - ref.double()).norm() / ref.double().norm())
        q = lambda t: (t.clamp(0, 1) * 255.0).round()
        mse = float(((q(out) - q(ref)) ** 2).mean())
        parity = {'vs': 'exact-fp32 CUDA-core path of this library, same input/weights, full size', 'rel_l2': err,
                  'psnr_db_between_uint8_outputs': None if mse == 0 else 20 * __import__('math').log10(255.0 / mse ** 0.5),
                  'tolerance': 1e-3}
        del eng_ref

    if rank == 0:
        cpu = None if args.no_cpu_baseline else cpu_adapt_sample(2, 1, budget_s=40.0)
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': max(args.warmup, 3, P), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': ('bf16x3 split operands, fp32 accumulate (tcgen05)' if ops._backend['precision'] == 'bf16x3' else 'tf32, fp32 accumulate (tcgen05)') if use_tc else 'f32',
                'data': 'synthetic',
                'config': {'workload': ('adapt2_sgd_l2+final_forward' if args.workload == 'adapt' else 'inference_only') +
                           ' EDVR-M 4x + MFDN, REDS4-shaped 5x3x180x320 window cropped to %dx%d -> 3x%dx%d' % (H, W, SCALE * H, SCALE * W),
                           'inner': INNER, 'inner_conv_precision': {'forward': args.inner_precision or 'bf16x3',
                                                    'backward': args.inner_backward_precision or args.inner_precision or 'bf16x3'}, 'clips_per_rank': n_clips,
                           'cuda_graphs': not args.no_graphs,
                           'frames_in_flight_per_gpu': P, 'conv_min_tiles_per_cta': min_tiles, 'cta_budget_per_launch': budget,
                           'l2': 'per-step working set (activations ~GBs) exceeds the 126 MB L2; inputs rotate over %d clips' % n_clips,
                           'parallelism': 'clip-sharded dp%d, no data-path collective' % world},
                'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': NFR * 3 * H * W * 4,
                        'd2h_bytes_per_step': 3 * SCALE * H * SCALE * W * 4, 'ms_per_step': ms_e2e / args.steps},
                'gpu_launches': int(per_step * args.steps), 'clocks': clk, 'roofline': roofline, 'cpu_baseline': cpu,
                'parity': parity}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
