#!/usr/bin/env python
"""bench.py -- HR frames/s at 4x including the 2-step inner adaptation (BASELINE.json metric).

One "step" = one output HR frame of the DynaVSR test-time path (codes/test_dynavsr.py:197-283) on a
synthetic REDS4-shaped window: restore meta-weights, 2 x {MFDN(LR)->SLR, EDVR(SLR), L2 + 10*L1(SLR),
backward, fused SGD on EDVR+MFDN}, final EDVR(LR) -> HR.  The 5x3x180x320 LR window is cropped to
5x3x176x320 exactly as the reference loader does (video_test_dataset_int.py:185-189: SLR must be a
multiple of 4), so HR is 3x704x1280.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-cuda]
                    [--workload adapt|infer|meta]

Arms
  ours            this library (libdvsr_b200.so) through its public API
  reference       the reference ALGORITHM on the box's host cores (oracle port, torch CPU, all threads)
  reference-cuda  the reference's OWN CUDA path on the GPU: its unmodified Python modules (baseline/_ref/codes) driving its
                  own deform_conv_cuda extension compiled for sm_100a (oracle/_ref), stock eager PyTorch everywhere else --
                  the "kernel to beat".  Does not import dynavsr_b200.
Workloads
  adapt  the metric (default);  infer  plain EDVR inference (BASELINE config 2);
  meta   BASELINE config 4: EDVR-L + MFDN meta-training outer steps, clip-sharded over ranks, ONE exchange of the flat
         meta-gradient per outer step fused with the outer Adam update over NVLink peer memory (tasks/s).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import glob
import hashlib
import importlib.util
import json
import math
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
ADAPTED_FRAME_TFLOP = 1.72                     # 2 x 386 GF inner steps + 952 GF final forward at 176x320


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=60)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference', 'reference-cuda'])
    ap.add_argument('--workload', default='adapt', choices=['adapt', 'infer', 'meta'])
    ap.add_argument('--no-graphs', action='store_true')
    ap.add_argument('--inner-steps', type=int, default=None, help='diagnostic only: override the 2 inner adaptation steps of the metric')
    ap.add_argument('--cta-budget', type=int, default=None, help='CTAs per launch of the persistent kernels (default: AdaptationPool policy)')
    ap.add_argument('--wg-chunks', type=int, default=None, help='wgrad split-K policy: min chunks per CTA')
    ap.add_argument('--min-tiles', type=int, default=None, help='conv_tc2 grid policy: tiles per CTA (default: AdaptationPool policy)')
    ap.add_argument('--pipelines', type=int, default=None, help='independent frames kept in flight per GPU (adapt.AdaptationPool)')
    ap.add_argument('--ring', type=int, default=2, help='pinned output buffers per pipeline in the e2e leg')
    ap.add_argument('--no-tc', action='store_true', help='exact-fp32 CUDA-core convolutions only')
    ap.add_argument('--inner-precision', default='bf16', choices=['bf16x3', 'bf16', 'tf32'],
                    help='operand precision of the tensor-core convs during the inner steps (final forward: always bf16x3). '
                         'Default bf16 = single product: admissible for the 2-step SGD setting (parity block decides)')
    ap.add_argument('--inner-backward-precision', default=None, choices=['bf16x3', 'bf16', 'tf32'],
                    help='... of their backward passes only (default: --inner-precision)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-reference-cuda', action='store_true', help='skip the reference-CUDA sub-run that fills `reference_cuda`')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--no-roofline', action='store_true', help='skip the isolated kernel timings (for an ncu launch list of the step itself)')
    ap.add_argument('--height', type=int, default=LR_H)
    ap.add_argument('--width', type=int, default=LR_W)
    # meta workload (BASELINE config 4)
    ap.add_argument('--tasks-per-rank', type=int, default=1)
    ap.add_argument('--task-lanes', type=int, default=1, help='meta workload: tasks of an outer step run on this many lanes side by side (meta.MetaPool)')
    ap.add_argument('--nf', type=int, default=128)
    ap.add_argument('--back-rbs', type=int, default=40)
    ap.add_argument('--exchange', default='peer', choices=['peer', 'peer-all', 'nccl'])
    ap.add_argument('--meta-precision', default='bf16x3', choices=['bf16', 'bf16x3'],
                    help='meta workload: operand precision of the <= 64-channel conv layers (the nf = 128 layers of EDVR-L run TF32 on the streaming kernel)')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
def _synth():
    """dynavsr_b200/synth.py loaded BY FILE PATH (it depends on torch only), so that the reference arms never import the
    dynavsr_b200 package."""
    spec = importlib.util.spec_from_file_location('_dvsr_synth', os.path.join(ROOT, 'dynavsr_b200', 'synth.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def synth_clip(seed, H, W, nfr=NFR):
    return _synth().synth_clip(seed, H, W, nfr)


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
    """(HBM GB/s, bf16 TF/s burst, bf16 TF/s sustained, source)"""
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), d.get('bf16_tflops_sustained', 1400.0), \
            'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1590.0, 1400.0, 'fallback (B200_PROFILING.md)'


def committed_traffic(key):
    """DRAM bytes per launch of kernel `key` from a COMMITTED `ncu --set full` capture: profiles/ncu_traffic.json maps
    key -> {dram_bytes_per_launch, source (a file under profiles/), sha256 of that file}.  The number is reported only when the
    cited file is present and its hash matches; otherwise None (nothing is hard-coded in this script)."""
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    try:
        ent = json.load(open(p)).get(key)
        src = os.path.join(ROOT, 'profiles', ent['source'])
        sha = hashlib.sha256(open(src, 'rb').read()).hexdigest()
        if sha != ent['sha256']:
            return None, None
        return float(ent['dram_bytes_per_launch']), {'file': 'profiles/' + ent['source'], 'sha256': sha[:16]}
    except Exception:
        return None, None


# ---------------------------------------------------------------------------------------------------
def cpu_adapt_sample(steps, warmup, budget_s=150.0, full_hw=(LR_H, LR_W), sds=None, clip=None):
    """The oracle port (reference algorithm, plain PyTorch CPU, all host threads) on a bounded sample of the same
    workload.  One warm-up on a 48x80 LR crop estimates the speed; if `steps + warmup` adapted frames at the full
    176x320 window fit in `budget_s` the full window is timed (no extrapolation), otherwise the crop is timed and
    frames/s are scaled by the pixel ratio (stated in `sample`).  With `sds` (EDVR, MFDN, fixed-MFDN state dicts) and `clip`
    the same weights / window as the GPU arm are used and the last full-size output is returned for the parity block."""
    import torch
    from oracle import edvr_oracle as O
    from oracle import params as P
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if sds is None:
        sds = (P.make_params(P.edvr_param_shapes(), seed=1234), P.make_params(P.mfdn_param_shapes(), seed=77),
               P.make_params(P.mfdn_param_shapes(), seed=78))
    sdG, sdE, sdF = sds
    last = [None]

    def run(c, n):
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            last[0] = O.adapt_and_infer(sdG, sdE, sdF, c, **INNER)
            ts.append(time.perf_counter() - t0)
        return ts

    ch, cw = 48, 80
    crop = synth_clip(0, ch, cw)
    run(crop, 1)                                  # cold start (thread pools, oneDNN primitives)
    t_crop = min(run(crop, 2))
    ratio = (full_hw[0] * full_hw[1]) / float(ch * cw)
    out_full = None
    if t_crop * ratio * (steps + warmup) <= budget_s:
        full = clip if clip is not None else synth_clip(0, full_hw[0], full_hw[1])
        run(full, warmup)
        ts = run(full, steps)
        out_full = last[0]
        t = sum(ts) / len(ts)
        value, what = 1.0 / t, 'the full %dx%d LR window: %.2f s per adapted frame (%d timed, %d warm-up)' % (
            full_hw[0], full_hw[1], t, steps, warmup)
    else:
        ts = run(crop, max(steps, 2))
        t = sum(ts) / len(ts)
        value, what = 1.0 / (t * ratio), 'an LR crop %dx%d: %.2f s per adapted frame, scaled by the pixel ratio %.1f to %dx%d' % (
            ch, cw, t, ratio, full_hw[0], full_hw[1])
    base = {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'oracle/edvr_oracle.adapt_and_infer (reference algorithm, torch CPU fp32, %d threads) on %s' % (
                torch.get_num_threads(), what), 'seconds_per_sample': t}
    return base, out_full


def workload_string(H, W, workload='adapt'):
    return ('adapt2_sgd_l2+final_forward' if workload == 'adapt' else 'inference_only') + \
        ' EDVR-M 4x + MFDN, REDS4-shaped 5x3x180x320 window cropped to %dx%d -> 3x%dx%d' % (H, W, SCALE * H, SCALE * W)


def run_reference(args, rank):
    if rank != 0:
        return
    base, _ = cpu_adapt_sample(max(1, args.steps), max(0, args.warmup))
    line = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 / base['value'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_string(LR_H, LR_W), 'inner': INNER,
                       'arm': 'reference algorithm (oracle port), host cores, bounded sample; ONE host process regardless of --gpus'},
            'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def reference_cuda_modules():
    """The reference's own modules + its own CUDA extension (oracle/build_ref.py built both; they travel with the snapshot).
    Returns (EDVR_arch, LRimg_estimator, ext) or raises RuntimeError with the reason."""
    so = sorted(glob.glob(os.path.join(ROOT, 'oracle', '_ref', 'deform_conv_cuda*.so')))
    codes = os.path.join(ROOT, 'baseline', '_ref', 'codes')
    if not so or not os.path.isdir(codes):
        raise RuntimeError('oracle/_ref/deform_conv_cuda*.so or baseline/_ref/codes missing: run `python oracle/build_ref.py` '
                           'in the build container (needs /root/reference)')
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location('deform_conv_cuda', so[0])
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    if codes not in sys.path:
        sys.path.insert(0, codes)
    sys.modules['models.archs.dcn.deform_conv_cuda'] = ext          # deform_conv.py:10 `from . import deform_conv_cuda`
    import models.archs.EDVR_arch as E
    import models.archs.LRimg_estimator as L
    return E, L, ext


def run_reference_cuda(args, rank):
    """The UNMODIFIED reference EDVR + MFDN adapted frame on the GPU, stock eager path: reference nn.Modules, reference
    deform_conv_cuda kernels (im2col + cuBLAS), cuDNN convs, torch.optim.SGD, deepcopy per frame -- the loop of
    test_dynavsr.py:208-283 restated around the reference's model classes (the driver script itself needs imageio / lmdb / files)."""
    if rank != 0:
        return
    import copy
    import torch
    import torch.nn.functional as F
    try:
        E, L, _ = reference_cuda_modules()
    except Exception as e:
        print(json.dumps({'impl': 'reference-cuda', 'unavailable': str(e).splitlines()[0][:300]}))
        return
    if not torch.cuda.is_available():
        print(json.dumps({'impl': 'reference-cuda', 'unavailable': 'no CUDA device'}))
        return
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = False          # the reference runs fp32 everywhere
    torch.backends.cudnn.allow_tf32 = False
    S = _synth()
    H, W = args.height, args.width
    netG = S.seed_parameters(E.EDVR(nf=64, nframes=NFR, groups=8, front_RBs=5, back_RBs=10), 1234).cuda()
    netE = S.seed_parameters(L.DirectKernelEstimatorVideo(64, 3, SCALE), 77).cuda()
    netF = S.seed_parameters(L.DirectKernelEstimatorVideo(64, 3, SCALE), 78).cuda().eval()
    clips = [synth_clip(100 + i, H, W).pin_memory() for i in range(4)]

    def frame(clip_host, adapt=True):
        clip = clip_host.cuda(non_blocking=True)
        G, Ecp = copy.deepcopy(netG), copy.deepcopy(netE)                      # test_dynavsr.py:208
        if adapt:
            opt = torch.optim.SGD(list(G.parameters()) + list(Ecp.parameters()), lr=INNER['lr_alpha'])   # :213-231
            gt = clip[:, NFR // 2]
            for _ in range(INNER['steps']):
                slr = Ecp(clip.transpose(1, 2)).transpose(1, 2)                # :238-241
                opt.zero_grad()
                loss = F.mse_loss(G(slr), gt)                                  # :262-264 (cri_pix = l2, weight 1)
                with torch.no_grad():
                    slr0 = netF(clip.transpose(1, 2)).transpose(1, 2)          # :267-270 (recomputed every step, as written)
                loss = loss + INNER['slr_weight'] * F.l1_loss(slr, slr0)       # :274
                loss.backward()
                opt.step()                                                     # :276-277
        with torch.no_grad():
            return G(clip).cpu()                                               # :282-283 + the host read of the result

    adapt = args.workload != 'infer'
    for i in range(max(1, args.warmup)):
        frame(clips[i % 4], adapt)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(args.steps):
        frame(clips[i % 4], adapt)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    v = 1000.0 / ms
    print(json.dumps({'impl': 'reference-cuda', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps,
                      'warmup': max(1, args.warmup), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                      'dtype': 'f32', 'data': 'synthetic',
                      'config': {'workload': workload_string(H, W, 'adapt' if adapt else 'infer'), 'inner': INNER,
                                 'arm': 'unmodified reference modules (baseline/_ref/codes) + reference deform_conv_cuda built for sm_100a '
                                        '(oracle/_ref), stock eager PyTorch %s / cuDNN, fp32 (TF32 off), one frame at a time' % torch.__version__},
                      'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': NFR * 3 * H * W * 4,
                              'd2h_bytes_per_step': 3 * SCALE * H * SCALE * W * 4}}))


# ---------------------------------------------------------------------------------------------------
def run_meta(args, rank, world, local):
    """BASELINE config 4: EDVR-L (nf 128, back_RBs 40) + MFDN meta-training outer steps.  Per task LR 5x3x64x64 / HR 3x256x256 /
    SLR 5x3x16x16 (train_dynavsr.py:252-438), inner Adam K = 1, Charbonnier loss; every rank runs its own tasks (DistIterSampler
    sharding), then ONE exchange of the flat meta-gradient (84.3 MB fp32) fused with the outer Adam update."""
    import torch
    import torch.distributed as dist
    from dynavsr_b200 import _lib
    from dynavsr_b200.meta import MetaLearner, MetaPool
    from dynavsr_b200.models.archs import EDVR_arch, LRimg_estimator
    from dynavsr_b200.synth import seed_parameters
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local))
    netG = seed_parameters(EDVR_arch.EDVR(nf=args.nf, nframes=5, groups=8, front_RBs=5, back_RBs=args.back_rbs, scale=4), 1).cuda()
    netE = seed_parameters(LRimg_estimator.DirectKernelEstimatorVideo(64, 3, 4), 2).cuda()
    mkw = dict(inner_steps=1, lr_alpha=1e-5, inner_optimizer='Adam', criterion='cb', outer_optimizer='Adam',
               lr_outer=1e-5, exchange=args.exchange, use_graphs=not args.no_graphs,
               precision=None if args.meta_precision == 'bf16x3' else args.meta_precision)
    lanes = max(1, min(args.task_lanes, args.tasks_per_rank))
    ml = MetaPool(netG, netE, lanes=lanes, **mkw) if lanes > 1 else MetaLearner(netG, netE, **mkw)
    g = torch.Generator().manual_seed(10 + rank)
    T = args.tasks_per_rank
    host = [{'LQs': torch.rand(1, 5, 3, 64, 64, generator=g).pin_memory(), 'GT': torch.rand(1, 3, 256, 256, generator=g).pin_memory(),
             'SuperLQs': torch.rand(1, 5, 3, 16, 16, generator=g).pin_memory()} for _ in range(2 * T)]
    dev = [{k: v.cuda() for k, v in t.items()} for t in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e):
        W_ = max(args.warmup, 3)
        loss = None
        for i in range(W_ + args.steps):
            if i == W_:
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                _lib.COUNTER[0] = 0
                ml.replayed_launches = 0
                ml.exchange_events = []
                e0.record()
            sl = slice((i % 2) * T, (i % 2) * T + T)
            tasks = [{k: v.cuda(non_blocking=True) for k, v in t.items()} for t in host[sl]] if e2e else dev[sl]
            loss = ml.outer_step(tasks)
            if e2e:
                loss = float(loss)                 # the host reads the step's query loss (train_dynavsr.py:440-449 logging)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, _lib.COUNTER[0] + ml.replayed_launches, float(loss)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, launches, lq = timed(False)
    clk = clocks.stop() if rank == 0 else None
    xch = ml.exchange_timing()
    ms_e2e, _, _ = timed(True)
    n_params = ml.theta.numel()
    if rank == 0:
        hbm_peak, tc_peak, tc_sus, peak_src = measured_peaks()
        task_tflop = 3 * 0.3009 + 3 * 0.0188 + 3 * 0.0038          # EDVR-L f+b on the LR patch + inner step on 16x16 SLR + MFDN
        value = world * T * args.steps / (ms / 1e3)
        bytes_in = sum(v.numel() * 4 for v in host[0].values()) * T
        line = {'metric': 'meta-training tasks/s (BASELINE config 4: EDVR-L + MFDN outer steps)', 'value': value, 'unit': 'tasks/s',
                'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': ml.dtype_string(), 'data': 'synthetic',
                'config': {'workload': 'meta-train EDVR(nf=%d, back_RBs=%d) + MFDN: task = LR 5x3x64x64 -> HR 3x256x256, SLR 5x3x16x16, inner Adam K=1, '
                                       'cb loss, outer Adam' % (args.nf, args.back_rbs), 'tasks_per_rank_per_outer_step': T, 'task_lanes': lanes,
                           'flat_params': n_params, 'flat_gradient_MB': n_params * 4 / 1e6, 'exchange': ml.exchange,
                           'cuda_graphs': ml.use_graphs, 'parallelism': 'clip-sharded dp%d, ONE fused exchange+update per outer step' % world,
                           'l2': 'EDVR-L weights + packs + activations of a task (~1 GB) exceed the 126 MB L2; tasks alternate between two sets'},
                'e2e': {'value': world * T * args.steps / (ms_e2e / 1e3), 'unit': 'tasks/s', 'h2d_bytes_per_step': bytes_in,
                        'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps},
                'gpu_launches': int(launches), 'clocks': clk, 'exchange': xch,
                'roofline': {'kernel': 'whole outer step (launch-bound small-patch work)', 'bound': 'tensor',
                             'achieved': task_tflop * T / (ms / args.steps / 1e3), 'peak': tc_sus, 'unit': 'TFLOP/s',
                             'frac': task_tflop * T / (ms / args.steps / 1e3) / tc_sus, 'traffic': None,
                             'peak_source': peak_src + ': bf16 dense sustained; algorithmic %.3f TFLOP per task' % task_tflop},
                'loss_q': lq}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def graph_time(fn, reps=20):
    """Average device time of one call of `fn`: `reps` back-to-back calls replayed from a CUDA graph, CUDA events on the
    replaying stream (the number is the kernel's, not the Python launch path's)."""
    import torch
    with torch.no_grad():
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
    return a.elapsed_time(b) / reps / 1000.0


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
    if args.impl == 'reference-cuda':
        run_reference_cuda(args, rank)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    if args.workload == 'meta':
        run_meta(args, rank, world, local)
        return
    import torch.distributed as dist
    from dynavsr_b200 import _lib, adapt, ops
    from dynavsr_b200.synth import seed_parameters
    from dynavsr_b200.models.archs import EDVR_arch, LRimg_estimator
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local))
    H, W = args.height, args.width
    use_tc = not args.no_tc
    ops.set_conv_backend(use_tc)                  # (the tcgen05 path is also the library default)
    inner_prec = (args.inner_precision, args.inner_backward_precision or args.inner_precision) if use_tc else None

    def build(seedG):
        netG = seed_parameters(EDVR_arch.EDVR(nf=64, nframes=NFR, groups=8, front_RBs=5, back_RBs=10, scale=SCALE), seedG)
        netE = seed_parameters(LRimg_estimator.DirectKernelEstimatorVideo(64, 3, SCALE), 77)
        netF = seed_parameters(LRimg_estimator.DirectKernelEstimatorVideo(64, 3, SCALE), 78)
        return netG.cuda(), netE.cuda(), netF.cuda()

    nets = build(1234)
    sds = tuple({k: v.detach().cpu().clone() for k, v in n.state_dict().items()} for n in nets)   # for the CPU oracle (parity)
    # frames in flight: 6 for the adaptation workload (latency-bound inner steps), 2 for plain inference (GPU-filling kernels)
    P = max(1, args.pipelines if args.pipelines is not None else (6 if args.workload == 'adapt' else 2))
    pool = adapt.AdaptationPool(*nets, pipelines=P, cta_budget=args.cta_budget, min_tiles_per_cta=args.min_tiles,
                                min_chunks_per_cta=args.wg_chunks, use_graphs=not args.no_graphs, inner_precision=inner_prec, **INNER)
    eng = pool.engines[0]
    # distinct windows per step and per rank (clip sharding: frame i -> rank i % world, train_dynavsr.py:509)
    n_clips = 4
    clips_host = [synth_clip(100 + rank * n_clips + i, H, W).pin_memory() for i in range(n_clips)]
    frames_dev = [ops.to_nhwc(c.cuda().reshape(NFR, 3, H, W)) for c in clips_host]
    # pinned output ring: two buffers per pipeline, so the host takes delivery of the frame a pipeline finished TWO rounds ago
    # before reusing its buffer and every pipeline's queue always holds its next frame (one buffer per pipeline left each
    # pipeline idle between its frame completing and the host's next submission: e2e 12 % below the resident-input number)
    RING = max(1, args.ring) * P
    hr_host = [torch.empty(1, 3, SCALE * H, SCALE * W).pin_memory() for _ in range(RING)]
    done = [None] * RING
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
        k = i % RING
        if done[k] is not None:
            done[k].synchronize()

        def work(e):
            x = clips_host[i % n_clips].cuda(non_blocking=True).reshape(NFR, 3, H, W)
            hr_host[k].copy_(ops.to_nchw(run(e, ops.to_nhwc(x))), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            return ev
        done[k] = pool.submit(work, pipeline=i % P)[1]

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

    # W untimed warm-up steps exactly as asked (>= 3); the pipelines' graphs were captured by pool.warm() above
    warm = max(args.warmup, 3)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, host_launches = timed(step_dev, args.steps, warm)
    clk = clocks.stop() if rank == 0 else None
    # the e2e leg fills its pinned-output ring once before the timed steps: the first pass through each ring slot / pipeline stream
    # allocates that stream's device staging blocks (cudaMalloc synchronises the device), which is set-up, not steady state
    # (tools/e2e_probe.py: 98.9-100.5 frames/s end to end once warm vs 77-88 when the first pass fell inside the timed region)
    e2e_warm = max(warm, RING)
    ms_e2e, _ = timed(step_e2e, args.steps, e2e_warm)
    value = world * args.steps / (ms / 1000.0)
    e2e = world * args.steps / (ms_e2e / 1000.0)
    # kernels launched per step: counted while the step was captured / run eagerly
    per_step = eng.launches_per_step if (getattr(eng, 'launches_per_step', None) and args.workload == 'adapt') \
        else host_launches / max(1, args.steps)

    # ---- rooflines, each kernel timed alone (whole GPU: default launch policy) with CUDA events, graph-replayed
    hbm_peak, tc_peak, tc_sus, peak_src = measured_peaks()
    prec = ops._backend['precision'] if use_tc else 'fp32'
    roofline = roofline_inner = roofline_dcn = None
    if not args.no_roofline:
        x = torch.randn(NFR, H, W, 64, device='cuda')
        wgt = torch.randn(64, 64, 3, 3, device='cuda') * 0.05
        bia = torch.zeros(64, device='cuda')
        t_conv = graph_time(lambda: ops.conv(x, wgt, bia, act=ops.ACT_RELU))
        flops = 18.0 * NFR * H * W * 64 * 64
        traffic, traffic_src = committed_traffic('conv_tc2_3x3_64_64_5x176x320_bf16x3') if (use_tc and prec == 'bf16x3' and (H, W) == (LR_H, LR_W)) else (None, None)
        roofline = {'kernel': 'conv3x3 64->64 fprop @%dx%dx%d (%s)' % (NFR, H, W, ('tcgen05 ' + prec) if use_tc else 'CUDA-core fp32'),
                    'bound': 'tensor', 'achieved': flops / t_conv / 1e12, 'peak': tc_peak, 'unit': 'TFLOP/s',
                    'frac': flops / t_conv / 1e12 / tc_peak, 'traffic': traffic, 'traffic_source': traffic_src,
                    'peak_source': peak_src + ': bf16 dense burst; algorithmic FLOPs = 18*N*H*W*Cin*Cout (the BF16x3 mode issues 3x '
                                              'that many tensor-core MACs: DESIGN.md section 3)',
                    'algorithmic_bytes': 4.0 * NFR * H * W * 128 + 36.0 * 64 * 64, 'launch_us': t_conv * 1e6,
                    'tensor_macs_issued_x': 3 if prec == 'bf16x3' else 1}
        # the single-product mode the inner steps run in (same kernel, same launch shape)
        roofline_inner = None
        if use_tc and inner_prec and inner_prec[0] != prec:
            with ops.conv_precision(inner_prec[0]):
                t_c1 = graph_time(lambda: ops.conv(x, wgt, bia, act=ops.ACT_RELU))
            roofline_inner = {'kernel': 'conv3x3 64->64 fprop @%dx%dx%d (tcgen05 %s)' % (NFR, H, W, inner_prec[0]), 'bound': 'tensor',
                              'achieved': flops / t_c1 / 1e12, 'peak': tc_peak, 'unit': 'TFLOP/s', 'frac': flops / t_c1 / 1e12 / tc_peak,
                              'launch_us': t_c1 * 1e6, 'traffic': None}
        # modulated deformable conv forward (HBM-bound): bytes = 4*N*H*W*(Cin + 3*dg*9 + Cout) + 4*Cout*Cin*9
        om = torch.cat([torch.randn(NFR, H, W, 144, device='cuda') * 1.5, torch.rand(NFR, H, W, 72, device='cuda')], 3).contiguous()
        t_dcn = graph_time(lambda: ops.mdcn(x, om, wgt, bia, 8, 1, 1, 1, ops.ACT_LRELU))
        dcn_bytes = 4.0 * NFR * H * W * (64 + 216 + 64) + 4.0 * 64 * 64 * 9
        dtraffic, dtraffic_src = committed_traffic('mdcn_fwd_5x176x320') if (use_tc and (H, W) == (LR_H, LR_W)) else (None, None)
        roofline_dcn = {'kernel': 'modulated DCN 3x3 64->64 dg8 forward @%dx%dx%d (%s)' % (NFR, H, W, 'tcgen05 gather+MMA' if use_tc else 'CUDA-core'),
                        'bound': 'hbm', 'achieved': dcn_bytes / t_dcn / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                        'frac': dcn_bytes / t_dcn / 1e9 / hbm_peak, 'traffic': dtraffic, 'traffic_source': dtraffic_src,
                        'algorithmic_bytes': dcn_bytes, 'launch_us': t_dcn * 1e6,
                        'tensor_equiv_tflops': 2.0 * NFR * H * W * 64 * 64 * 9 / t_dcn / 1e12}
    roofline_step = None
    if args.workload == 'adapt' and (H, W) == (LR_H, LR_W) and INNER['steps'] == 2:
        ach = ADAPTED_FRAME_TFLOP / (ms / args.steps / 1000.0)      # per GPU: each rank does `steps` frames in `ms`
        roofline_step = {'kernel': 'whole adapted frame (2 inner steps + final forward), %d frames in flight' % P, 'bound': 'tensor',
                         'achieved': ach, 'peak': tc_sus, 'unit': 'TFLOP/s', 'frac': ach / tc_sus,
                         'peak_source': peak_src + ': bf16 dense sustained; algorithmic %.2f TFLOP per adapted frame (SURVEY.md 8d)' % ADAPTED_FRAME_TFLOP}

    # ---- parity of the timed configuration at FULL SIZE against the CPU oracle (reference algorithm, fp32), same weights and
    # window; the library's own exact-fp32 CUDA-core path is reported beside it as a second field
    parity, cpu = None, None
    if rank == 0:
        fr = frames_dev[0]
        out = (eng.adapt_and_infer_nhwc(fr) if args.workload == 'adapt' else eng.infer_nhwc(fr)).detach().clone()
        out_nchw = ops.to_nchw(out).cpu()
        ref_cpu = None
        if not args.no_cpu_baseline and args.workload == 'adapt':
            cpu, ref_cpu = cpu_adapt_sample(2, 1, budget_s=60.0, full_hw=(H, W), sds=sds, clip=clips_host[0])
        if not args.no_parity:
            q = lambda t: (t.clamp(0, 1) * 255.0).round()

            def cmp(a, b):
                err = float((a.double() - b.double()).norm() / b.double().norm())
                mse = float(((q(a) - q(b)) ** 2).mean())
                return err, (None if mse == 0 else 20 * math.log10(255.0 / mse ** 0.5))
            if ref_cpu is None:
                from oracle import edvr_oracle as O
                torch.set_num_threads(os.cpu_count() or 1)
                if args.workload == 'adapt':
                    ref_cpu = O.adapt_and_infer(*sds, clips_host[0], **INNER)      # (the inner steps need autograd)
                else:
                    with torch.no_grad():
                        ref_cpu = O.edvr_forward(sds[0], clips_host[0])
            err, psnr = cmp(out_nchw, ref_cpu)
            parity = {'vs': 'oracle (reference algorithm, torch CPU fp32: oracle/edvr_oracle.%s), same weights and window, full size %dx%d' % (
                'adapt_and_infer' if args.workload == 'adapt' else 'edvr_forward', H, W), 'rel_l2': err,
                'psnr_db_between_uint8_outputs': psnr, 'tolerance': 1e-3, 'ok': err < 1e-3}
            if use_tc:
                ops.set_conv_backend(False)
                eng_ref = adapt.InnerLoopAdapter(*build(1234), use_graphs=False, **INNER)
                ref = (eng_ref.adapt_and_infer_nhwc(fr) if args.workload == 'adapt' else eng_ref.infer_nhwc(fr)).detach()
                ops.set_conv_backend(True)
                e2, p2 = cmp(out, ref)
                e3, _ = cmp(ops.to_nchw(ref).cpu(), ref_cpu)
                parity['vs_exact_fp32_path_of_this_library'] = {'rel_l2': e2, 'psnr_db_between_uint8_outputs': p2,
                                                                'exact_path_vs_oracle_rel_l2': e3}
                del eng_ref

    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_reference_cuda and args.workload == 'adapt':
        # the reference's own CUDA path on this GPU, in a separate process (it must not share this one's library state)
        try:
            torch.cuda.synchronize()
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference-cuda', '--steps', '5', '--warmup', '2',
                                '--height', str(H), '--width', str(W)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
            lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
            d = json.loads(lines[-1]) if lines else {'unavailable': (r.stderr or 'no output').strip().splitlines()[-1][:300]}
            ref_cuda = {k: d[k] for k in ('value', 'unit', 'ms_per_step', 'unavailable') if k in d}
            if 'value' in d:
                ref_cuda['arm'] = d['config']['arm']
                ref_cuda['ours_over_reference_cuda_e2e'] = e2e / world / d['value']
        except Exception as ex:
            ref_cuda = {'unavailable': str(ex)[:300]}

    if rank == 0:
        inner_fwd, inner_bwd = (inner_prec or ('fp32', 'fp32'))
        dtype = ('f32 storage; tcgen05 bf16 operands, fp32 accumulate: final forward BF16x3 (hi+lo split, 3 products), inner steps %s forward / %s backward'
                 % (inner_fwd, inner_bwd)) if use_tc else 'f32'
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': warm, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': dtype, 'data': 'synthetic',
                'config': {'workload': workload_string(H, W, args.workload),
                           'inner': INNER, 'inner_conv_precision': {'forward': inner_fwd, 'backward': inner_bwd}, 'clips_per_rank': n_clips,
                           'cuda_graphs': not args.no_graphs, 'frames_in_flight_per_gpu': P, 'e2e_pinned_output_ring': RING, 'e2e_warmup_steps': e2e_warm, 'launch_policy': eng.scope.policy.as_dict(),
                           'l2': 'per-step working set (activations ~GBs) exceeds the 126 MB L2; inputs rotate over %d clips' % n_clips,
                           'parallelism': 'clip-sharded dp%d, no data-path collective' % world},
                'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': NFR * 3 * H * W * 4,
                        'd2h_bytes_per_step': 3 * SCALE * H * SCALE * W * 4, 'ms_per_step': ms_e2e / args.steps},
                'gpu_launches': int(per_step * args.steps), 'clocks': clk, 'roofline': roofline, 'roofline_inner': roofline_inner,
                'roofline_dcn': roofline_dcn, 'roofline_step': roofline_step, 'cpu_baseline': cpu, 'reference_cuda': ref_cuda, 'parity': parity}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
