"""Parity at the BASELINE shapes against the CPU ORACLE (VERDICT r1 "What's weak" #1): the headline 5x3x176x320 REDS window,
the Vid4 shape 176x144 and the 64x64 config-1 case, BOTH kernel paths (exact-fp32 CUDA-core and tcgen05), eager and CUDA
graphs, single engine and the 6-pipeline pool, for both shipped inner settings (SGD-2-L2 and Adam-1-Charbonnier) and the
per-phase precision policy (single-product bf16 inner steps).  The oracle (oracle/edvr_oracle.py: the reference algorithm in
plain PyTorch fp32 on the host cores, pinned against the unmodified reference modules AND test driver) needs ~3-5 s per
adapted frame at 176x320 on the GPU box's cores; every oracle result is computed once per session.

Bar (north star): 1e-3 relative, PSNR within 0.01 dB.  These are the large-tile paths the 32x32 goldens do not reach: staged
DCN window (>= 2 tiles per SM), persistent multi-tile CTAs, the 37-CTA budget with 6 streams.
"""
import functools

import numpy as np
import pytest
import torch

from util import nchw, psnr_uint8, rel

pytestmark = pytest.mark.gpu
NS_TOL = 1e-3
SHAPES = {'reds_176x320': (176, 320), 'vid4_176x144': (176, 144), 'config1_64x64': (64, 64)}
SETTINGS = {'sgd2_l2': dict(steps=2, lr_alpha=1e-5, optimizer='SGD', criterion='l2', slr_weight=10.0),
            'adam1_cb': dict(steps=1, lr_alpha=1e-5, optimizer='Adam', betas=(0.9, 0.99), criterion='cb', slr_weight=10.0)}


@pytest.fixture(scope='module')
def mods():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dynavsr_b200.models.archs import EDVR_arch, LRimg_estimator
    from dynavsr_b200 import adapt, ops
    return EDVR_arch, LRimg_estimator, adapt, ops


def _sds():
    from oracle import params as P
    return (P.make_params(P.edvr_param_shapes(), seed=1234), P.make_params(P.mfdn_param_shapes(), seed=77),
            P.make_params(P.mfdn_param_shapes(), seed=78))


def _clip(shape):
    from dynavsr_b200.synth import synth_clip
    H, W = SHAPES[shape]
    return synth_clip(21, H, W)


@functools.lru_cache(maxsize=None)
def oracle_forward(shape):
    from oracle import edvr_oracle as O
    import os
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        return O.edvr_forward(_sds()[0], _clip(shape))


@functools.lru_cache(maxsize=None)
def oracle_adapted(shape, setting):
    from oracle import edvr_oracle as O
    import os
    torch.set_num_threads(os.cpu_count() or 1)
    out, losses, _, _ = O.adapt_and_infer(*_sds(), _clip(shape), return_losses=True, **SETTINGS[setting])
    return out, losses


def _nets(mods):
    E, L = mods[0], mods[1]
    sdG, sdE, sdF = _sds()
    netG = E.EDVR(nf=64, nframes=5, groups=8, front_RBs=5, back_RBs=10, scale=4)
    netG.load_state_dict(sdG, strict=True)
    netE, netF = L.DirectKernelEstimatorVideo(64, 3, 4), L.DirectKernelEstimatorVideo(64, 3, 4)
    netE.load_state_dict(sdE, strict=True)
    netF.load_state_dict(sdF, strict=True)
    return netG.cuda(), netE.cuda(), netF.cuda()


def _check(out, ref, x_center, what):
    assert out.shape == ref.shape, what
    e = rel(out, ref)
    assert e < NS_TOL, '%s: rel %.3e' % (what, e)
    base = torch.nn.functional.interpolate(x_center, scale_factor=4, mode='bicubic', align_corners=False)
    d = abs(psnr_uint8(out, base) - psnr_uint8(ref, base))
    assert d < 0.01, '%s: PSNR differs by %.4f dB' % (what, d)
    return e


@pytest.mark.parametrize('tc', [False, True], ids=['exact_fp32', 'tcgen05'])
@pytest.mark.parametrize('shape', list(SHAPES))
def test_edvr_forward_vs_oracle_at_baseline_shapes(mods, shape, tc):
    """EDVR.forward at full size vs the oracle (EDVR_arch.py:254-313)."""
    ops = mods[3]
    clip = _clip(shape)
    ref = oracle_forward(shape)
    ops.set_conv_backend(tc)
    try:
        netG = _nets(mods)[0]
        with torch.no_grad():
            out = netG(clip.cuda()).cpu()
    finally:
        ops.set_conv_backend(False)
    e = _check(out, ref, clip[:, 2], 'EDVR forward %s tc=%s' % (shape, tc))
    if not tc:
        assert e < 2e-5                       # the exact path is fp32-class at every size


@pytest.mark.parametrize('mode', ['exact_eager', 'tc_eager', 'tc_graph', 'tc_graph_inner_bf16'])
@pytest.mark.parametrize('setting', list(SETTINGS))
@pytest.mark.parametrize('shape', ['reds_176x320', 'vid4_176x144'])
def test_adapted_frame_vs_oracle_at_baseline_shapes(mods, shape, setting, mode):
    """test_dynavsr.py:208-283 for one frame at the BASELINE shapes vs the oracle: both kernel paths, eager and graph, and
    the per-phase precision policy -- single-product bf16 operands throughout the inner steps for SGD, in the inner BACKWARD
    only for Adam (its normalised step amplifies forward errors: profiles/r1_precision_study.md)."""
    adapt, ops = mods[2], mods[3]
    clip = _clip(shape)
    ref, ref_losses = oracle_adapted(shape, setting)
    tc = mode != 'exact_eager'
    prec = None
    if mode == 'tc_graph_inner_bf16':
        prec = ('bf16', 'bf16') if SETTINGS[setting]['optimizer'] == 'SGD' else (None, 'bf16')
    ops.set_conv_backend(tc)
    try:
        eng = adapt.InnerLoopAdapter(*_nets(mods), use_graphs=mode.startswith('tc_graph'), inner_precision=prec, **SETTINGS[setting])
        outs = [eng.adapt_and_infer(clip).cpu() for _ in range(2 if tc else 1)]       # twice: restore + replay
        losses = eng.last_losses.cpu().numpy()
    finally:
        ops.set_conv_backend(False)
    for i, out in enumerate(outs):
        _check(out, ref, clip[:, 2], 'adapted %s %s %s rep %d' % (shape, setting, mode, i))
    assert np.allclose(losses, ref_losses, rtol=5e-3 if prec else 2e-3)
    un = oracle_forward(shape)
    assert rel(outs[0], un) > 3 * rel(outs[0], ref)          # the adaptation moved the frame more than our error


def test_pool_of_six_pipelines_vs_oracle_full_size(mods):
    """adapt.AdaptationPool with 6 frames in flight (the bench configuration: 37-CTA budget, >= 2 tiles per CTA, 24 chunks per
    wgrad CTA, single-product inner steps) on the REDS window: every pipeline's frame must match the oracle."""
    adapt, ops = mods[2], mods[3]
    clip = _clip('reds_176x320')
    ref, _ = oracle_adapted('reds_176x320', 'sgd2_l2')
    ops.set_conv_backend(True)
    try:
        pool = adapt.AdaptationPool(*_nets(mods), pipelines=6, inner_precision=('bf16', 'bf16'), **SETTINGS['sgd2_l2'])
        fr = ops.to_nhwc(clip.cuda().reshape(5, 3, *clip.shape[-2:]))
        pool.warm(fr)
        outs = pool.adapt_and_infer_many([fr.clone() for _ in range(12)])
        torch.cuda.synchronize()
        outs = [nchw(o).cpu() for o in outs]
    finally:
        ops.set_conv_backend(False)
    for i, o in enumerate(outs):
        _check(o, ref, clip[:, 2], 'pool frame %d' % i)
    # frames handled by the same pipeline are bit-identical replays of one graph apart from the split-K summation order of the
    # weight gradients (atomics): bounded, far below the tolerance
    assert max(rel(outs[i], outs[i + 6]) for i in range(6)) < 1e-4


def test_two_pools_with_different_policies_do_not_interact(mods):
    """the launch policy travels in each descriptor (dvsr_policy), so two pools with different policies in one
    process give the results each gives alone -- interleaved frame by frame."""
    adapt, ops = mods[2], mods[3]
    H, W = 64, 96
    from dynavsr_b200.synth import synth_clip
    clips = [synth_clip(30 + i, H, W) for i in range(3)]
    frs = [ops.to_nhwc(c.cuda().reshape(5, 3, H, W)) for c in clips]
    kw = dict(use_graphs=True, **SETTINGS['sgd2_l2'])
    ops.set_conv_backend(True)
    try:
        def alone(**pol):
            pool = adapt.AdaptationPool(*_nets(mods), **pol, **kw)
            return [nchw(o).cpu() for o in pool.adapt_and_infer_many([f.clone() for f in frs])]
        polA = dict(pipelines=1)                                                    # whole GPU, 1 tile per CTA, 4 chunks
        polB = dict(pipelines=3, cta_budget=20, min_tiles_per_cta=3, min_chunks_per_cta=16)
        wantA, wantB = alone(**polA), alone(**polB)
        pa, pb = adapt.AdaptationPool(*_nets(mods), **polA, **kw), adapt.AdaptationPool(*_nets(mods), **polB, **kw)
        assert pa.engines[0].scope.policy.as_dict() != pb.engines[0].scope.policy.as_dict()
        gotA, gotB = [], []
        for f in frs:                                                              # interleave the two pools
            gotA.append(pa.adapt_and_infer_many([f.clone()])[0])
            gotB.append(pb.adapt_and_infer_many([f.clone()])[0])
        torch.cuda.synchronize()
        gotA, gotB = [nchw(o).cpu() for o in gotA], [nchw(o).cpu() for o in gotB]
    finally:
        ops.set_conv_backend(False)
    # forward kernels are deterministic for a given policy; the weight-gradient split-K order (atomics) leaves ~1e-5 jitter
    for w, g in zip(wantA + wantB, gotA + gotB):
        assert rel(g, w) < 1e-4
    assert max(rel(a, b) for a, b in zip(wantA, wantB)) < 1e-4                      # and the policy does not change the numbers
