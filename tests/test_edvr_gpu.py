"""GPU parity of the assembled hot path against (a) the golden vectors produced by the unmodified
reference modules (tests/golden, oracle/make_golden.py) and (b) the CPU oracle on fresh seeded inputs.
Tolerance: 1e-3 relative (north star), PSNR within 0.01 dB; the fp32 CUDA-core path is expected to be
~1e-5."""
import numpy as np
import pytest
import torch

from util import gold, nchw, nhwc, psnr_uint8, rel

pytestmark = pytest.mark.gpu
NS_TOL = 1e-3


@pytest.fixture(scope='module')
def mods():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dynavsr_b200.models.archs import EDVR_arch, LRimg_estimator
    from dynavsr_b200 import adapt, ops
    return EDVR_arch, LRimg_estimator, adapt, ops


def _edvr(mods, seed, **kw):
    from oracle import params as P
    E = mods[0]
    net = E.EDVR(nf=64, nframes=5, groups=8, front_RBs=5, back_RBs=10, scale=4, **kw)
    sd = P.make_params(P.edvr_param_shapes(), seed=seed)
    net.load_state_dict(sd, strict=True)          # reference key names / shapes
    return net.cuda(), sd


def _mfdn(mods, seed):
    from oracle import params as P
    net = mods[1].DirectKernelEstimatorVideo(nf=64, in_nc=3, scale=4)
    sd = P.make_params(P.mfdn_param_shapes(), seed=seed)
    net.load_state_dict(sd, strict=True)
    return net.cuda(), sd


def test_edvr_forward_matches_reference_golden(mods):
    g = gold('edvr_m_32.npz')
    net, _ = _edvr(mods, int(g['seed']))
    x = torch.from_numpy(g['x']).cuda()
    with torch.no_grad():
        out = net(x)
    ref = torch.from_numpy(g['out'])
    assert out.shape == ref.shape
    assert rel(out, ref) < NS_TOL
    assert abs(psnr_uint8(out, ref)) > 60 or psnr_uint8(out, ref) == float('inf')
    # PSNR parity against a fixed pseudo ground truth (bicubic-free plumbing check, SURVEY config 1)
    base = torch.nn.functional.interpolate(x[:, 2], scale_factor=4, mode='bicubic', align_corners=False).cpu()
    assert abs(psnr_uint8(out, base) - psnr_uint8(ref, base)) < 0.01


def test_edvr_stage_outputs_match_reference_golden(mods):
    """Localises a failure: PCD-aligned features and the TSA output of the reference modules."""
    E, _, _, ops = mods
    g = gold('edvr_m_32.npz')
    net, _ = _edvr(mods, int(g['seed']))
    x = torch.from_numpy(g['x']).cuda()
    B, N = 1, 5
    with torch.no_grad():
        frames = ops.to_nhwc(x.reshape(B * N, 3, 32, 32))
        L1 = ops.conv(frames, net.conv_first.weight, net.conv_first.bias, act=ops.ACT_LRELU)
        L1 = net.feature_extraction(L1)
        L2 = E._c(E._c(L1, net.fea_L2_conv1, ops.ACT_LRELU), net.fea_L2_conv2, ops.ACT_LRELU)
        L3 = E._c(E._c(L2, net.fea_L3_conv1, ops.ACT_LRELU), net.fea_L3_conv2, ops.ACT_LRELU)
        assert rel(nchw(L3), torch.from_numpy(g['L3_fea'])) < NS_TOL, 'feature pyramid'
        ref = [ops.Seg(t.view(B, N, *t.shape[1:])[:, 2], T=N, Tsrc=1, t_fixed=0) for t in (L1, L2, L3)]
        aligned = net.pcd_align.forward_nhwc([L1, L2, L3], ref)
        assert rel(nchw(aligned[2:3]), torch.from_numpy(g['aligned_center'])) < NS_TOL, 'PCD (centre frame)'
        assert rel(nchw(aligned[0:1]), torch.from_numpy(g['aligned_0'])) < NS_TOL, 'PCD (frame 0)'
        tsa = net.tsa_fusion.forward_nhwc(aligned, B, N)
        assert rel(nchw(tsa), torch.from_numpy(g['tsa'])) < NS_TOL, 'TSA'


def test_edvr_reference_shaped_submodules(mods):
    """PCD_Align.forward / TSA_Fusion.forward keep the reference's NCHW call contract."""
    from oracle import edvr_oracle as O
    net, sd = _edvr(mods, 99)
    g = torch.Generator().manual_seed(1)
    nbr = [torch.randn(1, 64, 16 >> i, 24 >> i, generator=g) for i in range(3)]
    ref = [torch.randn(1, 64, 16 >> i, 24 >> i, generator=g) for i in range(3)]
    with torch.no_grad():
        y = net.pcd_align([t.cuda() for t in nbr], [t.cuda() for t in ref])
        y_ref = O.pcd_align({k: v.double() for k, v in sd.items()}, 'pcd_align.', [t.double() for t in nbr],
                            [t.double() for t in ref], 8)
        assert rel(y, y_ref) < NS_TOL
        al = torch.randn(2, 5, 64, 8, 12, generator=g)
        t = net.tsa_fusion(al.cuda())
        t_ref = O.tsa_fusion({k: v.double() for k, v in sd.items()}, 'tsa_fusion.', al.double(), 2)
        assert rel(t, t_ref) < NS_TOL


def test_edvr_batch2_and_gradients_vs_oracle(mods):
    """B=2 (exercises the strided centre-frame views) and the full backward pass against autograd
    through the CPU oracle (float64)."""
    from oracle import edvr_oracle as O
    net, sd = _edvr(mods, 321)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(2, 5, 3, 16, 16, generator=g)
    gt = torch.rand(2, 3, 64, 64, generator=g)
    sdd = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    out_ref = O.edvr_forward(sdd, x.double())
    loss_ref = ((out_ref - gt.double()) ** 2).mean()
    probes = ['conv_first.weight', 'pcd_align.L3_dcnpack.conv_offset_mask.weight', 'pcd_align.L1_dcnpack.weight',
              'pcd_align.cas_offset_conv1.weight', 'tsa_fusion.tAtt_2.weight', 'tsa_fusion.fea_fusion.weight',
              'tsa_fusion.sAtt_L2.weight', 'recon_trunk.0.conv1.weight', 'upconv1.weight', 'upconv2.bias',
              'conv_last.weight', 'conv_last.bias', 'fea_L2_conv1.weight', 'pcd_align.L2_offset_conv2.bias']
    gref = torch.autograd.grad(loss_ref, [sdd[k] for k in probes])
    out = net(x.cuda())
    assert rel(out, out_ref) < NS_TOL
    loss = ((out - gt.cuda()) ** 2).mean()
    loss.backward()
    named = dict(net.named_parameters())
    errs = {k: rel(named[k].grad, gr) for k, gr in zip(probes, gref)}
    bad = {k: v for k, v in errs.items() if not v < NS_TOL}
    assert not bad, 'gradient mismatches: %s' % bad


def test_mfdn_matches_reference_golden(mods):
    g = gold('mfdn_32x48.npz')
    net, _ = _mfdn(mods, int(g['seed']))
    lr = torch.from_numpy(g['lr']).cuda()
    with torch.no_grad():
        slr = net(lr.transpose(1, 2))                 # [B, C, T, h, w], the reference module's own contract
    assert rel(slr, torch.from_numpy(g['slr'])) < NS_TOL


@pytest.mark.parametrize('tc', [False, True], ids=['exact_fp32', 'tcgen05'])
def test_sfdn_matches_reference_golden(mods, tc):
    """SFDN (DirectKernelEstimator_CMS, LRimg_estimator.py:38-67; networks.define_E 'SFDN') vs the golden of the unmodified
    reference module: output, input gradient (mean kept in the graph) and weight / bias gradients."""
    from oracle import params as P
    L, ops = mods[1], mods[3]
    g = gold('sfdn_32x48.npz')
    net = L.DirectKernelEstimator_CMS(nf=64)
    net.load_state_dict(P.make_params(P.sfdn_param_shapes(64), seed=int(g['seed'])), strict=True)
    net = net.cuda()
    x = torch.from_numpy(g['x']).cuda().requires_grad_(True)
    ops.set_conv_backend(tc)
    try:
        out = net(x)
        (out * torch.from_numpy(g['probe']).cuda()).sum().backward()
    finally:
        ops.set_conv_backend(False)
    tol = 1e-3 if tc else 2e-5
    assert out.shape == (3, 3, 16, 24) and rel(out, torch.from_numpy(g['out'])) < tol
    # gradients of the reduced-precision path against an fp32 reference have an inherent floor: a pre-activation within the
    # rounding error of zero flips its LeakyReLU mask (DESIGN.md "Numerics"), six layers deep here -- wide band for the
    # tensor-core path, the exact path pins the gradient formulas (incl. the mean term: measured 2e-2 vs 1e-6)
    gtol = 5e-2 if tc else 1e-4
    assert rel(x.grad, torch.from_numpy(g['gx'])) < gtol
    assert rel(net.conv0.weight.grad, torch.from_numpy(g['g_conv0_w'])) < gtol
    assert rel(net.conv3.weight.grad, torch.from_numpy(g['g_conv3_w'])) < gtol
    assert rel(net.conv6.bias.grad, torch.from_numpy(g['g_conv6_b'])) < gtol


def test_mfdn_gradients_vs_oracle(mods):
    from oracle import edvr_oracle as O
    net, sd = _mfdn(mods, 5)
    g = torch.Generator().manual_seed(3)
    lr = torch.rand(1, 5, 3, 16, 24, generator=g)
    tgt = torch.rand(1, 5, 3, 4, 6, generator=g)
    sdd = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    ref = O.mfdn_forward(sdd, lr.double().transpose(1, 2)).transpose(1, 2)
    gref = torch.autograd.grad((ref - tgt.double()).abs().mean(), list(sdd.values()))
    out = net(lr.cuda().transpose(1, 2)).transpose(1, 2)
    assert rel(out, ref) < NS_TOL
    (out - tgt.cuda()).abs().mean().backward()
    for (k, p), gr in zip(net.named_parameters(), gref):
        assert rel(p.grad, gr) < NS_TOL, k


@pytest.mark.parametrize('tag,optimizer,crit', [('sgd2_l2', 'SGD', 'l2'), ('adam1_cb', 'Adam', 'cb')])
@pytest.mark.parametrize('graphs', [False, True], ids=['eager', 'graph'])
def test_adaptation_matches_reference_golden(mods, tag, optimizer, crit, graphs):
    """test_dynavsr.py:208-283 (2-step SGD/L2 and the shipped 1-step Adam/Charbonnier setting)."""
    adapt = mods[2]
    g = gold('adapt_%s.npz' % tag)
    netG, sdG = _edvr(mods, int(g['seed_G']))
    netE, sdE = _mfdn(mods, int(g['seed_E']))
    netF, _ = _mfdn(mods, int(g['seed_E_fixed']))
    eng = adapt.InnerLoopAdapter(netG, netE, netF, steps=int(g['steps']), lr_alpha=float(g['lr_alpha']),
                                 optimizer=optimizer, betas=(0.9, 0.99), criterion=crit, slr_weight=10.0,
                                 use_graphs=graphs)
    lr = torch.from_numpy(g['lr'])
    ref = torch.from_numpy(g['out'])
    for rep in range(2):                       # second frame must restart from the meta-weights (:208)
        hr = eng.adapt_and_infer(lr)
        assert rel(hr, ref) < NS_TOL, 'rep %d' % rep
        assert np.allclose(eng.last_losses.cpu().numpy(), g['losses'], rtol=1e-3), 'rep %d' % rep
        # the adaptation must actually have moved the output (sensitivity of this test)
        assert rel(hr, torch.from_numpy(g['out_unadapted'])) > 5 * rel(hr, ref)
    # parameter deltas of two probes = lr * gradient, checked against the reference run
    d_first = netG.conv_first.weight.detach().cpu() - sdG['conv_first.weight']
    d_conv6 = netE.conv6.weight.detach().cpu() - sdE['conv6.weight']
    assert rel(d_first, torch.from_numpy(g['d_conv_first'])) < 2e-2
    assert rel(d_conv6, torch.from_numpy(g['d_conv6'])) < 2e-2
    # and un-adapted inference still reproduces the meta-weights result
    frames = nhwc(lr.reshape(5, 3, *lr.shape[-2:]).cuda())
    assert rel(nchw(eng.infer_nhwc(frames)), torch.from_numpy(g['out_unadapted'])) < NS_TOL


def test_adaptation_pool_frames_in_flight_match_golden(mods):
    """adapt.AdaptationPool: 3 pipelines (own parameter copy / packs / graphs / streams) process 7 windows concurrently;
    every output must equal the reference golden of its window, and the windows must not bleed into each other."""
    adapt = mods[2]
    g = gold('adapt_sgd2_l2.npz')
    netG, _ = _edvr(mods, int(g['seed_G']))
    netE, _ = _mfdn(mods, int(g['seed_E']))
    netF, _ = _mfdn(mods, int(g['seed_E_fixed']))
    pool = adapt.AdaptationPool(netG, netE, netF, pipelines=3, steps=int(g['steps']), lr_alpha=float(g['lr_alpha']),
                                optimizer='SGD', criterion='l2', slr_weight=10.0, use_graphs=True)
    lr = torch.from_numpy(g['lr'])
    ref = torch.from_numpy(g['out'])
    fr = nhwc(lr.reshape(5, 3, *lr.shape[-2:]).cuda())
    other = fr.flip(0).contiguous()                     # a different window (reversed frame order)
    single = nchw(pool.engines[0].adapt_and_infer_nhwc(other).clone())
    outs = pool.adapt_and_infer_many([fr, other, fr, fr, other, fr, other])
    torch.cuda.synchronize()
    for i, o in enumerate(outs):
        want = ref if i in (0, 2, 3, 5) else single
        assert rel(nchw(o), want) < (NS_TOL if want is ref else 1e-5), 'window %d' % i
    assert rel(single, ref) > 1e-2                      # the two windows really differ


@pytest.mark.parametrize('tag,kw', [('predeblur', dict(predeblur=True, HR_in=False, w_TSA=True)),
                                    ('hrin_notsa', dict(predeblur=False, HR_in=True, w_TSA=False)),
                                    ('predeblur_hrin', dict(predeblur=True, HR_in=True, w_TSA=True))])
def test_edvr_variants_match_reference_golden(mods, tag, kw):
    """Constructor variants no YML uses (EDVR_arch.py:13-57,208-239): strict state_dict loading, forward vs the golden of the
    unmodified reference module, gradients of two probes vs the oracle."""
    from oracle import edvr_oracle as O
    from oracle import params as P
    g = gold('edvr_variants.npz')
    cfg = dict(nf=64, nframes=5, groups=8, front_RBs=1, back_RBs=1, scale=4)
    sd = P.make_params(P.edvr_param_shapes(**cfg, **kw), seed=int(g[tag + '_seed']))
    net = mods[0].EDVR(**cfg, **kw)
    net.load_state_dict(sd, strict=True)
    net.cuda()
    x = torch.from_numpy(g[tag + '_x'])
    y = net(x.cuda())
    assert rel(y, torch.from_numpy(g[tag + '_out'])) < 1e-5
    # float64 oracle: two fp32 implementations flip different ReLU masks at pre-activations within rounding of zero
    ref = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    yr = O.edvr_forward(ref, x.double(), front_RBs=1, back_RBs=1, predeblur_=kw['predeblur'], HR_in=kw['HR_in'], w_TSA=kw['w_TSA'])
    gy = torch.randn(yr.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    first = next(iter(sd.keys()))
    fus = 'tsa_fusion.weight' if not kw['w_TSA'] else 'tsa_fusion.fea_fusion.weight'
    gr = torch.autograd.grad(yr, [ref[first], ref[fus]], gy)
    params = dict(net.named_parameters())
    gd = torch.autograd.grad(y, [params[first], params[fus]], gy.float().cuda())
    # band, not NS_TOL: with only ~65k-260k activations per layer at this size, each ReLU / LeakyReLU mask that flips
    # between the fp32 kernels and the fp64 oracle (pre-activation within ~1e-6 of zero) moves a weight gradient by
    # 2e-3 - 4e-3; a seed sweep gave 0-4 flips per run.  The main configuration is held to NS_TOL above.
    assert rel(gd[0], gr[0]) < 3e-2 and rel(gd[1], gr[1]) < 3e-2


def test_full_size_properties(mods):
    """BASELINE size (5x3x180x320 -> 3x720x1280): properties that need no oracle run --
    determinism, batch-vs-single consistency of the batched PCD pass, and shape."""
    ops = mods[3]
    net, _ = _edvr(mods, 7)
    g = torch.Generator().manual_seed(4)
    x = torch.rand(1, 5, 3, 180, 320, generator=g).cuda()
    with torch.no_grad():
        a = net(x)
        b = net(x)
        assert a.shape == (1, 3, 720, 1280)
        assert torch.equal(a, b)                                   # forward is deterministic (no atomics)
        # a crop processed alone must equal the crop of the full frame away from the borders
        # (receptive field of EDVR-M is large; compare a centre window with generous margin at L3 scale)
        assert torch.isfinite(a).all()
    xs = torch.cat([x, x.flip(1)], 0)
    with torch.no_grad():
        c = net(xs)
    assert rel(c[0:1], a) < 1e-6                                   # batching does not change results


# ---------------------------------------------------------------------------------------------------
# tensor-core path (TF32 operands rounded to nearest, fp32 accumulation): same goldens, north-star tolerance.
@pytest.fixture()
def tc(mods):
    ops = mods[3]
    ops.set_conv_backend(True)
    yield ops
    ops.set_conv_backend(False)


def test_edvr_tc_forward_matches_reference_golden(mods, tc):
    g = gold('edvr_m_32.npz')
    net, _ = _edvr(mods, int(g['seed']))
    x = torch.from_numpy(g['x']).cuda()
    with torch.no_grad():
        out = net(x)
    ref = torch.from_numpy(g['out'])
    assert rel(out, ref) < NS_TOL                     # BF16x3: measured 1.7e-4 (the round-1 TF32 build sat at 9.3e-4)
    base = torch.nn.functional.interpolate(x[:, 2], scale_factor=4, mode='bicubic', align_corners=False).cpu()
    assert abs(psnr_uint8(out, base) - psnr_uint8(ref, base)) < 0.01


def test_adaptation_tc_matches_reference_golden(mods, tc):
    """The benchmark configuration (2 SGD steps, L2 + 10*L1(SLR)) on the tcgen05 path, CUDA graphs on."""
    adapt = mods[2]
    g = gold('adapt_sgd2_l2.npz')
    netG, _ = _edvr(mods, int(g['seed_G']))
    netE, _ = _mfdn(mods, int(g['seed_E']))
    netF, _ = _mfdn(mods, int(g['seed_E_fixed']))
    eng = adapt.InnerLoopAdapter(netG, netE, netF, steps=2, lr_alpha=float(g['lr_alpha']), optimizer='SGD',
                                 criterion='l2', slr_weight=10.0, use_graphs=True)
    ref = torch.from_numpy(g['out'])
    for rep in range(2):
        hr = eng.adapt_and_infer(torch.from_numpy(g['lr']))
        assert rel(hr, ref) < NS_TOL, 'rep %d' % rep  # BF16x3: measured 1.7e-4 (TF32 build: 9.1e-4)
        assert np.allclose(eng.last_losses.cpu().numpy(), g['losses'], rtol=2e-3)
        assert rel(hr, torch.from_numpy(g['out_unadapted'])) > 5 * rel(hr, ref)


def test_full_size_tc_vs_fp32_cuda_core_path(mods):
    """BASELINE size (5x3x176x320 -> 3x704x1280): the tcgen05 path against the exact-fp32 CUDA-core path of
    this library on the same weights / input (the oracle is too slow at this size).  Bar: 1e-3 relative and
    0.01 dB PSNR."""
    ops = mods[3]
    net, _ = _edvr(mods, 7)
    g = torch.Generator().manual_seed(4)
    x = torch.rand(1, 5, 3, 176, 320, generator=g).cuda()
    with torch.no_grad():
        ops.set_conv_backend(False)
        ref = net(x)
        ops.set_conv_backend(True)
        try:
            out = net(x)
            again = net(x)
        finally:
            ops.set_conv_backend(False)
    assert torch.equal(out, again)                    # deterministic
    assert rel(out, ref) < NS_TOL
    base = torch.nn.functional.interpolate(x[:, 2], scale_factor=4, mode='bicubic', align_corners=False)
    assert abs(psnr_uint8(out, base) - psnr_uint8(ref, base)) < 0.01


def test_vid4_shaped_window_tc_vs_exact_path(mods):
    """BASELINE config 3 shape (Vid4 'city' LR 176x144 -> SLR 44x36): the tensor-core adaptation path against this library's
    exact-fp32 path on the same window (CUDA graphs on), north-star tolerance."""
    adapt, ops = mods[2], mods[3]
    g = torch.Generator().manual_seed(9)
    clip = torch.rand(1, 5, 3, 176, 144, generator=g)

    def run(tc):
        ops.set_conv_backend(tc)
        try:
            netG, _ = _edvr(mods, 1234)
            netE, _ = _mfdn(mods, 77)
            netF, _ = _mfdn(mods, 78)
            eng = adapt.InnerLoopAdapter(netG, netE, netF, steps=2, lr_alpha=1e-5, optimizer='SGD', criterion='l2', use_graphs=tc)
            out = eng.adapt_and_infer(clip).cpu()
            return out, eng.last_losses.cpu()
        finally:
            ops.set_conv_backend(False)
    ref, lref = run(False)
    out, lout = run(True)
    assert out.shape == (1, 3, 704, 576)
    assert rel(out, ref) < NS_TOL and abs(psnr_uint8(out, ref)) > 60
    assert np.allclose(lout.numpy(), lref.numpy(), rtol=2e-3)


def test_graph_engine_follows_new_meta_weights(mods):
    """New meta-weights after the graphs were captured (e.g. after a meta-training step): the captured restore-copy must pick
    up the refreshed pack snapshot -- outputs equal those of a fresh engine built on the new weights (up to the run-to-run
    jitter of the atomically accumulated weight gradients, ~3e-5 here)."""
    adapt, ops = mods[2], mods[3]
    from oracle import params as P
    g = torch.Generator().manual_seed(3)
    clip = torch.rand(1, 5, 3, 32, 48, generator=g)
    ops.set_conv_backend(True)
    try:
        netG, _ = _edvr(mods, 1)
        netE, _ = _mfdn(mods, 2)
        netF, _ = _mfdn(mods, 3)
        eng = adapt.InnerLoopAdapter(netG, netE, netF, steps=1, lr_alpha=1e-4, optimizer='SGD', criterion='l2', use_graphs=True)
        a = eng.adapt_and_infer(clip).clone()
        new = P.make_params(P.edvr_param_shapes(), seed=99)
        eng.set_meta_weights(state_dict_G=new)       # restore (drop the last frame's adaptation), load, snapshot
        b = eng.adapt_and_infer(clip).clone()
        b2 = eng.adapt_and_infer(clip).clone()       # and again: restart from the NEW meta-weights
        netG2, _ = _edvr(mods, 99)
        netE2, _ = _mfdn(mods, 2)
        netF2, _ = _mfdn(mods, 3)
        fresh = adapt.InnerLoopAdapter(netG2, netE2, netF2, steps=1, lr_alpha=1e-4, optimizer='SGD', criterion='l2', use_graphs=False)
        c = fresh.adapt_and_infer(clip)
    finally:
        ops.set_conv_backend(False)
    assert rel(b, c) < 1e-4 and rel(b2, c) < 1e-4
    assert rel(a, c) > 1e-2


def _driver_engine(mods, g, **kw):
    from oracle import params as P
    adapt = mods[2]
    gain = float(g['head_gain'])
    nets = []
    for seed in (int(g['seed_G']), int(g['seed_baseline_G'])):
        net, sd = _edvr(mods, seed)
        sd['conv_last.weight'] *= gain             # oracle/make_golden_driver.tame_head: outputs inside [0, 1]
        sd['conv_last.bias'] *= gain
        net.load_state_dict(sd, strict=True)
        nets.append(net)
    netE, _ = _mfdn(mods, int(g['seed_E']))
    netF, _ = _mfdn(mods, int(g['seed_E_fixed']))
    eng = adapt.InnerLoopAdapter(nets[0], netE, netF, steps=int(g['steps']), lr_alpha=float(g['lr_alpha']),
                                 optimizer=str(g['optimizer']), betas=(0.9, 0.99), criterion=str(g['criterion']),
                                 slr_weight=10.0, **kw)
    return eng, nets[1]


@pytest.mark.parametrize('backend', ['fp32', 'tensor-core'])
@pytest.mark.parametrize('tag', ['adam1_cb', 'sgd2_l2'])
def test_adaptation_vs_reference_test_driver(mods, tag, backend):
    """The north-star criterion against the reference's OWN driver: tests/golden/driver_<tag>.npz holds the frame, the PNG
    bytes and the PSNR numbers the unmodified test_dynavsr.py main() produced for one clip (oracle/make_golden_driver.py).
    Product: frame within 1e-3 relative, PSNR within 0.01 dB of the driver's psnr_update.csv, for both kernel paths."""
    ops = mods[3]
    g = gold('driver_%s.npz' % tag)
    lq, gt = torch.from_numpy(g['lq']), torch.from_numpy(g['gt'])
    ops.set_conv_backend(backend == 'tensor-core')
    try:
        eng, baseline = _driver_engine(mods, g, use_graphs=(backend == 'tensor-core'))
        for rep in range(2):
            hr = eng.adapt_and_infer(lq)[0].float().cpu()
            assert rel(hr.clamp(0, 1), torch.from_numpy(g['out'])) < NS_TOL, rep
            assert abs(psnr_uint8(hr, gt) - float(g['psnr_adapted'])) < 0.01, rep
            image = (hr.clamp(0, 1) * 255.0).round().permute(1, 2, 0).numpy().astype(np.int32)
            assert int(np.abs(image - g['image'].astype(np.int32)).max()) <= 1       # PNG bytes: at most one level off
        with torch.no_grad():
            base = baseline(lq.cuda())[0].float().cpu()
        assert abs(psnr_uint8(base, gt) - float(g['psnr_baseline'])) < 0.01
    finally:
        ops.set_conv_backend(False)


@pytest.mark.parametrize('backend', ['fp32', 'tensor-core'])
@pytest.mark.parametrize('tag', ['sgd2_l2_patch', 'adam1_cb_real'])
def test_optional_inner_loop_branches_vs_reference_test_driver(mods, tag, backend):
    """``maml.use_patch`` (B = num_patch random crops, test_dynavsr.py:118-145,255-260) and ``train.use_real`` (EDVR-only
    adaptation on the dataset's super-LR clip, :218-221,243-244) against what the UNMODIFIED driver produced.  use_patch is
    checked twice: with the logged crop positions passed in, and with Python's ``random`` seeded like the harness did -- the
    engine must then draw the very same positions (py before px, per patch, per step)."""
    import random
    ops = mods[3]
    g = gold('driver_%s.npz' % tag)
    lq, gt = torch.from_numpy(g['lq']), torch.from_numpy(g['gt'])
    kw, call_kw = {}, [{}]
    if bool(g['use_real']):
        kw.update(use_real=True)
        call_kw = [dict(slr_clip=torch.from_numpy(g['slq']))]
    if bool(g['use_patch']):
        n, c = int(g['num_patch']), g['crops'].tolist()
        pos = [(c[2 * i], c[2 * i + 1]) for i in range(len(c) // 2)]
        kw.update(use_patch=True, num_patch=n, patch_size=int(g['patch_size']))
        call_kw = [dict(patch_positions=[pos[k * n:(k + 1) * n] for k in range(int(g['steps']))]), {}]
    ops.set_conv_backend(backend == 'tensor-core')
    try:
        eng, _ = _driver_engine(mods, g, use_graphs=(backend == 'tensor-core'), **kw)
        e0 = {k: v.detach().clone() for k, v in eng.netE.state_dict().items()}
        for ck in call_kw:
            random.seed(int(g['patch_seed']))
            hr = eng.adapt_and_infer(lq, **ck)[0].float().cpu()
            assert rel(hr.clamp(0, 1), torch.from_numpy(g['out'])) < NS_TOL
            assert abs(psnr_uint8(hr, gt) - float(g['psnr_adapted'])) < 0.01
            if bool(g['use_real']):       # MFDN receives no gradient: bit-identical after the update
                assert all(torch.equal(v, e0[k]) for k, v in eng.netE.state_dict().items())
        with pytest.raises(RuntimeError):
            if bool(g['use_real']):
                eng.adapt_and_infer(lq)                      # use_real without the super-LR clip
            else:
                raise RuntimeError('n/a')
    finally:
        ops.set_conv_backend(False)
