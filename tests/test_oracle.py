"""CPU tests: the oracle against the golden vectors generated from the unmodified reference
(oracle/make_golden.py) and against torchvision's implementation of the same DCN algorithm."""
import numpy as np
import pytest
import torch

from oracle import edvr_oracle as O
from oracle import mdcn_c
from oracle import params as P
from oracle.torch_ops import mdcn_torch
from util import gold, rel


def test_c_oracle_matches_golden_dcn():
    g = gold('dcn_small.npz')
    dg = int(g['dg'])
    y = mdcn_c.forward(g['x'], g['offset'], g['mask'], g['weight'], g['bias'], 1, 1, 1, 1, dg)
    assert np.abs(y - g['y']).max() < 5e-6
    outs = mdcn_c.backward(g['x'], g['offset'], g['mask'], g['weight'], g['gy'], 1, 1, 1, 1, dg)
    for name, a in zip(['gx', 'goffset', 'gmask', 'gweight', 'gbias'], outs):
        assert np.abs(a - g[name]).max() <= 1e-5 * max(1.0, np.abs(g[name]).max()), name


def test_torch_oracle_matches_golden_dcn_and_grads():
    g = gold('dcn_small.npz')
    t = {k: torch.from_numpy(g[k]) for k in ('x', 'offset', 'mask', 'weight', 'bias', 'gy')}
    leaves = [t[k].clone().requires_grad_(True) for k in ('x', 'offset', 'mask', 'weight', 'bias')]
    y = mdcn_torch(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], 1, 1, 1, 1, int(g['dg']))
    assert rel(y, torch.from_numpy(g['y'])) < 1e-6
    grads = torch.autograd.grad(y, leaves, t['gy'])
    for name, a in zip(['gx', 'goffset', 'gmask', 'gweight', 'gbias'], grads):
        assert rel(a, torch.from_numpy(g[name])) < 1e-5, name


@pytest.mark.parametrize('stride,pad,dil', [(1, 1, 1), (2, 1, 1), (1, 2, 2), (1, 0, 1)])
def test_torch_oracle_vs_torchvision_geometry(stride, pad, dil):
    torchvision = pytest.importorskip('torchvision')
    g = torch.Generator().manual_seed(3)
    B, C, H, W, Co, dg = 1, 8, 10, 7, 6, 2
    Ho = (H + 2 * pad - (dil * 2 + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * 2 + 1)) // stride + 1
    x = torch.randn(B, C, H, W, generator=g)
    off = torch.randn(B, dg * 18, Ho, Wo, generator=g) * 4
    m = torch.rand(B, dg * 9, Ho, Wo, generator=g)
    w = torch.randn(Co, C, 3, 3, generator=g)
    ref = torchvision.ops.deform_conv2d(x, off, w, None, stride=stride, padding=pad, dilation=dil, mask=m)
    assert rel(mdcn_torch(x, off, m, w, None, stride, pad, dil, 1, dg), ref) < 1e-5
    yc = mdcn_c.forward(x.numpy(), off.numpy(), m.numpy(), w.numpy(), None, stride, pad, dil, 1, dg)
    assert rel(torch.from_numpy(yc), ref) < 1e-5


def test_empty_and_border_taps():
    # all offsets far outside: output must be exactly the bias (kernel.cu:617)
    x = torch.randn(1, 4, 5, 5)
    off = torch.full((1, 18, 5, 5), 100.0)
    m = torch.ones(1, 9, 5, 5)
    w = torch.randn(3, 4, 3, 3)
    b = torch.randn(3)
    y = mdcn_torch(x, off, m, w, b, 1, 1, 1, 1, 1)
    assert torch.equal(y, b.view(1, 3, 1, 1).expand_as(y))
    yc = mdcn_c.forward(x.numpy(), off.numpy(), m.numpy(), w.numpy(), b.numpy(), 1, 1, 1, 1, 1)
    assert np.array_equal(yc, y.numpy())


def test_edvr_oracle_matches_reference_golden():
    g = gold('edvr_m_32.npz')
    sd = P.make_params(P.edvr_param_shapes(), seed=int(g['seed']))
    with torch.no_grad():
        out, inter = O.edvr_forward(sd, torch.from_numpy(g['x']), return_intermediates=True)
    assert rel(out, torch.from_numpy(g['out'])) < 1e-6
    assert rel(inter['aligned'][:, 2], torch.from_numpy(g['aligned_center'])) < 1e-6
    assert rel(inter['tsa'], torch.from_numpy(g['tsa'])) < 1e-6


def test_sfdn_oracle_matches_reference_golden():
    """SFDN restatement (LRimg_estimator.py:38-67) vs the unmodified reference module (oracle/make_golden_sfdn.py), forward
    and the input gradient (the spatial mean stays in the graph)."""
    g = gold('sfdn_32x48.npz')
    sd = P.make_params(P.sfdn_param_shapes(64), seed=int(g['seed']))
    x = torch.from_numpy(g['x']).requires_grad_(True)
    out = O.sfdn_forward(sd, x)
    assert rel(out, torch.from_numpy(g['out'])) < 1e-6
    (out * torch.from_numpy(g['probe'])).sum().backward()
    assert rel(x.grad, torch.from_numpy(g['gx'])) < 1e-5


def test_mfdn_oracle_matches_reference_golden():
    g = gold('mfdn_32x48.npz')
    sd = P.make_params(P.mfdn_param_shapes(), seed=int(g['seed']))
    with torch.no_grad():
        slr = O.mfdn_forward(sd, torch.from_numpy(g['lr']).transpose(1, 2))
    assert rel(slr, torch.from_numpy(g['slr'])) < 1e-6


@pytest.mark.parametrize('tag,optimizer,crit', [('sgd2_l2', 'SGD', 'l2'), ('adam1_cb', 'Adam', 'cb')])
def test_adapt_oracle_matches_reference_golden(tag, optimizer, crit):
    g = gold('adapt_%s.npz' % tag)
    sdG = P.make_params(P.edvr_param_shapes(), seed=int(g['seed_G']))
    sdE = P.make_params(P.mfdn_param_shapes(), seed=int(g['seed_E']))
    sdF = P.make_params(P.mfdn_param_shapes(), seed=int(g['seed_E_fixed']))
    out, losses, pG, pE = O.adapt_and_infer(sdG, sdE, sdF, torch.from_numpy(g['lr']), steps=int(g['steps']),
                                            lr_alpha=float(g['lr_alpha']), optimizer=optimizer, criterion=crit,
                                            return_losses=True)
    assert np.allclose(losses, g['losses'], rtol=1e-5)
    assert rel(out, torch.from_numpy(g['out'])) < 1e-5
    assert rel(pG['conv_first.weight'] - sdG['conv_first.weight'], torch.from_numpy(g['d_conv_first'])) < 1e-3
    assert rel(pE['conv6.weight'] - sdE['conv6.weight'], torch.from_numpy(g['d_conv6'])) < 1e-3


def test_param_inventory_counts():
    # EDVR-M 3 300 131 params / 144 tensors, MFDN 4x 452 291 / 14, EDVR-L 20 633 827 / 264
    n = lambda d: sum(int(np.prod(s)) for s in d.values())
    assert (n(P.edvr_param_shapes()), len(P.edvr_param_shapes())) == (3300131, 144)
    assert (n(P.mfdn_param_shapes()), len(P.mfdn_param_shapes())) == (452291, 14)
    L = P.edvr_param_shapes(nf=128, back_RBs=40)
    assert (n(L), len(L)) == (20633827, 264)


@pytest.mark.parametrize('tag,kw', [('predeblur', dict(predeblur=True, HR_in=False, w_TSA=True)),
                                    ('hrin_notsa', dict(predeblur=False, HR_in=True, w_TSA=False)),
                                    ('predeblur_hrin', dict(predeblur=True, HR_in=True, w_TSA=True))])
def test_edvr_variant_oracle_matches_reference_golden(tag, kw):
    """predeblur / HR_in / w_TSA=False (EDVR_arch.py:13-57,208-239) through the unmodified reference module."""
    g = gold('edvr_variants.npz')
    cfg = dict(nf=64, nframes=5, groups=8, front_RBs=1, back_RBs=1, scale=4)
    sd = P.make_params(P.edvr_param_shapes(**cfg, **kw), seed=int(g[tag + '_seed']))
    y = O.edvr_forward(sd, torch.from_numpy(g[tag + '_x']), front_RBs=1, back_RBs=1, predeblur_=kw['predeblur'],
                       HR_in=kw['HR_in'], w_TSA=kw['w_TSA'])
    assert torch.equal(y, torch.from_numpy(g[tag + '_out']))


def _meta_loop_fixture():
    g = gold('meta_loop.npz')
    nf, nframes, groups, front, back, nf_e, scale, batch, iters, inner = [int(v) for v in g['cfg']]
    shapes_G = P.edvr_param_shapes(nf=nf, nframes=nframes, groups=groups, front_RBs=front, back_RBs=back, scale=scale)
    shapes_E = P.mfdn_param_shapes(nf=nf_e, scale=scale)
    assert list(g['keys_G']) == list(shapes_G) and list(g['keys_E']) == list(shapes_E)

    def unflat(vec, shapes):
        out, o = {}, 0
        for k, s in shapes.items():
            n = int(np.prod(s))
            out[k] = torch.from_numpy(vec[o:o + n].reshape(s).copy())
            o += n
        assert o == vec.size
        return out

    cfg = dict(nf=nf, nframes=nframes, groups=groups, front_RBs=front, back_RBs=back)
    return g, shapes_G, shapes_E, unflat, cfg, scale, batch, iters, inner


def test_meta_oracle_vs_reference_training_loop():
    """The loop structure of the outer step, pinned: tests/golden/meta_loop.npz holds the weights the UNMODIFIED
    train_dynavsr.py main() leaves after each of two outer iterations (oracle/make_golden_meta.py ran it on CPU);
    meta_oracle's ``as_written`` mode must land on the same weights from the same start, tasks and hyper-parameters."""
    from oracle.meta_oracle import meta_outer_step
    g, shapes_G, shapes_E, unflat, cfg, scale, batch, iters, inner = _meta_loop_fixture()
    sdG, sdE = unflat(g['G0'], shapes_G), unflat(g['E0'], shapes_E)
    # the start weights are the seeded ones (make_golden_meta.py re-seeds the reference modules with seeds 21 / 22)
    for k, v in P.make_params(shapes_G, 21).items():
        assert torch.equal(v, sdG[k])
    state = None
    for it in range(iters):
        tasks = [{'LQs': torch.from_numpy(g['it%d_LQs' % it][b:b + 1]), 'GT': torch.from_numpy(g['it%d_GT' % it][b:b + 1]),
                  'SuperLQs': torch.from_numpy(g['it%d_SuperLQs' % it][b:b + 1])} for b in range(batch)]
        sdG, sdE, info = meta_outer_step(sdG, sdE, tasks, inner_steps=inner, lr_alpha=1e-3, lr_alpha_est=2e-3,
                                         inner_optimizer='Adam', inner_betas=(0.9, 0.99), criterion='cb', est_loss='l1',
                                         outer='Adam', lr_outer=1e-3, outer_betas=(0.9, 0.99), outer_state=state,
                                         mode='as_written', scale=scale, edvr_cfg=cfg)
        state = info['outer_state']
        # 'Train loss' scalar of train_dynavsr.py:432 = sum of loss_q / batch
        assert abs(sum(info['loss_q']) / batch - float(g['train_loss'][it])) < 1e-5
        refG, refE = unflat(g['G%d' % (it + 1)], shapes_G), unflat(g['E%d' % (it + 1)], shapes_E)
        prevG, prevE = unflat(g['G%d' % it], shapes_G), unflat(g['E%d' % it], shapes_E)
        # compare the UPDATES (Adam's first steps are ~lr * sign(g): the weights themselves would agree trivially)
        dG = torch.cat([(sdG[k] - prevG[k]).reshape(-1) for k in shapes_G])
        dE = torch.cat([(sdE[k] - prevE[k]).reshape(-1) for k in shapes_E])
        rG = torch.cat([(refG[k] - prevG[k]).reshape(-1) for k in shapes_G])
        rE = torch.cat([(refE[k] - prevE[k]).reshape(-1) for k in shapes_E])
        assert rel(dG, rG) < 2e-3, (it, rel(dG, rG))
        assert rel(dE, rE) < 2e-3, (it, rel(dE, rE))
        sdG, sdE = refG, refE            # continue from the reference's own weights (no drift accumulation)


def test_meta_loop_golden_is_not_first_order_maml():
    """Documents WHY there are two modes: on the same inputs the ``fomaml`` mode gives a different outer update than the
    reference's training loop -- the reference loop does not adapt its working copy."""
    from oracle.meta_oracle import meta_outer_step
    g, shapes_G, shapes_E, unflat, cfg, scale, batch, iters, inner = _meta_loop_fixture()
    sdG, sdE = unflat(g['G0'], shapes_G), unflat(g['E0'], shapes_E)
    tasks = [{'LQs': torch.from_numpy(g['it0_LQs'][b:b + 1]), 'GT': torch.from_numpy(g['it0_GT'][b:b + 1]),
              'SuperLQs': torch.from_numpy(g['it0_SuperLQs'][b:b + 1])} for b in range(batch)]
    _, _, info = meta_outer_step(sdG, sdE, tasks, inner_steps=inner, lr_alpha=1e-3, lr_alpha_est=2e-3, criterion='cb',
                                 outer='Adam', lr_outer=1e-3, mode='fomaml', scale=scale, edvr_cfg=cfg)
    assert abs(sum(info['loss_q']) / batch - float(g['train_loss'][0])) > 1e-4


def test_meta_oracle_outer_gradient_vs_reference_training_loop():
    """Outer SGD with lr 1 makes the reference's update equal to its accumulated outer gradient: compare it with
    meta_oracle's ``as_written`` gradient tensor by tensor (the Adam pin above is insensitive to gradient magnitudes)."""
    from oracle.meta_oracle import meta_outer_step
    g, shapes_G, shapes_E, unflat, cfg, scale, batch, iters, inner = _meta_loop_fixture()
    s = gold('meta_loop_sgd.npz')
    sdG, sdE = unflat(g['G0'], shapes_G), unflat(g['E0'], shapes_E)
    refG, refE = unflat(s['G1'], shapes_G), unflat(s['E1'], shapes_E)
    tasks = [{'LQs': torch.from_numpy(g['it0_LQs'][b:b + 1]), 'GT': torch.from_numpy(g['it0_GT'][b:b + 1]),
              'SuperLQs': torch.from_numpy(g['it0_SuperLQs'][b:b + 1])} for b in range(batch)]
    nG, nE, info = meta_outer_step(sdG, sdE, tasks, inner_steps=inner, lr_alpha=1e-3, lr_alpha_est=2e-3,
                                   inner_optimizer='Adam', inner_betas=(0.9, 0.99), criterion='cb', est_loss='l1',
                                   outer='SGD', lr_outer=float(s['lr_G']), mode='as_written', scale=scale, edvr_cfg=cfg)
    assert abs(sum(info['loss_q']) / batch - float(s['train_loss'][0])) < 1e-5
    worst = 0.0
    for sd, ref, grads in ((sdG, refG, info['gG']), (sdE, refE, info['gE'])):
        for k in sd:
            g_ref = (sd[k] - ref[k]) / float(s['lr_G'])
            scale_k = float(g_ref.abs().max())
            assert scale_k > 0, k                      # every tensor receives gradient in the reference loop
            # absolute error relative to the tensor's largest gradient entry; the reference's w - lr*g is rounded to fp32
            err = float((grads[k] - g_ref).abs().max()) / max(scale_k, 1e-4)
            worst = max(worst, err)
            assert err < 1e-4, (k, err, scale_k)
    assert rel(torch.cat([v.reshape(-1) for v in nG.values()]), torch.cat([v.reshape(-1) for v in refG.values()])) < 1e-6


def _driver_weights(g):
    gain = float(g['head_gain'])

    def edvr(seed):
        sd = P.make_params(P.edvr_param_shapes(), seed=seed)
        sd['conv_last.weight'] *= gain             # make_golden_driver.tame_head: outputs inside [0, 1]
        sd['conv_last.bias'] *= gain
        return sd

    return (edvr(int(g['seed_G'])), P.make_params(P.mfdn_param_shapes(), seed=int(g['seed_E'])),
            P.make_params(P.mfdn_param_shapes(), seed=int(g['seed_E_fixed'])), edvr(int(g['seed_baseline_G'])))


@pytest.mark.parametrize('tag', ['adam1_cb', 'sgd2_l2'])
def test_adapt_oracle_vs_reference_test_driver(tag):
    """tests/golden/driver_<tag>.npz is what the UNMODIFIED test_dynavsr.py main() produced for one clip (baseline
    inference, deepcopy, K inner steps, final inference, tensor2img, PSNR -- oracle/make_golden_driver.py): the oracle's
    adapt_and_infer must give the same frame, the same PNG bytes and the same PSNR."""
    from util import psnr_uint8
    g = gold('driver_%s.npz' % tag)
    sdG, sdE, sdF, sdB = _driver_weights(g)
    lq, gt = torch.from_numpy(g['lq']), torch.from_numpy(g['gt'])
    out, losses, pG, pE = O.adapt_and_infer(sdG, sdE, sdF, lq, steps=int(g['steps']), lr_alpha=float(g['lr_alpha']),
                                            optimizer=str(g['optimizer']), criterion=str(g['criterion']),
                                            return_losses=True)
    out = out[0].clamp(0, 1)                      # the driver's tensor2img clamps in place on CPU (utils/util.py:118)
    assert rel(out, torch.from_numpy(g['out'])) < 1e-6
    image = (out * 255.0).round().permute(1, 2, 0).numpy().astype(np.uint8)
    assert int((image != g['image']).sum()) <= 16            # rounding ties only (measured: 0-6 of 73 728 bytes)
    assert abs(psnr_uint8(out, gt) - float(g['psnr_adapted'])) < 1e-3
    assert rel(pG['conv_first.weight'] - sdG['conv_first.weight'], torch.from_numpy(g['d_conv_first'])) < 1e-4
    assert rel(pE['conv6.weight'] - sdE['conv6.weight'], torch.from_numpy(g['d_conv6'])) < 1e-4
    with torch.no_grad():
        base = O.edvr_forward(sdB, lq)[0]
    assert abs(psnr_uint8(base, gt) - float(g['psnr_baseline'])) < 1e-3
    assert float(g['psnr_adapted']) > float(g['psnr_baseline'])


@pytest.mark.parametrize('tag', ['sgd2_l2_patch', 'adam1_cb_real'])
def test_adapt_oracle_optional_branches_vs_reference_test_driver(tag):
    """The ``maml.use_patch`` (test_dynavsr.py:118-145,255-260) and ``train.use_real`` (:218-221,243-244) branches of the inner
    loop, pinned the same way: the unmodified driver ran them (crop positions = what its ``random.randrange`` calls returned,
    logged by the harness; 'SuperLQs' supplied by the loader) and the oracle reproduces frame, PSNR and parameter deltas --
    with use_real MFDN must not move at all."""
    from util import psnr_uint8
    g = gold('driver_%s.npz' % tag)
    sdG, sdE, sdF, _ = _driver_weights(g)
    lq, gt = torch.from_numpy(g['lq']), torch.from_numpy(g['gt'])
    extra = {}
    if bool(g['use_real']):
        extra['slr_given'] = torch.from_numpy(g['slq'])
    if bool(g['use_patch']):
        n, c = int(g['num_patch']), g['crops'].tolist()
        pos = [(c[2 * i], c[2 * i + 1]) for i in range(len(c) // 2)]
        assert len(pos) == n * int(g['steps'])
        extra.update(patches=[pos[k * n:(k + 1) * n] for k in range(int(g['steps']))], patch_size=int(g['patch_size']))
    out, losses, pG, pE = O.adapt_and_infer(sdG, sdE, sdF, lq, steps=int(g['steps']), lr_alpha=float(g['lr_alpha']),
                                            optimizer=str(g['optimizer']), criterion=str(g['criterion']), return_losses=True, **extra)
    out = out[0].clamp(0, 1)
    assert rel(out, torch.from_numpy(g['out'])) < 1e-6
    assert abs(psnr_uint8(out, gt) - float(g['psnr_adapted'])) < 1e-3
    assert rel(pG['conv_first.weight'] - sdG['conv_first.weight'], torch.from_numpy(g['d_conv_first'])) < 1e-4
    if bool(g['use_real']):
        assert float(np.abs(g['d_conv6']).max()) == 0.0 and torch.equal(pE['conv6.weight'], sdE['conv6.weight'])
    else:
        assert rel(pE['conv6.weight'] - sdE['conv6.weight'], torch.from_numpy(g['d_conv6'])) < 1e-4


def test_precision_study_operand_emulation():
    """The operand roundings oracle/precision_study.py emulates (its conclusions feed DESIGN.md section 7): TF32 = round to
    nearest even on 10 explicit mantissa bits, bf16 split = hi + lo reconstructs 16 mantissa bits; the three-product scheme
    of the kernels equals the exact product up to the dropped lo.lo term."""
    from oracle import precision_study as S
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4096, generator=g) * torch.logspace(-6, 6, 4096)
    t = S._tf32(x)
    assert int((t.view(torch.int32) & 0x1FFF).abs().max()) == 0                       # 13 low mantissa bits cleared
    assert float(((t - x).abs() / x.abs()).max()) <= 2.0 ** -11 * (1 + 1e-6)
    tie = torch.tensor([1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11])                      # exactly half way between neighbours
    assert S._tf32(tie).tolist() == [1.0, 1.0 + 2.0 ** -9]                            # ties go to the even mantissa
    hi, lo = S._bf16_split(x)
    assert float(((hi + lo - x).abs() / x.abs()).max()) <= 2.0 ** -16
    a, w = torch.randn(2, 8, 9, 9, generator=g), torch.randn(4, 8, 3, 3, generator=g) * 0.1
    exact = torch.nn.functional.conv2d(a.double(), w.double(), None, padding=1)
    err = lambda scheme: rel(S.emulated_conv(scheme, a, w, None, 1, 1), exact)
    assert err('x3') < 3e-5 < err('tf32') < 2e-3 and err('tf32') < err('x2w') < 2 * err('x1') and err('x1') < 1e-2
    assert S.group_of('recon_trunk.3.conv1') == 'trunk' and S.group_of('pcd_align.L1_dcnpack.conv_offset_mask') == 'dcn_offset_mask'


@pytest.mark.parametrize('case', ['plain_adam_cb', 'small_offset_sgd_l2', 'ft_tsa_only3_sgd_l1', 'ft_tsa_and_small_offset',
                                  'weight_decay_sgd'])
def test_training_iterations_vs_reference_wrapper_golden(case):
    """The host logic of the training path put together on CPU -- ``param_group_spec`` + the chained schedules +
    ``update_learning_rate`` ordering + the ft_tsa_only freeze -- around the oracle's forward / loss and torch.optim, against
    what the UNMODIFIED reference VideoBaseModel did (tests/golden/wrapper_train.*): losses, learning rates, probe deltas."""
    import json
    import os
    from util import GOLD
    from dynavsr_b200.models.Video_base_model import param_group_spec
    from dynavsr_b200.models.lr_scheduler import MultiStepLR_Restart
    from dynavsr_b200.options import dict_to_nonedict
    g = json.load(open(os.path.join(GOLD, 'wrapper_train.json')))
    arr = np.load(os.path.join(GOLD, 'wrapper_train.npz'))
    c = g['cases'][case]
    t = dict_to_nonedict(dict(c['train'], lr_steps=[2], lr_gamma=0.5))
    sd0 = P.make_params(P.edvr_param_shapes(scale=4, **g['net']), seed=int(g['seed']))
    w = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    groups = [{'params': [w[k] for k in names], 'lr': lr} for lr, names in param_group_spec(list(w), t)]
    wd = t['weight_decay_G'] or 0
    opt = torch.optim.SGD(groups, lr=t['lr_G'], weight_decay=wd) if t['optim'] == 'SGD' else \
        torch.optim.Adam(groups, lr=t['lr_G'], weight_decay=wd, betas=(0.9, 0.99))
    sch = MultiStepLR_Restart(opt, t['lr_steps'], gamma=t['lr_gamma'])
    x, gt = torch.from_numpy(arr['LQs']), torch.from_numpy(arr['GT'])
    losses, lrs = [], []
    for step in range(1, int(g['steps']) + 1):
        sch.step()                                                    # update_learning_rate (no warm-up)
        if t['ft_tsa_only'] and step < t['ft_tsa_only']:
            opt.param_groups[0]['lr'] = 0                             # set_params_lr_zero
        opt.zero_grad()
        loss = O.pixel_loss(t['pixel_criterion'], O.edvr_forward(w, x, **g['net']), gt)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
        lrs.append([grp['lr'] for grp in opt.param_groups])
    assert lrs == [pytest.approx(v) for v in c['lrs_after_step']]
    assert losses == pytest.approx(c['losses'], rel=1e-5)
    for k in g['probes']:
        want = torch.from_numpy(arr['%s/%s' % (case, k)])
        got = w[k].detach() - sd0[k]
        if float(want.abs().max()) == 0.0:
            assert float(got.abs().max()) == 0.0, k
        else:
            assert rel(got, want) < 1e-3, k
