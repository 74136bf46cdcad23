"""Evaluation loop on the GPU (dynavsr_b200/driver.py): the 8-bit image + squared-error kernel against torch, and the whole
loop against the table the reference's own test driver wrote for the same clip (tests/golden/driver_*.npz)."""
import os

import numpy as np
import pytest
import torch

from util import gold

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cuda():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')


@pytest.mark.parametrize('shape', [(1, 7, 9, 3), (33, 47, 3), (2, 16, 16, 1), (5, 5, 4)])
@pytest.mark.parametrize('bgr', [False, True])
def test_frame_to_u8_bit_exact(cuda, shape, bgr):
    from dynavsr_b200 import ops
    g = torch.Generator().manual_seed(sum(shape) + bgr)
    x = torch.rand(shape, generator=g) * 1.4 - 0.2                      # below 0 and above 1 too
    flat = x.view(-1)
    k = torch.arange(0, min(64, flat.numel()))
    flat[k] = (k.float() + 0.5) / 255.0                                 # exact .5 ties: round half to even
    flat[-1], flat[-2] = float('inf'), -float('inf')
    ref_img = (torch.rand(shape, generator=g) * 255).to(torch.uint8)
    want = (x.clamp(0, 1) * 255.0).round().to(torch.uint8)
    if bgr:
        want = want.flip(-1)
    sse = torch.zeros(1, dtype=torch.int64, device='cuda')
    got = ops.frame_to_u8(x.cuda(), ref=ref_img.cuda(), sse=sse, bgr=bgr)
    assert got.dtype == torch.uint8 and got.shape == x.shape
    assert torch.equal(got.cpu(), want)
    want_sse = int(((want.long() - ref_img.long()) ** 2).sum())
    assert int(sse.item()) == want_sse
    ops.frame_to_u8(x.cuda(), ref=ref_img.cuda(), sse=sse, bgr=bgr)    # accumulates
    assert int(sse.item()) == 2 * want_sse
    out = torch.zeros(shape, dtype=torch.uint8, device='cuda')
    assert ops.frame_to_u8(x.cuda(), out=out, bgr=bgr) is out and torch.equal(out.cpu(), want)


def test_frame_to_u8_argument_errors(cuda):
    from dynavsr_b200 import ops
    x = torch.rand(4, 4, 3, device='cuda')
    with pytest.raises(NotImplementedError):
        ops.frame_to_u8(x.cpu())
    with pytest.raises(RuntimeError):
        ops.frame_to_u8(x, out=torch.zeros(4, 4, 3, device='cuda'))                     # float image buffer
    with pytest.raises(RuntimeError):
        ops.frame_to_u8(x, sse=torch.zeros(1, dtype=torch.int64, device='cuda'))         # accumulator without a reference
    with pytest.raises(RuntimeError):
        ops.frame_to_u8(x, ref=torch.zeros(4, 4, 3, dtype=torch.uint8, device='cuda'), sse=torch.zeros(1, device='cuda'))


def _nets(g):
    from oracle import params as P
    from dynavsr_b200.models.archs import EDVR_arch, LRimg_estimator
    gain = float(g['head_gain'])

    def edvr(seed):
        sd = P.make_params(P.edvr_param_shapes(), seed=seed)
        sd['conv_last.weight'] *= gain
        sd['conv_last.bias'] *= gain
        net = EDVR_arch.EDVR(nf=64, nframes=5, groups=8, front_RBs=5, back_RBs=10, scale=4)
        net.load_state_dict(sd, strict=True)
        return net.cuda()

    def mfdn(seed):
        net = LRimg_estimator.DirectKernelEstimatorVideo(nf=64, in_nc=3, scale=4)
        net.load_state_dict(P.make_params(P.mfdn_param_shapes(), seed=seed), strict=True)
        return net.cuda()

    return edvr(int(g['seed_G'])), mfdn(int(g['seed_E'])), mfdn(int(g['seed_E_fixed'])), edvr(int(g['seed_baseline_G']))


def _loader(g, n):
    lq = torch.from_numpy(g['lq'])
    gt = torch.from_numpy(g['gt'])[None, None].expand(1, lq.shape[1], -1, -1, -1)
    return [{'LQs': lq, 'GT': gt, 'folder': ['clip'], 'idx': ['%d/%d' % (i, n)]} for i in range(n)]


@pytest.mark.parametrize('mode', ['engine-fp32', 'pool-tensor-core'])
def test_evaluate_reproduces_reference_driver_table(cuda, mode, tmp_path):
    """Same clip, same weights, same settings as the run of the unmodified test_dynavsr.py main() recorded in
    tests/golden/driver_sgd2_l2.npz: every row of the table must match the reference's psnr_update.csv (PSNR within 0.01 dB
    -- the north-star criterion -- SSIM within 1e-3) and the PNG on disk must be the reference's image within one level."""
    import cv2
    from dynavsr_b200 import adapt, driver, ops
    g = gold('driver_sgd2_l2.npz')
    netG, netE, netF, base = _nets(g)
    kw = dict(steps=int(g['steps']), lr_alpha=float(g['lr_alpha']), optimizer=str(g['optimizer']), criterion=str(g['criterion']),
              slr_weight=10.0)
    tc = mode == 'pool-tensor-core'
    ops.set_conv_backend(tc)
    try:
        if tc:
            engine = adapt.AdaptationPool(netG, netE, netF, pipelines=3, use_graphs=True, **kw)
        else:
            engine = adapt.InnerLoopAdapter(netG, netE, netF, use_graphs=False, **kw)
        n = 7 if tc else 2
        rows = driver.evaluate(engine, _loader(g, n), baseline_netG=base, save_dir=str(tmp_path), ring_slots=2)
        assert list(rows) == ['clip/%08d' % i for i in range(n)]
        for name, (p0, p1, s0, s1) in rows.items():
            assert abs(p0 - float(g['psnr_baseline'])) < 0.01, (name, p0)
            assert abs(p1 - float(g['psnr_adapted'])) < 0.01, (name, p1)
            assert abs(s0 - float(g['ssim_baseline'])) < 1e-3 and abs(s1 - float(g['ssim_adapted'])) < 1e-3, (name, s0, s1)
        for i in range(n):
            img = cv2.imread(os.path.join(str(tmp_path), 'clip', 'DynaVSR', '%08d.png' % i))[..., ::-1]
            assert int(np.abs(img.astype(np.int32) - g['image'].astype(np.int32)).max()) <= 1
        # sharded over two ranks: the union of the two tables is the whole table, no frame twice
        r0 = driver.evaluate(engine, _loader(g, n), with_GT=True, compute_ssim=False, rank=0, world_size=2)
        r1 = driver.evaluate(engine, _loader(g, n), with_GT=True, compute_ssim=False, rank=1, world_size=2)
        assert sorted(list(r0) + list(r1)) == list(rows) and not set(r0) & set(r1)
        assert all(np.isnan(v[0]) and np.isnan(v[3]) and abs(v[1] - float(g['psnr_adapted'])) < 0.01 for v in r0.values())
        # no ground truth (the 'demo' mode of the reference, with_GT = False): images only
        demo = driver.evaluate(engine, [{k: v for k, v in d.items() if k != 'GT'} for d in _loader(g, 2)], with_GT=False,
                               save_dir=str(tmp_path / 'demo'))
        assert all(np.isnan(v).all() for v in demo.values()) and os.path.exists(str(tmp_path / 'demo' / 'clip' / 'DynaVSR' / '00000001.png'))
    finally:
        ops.set_conv_backend(False)       # (the pool's launch policy lives in its engines' scopes: nothing global to undo)


def test_resident_clips_feed_the_evaluation_loop(cuda):
    """clips.ResidentClips (device-resident clip, GPU degradation, window gather) -> driver.evaluate: every row must equal
    what the plain per-window API gives for that window (adapt_and_infer + host-side tensor2img / PSNR)."""
    import torch.nn.functional as F
    from util import psnr_uint8
    from dynavsr_b200 import adapt, clips, driver
    from dynavsr_b200.degradation import Degradation
    g = gold('driver_sgd2_l2.npz')
    netG, netE, netF, _ = _nets(g)
    gen = torch.Generator().manual_seed(5)
    hr = F.interpolate(torch.rand(1, 3, 10, 14, generator=gen), size=(128 + 8, 192 + 8), mode='bicubic', align_corners=False).clamp(0, 1)
    hr = torch.stack([hr[0, :, t:t + 128, t:t + 192] for t in range(6)])          # a 6-frame panning clip [6, 3, 128, 192]
    store = clips.ResidentClips(n_frames=5, padding='new_info', scale=4, device='cuda')
    store.add_degraded('pan', hr, Degradation(21, 4, sigma=[1.6, 1.6]))
    assert len(store) == 6
    it = store.item(0)
    assert it['LQs'].is_cuda and it['LQs'].shape == (5, 3, 32, 48) and it['GT'].shape == (5, 3, 128, 192)
    assert it['SuperLQs'].shape == (5, 3, 8, 12) and store.window_indices(0) == [4, 3, 0, 1, 2]
    eng = adapt.InnerLoopAdapter(netG, netE, netF, steps=2, lr_alpha=1e-4, optimizer='SGD', criterion='l2', slr_weight=10.0,
                                 use_graphs=False)
    rows = driver.evaluate(eng, store, compute_ssim=False)
    assert list(rows) == ['pan/%08d' % i for i in range(6)]
    for i in (0, 3, 5):
        d = store[i]
        out = eng.adapt_and_infer(d['LQs'])[0]
        want = psnr_uint8(out, d['GT'][0, 2])
        assert abs(rows['pan/%08d' % i][1] - want) < 1e-3, (i, rows['pan/%08d' % i][1], want)
    assert len({round(v[1], 3) for v in rows.values()}) > 1                       # the windows really differ
