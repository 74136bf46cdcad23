"""GPU parity tests of every kernel family against a plain PyTorch reference computed on CPU in
float64 (tolerances are written at each assert; the north-star bar is 1e-3 relative fp32)."""
import pytest
import torch
import torch.nn.functional as F

from util import gold, maxabs, nchw, nhwc, rel

pytestmark = pytest.mark.gpu
TOL = 2e-5          # fp32 kernels vs fp64 truth, relative L2
GTOL = 1e-4         # gradients (atomics / longer reductions)


@pytest.fixture(scope='module')
def ops():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dynavsr_b200 import ops as _ops
    return _ops


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64) * scale


def _dev(t):
    return t.float().cuda()


def _act_ref(v, act, sig_split=0):
    if act == 1:
        return F.relu(v)
    if act == 2:
        return F.leaky_relu(v, 0.1)
    if act == 3:
        return torch.cat([v[:, :sig_split], torch.sigmoid(v[:, sig_split:])], 1)
    return v


CONV_CASES = [
    # name, N, H, W, segC, Co, k, stride, pad, act, res, shuffle
    ('3x3_64', 2, 12, 20, [64], 64, 3, 1, 1, 2, False, 0),
    ('3x3_relu_res', 1, 9, 13, [64], 64, 3, 1, 1, 0, True, 0),
    ('3x3_relu', 1, 9, 13, [64], 64, 3, 1, 1, 1, False, 0),
    ('3x3_lrelu_then_res', 2, 9, 13, [64], 64, 3, 1, 1, 2, True, 0),
    ('cat2', 2, 8, 8, [64, 64], 64, 3, 1, 1, 2, False, 0),
    ('stride2', 2, 12, 16, [64], 64, 3, 2, 1, 2, False, 0),
    ('cin3', 3, 10, 14, [3], 64, 3, 1, 1, 2, False, 0),
    ('cout3', 1, 10, 14, [64], 3, 3, 1, 1, 0, True, 0),
    ('offmask216', 1, 7, 9, [64], 216, 3, 1, 1, 3, False, 0),
    ('1x1_320', 1, 8, 12, [320], 64, 1, 1, 0, 2, False, 0),
    ('1x1_cout3', 2, 6, 6, [64], 3, 1, 1, 0, 0, False, 0),
    ('4x4s2', 2, 10, 14, [64], 128, 4, 2, 0, 2, False, 0),
    ('4x4s2_128', 1, 10, 14, [128], 64, 4, 2, 0, 2, False, 0),
    ('shuffle', 1, 6, 10, [64], 256, 3, 1, 1, 2, False, 2),
    ('big_M_tail', 1, 23, 37, [16], 24, 3, 1, 1, 2, False, 0),
]


@pytest.mark.parametrize('case', CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_forward_backward(ops, case):
    name, N, H, W, segC, Co, k, stride, pad, act, use_res, shuffle = case
    xs = [_rand(N, c, H, W, seed=i + 1) for i, c in enumerate(segC)]
    w = _rand(Co, sum(segC), k, k, seed=10, scale=0.1)
    b = _rand(Co, seed=11, scale=0.1)
    sig_split = 144 if act == 3 else 0
    # reference (float64 CPU)
    xr = [t.clone().requires_grad_(True) for t in xs]
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = _act_ref(F.conv2d(torch.cat(xr, 1), wr, br, stride=stride, padding=pad), act, sig_split)
    if shuffle:
        y = F.pixel_shuffle(y, 2)
    res = _rand(*y.shape, seed=12) if use_res else None
    resr = res.clone().requires_grad_(True) if use_res else None
    if use_res:
        y = y + resr
    gy = _rand(*y.shape, seed=13)
    grads = torch.autograd.grad(y, xr + [wr, br] + ([resr] if use_res else []), gy)
    # device
    xd = [nhwc(_dev(t)).requires_grad_(True) for t in xs]
    wd, bd = _dev(w).requires_grad_(True), _dev(b).requires_grad_(True)
    resd = nhwc(_dev(res)).requires_grad_(True) if use_res else None
    yd = ops.conv(xd, wd, bd, stride=stride, pad=pad, act=act, sig_split=sig_split, res=resd, shuffle=shuffle)
    assert rel(nchw(yd), y) < TOL, 'forward'
    gd = torch.autograd.grad(yd, xd + [wd, bd] + ([resd] if use_res else []), nhwc(_dev(gy)))
    for i in range(len(xs)):
        assert rel(nchw(gd[i]), grads[i]) < GTOL, 'grad input %d' % i
    assert rel(gd[len(xs)], grads[len(xs)]) < GTOL, 'grad weight'
    assert rel(gd[len(xs) + 1], grads[len(xs) + 1]) < GTOL, 'grad bias'
    if use_res:
        assert rel(nchw(gd[-1]), grads[-1]) < GTOL, 'grad residual'


def test_conv_broadcast_reference_segment(ops):
    """PCD: cat([nbr_f, ref]) for all N frames of each clip with ONE reference feature per clip."""
    B, N, H, W, C = 2, 5, 6, 8, 64
    nbr, ref = _rand(B * N, C, H, W, seed=1), _rand(B, C, H, W, seed=2)
    w, b = _rand(C, 2 * C, 3, 3, seed=3, scale=0.05), _rand(C, seed=4, scale=0.1)
    nr, rr, wr, br = (t.clone().requires_grad_(True) for t in (nbr, ref, w, b))
    y = F.leaky_relu(F.conv2d(torch.cat([nr, rr.repeat_interleave(N, 0)], 1), wr, br, padding=1), 0.1)
    gy = _rand(*y.shape, seed=5)
    gr = torch.autograd.grad(y, [nr, rr, wr, br], gy)
    nd = nhwc(_dev(nbr)).requires_grad_(True)
    # the reference features live inside a [B, N, ...] tensor: take the strided centre view, as EDVR does
    full = torch.zeros(B, N, H, W, C, device='cuda')
    full[:, 2] = nhwc(_dev(ref))
    full.requires_grad_(True)
    wd, bd = _dev(w).requires_grad_(True), _dev(b).requires_grad_(True)
    yd = ops.conv([nd, ops.Seg(full[:, 2], T=N, Tsrc=1, t_fixed=0)], wd, bd, act=ops.ACT_LRELU)
    assert rel(nchw(yd), y) < TOL
    gd = torch.autograd.grad(yd, [nd, full, wd, bd], nhwc(_dev(gy)))
    assert rel(nchw(gd[0]), gr[0]) < GTOL
    assert rel(nchw(gd[1][:, 2]), gr[1]) < GTOL and float(gd[1][:, 0].abs().max()) == 0.0
    assert rel(gd[2], gr[2]) < GTOL and rel(gd[3], gr[3]) < GTOL


@pytest.mark.parametrize('Cin,Co', [(3, 64), (64, 64)])
def test_conv3d_replicate_padded(ops, Cin, Co):
    B, T, H, W = 2, 5, 6, 7
    x = _rand(B, Cin, T, H, W, seed=1)
    w, b = _rand(Co, Cin, 3, 3, 3, seed=2, scale=0.1), _rand(Co, seed=3, scale=0.1)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    y = F.leaky_relu(F.conv3d(F.pad(xr, (1,) * 6, mode='replicate'), wr, br), 0.1)      # B Co T H W
    gy = _rand(*y.shape, seed=4)
    gr = torch.autograd.grad(y, [xr, wr, br], gy)
    frames = _dev(x).permute(0, 2, 3, 4, 1).reshape(B * T, H, W, Cin).contiguous().requires_grad_(True)
    wd, bd = _dev(w).requires_grad_(True), _dev(b).requires_grad_(True)
    yd = ops.conv3d_padded(ops.pad3d_replicate(frames, T), wd, bd, T, act=ops.ACT_LRELU)   # [B*T, H, W, Co]
    y_ref = y.permute(0, 2, 3, 4, 1).reshape(B * T, H, W, Co)
    assert rel(yd, y_ref) < TOL
    gd = torch.autograd.grad(yd, [frames, wd, bd], _dev(gy.permute(0, 2, 3, 4, 1).reshape(B * T, H, W, Co)).contiguous())
    assert rel(gd[0], gr[0].permute(0, 2, 3, 4, 1).reshape(B * T, H, W, Cin)) < GTOL
    assert rel(gd[1], gr[1]) < GTOL and rel(gd[2], gr[2]) < GTOL


def _mdcn_case(ops, N, C, H, W, Co, dg, off_scale, act, seed):
    from oracle.torch_ops import mdcn_torch
    K = 9
    x = _rand(N, C, H, W, seed=seed)
    off = _rand(N, 2 * dg * K, H, W, seed=seed + 1, scale=off_scale)
    m = torch.sigmoid(_rand(N, dg * K, H, W, seed=seed + 2))
    w, b = _rand(Co, C, 3, 3, seed=seed + 3, scale=0.1), _rand(Co, seed=seed + 4, scale=0.1)
    leaves = [t.clone().requires_grad_(True) for t in (x, off, m, w, b)]
    y = mdcn_torch(*leaves, 1, 1, 1, 1, dg)
    if act:
        y = F.leaky_relu(y, 0.1)
    gy = _rand(*y.shape, seed=seed + 5)
    gr = torch.autograd.grad(y, leaves, gy)
    xd = nhwc(_dev(x)).requires_grad_(True)
    om = torch.cat([nhwc(_dev(off)), nhwc(_dev(m))], 3).contiguous().requires_grad_(True)
    wd, bd = _dev(w).requires_grad_(True), _dev(b).requires_grad_(True)
    yd = ops.mdcn(xd, om, wd, bd, dg, 1, 1, 1, ops.ACT_LRELU if act else ops.ACT_NONE)
    assert rel(nchw(yd), y) < TOL, 'forward'
    gd = torch.autograd.grad(yd, [xd, om, wd, bd], nhwc(_dev(gy)))
    assert rel(nchw(gd[0]), gr[0]) < GTOL, 'grad input'
    assert rel(nchw(gd[1][..., :2 * dg * K]), gr[1]) < GTOL, 'grad offset'
    assert rel(nchw(gd[1][..., 2 * dg * K:]), gr[2]) < GTOL, 'grad mask'
    assert rel(gd[2], gr[3]) < GTOL, 'grad weight'
    assert rel(gd[3], gr[4]) < GTOL, 'grad bias'


@pytest.mark.parametrize('cfg', [(1, 64, 11, 13, 64, 8, 2.0, True), (2, 64, 8, 8, 64, 8, 6.0, False),
                                 (1, 16, 9, 11, 8, 4, 3.0, False), (1, 128, 6, 7, 128, 8, 1.5, True),
                                 (1, 12, 7, 7, 8, 4, 2.0, False)],
                         ids=['edvr_m', 'oob_heavy', 'small_cpg4', 'edvr_l', 'cpg3_scalar'])
def test_mdcn_nhwc(ops, cfg):
    N, C, H, W, Co, dg, s, act = cfg
    _mdcn_case(ops, N, C, H, W, Co, dg, s, act, seed=7)


@pytest.mark.parametrize('shape,off_scale,kw', [((2, 19, 21), 1.0, {}), ((1, 33, 40), 3.0, {}), ((1, 16, 8), 12.0, {}),
                                                ((5, 44, 80), 2.0, {}), ((2, 24, 20), 1.5, dict(stride=2)),
                                                ((1, 20, 28), 1.5, dict(dil=2, pad=2))],
                         ids=['small_offsets', 'ragged_last_tile', 'mostly_outside_image', 'slr_size', 'stride2', 'dilated'])
def test_mdcn_tensor_core_backward(ops, shape, off_scale, kw):
    """dvsr_mdcn_bwd_tc (mdcn_bwd_tc.cu): grad_col = gy . W^T as tcgen05 MMAs into TMEM, consumed in-kernel for the input / offset /
    mask gradients, and the weight gradient from the modulated samples rebuilt on the way -- all five gradients against the
    float64 oracle (deform_conv_cuda.cpp:566-679 semantics), BF16x3 = fp32-class.  Also: gx accumulates into what it holds, and
    the launch-policy CTA budget does not change the numbers beyond the atomics' summation order."""
    from oracle.torch_ops import mdcn_torch
    N, H, W = shape
    stride, pad, dil = kw.get('stride', 1), kw.get('pad', 1), kw.get('dil', 1)
    x = _rand(N, 64, H, W, seed=1)
    Ho, Wo = (H + 2 * pad - (dil * 2 + 1)) // stride + 1, (W + 2 * pad - (dil * 2 + 1)) // stride + 1
    off = _rand(N, 144, Ho, Wo, seed=2, scale=off_scale)
    m = torch.sigmoid(_rand(N, 72, Ho, Wo, seed=3))
    w, b = _rand(64, 64, 3, 3, seed=4, scale=0.1), _rand(64, seed=5, scale=0.1)
    leaves = [t.clone().requires_grad_(True) for t in (x, off, m, w, b)]
    y = mdcn_torch(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], stride, pad, dil, 1, 8)
    gy = _rand(*y.shape, seed=6)
    gref = torch.autograd.grad(y, leaves, gy)
    ops.set_conv_backend(True)
    try:
        outs = []
        for budget in (0, 3):
            with ops.scope(ops.new_scope(ops.LaunchPolicy(cta_budget=budget))):
                xs = nhwc(_dev(x)).requires_grad_(True)
                om = torch.cat([nhwc(_dev(off)), nhwc(_dev(m))], 3).contiguous().requires_grad_(True)
                ws, bs = _dev(w).requires_grad_(True), _dev(b).requires_grad_(True)
                yd = ops.mdcn(xs, om, ws, bs, 8, stride, pad, dil)
                assert ops._lib.lib().dvsr_mdcn_bwd_tc_supported is not None
                outs.append(torch.autograd.grad(yd, [xs, om, ws, bs], nhwc(_dev(gy))))
    finally:
        ops.set_conv_backend(False)
    from util import rel_robust
    for gx, gom, gw, gb in outs:
        assert rel(nchw(gx), gref[0]) < 2e-4
        # the offset gradient is discontinuous where a sampling coordinate crosses an integer: ignore the 1e-4 worst elements
        # (fp32 vs fp64 floor of the same coordinate; measured 1.1e-3 plain relative L2 at 2.5 M elements, all from such flips)
        assert rel_robust(nchw(gom[..., :144]), gref[1]) < 2e-4 and rel(nchw(gom[..., :144]), gref[1]) < 5e-3
        assert rel(nchw(gom[..., 144:]), gref[2]) < 2e-4, rel(nchw(gom[..., 144:]), gref[2])
        assert rel(gw, gref[3]) < 2e-4, rel(gw, gref[3])
        assert rel(gb, gref[4]) < 2e-5, rel(gb, gref[4])
    # one CTA per tap group (budget 3) vs 74: same numbers up to the fp32 summation order of the in-TMEM / atomic accumulation
    assert rel(outs[1][0], outs[0][0]) < 1e-4, rel(outs[1][0], outs[0][0])
    assert rel(outs[1][2], outs[0][2]) < 1e-4, rel(outs[1][2], outs[0][2])


@pytest.mark.parametrize('staged', [2, 1], ids=['staged_window', 'direct_gather'])
@pytest.mark.parametrize('shape,off_scale', [((2, 19, 21), 1.0), ((1, 33, 40), 3.0), ((1, 16, 8), 12.0), ((5, 44, 80), 2.0)],
                         ids=['small_offsets', 'edge_of_window', 'mostly_outside_window', 'slr_size'])
def test_mdcn_tensor_core_forward(ops, staged, shape, off_scale):
    """tcgen05 DCN forward, both gather variants, against the float64 oracle: offsets inside the staged window (margin 3),
    straddling it, and far outside (global fallback + out-of-image corners)."""
    from oracle.torch_ops import mdcn_torch
    N, H, W = shape
    x = _rand(N, 64, H, W, seed=1)
    off = _rand(N, 144, H, W, seed=2, scale=off_scale)
    m = torch.sigmoid(_rand(N, 72, H, W, seed=3))
    w, b = _rand(64, 64, 3, 3, seed=4, scale=0.1), _rand(64, seed=5, scale=0.1)
    y = F.leaky_relu(mdcn_torch(x, off, m, w, b, 1, 1, 1, 1, 8), 0.1)
    om = torch.cat([nhwc(_dev(off)), nhwc(_dev(m))], 3).contiguous()
    ops.set_conv_backend(True)
    try:
        with ops.scope(ops.new_scope(ops.LaunchPolicy(mdcn_staged=staged))):       # dvsr_policy.mdcn_staged: 2 = staged, 1 = direct
            yd = ops.mdcn(nhwc(_dev(x)), om, _dev(w), _dev(b), 8, 1, 1, 1, ops.ACT_LRELU)
    finally:
        ops.set_conv_backend(False)
    assert rel(nchw(yd), y) < 1e-4          # BF16x3: fp32-class


def test_mdcn_all_taps_outside_gives_bias(ops):
    x = nhwc(_dev(_rand(1, 64, 5, 5)))
    om = torch.cat([torch.full((1, 5, 5, 144), 100.0), torch.ones(1, 5, 5, 72)], 3).cuda()
    w, b = _dev(_rand(64, 64, 3, 3, seed=1)), _dev(_rand(64, seed=2))
    y = ops.mdcn(x, om, w, b, 8)
    assert torch.equal(y, b.view(1, 1, 1, 64).expand_as(y))


def test_mdcn_nchw_boundary_matches_reference_golden(ops):
    """The reference operator API on NCHW tensors against outputs of torchvision's implementation of
    the reference algorithm (tests/golden/dcn_small.npz, generated by oracle/make_golden.py)."""
    from dynavsr_b200.models.archs.dcn import modulated_deform_conv
    g = gold('dcn_small.npz')
    t = {k: torch.from_numpy(g[k]).cuda() for k in ('x', 'offset', 'mask', 'weight', 'bias', 'gy')}
    leaves = [t[k].clone().requires_grad_(True) for k in ('x', 'offset', 'mask', 'weight', 'bias')]
    y = modulated_deform_conv(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], 1, 1, 1, 1, int(g['dg']))
    assert rel(y, torch.from_numpy(g['y'])) < TOL
    grads = torch.autograd.grad(y, leaves, t['gy'])
    for name, a in zip(['gx', 'goffset', 'gmask', 'gweight', 'gbias'], grads):
        assert rel(a, torch.from_numpy(g[name])) < GTOL, name


def test_mdcn_nchw_boundary_errors(ops):
    from dynavsr_b200.models.archs.dcn import modulated_deform_conv
    x = torch.randn(1, 8, 4, 4)
    with pytest.raises(NotImplementedError):      # CPU tensors: deform_conv.py:109-110
        modulated_deform_conv(x, torch.zeros(1, 18, 4, 4), torch.ones(1, 9, 4, 4), torch.randn(4, 8, 3, 3), None, 1, 1, 1, 1, 1)
    xc = x.cuda()
    with pytest.raises(RuntimeError):             # channel mismatch: deform_conv_cuda.cpp:509-511
        modulated_deform_conv(xc, torch.zeros(1, 18, 4, 4).cuda(), torch.ones(1, 9, 4, 4).cuda(),
                              torch.randn(4, 6, 3, 3).cuda(), None, 1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError):             # groups != 1 is rejected loudly (never used by the reference)
        modulated_deform_conv(xc, torch.zeros(1, 18, 4, 4).cuda(), torch.ones(1, 9, 4, 4).cuda(),
                              torch.randn(4, 4, 3, 3).cuda(), None, 1, 1, 1, 2, 1)


def test_dcn_pack_module_matches_oracle(ops):
    """ModulatedDeformConvPack through its reference-shaped NCHW forward (deform_conv.py:274-291)."""
    from dynavsr_b200.models.archs.dcn import ModulatedDeformConvPack
    from oracle import edvr_oracle as O
    torch.manual_seed(0)
    m = ModulatedDeformConvPack(64, 64, 3, stride=1, padding=1, dilation=1, deformable_groups=8,
                                extra_offset_mask=True).cuda()
    assert float(m.conv_offset_mask.weight.abs().max()) == 0.0          # zero init, deform_conv.py:270-272
    m.conv_offset_mask.weight.data.normal_(0, 0.02)
    m.conv_offset_mask.bias.data.normal_(0, 0.5)
    x, feat = _rand(2, 64, 9, 10, seed=1), _rand(2, 64, 9, 10, seed=2)
    sd = {'p.' + k: v.detach().double().cpu() for k, v in m.state_dict().items()}
    ref = O.dcn_pack(sd, 'p', x, feat, 8)
    y = m([_dev(x), _dev(feat)])
    assert rel(y, ref) < TOL


def _dcn_v1_case(seed=3):
    from oracle.torch_ops import mdcn_torch
    x, w = _rand(2, 16, 9, 11, seed=seed), _rand(12, 16, 3, 3, seed=seed + 1, scale=0.2)
    off = _rand(2, 2 * 2 * 9, 9, 11, seed=seed + 2, scale=1.5)
    gy = _rand(2, 12, 9, 11, seed=seed + 3)
    leaves = [t.clone().requires_grad_(True) for t in (x, off, w)]
    y = mdcn_torch(leaves[0], leaves[1], torch.ones(2, 18, 9, 11, dtype=x.dtype), leaves[2], None, 1, 1, 1, 1, 2)   # DCNv1 = mask of ones
    return x, off, w, gy, y.detach(), torch.autograd.grad(y, leaves, gy)


def test_dcn_v1_function_and_modules_vs_oracle(ops):
    """DeformConv* (deform_conv.py:15-94,161-218; kernel.cu:189-464): exported by the reference, used by no YML."""
    from dynavsr_b200.models.archs.dcn import DeformConv, DeformConvPack, deform_conv
    x, off, w, gy, y_ref, g_ref = _dcn_v1_case()
    leaves = [_dev(t).requires_grad_(True) for t in (x, off, w)]
    y = deform_conv(leaves[0], leaves[1], leaves[2], 1, 1, 1, 1, 2)
    assert rel(y, y_ref) < TOL
    for name, a, b in zip(('gx', 'goffset', 'gweight'), torch.autograd.grad(y, leaves, _dev(gy)), g_ref):
        assert rel(a, b) < GTOL, name
    m = DeformConv(16, 12, 3, stride=1, padding=1, deformable_groups=2).cuda()
    assert list(m.state_dict().keys()) == ['weight']
    m.weight.data.copy_(_dev(w))
    assert rel(m(_dev(x), _dev(off)), y_ref) < TOL
    pk = DeformConvPack(16, 12, 3, stride=1, padding=1, deformable_groups=2).cuda()
    assert list(pk.state_dict().keys()) == ['weight', 'conv_offset.weight', 'conv_offset.bias']
    assert float(pk.conv_offset.weight.abs().max()) == 0.0                                    # zero init -> plain convolution
    pk.weight.data.copy_(_dev(w))
    assert rel(pk(_dev(x)), F.conv2d(x, w, None, 1, 1)) < TOL
    with pytest.raises(NotImplementedError):
        deform_conv(x.float(), off.float(), w.float(), 1, 1, 1, 1, 2)
    with pytest.raises(ValueError):
        deform_conv(_dev(x)[0], _dev(off), _dev(w))


def test_legacy_deform_conv_cuda_module(ops):
    """The five pybind names of the reference extension (deform_conv_cuda.cpp:681-695) with their positional signatures,
    caller-allocated outputs and accumulate-into-gradient behaviour."""
    from dynavsr_b200.models.archs.dcn import deform_conv_cuda as D
    g = gold('dcn_small.npz')
    t = {k: torch.from_numpy(g[k]).cuda() for k in ('x', 'offset', 'mask', 'weight', 'bias', 'gy')}
    dg = int(g['dg'])
    empty = t['x'].new_empty(0)
    out = t['x'].new_empty(tuple(g['y'].shape))
    D.modulated_deform_conv_cuda_forward(t['x'], t['weight'], t['bias'], empty, t['offset'], t['mask'], out, empty,
                                         3, 3, 1, 1, 1, 1, 1, 1, 1, dg, True)
    assert rel(out, torch.from_numpy(g['y'])) < TOL
    gi, gw, gb = torch.zeros_like(t['x']), torch.ones_like(t['weight']), torch.zeros_like(t['bias'])
    go, gm = torch.zeros_like(t['offset']), torch.zeros_like(t['mask'])
    D.modulated_deform_conv_cuda_backward(t['x'], t['weight'], t['bias'], empty, t['offset'], t['mask'], empty, gi, gw, gb, go, gm,
                                          t['gy'], 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, True)
    for name, a in (('gx', gi), ('goffset', go), ('gmask', gm), ('gbias', gb)):
        assert rel(a, torch.from_numpy(g[name])) < GTOL, name
    assert rel(gw - 1.0, torch.from_numpy(g['gweight'])) < GTOL                               # accumulated onto the caller's values
    with pytest.raises(RuntimeError):                                                         # AT_CHECK(is_contiguous)
        D.modulated_deform_conv_cuda_forward(t['x'].transpose(2, 3), t['weight'], t['bias'], empty, t['offset'], t['mask'], out,
                                             empty, 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, True)
    # DCNv1 trio
    x, off, w, gy, y_ref, g_ref = _dcn_v1_case(7)
    xd, od, wd, gyd = _dev(x), _dev(off), _dev(w), _dev(gy)
    out = xd.new_empty(0)
    D.deform_conv_forward_cuda(xd, wd, od, out, empty, empty, 3, 3, 1, 1, 1, 1, 1, 1, 1, 2, 2)
    assert tuple(out.shape) == tuple(y_ref.shape) and rel(out, y_ref) < TOL
    gi, go, gw = torch.zeros_like(xd), torch.zeros_like(od), torch.zeros_like(wd)
    D.deform_conv_backward_input_cuda(xd, od, gyd, gi, go, wd, empty, 3, 3, 1, 1, 1, 1, 1, 1, 1, 2, 2)
    D.deform_conv_backward_parameters_cuda(xd, od, gyd, gw, empty, empty, 3, 3, 1, 1, 1, 1, 1, 1, 1, 2, 1, 2)
    for name, a, b in zip(('gx', 'goffset', 'gweight'), (gi, go, gw), g_ref):
        assert rel(a, b) < GTOL, name


@pytest.mark.parametrize('scale,C,mul', [(2, 64, 1.0), (2, 64, 2.0), (4, 3, 1.0), (2, 6, 1.0)])
def test_upsample(ops, scale, C, mul):
    x = _rand(2, C, 7, 9, seed=1)
    xr = x.clone().requires_grad_(True)
    y = F.interpolate(xr, scale_factor=scale, mode='bilinear', align_corners=False) * mul
    gy = _rand(*y.shape, seed=2)
    gr, = torch.autograd.grad(y, xr, gy)
    xd = nhwc(_dev(x)).requires_grad_(True)
    yd = ops.upsample(xd, scale, mul)
    assert rel(nchw(yd), y) < TOL
    gd, = torch.autograd.grad(yd, xd, nhwc(_dev(gy)))
    assert rel(nchw(gd), gr) < TOL


@pytest.mark.parametrize('H,W,C', [(8, 12, 64), (9, 7, 64), (6, 6, 5)])
def test_pool_maxavg(ops, H, W, C):
    x = _rand(2, C, H, W, seed=1)
    xr = x.clone().requires_grad_(True)
    y = torch.cat([F.max_pool2d(xr, 3, 2, 1), F.avg_pool2d(xr, 3, 2, 1)], 1)
    gy = _rand(*y.shape, seed=2)
    gr, = torch.autograd.grad(y, xr, gy)
    xd = nhwc(_dev(x)).requires_grad_(True)
    yd = ops.pool_maxavg(xd)
    assert rel(nchw(yd), y) < TOL
    gd, = torch.autograd.grad(yd, xd, nhwc(_dev(gy)))
    assert rel(nchw(gd), gr) < TOL


@pytest.mark.parametrize('mode', ['reflect', 'replicate'])
@pytest.mark.parametrize('C', [64, 3])
def test_pad2d(ops, mode, C):
    x = _rand(2, C, 5, 6, seed=1)
    xr = x.clone().requires_grad_(True)
    y = F.pad(xr, (1, 1, 1, 1), mode=mode)
    gy = _rand(*y.shape, seed=2)
    gr, = torch.autograd.grad(y, xr, gy)
    xd = nhwc(_dev(x)).requires_grad_(True)
    yd = ops.pad2d(xd, 1, mode)
    assert maxabs(nchw(yd), y.float()) == 0.0
    gd, = torch.autograd.grad(yd, xd, nhwc(_dev(gy)))
    assert rel(nchw(gd), gr) < TOL


def test_pad3d_replicate(ops):
    B, T, C, H, W = 2, 5, 64, 4, 5
    x = _rand(B, C, T, H, W, seed=1)
    xr = x.clone().requires_grad_(True)
    y = F.pad(xr, (1,) * 6, mode='replicate')
    gy = _rand(*y.shape, seed=2)
    gr, = torch.autograd.grad(y, xr, gy)
    fr = _dev(x).permute(0, 2, 3, 4, 1).reshape(B * T, H, W, C).contiguous().requires_grad_(True)
    yd = ops.pad3d_replicate(fr, T)
    assert maxabs(yd, y.permute(0, 2, 3, 4, 1).reshape(B * (T + 2), H + 2, W + 2, C).float()) == 0.0
    gd, = torch.autograd.grad(yd, fr, _dev(gy.permute(0, 2, 3, 4, 1).reshape(B * (T + 2), H + 2, W + 2, C)).contiguous())
    assert rel(gd, gr.permute(0, 2, 3, 4, 1).reshape(B * T, H, W, C)) < TOL


def test_tsa_temporal_and_combine(ops):
    B, N, C, H, W = 2, 5, 64, 6, 7
    al, emb, ref = _rand(B, N, C, H, W, seed=1), _rand(B, N, C, H, W, seed=2, scale=0.2), _rand(B, C, H, W, seed=3, scale=0.2)
    a, e, r = (t.clone().requires_grad_(True) for t in (al, emb, ref))
    prob = torch.sigmoid((e * r.unsqueeze(1)).sum(2))                    # B N H W
    out = (a * prob.unsqueeze(2)).reshape(B, N * C, H, W)
    gy = _rand(*out.shape, seed=4)
    gr = torch.autograd.grad(out, [a, e, r], gy)
    ad = nhwc(_dev(al.reshape(B * N, C, H, W))).requires_grad_(True)
    ed = nhwc(_dev(emb.reshape(B * N, C, H, W))).requires_grad_(True)
    rd = nhwc(_dev(ref)).requires_grad_(True)
    od = ops.tsa_temporal(ad, ed, rd, N)
    assert rel(nchw(od), out) < TOL
    gd = torch.autograd.grad(od, [ad, ed, rd], nhwc(_dev(gy)))
    assert rel(nchw(gd[0]), gr[0].reshape(B * N, C, H, W)) < GTOL
    assert rel(nchw(gd[1]), gr[1].reshape(B * N, C, H, W)) < GTOL
    assert rel(nchw(gd[2]), gr[2]) < GTOL
    # combine
    f, t, ad2 = _rand(2, 64, 5, 5, seed=5), _rand(2, 64, 5, 5, seed=6), _rand(2, 64, 5, 5, seed=7)
    fr, tr, ar = (v.clone().requires_grad_(True) for v in (f, t, ad2))
    o = fr * torch.sigmoid(tr) * 2 + ar
    g = _rand(*o.shape, seed=8)
    gr = torch.autograd.grad(o, [fr, tr, ar], g)
    fd, td, dd = (_dev(v).requires_grad_(True) for v in (f, t, ad2))
    od = ops.tsa_combine(fd, td, dd)
    assert rel(od, o) < TOL
    gd = torch.autograd.grad(od, [fd, td, dd], _dev(g))
    for i in range(3):
        assert rel(gd[i], gr[i]) < TOL


@pytest.mark.parametrize('kind', ['l1', 'l2', 'cb', 'huber'])
def test_pixel_loss(ops, kind):
    from oracle import edvr_oracle as O
    sc = 0.02 if kind == 'huber' else 1.0        # straddle the Huber delta (1e-2, loss.py:8)
    a, b = _rand(2, 8, 9, 3, seed=1) * sc, _rand(2, 8, 9, 3, seed=2) * sc
    ar = a.clone().requires_grad_(True)
    ref = O.pixel_loss(kind, ar, b) * 10.0
    gr, = torch.autograd.grad(ref * 3.0, ar)
    ad = _dev(a).requires_grad_(True)
    l = ops.pixel_loss(ad, _dev(b), kind, 10.0, 1e-2 if kind == 'huber' else 1e-6)
    assert abs(float(l) - float(ref)) < 1e-5 * abs(float(ref))
    gd, = torch.autograd.grad(l * 3.0, ad)
    assert rel(gd, gr) < TOL


def test_fused_updates_match_torch_optim(ops):
    from dynavsr_b200.adapt import FlatParams
    torch.manual_seed(0)
    m1, m2 = torch.nn.Conv2d(8, 8, 3).cuda(), torch.nn.Conv2d(8, 4, 1).cuda()
    r1, r2 = torch.nn.Conv2d(8, 8, 3).cuda(), torch.nn.Conv2d(8, 4, 1).cuda()
    r1.load_state_dict(m1.state_dict()); r2.load_state_dict(m2.state_dict())
    for opt_name, wd in (('SGD', 0.0), ('Adam', 0.0), ('SGD', 0.05), ('Adam', 0.05)):
        flat = FlatParams([m1, m2])
        if opt_name == 'SGD':
            ref = torch.optim.SGD([{'params': r1.parameters(), 'lr': 0.1}, {'params': r2.parameters(), 'lr': 0.01}], weight_decay=wd)
        else:
            ref = torch.optim.Adam([{'params': r1.parameters(), 'lr': 0.1}, {'params': r2.parameters(), 'lr': 0.01}], betas=(0.9, 0.99),
                                   weight_decay=wd)
        for step in range(3):
            flat.zero_grad()
            for (p, q) in zip(list(m1.parameters()) + list(m2.parameters()), list(r1.parameters()) + list(r2.parameters())):
                g = torch.randn_like(p)
                p._dvsr_grad.copy_(g)
                q.grad = g.clone()
            if opt_name == 'SGD':
                flat.sgd_step(0.1, 0.01, weight_decay=wd)
            else:
                flat.adam_step(0.1, 0.01, (0.9, 0.99), weight_decay=wd)
            ref.step()
        for (p, q) in zip(list(m1.parameters()) + list(m2.parameters()), list(r1.parameters()) + list(r2.parameters())):
            assert rel(p, q) < 1e-5, opt_name
        flat.restore()
        flat.release()                      # hand the parameters back before the next FlatParams adopts them (single owner)
        r1.load_state_dict(m1.state_dict()); r2.load_state_dict(m2.state_dict())


@pytest.mark.parametrize('adam', [True, False], ids=['adam', 'sgd'])
@pytest.mark.parametrize('world', [1, 2, 3], ids=['1rank', '2ranks', '3ranks'])
def test_sliced_peer_exchange_kernel_in_one_process(ops, world, adam):
    """dvsr_update_peers_sliced (update.cu), the reduce-scatter form of the meta step's exchange + outer update, with the
    `world` exchange buffers living on ONE GPU (what peer pointers of a symmetric-memory allocation look like to the kernel):
    after every rank's launch each buffer holds the complete new weights = torch.optim step on the rank-ordered mean gradient,
    bit-identical across the buffers; each rank only touches its slice of the moments."""
    import ctypes
    from dynavsr_b200._lib import call
    n = 4 * 1000 + 64                  # not a multiple of the slice size for 3 ranks
    g = torch.Generator().manual_seed(world)
    p0 = torch.randn(n, generator=g).cuda()
    grads = [torch.randn(n, generator=g).cuda() for _ in range(world)]
    bufs = [t.clone() for t in grads]
    table = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device='cuda')
    m, v = [torch.zeros(n, device='cuda') for _ in range(world)], [torch.zeros(n, device='cuda') for _ in range(world)]
    lr, b1, b2, step = 1e-2, 0.9, 0.99, 1
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for r in range(world):            # every rank reads its slice of all (still unreduced) buffers, then writes it back everywhere
        call('dvsr_update_peers_sliced', P(p0), P(table), world, r, 0, 1.0 / world, P(m[r]), P(v[r]), n, n, lr, lr, b1, b2, 1e-8,
             1 - b1 ** step, 1 - b2 ** step, 0.0, 1 if adam else 0, st)
    torch.cuda.synchronize()
    ref = torch.nn.Parameter(p0.clone())
    opt = (torch.optim.Adam([ref], lr=lr, betas=(b1, b2)) if adam else torch.optim.SGD([ref], lr=lr))
    mean = grads[0].clone()
    for t in grads[1:]:
        mean += t
    ref.grad = mean / world
    opt.step()
    for b in bufs:
        assert rel(b, ref.detach()) < 1e-6
        assert torch.equal(b, bufs[0])
    per = ((n // 4 + world - 1) // world) * 4
    for r in range(world):
        if adam:
            lo, hi = r * per, min(n, (r + 1) * per)
            assert float(m[r][lo:hi].abs().sum()) > 0 and float(m[r][:lo].abs().sum()) == 0 and float(m[r][hi:].abs().sum()) == 0


def test_layout_roundtrip_and_cpu_rejection(ops):
    x = torch.randn(3, 5, 7, 9).cuda()
    y = ops.to_nhwc(x)
    assert torch.equal(y, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.to_nchw(y), x)
    with pytest.raises(NotImplementedError):
        ops.conv(torch.randn(1, 4, 4, 8), torch.randn(8, 8, 3, 3))


# ---------------------------------------------------------------------------------------------------
# tensor-core (tcgen05 / TMEM / TMA, TF32 operands rounded to nearest, fp32 accumulation) convolution family.
# Per-layer tolerance 1e-3 relative L2 against the float64 CPU reference (measured ~3e-4 forward / data
# gradient, ~8e-4 weight gradient).
# Backward cases use act = 0 (or the smooth sigmoid): with a ReLU-family activation a pre-activation within the TF32 error of
# zero flips the derivative mask relative to the fp32 CPU reference (about one element per 5 000 outputs), which is a
# property of the comparison, not of the kernels; activation derivatives are covered by the exact-fp32 cases above.
TC_TOL = 1e-3
TC_CASES = [
    # name, N, H, W, segC, Co, k, stride, pad, act, res, shuffle
    ('tc_3x3_64', 2, 20, 40, [64], 64, 3, 1, 1, 0, False, 0),
    ('tc_3x3_res', 1, 9, 13, [64], 64, 3, 1, 1, 0, True, 0),
    ('tc_cat2', 5, 11, 20, [64, 64], 64, 3, 1, 1, 0, False, 0),
    ('tc_offmask216', 1, 12, 20, [64], 216, 3, 1, 1, 3, False, 0),
    ('tc_1x1_320', 1, 16, 24, [320], 64, 1, 1, 0, 0, False, 0),
    ('tc_shuffle', 1, 8, 16, [64], 256, 3, 1, 1, 0, False, 2),
    ('tc_valid_pad0', 2, 18, 34, [64], 64, 3, 1, 0, 0, False, 0),
    ('tc_4x4s2', 1, 18, 34, [64], 128, 4, 2, 0, 0, False, 0),
    ('tc_4x4s2_128', 2, 10, 18, [128], 64, 4, 2, 0, 0, False, 0),
    ('tc_3x3s2', 2, 16, 24, [64], 64, 3, 2, 1, 0, False, 0),
    ('tc_4x4s2_p1', 1, 16, 24, [64], 128, 4, 2, 1, 0, False, 0),
    ('tc_3x3s2_odd', 1, 15, 21, [64], 64, 3, 2, 1, 0, False, 0),
    ('tc_last_64to3_res', 1, 24, 40, [64], 3, 3, 1, 1, 0, True, 0),       # conv_last + bilinear skip: narrow scalar epilogue
    ('tc_1x1_64to3', 2, 9, 13, [64], 3, 1, 1, 0, 0, False, 0),            # MFDN conv6
    ('tc_3x3_64to24', 1, 10, 12, [64], 24, 3, 1, 1, 0, False, 0),
]


@pytest.mark.parametrize('case', TC_CASES, ids=[c[0] for c in TC_CASES])
def test_conv_tensor_core_forward_backward(ops, case):
    name, N, H, W, segC, Co, k, stride, pad, act, use_res, shuffle = case
    xs = [_rand(N, c, H, W, seed=i + 1) for i, c in enumerate(segC)]
    w = _rand(Co, sum(segC), k, k, seed=10, scale=(2.0 / (sum(segC) * k * k)) ** 0.5)
    b = _rand(Co, seed=11, scale=0.1)
    sig_split = 144 if act == 3 else 0
    xr = [t.clone().requires_grad_(True) for t in xs]
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = _act_ref(F.conv2d(torch.cat(xr, 1), wr, br, stride=stride, padding=pad), act, sig_split)
    if shuffle:
        y = F.pixel_shuffle(y, 2)
    res = _rand(*y.shape, seed=12) if use_res else None
    if use_res:
        y = y + res
    gy = _rand(*y.shape, seed=13)
    grads = torch.autograd.grad(y, xr + [wr, br], gy)
    ops.set_conv_backend(True)
    try:
        xd = [nhwc(_dev(t)).requires_grad_(True) for t in xs]
        wd, bd = _dev(w).requires_grad_(True), _dev(b).requires_grad_(True)
        resd = nhwc(_dev(res)) if use_res else None
        n0 = ops._lib.COUNTER[0]
        yd = ops.conv(xd, wd, bd, stride=stride, pad=pad, act=act, sig_split=sig_split, res=resd, shuffle=shuffle)
        assert rel(nchw(yd), y) < TC_TOL, 'forward'
        gd = torch.autograd.grad(yd, xd + [wd, bd], nhwc(_dev(gy)))
    finally:
        ops.set_conv_backend(False)
    for i in range(len(xs)):
        assert rel(nchw(gd[i]), grads[i]) < TC_TOL, 'grad input %d' % i
    assert rel(gd[len(xs)], grads[len(xs)]) < 2 * TC_TOL, 'grad weight'
    assert rel(gd[len(xs) + 1], grads[len(xs) + 1]) < 1e-4, 'grad bias'


def test_conv_tensor_core_broadcast_segment(ops):
    B, N, H, W, C = 2, 5, 12, 16, 64
    nbr, ref = _rand(B * N, C, H, W, seed=1), _rand(B, C, H, W, seed=2)
    w, b = _rand(C, 2 * C, 3, 3, seed=3, scale=0.03), _rand(C, seed=4, scale=0.1)
    nr, rr, wr, br = (t.clone().requires_grad_(True) for t in (nbr, ref, w, b))
    y = F.conv2d(torch.cat([nr, rr.repeat_interleave(N, 0)], 1), wr, br, padding=1)
    gy = _rand(*y.shape, seed=5)
    gr = torch.autograd.grad(y, [nr, rr, wr, br], gy)
    ops.set_conv_backend(True)
    try:
        nd = nhwc(_dev(nbr)).requires_grad_(True)
        full = torch.zeros(B, N, H, W, C, device='cuda')
        full[:, 2] = nhwc(_dev(ref))
        full.requires_grad_(True)
        wd, bd = _dev(w).requires_grad_(True), _dev(b).requires_grad_(True)
        yd = ops.conv([nd, ops.Seg(full[:, 2], T=N, Tsrc=1, t_fixed=0)], wd, bd)
        assert rel(nchw(yd), y) < TC_TOL
        gd = torch.autograd.grad(yd, [nd, full, wd, bd], nhwc(_dev(gy)))
    finally:
        ops.set_conv_backend(False)
    assert rel(nchw(gd[0]), gr[0]) < TC_TOL
    assert rel(nchw(gd[1][:, 2]), gr[1]) < TC_TOL
    assert rel(gd[2], gr[2]) < 2 * TC_TOL and rel(gd[3], gr[3]) < 1e-4


def test_conv3d_tensor_core(ops):
    B, T, H, W, Cin, Co = 1, 5, 10, 14, 64, 64
    x = _rand(B, Cin, T, H, W, seed=1)
    w, b = _rand(Co, Cin, 3, 3, 3, seed=2, scale=0.03), _rand(Co, seed=3, scale=0.1)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    y = F.conv3d(F.pad(xr, (1,) * 6, mode='replicate'), wr, br)
    gy = _rand(*y.shape, seed=4)
    gr = torch.autograd.grad(y, [xr, wr, br], gy)
    ops.set_conv_backend(True)
    try:
        frames = _dev(x).permute(0, 2, 3, 4, 1).reshape(B * T, H, W, Cin).contiguous().requires_grad_(True)
        wd, bd = _dev(w).requires_grad_(True), _dev(b).requires_grad_(True)
        yd = ops.conv3d_padded(ops.pad3d_replicate(frames, T), wd, bd, T)
        assert rel(yd, y.permute(0, 2, 3, 4, 1).reshape(B * T, H, W, Co)) < TC_TOL
        gd = torch.autograd.grad(yd, [frames, wd, bd], _dev(gy.permute(0, 2, 3, 4, 1).reshape(B * T, H, W, Co)).contiguous())
    finally:
        ops.set_conv_backend(False)
    assert rel(gd[0], gr[0].permute(0, 2, 3, 4, 1).reshape(B * T, H, W, Cin)) < TC_TOL
    assert rel(gd[1], gr[1]) < 2 * TC_TOL and rel(gd[2], gr[2]) < 1e-4


def test_conv3d_rgb_tensor_core(ops):
    """MFDN conv0 (Conv3d 3 -> 64 over the replication-padded RGB clip): temporal taps folded into channels."""
    B, T, H, W, Cin, Co = 2, 5, 12, 18, 3, 64
    x = _rand(B, Cin, T, H, W, seed=1)
    w, b = _rand(Co, Cin, 3, 3, 3, seed=2, scale=0.1), _rand(Co, seed=3, scale=0.1)
    wr, br = (t.clone().requires_grad_(True) for t in (w, b))
    y = F.conv3d(F.pad(x, (1,) * 6, mode='replicate'), wr, br)
    gy = _rand(*y.shape, seed=4)
    gr = torch.autograd.grad(y, [wr, br], gy)
    frames = _dev(x).permute(0, 2, 3, 4, 1).reshape(B * T, H, W, Cin).contiguous()
    wd, bd = _dev(w).requires_grad_(True), _dev(b).requires_grad_(True)
    ops.set_conv_backend(True)
    try:
        yd = ops.conv3d_rgb(frames, wd, bd, T)
        assert rel(yd, y.permute(0, 2, 3, 4, 1).reshape(B * T, H, W, Co)) < 1e-4        # BF16x3: fp32-class
        gd = torch.autograd.grad(yd, [wd, bd], _dev(gy.permute(0, 2, 3, 4, 1).reshape(B * T, H, W, Co)).contiguous())
        with pytest.raises(NotImplementedError):
            f2 = frames.clone().requires_grad_(True)
            torch.autograd.grad(ops.conv3d_rgb(f2, wd, bd, T).sum(), [f2])
    finally:
        ops.set_conv_backend(False)
    assert rel(gd[0], gr[0]) < 2 * TC_TOL and rel(gd[1], gr[1]) < 1e-4


def test_packed_weight_cache_never_aliases_freed_weights(ops):
    """Regression: packs are keyed by the weight OBJECT, not its device address (addresses get recycled)."""
    x = torch.randn(1, 8, 8, 64, device='cuda')
    outs = []
    for seed in range(4):
        w = _dev(_rand(64, 64, 3, 3, seed=seed, scale=0.05))    # freed after each iteration -> address reuse
        outs.append(ops.conv(x, w).clone())
        ref = F.conv2d(x.permute(0, 3, 1, 2).double().cpu(), w.double().cpu(), padding=1)
        assert rel(nchw(outs[-1]), ref) < TOL
        del w


@pytest.mark.parametrize('shape', [(5, 44, 80, 64, 64), (1, 33, 40, 128, 64), (2, 16, 24, 64, 216)],
                         ids=['slr_trunk', 'two_segments_worth_of_K', 'offset_mask_conv'])
def test_conv_tc2_single_product_mode(ops, shape):
    """``set_conv_backend(True, 'bf16')``: the resident-weight kernel issues only x_hi . w_hi on the BF16x3 layouts.  Expected:
    the result of bf16-rounded operands with fp32 accumulation (so ~1e-6 against a reference computed from rounded operands,
    ~3e-3 against fp32), and the default mode unchanged afterwards."""
    N, H, W, Ci, Co = shape
    x, w, b = _rand(N, Ci, H, W, seed=11), _rand(Co, Ci, 3, 3, seed=12, scale=0.05), _rand(Co, seed=13, scale=0.1)
    exact = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    rounded = F.conv2d(x.bfloat16().double(), w.bfloat16().double(), b.double(), padding=1)
    # 128 input channels reach the resident-weight kernel as two 64-channel K-segments (torch.cat feeding a conv, as in PCD):
    # the host K-splits them over two launches; a single 128-channel tensor would go to the streaming TF32 kernel instead
    segs = lambda t: [nhwc(_dev(t[:, i:i + 64])) for i in range(0, Ci, 64)] if Ci > 64 else nhwc(_dev(t))
    try:
        ops.set_conv_backend(True, 'bf16')
        y1 = nchw(ops.conv(segs(x), _dev(w), _dev(b), stride=1, pad=1))
        ops.set_conv_backend(True, 'bf16x3')
        y3 = nchw(ops.conv(segs(x), _dev(w), _dev(b), stride=1, pad=1))
    finally:
        ops.set_conv_backend(False, 'bf16x3')
    assert rel(y1, rounded) < 2e-5                       # exactly the single product (fp32 accumulation order aside)
    assert 5e-4 < rel(y1, exact) < 1e-2                  # bf16-level error against fp32
    assert rel(y3, exact) < 5e-5                         # the default mode is untouched
