"""oracle/edvr_oracle.py -- TEST INFRASTRUCTURE ONLY.

Functional, plain-PyTorch CPU restatement of the reference hot path: EDVR forward
(feature pyramid, PCD alignment, TSA fusion, reconstruction trunk, PixelShuffle head),
the MFDN down-scaling net and the test-time inner adaptation step.  Every function works
on a flat ``dict`` of tensors that uses the reference's ``state_dict`` key names, so the
same weights can be handed to the reference modules, to this oracle and to the CUDA
product path.  Autograd supplies the backward pass.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference legs may
import this module.  The product package must never route through it.

Reference files restated:
  codes/models/archs/EDVR_arch.py:95-128    PCD_Align.forward
  codes/models/archs/EDVR_arch.py:163-203   TSA_Fusion.forward
  codes/models/archs/EDVR_arch.py:254-313   EDVR.forward
  codes/models/archs/arch_util.py:48-52     ResidualBlock_noBN.forward
  codes/models/archs/dcn/deform_conv.py:274-291  ModulatedDeformConvPack.forward
  codes/models/archs/LRimg_estimator.py:92-117   DirectKernelEstimatorVideo.forward (MFDN)
  codes/models/loss.py:19-30                CharbonnierLoss
  codes/test_dynavsr.py:235-283             inner adaptation loop + final forward

Parity pin: checked against the unmodified reference modules (DCN routed to
torchvision.ops.deform_conv2d) by oracle/make_golden.py -> tests/golden/*.npz.
"""
import torch
import torch.nn.functional as F

from .torch_ops import mdcn_torch


def _conv(sd, name, x, stride=1, padding=1):
    return F.conv2d(x, sd[name + '.weight'], sd[name + '.bias'], stride=stride, padding=padding)


def _lrelu(x):
    return F.leaky_relu(x, 0.1)


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)


def res_block(sd, name, x):
    """arch_util.py:48-52: x + conv2(relu(conv1(x)))."""
    return x + _conv(sd, name + '.conv2', F.relu(_conv(sd, name + '.conv1', x)))


def dcn_pack(sd, name, x, feat, groups):
    """deform_conv.py:274-291 with extra_offset_mask=True (EDVR_arch.py:70-91)."""
    out = _conv(sd, name + '.conv_offset_mask', feat)
    o1, o2, m = torch.chunk(out, 3, dim=1)
    offset = torch.cat((o1, o2), dim=1)
    mask = torch.sigmoid(m)
    return mdcn_torch(x, offset, mask, sd[name + '.weight'], sd[name + '.bias'], 1, 1, 1, 1, groups)


def pcd_align(sd, p, nbr, ref, groups):
    """EDVR_arch.py:95-128.  nbr/ref: [L1, L2, L3] feature lists."""
    # L3
    o3 = _lrelu(_conv(sd, p + 'L3_offset_conv1', torch.cat([nbr[2], ref[2]], 1)))
    o3 = _lrelu(_conv(sd, p + 'L3_offset_conv2', o3))
    f3 = _lrelu(dcn_pack(sd, p + 'L3_dcnpack', nbr[2], o3, groups))
    # L2
    o2 = _lrelu(_conv(sd, p + 'L2_offset_conv1', torch.cat([nbr[1], ref[1]], 1)))
    o2 = _lrelu(_conv(sd, p + 'L2_offset_conv2', torch.cat([o2, _up2(o3) * 2], 1)))
    o2 = _lrelu(_conv(sd, p + 'L2_offset_conv3', o2))
    f2 = dcn_pack(sd, p + 'L2_dcnpack', nbr[1], o2, groups)
    f2 = _lrelu(_conv(sd, p + 'L2_fea_conv', torch.cat([f2, _up2(f3)], 1)))
    # L1
    o1 = _lrelu(_conv(sd, p + 'L1_offset_conv1', torch.cat([nbr[0], ref[0]], 1)))
    o1 = _lrelu(_conv(sd, p + 'L1_offset_conv2', torch.cat([o1, _up2(o2) * 2], 1)))
    o1 = _lrelu(_conv(sd, p + 'L1_offset_conv3', o1))
    f1 = dcn_pack(sd, p + 'L1_dcnpack', nbr[0], o1, groups)
    f1 = _conv(sd, p + 'L1_fea_conv', torch.cat([f1, _up2(f2)], 1))
    # cascade
    oc = _lrelu(_conv(sd, p + 'cas_offset_conv1', torch.cat([f1, ref[0]], 1)))
    oc = _lrelu(_conv(sd, p + 'cas_offset_conv2', oc))
    return _lrelu(dcn_pack(sd, p + 'cas_dcnpack', f1, oc, groups))


def tsa_fusion(sd, p, aligned, center):
    """EDVR_arch.py:163-203.  aligned: [B, N, C, H, W]."""
    B, N, C, H, W = aligned.shape
    emb_ref = _conv(sd, p + 'tAtt_2', aligned[:, center])
    emb = _conv(sd, p + 'tAtt_1', aligned.reshape(-1, C, H, W)).view(B, N, -1, H, W)
    cor = torch.stack([(emb[:, i] * emb_ref).sum(1) for i in range(N)], 1)      # B, N, H, W
    prob = torch.sigmoid(cor).unsqueeze(2)                                       # B, N, 1, H, W
    al = (aligned * prob).reshape(B, N * C, H, W)
    fea = _lrelu(_conv(sd, p + 'fea_fusion', al, padding=0))
    att = _lrelu(_conv(sd, p + 'sAtt_1', al, padding=0))
    att = _lrelu(_conv(sd, p + 'sAtt_2',
                       torch.cat([F.max_pool2d(att, 3, 2, 1), F.avg_pool2d(att, 3, 2, 1)], 1), padding=0))
    att_L = _lrelu(_conv(sd, p + 'sAtt_L1', att, padding=0))
    att_L = _lrelu(_conv(sd, p + 'sAtt_L2',
                         torch.cat([F.max_pool2d(att_L, 3, 2, 1), F.avg_pool2d(att_L, 3, 2, 1)], 1)))
    att_L = _up2(_lrelu(_conv(sd, p + 'sAtt_L3', att_L)))
    att = _lrelu(_conv(sd, p + 'sAtt_3', att)) + att_L
    att = _up2(_lrelu(_conv(sd, p + 'sAtt_4', att, padding=0)))
    att = _conv(sd, p + 'sAtt_5', att)
    att_add = _conv(sd, p + 'sAtt_add_2', _lrelu(_conv(sd, p + 'sAtt_add_1', att, padding=0)), padding=0)
    return fea * torch.sigmoid(att) * 2 + att_add


def _first(sd, p, x, HR_in):
    """EDVR_arch.py:45-50 / :266-272: conv_first, or the three-conv 4x down-sampling stem of HR-sized inputs."""
    if HR_in:
        x = _lrelu(_conv(sd, p + 'conv_first_1', x))
        x = _lrelu(_conv(sd, p + 'conv_first_2', x, stride=2))
        return _lrelu(_conv(sd, p + 'conv_first_3', x, stride=2))
    return _lrelu(_conv(sd, p + 'conv_first', x))


def predeblur(sd, p, x, HR_in):
    """Predeblur_ResNet_Pyramid.forward, EDVR_arch.py:43-57."""
    L1 = _first(sd, p, x, HR_in)
    L2 = _lrelu(_conv(sd, p + 'deblur_L2_conv', L1, stride=2))
    L3 = _lrelu(_conv(sd, p + 'deblur_L3_conv', L2, stride=2))
    L3 = _up2(res_block(sd, p + 'RB_L3_1', L3))
    L2 = res_block(sd, p + 'RB_L2_1', L2) + L3
    L2 = _up2(res_block(sd, p + 'RB_L2_2', L2))
    L1 = res_block(sd, p + 'RB_L1_2', res_block(sd, p + 'RB_L1_1', L1)) + L2
    return res_block(sd, p + 'RB_L1_5', res_block(sd, p + 'RB_L1_4', res_block(sd, p + 'RB_L1_3', L1)))


def edvr_forward(sd, x, nf=64, nframes=5, groups=8, front_RBs=5, back_RBs=10, center=None, scale=4,
                 return_intermediates=False, predeblur_=False, HR_in=False, w_TSA=True):
    """EDVR_arch.py:254-313.  Every shipped YML uses predeblur=False, HR_in=False, w_TSA=True; the other variants of the
    constructor (:208-239) are restated too."""
    B, N, C, H, W = x.shape
    center = N // 2 if center is None else center
    inter = {}
    if predeblur_:
        f1 = _conv(sd, 'conv_1x1', predeblur(sd, 'pre_deblur.', x.reshape(-1, C, H, W), HR_in), padding=0)
    else:
        f1 = _first(sd, '', x.reshape(-1, C, H, W), HR_in)
    if HR_in:
        H, W = H // 4, W // 4
    for i in range(front_RBs):
        f1 = res_block(sd, 'feature_extraction.%d' % i, f1)
    f2 = _lrelu(_conv(sd, 'fea_L2_conv2', _lrelu(_conv(sd, 'fea_L2_conv1', f1, stride=2))))
    f3 = _lrelu(_conv(sd, 'fea_L3_conv2', _lrelu(_conv(sd, 'fea_L3_conv1', f2, stride=2))))
    f1 = f1.view(B, N, -1, H, W)
    f2 = f2.view(B, N, -1, H // 2, W // 2)
    f3 = f3.view(B, N, -1, H // 4, W // 4)
    inter['L1_fea'], inter['L2_fea'], inter['L3_fea'] = f1, f2, f3
    ref = [f1[:, center], f2[:, center], f3[:, center]]
    aligned = torch.stack([pcd_align(sd, 'pcd_align.', [f1[:, i], f2[:, i], f3[:, i]], ref, groups)
                           for i in range(N)], 1)
    inter['aligned'] = aligned
    if w_TSA:
        fea = tsa_fusion(sd, 'tsa_fusion.', aligned, center)
    else:
        fea = _conv(sd, 'tsa_fusion', aligned.reshape(B, -1, H, W), padding=0)          # :299-301
    inter['tsa'] = fea
    out = fea
    for i in range(back_RBs):
        out = res_block(sd, 'recon_trunk.%d' % i, out)
    inter['trunk'] = out
    if scale == 4:
        out = _lrelu(F.pixel_shuffle(_conv(sd, 'upconv1', out), 2))
    out = _lrelu(F.pixel_shuffle(_conv(sd, 'upconv2', out), 2))
    out = _lrelu(_conv(sd, 'HRconv', out))
    out = _conv(sd, 'conv_last', out)
    if HR_in:
        out = out + x[:, center]                                                         # :308-309
    else:
        out = out + F.interpolate(x[:, center], scale_factor=scale, mode='bilinear', align_corners=False)
    if return_intermediates:
        return out, inter
    return out


def mfdn_forward(sd, x, scale=4):
    """LRimg_estimator.py:92-117.  x: [B, C, T, H, W] -> [B, C, T, H/scale, W/scale]."""
    B, C, T, H, W = x.shape
    m = x.mean(-1, keepdim=True).mean(-2, keepdim=True)
    x = x - m
    x = _lrelu(F.conv3d(F.pad(x, (1,) * 6, mode='replicate'), sd['conv0.weight'], sd['conv0.bias']))
    fea = x.transpose(1, 2).reshape(B * T, -1, H, W)
    rp = lambda t: F.pad(t, (1, 1, 1, 1), mode='reflect')
    fea = _lrelu(F.conv2d(rp(fea), sd['conv1.weight'], sd['conv1.bias']))
    fea = _lrelu(F.conv2d(rp(fea), sd['conv2.weight'], sd['conv2.bias'], stride=2))
    s3 = 2 if scale == 4 else 1
    fea = _lrelu(F.conv2d(rp(fea), sd['conv3.weight'], sd['conv3.bias'], stride=s3))
    fea = _lrelu(F.conv2d(rp(fea), sd['conv4.weight'], sd['conv4.bias']))
    fea = fea.reshape(B, T, -1, H // scale, W // scale).transpose(1, 2)
    fea = _lrelu(F.conv3d(F.pad(fea, (1,) * 6, mode='replicate'), sd['conv5.weight'], sd['conv5.bias']))
    fea = fea.transpose(1, 2).reshape(B * T, -1, H // scale, W // scale)
    fea = F.conv2d(fea, sd['conv6.weight'], sd['conv6.bias'])
    fea = fea.reshape(B, T, -1, H // scale, W // scale).transpose(1, 2)
    return fea + m


def sfdn_forward(sd, x):
    """SFDN, LRimg_estimator.py:38-67 (DirectKernelEstimator_CMS).  x: [N, 3, H, W] -> [N, 3, H/2, W/2]."""
    m = x.mean(2, keepdim=True).mean(3, keepdim=True)
    rp = lambda t: F.pad(t, (1, 1, 1, 1), mode='reflect')
    fea = x - m
    for i in range(6):
        fea = _lrelu(F.conv2d(rp(fea), sd['conv%d.weight' % i], sd['conv%d.bias' % i], stride=2 if i == 3 else 1))
    return F.conv2d(fea, sd['conv6.weight'], sd['conv6.bias']) + m


def pixel_loss(kind, a, b):
    """Video_base_model.py:39-50, loss.py:5-30."""
    if kind == 'l1':
        return (a - b).abs().mean()
    if kind == 'l2':
        return ((a - b) ** 2).mean()
    if kind == 'cb':
        d = a - b
        return torch.sqrt(d * d + 1e-6).mean()
    if kind == 'huber':                       # loss.py:5-17, delta = 1e-2
        delta = 1e-2
        diff = (a - b).abs()
        return torch.where(diff > delta, delta * diff - 0.5 * delta * delta, 0.5 * diff * diff).mean()
    raise NotImplementedError(kind)


def adapt_and_infer(sd_G, sd_E, sd_E_fixed, lr_clip, steps=2, lr_alpha=1e-5, optimizer='SGD',
                    betas=(0.9, 0.99), criterion='l2', slr_weight=10.0, scale=4, edvr_cfg=None,
                    return_losses=False, slr_given=None, patches=None, patch_size=None):
    """test_dynavsr.py:208-283 for one output frame.

    slr_given ([1, N, 3, H/s, W/s]): the ``train.use_real`` branch (:218-221,243-244) -- the dataset's pre-generated super-LR
    clip replaces MFDN(LR) and only EDVR is optimised (the frozen-MFDN L1 term stays in the loss value, :267-274).
    patches (per step, a list of (py, px) in SLR pixels) + patch_size: the ``maml.use_patch`` branch (:118-145,255-260) -- the
    pixel loss is taken on SLR crops of (patch_size // 2)^2 and the matching LR crops (preprocessing.common_crop :57-85).

    lr_clip: [1, N, 3, H, W] LR window.  Returns the adapted HR estimate [1, 3, sH, sW]
    (and the per-step losses).  sd_* are not modified (the reference deep-copies, :208).
    """
    cfg = dict(edvr_cfg or {})
    pG = {k: v.detach().clone().requires_grad_(True) for k, v in sd_G.items()}
    pE = {k: v.detach().clone().requires_grad_(True) for k, v in sd_E.items()}
    params = list(pG.values()) + (list(pE.values()) if slr_given is None else [])       # :213-221
    if optimizer == 'SGD':
        opt = torch.optim.SGD(params, lr=lr_alpha)
    elif optimizer == 'Adam':
        opt = torch.optim.Adam(params, lr=lr_alpha, betas=betas)
    else:
        raise NotImplementedError(optimizer)
    center = lr_clip.shape[1] // 2
    gt = lr_clip[:, center]
    xin = lr_clip.transpose(1, 2)                                     # B C T H W (LRestimator_model.py:103)
    with torch.no_grad():
        slr_fixed = mfdn_forward(sd_E_fixed, xin, scale).transpose(1, 2)   # :267-270 (constant over steps)
    losses = []
    for k in range(steps):
        slr = mfdn_forward(pE, xin, scale).transpose(1, 2) if slr_given is None else slr_given      # :238-244
        opt.zero_grad()
        if patches is not None:                                            # :255-260 crop() -> common_crop
            q = patch_size // 2
            s_in = torch.stack([slr[0, :, :, py:py + q, px:px + q] for py, px in patches[k]])
            s_gt = torch.stack([gt[0, :, scale * py:scale * (py + q), scale * px:scale * (px + q)] for py, px in patches[k]])
            loss = pixel_loss(criterion, edvr_forward(pG, s_in, scale=scale, **cfg), s_gt)
        else:
            sr = edvr_forward(pG, slr, scale=scale, **cfg)                 # :262-264
            loss = pixel_loss(criterion, sr, gt)
        loss = loss + slr_weight * F.l1_loss(slr, slr_fixed)               # :274
        loss.backward()                                                    # :276
        opt.step()                                                         # :277
        losses.append(float(loss.detach()))
    with torch.no_grad():
        out = edvr_forward({k: v.detach() for k, v in pG.items()}, lr_clip, scale=scale, **cfg)   # :282-283
    if return_losses:
        return out, losses, {k: v.detach() for k, v in pG.items()}, {k: v.detach() for k, v in pE.items()}
    return out
