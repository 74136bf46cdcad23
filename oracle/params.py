"""oracle/params.py -- TEST INFRASTRUCTURE ONLY.

Parameter inventories (reference ``state_dict`` key -> shape) for EDVR and MFDN and a
deterministic, module-independent weight generator, so that golden fixtures can be
regenerated without depending on ``nn.Module`` construction order.

Key names / shapes follow (and are asserted key-for-key against, in oracle/make_golden.py)
  codes/models/archs/EDVR_arch.py:60-93,131-161,206-252
  codes/models/archs/arch_util.py:40-46
  codes/models/archs/dcn/deform_conv.py:221-272
  codes/models/archs/LRimg_estimator.py:70-90
"""
import math
from collections import OrderedDict

import torch


def _conv(d, name, co, ci, k, k2=None):
    d[name + '.weight'] = (co, ci, k, k if k2 is None else k2)
    d[name + '.bias'] = (co,)


def edvr_param_shapes(nf=64, nframes=5, groups=8, front_RBs=5, back_RBs=10, scale=4, predeblur=False, HR_in=False,
                      w_TSA=True):
    d = OrderedDict()

    def first(prefix):                      # EDVR_arch.py:21-26 / :224-229
        if HR_in:
            _conv(d, prefix + 'conv_first_1', nf, 3, 3)
            _conv(d, prefix + 'conv_first_2', nf, nf, 3)
            _conv(d, prefix + 'conv_first_3', nf, nf, 3)
        else:
            _conv(d, prefix + 'conv_first', nf, 3, 3)

    if predeblur:                           # Predeblur_ResNet_Pyramid, EDVR_arch.py:13-39, + conv_1x1 :221
        first('pre_deblur.')
        for rb in ('RB_L1_1', 'RB_L1_2', 'RB_L1_3', 'RB_L1_4', 'RB_L1_5', 'RB_L2_1', 'RB_L2_2', 'RB_L3_1'):
            _conv(d, 'pre_deblur.%s.conv1' % rb, nf, nf, 3)
            _conv(d, 'pre_deblur.%s.conv2' % rb, nf, nf, 3)
        _conv(d, 'pre_deblur.deblur_L2_conv', nf, nf, 3)
        _conv(d, 'pre_deblur.deblur_L3_conv', nf, nf, 3)
        _conv(d, 'conv_1x1', nf, nf, 1)
    else:
        first('')
    for i in range(front_RBs):
        _conv(d, 'feature_extraction.%d.conv1' % i, nf, nf, 3)
        _conv(d, 'feature_extraction.%d.conv2' % i, nf, nf, 3)
    for n in ('fea_L2_conv1', 'fea_L2_conv2', 'fea_L3_conv1', 'fea_L3_conv2'):
        _conv(d, n, nf, nf, 3)

    def dcn(name):
        d[name + '.weight'] = (nf, nf, 3, 3)
        d[name + '.bias'] = (nf,)
        _conv(d, name + '.conv_offset_mask', groups * 27, nf, 3)

    p = 'pcd_align.'
    _conv(d, p + 'L3_offset_conv1', nf, 2 * nf, 3)
    _conv(d, p + 'L3_offset_conv2', nf, nf, 3)
    dcn(p + 'L3_dcnpack')
    _conv(d, p + 'L2_offset_conv1', nf, 2 * nf, 3)
    _conv(d, p + 'L2_offset_conv2', nf, 2 * nf, 3)
    _conv(d, p + 'L2_offset_conv3', nf, nf, 3)
    dcn(p + 'L2_dcnpack')
    _conv(d, p + 'L2_fea_conv', nf, 2 * nf, 3)
    _conv(d, p + 'L1_offset_conv1', nf, 2 * nf, 3)
    _conv(d, p + 'L1_offset_conv2', nf, 2 * nf, 3)
    _conv(d, p + 'L1_offset_conv3', nf, nf, 3)
    dcn(p + 'L1_dcnpack')
    _conv(d, p + 'L1_fea_conv', nf, 2 * nf, 3)
    _conv(d, p + 'cas_offset_conv1', nf, 2 * nf, 3)
    _conv(d, p + 'cas_offset_conv2', nf, nf, 3)
    dcn(p + 'cas_dcnpack')
    t = 'tsa_fusion.'
    if not w_TSA:                           # EDVR_arch.py:239: a plain 1x1 fusion conv
        _conv(d, 'tsa_fusion', nf, nframes * nf, 1)
    else:
        _conv(d, t + 'tAtt_1', nf, nf, 3)
        _conv(d, t + 'tAtt_2', nf, nf, 3)
        _conv(d, t + 'fea_fusion', nf, nframes * nf, 1)
        _conv(d, t + 'sAtt_1', nf, nframes * nf, 1)
        _conv(d, t + 'sAtt_2', nf, 2 * nf, 1)
        _conv(d, t + 'sAtt_3', nf, nf, 3)
        _conv(d, t + 'sAtt_4', nf, nf, 1)
        _conv(d, t + 'sAtt_5', nf, nf, 3)
        _conv(d, t + 'sAtt_L1', nf, nf, 1)
        _conv(d, t + 'sAtt_L2', nf, 2 * nf, 3)
        _conv(d, t + 'sAtt_L3', nf, nf, 3)
        _conv(d, t + 'sAtt_add_1', nf, nf, 1)
        _conv(d, t + 'sAtt_add_2', nf, nf, 1)
    for i in range(back_RBs):
        _conv(d, 'recon_trunk.%d.conv1' % i, nf, nf, 3)
        _conv(d, 'recon_trunk.%d.conv2' % i, nf, nf, 3)
    if scale == 4:
        _conv(d, 'upconv1', nf * 4, nf, 3)
    _conv(d, 'upconv2', 64 * 4, nf, 3)
    _conv(d, 'HRconv', 64, 64, 3)
    _conv(d, 'conv_last', 3, 64, 3)
    return d


def mfdn_param_shapes(nf=64, in_nc=3, scale=4):
    d = OrderedDict()
    d['conv0.weight'] = (nf, in_nc, 3, 3, 3)
    d['conv0.bias'] = (nf,)
    _conv(d, 'conv1', nf, nf, 3)
    _conv(d, 'conv2', nf * 2, nf, 4)
    _conv(d, 'conv3', nf, nf * 2, 4 if scale == 4 else 3)
    _conv(d, 'conv4', nf, nf, 3)
    d['conv5.weight'] = (nf, nf, 3, 3, 3)
    d['conv5.bias'] = (nf,)
    _conv(d, 'conv6', in_nc, nf, 1)
    return d


def sfdn_param_shapes(nf=64):
    """DirectKernelEstimator_CMS, LRimg_estimator.py:39-50."""
    d = OrderedDict()
    _conv(d, 'conv0', nf, 3, 3)
    _conv(d, 'conv1', nf, nf, 3)
    _conv(d, 'conv2', nf, nf, 3)
    _conv(d, 'conv3', nf * 2, nf, 4)
    _conv(d, 'conv4', nf * 2, nf * 2, 3)
    _conv(d, 'conv5', nf, nf * 2, 3)
    _conv(d, 'conv6', 3, nf, 1)
    return d


def make_params(shapes, seed, residual_scale=0.1, offset_std=0.02, dtype=torch.float32):
    """Deterministic weights: kaiming-like N(0, 2/fan_in) (x``residual_scale`` for residual-block
    convs, as arch_util.py:7-24 does), small random biases, and NON-zero ``conv_offset_mask``
    (the reference zero-initialises it, deform_conv.py:270-272, which would hide gather bugs).
    Values depend only on (key order, shape, seed)."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, shp in shapes.items():
        if k.endswith('.bias'):
            v = torch.randn(shp, generator=g, dtype=torch.float64) * 0.05
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            std = math.sqrt(2.0 / fan_in)
            if 'feature_extraction' in k or 'recon_trunk' in k or '.RB_L' in k:
                std *= residual_scale
            if 'conv_offset_mask' in k:
                std = offset_std
            v = torch.randn(shp, generator=g, dtype=torch.float64) * std
        out[k] = v.to(dtype)
    return out
