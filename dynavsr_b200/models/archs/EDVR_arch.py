"""Mirror of codes/models/archs/EDVR_arch.py: same classes, constructor arguments, parameter names /
shapes / initialisation (so released ``.pth`` files load with ``strict=True``) and the same
``EDVR.forward([B, N, 3, H, W]) -> [B, 3, sH, sW]`` contract -- but every layer runs in the
hand-written sm_100a kernels of libdvsr_b200.so on channels-last tensors:

  * the 5 per-frame PCD passes (python loop EDVR_arch.py:290-297) run as ONE batch of N frames (the
    weights are shared); the reference feature of the centre frame is read in place through a
    broadcast input segment instead of 18 ``.clone()``s (:287-294);
  * ``torch.cat`` (65 per forward) never materialises: a conv reads its inputs as K-segments;
  * LeakyReLU / ReLU / residual add / sigmoid / PixelShuffle / the final bilinear skip are conv epilogues;
  * ``offset * 2`` is folded into the bilinear up-sampling kernel (:108,:117);
  * TSA's temporal attention and final modulation are single fused kernels.
"""
import functools

import torch.nn as nn

from . import arch_util
from ... import ops
from ...ops import ACT_LRELU, Seg

try:
    from .dcn.deform_conv import ModulatedDeformConvPack as DCN
except ImportError:  # pragma: no cover
    raise ImportError('Failed to import DCNv2 module.')


def _c(x, m, act=ops.ACT_NONE, res=None, stride=None, shuffle=0):
    """Run nn.Conv2d parameter holder ``m`` through the NHWC conv kernel."""
    return ops.conv(x, m.weight, m.bias, stride=m.stride[0] if stride is None else stride, pad=m.padding[0],
                    act=act, slope=0.1, res=res, shuffle=shuffle)


class _Stem(object):
    """Shared by EDVR and Predeblur: ``conv_first`` or, for HR-sized inputs, the three-conv 4x down-sampling stem
    (EDVR_arch.py:21-26,45-50 / :224-229,266-272)."""

    @staticmethod
    def build(mod, nf, HR_in):
        if HR_in:
            mod.conv_first_1 = nn.Conv2d(3, nf, 3, 1, 1, bias=True)
            mod.conv_first_2 = nn.Conv2d(nf, nf, 3, 2, 1, bias=True)
            mod.conv_first_3 = nn.Conv2d(nf, nf, 3, 2, 1, bias=True)
        else:
            mod.conv_first = nn.Conv2d(3, nf, 3, 1, 1, bias=True)

    @staticmethod
    def run(mod, frames, HR_in):
        if HR_in:
            x = _c(frames, mod.conv_first_1, ACT_LRELU)
            x = _c(x, mod.conv_first_2, ACT_LRELU)
            return _c(x, mod.conv_first_3, ACT_LRELU)
        return _c(frames, mod.conv_first, ACT_LRELU)


class Predeblur_ResNet_Pyramid(nn.Module):
    """Pre-deblur pyramid of EDVR_arch.py:13-57 (same parameter names; used by no YML of the reference)."""

    def __init__(self, nf=128, HR_in=False):
        super(Predeblur_ResNet_Pyramid, self).__init__()
        self.HR_in = True if HR_in else False
        _Stem.build(self, nf, self.HR_in)
        basic_block = functools.partial(arch_util.ResidualBlock_noBN, nf=nf)
        for name in ('RB_L1_1', 'RB_L1_2', 'RB_L1_3', 'RB_L1_4', 'RB_L1_5', 'RB_L2_1', 'RB_L2_2', 'RB_L3_1'):
            setattr(self, name, basic_block())
        self.deblur_L2_conv = nn.Conv2d(nf, nf, 3, 2, 1, bias=True)
        self.deblur_L3_conv = nn.Conv2d(nf, nf, 3, 2, 1, bias=True)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)

    def forward(self, x):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))

    def forward_nhwc(self, frames):
        L1 = _Stem.run(self, frames, self.HR_in)
        L2 = _c(L1, self.deblur_L2_conv, ACT_LRELU)
        L3 = _c(L2, self.deblur_L3_conv, ACT_LRELU)
        L3 = ops.upsample(self.RB_L3_1(L3), 2)
        L2 = self.RB_L2_1(L2) + L3
        L2 = ops.upsample(self.RB_L2_2(L2), 2)
        L1 = self.RB_L1_2(self.RB_L1_1(L1)) + L2
        return self.RB_L1_5(self.RB_L1_4(self.RB_L1_3(L1)))


class PCD_Align(nn.Module):
    """Alignment module using Pyramid, Cascading and Deformable convolution with 3 pyramid levels
    (EDVR_arch.py:60-128)."""

    def __init__(self, nf=64, groups=8):
        super(PCD_Align, self).__init__()
        self.L3_offset_conv1 = nn.Conv2d(nf * 2, nf, 3, 1, 1, bias=True)
        self.L3_offset_conv2 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.L3_dcnpack = DCN(nf, nf, 3, stride=1, padding=1, dilation=1, deformable_groups=groups,
                              extra_offset_mask=True)
        self.L2_offset_conv1 = nn.Conv2d(nf * 2, nf, 3, 1, 1, bias=True)
        self.L2_offset_conv2 = nn.Conv2d(nf * 2, nf, 3, 1, 1, bias=True)
        self.L2_offset_conv3 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.L2_dcnpack = DCN(nf, nf, 3, stride=1, padding=1, dilation=1, deformable_groups=groups,
                              extra_offset_mask=True)
        self.L2_fea_conv = nn.Conv2d(nf * 2, nf, 3, 1, 1, bias=True)
        self.L1_offset_conv1 = nn.Conv2d(nf * 2, nf, 3, 1, 1, bias=True)
        self.L1_offset_conv2 = nn.Conv2d(nf * 2, nf, 3, 1, 1, bias=True)
        self.L1_offset_conv3 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.L1_dcnpack = DCN(nf, nf, 3, stride=1, padding=1, dilation=1, deformable_groups=groups,
                              extra_offset_mask=True)
        self.L1_fea_conv = nn.Conv2d(nf * 2, nf, 3, 1, 1, bias=True)
        self.cas_offset_conv1 = nn.Conv2d(nf * 2, nf, 3, 1, 1, bias=True)
        self.cas_offset_conv2 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.cas_dcnpack = DCN(nf, nf, 3, stride=1, padding=1, dilation=1, deformable_groups=groups,
                               extra_offset_mask=True)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)

    def forward(self, nbr_fea_l, ref_fea_l):
        """Reference-shaped entry: lists [L1, L2, L3] of NCHW features -> aligned NCHW feature."""
        nbr = [ops.to_nhwc(t) for t in nbr_fea_l]
        ref = [ops.to_nhwc(t) for t in ref_fea_l]
        return ops.to_nchw(self.forward_nhwc(nbr, ref))

    def forward_nhwc(self, nbr, ref, stats=None):
        """nbr: [L1, L2, L3] NHWC features of all frames to align; ref: matching tensors or ``Seg``s
        (a broadcast Seg when one reference serves several frames)."""
        L = ACT_LRELU
        # L3
        o3 = _c([nbr[2], ref[2]], self.L3_offset_conv1, L)
        o3 = _c(o3, self.L3_offset_conv2, L)
        f3 = self.L3_dcnpack.forward_nhwc(nbr[2], o3, act=L, stats=stats)
        # L2
        o2 = _c([nbr[1], ref[1]], self.L2_offset_conv1, L)
        o2 = _c([o2, ops.upsample(o3, 2, mul=2.0)], self.L2_offset_conv2, L)
        o2 = _c(o2, self.L2_offset_conv3, L)
        f2 = self.L2_dcnpack.forward_nhwc(nbr[1], o2, stats=stats)
        f2 = _c([f2, ops.upsample(f3, 2)], self.L2_fea_conv, L)
        # L1
        o1 = _c([nbr[0], ref[0]], self.L1_offset_conv1, L)
        o1 = _c([o1, ops.upsample(o2, 2, mul=2.0)], self.L1_offset_conv2, L)
        o1 = _c(o1, self.L1_offset_conv3, L)
        f1 = self.L1_dcnpack.forward_nhwc(nbr[0], o1, stats=stats)
        f1 = _c([f1, ops.upsample(f2, 2)], self.L1_fea_conv)
        # cascade
        oc = _c([f1, ref[0]], self.cas_offset_conv1, L)
        oc = _c(oc, self.cas_offset_conv2, L)
        return self.cas_dcnpack.forward_nhwc(f1, oc, act=L, stats=stats)


class TSA_Fusion(nn.Module):
    """Temporal Spatial Attention fusion (EDVR_arch.py:131-203)."""

    def __init__(self, nf=64, nframes=5, center=2):
        super(TSA_Fusion, self).__init__()
        self.center = center
        self.nframes = nframes
        self.tAtt_1 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.tAtt_2 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.fea_fusion = nn.Conv2d(nframes * nf, nf, 1, 1, bias=True)
        self.sAtt_1 = nn.Conv2d(nframes * nf, nf, 1, 1, bias=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.avgpool = nn.AvgPool2d(3, stride=2, padding=1)
        self.sAtt_2 = nn.Conv2d(nf * 2, nf, 1, 1, bias=True)
        self.sAtt_3 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.sAtt_4 = nn.Conv2d(nf, nf, 1, 1, bias=True)
        self.sAtt_5 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.sAtt_L1 = nn.Conv2d(nf, nf, 1, 1, bias=True)
        self.sAtt_L2 = nn.Conv2d(nf * 2, nf, 3, 1, 1, bias=True)
        self.sAtt_L3 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.sAtt_add_1 = nn.Conv2d(nf, nf, 1, 1, bias=True)
        self.sAtt_add_2 = nn.Conv2d(nf, nf, 1, 1, bias=True)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)

    def forward(self, aligned_fea):
        """Reference-shaped entry: [B, N, C, H, W] -> [B, C, H, W]."""
        B, N, C, H, W = aligned_fea.shape
        a = ops.to_nhwc(aligned_fea.reshape(B * N, C, H, W))
        return ops.to_nchw(self.forward_nhwc(a, B, N))

    def forward_nhwc(self, aligned, B, N):
        """aligned: [B*N, H, W, C] -> [B, H, W, C]."""
        L = ACT_LRELU
        _, H, W, C = aligned.shape
        center = aligned.view(B, N, H, W, C)[:, self.center]
        emb_ref = _c(center, self.tAtt_2)
        emb = _c(aligned, self.tAtt_1)
        al = ops.tsa_temporal(aligned, emb, emb_ref, N)            # [B, H, W, N*C]
        fea = _c(al, self.fea_fusion, L)
        att = _c(al, self.sAtt_1, L)
        att = _c(ops.pool_maxavg(att), self.sAtt_2, L)
        att_L = _c(att, self.sAtt_L1, L)
        att_L = _c(ops.pool_maxavg(att_L), self.sAtt_L2, L)
        att_L = ops.upsample(_c(att_L, self.sAtt_L3, L), 2)
        att = _c(att, self.sAtt_3, L, res=att_L)                   # lrelu(sAtt_3(att)) + att_L
        att = ops.upsample(_c(att, self.sAtt_4, L), 2)
        att = _c(att, self.sAtt_5)
        att_add = _c(_c(att, self.sAtt_add_1, L), self.sAtt_add_2)
        return ops.tsa_combine(fea, att, att_add)


class EDVR(nn.Module):
    def __init__(self, nf=64, nframes=5, groups=8, front_RBs=5, back_RBs=10, center=None, predeblur=False,
                 HR_in=False, w_TSA=True, scale=4):
        super(EDVR, self).__init__()
        if scale not in (2, 4):
            raise NotImplementedError('scale must be 2 or 4')
        self.nf = nf
        self.nframes = nframes
        self.center = nframes // 2 if center is None else center
        self.is_predeblur = True if predeblur else False
        self.HR_in = True if HR_in else False
        self.w_TSA = w_TSA
        self.scale = scale
        if not w_TSA and nframes > 5:
            raise NotImplementedError('w_TSA=False reads the aligned frames as K-segments: at most 5 frames')
        ResidualBlock_noBN_f = functools.partial(arch_util.ResidualBlock_noBN, nf=nf)

        if self.is_predeblur:
            self.pre_deblur = Predeblur_ResNet_Pyramid(nf=nf, HR_in=self.HR_in)
            self.conv_1x1 = nn.Conv2d(nf, nf, 1, 1, bias=True)
        else:
            _Stem.build(self, nf, self.HR_in)
        self.feature_extraction = arch_util.make_layer(ResidualBlock_noBN_f, front_RBs)
        self.fea_L2_conv1 = nn.Conv2d(nf, nf, 3, 2, 1, bias=True)
        self.fea_L2_conv2 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.fea_L3_conv1 = nn.Conv2d(nf, nf, 3, 2, 1, bias=True)
        self.fea_L3_conv2 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)

        self.pcd_align = PCD_Align(nf=nf, groups=groups)
        if self.w_TSA:
            self.tsa_fusion = TSA_Fusion(nf=nf, nframes=nframes, center=self.center)
        else:
            self.tsa_fusion = nn.Conv2d(nframes * nf, nf, 1, 1, bias=True)

        self.recon_trunk = arch_util.make_layer(ResidualBlock_noBN_f, back_RBs)
        if self.scale == 4:
            self.upconv1 = nn.Conv2d(nf, nf * 4, 3, 1, 1, bias=True)
        self.upconv2 = nn.Conv2d(nf, 64 * 4, 3, 1, 1, bias=True)
        self.pixel_shuffle = nn.PixelShuffle(2)
        self.HRconv = nn.Conv2d(64, 64, 3, 1, 1, bias=True)
        self.conv_last = nn.Conv2d(64, 3, 3, 1, 1, bias=True)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)
        # device-side [sum |offset|, unused] accumulator for the offset-magnitude warning
        self.offset_stats = None

    def forward(self, x):
        """x: [B, N, 3, H, W] (NCHW frames, as the reference) -> [B, 3, scale*H, scale*W]."""
        B, N, C, H, W = x.size()
        frames = ops.to_nhwc(x.reshape(B * N, C, H, W))
        return ops.to_nchw(self.forward_nhwc(frames, B, N))

    def forward_nhwc(self, frames, B, N):
        """frames: [B*N, H, W, 3] channels-last -> [B, scale*H, scale*W, 3]."""
        L = ACT_LRELU
        _, H, W, C = frames.shape
        need = 16 if self.HR_in else 4
        if H % need or W % need:
            raise RuntimeError('EDVR needs H and W to be multiples of %d, got %dx%d' % (need, H, W))
        # ---- per-frame feature pyramid (EDVR_arch.py:272-283), all N frames batched
        if self.is_predeblur:
            L1 = _c(self.pre_deblur.forward_nhwc(frames), self.conv_1x1)
        else:
            L1 = _Stem.run(self, frames, self.HR_in)
        L1 = self.feature_extraction(L1)
        L2 = _c(_c(L1, self.fea_L2_conv1, L), self.fea_L2_conv2, L)
        L3 = _c(_c(L2, self.fea_L3_conv1, L), self.fea_L3_conv2, L)
        # ---- PCD alignment of every frame against the centre frame (:285-297), one batched pass
        ref = [Seg(t.view(B, N, *t.shape[1:])[:, self.center], T=N, Tsrc=1, t_fixed=0) for t in (L1, L2, L3)]
        aligned = self.pcd_align.forward_nhwc([L1, L2, L3], ref, stats=self.offset_stats)
        # ---- TSA fusion, reconstruction trunk, PixelShuffle head (:299-312)
        if self.w_TSA:
            fea = self.tsa_fusion.forward_nhwc(aligned, B, N)
        else:
            # aligned_fea.view(B, -1, H, W) -> 1x1 conv (:299-301): the N frames are the K-segments of the conv, read in place
            al = aligned.view(B, N, *aligned.shape[1:])
            fea = _c([al[:, i] for i in range(N)], self.tsa_fusion)
        out = self.recon_trunk(fea)
        if self.scale == 4:
            out = _c(out, self.upconv1, L, shuffle=2)   # lrelu commutes with the PixelShuffle permutation
        out = _c(out, self.upconv2, L, shuffle=2)
        out = _c(out, self.HRconv, L)
        x_center = frames.view(B, N, H, W, C)[:, self.center]
        base = x_center.contiguous() if self.HR_in else ops.upsample(x_center, self.scale)     # :308-311
        return _c(out, self.conv_last, res=base)        # conv_last(out) + bilinear(x_center)
