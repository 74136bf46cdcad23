"""Mirror of codes/models/archs/LRimg_estimator.py:38-117 -- MFDN (``DirectKernelEstimatorVideo``, :70-117), the
multi-frame down-scaling network that maps an LR clip to its "super-LR" version inside the inner
adaptation step, and SFDN (``DirectKernelEstimator_CMS``, :38-67), its single-frame 2x sibling
(options/train/MAML/EDVR/EDVR_{REDS,Vimeo}_SFDN.yml).  Same constructors, parameter names and shapes; kernels from libdvsr_b200.so:
Conv3d = three temporal K-segments of one implicit GEMM over an explicitly replication-padded clip,
reflection pads are explicit (their adjoints are gather kernels), activations are conv epilogues.
"""
import ctypes

import torch
import torch.nn as nn
from torch.autograd import Function

from ... import ops
from ..._lib import call
from ...ops import ACT_LRELU


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _AddFrameMean(Function):
    """y[n, h, w, c] = x[n, h, w, c] + sign * m[n, c].  Gradients: g to x, sign * sum_hw(g) to m -- so that when the input
    clip itself requires grad the mean stays in the graph exactly as the reference's ``x - m ... + m`` does
    (LRimg_estimator.py:58-66,98-100,116); in the hot path the clip is data and neither term is computed."""

    @staticmethod
    def forward(ctx, x, m, sign):
        x = x.contiguous()
        N, H, W, C = x.shape
        y = torch.empty_like(x)
        call('dvsr_add_channel_bias', _p(x), _p(m), _p(y), N, H * W, C, float(sign), _stream())
        ctx.sign = sign
        return y

    @staticmethod
    def backward(ctx, g):
        gm = None
        if ctx.needs_input_grad[1]:
            g = g.contiguous()
            N, H, W, C = g.shape
            gm = torch.empty(N, C, device=g.device, dtype=torch.float32)
            call('dvsr_spatial_mean', _p(g), _p(gm), N, H * W, C, _stream())
            gm = gm * (ctx.sign * H * W)
        return g, gm, None


class _FrameMean(Function):
    """m[n, c] = mean_hw x[n, h, w, c]; backward spreads gm / (H W) over the frame."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        N, H, W, C = x.shape
        m = torch.empty(N, C, device=x.device, dtype=torch.float32)
        call('dvsr_spatial_mean', _p(x), _p(m), N, H * W, C, _stream())
        ctx.shape = (N, H, W, C)
        return m

    @staticmethod
    def backward(ctx, gm):
        N, H, W, C = ctx.shape
        z = torch.zeros(N, H, W, C, device=gm.device, dtype=torch.float32)
        gx = torch.empty_like(z)
        call('dvsr_add_channel_bias', _p(z), _p((gm / (H * W)).contiguous()), _p(gx), N, H * W, C, 1.0, _stream())
        return gx


def frame_mean(frames):
    """Per-frame, per-channel spatial mean [N, C] (LRimg_estimator.py:58,99); differentiable when ``frames`` requires grad."""
    return _FrameMean.apply(frames)


class DirectKernelEstimator_CMS(nn.Module):
    """SFDN (LRimg_estimator.py:38-67): per-image mean removal, six reflection-padded convs (one 4x4 stride 2: the output is
    H/2 x W/2), 1x1 to RGB, mean added back.  ``forward`` takes the reference's [N, 3, H, W]."""

    def __init__(self, nf):
        super(DirectKernelEstimator_CMS, self).__init__()
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)
        self.pad = nn.ReflectionPad2d(1)
        self.conv0 = nn.Conv2d(3, nf, 3, 1, 0, bias=True)
        self.conv1 = nn.Conv2d(nf, nf, 3, 1, 0, bias=True)
        self.conv2 = nn.Conv2d(nf, nf, 3, 1, padding=0, bias=True)
        self.conv3 = nn.Conv2d(nf, nf * 2, 4, 2, padding=0, bias=True)
        self.conv4 = nn.Conv2d(nf * 2, nf * 2, 3, 1, padding=0, bias=True)
        self.conv5 = nn.Conv2d(nf * 2, nf, 3, 1, padding=0, bias=True)
        self.conv6 = nn.Conv2d(nf, 3, 1, stride=1, padding=0, bias=True)
        self.scale = 2

    def forward(self, x):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))

    def forward_nhwc(self, frames, B=None, T=None):
        """frames: [N, H, W, 3] channels-last -> [N, H/2, W/2, 3] (B, T accepted for interface parity with MFDN)."""
        L = ACT_LRELU
        c2 = lambda t, m: ops.conv(ops.pad2d(t, 1, 'reflect'), m.weight, m.bias, stride=m.stride[0], pad=0, act=L)
        m = frame_mean(frames)
        fea = _AddFrameMean.apply(frames, m, -1.0)
        for conv in (self.conv0, self.conv1, self.conv2, self.conv3, self.conv4, self.conv5):
            fea = c2(fea, conv)
        fea = ops.conv(fea, self.conv6.weight, self.conv6.bias, stride=1, pad=0)
        return _AddFrameMean.apply(fea, m, 1.0)


class DirectKernelEstimatorVideo(nn.Module):
    def __init__(self, nf, in_nc=3, scale=2):
        super(DirectKernelEstimatorVideo, self).__init__()
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)
        self.pad3d = nn.ReplicationPad3d(1)
        self.pad = nn.ReflectionPad2d(1)
        self.conv0 = nn.Conv3d(in_nc, nf, 3, 1, 0, bias=True)
        self.conv1 = nn.Conv2d(nf, nf, 3, 1, 0, bias=True)
        self.conv2 = nn.Conv2d(nf, nf * 2, 4, 2, 0, bias=True)
        if scale == 2:
            self.conv3 = nn.Conv2d(nf * 2, nf, 3, 1, 0, bias=True)
        elif scale == 4:
            self.conv3 = nn.Conv2d(nf * 2, nf, 4, 2, 0, bias=True)
        else:
            raise NotImplementedError()
        self.conv4 = nn.Conv2d(nf, nf, 3, 1, 0, bias=True)
        self.conv5 = nn.Conv3d(nf, nf, 3, 1, 0, bias=True)
        self.conv6 = nn.Conv2d(nf, in_nc, 1, 1, 0, bias=True)
        self.scale = scale

    def forward(self, x):
        """x: [B, C, T, H, W] -> [B, C, T, H/scale, W/scale] (reference contract)."""
        B, C, T, H, W = x.shape
        frames = ops.to_nhwc(x.transpose(1, 2).reshape(B * T, C, H, W))
        out = self.forward_nhwc(frames, B, T)
        h, w = out.shape[1], out.shape[2]
        return ops.to_nchw(out).view(B, T, C, h, w).transpose(1, 2)

    def forward_nhwc(self, frames, B, T):
        """frames: [B*T, H, W, C] channels-last LR frames -> [B*T, H/scale, W/scale, C] super-LR frames."""
        L = ACT_LRELU
        c2 = lambda t, m: ops.conv(ops.pad2d(t, 1, 'reflect'), m.weight, m.bias, stride=m.stride[0], pad=0, act=L)
        m = frame_mean(frames)
        x = _AddFrameMean.apply(frames, m, -1.0)
        if ops._tc() and x.shape[3] <= 4 and not x.requires_grad:
            # RGB clip: temporal taps folded into channels, one tensor-core conv (ops.conv3d_rgb)
            x = ops.conv3d_rgb(x, self.conv0.weight, self.conv0.bias, T, act=L)
        else:
            x = ops.conv3d_padded(ops.pad3d_replicate(x, T), self.conv0.weight, self.conv0.bias, T, act=L)
        fea = c2(x, self.conv1)
        fea = c2(fea, self.conv2)
        fea = c2(fea, self.conv3)
        fea = c2(fea, self.conv4)
        fea = ops.conv3d_padded(ops.pad3d_replicate(fea, T), self.conv5.weight, self.conv5.bias, T, act=L)
        fea = ops.conv(fea, self.conv6.weight, self.conv6.bias, stride=1, pad=0)
        return _AddFrameMean.apply(fea, m, 1.0)
