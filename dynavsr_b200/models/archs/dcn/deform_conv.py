"""Mirror of codes/models/archs/dcn/deform_conv.py:15-291 (deformable convolution: DCNv2 = modulated, on the EDVR hot
path; DCNv1 = the same sampling without the modulation mask, exported by the reference but used by no YML).

Two entry levels:
  * ``modulated_deform_conv(input, offset, mask, weight, bias, stride, padding, dilation, groups,
    deformable_groups)`` -- the reference operator on NCHW tensors (deform_conv.py:99-100), through
    ``dvsr_mdcn_forward_nchw`` / ``dvsr_mdcn_backward_nchw``;
  * ``ModulatedDeformConvPack`` -- same constructor, initialisation and ``state_dict`` keys as the
    reference (deform_conv.py:221-272).  Inside EDVR it is driven through ``forward_nhwc`` which keeps
    everything channels-last and fuses chunk/cat/sigmoid into the offset-conv epilogue.
"""
import ctypes
import logging
import math

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from .... import ops
from ...._lib import call

logger = logging.getLogger('base')


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_workspaces = {}


def _workspace(nbytes, device):
    """Persistent caller-owned scratch (the reference re-allocates and memsets 132.7 MB per call,
    deform_conv_cuda.cpp:525-529)."""
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


class ModulatedDeformConvFunction(Function):
    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                deformable_groups=1):
        if not input.is_cuda:
            raise NotImplementedError
        ctx.cfg = (stride, padding, dilation, groups, deformable_groups)
        ctx.with_bias = bias is not None
        input, offset, mask, weight = (t.contiguous() for t in (input, offset, mask, weight))
        B, C, H, W = input.shape
        Co, _, kh, kw = weight.shape
        Ho = (H + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
        Wo = (W + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
        if C != weight.shape[1] * groups:
            raise RuntimeError('Input shape and kernel channels wont match: (%d vs %d).' % (C, weight.shape[1] * groups))
        if offset.shape != (B, 2 * deformable_groups * kh * kw, Ho, Wo) or \
                mask.shape != (B, deformable_groups * kh * kw, Ho, Wo):
            raise RuntimeError('offset/mask shapes %s %s do not match the output size' % (tuple(offset.shape), tuple(mask.shape)))
        output = input.new_empty((B, Co, Ho, Wo))
        nbytes = ops._lib.lib().dvsr_mdcn_workspace_bytes(B, C, H, W, Co, kh, kw, stride, stride, padding, padding, dilation,
                                                          dilation, deformable_groups, 0)
        ws = _workspace(nbytes, input.device)
        call('dvsr_mdcn_forward_nchw', _p(input), _p(offset), _p(mask), _p(weight), _p(bias), _p(output),
             B, C, H, W, Co, kh, kw, stride, stride, padding, padding, dilation, dilation, groups, deformable_groups, _p(ws), ws.numel(),
             _stream())
        if weight.requires_grad or mask.requires_grad or offset.requires_grad or input.requires_grad:
            ctx.save_for_backward(input, offset, mask, weight)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        input, offset, mask, weight = ctx.saved_tensors
        stride, padding, dilation, groups, dg = ctx.cfg
        B, C, H, W = input.shape
        Co, _, kh, kw = weight.shape
        grad_output = grad_output.contiguous()
        grad_input, grad_offset, grad_mask = torch.empty_like(input), torch.empty_like(offset), torch.empty_like(mask)
        grad_weight = torch.empty_like(weight)
        grad_bias = input.new_empty(Co) if ctx.with_bias else None
        nbytes = ops._lib.lib().dvsr_mdcn_workspace_bytes(B, C, H, W, Co, kh, kw, stride, stride, padding, padding, dilation, dilation, dg, 1)
        ws = _workspace(nbytes, input.device)
        call('dvsr_mdcn_backward_nchw', _p(input), _p(offset), _p(mask), _p(weight), _p(grad_output),
             _p(grad_input), _p(grad_offset), _p(grad_mask), _p(grad_weight), _p(grad_bias),
             B, C, H, W, Co, kh, kw, stride, stride, padding, padding, dilation, dilation, groups, dg, _p(ws), ws.numel(), _stream())
        return (grad_input, grad_offset, grad_mask, grad_weight, grad_bias, None, None, None, None, None)


modulated_deform_conv = ModulatedDeformConvFunction.apply


def _one(v, what):
    a, b = _pair(v)
    if a != b:
        raise NotImplementedError('%s must be the same along H and W (got %s)' % (what, (a, b)))
    return a


class DeformConvFunction(Function):
    """DCNv1 (deform_conv.py:15-94): y = W * bilinear(x, p + dk + offset).  Runs on the modulated kernels with a mask of
    ones -- the sampling rule of deformable_im2col (deform_conv_cuda_kernel.cu:189-262) is the modulated one
    (:569-632) without the multiply; the offset layout [dg][k][(dh, dw)] is the same.  ``im2col_step`` only sets the
    reference's scratch batching and is accepted for signature parity."""

    @staticmethod
    def forward(ctx, input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1, im2col_step=64):
        if input is not None and input.dim() != 4:
            raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
        if not input.is_cuda:
            raise NotImplementedError
        cur_im2col_step = min(im2col_step, input.shape[0])
        assert (input.shape[0] % cur_im2col_step) == 0, 'im2col step must divide batchsize'
        stride, padding, dilation = _one(stride, 'stride'), _one(padding, 'padding'), _one(dilation, 'dilation')
        kh, kw = weight.shape[2:]
        B = input.shape[0]
        Ho = (input.shape[2] + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
        Wo = (input.shape[3] + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
        if Ho <= 0 or Wo <= 0:
            raise ValueError("convolution input is too small (output would be {})".format('x'.join(map(str, (B, weight.shape[0], Ho, Wo)))))
        mask = input.new_ones((B, deformable_groups * kh * kw, Ho, Wo))
        ctx.cfg = (stride, padding, dilation, groups, deformable_groups)
        ctx.save_for_backward(input, offset, weight, mask)
        with torch.no_grad():
            return ModulatedDeformConvFunction.apply(input, offset, mask, weight, None, stride, padding, dilation, groups,
                                                     deformable_groups)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        input, offset, weight, mask = ctx.saved_tensors
        stride, padding, dilation, groups, dg = ctx.cfg
        B, C, H, W = input.shape
        Co, _, kh, kw = weight.shape
        input, offset, weight, grad_output = (t.contiguous() for t in (input, offset, weight, grad_output))
        grad_input, grad_offset, grad_mask = torch.empty_like(input), torch.empty_like(offset), torch.empty_like(mask)
        grad_weight = torch.empty_like(weight)
        nbytes = ops._lib.lib().dvsr_mdcn_workspace_bytes(B, C, H, W, Co, kh, kw, stride, stride, padding, padding, dilation, dilation, dg, 1)
        ws = _workspace(nbytes, input.device)
        call('dvsr_mdcn_backward_nchw', _p(input), _p(offset), _p(mask), _p(weight), _p(grad_output),
             _p(grad_input), _p(grad_offset), _p(grad_mask), _p(grad_weight), None,
             B, C, H, W, Co, kh, kw, stride, stride, padding, padding, dilation, dilation, groups, dg, _p(ws), ws.numel(), _stream())
        return (grad_input, grad_offset, grad_weight, None, None, None, None, None, None)


deform_conv = DeformConvFunction.apply


class DeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=False):
        super(DeformConv, self).__init__()
        assert not bias
        assert in_channels % groups == 0, 'in_channels {} cannot be divisible by groups {}'.format(in_channels, groups)
        assert out_channels % groups == 0, 'out_channels {} cannot be divisible by groups {}'.format(out_channels, groups)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // self.groups, *self.kernel_size))
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)

    def forward(self, x, offset):
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                           self.deformable_groups)


class DeformConvPack(DeformConv):
    def __init__(self, *args, **kwargs):
        super(DeformConvPack, self).__init__(*args, **kwargs)
        self.conv_offset = nn.Conv2d(self.in_channels, self.deformable_groups * 2 * self.kernel_size[0] * self.kernel_size[1],
                                     kernel_size=self.kernel_size, stride=_pair(self.stride), padding=_pair(self.padding), bias=True)
        self.init_offset()

    def init_offset(self):
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()

    def forward(self, x):
        m = self.conv_offset
        offset = ops.to_nchw(ops.conv(ops.to_nhwc(x), m.weight, m.bias, stride=m.stride[0], pad=m.padding[0]))
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                           self.deformable_groups)


class ModulatedDeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=True):
        super(ModulatedDeformConv, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.with_bias = bias
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, offset, mask):
        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding,
                                     self.dilation, self.groups, self.deformable_groups)


class ModulatedDeformConvPack(ModulatedDeformConv):
    # device-side accumulator of sum|offset| / count for the "Offset mean ... larger than 100" warning
    # (deform_conv.py:285-287) -- read once per forward by the caller instead of a host sync per call.
    check_offsets = False

    def __init__(self, *args, extra_offset_mask=False, **kwargs):
        super(ModulatedDeformConvPack, self).__init__(*args, **kwargs)
        self.extra_offset_mask = extra_offset_mask
        self.conv_offset_mask = nn.Conv2d(
            self.in_channels, self.deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1],
            kernel_size=self.kernel_size, stride=_pair(self.stride), padding=_pair(self.padding), bias=True)
        self.init_offset()

    def init_offset(self):
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()

    # ---- reference-shaped entry (NCHW), deform_conv.py:274-291
    def forward(self, x):
        if self.extra_offset_mask:
            feat, x = x[1], x[0]
        else:
            feat = x
        y = self.forward_nhwc(ops.to_nhwc(x), ops.to_nhwc(feat), host_check=True)
        return ops.to_nchw(y)

    # ---- channels-last entry used inside EDVR
    def forward_nhwc(self, x, feat, act=ops.ACT_NONE, slope=0.1, host_check=False, stats=None):
        K = self.kernel_size[0] * self.kernel_size[1]
        n_off = 2 * self.deformable_groups * K
        # chunk/cat (a no-op re-ordering: offset = out[:, :2*dg*K]) and sigmoid(mask) are fused into the epilogue
        om = ops.conv(feat, self.conv_offset_mask.weight, self.conv_offset_mask.bias, stride=self.stride,
                      pad=self.padding, act=ops.ACT_SIGMOID_SPLIT, sig_split=n_off)
        if host_check or stats is not None:
            acc = stats if stats is not None else torch.zeros(2, device=x.device)
            ops.abs_sum(om.detach(), 0, n_off, acc)
            if host_check:
                offset_mean = float(acc[0]) / (om.numel() // om.shape[3] * n_off)
                if offset_mean > 100:
                    logger.warning('Offset mean is {}, larger than 100.'.format(offset_mean))
        return ops.mdcn(x, om, self.weight, self.bias, self.deformable_groups, self.stride, self.padding,
                        self.dilation, act, slope)
