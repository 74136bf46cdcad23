"""Drop-in for the reference's pybind11 extension module ``deform_conv_cuda`` (codes/models/archs/dcn/src/
deform_conv_cuda.cpp:681-695): the same five function names with the same positional tensor / int arguments, implemented on
the C ABI of libdvsr_b200.so (``dvsr_mdcn_forward_nchw`` / ``dvsr_mdcn_backward_nchw``).  A caller that does
``from . import deform_conv_cuda`` (deform_conv.py:10) keeps working unchanged.

Reference conventions kept: tensors are NCHW, fp32, contiguous (``AT_CHECK(is_contiguous)``, .cpp:493-494 -> RuntimeError
here); outputs and gradients are CALLER-allocated (deform_conv.py:114,128-132) and written in place; gradients are
ACCUMULATED into the caller's (zero-initialised) tensors as the reference's ``addmm_`` / ``atomicAdd`` do; the scratch
arguments (``columns``, ``ones``, ``bufs``) are accepted and ignored -- nothing like the 132.7 MB ``columns`` buffer
exists on this path; CUDA launch errors raise instead of being printed (.cu:793-797).  groups must be 1 and stride /
padding / dilation square (every EDVR use; anything else raises NotImplementedError).
"""
import ctypes

import torch

from .... import _lib
from ...._lib import call


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(*ts):
    for t in ts:
        if not t.is_cuda:
            raise NotImplementedError('deform_conv_cuda: CUDA tensors only (no CPU fallback)')
        if not t.is_contiguous():
            raise RuntimeError('input tensor has to be contiguous')
        if t.dtype != torch.float32:
            raise TypeError('deform_conv_cuda: float32 only, got %s' % t.dtype)


def _square(a, b, what):
    if a != b:
        raise NotImplementedError('deform_conv_cuda: %s must be equal along H and W (got %d, %d)' % (what, a, b))
    return a


_ws = {}


def _workspace(nbytes, device):
    ws = _ws.get(device.index)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws[device.index] = ws
    return ws


def _fwd(input, weight, bias, offset, mask, output, kh, kw, stride, pad, dil, group, dg):
    B, C, H, W = input.shape
    Co = weight.shape[0]
    if weight.shape[2] != kh or weight.shape[3] != kw:
        raise RuntimeError('Input shape and kernel shape wont match: (%d x %d vs %d x %d).' % (kh, kw, weight.shape[2], weight.shape[3]))
    if C != weight.shape[1] * group:
        raise RuntimeError('Input shape and kernel channels wont match: (%d vs %d).' % (C, weight.shape[1] * group))
    Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    if tuple(output.shape) != (B, Co, Ho, Wo):
        output.resize_(B, Co, Ho, Wo)                     # the reference views / resizes the caller's tensor (.cpp:525)
    ws = _workspace(_lib.lib().dvsr_mdcn_workspace_bytes(B, C, H, W, Co, kh, kw, stride, stride, pad, pad, dil, dil, dg, 0), input.device)
    call('dvsr_mdcn_forward_nchw', _p(input), _p(offset), _p(mask), _p(weight), _p(bias), _p(output), B, C, H, W, Co, kh, kw,
         stride, stride, pad, pad, dil, dil, group, dg, _p(ws), ws.numel(), _stream())
    return Ho, Wo


def _bwd(input, weight, offset, mask, grad_output, kh, kw, stride, pad, dil, group, dg, want_bias):
    B, C, H, W = input.shape
    Co = weight.shape[0]
    gi, go, gm = torch.empty_like(input), torch.empty_like(offset), torch.empty_like(mask)
    gw = torch.empty_like(weight)
    gb = input.new_empty(Co) if want_bias else None
    ws = _workspace(_lib.lib().dvsr_mdcn_workspace_bytes(B, C, H, W, Co, kh, kw, stride, stride, pad, pad, dil, dil, dg, 1), input.device)
    call('dvsr_mdcn_backward_nchw', _p(input), _p(offset), _p(mask), _p(weight), _p(grad_output), _p(gi), _p(go), _p(gm),
         _p(gw), _p(gb), B, C, H, W, Co, kh, kw, stride, stride, pad, pad, dil, dil, group, dg, _p(ws), ws.numel(), _stream())
    return gi, go, gm, gw, gb


# ---- DCNv2 (deform_conv_cuda.cpp:486-492, :566-573) ---------------------------------------------------------------------
def modulated_deform_conv_cuda_forward(input, weight, bias, ones, offset, mask, output, columns, kernel_h, kernel_w,
                                       stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group,
                                       with_bias):
    _check(input, weight, offset, mask, output)
    _fwd(input, weight, bias if with_bias else None, offset, mask, output, kernel_h, kernel_w,
         _square(stride_h, stride_w, 'stride'), _square(pad_h, pad_w, 'padding'), _square(dilation_h, dilation_w, 'dilation'),
         group, deformable_group)


def modulated_deform_conv_cuda_backward(input, weight, bias, ones, offset, mask, columns, grad_input, grad_weight, grad_bias,
                                        grad_offset, grad_mask, grad_output, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                                        pad_w, dilation_h, dilation_w, group, deformable_group, with_bias):
    _check(input, weight, offset, mask, grad_output, grad_input, grad_weight, grad_offset, grad_mask)
    gi, go, gm, gw, gb = _bwd(input, weight, offset, mask, grad_output, kernel_h, kernel_w,
                              _square(stride_h, stride_w, 'stride'), _square(pad_h, pad_w, 'padding'),
                              _square(dilation_h, dilation_w, 'dilation'), group, deformable_group, bool(with_bias))
    grad_input.add_(gi)
    grad_offset.add_(go)
    grad_mask.add_(gm)
    grad_weight.add_(gw)
    if with_bias:
        grad_bias.add_(gb)


# ---- DCNv1 (deform_conv_cuda.cpp:151-484): the modulated kernels with a mask of ones ------------------------------------
def _ones_mask(input, weight, kH, kW, stride, pad, dil, dg):
    B, _, H, W = input.shape
    Ho = (H + 2 * pad - (dil * (kH - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kW - 1) + 1)) // stride + 1
    return input.new_ones((B, dg * kH * kW, Ho, Wo))


def deform_conv_forward_cuda(input, weight, offset, output, columns, ones, kW, kH, dW, dH, padW, padH, dilationW, dilationH,
                             group, deformable_group, im2col_step):
    _check(input, weight, offset, output)
    s, p, d = _square(dH, dW, 'stride'), _square(padH, padW, 'padding'), _square(dilationH, dilationW, 'dilation')
    _fwd(input, weight, None, offset, _ones_mask(input, weight, kH, kW, s, p, d, deformable_group), output, kH, kW, s, p, d,
         group, deformable_group)
    return 1


def deform_conv_backward_input_cuda(input, offset, gradOutput, gradInput, gradOffset, weight, columns, kW, kH, dW, dH, padW,
                                    padH, dilationW, dilationH, group, deformable_group, im2col_step):
    _check(input, weight, offset, gradOutput, gradInput, gradOffset)
    s, p, d = _square(dH, dW, 'stride'), _square(padH, padW, 'padding'), _square(dilationH, dilationW, 'dilation')
    gi, go, _, _, _ = _bwd(input, weight, offset, _ones_mask(input, weight, kH, kW, s, p, d, deformable_group), gradOutput,
                           kH, kW, s, p, d, group, deformable_group, False)
    gradInput.add_(gi)
    gradOffset.add_(go)
    return 1


def deform_conv_backward_parameters_cuda(input, offset, gradOutput, gradWeight, columns, ones, kW, kH, dW, dH, padW, padH,
                                         dilationW, dilationH, group, deformable_group, scale, im2col_step):
    _check(input, offset, gradOutput, gradWeight)
    s, p, d = _square(dH, dW, 'stride'), _square(padH, padW, 'padding'), _square(dilationH, dilationW, 'dilation')
    _, _, _, gw, _ = _bwd(input, gradWeight.new_zeros(gradWeight.shape), offset,
                          _ones_mask(input, gradWeight, kH, kW, s, p, d, deformable_group), gradOutput, kH, kW, s, p, d, group,
                          deformable_group, False)
    gradWeight.add_(gw, alpha=float(scale))
    return 1
