"""oracle/torch_ops.py -- TEST INFRASTRUCTURE ONLY.

Plain-PyTorch fp32 (or fp64) CPU restatement of the modulated deformable convolution used
by the reference, written as dense tensor algebra so that ``torch.autograd`` provides the
backward pass.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline
legs may import this module; the product package (``dynavsr_b200``) never does.

Algorithm restated (not copied) from
  codes/models/archs/dcn/src/deform_conv_cuda_kernel.cu:466-496  bilinear sample with per-corner bounds
  codes/models/archs/dcn/src/deform_conv_cuda_kernel.cu:569-632  modulated im2col (offset [dg][k][dy,dx], mask [dg][k])
  codes/models/archs/dcn/src/deform_conv_cuda.cpp:545-563        out = W . col + bias

Pinned against torchvision.ops.deform_conv2d and oracle/mdcn_oracle.c in tests/test_oracle.py.
"""
import torch


def mdcn_columns(x, offset, mask, kh, kw, stride=1, padding=0, dilation=1, deformable_groups=1):
    """Return the modulated, bilinearly sampled column tensor [B, C, kh*kw, Ho, Wo]."""
    B, C, H, W = x.shape
    K = kh * kw
    dg = deformable_groups
    Ho = (H + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    dt, dev = x.dtype, x.device
    off = offset.reshape(B, dg, K, 2, Ho, Wo)
    msk = mask.reshape(B, dg, K, Ho, Wo)
    ki = torch.arange(K, device=dev) // kw
    kj = torch.arange(K, device=dev) % kw
    base_h = (torch.arange(Ho, device=dev) * stride - padding).to(dt)
    base_w = (torch.arange(Wo, device=dev) * stride - padding).to(dt)
    # sampling positions [B, dg, K, Ho, Wo]
    ph = base_h.view(1, 1, 1, Ho, 1) + (ki * dilation).to(dt).view(1, 1, K, 1, 1) + off[:, :, :, 0]
    pw = base_w.view(1, 1, 1, 1, Wo) + (kj * dilation).to(dt).view(1, 1, K, 1, 1) + off[:, :, :, 1]
    inside = (ph > -1) & (pw > -1) & (ph < H) & (pw < W)          # kernel.cu:617
    h0 = torch.floor(ph)
    w0 = torch.floor(pw)
    lh, lw = ph - h0, pw - w0
    h0, w0 = h0.long(), w0.long()
    h1, w1 = h0 + 1, w0 + 1
    cpg = C // dg
    xg = x.reshape(B, dg, cpg, H * W)
    col = torch.zeros(B, dg, cpg, K, Ho, Wo, dtype=dt, device=dev)
    for hh, ww, wt in ((h0, w0, (1 - lh) * (1 - lw)), (h0, w1, (1 - lh) * lw),
                       (h1, w0, lh * (1 - lw)), (h1, w1, lh * lw)):
        ok = inside & (hh >= 0) & (hh <= H - 1) & (ww >= 0) & (ww <= W - 1)   # kernel.cu:478-490
        idx = (hh.clamp(0, H - 1) * W + ww.clamp(0, W - 1)).reshape(B, dg, 1, K * Ho * Wo)
        v = torch.gather(xg, 3, idx.expand(B, dg, cpg, K * Ho * Wo)).reshape(B, dg, cpg, K, Ho, Wo)
        col = col + v * (wt * ok.to(dt)).unsqueeze(2)
    col = col * msk.unsqueeze(2)
    return col.reshape(B, C, K, Ho, Wo)


def mdcn_torch(x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
               deformable_groups=1):
    """Same signature/semantics as the reference ``modulated_deform_conv`` (deform_conv.py:99-100)."""
    Co, Cg, kh, kw = weight.shape
    B, C = x.shape[:2]
    col = mdcn_columns(x, offset, mask, kh, kw, stride, padding, dilation, deformable_groups)
    Ho, Wo = col.shape[-2:]
    col = col.reshape(B, groups, Cg * kh * kw, Ho * Wo)
    w = weight.reshape(groups, Co // groups, Cg * kh * kw)
    y = torch.einsum('gok,bgkp->bgop', w, col).reshape(B, Co, Ho, Wo)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y
