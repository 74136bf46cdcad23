"""oracle/precision_study.py -- TEST INFRASTRUCTURE / ANALYSIS ONLY (CPU, plain PyTorch).

How much operand precision does each group of EDVR's plain convolutions need?  The tcgen05 path multiplies bf16 pieces of
fp32 operands (x = x_hi + x_lo, w = w_hi + w_lo) and issues THREE products per MAC (x_hi.w_hi + x_hi.w_lo + x_lo.w_hi); the
kernel is bounded by the tensor pipe + shared-memory operand traffic of those three products.  This script
emulates cheaper schemes per layer group inside the oracle forward and reports the end-to-end relative L2 error of the
output frame against the exact-fp32 oracle (north-star tolerance: 1e-3):

    x3      the three products of today's kernel
    x2w     x_hi.(w_hi + w_lo)            -- activations rounded to bf16, weights kept to 16 bits   (2 products)
    x2x     (x_hi + x_lo).w_hi            -- weights rounded to bf16, activations kept to 16 bits   (2 products)
    x1      x_hi.w_hi                     -- plain bf16                                             (1 product)
    tf32    tf32(x).tf32(w) round-to-nearest (one product at half the bf16 rate)

    python -m oracle.precision_study [H W [seed]]   # default 32 x 32 LR, weight seed 1234 (the goldens')
    python -m oracle.precision_study adapt [adam]   # second table: precision of the INNER adaptation steps (forward and
                                                # backward of the 2 SGD steps) vs the adapted frame; final forward exact
"""
import sys

import torch
import torch.nn.functional as F

from . import edvr_oracle as O
from . import params as P


def _bf16_split(t):
    hi = t.to(torch.bfloat16).to(torch.float32)
    lo = (t - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def _tf32(t):
    # round to nearest, 10 explicit mantissa bits
    i = t.contiguous().view(torch.int32)
    r = ((i + 0x00000FFF + ((i >> 13) & 1)) & ~0x1FFF)
    return r.view(torch.float32)


def emulated_conv(scheme, x, w, b, stride, padding):
    if scheme == 'exact':
        return F.conv2d(x, w, b, stride=stride, padding=padding)
    if scheme == 'tf32':
        return F.conv2d(_tf32(x), _tf32(w), b, stride=stride, padding=padding)
    xh, xl = _bf16_split(x)
    wh, wl = _bf16_split(w)
    conv = lambda a, k: F.conv2d(a.double(), k.double(), None, stride=stride, padding=padding)
    if scheme == 'x3':
        y = conv(xh, wh) + conv(xh, wl) + conv(xl, wh)
    elif scheme == 'x2w':
        y = conv(xh, wh + wl)
    elif scheme == 'x2x':
        y = conv(xh + xl, wh)
    elif scheme == 'x1':
        y = conv(xh, wh)
    else:
        raise ValueError(scheme)
    return (y + (b.double().view(1, -1, 1, 1) if b is not None else 0)).float()      # fp32 accumulate / epilogue


GROUPS = [('front', lambda n: n.startswith(('conv_first', 'feature_extraction', 'fea_L'))),
          ('pcd_offset', lambda n: n.startswith('pcd_align') and 'dcnpack' not in n),
          ('dcn_offset_mask', lambda n: 'conv_offset_mask' in n),
          ('tsa', lambda n: n.startswith('tsa_fusion')),
          ('trunk', lambda n: n.startswith('recon_trunk')),
          ('upconv', lambda n: n.startswith('upconv')),
          ('HRconv', lambda n: n.startswith('HRconv')),
          ('conv_last', lambda n: n.startswith('conv_last'))]


def group_of(name):
    for g, f in GROUPS:
        if f(name):
            return g
    return 'other'


def run(sd, x, plan):
    """plan: {group: scheme}; groups not named run exact."""
    orig = O._conv

    def conv(sd_, name, t, stride=1, padding=1):
        return emulated_conv(plan.get(group_of(name), 'exact'), t, sd_[name + '.weight'], sd_[name + '.bias'], stride, padding)

    O._conv = conv
    try:
        with torch.no_grad():
            return O.edvr_forward(sd, x)
    finally:
        O._conv = orig


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


class _EmuConv(torch.autograd.Function):
    """Convolution whose forward and backward products use emulated operand precision (bias gradient exact)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, padding, fwd, bwd):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, bwd)
        return emulated_conv(fwd, x, w, b, stride, padding)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding, bwd = ctx.cfg
        rnd = {'exact': lambda t: t, 'tf32': _tf32, 'x1': lambda t: _bf16_split(t)[0],
               'x3': lambda t: sum(_bf16_split(t))}[bwd]                      # x3 ~ 16-bit operands (lo.lo dropped: negligible)
        gx = torch.nn.grad.conv2d_input(x.shape, rnd(w), rnd(gy), stride=stride, padding=padding)
        gw = torch.nn.grad.conv2d_weight(rnd(x), w.shape, rnd(gy), stride=stride, padding=padding)
        return gx, gw, gy.sum((0, 2, 3)), None, None, None, None


def adapt_study():
    """Adapted frame (2 SGD steps, L2 + 10 L1(SLR), lr 1e-4 -- the setting of the adapt_sgd2_l2 golden) with the inner steps'
    EDVR convolutions on cheaper schemes; the final forward is exact, so the error isolates what the adaptation loses."""
    g = torch.Generator().manual_seed(7)
    lr_clip = torch.rand(1, 5, 3, 32, 48, generator=g)
    sdG = P.make_params(P.edvr_param_shapes(), seed=1234)
    sdE = P.make_params(P.mfdn_param_shapes(), seed=77)
    sdF = P.make_params(P.mfdn_param_shapes(), seed=78)
    if len(sys.argv) > 2 and sys.argv[2] == 'adam':              # the shipped test YMLs: one Adam step, Charbonnier, lr 1e-5
        kw = dict(steps=1, lr_alpha=1e-5, optimizer='Adam', criterion='cb')
    else:
        kw = dict(steps=2, lr_alpha=1e-4, optimizer='SGD', criterion='l2')
    print('inner loop:', kw)
    ref = O.adapt_and_infer(sdG, sdE, sdF, lr_clip, **kw)
    with torch.no_grad():
        unadapted = O.edvr_forward(sdG, lr_clip)
    print('adapted vs un-adapted frame (how far the adaptation moves the output): %.2e' % rel(ref, unadapted))
    print('| inner steps: forward scheme / backward scheme ; final forward | adapted-frame error vs exact |')
    print('|---|---:|')
    orig = O._conv
    try:
        mixed = {g: 'x3' for g, _ in GROUPS}
        mixed.update(trunk='x1', dcn_offset_mask='x1')
        for fwd, bwd, final in (('x3', 'x3', None), ('x3', 'tf32', None), ('x3', 'x1', None), ('tf32', 'tf32', None),
                                ('x1', 'x1', None), ('x1', 'x1', {g: 'x3' for g, _ in GROUPS}), ('x1', 'x1', mixed)):
            def conv(sd_, name, t, stride=1, padding=1, fwd=fwd, bwd=bwd, final=final):
                if not torch.is_grad_enabled():                               # the final forward
                    if final is None:
                        return orig(sd_, name, t, stride, padding)
                    return emulated_conv(final.get(group_of(name), 'exact'), t, sd_[name + '.weight'], sd_[name + '.bias'],
                                         stride, padding)
                return _EmuConv.apply(t, sd_[name + '.weight'], sd_[name + '.bias'], stride, padding, fwd, bwd)
            O._conv = conv
            out = O.adapt_and_infer(sdG, sdE, sdF, lr_clip, **kw)
            what = 'exact' if final is None else ('x3' if final is not mixed else 'x3, trunk + offset/mask convs x1')
            print('| %s / %s ; final forward %s | %.2e |' % (fwd, bwd, what, rel(out, ref)))
        # the deformable convolutions' own products too: bf16 features and weights while gradients are enabled (the cast's
        # backward rounds the gradients to bf16 as well); offsets and masks stay fp32 as on the GPU
        real_mdcn = O.mdcn_torch
        bf = lambda t: t.to(torch.bfloat16).to(torch.float32)

        def mdcn(x, offset, mask, w, b, *a):
            if torch.is_grad_enabled():
                return real_mdcn(bf(x), offset, mask, bf(w), b, *a)
            return real_mdcn(x, offset, mask, w, b, *a)
        O.mdcn_torch = mdcn
        try:
            out = O.adapt_and_infer(sdG, sdE, sdF, lr_clip, **kw)
        finally:
            O.mdcn_torch = real_mdcn
        print('| x1 / x1 incl. the deformable convolutions\' products ; final forward %s | %.2e |' % (what, rel(out, ref)))
    finally:
        O._conv = orig


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'adapt':
        return adapt_study()
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 32)
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1234
    torch.manual_seed(0)
    sd = P.make_params(P.edvr_param_shapes(), seed=seed)
    x = torch.rand(1, 5, 3, H, W, generator=torch.Generator().manual_seed(7))
    ref = run(sd, x, {})
    names = [g for g, _ in GROUPS]
    print('EDVR-M 4x, %dx%d LR, seeded weights; relative L2 error of the output frame vs exact fp32' % (H, W))
    print('| layers on the cheaper scheme (all others: x3) | x2w | x2x | x1 | tf32 |')
    print('|---|---:|---:|---:|---:|')
    base = {g: 'x3' for g in names}
    print('| none (today: x3 everywhere) | %.2e | | | |' % rel(run(sd, x, base), ref))
    for g in names + ['ALL']:
        row = []
        for scheme in ('x2w', 'x2x', 'x1', 'tf32'):
            plan = dict(base)
            for k in (names if g == 'ALL' else [g]):
                plan[k] = scheme
            row.append('%.2e' % rel(run(sd, x, plan), ref))
        print('| %s | %s |' % (g, ' | '.join(row)))


if __name__ == '__main__':
    main()
