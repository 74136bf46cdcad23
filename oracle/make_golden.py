"""oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY; run in the build container, not on the GPU box.

Generates tests/golden/*.npz by importing the UNMODIFIED reference Python modules from
/root/reference/codes and running them on CPU with the CUDA-only DCN op routed to
``torchvision.ops.deform_conv2d`` (the recipe of SURVEY.md section 8c / Appendix A).  The reference
ships no golden vectors of its own, so these files ARE the parity pin; the
weights are regenerated at test time from (shapes, seed) by oracle/params.py.

    python -m oracle.make_golden            # writes tests/golden/, prints oracle-vs-reference errors
"""
import os
import sys
import types

import numpy as np
import torch
import torchvision

REF = os.environ.get('DVSR_REFERENCE', '/root/reference/codes')
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def import_reference():
    sys.path.insert(0, REF)
    sys.modules['models.archs.dcn.deform_conv_cuda'] = types.ModuleType('deform_conv_cuda_stub')
    import models.archs.dcn  # noqa: F401
    dc = sys.modules['models.archs.dcn.deform_conv']

    def tv_mdcn(x, off, m, w, b, s, p, d, g, dg):
        assert g == 1
        return torchvision.ops.deform_conv2d(x, off, w, b, stride=s, padding=p, dilation=d, mask=m)

    dc.modulated_deform_conv = tv_mdcn
    import models.archs.EDVR_arch as E
    import models.archs.LRimg_estimator as L
    import models.loss as loss
    return E, L, loss, dc


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import edvr_oracle as O
    from oracle import params as P
    from oracle.torch_ops import mdcn_torch
    E, L, loss_mod, dc = import_reference()

    # ---------------- 1. DCN op: forward + all five gradients (torchvision = reference algorithm) ----
    g = torch.Generator().manual_seed(11)
    B, C, H, W, Co, dg = 2, 16, 9, 11, 8, 4
    x = torch.randn(B, C, H, W, generator=g)
    off = torch.randn(B, dg * 18, H, W, generator=g) * 3.0       # spans in-bounds, border and OOB taps
    msk = torch.sigmoid(torch.randn(B, dg * 9, H, W, generator=g))
    w = torch.randn(Co, C, 3, 3, generator=g) * 0.1
    b = torch.randn(Co, generator=g) * 0.1
    gy = torch.randn(B, Co, H, W, generator=g)
    leaves = [t.clone().requires_grad_(True) for t in (x, off, msk, w, b)]
    y = torchvision.ops.deform_conv2d(leaves[0], leaves[1], leaves[3], leaves[4], stride=1, padding=1,
                                      dilation=1, mask=leaves[2])
    grads = torch.autograd.grad(y, leaves, gy)
    y2 = mdcn_torch(x, off, msk, w, b, 1, 1, 1, 1, dg)
    print('dcn: oracle(torch) vs torchvision rel', rel(y2, y.detach()))
    np.savez(os.path.join(GOLD, 'dcn_small.npz'), x=x.numpy(), offset=off.numpy(), mask=msk.numpy(),
             weight=w.numpy(), bias=b.numpy(), gy=gy.numpy(), y=y.detach().numpy(),
             gx=grads[0].numpy(), goffset=grads[1].numpy(), gmask=grads[2].numpy(),
             gweight=grads[3].numpy(), gbias=grads[4].numpy(), dg=dg)

    # ---------------- 2. EDVR-M forward through the reference module --------------------------------
    shapes = P.edvr_param_shapes()
    net = E.EDVR(nf=64, nframes=5, groups=8, front_RBs=5, back_RBs=10, scale=4)
    ref_sd = net.state_dict()
    assert list(ref_sd.keys()) == list(shapes.keys()), 'EDVR key inventory mismatch'
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), (k, v.shape, shapes[k])
    sd = P.make_params(shapes, seed=1234)
    net.load_state_dict(sd, strict=True)
    net.eval()
    xg = torch.Generator().manual_seed(5)
    xin = torch.rand(1, 5, 3, 32, 32, generator=xg)
    with torch.no_grad():
        out_ref = net(xin)
        out_orc, inter = O.edvr_forward(sd, xin, return_intermediates=True)
    print('edvr: oracle vs reference rel', rel(out_orc, out_ref), 'abs', float((out_orc - out_ref).abs().max()))
    # a few intermediates from the reference modules, for localising failures
    with torch.no_grad():
        L1 = net.lrelu(net.conv_first(xin.view(-1, 3, 32, 32)))
        L1 = net.feature_extraction(L1)
        L2 = net.lrelu(net.fea_L2_conv2(net.lrelu(net.fea_L2_conv1(L1))))
        L3 = net.lrelu(net.fea_L3_conv2(net.lrelu(net.fea_L3_conv1(L2))))
        f = [L1.view(1, 5, 64, 32, 32), L2.view(1, 5, 64, 16, 16), L3.view(1, 5, 64, 8, 8)]
        refl = [t[:, 2].clone() for t in f]
        aligned = torch.stack([net.pcd_align([t[:, i].clone() for t in f], refl) for i in range(5)], 1)
        tsa = net.tsa_fusion(aligned)
    print('edvr: aligned rel', rel(inter['aligned'], aligned), 'tsa rel', rel(inter['tsa'], tsa))
    np.savez(os.path.join(GOLD, 'edvr_m_32.npz'), seed=1234, x=xin.numpy(), out=out_ref.numpy(),
             L3_fea=L3.numpy(), aligned_center=aligned[:, 2].numpy(), aligned_0=aligned[:, 0].numpy(),
             tsa=tsa.numpy())

    # ---------------- 3. MFDN forward ----------------------------------------------------------------
    mshapes = P.mfdn_param_shapes()
    mnet = L.DirectKernelEstimatorVideo(nf=64, in_nc=3, scale=4)
    msd_ref = mnet.state_dict()
    assert list(msd_ref.keys()) == list(mshapes.keys()), 'MFDN key inventory mismatch'
    for k, v in msd_ref.items():
        assert tuple(v.shape) == tuple(mshapes[k]), (k, v.shape, mshapes[k])
    msd = P.make_params(mshapes, seed=77)
    mnet.load_state_dict(msd, strict=True)
    lr = torch.rand(1, 5, 3, 32, 48, generator=xg)
    with torch.no_grad():
        slr_ref = mnet(lr.transpose(1, 2))
        slr_orc = O.mfdn_forward(msd, lr.transpose(1, 2))
    print('mfdn: oracle vs reference rel', rel(slr_orc, slr_ref))
    np.savez(os.path.join(GOLD, 'mfdn_32x48.npz'), seed=77, lr=lr.numpy(), slr=slr_ref.numpy())

    # ---------------- 4. Inner adaptation (test_dynavsr.py:208-283) around the reference modules ----
    import copy
    import torch.nn.functional as F
    msd_fixed = P.make_params(mshapes, seed=78)
    for tag, optimizer, steps, crit, lr_alpha in (('sgd2_l2', 'SGD', 2, 'l2', 1e-4), ('adam1_cb', 'Adam', 1, 'cb', 1e-5)):
        netG = copy.deepcopy(net).train()
        netE = copy.deepcopy(mnet).train()
        netF = copy.deepcopy(mnet)
        netF.load_state_dict(msd_fixed)
        ps = list(netG.parameters()) + list(netE.parameters())
        opt = torch.optim.SGD(ps, lr=lr_alpha) if optimizer == 'SGD' else \
            torch.optim.Adam(ps, lr=lr_alpha, betas=(0.9, 0.99))
        cri = {'l2': torch.nn.MSELoss(reduction='mean'), 'cb': loss_mod.CharbonnierLoss()}[crit]
        losses = []
        for _ in range(steps):
            slr = netE(lr.transpose(1, 2)).transpose(1, 2)
            opt.zero_grad()
            l = cri(netG(slr), lr[:, 2])
            with torch.no_grad():
                slr_f = netF(lr.transpose(1, 2)).transpose(1, 2)
            l = l + 10 * F.l1_loss(slr, slr_f)
            l.backward()
            opt.step()
            losses.append(float(l.detach()))
        netG.eval()
        with torch.no_grad():
            hr = netG(lr)
            hr0 = net(lr)
        o_hr, o_losses, _, _ = O.adapt_and_infer(sd, msd, msd_fixed, lr, steps=steps, lr_alpha=lr_alpha,
                                                 optimizer=optimizer, criterion=crit, return_losses=True)
        print('adapt[%s]: losses ref %s oracle %s ; out rel %.3e ; adapted-vs-unadapted rel %.3e' % (
            tag, losses, o_losses, rel(o_hr, hr), rel(hr, hr0)))
        # parameter deltas of two probes, so a test can check the update itself
        dG = (netG.state_dict()['conv_first.weight'] - sd['conv_first.weight']).numpy()
        dE = (netE.state_dict()['conv6.weight'] - msd['conv6.weight']).numpy()
        np.savez(os.path.join(GOLD, 'adapt_%s.npz' % tag), seed_G=1234, seed_E=77, seed_E_fixed=78,
                 lr=lr.numpy(), out=hr.numpy(), out_unadapted=hr0.numpy(), losses=np.array(losses),
                 steps=steps, lr_alpha=lr_alpha, d_conv_first=dG, d_conv6=dE)
    print('golden written to', GOLD)


if __name__ == '__main__':
    main()
