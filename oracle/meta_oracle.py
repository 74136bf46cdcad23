"""oracle/meta_oracle.py -- TEST INFRASTRUCTURE ONLY.

Plain-PyTorch CPU restatement of one outer (meta) step of codes/train_dynavsr.py:252-438 for EDVR + MFDN:
``fomaml`` = the evident intent (working copy adapted K steps, query gradients taken at the adapted weights, applied to the
base weights; what the reference's own validation loop :633-677 does), ``as_written`` = the literal behaviour of the
training loop (the inner optimiser is bound to the copy, the loss to the original, :322-399 -- nothing adapts and the
inner gradients pile up unscaled in the outer gradient).  See SURVEY.md section 3.2.

Parity pin: oracle/make_golden_meta.py runs the UNMODIFIED ``main()`` of train_dynavsr.py on CPU (stand-ins only for
absent modules, the data loader, output paths and the hard-coded ``.to('cuda')``) and stores the weights it leaves after
outer Adam (two iterations) and outer SGD (lr 1: update = accumulated gradient) steps in tests/golden/meta_loop*.npz;
``as_written`` reproduces the reference's accumulated outer gradient to 8e-8 relative
(tests/test_oracle.py::test_meta_oracle_*).  ``fomaml`` has no reference counterpart to pin against (the reference's
training loop never runs it): it is built from the pinned pieces (edvr_forward / mfdn_forward / pixel_loss, torch.optim)
and differs from ``as_written`` only in which weights the losses are evaluated at.
"""
import torch
import torch.nn.functional as F

from .edvr_oracle import edvr_forward, mfdn_forward, pixel_loss


def meta_outer_step(sd_G, sd_E, tasks, inner_steps=1, lr_alpha=1e-5, lr_alpha_est=None, inner_optimizer='Adam',
                    inner_betas=(0.9, 0.99), criterion='cb', est_loss='l1', outer='Adam', lr_outer=1e-5,
                    outer_betas=(0.9, 0.99), outer_state=None, mode='fomaml', scale=4, world_grads=None, edvr_cfg=None):
    """Returns (new sd_G, new sd_E, info).  ``tasks``: list of dicts LQs [1,N,3,h,w], GT [1,3,sh,sw], SuperLQs [1,N,3,h/s,w/s].
    ``world_grads``: optional list of (gG, gE) dicts from other ranks, averaged with this rank's (the all-reduce)."""
    cfg = dict(edvr_cfg or {})
    lr_alpha_est = lr_alpha if lr_alpha_est is None else lr_alpha_est
    B = len(tasks)
    gG = {k: torch.zeros_like(v) for k, v in sd_G.items()}
    gE = {k: torch.zeros_like(v) for k, v in sd_E.items()}
    lq_all, le_all, inner_all = [], [], []
    lossf = {'l1': F.l1_loss, 'l2': F.mse_loss}[est_loss]
    for t in tasks:
        LQs, GT, SLQ = t['LQs'], t['GT'], t['SuperLQs']
        center = LQs.shape[1] // 2
        pG = {k: v.detach().clone().requires_grad_(True) for k, v in sd_G.items()}      # :322
        pE = {k: v.detach().clone().requires_grad_(True) for k, v in sd_E.items()}
        groups = [{'params': list(pG.values()), 'lr': lr_alpha}, {'params': list(pE.values()), 'lr': lr_alpha_est}]   # :335-344
        opt = torch.optim.SGD(groups, lr=lr_alpha) if inner_optimizer == 'SGD' else \
            torch.optim.Adam(groups, lr=lr_alpha, betas=inner_betas)
        xin = LQs.transpose(1, 2)
        inner = []
        for _ in range(inner_steps):
            if mode == 'fomaml':
                opt.zero_grad()
            slr = mfdn_forward(pE, xin, scale).transpose(1, 2)                            # :360-363
            loss = pixel_loss(criterion, edvr_forward(pG, slr, scale=scale, **cfg), LQs[:, center]) + F.l1_loss(slr, SLQ)   # :385,:395
            if mode == 'fomaml':
                loss.backward()
                opt.step()                                                                # :397-399
            else:
                gs = torch.autograd.grad(loss, list(pG.values()) + list(pE.values()))     # lands in the ORIGINAL's .grad
                for k, g in zip(gG.keys(), gs[:len(gG)]):
                    gG[k] += g
                for k, g in zip(gE.keys(), gs[len(gG):]):
                    gE[k] += g
            inner.append(float(loss.detach()))
        lq = pixel_loss(criterion, edvr_forward(pG, LQs, scale=scale, **cfg), GT)         # :404-405
        for k, g in zip(gG.keys(), torch.autograd.grad(lq / B, list(pG.values()))):       # :413-415
            gG[k] += g
        le = lossf(mfdn_forward(pE, xin, scale).transpose(1, 2), SLQ)                     # :417-419
        for k, g in zip(gE.keys(), torch.autograd.grad(le / (B * 10), list(pE.values()))):   # :423-426
            gE[k] += g
        lq_all.append(float(lq.detach())); le_all.append(float(le.detach())); inner_all.append(inner)
    if world_grads:
        n = 1 + len(world_grads)
        for k in gG:
            gG[k] = (gG[k] + sum(w[0][k] for w in world_grads)) / n
        for k in gE:
            gE[k] = (gE[k] + sum(w[1][k] for w in world_grads)) / n
    # outer update (:438)
    base = [v.detach().clone().requires_grad_(True) for v in list(sd_G.values()) + list(sd_E.values())]
    oo = torch.optim.SGD(base, lr=lr_outer) if outer == 'SGD' else torch.optim.Adam(base, lr=lr_outer, betas=outer_betas)
    if outer_state is not None:
        oo.load_state_dict(outer_state)
    for p, g in zip(base, list(gG.values()) + list(gE.values())):
        p.grad = g.clone()
    oo.step()
    nG = {k: p.detach() for k, p in zip(sd_G.keys(), base[:len(sd_G)])}
    nE = {k: p.detach() for k, p in zip(sd_E.keys(), base[len(sd_G):])}
    return nG, nE, {'loss_q': lq_all, 'loss_e': le_all, 'inner': inner_all, 'gG': gG, 'gE': gE, 'outer_state': oo.state_dict()}
