"""oracle/make_golden_sfdn.py -- TEST INFRASTRUCTURE ONLY; run in the build container.

SFDN (DirectKernelEstimator_CMS, LRimg_estimator.py:38-67) through the UNMODIFIED reference module on seeded weights / input
-> tests/golden/sfdn_32x48.npz: output, and the gradients of sum(out * probe) w.r.t. the input and two weights (the mean
m = x.mean(2).mean(3) stays in the graph, :58,66).  The oracle restatement (edvr_oracle.sfdn_forward) is asserted against it
here and in tests/test_oracle.py.

    python -m oracle.make_golden_sfdn
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import edvr_oracle as O
    from oracle import params as P
    from oracle.make_golden import import_reference, rel
    _, L, _, _ = import_reference()
    net = L.DirectKernelEstimator_CMS(nf=64)
    shapes = P.sfdn_param_shapes(64)
    ref_sd = net.state_dict()
    assert list(ref_sd.keys()) == list(shapes.keys()) and all(tuple(ref_sd[k].shape) == tuple(shapes[k]) for k in shapes)
    seed = 41
    sd = P.make_params(shapes, seed=seed)
    net.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(8)
    x = torch.rand(3, 3, 32, 48, generator=g).requires_grad_(True)
    probe = torch.randn(3, 3, 16, 24, generator=g)
    out = net(x)
    (out * probe).sum().backward()
    with torch.no_grad():
        mine = O.sfdn_forward(sd, x.detach())
    assert rel(mine, out.detach()) < 1e-6, rel(mine, out.detach())
    np.savez_compressed(os.path.join(GOLD, 'sfdn_32x48.npz'), seed=seed, x=x.detach().numpy(), probe=probe.numpy(),
                        out=out.detach().numpy(), gx=x.grad.numpy(), g_conv0_w=net.conv0.weight.grad.numpy(),
                        g_conv3_w=net.conv3.weight.grad.numpy(), g_conv6_b=net.conv6.bias.grad.numpy())
    print('sfdn golden written; oracle rel %.2e' % rel(mine, out.detach()))


if __name__ == '__main__':
    main()
