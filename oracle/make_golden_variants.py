"""oracle/make_golden_variants.py -- TEST INFRASTRUCTURE ONLY; run in the build container.

EDVR constructor variants no YML uses (predeblur / HR_in / w_TSA=False, EDVR_arch.py:13-57,208-239) through the UNMODIFIED
reference module -> tests/golden/edvr_variants.npz.

    python -m oracle.make_golden_variants
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
VARIANTS = {'predeblur': dict(predeblur=True, HR_in=False, w_TSA=True),
            'hrin_notsa': dict(predeblur=False, HR_in=True, w_TSA=False),
            'predeblur_hrin': dict(predeblur=True, HR_in=True, w_TSA=True)}
CFG = dict(nf=64, nframes=5, groups=8, front_RBs=1, back_RBs=1, scale=4)


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import edvr_oracle as O
    from oracle import params as P
    from oracle.make_golden import import_reference, rel
    E, _, _, _ = import_reference()
    out = {}
    g = torch.Generator().manual_seed(17)
    for tag, v in VARIANTS.items():
        shapes = P.edvr_param_shapes(**CFG, **v)
        net = E.EDVR(**CFG, **v)
        ref_sd = net.state_dict()
        assert list(ref_sd.keys()) == list(shapes.keys()), (tag, [a for a, b in zip(ref_sd.keys(), shapes.keys()) if a != b][:4])
        for k, t in ref_sd.items():
            assert tuple(t.shape) == tuple(shapes[k]), (tag, k)
        sd = P.make_params(shapes, seed=300 + len(out))
        net.load_state_dict(sd, strict=True)
        net.eval()
        hw = 64 if v['HR_in'] else 16
        x = torch.rand(1, 5, 3, hw, hw, generator=g)
        with torch.no_grad():
            y = net(x)
            yo = O.edvr_forward(sd, x, front_RBs=1, back_RBs=1, predeblur_=v['predeblur'], HR_in=v['HR_in'], w_TSA=v['w_TSA'])
        print(tag, tuple(y.shape), 'oracle vs reference rel', rel(yo, y))
        out.update({tag + '_x': x.numpy(), tag + '_out': y.numpy(), tag + '_seed': 300 + len(out)})
    np.savez_compressed(os.path.join(GOLD, 'edvr_variants.npz'), **out)


if __name__ == '__main__':
    main()
