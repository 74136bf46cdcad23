"""oracle/make_golden_degradation.py -- TEST INFRASTRUCTURE ONLY; run in the build container.

Runs the UNMODIFIED reference ``Degradation`` class (codes/data/random_kernel_generator.py) on seeded inputs and writes
tests/golden/degradation.npz.  ``np.int`` (removed from numpy >= 1.24, used at :72) and the removed
``scipy.ndimage.measurements / interpolation`` aliases (:2) are provided as aliases, nothing else is touched.

    python -m oracle.make_golden_degradation
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('DVSR_REFERENCE', '/root/reference/codes')
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def main():
    np.int = int
    import scipy.ndimage as ndi
    for name in ('measurements', 'interpolation'):
        if not hasattr(ndi, name):
            m = types.ModuleType('scipy.ndimage.' + name)
            m.center_of_mass, m.shift = ndi.center_of_mass, ndi.shift
            setattr(ndi, name, m)
            sys.modules['scipy.ndimage.' + name] = m
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.dirname(HERE))
    from data.random_kernel_generator import Degradation
    from oracle import degradation_oracle as DO
    g = torch.Generator().manual_seed(5)
    out = {}
    cases = [('iso', 21, 4, [1.6, 1.6], 0.0, (5, 3, 48, 64)), ('aniso', 21, 4, [2.4, 0.9], 0.6, (3, 3, 40, 56)),
             ('x2', 11, 2, [0.8, 1.3], 1.1, (2, 3, 30, 26)), ('delta', 21, 4, [0.0, 0.0], 0.0, (1, 3, 32, 32))]
    for tag, ks, sc, sigma, theta, shape in cases:
        d = Degradation(ks, sc, theta=theta, sigma=sigma)
        img = torch.rand(*shape, generator=g)
        y = d.apply(img)
        yo = DO.degrade(img, DO.gaussian_kernel(ks, sigma, theta), sc)
        print(tag, tuple(y.shape), 'oracle vs reference max abs', float((y - yo).abs().max()),
              'kernel', float(np.abs(d.get_kernel() - DO.gaussian_kernel(ks, sigma, theta)).max()))
        out.update({tag + '_img': img.numpy(), tag + '_kernel': d.get_kernel(), tag + '_out': y.numpy(),
                    tag + '_cfg': np.array([ks, sc, sigma[0], sigma[1], theta], dtype=np.float64)})
    # per-frame kernels (kernel.ndim == 3, :106-124): T == Tk and T == Tk + 2
    ker = np.stack([DO.gaussian_kernel(21, [1.0 + 0.4 * i, 2.0 - 0.3 * i], 0.3 * i) for i in range(3)])
    for tag, T in (('perframe3', 3), ('perframe5', 5)):
        d = Degradation(21, 4)
        d.set_kernel_directly(ker)
        img = torch.rand(T, 3, 36, 44, generator=g)
        y = d.apply(img)
        yo = DO.degrade(img, ker, 4)
        print(tag, 'oracle vs reference max abs', float((y - yo).abs().max()))
        out.update({tag + '_img': img.numpy(), tag + '_kernel': ker, tag + '_out': y.numpy()})
    np.savez_compressed(os.path.join(GOLD, 'degradation.npz'), **out)
    print('wrote', os.path.join(GOLD, 'degradation.npz'))


if __name__ == '__main__':
    main()
