"""oracle/degradation_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's blur-and-subsample degradation that turns HR frames into LR inputs
(codes/data/random_kernel_generator.py): anisotropic Gaussian (:20-48), centre-of-mass kernel shift (:50-75), and
``Degradation.apply`` = ReflectionPad2d(k // 2) + depthwise conv with stride = scale (:83-130), followed by the
dataset's 8-bit quantisation ``round(x * 255) / 255`` (data/meta_learner/vsrbase.py:188).

Parity pin: oracle/make_golden_degradation.py runs the unmodified reference class (with ``np.int`` aliased -- numpy
removed it) and stores kernels + outputs in tests/golden/degradation.npz.
"""
import numpy as np
import torch
import torch.nn.functional as F
from scipy import ndimage


def gaussian_kernel(kernel_size, sigma, theta):
    """random_kernel_generator.py:20-48."""
    if sigma[0] == 0 and sigma[1] == 0:
        k = np.zeros((kernel_size, kernel_size))
        k[kernel_size // 2, kernel_size // 2] = 1
        return k
    r = kernel_size // 2
    rng = np.linspace(-r, r, kernel_size)
    xx, yy = np.meshgrid(rng, rng)
    ct, st = np.cos(theta), np.sin(theta)
    sx2, sy2 = 2.0 * sigma[0] ** 2, 2.0 * sigma[1] ** 2
    a = ct * ct / sx2 + st * st / sy2
    b = st * ct * (1.0 / sy2 - 1.0 / sx2)
    c = st * st / sx2 + ct * ct / sy2
    k = np.exp(-(a * xx ** 2 + 2.0 * b * xx * yy + c * yy ** 2))
    return k / k.sum()


def shift_kernel(kernel, scale):
    """random_kernel_generator.py:50-75."""
    com = np.array(ndimage.center_of_mass(kernel))
    wanted = np.array(kernel.shape) // 2 + 0.5 * (scale - (kernel.shape[0] % 2))
    shift_vec = wanted - com
    kernel = np.pad(kernel, int(np.ceil(np.max(shift_vec))) + 1, 'constant')
    return ndimage.shift(kernel, shift_vec)


def degrade(img, kernel, scale, quantize=False):
    """random_kernel_generator.py:83-130.  img: [T, C, H, W]; kernel: [K, K] or [Tk, K, K] (per-frame; T == Tk or Tk + 2)."""
    outs = []
    for i in range(img.shape[0]):
        if kernel.ndim == 2:
            k = kernel
        elif img.shape[0] == kernel.shape[0]:
            k = kernel[i]
        else:
            k = kernel[(i - 1) % kernel.shape[0]]
        ks = torch.from_numpy(shift_kernel(k, scale)).float()
        L = ks.shape[0]
        x = F.pad(img[i:i + 1], (L // 2,) * 4, mode='reflect')
        outs.append(F.conv2d(x, ks.repeat(img.shape[1], 1, 1, 1), groups=img.shape[1], stride=int(scale)))
    y = torch.cat(outs, 0)
    return (y * 255).round() / 255 if quantize else y
