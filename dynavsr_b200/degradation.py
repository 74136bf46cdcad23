"""On-GPU blur-and-subsample degradation (HR frames -> LR network inputs), same API as the reference's
``Degradation`` (codes/data/random_kernel_generator.py:7-130): ``Degradation(kernel_size, scale_factor, theta, sigma)``,
``set_parameters``, ``build_kernel``, ``kernel_shift``, ``get_kernel``, ``set_kernel_directly``, ``apply(img)``.

The kernel construction is a few hundred host flops (numpy + scipy.ndimage, as in the reference; ``np.int`` at :72 is
replaced by ``int``); the image work -- reflection pad, depthwise conv with stride = scale over every frame, optional
8-bit quantisation (vsrbase.py:188) -- is ONE launch of ``dvsr_degrade`` on channels-last frames, so synthetic or decoded
HR clips can be degraded where they already live and fed to EDVR / MFDN without a host round trip
(``apply_nhwc``).  CPU tensors raise NotImplementedError (no CPU fallback).
"""
import ctypes

import numpy as np
import torch
from scipy import ndimage

from . import ops
from ._lib import call


class Degradation(object):
    def __init__(self, kernel_size, scale_factor, theta=0.0, sigma=[1.0, 1.0]):
        self.kernel_size = kernel_size
        self.scale = scale_factor
        self.theta = theta
        self.sigma = sigma
        self._dev = None
        self.build_kernel()

    def set_parameters(self, sigma, theta):
        self.sigma = sigma
        self.theta = theta

    def build_kernel(self):
        """Anisotropic Gaussian with covariance rotated by theta (:20-48); sigma = (0, 0) gives a delta."""
        n = self.kernel_size
        if self.sigma[0] == 0 and self.sigma[1] == 0:
            kernel = np.zeros((n, n))
            kernel[n // 2, n // 2] = 1
        else:
            r = n // 2
            axis = np.linspace(-r, r, n)
            xx, yy = np.meshgrid(axis, axis)
            c, s = np.cos(self.theta), np.sin(self.theta)
            ax, ay = 2.0 * self.sigma[0] ** 2, 2.0 * self.sigma[1] ** 2
            qa = c * c / ax + s * s / ay
            qb = s * c * (1.0 / ay - 1.0 / ax)
            qc = s * s / ax + c * c / ay
            kernel = np.exp(-(qa * xx ** 2 + 2.0 * qb * xx * yy + qc * yy ** 2))
            kernel = kernel / kernel.sum()
        self.kernel = kernel
        self._dev = None

    def kernel_shift(self, kernel):
        """Centre of mass moved to the middle of the first scale x scale block (:50-75), so that the LR grid is aligned
        with the HR one."""
        com = np.array(ndimage.center_of_mass(kernel))
        wanted = np.array(kernel.shape) // 2 + 0.5 * (self.scale - (kernel.shape[0] % 2))
        shift_vec = wanted - com
        kernel = np.pad(kernel, int(np.ceil(np.max(shift_vec))) + 1, 'constant')
        return ndimage.shift(kernel, shift_vec)

    def get_kernel(self):
        return self.kernel

    def set_kernel_directly(self, kernel):
        self.kernel = kernel
        self._dev = None

    def _device_kernels(self, device):
        if self._dev is None or self._dev[0].device != device:
            ks = [self.kernel] if self.kernel.ndim == 2 else list(self.kernel)
            # per-frame kernels: the padding kernel_shift adds depends on each kernel's own shift, so the shifted kernels can differ
            # in size (always by an even amount: size + 2 * pad).  Zero-padding the smaller ones symmetrically to the common size
            # keeps every centre, i.e. gives the result of applying each kernel on its own as the reference does (:100-118).
            sk = [self.kernel_shift(k) for k in ks]
            Lmax = max(k.shape[0] for k in sk)
            sk = [np.pad(k, (Lmax - k.shape[0]) // 2, 'constant') for k in sk]
            shifted = np.stack(sk).astype(np.float32)
            self._dev = (torch.from_numpy(shifted).to(device), shifted.shape[1])
        return self._dev

    def apply_nhwc(self, frames, quantize=False):
        """frames: [T, H, W, C] CUDA tensor (C <= 4) -> [T, H/scale, W/scale, C]."""
        if not frames.is_cuda:
            raise NotImplementedError('dynavsr_b200.degradation is CUDA-only (no CPU fallback)')
        frames = frames.contiguous().float()
        T, H, W, C = frames.shape
        kdev, L = self._device_kernels(frames.device)
        Tk = kdev.shape[0]
        if self.kernel.ndim == 2:
            kmode = 0
        else:
            assert T == Tk or T == Tk + 2, 'per-frame kernels: T must be len(kernels) or len(kernels) + 2'   # :108 (EDVR, DUF)
            kmode = 1 if T == Tk else 2
        s = int(self.scale)
        Ho, Wo = (H + 2 * (L // 2) - L) // s + 1, (W + 2 * (L // 2) - L) // s + 1
        y = torch.empty(T, Ho, Wo, C, device=frames.device, dtype=torch.float32)
        call('dvsr_degrade', ctypes.c_void_p(frames.data_ptr()), ctypes.c_void_p(kdev.data_ptr()), ctypes.c_void_p(y.data_ptr()),
             T, H, W, C, L, s, Tk, kmode, 1 if quantize else 0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        return y

    def apply(self, img):
        """Reference contract (:83-130): img [T, C, H, W] or [C, H, W] -> LR image(s), same rank."""
        if not img.is_cuda:
            raise NotImplementedError('dynavsr_b200.degradation is CUDA-only (no CPU fallback)')
        single = img.ndim == 3
        x = img[None] if single else img
        y = ops.to_nchw(self.apply_nhwc(ops.to_nhwc(x.float())))
        return y[0] if single else y
