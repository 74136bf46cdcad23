"""Test helpers (layout conversions are done with torch permutes here, on purpose: the product's own
layout kernels are among the things under test)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def maxabs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def gold(name):
    return np.load(os.path.join(GOLD, name))


def psnr_uint8(a, b):
    """utils/util.py:112-142 (tensor2img: clamp, *255, round) + :262-269 (calculate_psnr) restated."""
    a8 = (a.detach().float().cpu().clamp(0, 1) * 255.0).round().double()
    b8 = (b.detach().float().cpu().clamp(0, 1) * 255.0).round().double()
    mse = float(((a8 - b8) ** 2).mean())
    return float('inf') if mse == 0 else 20 * np.log10(255.0 / np.sqrt(mse))


def rel_robust(a, b, drop=1e-4):
    """Relative L2 error after discarding the ``drop`` fraction of elements with the largest absolute difference.  For
    quantities that are DISCONTINUOUS in their inputs -- the deformable-conv offset gradient jumps when a sampling coordinate
    crosses an integer, and an fp32 kernel and the fp64 oracle floor a handful of coordinates (|h| ~ 50, ulp 4e-6) differently --
    a few O(1) outliers out of millions of elements say nothing about the kernel; the rest must agree tightly."""
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    d = (a - b).abs()
    k = int(d.numel() * drop)
    if k > 0:
        keep = d.argsort()[:d.numel() - k]
        a, b = a[keep], b[keep]
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
