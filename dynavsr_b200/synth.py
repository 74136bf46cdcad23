"""Seeded synthetic weights and LR windows for benchmarks and smoke runs (no datasets / checkpoints exist offline).

Product-side helper: it depends on nothing but torch and the module it is given (the test oracle under ``oracle/`` has
its own, independent generator).
"""
import math

import torch


def seed_parameters(module, seed, residual_scale=0.1, offset_std=0.02, bias_std=0.05):
    """Fill ``module``'s parameters (in ``named_parameters`` order) with seeded values: N(0, 2/fan_in) weights (x
    ``residual_scale`` inside residual blocks, as arch_util.py:7-24 scales them), small random biases and NON-zero
    ``conv_offset_mask`` weights -- the reference zero-initialises those (deform_conv.py:270-272), which would make every
    deformable sample hit an integer pixel and hide the bilinear gather cost."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in module.named_parameters():
            if k.endswith('.bias'):
                v = torch.randn(p.shape, generator=g, dtype=torch.float64) * bias_std
            else:
                fan_in = 1
                for s in p.shape[1:]:
                    fan_in *= s
                std = math.sqrt(2.0 / fan_in)
                if 'feature_extraction' in k or 'recon_trunk' in k:
                    std *= residual_scale
                if 'conv_offset_mask' in k:
                    std = offset_std
                v = torch.randn(p.shape, generator=g, dtype=torch.float64) * std
            p.copy_(v.to(p.dtype))
    return module


def synth_clip(seed, H, W, nfr=5):
    """Seeded band-limited noise with a global translation of <= 2 px/frame, in [0, 1], quantised to 8 bits like the
    reference pipeline (vsrbase.py:188).  [1, nfr, 3, H, W] float32 (CPU)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(1, 3, H // 4 + 8, W // 4 + 8, generator=g)
    base = F.interpolate(base, scale_factor=4, mode='bicubic', align_corners=False).clamp(0, 1)
    frames = []
    for t in range(nfr):
        dy, dx = 8 + (t * 2) % 5, 8 + (t * 3) % 7
        frames.append(base[:, :, dy:dy + H, dx:dx + W])
    clip = torch.stack(frames, 1)
    return (clip * 255).round() / 255
