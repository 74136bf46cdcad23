"""dynavsr_b200 -- B200 (sm_100a) implementation of DynaVSR's data-parallel hot path.

EDVR forward/backward (PCD deformable alignment, TSA fusion, 3x3 trunk, PixelShuffle head) and the
MAML inner-loop step, behind the reference's own model API (``models.archs.EDVR_arch.EDVR``,
``models.archs.dcn.ModulatedDeformConvPack``, ``models.archs.LRimg_estimator.DirectKernelEstimatorVideo``).
All arithmetic runs in hand-written CUDA kernels (``libdvsr_b200.so``, C ABI in include/dvsr_b200.h);
there is no CPU or library fallback.
"""
from . import _lib  # noqa: F401

__version__ = '0.1.0'
