"""tools/profile_step.py -- per-entry-point GPU time of one adapted frame (CUDA events around every C-ABI call,
eager mode).  DVSR_PROFILE=1 python tools/profile_step.py [--no-tc]"""
import os
import sys

os.environ['DVSR_PROFILE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import INNER, LR_H, LR_W, NFR, SCALE, synth_clip  # noqa: E402
from dynavsr_b200 import _lib, adapt, ops  # noqa: E402
from dynavsr_b200.models.archs import EDVR_arch, LRimg_estimator  # noqa: E402
from dynavsr_b200.synth import seed_parameters  # noqa: E402

ops.set_conv_backend('--no-tc' not in sys.argv)
netG = seed_parameters(EDVR_arch.EDVR(), 1234)
netE = seed_parameters(LRimg_estimator.DirectKernelEstimatorVideo(64, 3, SCALE), 77)
netF = seed_parameters(LRimg_estimator.DirectKernelEstimatorVideo(64, 3, SCALE), 78)
eng = adapt.InnerLoopAdapter(netG.cuda(), netE.cuda(), netF.cuda(), use_graphs=False, **INNER)
fr = ops.to_nhwc(synth_clip(1, LR_H, LR_W).cuda().reshape(NFR, 3, LR_H, LR_W))
for _ in range(2):
    eng.adapt_and_infer_nhwc(fr)
torch.cuda.synchronize()
_lib.PROFILE['events'].clear()
eng.adapt_and_infer_nhwc(fr)
print(_lib.profile_report(top=60))
_lib.PROFILE['events'].clear()
eng.adapt_and_infer_nhwc(fr)
print(_lib.profile_report(top=40, by_tag=False))
