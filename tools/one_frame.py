"""tools/one_frame.py -- one eagerly launched adapted frame (target for `ncu -k regex:<kernel>`): restore, MFDN_fixed, 2 inner
steps at the SLR resolution, final EDVR forward at 176x320."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import INNER, LR_H, LR_W, NFR, SCALE, synth_clip  # noqa: E402
from dynavsr_b200 import adapt, ops  # noqa: E402
from dynavsr_b200.models.archs import EDVR_arch, LRimg_estimator  # noqa: E402
from dynavsr_b200.synth import seed_parameters  # noqa: E402

ops.set_conv_backend(True)
netG = seed_parameters(EDVR_arch.EDVR(), 1234).cuda()
netE = seed_parameters(LRimg_estimator.DirectKernelEstimatorVideo(64, 3, SCALE), 77).cuda()
netF = seed_parameters(LRimg_estimator.DirectKernelEstimatorVideo(64, 3, SCALE), 78).cuda()
eng = adapt.InnerLoopAdapter(netG, netE, netF, use_graphs=False, **INNER)
fr = ops.to_nhwc(synth_clip(1, LR_H, LR_W).cuda().reshape(NFR, 3, LR_H, LR_W))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(n):
    eng.adapt_and_infer_nhwc(fr)
torch.cuda.synchronize()
print('done')
