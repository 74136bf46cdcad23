"""oracle/make_golden_wrapper.py -- TEST INFRASTRUCTURE ONLY; run in the build container.

Training-side behaviour of the UNMODIFIED reference ``VideoBaseModel`` (codes/models/Video_base_model.py:16-251) on CPU:
which parameters end up in which optimiser group at which learning rate for the ``ft_tsa_only`` / ``small_offset_lr``
options, and four iterations of the training loop's per-iteration calls (``update_learning_rate`` -> ``feed_data`` ->
``optimize_parameters``, train.py order) on a narrow EDVR: losses, learning rates, parameter deltas of probe tensors
-> tests/golden/wrapper_train.json / wrapper_train.npz.

    python -m oracle.make_golden_wrapper
"""
import json
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch
import torchvision

REF = os.environ.get('DVSR_REFERENCE', '/root/reference/codes')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, 'tests', 'golden')

NET = dict(nf=8, nframes=3, groups=2, front_RBs=1, back_RBs=1)
PROBES = ['conv_first.weight', 'pcd_align.L1_dcnpack.conv_offset_mask.weight', 'tsa_fusion.fea_fusion.weight',
          'recon_trunk.0.conv2.bias', 'conv_last.weight']
CASES = OrderedDict([
    ('plain_adam_cb', dict(optim='Adam', pixel_criterion='cb', lr_G=1e-3)),
    ('small_offset_sgd_l2', dict(optim='SGD', pixel_criterion='l2', lr_G=1e-2, small_offset_lr=True)),
    ('ft_tsa_only3_sgd_l1', dict(optim='SGD', pixel_criterion='l1', lr_G=1e-2, ft_tsa_only=3)),
    ('ft_tsa_and_small_offset', dict(optim='SGD', pixel_criterion='l2', lr_G=1e-2, ft_tsa_only=2, small_offset_lr=True)),
    ('weight_decay_sgd', dict(optim='SGD', pixel_criterion='l2', lr_G=1e-2, weight_decay_G=1e-2)),
])
STEPS = 4


def main():
    torch.set_num_threads(os.cpu_count())
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    for name in ('imageio', 'lmdb'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['models.archs.dcn.deform_conv_cuda'] = types.ModuleType('deform_conv_cuda_stub')
    import models.archs.dcn  # noqa: F401
    dc = sys.modules['models.archs.dcn.deform_conv']
    dc.modulated_deform_conv = lambda x, off, m, w, b, s, p, d, g, dg: torchvision.ops.deform_conv2d(
        x, off, w, b, stride=s, padding=p, dilation=d, mask=m)
    import options.options as option
    from models.Video_base_model import VideoBaseModel
    from oracle import params as P

    shapes = P.edvr_param_shapes(scale=4, **NET)
    sd0 = P.make_params(shapes, seed=51)
    g = torch.Generator().manual_seed(52)
    data = {'LQs': torch.rand(1, NET['nframes'], 3, 16, 16, generator=g), 'GT': torch.rand(1, 3, 64, 64, generator=g)}
    meta, arrays = OrderedDict(), {'LQs': data['LQs'].numpy(), 'GT': data['GT'].numpy()}
    for case, train in CASES.items():
        t = dict(pixel_weight=1.0, beta1=0.9, beta2=0.99, lr_scheme='MultiStepLR', lr_steps=[2], lr_gamma=0.5,
                 warmup_iter=-1)
        t.update(train)
        opt = option.dict_to_nonedict({
            'model': 'video_base', 'scale': 4, 'gpu_ids': None, 'dist': False, 'is_train': True,
            'network_G': dict(which_model_G='EDVR', predeblur=False, HR_in=False, w_TSA=True, **NET),
            'path': {'strict_load': True, 'pretrain_model_G': None}, 'train': t})
        model = VideoBaseModel(opt)
        net = model.netG.module
        net.load_state_dict(sd0, strict=True)
        names = {id(p): k for k, p in net.named_parameters()}
        groups = [{'lr': grp['lr'], 'names': [names[id(p)] for p in grp['params']]} for grp in model.optimizer_G.param_groups]
        losses, lrs = [], []
        for step in range(1, STEPS + 1):
            model.update_learning_rate(step, warmup_iter=t['warmup_iter'])
            model.feed_data(data)
            model.optimize_parameters(step)
            losses.append(float(model.get_current_log()['l_pix']))
            lrs.append([grp['lr'] for grp in model.optimizer_G.param_groups])
        new = net.state_dict()
        for k in PROBES:
            arrays['%s/%s' % (case, k)] = (new[k] - sd0[k]).numpy()
        meta[case] = dict(train=train, groups=groups, losses=losses, lrs_after_step=lrs)
        print(case, 'groups', [(grp['lr'], len(grp['names'])) for grp in groups], 'losses', ['%.5f' % v for v in losses],
              'lrs', lrs)
    with open(os.path.join(GOLD, 'wrapper_train.json'), 'w') as f:
        json.dump(dict(net=NET, seed=51, steps=STEPS, probes=PROBES, cases=meta), f)
    np.savez_compressed(os.path.join(GOLD, 'wrapper_train.npz'), **arrays)


if __name__ == '__main__':
    main()
