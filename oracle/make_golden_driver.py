"""oracle/make_golden_driver.py -- TEST INFRASTRUCTURE ONLY; run in the build container, not on the GPU box.

Pins the north-star path against the reference's OWN test-time driver: runs the UNMODIFIED ``main()`` of
/root/reference/codes/test_dynavsr.py (:33-365 -- baseline inference, per-clip deepcopy, K inner steps on EDVR + MFDN with
the frozen-MFDN L1 term, final inference, tensor2img, PSNR / SSIM) on CPU for one synthetic clip, with EDVR-M 4x + MFDN at
full width and seeded weights, and stores what it produced in tests/golden/driver_<tag>.npz:

    out            float32 [3, 4h, 4w]   the adapted SR frame (``modelcp.fake_H`` when the driver writes its PNG; on CPU the
                                         driver's tensor2img has clamped it to [0, 1] in place by then, utils/util.py:118)
    image          uint8  [4h, 4w, 3]    the PNG the driver hands to imageio.imwrite
    psnr / ssim    the four numbers of the driver's own psnr_update.csv (baseline model, adapted model)
    d_conv_first, d_conv6                parameter deltas of two probe tensors after adaptation

Nothing under /root/reference is modified or written to.  What the harness supplies around the script: stand-ins for the
absent ``imageio`` / ``lmdb``; the DCN op routed to ``torchvision.ops.deform_conv2d``; a one-clip data loader; the four
checkpoint files the YML names (written to a scratch directory in this repo from oracle/params.py seeds); the working
directory set inside the scratch directory (the script writes ``../test_results``); ``Tensor.to('cuda')`` answered on CPU.

    python -m oracle.make_golden_driver [tag ...]       # all configurations, or the named ones
"""
import os
import shutil
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F
import torchvision

REF = os.environ.get('DVSR_REFERENCE', '/root/reference/codes')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, 'tests', 'golden')
SCRATCH = os.path.join(HERE, '_ref', 'driver_scratch')

SEED_G, SEED_E, SEED_E_FIXED, SEED_BASELINE_G, SEED_CLIP = 1234, 77, 78, 1235, 4321
NFRAMES, SCALE, H, W = 5, 4, 32, 48
CONFIGS = {'adam1_cb': dict(optimizer='Adam', steps=1, criterion='cb', lr_alpha='1e-5'),      # the shipped test YMLs
           'sgd2_l2': dict(optimizer='SGD', steps=2, criterion='l2', lr_alpha='1e-4'),        # BASELINE.json's 2-step case
           # the two optional branches of the inner loop (test_dynavsr.py:118-145,218-221,237-260)
           'sgd2_l2_patch': dict(optimizer='SGD', steps=2, criterion='l2', lr_alpha='1e-4', use_patch='true', num_patch=2,
                                 patch_size=16),          # SLR crops of 8x8 (patch_size // 2) out of 8x12, LR crops of 32x32
           'adam1_cb_real': dict(optimizer='Adam', steps=1, criterion='cb', lr_alpha='1e-5', use_real='true')}
DEFAULTS = dict(use_patch='false', num_patch=1, patch_size=64, use_real='false')
PATCH_SEED = 2024          # random.seed() right before main(): the driver draws its crop positions from `random` (preprocessing.py:76-77)

YML = """
name: driver_pin
use_tb_logger: false
model: video_base+lrimgestimator
distortion: sr
scale: {scale}
gpu_ids: ~
datasets:
  train:
    name: unused
    mode: synthetic
    N_frames: {nframes}
    kernel_size: 21
    batch_size: 1
    patch_size: 64
  val:
    name: synthetic
    mode: synthetic
    N_frames: {nframes}
network_G:
  which_model_G: EDVR
  nf: 64
  nframes: {nframes}
  groups: 8
  front_RBs: 5
  back_RBs: 10
  predeblur: false
  HR_in: false
  w_TSA: true
network_E:
  which_model_E: MFDN
  mode: video
  nf: 64
  in_nc: 3
path:
  bicubic_G: {dir}/baseline_G.pth
  fixed_E: {dir}/fixed_E.pth
  pretrain_model_G: {dir}/meta_G.pth
  pretrain_model_E: {dir}/meta_E.pth
  strict_load: true
  resume_state: ~
train:
  lr_C: !!float 1e-4
  lr_G: !!float 1e-5
  lr_scheme: MultiStepLR
  optim: Adam
  beta1: 0.9
  beta2: 0.99
  niter: 1
  warmup_iter: -1
  lr_steps: [1000]
  lr_gamma: 0.2
  loss_ftn: l1
  use_real: {use_real}
  maml:
    use_patch: {use_patch}
    num_patch: {num_patch}
    patch_size: {patch_size}
    optimizer: {optimizer}
    lr_alpha: !!float {lr_alpha}
    beta1: 0.9
    beta2: 0.99
    adapt_iter: {steps}
  pixel_criterion: {criterion}
  pixel_weight: 1.0
  val_freq: !!float 500
  manual_seed: 0
logger:
  print_freq: 100
  save_checkpoint_freq: !!float 1e3
"""


HEAD_GAIN = 0.02


def tame_head(sd):
    """Scale ``conv_last`` of seeded EDVR weights in place (tests apply the same factor)."""
    sd['conv_last.weight'] *= HEAD_GAIN
    sd['conv_last.bias'] *= HEAD_GAIN
    return sd


def synthetic_clip():
    """A smooth moving scene: HR frames are a bicubically enlarged random field shifted by one pixel per frame; the LR
    clip is their 4x bicubic reduction plus a little noise, so PSNR values are in a natural range (not ~6 dB of noise)."""
    g = torch.Generator().manual_seed(SEED_CLIP)
    base = torch.rand(1, 3, 12, 16, generator=g)
    big = F.interpolate(base, size=(H * SCALE + 16, W * SCALE + 16), mode='bicubic', align_corners=False).clamp(0, 1)
    hr = torch.stack([big[0, :, 8 + t:8 + t + H * SCALE, 8 + 2 * t:8 + 2 * t + W * SCALE] for t in range(NFRAMES)])
    lq = F.interpolate(hr, scale_factor=1.0 / SCALE, mode='bicubic', align_corners=False)
    lq = (lq + 0.01 * torch.randn(lq.shape, generator=g)).clamp(0, 1)
    return lq.unsqueeze(0).contiguous(), hr.unsqueeze(0).contiguous()


def run(tag, cfg, T, P, state):
    yml = os.path.join(SCRATCH, 'driver_%s.yml' % tag)
    with open(yml, 'w') as f:
        f.write(YML.format(scale=SCALE, nframes=NFRAMES, dir=SCRATCH, **dict(DEFAULTS, **cfg)))
    lq, hr = synthetic_clip()
    # use_real: the dataset supplies the pre-generated super-LR clip (video_test_dataset_int.py 'SuperLQs'); here a bicubic
    # 4x reduction of the LR clip
    slq = F.interpolate(lq[0], scale_factor=1.0 / SCALE, mode='bicubic', align_corners=False).clamp(0, 1).unsqueeze(0).contiguous()
    state.update(models=[], written=[], crops=[])
    import random
    random.seed(PATCH_SEED)
    real_randrange = random.randrange

    def logged_randrange(*a, **k):
        v = real_randrange(*a, **k)
        state['crops'].append(v)
        return v
    random.randrange = logged_randrange

    class Loader(object):
        def __len__(self):
            return 1

        def __iter__(self):
            yield {'LQs': lq.clone(), 'GT': hr.clone(), 'SuperLQs': slq.clone(), 'folder': ['clip'], 'idx': ['0/1']}

    T.create_dataset = lambda dataset_opt, **kw: [0]
    T.create_dataloader = lambda dataset, dataset_opt, opt=None, sampler=None: Loader()
    argv = sys.argv
    sys.argv = ['test_dynavsr.py', '-opt', yml, '--exp_name', 'pin_' + tag]
    try:
        T.main()
    finally:
        sys.argv = argv
        random.randrange = real_randrange
    assert len(state['written']) == 1
    image, out, dG, dE = state['written'][0]
    import pandas as pd
    csv = pd.read_csv(os.path.join(SCRATCH, 'test_results', 'pin_' + tag, 'psnr_update.csv'), index_col=0)
    row = csv.iloc[0]
    np.savez_compressed(os.path.join(GOLD, 'driver_%s.npz' % tag), lq=lq.numpy(), gt=hr[0, NFRAMES // 2].numpy(),
                        out=out.numpy(), image=image, d_conv_first=dG.numpy(), d_conv6=dE.numpy(),
                        psnr_baseline=float(row['PSNR_Bicubic']), psnr_adapted=float(row['PSNR_Ours']),
                        ssim_baseline=float(row['SSIM_Bicubic']), ssim_adapted=float(row['SSIM_Ours']),
                        seed_G=SEED_G, seed_E=SEED_E, seed_E_fixed=SEED_E_FIXED, seed_baseline_G=SEED_BASELINE_G,
                        steps=cfg['steps'], lr_alpha=float(cfg['lr_alpha']), optimizer=cfg['optimizer'],
                        criterion=cfg['criterion'], head_gain=HEAD_GAIN, slq=slq.numpy(), crops=np.array(state['crops'], dtype=np.int64),
                        patch_seed=PATCH_SEED, use_patch=cfg.get('use_patch', 'false') == 'true', num_patch=int(cfg.get('num_patch', 1)),
                        patch_size=int(cfg.get('patch_size', 64)), use_real=cfg.get('use_real', 'false') == 'true')
    print('[%s] driver PSNR baseline %.4f adapted %.4f ; SSIM %.4f / %.4f' % (
        tag, row['PSNR_Bicubic'], row['PSNR_Ours'], row['SSIM_Bicubic'], row['SSIM_Ours']))
    return lq, hr, out, slq, list(state['crops'])


def main():
    torch.set_num_threads(os.cpu_count())
    shutil.rmtree(SCRATCH, ignore_errors=True)
    os.makedirs(os.path.join(SCRATCH, 'work'))
    os.makedirs(GOLD, exist_ok=True)
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    from oracle import edvr_oracle as O
    from oracle import params as P

    sdG = P.make_params(P.edvr_param_shapes(), seed=SEED_G)
    sdE = P.make_params(P.mfdn_param_shapes(), seed=SEED_E)
    sdF = P.make_params(P.mfdn_param_shapes(), seed=SEED_E_FIXED)
    sdB = P.make_params(P.edvr_param_shapes(), seed=SEED_BASELINE_G)
    for sd in (sdG, sdB):          # a small residual on top of the bilinear skip: outputs stay inside [0, 1] like a trained
        tame_head(sd)              # model's, so the uint8 image and its PSNR are informative
    for name, sd in (('meta_G', sdG), ('meta_E', sdE), ('fixed_E', sdF), ('baseline_G', sdB)):
        torch.save(sd, os.path.join(SCRATCH, name + '.pth'))

    state = {}
    imageio = types.ModuleType('imageio')

    def imwrite(path, image):
        # called once per clip with the adapted result (test_dynavsr.py:283); the second create_model() result is the
        # working copy the driver adapts
        modelcp, est_modelcp = state['models'][1]
        g_sd, e_sd = modelcp.netG.module.state_dict(), est_modelcp.netE.module.state_dict()
        state['written'].append((np.array(image), modelcp.fake_H.detach()[0].float().cpu().clone(),
                                 (g_sd['conv_first.weight'] - sdG['conv_first.weight']).clone(),
                                 (e_sd['conv6.weight'] - sdE['conv6.weight']).clone()))

    imageio.imwrite = imwrite
    sys.modules['imageio'] = imageio
    sys.modules.setdefault('lmdb', types.ModuleType('lmdb'))
    sys.modules['models.archs.dcn.deform_conv_cuda'] = types.ModuleType('deform_conv_cuda_stub')
    import models.archs.dcn  # noqa: F401
    dc = sys.modules['models.archs.dcn.deform_conv']
    dc.modulated_deform_conv = lambda x, off, m, w, b, s, p, d, g, dg: torchvision.ops.deform_conv2d(
        x, off, w, b, stride=s, padding=p, dilation=d, mask=m)

    cwd = os.getcwd()
    os.chdir(os.path.join(SCRATCH, 'work'))            # the driver writes ../test_results/<exp_name>/...
    import test_dynavsr as T
    real_create = T.create_model

    def create_model(opt):
        models = real_create(opt)
        state['models'].append(models)
        return models

    T.create_model = create_model
    real_to = torch.Tensor.to

    def to(self, *a, **k):
        if a and isinstance(a[0], str) and a[0].startswith('cuda'):
            a = ('cpu',) + a[1:]
        return real_to(self, *a, **k)

    torch.Tensor.to = to
    try:
        tags = [a for a in sys.argv[1:] if a in CONFIGS] or list(CONFIGS)
        for tag in tags:
            cfg = CONFIGS[tag]
            lq, hr, out, slq, crops = run(tag, cfg, T, P, state)
            extra = {}
            if cfg.get('use_real') == 'true':
                extra['slr_given'] = slq
            if cfg.get('use_patch') == 'true':
                n = int(cfg['num_patch'])
                pos = [(crops[2 * i], crops[2 * i + 1]) for i in range(len(crops) // 2)]
                extra.update(patches=[pos[k * n:(k + 1) * n] for k in range(cfg['steps'])], patch_size=int(cfg['patch_size']))
            o_hr = O.adapt_and_infer(sdG, sdE, sdF, lq, steps=cfg['steps'], lr_alpha=float(cfg['lr_alpha']),
                                     optimizer=cfg['optimizer'], criterion=cfg['criterion'], **extra)
            o_hr = o_hr[0] if isinstance(o_hr, tuple) else o_hr
            print('[%s] oracle vs driver: rel %.3e ; fraction clamped %.4f' % (
                tag, float((o_hr[0].clamp(0, 1) - out).norm() / out.norm()), float(((o_hr < 0) | (o_hr > 1)).float().mean())))
    finally:
        torch.Tensor.to = real_to
        os.chdir(cwd)
    shutil.rmtree(SCRATCH, ignore_errors=True)


if __name__ == '__main__':
    main()
