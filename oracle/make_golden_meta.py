"""oracle/make_golden_meta.py -- TEST INFRASTRUCTURE ONLY; run in the build container, not on the GPU box.

Pins the LOOP STRUCTURE of the outer (meta) step: runs the UNMODIFIED ``main()`` of
/root/reference/codes/train_dynavsr.py on CPU for a few outer iterations and records the weights it leaves behind, so
that oracle/meta_oracle.py (mode ``as_written``) can be checked against what the reference's own training driver does
(tests/test_oracle.py::test_meta_oracle_vs_reference_training_loop).

Nothing under /root/reference is modified or written to.  What the harness supplies around the script:
  * empty stand-ins for modules that are absent in this image (imageio, lmdb) and for tensorboard's SummaryWriter;
  * the DCN op routed to ``torchvision.ops.deform_conv2d`` (same recipe as make_golden.py);
  * ``option.parse`` post-processed so every output path lies in a scratch directory inside this repo
    (the script would otherwise create ``experiments/`` under the reference tree);
  * ``loader.get_dataset`` / ``create_dataloader`` replaced by a fixed list of synthetic batches; asking for the batch
    after the last one snapshots the weights and stops the script;
  * ``Tensor.to('cuda')`` (hard-coded at train_dynavsr.py:393-395) answered on the CPU;
  * the first ``create_model`` result is re-seeded deterministically (non-zero offset convs) and its weights recorded.

    python -m oracle.make_golden_meta       # outer Adam, two iterations  -> tests/golden/meta_loop.npz
    python -m oracle.make_golden_meta sgd   # outer SGD (lr 1: the update IS the accumulated gradient), one iteration on
                                            # the same start and first batch  -> tests/golden/meta_loop_sgd.npz
"""
import os
import shutil
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
SCRATCH = os.path.join(HERE, '_ref', 'meta_scratch')

NF, NFRAMES, GROUPS, FRONT, BACK, NF_E, SCALE = 8, 3, 2, 1, 1, 8, 4
H = W = 32
BATCH, ITERS, INNER = 2, 2, 2

YML = """
name: meta_pin
use_tb_logger: true
model: video_base+lrimgestimator
distortion: sr
scale: {scale}
gpu_ids: ~
datasets:
  train:
    name: synthetic
    mode: synthetic
    N_frames: {nframes}
    batch_size: {batch}
    n_workers: 0
    kernel_size: 21
    patch_size: {h}
  val:
    name: none
    mode: synthetic
network_G:
  which_model_G: EDVR
  nf: {nf}
  nframes: {nframes}
  groups: {groups}
  front_RBs: {front}
  back_RBs: {back}
  predeblur: false
  HR_in: false
  w_TSA: true
network_E:
  which_model_E: MFDN
  mode: video
  nf: {nf_e}
  in_nc: 3
path:
  pretrain_model_G: ~
  pretrain_model_E: ~
  strict_load: true
  resume_state: ~
train:
  lr_G: !!float {lr_G}
  lr_C: !!float 1e-3
  lr_scheme: MultiStepLR
  optim: {optim}
  beta1: 0.9
  beta2: 0.99
  niter: {iters}
  warmup_iter: -1
  lr_steps: [1000]
  lr_gamma: 0.5
  pixel_criterion: cb
  pixel_weight: 1.0
  loss_ftn: l1
  val_freq: !!float 1e9
  manual_seed: 0
  use_real: false
  maml:
    optimizer: Adam
    beta1: 0.9
    beta2: 0.99
    lr_alpha: !!float 1e-3
    lr_alpha_est: !!float 2e-3
    use_patch: false
    num_patch: 1
    patch_size: {h}
    adapt_iter: {inner}
logger:
  print_freq: 1
  save_checkpoint_freq: !!float 1e9
"""


class _Stop(Exception):
    pass


def main():
    sgd = len(sys.argv) > 1 and sys.argv[1] == 'sgd'
    ITERS = 1 if sgd else globals()['ITERS']
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    shutil.rmtree(SCRATCH, ignore_errors=True)
    os.makedirs(SCRATCH)
    yml = os.path.join(SCRATCH, 'meta_pin.yml')
    with open(yml, 'w') as f:
        f.write(YML.format(scale=SCALE, nframes=NFRAMES, batch=BATCH, h=H, nf=NF, groups=GROUPS, front=FRONT, back=BACK,
                           nf_e=NF_E, iters=ITERS, inner=INNER, optim='SGD' if sgd else 'Adam',
                           lr_G='1.0' if sgd else '1e-3'))

    # ---- stand-ins for what the image lacks / what must not touch the reference tree
    for name in ('imageio', 'lmdb'):
        sys.modules.setdefault(name, types.ModuleType(name))
    tb = types.ModuleType('torch.utils.tensorboard')
    scalars = []

    class SummaryWriter(object):
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, tag, value, global_step=None):
            scalars.append((tag, float(value), global_step))

    tb.SummaryWriter = SummaryWriter
    sys.modules['torch.utils.tensorboard'] = tb

    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    sys.modules['models.archs.dcn.deform_conv_cuda'] = types.ModuleType('deform_conv_cuda_stub')
    import models.archs.dcn  # noqa: F401
    dc = sys.modules['models.archs.dcn.deform_conv']
    dc.modulated_deform_conv = lambda x, off, m, w, b, s, p, d, g, dg: torchvision.ops.deform_conv2d(
        x, off, w, b, stride=s, padding=p, dilation=d, mask=m)

    cwd = os.getcwd()
    os.chdir(SCRATCH)                      # any relative path the script may use resolves inside the scratch directory
    import train_dynavsr as T

    # ---- output paths into the scratch directory
    real_parse = T.option.parse

    def parse(path, is_train=True, exp_name=None):
        opt = real_parse(path, is_train=is_train, exp_name=exp_name)
        exp = os.path.join(SCRATCH, 'experiments', opt['name'])
        opt['path'].update(root=SCRATCH, experiments_root=exp, models=os.path.join(exp, 'models'),
                           training_state=os.path.join(exp, 'training_state'), log=exp,
                           val_images=os.path.join(exp, 'val_images'))
        return opt

    T.option.parse = parse

    # ---- synthetic tasks (one clip per task): LQs [B, N, 3, h, w], GT [B, N, 3, s*h, s*w], SuperLQs [B, N, 3, h/s, w/s]
    g = torch.Generator().manual_seed(1234)
    batches = []
    for _ in range(ITERS):
        batches.append({'LQs': torch.rand(BATCH, NFRAMES, 3, H, W, generator=g),
                        'GT': torch.rand(BATCH, NFRAMES, 3, H * SCALE, W * SCALE, generator=g),
                        'SuperLQs': torch.rand(BATCH, NFRAMES, 3, H // SCALE, W // SCALE, generator=g)})
    created, snapshots = [], []

    def state():
        model, est = created[0]
        return ({k: v.detach().clone() for k, v in model.netG.module.state_dict().items()},
                {k: v.detach().clone() for k, v in est.netE.module.state_dict().items()})

    class Loader(object):
        def __len__(self):
            return ITERS * BATCH

        def __iter__(self):
            for i, b in enumerate(batches):
                if i:
                    snapshots.append(state())          # weights after outer iteration i
                yield b
            snapshots.append(state())
            raise _Stop()

    T.loader.get_dataset = lambda opt, train=True: list(range(ITERS * BATCH))
    T.create_dataloader = lambda dataset, dataset_opt, opt=None, sampler=None: Loader()
    T.create_dataset = lambda dataset_opt, **kw: []

    real_create = T.create_model
    from oracle import params as P

    def create_model(opt):
        models = real_create(opt)
        if not created:
            # deterministic, non-degenerate weights (the default zero-initialised offset convs would hide the DCN)
            for net, seed in ((models[0].netG.module, 21), (models[1].netE.module, 22)):
                shapes = OrderedDict((k, tuple(v.shape)) for k, v in net.state_dict().items())
                net.load_state_dict(P.make_params(shapes, seed))
            created.append(models)
            created.append(state())
        return models

    T.create_model = create_model

    real_to = torch.Tensor.to

    def to(self, *a, **k):
        if a and isinstance(a[0], str) and a[0].startswith('cuda'):
            a = ('cpu',) + a[1:]
        return real_to(self, *a, **k)

    torch.Tensor.to = to
    argv = sys.argv
    sys.argv = ['train_dynavsr.py', '-opt', yml]
    try:
        T.main()
        raise SystemExit('the reference loop ended without reaching the stop marker')
    except _Stop:
        pass
    finally:
        sys.argv = argv
        torch.Tensor.to = real_to
        os.chdir(cwd)

    sd0_G, sd0_E = created[1]
    assert len(snapshots) == ITERS
    flat = lambda sd: torch.cat([v.reshape(-1) for v in sd.values()]).numpy()
    if sgd:                                # start weights and the batch are those of meta_loop.npz (same seeds)
        out = {'G1': flat(snapshots[0][0]), 'E1': flat(snapshots[0][1]), 'lr_G': np.float32(1.0)}
    else:
        out = {'cfg': np.array([NF, NFRAMES, GROUPS, FRONT, BACK, NF_E, SCALE, BATCH, ITERS, INNER])}
        for i, b in enumerate(batches):
            for k, v in b.items():          # only the centre GT frame is read by the loop (train_dynavsr.py:288)
                out['it%d_%s' % (i, k)] = (v[:, NFRAMES // 2] if k == 'GT' else v).numpy()
        out['keys_G'] = np.array(list(sd0_G.keys()))
        out['keys_E'] = np.array(list(sd0_E.keys()))
        out['G0'], out['E0'] = flat(sd0_G), flat(sd0_E)
        for i, (sg, se) in enumerate(snapshots):
            out['G%d' % (i + 1)], out['E%d' % (i + 1)] = flat(sg), flat(se)
    out['train_loss'] = np.array([v for t, v, _ in scalars if t == 'Train loss'])
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, 'meta_loop_sgd.npz' if sgd else 'meta_loop.npz'), **out)
    print('train loss per outer iteration', out['train_loss'])
    print('parameters: G %d, E %d' % (flat(sd0_G).size, flat(sd0_E).size))
    shutil.rmtree(SCRATCH, ignore_errors=True)


if __name__ == '__main__':
    main()
