"""tools/meta_bench.py -- BASELINE config 4 (diagnostic, not the bench.py contract): meta-training outer steps of EDVR-L
(nf = 128, back_RBs = 40) + MFDN, per task LR 5x3x64x64 / HR 3x256x256 / SLR 5x3x16x16, inner Adam K = 1, Charbonnier loss,
one task per rank per outer step, exchange of the flat meta-gradient fused with the outer Adam update over peer memory.

    python tools/meta_bench.py [--steps K] [--tasks-per-rank T] [--nf 128 --back 40]      (torchrun for N > 1)
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from dynavsr_b200 import ops  # noqa: E402
from dynavsr_b200.meta import MetaLearner  # noqa: E402
from dynavsr_b200.models.archs import EDVR_arch, LRimg_estimator  # noqa: E402
from dynavsr_b200.synth import seed_parameters  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--tasks-per-rank', type=int, default=1)
ap.add_argument('--nf', type=int, default=128)
ap.add_argument('--back', type=int, default=40)
ap.add_argument('--exchange', default='peer')
args = ap.parse_args()
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local))
ops.set_conv_backend(True)
netG = seed_parameters(EDVR_arch.EDVR(nf=args.nf, nframes=5, groups=8, front_RBs=5, back_RBs=args.back, scale=4), 1).cuda()
netE = seed_parameters(LRimg_estimator.DirectKernelEstimatorVideo(64, 3, 4), 2).cuda()
ml = MetaLearner(netG, netE, inner_steps=1, lr_alpha=1e-5, inner_optimizer='Adam', criterion='cb', outer_optimizer='Adam',
                 lr_outer=1e-5, exchange=args.exchange)
g = torch.Generator().manual_seed(10 + rank)
tasks = [{'LQs': torch.rand(1, 5, 3, 64, 64, generator=g).cuda(), 'GT': torch.rand(1, 3, 256, 256, generator=g).cuda(),
          'SuperLQs': torch.rand(1, 5, 3, 16, 16, generator=g).cuda()} for _ in range(args.tasks_per_rank)]
for _ in range(args.warmup):
    ml.outer_step(tasks)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.steps):
    ml.outer_step(tasks)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b)
if world > 1:
    t = torch.tensor([ms], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
if rank == 0:
    n_params = ml.theta.numel()
    print(json.dumps({'metric': 'meta-training tasks/s (config 4, diagnostic)', 'value': world * args.tasks_per_rank * args.steps / (ms / 1e3),
                      'unit': 'tasks/s', 'n_gpus': world, 'ms_per_outer_step': ms / args.steps, 'exchange': ml.exchange,
                      'config': {'model': 'EDVR(nf=%d, back_RBs=%d) + MFDN' % (args.nf, args.back), 'flat_params': n_params,
                                 'flat_gradient_MB': n_params * 4 / 1e6, 'task': 'LR 5x3x64x64 -> HR 3x256x256, SLR 5x3x16x16, inner Adam K=1, cb loss',
                                 'tasks_per_rank_per_outer_step': args.tasks_per_rank, 'precision': 'tcgen05 (BF16x3 / TF32), fp32 accumulate'},
                      'loss_q': float(ml.last['loss_q'].mean())}))
if world > 1:
    dist.destroy_process_group()
