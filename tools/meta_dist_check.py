"""tools/meta_dist_check.py -- 2-rank check of the meta step's single exchange (torchrun --nproc-per-node 2): every rank runs
MetaLearner.outer_step on ITS task; the all-reduced result must equal a 1-process run over both tasks (same B = 1 per rank
scaling as the reference: loss_q / per-rank batch, then the mean over ranks)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynavsr_b200.meta import MetaLearner  # noqa: E402
from dynavsr_b200.models.archs import EDVR_arch, LRimg_estimator  # noqa: E402
from dynavsr_b200.synth import seed_parameters  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', rank=rank, world_size=world)


def nets():
    g = seed_parameters(EDVR_arch.EDVR(nf=64, front_RBs=2, back_RBs=2), 5).cuda()
    e = seed_parameters(LRimg_estimator.DirectKernelEstimatorVideo(64, 3, 4), 6).cuda()
    return g, e


gen = torch.Generator().manual_seed(9)
tasks = [{'LQs': torch.rand(1, 5, 3, 32, 32, generator=gen).cuda(), 'GT': torch.rand(1, 3, 128, 128, generator=gen).cuda(),
          'SuperLQs': torch.rand(1, 5, 3, 8, 8, generator=gen).cuda()} for _ in range(world)]
kw = dict(inner_steps=1, lr_alpha=1e-3, inner_optimizer='SGD', criterion='l2', outer_optimizer='SGD', lr_outer=1e-2)
ml = MetaLearner(*nets(), **kw)
ml.outer_step([tasks[rank]])                       # sharded: one task per rank, one all-reduce
mine = ml.theta.clone()
# reference on every rank without the collective: both tasks locally.  per-task scaling 1/B with B = 2, and no averaging,
# equals (1/1 per rank) averaged over 2 ranks.
dist.barrier()
import dynavsr_b200.dist as dd  # noqa: E402
orig = dd.allreduce_flat_gradient
dd.allreduce_flat_gradient = lambda g, average=True: g
ml2 = MetaLearner(*nets(), **kw)
ml2.outer_step(tasks)
dd.allreduce_flat_gradient = orig
theta0 = MetaLearner(*nets(), **kw).theta
err = float(((mine - theta0) - (ml2.theta - theta0)).norm() / (ml2.theta - theta0).norm())
all_same = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(all_same, mine)
drift = max(float((t - mine).abs().max()) for t in all_same)
if rank == 0:
    print('meta step, %d ranks: sharded+allreduce vs single-process update rel %.3e; rank drift %.3e' % (world, err, drift))
    assert err < 5e-3 and drift == 0.0
dist.destroy_process_group()
