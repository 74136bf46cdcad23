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
kw = dict(inner_steps=1, lr_alpha=1e-3, inner_optimizer='SGD', criterion='l2', lr_outer=1e-2)
theta0 = MetaLearner(*nets(), exchange='none', **kw).theta.clone()
for outer in ('SGD', 'Adam'):
    # single-process truth: both tasks locally, no collective.  Per-task scaling 1/B with B = world equals B = 1 per rank averaged
    # over the ranks.
    ref = MetaLearner(*nets(), outer_optimizer=outer, exchange='none', **kw)
    for _ in range(2):
        ref.outer_step(tasks)
    for exch in ('peer', 'peer-all', 'nccl'):
        ml = MetaLearner(*nets(), outer_optimizer=outer, exchange=exch, **kw)
        for _ in range(2):
            ml.outer_step([tasks[rank]])          # sharded: one task per rank, one exchange per outer step
        torch.cuda.synchronize()
        err = float(((ml.theta - theta0) - (ref.theta - theta0)).norm() / (ref.theta - theta0).norm())
        gathered = [torch.zeros_like(ml.theta) for _ in range(world)]
        dist.all_gather(gathered, ml.theta)
        drift = max(float((t - ml.theta).abs().max()) for t in gathered)
        if rank == 0:
            print('meta step x2, %d ranks, outer %s, exchange requested %s -> used %s: update vs single-process rel %.3e; rank drift %.3e'
                  % (world, outer, exch, ml.exchange, err, drift), flush=True)
            assert err < (5e-3 if outer == 'SGD' else 5e-2) and drift == 0.0
        dist.barrier()
# timing of the exchange + update alone (EDVR-M + MFDN flat buffer), both paths
ml_p = MetaLearner(*nets(), outer_optimizer='Adam', exchange='peer', **kw)
ml_n = MetaLearner(*nets(), outer_optimizer='Adam', exchange='nccl', **kw)
import ctypes  # noqa: E402
from dynavsr_b200._lib import call  # noqa: E402
from dynavsr_b200 import dist as dd  # noqa: E402


def t_peer():
    # the sliced (reduce-scatter) fused kernel: reduce + Adam on this rank's slice, new weights written to every rank's buffer
    h = ml_p._peer
    h.barrier(channel=0)
    call('dvsr_update_peers_sliced', ctypes.c_void_p(ml_p.theta.data_ptr()), ctypes.c_void_p(int(h.buffer_ptrs_dev)), world, rank, 0, 1.0 / world,
         ctypes.c_void_p(ml_p.m.data_ptr()), ctypes.c_void_p(ml_p.v.data_ptr()), ml_p.theta.numel(), ml_p.theta.numel(), 1e-5, 1e-5, 0.9, 0.99,
         1e-8, 0.1, 0.01, 0.0, 1, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    h.barrier(channel=1)
    ml_p.theta.copy_(ml_p.meta_grad)


def t_peer_all():
    h = ml_p._peer
    h.barrier(channel=0)
    call('dvsr_update_peers', ctypes.c_void_p(ml_p.theta.data_ptr()), ctypes.c_void_p(int(h.buffer_ptrs_dev)), world, 0, 1.0 / world,
         ctypes.c_void_p(ml_p.m.data_ptr()), ctypes.c_void_p(ml_p.v.data_ptr()), ml_p.theta.numel(), ml_p.theta.numel(), 1e-5, 1e-5, 0.9, 0.99,
         1e-8, 0.1, 0.01, 0.0, 1, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    h.barrier(channel=1)


def t_nccl():
    dd.allreduce_flat_gradient(ml_n.meta_grad, average=True)
    call('dvsr_update_adam', ctypes.c_void_p(ml_n.theta.data_ptr()), ctypes.c_void_p(ml_n.meta_grad.data_ptr()), ctypes.c_void_p(ml_n.m.data_ptr()),
         ctypes.c_void_p(ml_n.v.data_ptr()), ml_n.theta.numel(), ml_n.theta.numel(), 1e-5, 1e-5, 0.9, 0.99, 1e-8, 0.1, 0.01, 0.0,
         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))


for name, fn in (('peer-memory fused exchange+Adam (sliced)', t_peer), ('peer-memory fused exchange+Adam (read-all)', t_peer_all),
                 ('NCCL all-reduce + Adam launch', t_nccl)):
    if name.startswith('peer') and ml_p._peer is None:
        continue
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        fn()
    b.record()
    torch.cuda.synchronize()
    tt = torch.tensor([a.elapsed_time(b) / 50 * 1e3], device='cuda')
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print('%-44s %7.1f us per outer exchange (%.1f MB flat gradient, max over ranks)' % (name, float(tt), ml_p.theta.numel() * 4 / 1e6), flush=True)
dist.destroy_process_group()
