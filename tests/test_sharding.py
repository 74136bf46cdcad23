"""CPU tests of the N>1 path: frame sharding, DistIterSampler semantics and the two collectives,
with the gloo backend and world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dynavsr_b200 import dist as D


def test_shard_indices_partition():
    for n in (0, 1, 7, 100):
        for world in (1, 2, 8):
            parts = [D.shard_indices(n, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_dist_iter_sampler_semantics():
    # restated from data/data_sampler.py:46-59: deterministic per epoch, disjoint cover of the enlarged index space
    a0 = D.dist_iter_sampler_indices(10, 2, 0, epoch=3, ratio=5)
    a1 = D.dist_iter_sampler_indices(10, 2, 1, epoch=3, ratio=5)
    assert len(a0) == len(a1) == 25
    assert a0 == D.dist_iter_sampler_indices(10, 2, 0, epoch=3, ratio=5)
    assert a0 != D.dist_iter_sampler_indices(10, 2, 0, epoch=4, ratio=5)
    g = torch.Generator(); g.manual_seed(3)
    perm = [v % 10 for v in torch.randperm(50, generator=g).tolist()]
    assert a0 == perm[0::2] and a1 == perm[1::2]
    counts = [0] * 10
    for v in a0 + a1:
        counts[v] += 1
    assert counts == [5] * 10


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_frames):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    psnr = torch.zeros(n_frames)
    for i in D.shard_indices(n_frames, rank, world):
        psnr[i] = 30.0 + i                         # this rank's frames only
    D.reduce_metric_vectors([psnr], dst=0)
    if rank == 0:
        assert torch.equal(psnr, 30.0 + torch.arange(n_frames, dtype=torch.float32))
    g = torch.full((1000,), float(rank + 1))
    D.allreduce_flat_gradient(g, average=True)
    assert torch.allclose(g, torch.full((1000,), (1 + 2) / 2.0))
    # evaluation tables (dynavsr_b200/driver.py): every rank holds its own frames' rows, rank 0 ends up with all of them
    from collections import OrderedDict
    from dynavsr_b200 import driver
    rows = OrderedDict(('clip/%08d' % i, [28.0 + i, 31.0 + i, float('nan'), 0.9]) for i in D.shard_indices(n_frames, rank, world))
    merged = driver.gather_rows(rows, dst=0)
    if rank == 0:
        assert list(merged) == ['clip/%08d' % i for i in range(n_frames)]
        assert all(merged['clip/%08d' % i][1] == 31.0 + i for i in range(n_frames))
    else:
        assert merged is None
    dist.destroy_process_group()


def test_two_rank_sharding_and_collectives():
    mp.spawn(_worker, args=(2, _free_port(), 11), nprocs=2, join=True)
