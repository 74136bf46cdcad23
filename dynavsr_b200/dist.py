"""Clip / frame sharding across the GPUs of one box (one process per GPU, torch.distributed).

The hot path shards with NO data-path collective: every output frame restarts from the same
meta-weights and reads only its own 5-frame window (test_dynavsr.py:208), so ranks simply take
strided frame indices exactly as the reference's distributed validation does
(train_dynavsr.py:509) and reduce the per-frame PSNR vectors to rank 0 at the end (:722-728).
Meta-training shards clips with ``DistIterSampler`` semantics (data/data_sampler.py:46-59) and needs
ONE all-reduce of the flat meta-gradient per outer step.
"""
import math

import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world_size):
    """train_dynavsr.py:509 -- ``for idx in range(rank, len(val_set_frag), world_size)``."""
    return list(range(rank, n_items, world_size))


def dist_iter_sampler_indices(dataset_len, num_replicas, rank, epoch=0, ratio=100):
    """Index stream of the reference's DistIterSampler (data/data_sampler.py:30-59): epoch-seeded
    randperm over the enlarged index space, modulo the dataset size, rank-strided."""
    num_samples = int(math.ceil(dataset_len * ratio / num_replicas))
    total = num_samples * num_replicas
    g = torch.Generator()
    g.manual_seed(epoch)
    idx = [v % dataset_len for v in torch.randperm(total, generator=g).tolist()]
    idx = idx[rank:total:num_replicas]
    assert len(idx) == num_samples
    return idx


def reduce_metric_vectors(vectors, dst=0):
    """train_dynavsr.py:722-728: every rank fills only its own frames of each per-folder vector (zeros
    elsewhere); ``dist.reduce`` sums them onto rank ``dst``; a barrier closes the phase."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for v in vectors:
            dist.reduce(v, dst)
        dist.barrier()
    return vectors


def allreduce_flat_gradient(flat_grad, average=True):
    """The single exchange step of meta-training: sum (then average) the flat EDVR+MFDN meta-gradient
    across ranks before the outer optimiser step (NCCL over NVLink on GPU, gloo in CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        if average:
            flat_grad.div_(dist.get_world_size())
    return flat_grad
