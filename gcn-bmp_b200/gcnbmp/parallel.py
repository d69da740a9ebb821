"""Data-parallel plumbing: the pair batch shards across ranks with no data-path collective;
training has exactly one exchange step per iteration -- an allreduce(SUM) of the flat fp32
gradient buffer (the loss is normalised by the GLOBAL element count, so the sum of the shard
gradients equals the single-GPU gradient).  Replaces chainer.training.ParallelUpdater
(train_binary.py:546-549: two hard-wired devices, per-parameter addgrads)."""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous split of n pairs; the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_sum_(flat, group=None):
    """In-place SUM allreduce of one flat buffer (NCCL on GPUs, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_(flat, src=0, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat
