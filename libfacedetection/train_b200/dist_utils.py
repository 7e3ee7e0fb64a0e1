"""Collectives of the data-parallel step (one process per GPU, ``torch.distributed``).

Reference semantics being reproduced:
  * ``reduce_mean`` of the positive count — mmdet/core/utils/dist_utils.py:68-74, called at
    mmdet/models/dense_heads/yunet_head.py:493-497 (divide by world, all-reduce SUM);
  * DDP gradient averaging — mmdet/apis/train.py:156-161 (sum over ranks, divide by world); here
    ONE all-reduce of the flat gradient bucket, the division folded into the SGD kernel's
    ``grad_scale``.
Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def reduce_mean_(t):
    """In-place mean over ranks of a (1,) tensor; identity for a single process."""
    w = world_size()
    if w == 1:
        return t
    t.div_(w)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_bucket_(bucket):
    """In-place SUM of the flat gradient bucket over ranks; returns the scale (1/world) the
    optimiser applies so that the update uses the mean gradient."""
    w = world_size()
    if w > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
    return 1.0 / w


def shard_batch(global_batch, rank, world):
    """Images [lo, hi) of a global batch owned by ``rank`` (contiguous, equal shards)."""
    if global_batch % world != 0:
        raise ValueError(f'global batch {global_batch} is not divisible by world size {world}')
    per = global_batch // world
    return rank * per, (rank + 1) * per


def init_from_env(backend=None):
    """``(rank, world)`` from the torchrun environment (RANK / WORLD_SIZE / MASTER_*); initialises the
    process group when WORLD_SIZE > 1 (NCCL if CUDA is available, else gloo).  Single process: (0, 1)."""
    import os
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world
