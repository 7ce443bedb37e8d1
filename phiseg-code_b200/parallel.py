"""Data-parallel plumbing: one process per GPU, contiguous batch shards, identical replicas,
ONE all-reduce(sum) over the flat gradient buffer per step; the 1/world scaling is applied inside the fused
optimizer kernel (phs_adam_step grad_scale).  The reference has no distributed path at all."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun); no-op for a single process."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return 0, 1, 0
    rank = int(os.environ['RANK'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_slice(global_batch, rank, world):
    """Contiguous, equal shards; the loss is a mean over the batch (phiseg_model.py:221,236) so averaging the
    replica gradients of equal shards reproduces the large-batch gradient (exactly under group norm)."""
    if global_batch % world:
        raise ValueError('global batch %d is not divisible by %d ranks' % (global_batch, world))
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)


def allreduce_sum_(flat):
    """The single collective of the data path."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
