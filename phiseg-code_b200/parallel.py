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


COMM_LANE = 20


def gradient_buckets(params):
    """Contiguous ranges [lo, hi) (in floats) of the flat gradient buffer, grouped by WHEN their gradients become final
    during the backward pass: the likelihood's top-down merge and heads (the first and largest part of the backward),
    its per-level towers, the latent hierarchy of prior / posterior, their 192-channel encoder levels (most of the
    parameters, final ~1 ms before the end) and finally the 32..128-channel encoder levels (few parameters, the last
    launches of the step: the only all-reduce that is not hidden behind compute)."""
    items = sorted(params.table.items(), key=lambda kv: kv[1][0])

    def group(name):
        net, rest = name.split('/', 1)
        scope = rest.split('/', 1)[0]
        if net == 'likelihood':
            if scope in ('encoder', 'decoder'):
                return 'likelihood-' + scope
            tower = scope.startswith('z') or scope.startswith('preups_')
            return 'likelihood-towers' if tower else 'likelihood-merge'
        enc = None
        if scope.startswith('z') and '_pre_' in scope:            # phiseg encoders: z{r}_pre_{t}
            enc = int(scope[1:scope.index('_')])
        elif scope.startswith('conv_'):                           # prob. U-Net encoders: conv_{r}_{t}
            enc = int(scope.split('_')[1])
        if enc is None:
            return net + '-latent'
        return net + ('-enc-lo' if enc <= 2 else '-enc-hi')

    buckets = []
    for name, (off, shape, kind) in items:
        n = 1
        for d in shape:
            n *= int(d)
        hi = off + (n + 3) // 4 * 4
        g = group(name)
        if buckets and buckets[-1][0] == g:
            buckets[-1][2] = hi
        else:
            buckets.append([g, off, hi])
    return [(g, lo, hi) for g, lo, hi in buckets]


def insert_gradient_allreduce(bwd_steps, params, world, allreduce=None):
    """Returns the backward launch list with one all-reduce(sum) per gradient bucket inserted right behind the last launch
    that writes into the bucket, on a communication lane that waits for exactly the lanes that wrote into it: the
    collective of a bucket runs (NCCL over NVLink, inside the captured graph) while the rest of the backward pass is
    still computing.  Writers are found by scanning every launch's arguments for pointers into the gradient buffer.
    allreduce(tensor) -> step callable; default: torch.distributed.all_reduce on the launch's stream."""
    from .engine import Step
    g = params.g
    base, nbytes = g.data_ptr(), g.numel() * 4
    buckets = gradient_buckets(params)
    last = {}          # bucket index -> (position of the last writer, lanes of all writers)
    for i, st in enumerate(bwd_steps):
        fn, args, name = st
        if fn is None:
            continue
        for a in args:
            if isinstance(a, int) and base <= a < base + nbytes:
                off = (a - base) // 4
                for bi, (_, lo, hi) in enumerate(buckets):
                    if lo <= off < hi:
                        pos, lanes = last.get(bi, (-1, set()))
                        lanes.add(getattr(st, 'lane', 0))
                        last[bi] = (i, lanes)
                        break

    def make(t):
        if allreduce is not None:
            return allreduce(t)

        def run(stream):
            with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return 0
        return run

    out = list(bwd_steps)
    # insert from the back so that earlier positions stay valid; buckets nobody writes (dead branches only) are skipped
    for bi, (pos, lanes) in sorted(last.items(), key=lambda kv: -kv[1][0]):
        _, lo, hi = buckets[bi]
        ins = []
        for ln in sorted(lanes):
            dep = Step((None, ((ln, COMM_LANE),), 'after'))
            dep.lane = 0
            ins.append(dep)
        ar = Step((make(g[lo:hi]), (), 'allreduce[%s]' % buckets[bi][0]))
        ar.lane = COMM_LANE
        ins.append(ar)
        out[pos + 1:pos + 1] = ins
    join = Step((None, ((COMM_LANE,),), 'join'))
    join.lane = 0
    out.append(join)
    return out


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
