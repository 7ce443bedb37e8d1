"""Data-parallel plumbing on CPU: two processes, gloo backend.  Checks what the N-GPU path
relies on: contiguous equal shards, ONE all-reduce(sum) over a flat gradient buffer whose 1/world scaling reproduces
the large-batch mean gradient, max-over-ranks timing, and the environment-driven initialisation bench.py uses."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    from __graft_entry__ import load_package
    load_package()
    par = importlib.import_module('phiseg_code_b200.parallel')
    r, w, _ = par.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    # a "model" whose loss is the batch mean of 0.5*||x_i w||^2: gradient = mean_i x_i^T x_i w
    g = np.random.default_rng(0)
    B, D = 8, 5
    X = g.standard_normal((B, 3, D))
    wv = g.standard_normal(D)
    sl = par.shard_slice(B, rank, world)
    local = np.mean([x.T @ (x @ wv) for x in X[sl]], axis=0)          # replica's mean gradient over its shard
    flat = torch.tensor(local)
    par.allreduce_sum_(flat)
    flat /= world                                                      # grad_scale of phs_adam_step
    full = np.mean([x.T @ (x @ wv) for x in X], axis=0)
    out[rank] = (float(np.abs(flat.numpy() - full).max()), par.max_over_ranks(float(rank + 1), 'cpu'), (sl.start, sl.stop))
    torch.distributed.destroy_process_group()


def test_two_rank_gradient_allreduce():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert sorted(out.keys()) == [0, 1]
    for rank in range(world):
        err, mx, sl = out[rank]
        assert err < 1e-12
        assert mx == 2.0
        assert sl == (rank * 4, rank * 4 + 4)


def _bucket_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    from __graft_entry__ import load_package
    load_package()
    par = importlib.import_module('phiseg_code_b200.parallel')
    E = importlib.import_module('phiseg_code_b200.engine')
    par.init_from_env(backend='gloo')
    cfg = E.NetConfig(arch='phiseg', image_size=(64, 64, 1), n0=4, mode='fast')
    P = E.Params(cfg, torch.device('cpu'))
    sp = E.build_program(cfg, P, 2, 'train', torch.device('cpu'))
    bwd = sp.prog.steps[sp.n_fwd:]
    calls = []

    def allreduce(t):
        def run(stream):
            calls.append(t.numel())
            torch.distributed.all_reduce(t)
            return 0
        return run

    steps = par.insert_gradient_allreduce(bwd, P, world, allreduce=allreduce)
    # stand-in for the backward kernels: every replica's local gradient is (rank + 1) everywhere
    P.g.fill_(float(rank + 1))
    seen_after = set()
    for fn, args, name in steps:
        if fn is None and name == 'after':
            seen_after.add(args[0])
        if name.startswith('allreduce'):
            assert any(dst == par.COMM_LANE for _, dst in seen_after)
            fn(0)
    covered = torch.zeros(P.n, dtype=torch.bool)
    for _, lo, hi in par.gradient_buckets(P):
        assert not bool(covered[lo:hi].any()), 'overlapping buckets'
        covered[lo:hi] = True
    assert bool(covered.all())
    out[rank] = (float(P.g.min()), float(P.g.max()), len(calls), sum(calls))
    torch.distributed.destroy_process_group()


def test_bucketed_gradient_allreduce_covers_the_buffer_once():
    """The all-reduce launches that data parallelism inserts into the backward program (parallel.insert_gradient_allreduce):
    run on two gloo ranks over the real phiseg training program's launch list - every float of the flat gradient buffer
    is summed exactly once (1 + 2 = 3 everywhere), in a handful of buckets."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bucket_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for rank in range(world):
        lo, hi, ncalls, nfloats = out[rank]
        assert lo == 3.0 and hi == 3.0, (lo, hi)
        assert 4 <= ncalls <= 10
    assert out[0][3] == out[1][3]


def test_shard_slice_errors():
    from __graft_entry__ import load_package
    load_package()
    par = importlib.import_module('phiseg_code_b200.parallel')
    with pytest.raises(ValueError):
        par.shard_slice(10, 0, 4)
    assert par.shard_slice(12, 2, 3) == slice(8, 12)
    assert par.init_from_env() == (0, 1, 0) or os.environ.get('WORLD_SIZE', '1') != '1'
