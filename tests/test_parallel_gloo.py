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


def test_shard_slice_errors():
    from __graft_entry__ import load_package
    load_package()
    par = importlib.import_module('phiseg_code_b200.parallel')
    with pytest.raises(ValueError):
        par.shard_slice(10, 0, 4)
    assert par.shard_slice(12, 2, 3) == slice(8, 12)
    assert par.init_from_env() == (0, 1, 0) or os.environ.get('WORLD_SIZE', '1') != '1'
