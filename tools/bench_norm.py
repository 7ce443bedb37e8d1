#!/usr/bin/env python
"""Timing of the HBM-bound normalisation kernels on the step's main tensor shapes (CUDA events around a burst of 10
launches; achieved GB/s against the passes each kernel has to make).  Not a bench value."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from __graft_entry__ import load_package
load_package()
import importlib
L = importlib.import_module('phiseg_code_b200.lib')
from gpu_util import Caller
call = Caller(L)
SHAPES = [(64, 128, 128, 128), (64, 128, 128, 32), (64, 64, 64, 64), (64, 64, 64, 192), (64, 32, 32, 128), (64, 16, 16, 192)]
for (N, H, W, C) in SHAPES:
    y = torch.randn(N, H, W, C, device='cuda').to(torch.bfloat16)
    g = torch.randn(N, H, W, C, device='cuda').to(torch.bfloat16)
    a = torch.empty_like(y); dy = torch.empty_like(y)
    mean = torch.zeros(N * C, device='cuda'); rstd = torch.ones(N * C, device='cuda')
    gamma = torch.ones(C, device='cuda'); beta = torch.zeros(C, device='cuda')
    sums = torch.zeros(N * C * 2, device='cuda', dtype=torch.float64); coef = torch.zeros(N * C * 2, device='cuda')
    nbytes = N * H * W * C * 2
    out = '%3dx%-3d C=%-3d ' % (H, W, C)
    for name, passes, fn in (
            ('act_fwd', 2, lambda: call('phs_norm_act_fwd', call.T(y), mean, rstd, gamma, beta, 1, call.T(a))),
            ('bwd_reduce', 2, lambda: call('phs_norm_bwd_reduce', call.T(g), call.T(y), mean, rstd, gamma, beta, 1, sums)),
            ('bwd_apply', 3, lambda: call('phs_norm_bwd_apply', call.T(g), call.T(y), mean, rstd, gamma, beta, 1, coef, call.T(dy)))):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(10):
                fn()
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e2)
        call.keep.clear()
        t = sorted(ts)[len(ts) // 2]
        out += ' %s %6.1f us %5.0f GB/s |' % (name, t, passes * nbytes / t / 1e3)
    print(out)
