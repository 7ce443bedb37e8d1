#!/usr/bin/env python
"""Launch single tensor-core convolutions / normalisation kernels for an `ncu --set full` capture: every spec is warmed
up twice and then launched ONCE between cudaProfilerStart/Stop.
  ncu --profile-from-start off --set full --clock-control none --import-source on -o out \
      python tools/ncu_shapes.py KIND:N,H,W,Cin,Cout[,k] [...]      KIND in fwd | stats | dgrad | wgrad | norm, or the
CUDA-core kernels of layers with <= 8 channels on one side: sfwd | sdgrad | swgrad (that side float32, the other bf16)"""
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
rt = torch.cuda.cudart()
for spec in sys.argv[1:]:
    kind, dims = spec.split(':')
    dims = [int(v) for v in dims.split(',')]
    N, H, W, Cin, Cout = dims[:5]
    k = dims[5] if len(dims) > 5 else 3
    if kind in ('sfwd', 'sdgrad', 'swgrad'):
        dt = lambda c: torch.float32 if c <= 8 else torch.bfloat16
        xs = torch.randn(N, H, W, Cin, device='cuda').to(dt(Cin))
        ys = torch.randn(N, H, W, Cout, device='cuda').to(dt(Cout))
        wm = torch.randn(k, k, Cin, Cout, device='cuda') * 0.05
        dwm = torch.zeros(k, k, Cin, Cout, device='cuda')

        def run():
            if kind == 'sfwd':
                call('phs_conv2d', call.T(xs), wm, None, call.T(ys), k, 0, 0, L.IMPL_SIMT)
            elif kind == 'sdgrad':      # input gradient: reads ys (dy), writes xs (dx)
                call('phs_conv2d', call.T(ys), wm, None, call.T(xs), k, 1, 0, L.IMPL_SIMT)
            else:
                call('phs_conv2d_wgrad', call.T(xs), call.T(ys), dwm, None, k, 1, L.IMPL_SIMT)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        rt.cudaProfilerStart()
        run()
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
        call.keep.clear()
        continue
    x = torch.randn(N, H, W, Cin, device='cuda').to(torch.bfloat16)
    dy = torch.randn(N, H, W, Cout, device='cuda').to(torch.bfloat16)
    w = (torch.randn(Cout, 9 * Cin, device='cuda') * 0.05).to(torch.bfloat16)
    wd = (torch.randn(Cin, 9 * Cout, device='cuda') * 0.05).to(torch.bfloat16)
    y = torch.empty(N, H, W, Cout, device='cuda', dtype=torch.bfloat16)
    gx = torch.empty(N, H, W, Cin, device='cuda', dtype=torch.bfloat16)
    dw = torch.zeros(3, 3, Cin, Cout, device='cuda')
    stats = torch.zeros(N + 1, Cout, 2, device='cuda', dtype=torch.float64)
    mean = torch.zeros(N * Cout, device='cuda'); rstd = torch.ones(N * Cout, device='cuda')
    gamma = torch.ones(Cout, device='cuda'); beta = torch.zeros(Cout, device='cuda')
    coef = torch.zeros(N * Cout * 2, device='cuda'); sums = torch.zeros(N * Cout * 2, device='cuda', dtype=torch.float64)
    a = torch.empty_like(dy)

    def run():
        if kind == 'fwd':
            call('phs_conv2d', call.T(x), w, None, call.T(y), 3, 0, 0, L.IMPL_TC)
        elif kind == 'stats':
            call('phs_conv2d_stats_acc', call.T(x), w, None, call.T(y), 3, stats)
        elif kind == 'dgrad':
            call('phs_conv2d', call.T(dy), wd, None, call.T(gx), 3, 1, 0, L.IMPL_TC)
        elif kind == 'wgrad':
            call('phs_conv2d_wgrad', call.T(x), call.T(dy), dw, None, 3, 1, L.IMPL_TC)
        elif kind == 'norm':
            call('phs_norm_act_fwd_stats', call.T(dy), stats, L.NORM_BN_TRAIN, 1e-3, 0.99, None, None, mean, rstd, gamma, beta, 1, call.T(a))
            call('phs_norm_bwd_reduce', call.T(a), call.T(dy), mean, rstd, gamma, beta, 1, sums)
            call('phs_norm_bwd_apply', call.T(a), call.T(dy), mean, rstd, gamma, beta, 1, coef, call.T(y))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    run()
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
    call.keep.clear()
