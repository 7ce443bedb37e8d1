#!/usr/bin/env python
"""Timing of the HBM-bound small-channel head kernels (1x1 heads forward / input gradient / filter gradient) on the step's
shapes: CUDA events around a burst of 10 launches.  Not a bench value."""
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
for (N, H, C) in [(64, 128, 128), (64, 64, 192), (64, 32, 192), (64, 128, 32)]:
    x = torch.randn(N, H, H, C, device='cuda').to(torch.bfloat16)
    y2 = torch.randn(N, H, H, 2, device='cuda')
    w = torch.randn(1, 1, C, 2, device='cuda')
    dw = torch.zeros(1, 1, C, 2, device='cuda')
    gx = torch.empty_like(x)
    nbytes = x.numel() * 2
    out = '%3dx%-3d C=%-3d ' % (H, H, C)
    for name, fn in (('head fwd', lambda: call('phs_conv2d', call.T(x), w, None, call.T(y2), 1, 0, 0, L.IMPL_SIMT)),
                     ('head dgrad', lambda: call('phs_conv2d', call.T(y2), w, None, call.T(gx), 1, 1, 0, L.IMPL_SIMT)),
                     ('head wgrad', lambda: call('phs_conv2d_wgrad', call.T(x), call.T(y2), dw, None, 1, 1, L.IMPL_SIMT))):
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
        out += ' %s %6.1f us %5.0f GB/s |' % (name, t, nbytes / t / 1e3)
    print(out)
