#!/usr/bin/env python
"""Ablation of conv_halo_kernel through its PHS_HALO_DBG switches (1 = no TMA, 2 = no MMA, 4 = no epilogue stores):
which of the three pipelines bounds each layer shape.  CUDA events, median of 20 launches.  Not a bench value."""
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
SHAPES = [(64, 128, 128, 128, 128), (64, 128, 128, 32, 32), (64, 64, 64, 64, 64), (64, 64, 64, 192, 192),
          (64, 128, 128, 64, 128), (64, 128, 128, 192, 32), (64, 32, 32, 128, 128), (64, 16, 16, 192, 192)]
FLAGS = [0, 1, 2, 4, 3, 5, 6, 7]
print('shape              ' + ''.join('  dbg=%d ' % f for f in FLAGS) + '   (us; fwd+stats)')
for (N, H, W, Cin, Cout) in SHAPES:
    x = torch.randn(N, H, W, Cin, device='cuda').to(torch.bfloat16)
    w = (torch.randn(Cout, 9 * Cin, device='cuda') * 0.05).to(torch.bfloat16)
    y = torch.empty(N, H, W, Cout, device='cuda', dtype=torch.bfloat16)
    stats = torch.zeros(N, Cout, 2, device='cuda', dtype=torch.float64)
    out = '%3dx%-3d %3d->%-3d ' % (H, W, Cin, Cout)
    for stat in (True, False):
        for f in FLAGS:
            os.environ['PHS_HALO_DBG'] = str(f)
            def run():
                if stat:
                    call('phs_conv2d_stats', call.T(x), w, None, call.T(y), 3, stats)
                else:
                    call('phs_conv2d', call.T(x), w, None, call.T(y), 3, 0, 0, L.IMPL_TC)
            for _ in range(3):
                run()
            ts = []
            for _ in range(20):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            call.keep.clear()
            out += ' %7.1f' % sorted(ts)[len(ts) // 2]
        out += '  |'
    print(out)
os.environ['PHS_HALO_DBG'] = '0'
