#!/usr/bin/env python
"""Single-layer timing of the tensor-core convolutions through the C-ABI (CUDA events, median of 5 bursts of 50
back-to-back launches; the working set of every shape but the smallest exceeds L2 only partly: these are tuning numbers,
not bench values)."""
import os, sys, ctypes
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
kinds = sys.argv[1].split(',') if len(sys.argv) > 1 else ['fwd', 'stats', 'wgrad']
for (N, H, W, Cin, Cout) in SHAPES:
    x = torch.randn(N, H, W, Cin, device='cuda').to(torch.bfloat16)
    dy = torch.randn(N, H, W, Cout, device='cuda').to(torch.bfloat16)
    w = (torch.randn(Cout, 9 * Cin, device='cuda') * 0.05).to(torch.bfloat16)
    y = torch.empty(N, H, W, Cout, device='cuda', dtype=torch.bfloat16)
    dw = torch.zeros(3, 3, Cin, Cout, device='cuda')
    stats = torch.zeros(N, Cout, 2, device='cuda', dtype=torch.float64)
    fl = 2.0 * N * H * W * 9 * Cin * Cout
    out = '%3dx%-3d %3d->%-3d ' % (H, W, Cin, Cout)
    for kind in kinds:
        def run():
            if kind == 'fwd':
                call('phs_conv2d', call.T(x), w, None, call.T(y), 3, 0, 0, L.IMPL_TC)
            elif kind == 'stats':
                call('phs_conv2d_stats', call.T(x), w, None, call.T(y), 3, stats)
            else:
                call('phs_conv2d_wgrad', call.T(x), call.T(dy), dw, None, 3, 1, L.IMPL_TC)
        # back-to-back bursts: single short launches on an idle GPU are timed at ramping clocks and with the launch
        # latency inside the event pair; a burst of 50 shows the steady-state time the kernel has inside a step
        for _ in range(20):
            run()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                run()
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 50)
        call.keep.clear()
        t = sorted(ts)[len(ts) // 2]
        out += ' %s %7.1f us %6.1f TF' % (kind, t, fl / t / 1e6)
    print(out)
