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
N, H, W = 64, 128, 128
cases = [('wgrad', 3, 32, 3, torch.bfloat16, torch.bfloat16), ('wgrad', 1, 32, 3, torch.float32, torch.bfloat16),
         ('wgrad', 128, 2, 1, torch.bfloat16, torch.float32), ('fwd', 3, 32, 3, torch.bfloat16, torch.bfloat16),
         ('dgrad', 2, 128, 1, torch.float32, torch.bfloat16)]
for kind, Cin, Cout, k, dx, dy_t in cases:
    x = torch.randn(N, H, W, Cin, device='cuda').to(dx)
    dy = torch.randn(N, H, W, Cout, device='cuda').to(dy_t)
    w = torch.randn(k, k, Cin, Cout, device='cuda')
    dw = torch.zeros(k, k, Cin, Cout, device='cuda')
    ts = []
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if kind == 'wgrad':
            call('phs_conv2d_wgrad', call.T(x), call.T(dy), dw, None, k, 1, L.IMPL_SIMT)
        elif kind == 'fwd':
            call('phs_conv2d', call.T(x), w, None, call.T(dy), k, 0, 0, L.IMPL_SIMT)
        else:   # dgrad of a head: input dy-like tensor with Cin channels ... here x plays dy (Cin=2), output has Cout
            wt = torch.randn(k, k, Cout, Cin, device='cuda')
            call('phs_conv2d', call.T(x), wt, None, call.T(dy), k, 1, 0, L.IMPL_SIMT)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    print(kind, Cin, Cout, k, 'us:', ['%.1f' % t for t in ts])
