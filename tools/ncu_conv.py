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
kind = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
for (N, H, W, Cin, Cout) in [(64, 128, 128, 128, 128), (64, 128, 128, 32, 32), (64, 64, 64, 192, 192)]:
    x = torch.randn(N, H, W, Cin, device='cuda').to(torch.bfloat16)
    dy = torch.randn(N, H, W, Cout, device='cuda').to(torch.bfloat16)
    w = (torch.randn(Cout, 9 * Cin, device='cuda') * 0.05).to(torch.bfloat16)
    y = torch.empty(N, H, W, Cout, device='cuda', dtype=torch.bfloat16)
    dw = torch.zeros(3, 3, Cin, Cout, device='cuda')
    stats = torch.zeros(N, Cout, 2, device='cuda', dtype=torch.float64)
    for _ in range(2):
        if kind == 'fwd':
            call('phs_conv2d', call.T(x), w, None, call.T(y), 3, 0, 0, L.IMPL_TC)
        elif kind == 'stats':
            call('phs_conv2d_stats', call.T(x), w, None, call.T(y), 3, stats)
        else:
            call('phs_conv2d_wgrad', call.T(x), call.T(dy), dw, None, 3, 1, L.IMPL_TC)
    torch.cuda.synchronize()
