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
for (N, H, W, Cin, Cout) in [(64, 128, 128, 128, 128), (64, 128, 128, 32, 32), (64, 64, 64, 64, 64)]:
    x = torch.randn(N, H, W, Cin, device='cuda').to(torch.bfloat16)
    w = (torch.randn(Cout, 9 * Cin, device='cuda') * 0.05).to(torch.bfloat16)
    y = torch.empty(N, H, W, Cout, device='cuda', dtype=torch.bfloat16)
    stats = torch.zeros(N, Cout, 2, device='cuda', dtype=torch.float64)
    for _ in range(2):
        call('phs_conv2d_stats', call.T(x), w, None, call.T(y), 3, stats)
    torch.cuda.synchronize()
    s = stats.flatten()[:10].cpu().tolist()
    tiles = s[5]
    print('%dx%d %d->%d: MMA thread total %.0f cyc for %d tiles (%.0f/tile): wait tempty %.0f, a_full %.0f, b_full %.0f, issue+commit %.0f | epilogue warp: wait tfull %.0f, work %.0f (per tile)'
          % (H, W, Cin, Cout, s[0], tiles, s[0] / tiles, s[1] / tiles, s[2] / tiles, s[3] / tiles, s[4] / tiles, s[8] / tiles / 2, s[9] / tiles / 2))
