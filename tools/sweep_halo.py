#!/usr/bin/env python
"""Geometry sweep of conv_halo_kernel (PHS_HALO_CTAS / PHS_HALO_S / PHS_HALO_ACC) with an optional in-kernel timeline
(PHS_HALO_TRACE).  CUDA-event medians over 20 launches with a ~fixed host enqueue overhead subtracted by timing a
back-to-back burst of 10 launches instead of single ones.  Not a bench value."""
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
          (64, 128, 128, 64, 128), (64, 128, 128, 192, 32), (64, 128, 128, 32, 192), (64, 32, 32, 128, 128),
          (64, 16, 16, 192, 192), (64, 16, 16, 64, 64), (64, 32, 32, 256, 192)]
CFGS = [('2', None, None, '0'), ('2', None, None, None), ('2', None, None, '32'), ('1', None, None, None), ('1', '2', None, None),
        ('1', '4', None, None)]
if os.environ.get('SWEEP_SHAPES'):      # e.g. SWEEP_SHAPES=64,128,128,128,128;64,64,64,192,192
    SHAPES = [tuple(int(v) for v in sh.split(',')) for sh in os.environ['SWEEP_SHAPES'].split(';')]
if os.environ.get('SWEEP_CFGS'):
    CFGS = [tuple(None if v == '-' else v for v in c.split('/')) for c in os.environ['SWEEP_CFGS'].split(',')]
trace = torch.zeros(8 * 3 * 256, dtype=torch.int64, device='cuda')
do_trace = len(sys.argv) > 1 and sys.argv[1] == 'trace'


def setenv(k, v):
    if v is None:
        os.environ.pop(k, None)
    else:
        os.environ[k] = v


def dump_trace(tag):
    t = trace.cpu().view(8, 3, 256)
    for cta in (0, 5):
        base = int(t[cta, 0, 0])
        for role, nm in enumerate(('prod', 'mma ', 'epi ')):
            v = [int(x) - base for x in t[cta, role] if int(x) != 0]
            print('   trace %s cta%d %s: %s' % (tag, cta, nm, ' '.join(str(x) for x in v[:40])))


print('shape               ' + ''.join(' c%s/S%s/a%s/G%s' % (c, s or '-', a or '-', g or '-') for c, s, a, g in CFGS) + '   (us per launch, burst of 10; stats | no stats)')
for (N, H, W, Cin, Cout) in SHAPES:
    x = torch.randn(N, H, W, Cin, device='cuda').to(torch.bfloat16)
    w = (torch.randn(Cout, 9 * Cin, device='cuda') * 0.05).to(torch.bfloat16)
    y = torch.empty(N, H, W, Cout, device='cuda', dtype=torch.bfloat16)
    stats = torch.zeros(N, Cout, 2, device='cuda', dtype=torch.float64)
    out = '%3dx%-3d %3d->%-3d ' % (H, W, Cin, Cout)
    for stat in (True, False):
        for (c, s, a, g) in CFGS:
            setenv('PHS_HALO_CTAS', c); setenv('PHS_HALO_S', s); setenv('PHS_HALO_ACC', a); setenv('PHS_HALO_G', g)
            def run():
                if stat:
                    return call.rc('phs_conv2d_stats', call.T(x), w, None, call.T(y), 3, stats)
                return call.rc('phs_conv2d', call.T(x), w, None, call.T(y), 3, 0, 0, L.IMPL_TC)
            rc = run()
            if rc != 0:
                out += '        n/a'
                continue
            for _ in range(2):
                run()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(10):
                    run()
                e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e2)
            call.keep.clear()
            out += ' %10.1f' % sorted(ts)[len(ts) // 2]
            if do_trace and not stat:
                trace.zero_()
                os.environ['PHS_HALO_TRACE'] = hex(trace.data_ptr())
                run(); torch.cuda.synchronize()
                os.environ.pop('PHS_HALO_TRACE')
                print('%3dx%-3d %3d->%-3d cfg c%s/S%s/a%s/G%s' % (H, W, Cin, Cout, c, s, a, g))
                dump_trace('')
        out += '  |'
    print(out)
