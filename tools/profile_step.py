#!/usr/bin/env python
"""Per-launch timing of one training step (eager replay of the static launch program, CUDA events around every C-ABI
call, median of --reps): where the step time goes, per kernel and per convolution shape.  Not a bench value."""
import argparse
import collections
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='phiseg_7_5')
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--mode', default='fast')
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    import torch
    from __graft_entry__ import load_package
    load_package()
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
    exp = ex.load_experiment(ex.experiment_path(args.config))
    model = pm.phiseg(exp, mode=args.mode, use_cuda_graph=False)
    import importlib as _il; o = _il.import_module("phiseg_code_b200.data")
    H = model.cfg.H
    x, s = o.synthetic_batch(args.batch, H, H, model.cfg.nlabels, seed=1)
    for _ in range(2):
        model.training_step(x, s, 1e-3)
    sp = model._program('train', args.batch)
    st = torch.cuda.current_stream().cuda_stream
    steps = [s_ for s_ in sp.prog.steps if s_[0] is not None]
    times = [[] for _ in steps]
    for rep in range(args.reps):
        evs = []
        for fn, a, name in steps:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a, st)
            e1.record()
            assert rc == 0, name
            evs.append((e0, e1))
        torch.cuda.synchronize()
        for i, (e0, e1) in enumerate(evs):
            times[i].append(e0.elapsed_time(e1) * 1e3)
    med = [sorted(t)[len(t) // 2] for t in times]

    def desc(a):
        t = a._obj
        return '%dx%dx%dx%d%s' % (t.N, t.H, t.W, t.C, 'b' if t.dtype == 1 else 'f')

    by_kernel = collections.defaultdict(lambda: [0.0, 0])
    by_shape = collections.defaultdict(lambda: [0.0, 0, 0.0])
    for (fn, a, name), t in zip(steps, med):
        by_kernel[name][0] += t
        by_kernel[name][1] += 1
        if name in ('phs_conv2d', 'phs_conv2d_wgrad', 'phs_conv2d_stats', 'phs_conv2d_stats_acc'):
            if name in ('phs_conv2d_stats', 'phs_conv2d_stats_acc'):
                xd, yd, k, impl = a[0]._obj, a[3]._obj, a[4], 1
                kind = 'fwd+stats'
                cin, cout = xd.C, yd.C
            elif name == 'phs_conv2d':
                xd, yd, k, dgrad, impl = a[0]._obj, a[3]._obj, a[4], a[5], a[7]
                kind = 'dgrad' if dgrad else 'fwd'
                cin, cout = xd.C, yd.C
            else:
                xd, yd, k, impl = a[0]._obj, a[1]._obj, a[4], a[6]
                kind = 'wgrad'
                cin, cout = xd.C, yd.C
            fl = 2.0 * xd.N * xd.H * xd.W * k * k * cin * cout
            key = '%s %s %dx%d %d->%d k%d' % ('TC' if impl else 'SIMT', kind, xd.H, xd.W, cin, cout, k)
            by_shape[key][0] += t
            by_shape[key][1] += 1
            by_shape[key][2] += fl
    # memory-bound kernels by tensor shape: achieved GB/s against the passes they make
    passes = {'phs_norm_act_fwd': (0, 2), 'phs_norm_bwd_reduce': (0, 2), 'phs_norm_bwd_apply': (0, 3), 'phs_chan_stats': (0, 1),
              'phs_upsample2_fwd': (1, 1.25), 'phs_upsample2_bwd': (0, 1.25), 'phs_avgpool2_fwd': (0, 1.25),
              'phs_avgpool2_bwd': (1, 1.25)}
    by_mem = collections.defaultdict(lambda: [0.0, 0, 0.0])
    for (fn, a, name), t in zip(steps, med):
        if name in passes:
            ai, np_ = passes[name]
            td = a[ai]._obj
            nbytes = td.N * td.H * td.W * td.C * (2 if td.dtype == 1 else 4) * np_
            key = '%s %dx%d C=%d' % (name, td.H, td.W, td.C)
            by_mem[key][0] += t
            by_mem[key][1] += 1
            by_mem[key][2] += nbytes
    total = sum(med)
    print('total of per-launch medians: %.3f ms over %d launches (B=%d)' % (total / 1e3, len(steps), args.batch))
    print('\n-- by C-ABI entry point')
    for k, (t, n) in sorted(by_kernel.items(), key=lambda kv: -kv[1][0]):
        print('%-26s %5d launches %10.1f us  %5.1f%%' % (k, n, t, 100 * t / total))
    print('\n-- convolutions by shape (top %d)' % args.top)
    for k, (t, n, fl) in sorted(by_shape.items(), key=lambda kv: -kv[1][0])[:args.top]:
        print('%-40s x%-2d %9.1f us %5.1f%%  %7.1f TFLOP/s' % (k, n, t, 100 * t / total, fl / t / 1e6))
    print('\n-- memory-bound kernels by shape (top %d)' % args.top)
    for k, (t, n, nb) in sorted(by_mem.items(), key=lambda kv: -kv[1][0])[:args.top]:
        print('%-44s x%-2d %9.1f us %5.1f%%  %7.0f GB/s' % (k, n, t, 100 * t / total, nb / t / 1e3))
    if args.out:
        with open(args.out, 'w') as fh:
            json.dump({'total_us': total, 'by_kernel': {k: v for k, v in by_kernel.items()},
                       'by_shape': {k: v for k, v in by_shape.items()}}, fh, indent=1)


if __name__ == '__main__':
    main()
