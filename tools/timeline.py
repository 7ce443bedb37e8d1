#!/usr/bin/env python
"""Kernel timeline of one CUDA-graph replay of the training step (torch.profiler / CUPTI activity records):
per-kernel start, duration and stream, written as JSON for offline analysis (tools/timeline_report.py)."""
import argparse, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='phiseg_7_5')
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--out', default='gpurun_out/timeline.json')
    args = ap.parse_args()
    import torch
    from torch.profiler import profile, ProfilerActivity
    from __graft_entry__ import load_package
    load_package()
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
    exp = ex.load_experiment(ex.experiment_path(args.config))
    model = pm.phiseg(exp, mode='fast', use_cuda_graph=True)
    import importlib as _il; o = _il.import_module("phiseg_code_b200.data")
    x, s = o.synthetic_batch(args.batch, model.cfg.H, model.cfg.W, model.cfg.nlabels, seed=1)
    for _ in range(5):
        model.training_step(x, s, 1e-3)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            model.training_step(x, s, 1e-3)
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type is not None and str(e.device_type).endswith('CUDA'):
            ev.append({'name': e.name, 'ts': e.time_range.start, 'dur': e.time_range.end - e.time_range.start,
                       'stream': getattr(e, 'device_index', 0)})
    # kineto events carry the stream in the chrome trace; export that as well
    prof.export_chrome_trace(args.out.replace('.json', '_chrome.json'))
    with open(args.out, 'w') as fh:
        json.dump(ev, fh)
    print(len(ev), 'device events')


if __name__ == '__main__':
    main()
