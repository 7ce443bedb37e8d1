#!/usr/bin/env python
"""One eager, single-stream training step inside cudaProfilerStart/Stop with an NVTX range around every C-ABI call, for
  ncu --profile-from-start off --nvtx --print-nvtx-rename kernel --metrics gpu__time_duration.sum --clock-control none \
      --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py
The step list (index, entry point, tensor shapes) is also written to --out so launches can be joined by order."""
import argparse, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='phiseg_7_5')
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--out', default='gpurun_out/step_list.json')
    ap.add_argument('--kind', default='train', choices=['train', 'sample'],
                    help="'sample': one pass of the sampling program (prior encoder + one noise draw of 64 rows) instead")
    args = ap.parse_args()
    os.environ['PHS_NO_LANES'] = '1'
    import ctypes
    import torch
    from __graft_entry__ import load_package
    load_package()
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
    exp = ex.load_experiment(ex.experiment_path(args.config))
    model = pm.phiseg(exp, mode='fast', use_cuda_graph=False)
    import importlib as _il; o = _il.import_module("phiseg_code_b200.data")
    x, s = o.synthetic_batch(args.batch, model.cfg.H, model.cfg.W, model.cfg.nlabels, seed=1)
    if args.kind == 'sample':
        model.predict(x, num_samples=2)
        sp = model._program('sample', args.batch)
    else:
        for _ in range(2):
            model.training_step(x, s, 1e-3)
        sp = model._program('train', args.batch)
    st = torch.cuda.current_stream().cuda_stream
    steps = [s_ for s_ in sp.prog.steps if s_[0] is not None]

    def label(i, a, name):
        out = []
        for v in a:
            t = getattr(v, '_obj', None)
            if t is not None and hasattr(t, 'ld'):
                out.append('%dx%dx%dx%d%s' % (t.N, t.H, t.W, t.C, 'b' if t.dtype == 1 else 'f'))
            elif isinstance(v, int) and abs(v) < 1000:
                out.append(str(v))
        return '%04d|%s|%s' % (i, name, ','.join(out))

    labels = [label(i, a, name) for i, (fn, a, name) in enumerate(steps)]
    with open(args.out, 'w') as fh:
        json.dump(labels, fh, indent=0)
    torch.cuda.synchronize()
    cudart = torch.cuda.cudart()
    cudart.cudaProfilerStart()
    for (fn, a, name), lab in zip(steps, labels):
        torch.cuda.nvtx.range_push(lab)
        rc = fn(*a, st)
        torch.cuda.nvtx.range_pop()
        assert rc == 0, lab
    torch.cuda.synchronize()
    cudart.cudaProfilerStop()


if __name__ == '__main__':
    main()
