#!/usr/bin/env python
"""A/B timing of the whole training step (phiseg_7_5, B=64, fast mode, CUDA graph) under environment switches, one fresh
model per setting in ONE process.  usage: python tools/step_ab.py "A=1 B=2" "C=3" ...   ('-' = no switch).
Device time of 10 graph replays after 4 warm-up steps (CUDA events); tuning numbers, not bench values."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import faulthandler
import torch
from __graft_entry__ import load_package
if os.environ.get('AB_WATCHDOG'):          # dump the Python stack if a setting has not finished after that many seconds
    faulthandler.dump_traceback_later(int(os.environ['AB_WATCHDOG']), repeat=False, exit=True)
load_package()
pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
D = importlib.import_module('phiseg_code_b200.data')
B = int(os.environ.get('AB_BATCH', '64'))
x, s = D.synthetic_batch(B, 128, 128, 2, seed=1235)
for setting in sys.argv[1:] or ['-']:
    kv = [p.split('=', 1) for p in setting.split() if '=' in p]
    for k, v in kv:
        os.environ[k] = v
    exp = ex.load_experiment(ex.experiment_path(os.environ.get('AB_CONFIG', 'phiseg_7_5')))
    model = pm.phiseg(exp, mode='fast')
    if os.environ.get('AB_KIND') == 'sample':
        # sampling instead of training: predict(x[B], 8) -> host masks (bench.py's `sampling` entry), wall clock
        for _ in range(3):
            model.predict(x, num_samples=8)
        ts = []
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(5):
                m = model.predict(x, num_samples=8)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) / 5 * 1e3)
        print('%-40s  ms/predict(x[%d], 8) %s  mask sum %d' % (setting, B, ' '.join('%.3f' % t for t in ts), int(m.sum())), flush=True)
        for k, v in kv:
            os.environ.pop(k, None)
        del model
        torch.cuda.empty_cache()
        continue
    sp = model._program('train', B)
    model._stage_x(sp, x); model._stage_s(sp, s)
    for _ in range(4):
        model._draw_eps(sp); model._device_step(sp, 1e-3)
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            model._draw_eps(sp); model._device_step(sp, 1e-3)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 10)
    model._read_losses(sp)
    print('%-40s  ms/step %s  loss %.1f' % (setting, ' '.join('%.3f' % t for t in ts), model.loss_tot), flush=True)
    for k, v in kv:
        os.environ.pop(k, None)
    del model, sp
    torch.cuda.empty_cache()
