#!/usr/bin/env python
"""One eager fast-mode (tcgen05) training step + one sampling pass on a tiny problem, for compute-sanitizer:
  compute-sanitizer --tool {memcheck,racecheck,initcheck,synccheck} python tools/sanitize_step.py [exp] [size] [B]
PHS_NO_LANES=1 makes every launch serial (races between lanes then show up as a difference to the default run)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_package
load_package()
pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
D = importlib.import_module('phiseg_code_b200.data')
name = sys.argv[1] if len(sys.argv) > 1 else 'phiseg_7_5'
size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
exp = ex.load_experiment(ex.experiment_path(name))
exp.image_size = (size, size, 1)
model = pm.phiseg(exp, mode='fast', use_cuda_graph=False)
x, s = D.synthetic_batch(B, size, size, model.cfg.nlabels, seed=3)
eps = D.synthetic_eps(model.cfg.latent_shapes(B), seed=5)
loss = model.training_step(x, s, lr=1e-3, eps=eps)
seg = model.predict_segmentation_sample(x, eps=eps)
torch.cuda.synchronize()
print('sanitize_step %s %dx%d B=%d: loss %.6f, launches %d, mask sum %d' % (name, size, size, B, loss, model.gpu_launches, int(seg.sum())))
