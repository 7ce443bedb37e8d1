import sys, importlib, numpy as np, torch
sys.path.insert(0,'/root/repo')
from __graft_entry__ import load_package, load_oracle
load_package(); o=load_oracle()
pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
for name in ('phiseg_7_5','phiseg_7_5_gn'):
  for mode in ('fast','parity'):
    exp = ex.load_experiment(ex.experiment_path(name))
    m = pm.phiseg(exp, mode=mode, use_cuda_graph=True)
    x,s = o.synthetic_batch(16,128,128,2,seed=1)
    ls=[m.training_step(x,s,1e-3) for _ in range(12)]
    print(name, mode, ['%.4g'%l for l in ls])
