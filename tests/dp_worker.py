"""Worker of tests/test_gpu_multi.py (run under torchrun, one rank per GPU): data-parallel training step of the real
engine against the same step on ONE GPU with the concatenated batch.  Under group
norm every image is independent, so  sum_r grad_r / world == grad(single GPU, world * b images)  up to summation order."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package   # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else 'parity'
    tol = float(sys.argv[2]) if len(sys.argv) > 2 else 2e-5
    load_package()
    par = importlib.import_module('phiseg_code_b200.parallel')
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
    D = importlib.import_module('phiseg_code_b200.data')
    rank, world, local = par.init_from_env()
    assert world >= 2
    torch.cuda.set_device(local)
    exp = ex.load_experiment(ex.experiment_path('phiseg_7_5_gn'))
    size, b = 64, 2
    exp.image_size = (size, size, 1)
    exp.weight_decay_weight = 1e-5          # exercises the "added after the all-reduce, not summed" rule
    x, s = D.synthetic_batch(world * b, size, size, 2, seed=3)
    dp = pm.phiseg(exp, mode=mode, use_cuda_graph=True, seed=11)
    eps = D.synthetic_eps(dp.cfg.latent_shapes(world * b), seed=5)
    sl = par.shard_slice(world * b, rank, world)
    losses = []
    for it in range(3):                      # eager, capture + replay, replay: the collectives live inside the graph
        losses.append(dp.training_step(x[sl], s[sl], lr=0.0, eps=[e[sl] for e in eps]))
    assert dp._program('train', b).graphs, 'data-parallel step was not captured into a CUDA graph'
    print('rank %d: 3 data-parallel steps done (dp_mode=%s)' % (rank, dp.dp_mode), flush=True)
    assert abs(losses[0] - losses[2]) <= 1e-6 * abs(losses[0]), losses
    g_dp = dp.params.g.detach().clone() / world
    loss_sum = torch.tensor([losses[2]], dtype=torch.float64, device='cuda')
    torch.distributed.all_reduce(loss_sum)
    ok = 1
    if rank == 0:
        # same process group around, but this replica trains alone on the whole batch (its constructor must not issue the
        # weight broadcast the data-parallel replicas do: rank 1 is not there to take part)
        ref = pm.phiseg(exp, mode=mode, use_cuda_graph=False, seed=11, data_parallel=False)
        ref.set_weights(dp.get_weights())
        l_ref = ref.training_step(x, s, lr=0.0, eps=eps)
        g_ref = ref.params.g
        gmax = float(g_ref.abs().max())
        err = float((g_dp - g_ref).abs().max()) / gmax
        # weight decay appears once in either total (it is not a mean over replicas): compare the data terms
        wd = ref.loss_dict['weight_decay']
        l_dp = (float(loss_sum.item()) - world * wd) / world + wd
        lerr = abs(l_dp - l_ref) / abs(l_ref)
        print('dp-equivalence[%s] world=%d: grad err / max|g| = %.3e, loss rel err %.3e (loss %.4f)' % (mode, world, err, lerr, l_ref))
        ok = int(err <= tol and lerr <= tol)
    flag = torch.tensor([ok], device='cuda')
    torch.distributed.broadcast(flag, 0)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == '__main__':
    main()
