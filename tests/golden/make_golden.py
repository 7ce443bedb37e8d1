"""Generates tests/golden/oracle_golden.json: losses, gradient norms and logit samples of the CPU oracle on seeded
synthetic inputs (numpy Generator streams are stable across versions).  The reference itself cannot run here
(TensorFlow 1.12 absent) so these vectors pin the *oracle* against regressions; the GPU parity
tests compare the CUDA path with the same quantities.

  python tests/golden/make_golden.py       # rewrites the fixture
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    'phiseg_gn_64_n8': dict(arch='phiseg', norm='group_norm', n0=8, latent_levels=5, zdim0=2, nlabels=2),
    'phiseg_bn_64_n8': dict(arch='phiseg', norm='batch_norm', n0=8, latent_levels=5, zdim0=2, nlabels=2),
    'probunet_bn_64_n8': dict(arch='probunet', norm='batch_norm', n0=8, latent_levels=1, zdim0=6, nlabels=2),
    'phiseg_gn_64_n8_4cls': dict(arch='phiseg', norm='group_norm', n0=8, latent_levels=5, zdim0=2, nlabels=4),
}


def compute(oracle):
    out = {}
    for name, kw in CASES.items():
        orc = oracle.Oracle(image_size=(64, 64, 1), dtype=torch.float64, **kw)
        orc.init_params(seed=11)
        B = 2
        x, s = oracle.synthetic_batch(B, 64, 64, kw['nlabels'], seed=12)
        eps = [torch.tensor(e, dtype=torch.float64) for e in oracle.synthetic_eps(orc.latent_shapes(B), seed=13)]
        res, g, _ = orc.grads(torch.tensor(x), torch.tensor(s), eps)
        smp = orc.forward_sample(torch.tensor(x), eps, training=False)
        vals = {k: float(v.detach() if torch.is_tensor(v) else v) for k, v in res.loss_dict.items()}
        vals['grad_l2'] = float(sum((t.double() ** 2).sum() for t in g.values() if t is not None) ** 0.5)
        vals['sample_logits_probe'] = smp.s_out_eval[:, ::16, ::16, :].reshape(-1).tolist()
        vals['sample_argmax_sum'] = int(smp.s_out_eval.argmax(-1).sum())
        out[name] = vals
    return out


if __name__ == '__main__':
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from __graft_entry__ import load_oracle
    data = compute(load_oracle())
    with open(os.path.join(HERE, 'oracle_golden.json'), 'w') as fh:
        json.dump(data, fh, indent=1)
    print('wrote', len(data), 'cases')
