#!/usr/bin/env python
"""Which buffers of a fast-mode training step differ between two fresh, identical runs?  Prints, per run pair, the loss,
the gradient tensors that differ most, and the first differing program buffers with the launches that touch them.
  python tools/diag_determinism.py [exp] [size] [B]        (PHS_NO_LANES=1 for single-stream programs)
DIAG_ENV_A / DIAG_ENV_B = "K=V K=V": environment switches of the first / second run (e.g. PHS_FUSE_NORM=0 vs 1: which
buffers of the fused program differ from the unfused one's - the two programs allocate the same buffers in the same order)."""
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
size = int(sys.argv[2]) if len(sys.argv) > 2 else 128
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8


def run(env=''):
    for kv in env.split():
        k, v = kv.split('=', 1)
        os.environ[k] = v
    exp = ex.load_experiment(ex.experiment_path(name))
    exp.image_size = (size, size, 1)
    model = pm.phiseg(exp, mode='fast', use_cuda_graph=False, seed=7)
    x, s = D.synthetic_batch(B, size, size, model.cfg.nlabels, seed=3)
    eps = D.synthetic_eps(model.cfg.latent_shapes(B), seed=5)
    loss = model.training_step(x, s, lr=0.0, eps=eps)
    torch.cuda.synchronize()
    sp = model._program('train', B)
    bufs = [t for t in sp.prog.keep if torch.is_tensor(t)]
    return model, sp, loss, bufs


m1, sp1, l1, b1 = run(os.environ.get('DIAG_ENV_A', ''))
m2, sp2, l2, b2 = run(os.environ.get('DIAG_ENV_B', ''))
print('lanes=%s  loss %.8f vs %.8f  rel %.2e' % (os.environ.get('PHS_NO_LANES') is None, l1, l2, abs(l1 - l2) / abs(l1)))
g1, g2 = m1.params.g, m2.params.g
gmax = float(g1.abs().max())
rows = []
for n, (off, shape, kind) in m1.params.table.items():
    cnt = 1
    for d in shape:
        cnt *= d
    d = float((g1[off:off + cnt] - g2[off:off + cnt]).abs().max())
    if d > 0:
        rows.append((d / gmax, d / max(float(g1[off:off + cnt].abs().max()), 1e-30), n))
rows.sort(reverse=True)
print('%d of %d gradient tensors differ; worst (diff/max|g|, diff/max|g_tensor|, name):' % (len(rows), len(m1.params.table)))
for r in rows[:12]:
    print('   %.3e  %.3e  %s' % r)
# program buffers, in allocation order (= forward order, gradient twins are allocated while the backward is laid down)
ptr_steps = {}
for i, st in enumerate(sp1.prog.steps):
    fn, args, nm = st
    if fn is None:
        continue
    for a in args:
        t = getattr(a, '_obj', None)
        p = t.ptr if (t is not None and hasattr(t, 'ld')) else (a if isinstance(a, int) else None)
        if p:
            ptr_steps.setdefault(p, []).append('%d:%s@L%d' % (i, nm, getattr(st, 'lane', 0)))
assert len(b1) == len(b2)
nd = 0
for k, (t1, t2) in enumerate(zip(b1, b2)):
    a, b = t1.float().nan_to_num(0.0, 0.0, 0.0), t2.float().nan_to_num(0.0, 0.0, 0.0)
    d = float((a - b).abs().max())
    if d > 0:
        nd += 1
        if nd <= 25:
            base = t1.data_ptr()
            touch = []
            for p, lst in ptr_steps.items():
                if base <= p < base + t1.numel() * t1.element_size():
                    touch += lst
            touch.sort(key=lambda v: int(v.split(':')[0]))
            touch2 = []
            base2 = t2.data_ptr()
            for i, st in enumerate(sp2.prog.steps):
                if st[0] is None:
                    continue
                for a_ in st[1]:
                    o = getattr(a_, '_obj', None)
                    p_ = o.ptr if (o is not None and hasattr(o, 'ld')) else (a_ if isinstance(a_, int) else None)
                    if p_ and base2 <= p_ < base2 + t2.numel() * t2.element_size():
                        touch2.append('%d:%s@L%d' % (i, st[2], getattr(st, 'lane', 0)))
            print('buf %4d %-22s %-8s maxdiff %.3e (max %.3e)  launches: %s  ||  B: %s' % (k, tuple(t1.shape), str(t1.dtype)[6:], d, float(a.abs().max()), ' '.join(touch[:6]), ' '.join(touch2[:6])))
print('%d of %d buffers differ; n_fwd=%d of %d steps' % (nd, len(b1), sp1.n_fwd, len(sp1.prog.steps)))
