"""CPU: known-answer tests pinning the oracle (oracle/phiseg_oracle.py).

The reference ships no tests, golden vectors or checkpoints and TensorFlow 1.12 cannot run here ("parity unpinned",
SURVEY.md section 8c), so every TF-semantics assumption of the oracle is pinned by a hand-computed value or an
independent formula below, and the whole graph by tests/golden/oracle_golden.json (regression fixture)."""
import json
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))


def test_legacy_bilinear_hand_values(oracle):
    # TF1 resize_bilinear, align_corners=False, no half-pixel centres: src = dst * 0.5
    x = torch.tensor([[1.0, 3.0], [5.0, 11.0]]).reshape(1, 2, 2, 1)
    want = torch.tensor([[1, 2, 3, 3], [3, 5, 7, 7], [5, 8, 11, 11], [5, 8, 11, 11]], dtype=torch.float32)
    assert torch.equal(oracle.bilinear_upsample2d(x).reshape(4, 4), want)
    # it is NOT torch's half-pixel bilinear
    other = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode='bilinear', align_corners=False)
    assert not torch.allclose(other.reshape(4, 4), want)
    # generic formula on a longer row
    r = torch.arange(5.0).reshape(1, 1, 5, 1) ** 2
    out = oracle.bilinear_upsample2d(r.expand(1, 2, 5, 1))[0, 0, :, 0]
    src = np.arange(10) * 0.5
    lo = np.floor(src).astype(int)
    hi = np.minimum(lo + 1, 4)
    ref = (1 - (src - lo)) * (lo ** 2) + (src - lo) * (hi ** 2)
    assert np.allclose(out.numpy(), ref)


def test_nearest_and_avgpool(oracle):
    x = torch.arange(4.0).reshape(1, 2, 2, 1)
    up = oracle.nearest_upsample(x, 2).reshape(4, 4)
    assert torch.equal(up, torch.tensor([[0, 0, 1, 1], [0, 0, 1, 1], [2, 2, 3, 3], [2, 2, 3, 3]], dtype=torch.float32))
    y = torch.arange(16.0).reshape(1, 4, 4, 1)
    assert torch.equal(oracle.averagepool2d(y).reshape(2, 2), torch.tensor([[2.5, 4.5], [10.5, 12.5]]))


def test_conv_same_hand_values(oracle):
    # 3x3 ones filter on a 3x3 ones image, SAME zero padding: counts of in-bounds neighbours
    x = torch.ones(1, 3, 3, 1)
    w = torch.ones(3, 3, 1, 1)
    y = oracle.conv2d_same(x, w).reshape(3, 3)
    assert torch.equal(y, torch.tensor([[4., 6, 4], [6, 9, 6], [4, 6, 4]]))
    # HWIO orientation: filter tap (kh=0,kw=2) multiplies the pixel up-right (cross-correlation, no flip)
    w2 = torch.zeros(3, 3, 1, 1)
    w2[0, 2] = 1
    img = torch.arange(9.0).reshape(1, 3, 3, 1)
    y2 = oracle.conv2d_same(img, w2).reshape(3, 3)
    assert torch.equal(y2, torch.tensor([[0., 0, 0], [1, 2, 0], [4, 5, 0]]))


def test_batch_norm_train_infer_and_moving_stats(oracle):
    x = torch.tensor([1.0, 2.0, 3.0, 6.0]).reshape(4, 1, 1, 1)
    P = {'s/BatchNorm/gamma': torch.tensor([2.0]), 's/BatchNorm/beta': torch.tensor([0.5]),
         's/BatchNorm/moving_mean': torch.tensor([0.0]), 's/BatchNorm/moving_variance': torch.tensor([1.0])}
    ns = {}
    y = oracle.batch_norm(x, P, 's', True, ns)
    mean, var = 3.0, 3.5            # biased variance of (1,2,3,6)
    ref = (x - mean) / math.sqrt(var + 1e-3) * 2 + 0.5
    assert torch.allclose(y, ref, atol=1e-6)
    assert abs(float(ns['s/BatchNorm/moving_mean']) - 0.01 * 3.0) < 1e-7
    assert abs(float(ns['s/BatchNorm/moving_variance']) - (0.99 + 0.01 * 3.5 * 4 / 3)) < 1e-7   # Bessel corrected
    yi = oracle.batch_norm(x, P, 's', False)
    assert torch.allclose(yi, x / math.sqrt(1 + 1e-3) * 2 + 0.5, atol=1e-6)


def test_group_norm_grouping(oracle):
    # C=32 -> G=max(2, 32//16)=2 groups of 16 contiguous channels; C=8 -> G=2 groups of 4
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 3, 32, generator=g)
    P = {'s/gamma': torch.ones(1, 1, 1, 32), 's/beta': torch.zeros(1, 1, 1, 32)}
    y = oracle.group_norm2d(x, P, 's')
    for n in range(2):
        for grp in range(2):
            blk = x[n, :, :, grp * 16:(grp + 1) * 16]
            ref = (blk - blk.mean()) / torch.sqrt(blk.var(unbiased=False) + 1e-5)
            assert torch.allclose(y[n, :, :, grp * 16:(grp + 1) * 16], ref, atol=1e-5)
    x8 = torch.randn(1, 2, 2, 8, generator=g)
    y8 = oracle.group_norm2d(x8, {'s/gamma': torch.ones(1, 1, 1, 8), 's/beta': torch.zeros(1, 1, 1, 8)}, 's')
    blk = x8[0, :, :, 4:8]
    assert torch.allclose(y8[0, :, :, 4:8], (blk - blk.mean()) / torch.sqrt(blk.var(unbiased=False) + 1e-5), atol=1e-5)
    # 192 channels -> 12 groups of 16
    assert max(2, 192 // 16) == 12


def test_kl_closed_form_vs_distributions(oracle):
    g = torch.Generator().manual_seed(1)
    mu0, mu1 = torch.randn(3, 4, 4, 2, generator=g, dtype=torch.float64), torch.randn(3, 4, 4, 2, generator=g, dtype=torch.float64)
    s0 = torch.rand(3, 4, 4, 2, generator=g, dtype=torch.float64) + 0.2
    s1 = torch.rand(3, 4, 4, 2, generator=g, dtype=torch.float64) + 0.2
    kl = oracle.Oracle.KL_two_gauss_with_diag_cov(mu0, s0, mu1, s1)
    q = torch.distributions.Normal(mu0, s0)
    p = torch.distributions.Normal(mu1, s1)
    ref = torch.distributions.kl_divergence(q, p).reshape(3, -1).sum(1).mean()
    assert abs(float(kl) - float(ref)) < 1e-6 * float(ref)      # the reference adds 1e-10 inside the logs and the quotient


def test_multinoulli_and_residual_accumulation(oracle):
    orc = oracle.Oracle('phiseg', image_size=(64, 64, 1), latent_levels=2, resolution_levels=7, KL_weight=None)
    s = torch.tensor([[[0, 1]]], dtype=torch.uint8)                       # [1,1,2]
    l0 = torch.tensor([[[[0.0, 0.0], [2.0, 0.0]]]])
    l1 = torch.tensor([[[[1.0, 0.0], [0.0, 3.0]]]])
    ld = orc.losses(s, [l0, l1], None, None, None, None)
    # level 1: xent(l1); level 0: xent(l0 + l1); sum over pixels, mean over batch
    lse = lambda a, b: math.log(math.exp(a) + math.exp(b))
    want1 = (lse(1, 0) - 1) + (lse(0, 3) - 3)
    want0 = (lse(1, 0) - 1) + (lse(2, 3) - 3)
    assert abs(float(ld['residual_multinoulli_loss_lvl1']) - want1) < 1e-6
    assert abs(float(ld['residual_multinoulli_loss_lvl0']) - want0) < 1e-6
    assert abs(float(ld['total_loss']) - (want0 + want1)) < 1e-6


def test_tf_adam_one_step_hand_values(oracle):
    orc = oracle.Oracle('phiseg', image_size=(64, 64, 1), latent_levels=1)
    # TF: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m=(1-b1)g; v=(1-b2)g^2; first step moves every weight by ~lr*sign(g)
    g, lr = 0.5, 1e-3
    m, v = 0.1 * g, 0.001 * g * g
    lr_t = lr * math.sqrt(1 - 0.999) / (1 - 0.9)
    step = lr_t * m / (math.sqrt(v) + 1e-8)
    assert abs(step - lr) < 1e-9          # |first Adam step| == lr up to eps-hat


def test_he_normal_truncated(oracle):
    rng = np.random.default_rng(0)
    w = oracle.he_normal(rng, (3, 3, 64, 64))
    std = math.sqrt(1.3 * 2.0 / (9 * 64))
    assert np.abs(w).max() <= 2 * std + 1e-12
    assert abs(w.std() / std - 0.8796) < 0.01     # std of a +-2 sigma truncated normal


def test_topology_flops_and_params(oracle):
    """live conv FLOPs of the oracle's graph == SURVEY.md section 8d (25.03 / 22.04 GFLOP per image forward)"""
    def flops(orc):
        tot = 0
        names = {n: s for n, s, k in orc.spec.entries if k == 'W'}
        return names
    o = oracle.Oracle('phiseg')
    n_tr = sum(int(np.prod(s)) for n, s, k in o.spec.entries if k in ('W', 'b', 'gamma', 'beta'))
    assert n_tr == 18.7e6 or abs(n_tr - 18.71e6) < 0.05e6          # incl. dead branches
    o2 = oracle.Oracle('probunet', latent_levels=1, zdim0=6)
    n2 = sum(int(np.prod(s)) for n, s, k in o2.spec.entries if k in ('W', 'b', 'gamma', 'beta'))
    assert abs(n2 - 19.04e6) < 0.05e6


def test_oracle_gradients_fp64_finite_difference(oracle):
    """autograd of the oracle graph (stands for tf.gradients) against a central finite difference, fp64"""
    orc = oracle.Oracle('phiseg', image_size=(64, 64, 1), n0=4, norm='group_norm', dtype=torch.float64)
    orc.init_params(seed=3)
    x, s = oracle.synthetic_batch(2, 64, 64, 2, seed=1)
    eps = [torch.tensor(e, dtype=torch.float64) for e in oracle.synthetic_eps(orc.latent_shapes(2), seed=2)]
    xt, st = torch.tensor(x), torch.tensor(s)
    out, g, _ = orc.grads(xt, st, eps)
    for name, idx in (('likelihood/post_c_0_2/W', (1, 1, 2, 3)), ('posterior/z2_input_1/W', (0, 2, 1, 0)),
                      ('prior/z4_sigma/b', (1,)), ('posterior/z0_pre_1/group_norm/gamma', (0, 0, 0, 2))):
        h = 1e-7
        base = orc.P[name][idx].item()
        orc.P[name][idx] = base + h
        lp = float(orc.forward_train(xt, st, eps).loss_dict['total_loss'])
        orc.P[name][idx] = base - h
        lm = float(orc.forward_train(xt, st, eps).loss_dict['total_loss'])
        orc.P[name][idx] = base
        fd = (lp - lm) / (2 * h)
        assert abs(fd - float(g[name][idx])) <= 1e-2 * max(1.0, abs(fd)), (name, fd, float(g[name][idx]))


def test_golden_fixture(oracle):
    """regression fixture made by tests/golden/make_golden.py from this oracle (seeded numpy inputs/weights)"""
    path = os.path.join(HERE, 'golden', 'oracle_golden.json')
    gold = json.load(open(path))
    from golden.make_golden import compute
    now = compute(oracle)
    for case, vals in gold.items():
        for k, v in vals.items():
            got = now[case][k]
            assert np.allclose(got, v, rtol=1e-5, atol=1e-7), (case, k)
