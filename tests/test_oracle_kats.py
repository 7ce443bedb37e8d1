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
    """tf.train.AdamOptimizer (phiseg_model.py:136-141), first step, by hand: m = (1-b1) g, v = (1-b2) g^2,
    lr_t = lr sqrt(1-b2)/(1-b1)  =>  delta = lr * g / (|g| + eps_hat / sqrt(1-b2)): every weight with a gradient moves by
    ~lr against the gradient's sign.  Checked on the ORACLE's train_step (which is what the CUDA optimizer is compared to)."""
    orc = oracle.Oracle('phiseg', image_size=(64, 64, 1), n0=4, norm='group_norm', dtype=torch.float64)
    P0 = {k: v.clone() for k, v in orc.init_params(seed=11).items()}
    x, s = oracle.synthetic_batch(2, 64, 64, 2, seed=4)
    eps = [torch.tensor(e, dtype=torch.float64) for e in oracle.synthetic_eps(orc.latent_shapes(2), seed=6)]
    lr = 1e-3
    _, _, g = orc.train_step(torch.tensor(x), torch.tensor(s), eps, lr)
    checked = 0
    for name in ('likelihood/post_c_0_2/W', 'posterior/z0_pre_1/W', 'prior/z4_sigma/b', 'likelihood/y_lvl2/W',
                 'posterior/z2_input_1/group_norm/gamma'):
        gr = g[name]
        want = lr * gr / (gr.abs() + 1e-8 / math.sqrt(1 - 0.999))
        got = P0[name] - orc.P[name]
        assert float((got - want).abs().max()) < 1e-12, name
        big = gr.abs() > 1e-4
        assert float((got[big].abs() - lr).abs().max()) < 2e-6      # |first Adam step| == lr where eps_hat is negligible
        checked += int(big.sum())
    assert checked > 100
    # dead branches (posteriors.py:112-118) receive no gradient and do not move
    dead = 'posterior/z4_ups_to_3_c_1/W'
    assert g[dead] is None and torch.equal(P0[dead], orc.P[dead])
    # second step, by hand from the oracle's own gradients: m2 = b1 m1 + (1-b1) g2, v2 likewise, bias-corrected step size
    P1 = {k: v.clone() for k, v in orc.P.items()}
    _, _, g2 = orc.train_step(torch.tensor(x), torch.tensor(s), eps, lr)
    name = 'likelihood/post_c_0_2/W'
    m2 = 0.9 * 0.1 * g[name] + 0.1 * g2[name]
    v2 = 0.999 * 0.001 * g[name] ** 2 + 0.001 * g2[name] ** 2
    lr_t = lr * math.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    assert float((P1[name] - orc.P[name] - lr_t * m2 / (v2.sqrt() + 1e-8)).abs().max()) < 1e-12


def test_he_normal_truncated(oracle):
    rng = np.random.default_rng(0)
    w = oracle.he_normal(rng, (3, 3, 64, 64))
    std = math.sqrt(1.3 * 2.0 / (9 * 64))
    assert np.abs(w).max() <= 2 * std + 1e-12
    assert abs(w.std() / std - 0.8796) < 0.01     # std of a +-2 sigma truncated normal


def _count_conv_flops(oracle, orc, B=1):
    """2*MACs of every convolution the oracle actually executes in one training forward (dead branches are not run)."""
    tot = [0]
    real = oracle.conv2d_same

    def counting(x, W, b=None):
        tot[0] += 2 * x.shape[0] * x.shape[1] * x.shape[2] * W.shape[0] * W.shape[1] * W.shape[2] * W.shape[3]
        return real(x, W, b)

    oracle.conv2d_same = counting
    try:
        orc.init_params(seed=1)
        x, s = oracle.synthetic_batch(B, orc.H, orc.W, orc.nlabels, seed=1)
        eps = [torch.tensor(e) for e in oracle.synthetic_eps(orc.latent_shapes(B), seed=1)]
        orc.forward_train(torch.tensor(x), torch.tensor(s), eps)
    finally:
        oracle.conv2d_same = real
    return tot[0] / B


def test_topology_flops_and_params(oracle):
    """live conv FLOPs of the oracle's graph == SURVEY.md section 8d (25.03 / 22.04 GFLOP per image forward at 128x128,
    derived there from the reference's model_zoo), and its parameter counts (18.71 M incl. dead branches, 19.04 M)"""
    o = oracle.Oracle('phiseg')
    n_tr = sum(int(np.prod(s)) for n, s, k in o.spec.entries if k in ('W', 'b', 'gamma', 'beta'))
    assert abs(n_tr - 18.71e6) < 0.05e6          # incl. dead branches
    assert abs(_count_conv_flops(oracle, o) / 1e9 - 25.03) < 0.05
    o2 = oracle.Oracle('probunet', latent_levels=1, zdim0=6)
    n2 = sum(int(np.prod(s)) for n, s, k in o2.spec.entries if k in ('W', 'b', 'gamma', 'beta'))
    assert abs(n2 - 19.04e6) < 0.05e6
    assert abs(_count_conv_flops(oracle, o2) / 1e9 - 22.04) < 0.05


def test_engine_initialiser_and_synthetic_generator(oracle, pkg):
    """The ENGINE's own he_normal (engine.py, tfwrapper/utils.py:225-226: variance_scaling_initializer(factor=2, FAN_IN,
    normal) = N(0, 1.3*2/fan_in) truncated at 2 sigma) and Params.init defaults (gamma 1, beta 0, moving mean 0 /
    variance 1, zero biases), which the parity tests never exercise because they overwrite the weights from the oracle;
    and the product-side synthetic batch generator against the oracle's copy, bit for bit."""
    import importlib
    E = importlib.import_module('phiseg_code_b200.engine')
    gen = torch.Generator().manual_seed(5)
    shape = (3, 3, 64, 96)
    w = E.he_normal(gen, shape).numpy()
    std = math.sqrt(1.3 * 2.0 / (9 * 64))
    assert np.abs(w).max() <= 2 * std * (1 + 1e-6)
    assert abs(w.std() / std - 0.8796) < 0.01 and abs(w.mean()) < 0.01 * std
    cfg = E.NetConfig(image_size=(64, 64, 1), n0=4, mode='parity')
    P = E.Params(cfg, torch.device('cpu'))
    for name, shp, kind in P.spec:
        v = P.view(name)
        if kind in ('gamma', 'moving_variance'):
            assert bool((v == 1).all()), name
        elif kind in ('b', 'beta', 'moving_mean'):
            assert bool((v == 0).all()), name
        else:
            fan_in = shp[0] * shp[1] * shp[2]
            assert float(v.abs().max()) <= 2 * math.sqrt(1.3 * 2.0 / fan_in) * (1 + 1e-6), name
    D = importlib.import_module('phiseg_code_b200.data')
    for B, H, nl, seed in ((3, 64, 2, 3), (2, 128, 4, 1235)):
        xa, sa = D.synthetic_batch(B, H, H, nl, seed=seed)
        xb, sb = oracle.synthetic_batch(B, H, H, nl, seed=seed)
        assert np.array_equal(xa, xb) and np.array_equal(sa, sb) and xa.dtype == np.float32 and sa.dtype == np.uint8
    ea = D.synthetic_eps([(2, 4, 4, 2), (2, 2, 2, 2)], seed=9)
    eb = oracle.synthetic_eps([(2, 4, 4, 2), (2, 2, 2, 2)], seed=9)
    assert all(np.array_equal(a, b) for a, b in zip(ea, eb))


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
