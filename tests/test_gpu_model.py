"""-m gpu: the whole hot path (phiseg class surface -> engine -> C-ABI kernels) against the CPU oracle on identical
weights, inputs and injected eps.

Tolerances (BASELINE.json north_star): per-pixel logits within 1e-3, argmax masks bit-exact (asserted on pixels whose
logit margin exceeds the achieved tolerance; near-ties exist at random init).  Training-mode batch norm
at random init amplifies rounding differences ~1.2x per layer, so for BN the end-to-end training
assertions are on the losses and on relative gradient errors with a looser bound; GN is asserted tightly."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZE = 64   # smallest image the 7-level pyramid accepts


def _mods(pkg):
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
    return pm, ex


def _setup(pkg, oracle, exp_name, B, mode='parity', graph=False, size=SIZE, seed=7, fp64=True):
    """exp_name + '+gn' swaps the experiment's normalisation for group_norm2D (per-sample statistics: well conditioned,
    so parity can be asserted tightly; training-mode batch norm at random init is chaotic)"""
    pm, ex = _mods(pkg)
    gn = exp_name.endswith('+gn')
    exp = ex.load_experiment(ex.experiment_path(exp_name[:-3] if gn else exp_name))
    if gn:
        exp.layer_norm = ex.load_experiment(ex.experiment_path('phiseg_7_5_gn')).layer_norm
    exp.image_size = (size, size, 1)
    model = pm.phiseg(exp, mode=mode, use_cuda_graph=graph)
    cfg = model.cfg
    orc = oracle.Oracle(cfg.arch, image_size=(size, size, 1), nlabels=cfg.nlabels, zdim0=cfg.zdim0, n0=cfg.n0,
                        resolution_levels=cfg.R, latent_levels=cfg.L, norm=cfg.norm, KL_weight=cfg.KL_weight,
                        xent_weight=cfg.xent_weight, dtype=torch.float64 if fp64 else torch.float32)
    P = orc.init_params(seed=seed)
    model.set_weights({k: v.numpy() for k, v in P.items()})
    x, s = oracle.synthetic_batch(B, size, size, cfg.nlabels, seed=3)
    eps = oracle.synthetic_eps(orc.latent_shapes(B), seed=5)
    return model, orc, x, s, eps


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize('exp_name,mode,tol_loss,tol_grad', [
    ('phiseg_7_5_gn', 'parity', 1e-4, 5e-3), ('phiseg_7_5', 'parity', 1e-3, 5e-1), ('probunet', 'parity', 1e-3, 5e-1),
    ('phiseg_7_1', 'parity', 1e-3, 5e-1), ('detunet', 'parity', 1e-3, 5e-1), ('detunet+gn', 'parity', 1e-4, 5e-3),
    # the same contract on the tensor cores: three bf16 tcgen05 passes over a (hi, lo) operand split per convolution
    # (measured round 2: 4.1e-3 / 5.8e-1 / 5.7e-3 against 3.9e-3 / 1.8e-1 / - for the CUDA-core mode)
    ('phiseg_7_5_gn', 'parity_tc', 1e-4, 1e-2), ('phiseg_7_5', 'parity_tc', 1e-3, 1.0), ('probunet+gn', 'parity_tc', 1e-4, 1e-2)])
def test_training_step_parity(pkg, oracle, exp_name, mode, tol_loss, tol_grad):
    """The fp32-accurate modes ('parity': CUDA-core kernels; 'parity_tc': tcgen05 with split operands) against the fp64
    oracle.  Group norm is asserted tightly.  Training-mode batch norm at random init is chaotic (two correct
    fp32 implementations differ by >1e-3 end to end at depth 47, errors grow ~1.2x per layer), so for BN the gradient bound
    is loose and the tight checks are the per-kernel tests."""
    B = 3
    model, orc, x, s, eps = _setup(pkg, oracle, exp_name, B, mode=mode)
    loss = model.training_step(x, s, lr=1e-3, eps=eps)
    ref_loss, out, g = orc.train_step(torch.tensor(x), torch.tensor(s), [torch.tensor(e) for e in eps], 1e-3)
    # losses, level by level (phiseg_model.py:229-287)
    for k, v in out.loss_dict.items():
        got = model.loss_dict[k]
        assert abs(got - float(v)) <= tol_loss * max(1.0, abs(float(v))), (k, got, float(v))
    assert abs(loss - ref_loss) <= tol_loss * max(1.0, abs(ref_loss))
    # gradients: every trainable variable that receives one
    worst = (0.0, None)
    for name, gr in g.items():
        if gr is None:
            assert float(model.params.view(name, model.params.g).abs().max()) == 0.0, 'dead branch got a gradient: ' + name
            continue
        got = model.params.view(name, model.params.g).cpu().numpy().reshape(gr.shape)
        ref = gr.numpy()
        if np.abs(ref).max() < 1e-12:
            assert np.abs(got).max() < 1e-6, name
            continue
        r = _rel(got, ref)
        if r > worst[0]:
            worst = (r, name)
    print('worst relative gradient error %.3e at %s (%s, %s)' % (worst + (exp_name, mode)))
    assert worst[0] <= tol_grad, worst
    # moving statistics (normalisation.py:145-163; FusedBatchNorm Bessel-corrected variance)
    if model.cfg.norm == 'batch_norm':
        for name in model.params.state_table:
            if name.startswith('posterior/z0_pre_1') or name.startswith('likelihood/z0_post_1'):
                r = _rel(model.params.view(name).cpu().numpy(), orc.P[name].numpy())
                assert r < 1e-3, (name, r)


@pytest.mark.parametrize('exp_name', ['phiseg_7_5_gn', 'probunet_gn'])
def test_training_step_fast_mode(pkg, oracle, exp_name):
    """bf16 tensor-core mode (tcgen05 kernels, fp32 accumulation / statistics / losses) against the fp64 oracle under
    group norm: losses within 1e-2 relative; gradients judged statistically (bf16 activations carry 2^-9 relative rounding
    per layer through ~47 layers and B=3 images of 64x64 leave the deep 2x2 levels with 12 samples): every gradient
    tensor points the same way (cosine > 0.9; measured >= 0.94) and the whole flat gradient is within 20% in L2
    (measured 6-12%).  The tensor-core kernels themselves are checked exactly in test_gpu_conv_tc.py."""
    B = 3
    pm, ex = _mods(pkg)
    if exp_name == 'probunet_gn':
        import types
        base = ex.load_experiment(ex.experiment_path('probunet'))
        gn = ex.load_experiment(ex.experiment_path('phiseg_7_5_gn'))
        base.layer_norm = gn.layer_norm
        exp = base
    else:
        exp = ex.load_experiment(ex.experiment_path(exp_name))
    exp.image_size = (SIZE, SIZE, 1)
    model = pm.phiseg(exp, mode='fast', use_cuda_graph=False)
    cfg = model.cfg
    orc = oracle.Oracle(cfg.arch, image_size=(SIZE, SIZE, 1), nlabels=cfg.nlabels, zdim0=cfg.zdim0, n0=cfg.n0,
                        resolution_levels=cfg.R, latent_levels=cfg.L, norm=cfg.norm, dtype=torch.float64)
    P = orc.init_params(seed=7)
    model.set_weights({k: v.numpy() for k, v in P.items()})
    x, s = oracle.synthetic_batch(B, SIZE, SIZE, cfg.nlabels, seed=3)
    eps = oracle.synthetic_eps(orc.latent_shapes(B), seed=5)
    loss = model.training_step(x, s, lr=1e-3, eps=eps)
    ref_loss, out, g = orc.train_step(torch.tensor(x), torch.tensor(s), [torch.tensor(e) for e in eps], 1e-3)
    print('fast-mode loss %.6f oracle %.6f rel %.2e' % (loss, ref_loss, abs(loss - ref_loss) / abs(ref_loss)))
    assert abs(loss - ref_loss) <= 1e-2 * max(1.0, abs(ref_loss))
    # gradients: bf16 rounding noise is random, so judge direction and magnitude per tensor and over the whole buffer
    worst = (1.0, None)
    num = den = 0.0
    rows = []
    for name, gr in g.items():
        if gr is None or np.abs(gr.numpy()).max() < 1e-12:
            continue
        got = model.params.view(name, model.params.g).cpu().numpy().reshape(gr.shape).astype(np.float64).ravel()
        ref = gr.numpy().astype(np.float64).ravel()
        cos = float(got @ ref / max(np.linalg.norm(got) * np.linalg.norm(ref), 1e-300))
        rows.append((cos, name, _rel(got, ref)))
        num += float(((got - ref) ** 2).sum())
        den += float((ref ** 2).sum())
        if cos < worst[0]:
            worst = (cos, name)
    rows.sort()
    for r in rows[:5]:
        print('fast-mode gradient cosine %.4f  max-rel %.3f  %s' % (r[0], r[2], r[1]))
    print('fast-mode whole-gradient relative L2 error %.3e' % np.sqrt(num / den))
    assert np.sqrt(num / den) < 0.2
    assert worst[0] > 0.9, worst


def _fast_model(pkg, oracle, exp_name, size, B, graph, seed=7, fp64=False):
    """fast-mode model + fp32/fp64 oracle on the same weights, synthetic batch and eps."""
    return _setup(pkg, oracle, exp_name, B, mode='fast', graph=graph, size=size, seed=seed, fp64=fp64)


@pytest.mark.parametrize('exp_name,size,B,graph', [
    ('phiseg_7_5', 128, 8, False), ('phiseg_7_5', 128, 8, True), ('phiseg_7_5_gn', 64, 2, False),
    ('phiseg_7_5_gn', 64, 2, True), ('probunet', 64, 4, True)])
def test_fast_mode_reproducible(pkg, oracle, exp_name, size, B, graph):
    """Round-1 finding: the tcgen05 path gave 70766 / 70851 / 70962 for the same step.  Cause: the fused normalisation
    statistics were fp32 atomics, so their last bits depended on CTA arrival order and every bf16 rounding downstream
    amplified that.  They are fp64 atomics of fp32 partials now (exact in practice), so the SAME step on FRESH models -
    eager launches and CUDA-graph replays, all lanes on - must give the same loss (<= 1e-6 relative; the scalar loss
    reductions are still fp32 atomics) and the same flat gradient (<= 2e-6 of its largest entry: the split-K filter
    gradients are fp32 atomics).  lr = 0 keeps the weights fixed, so steps 1 (eager), 2 (capture + replay) and 3 (replay)
    of one model are comparable too.  Anything above these bounds is a race or an uninitialised read."""
    runs = []
    for rep in range(3):
        model, orc, x, s, eps = _fast_model(pkg, oracle, exp_name, size, B, graph)
        for it in range(3 if graph else 1):
            loss = model.training_step(x, s, lr=0.0, eps=eps)
            runs.append((loss, model.params.g.detach().clone()))
        if graph:
            assert model._program('train', B).graphs, 'the training step was not captured into a CUDA graph'
        del model
    l0, g0 = runs[0]
    gmax = float(g0.abs().max())
    assert np.isfinite(l0) and gmax > 0
    worst_l = max(abs(l - l0) / max(1.0, abs(l0)) for l, _ in runs)
    worst_g = max(float((g - g0).abs().max()) / gmax for _, g in runs)
    print('reproducibility %s %d^2 B=%d graph=%s: %d runs, loss %.6f, worst rel loss diff %.2e, worst grad diff / max|g| %.2e'
          % (exp_name, size, B, graph, len(runs), l0, worst_l, worst_g))
    assert worst_l <= 1e-6, [l for l, _ in runs]
    assert worst_g <= 2e-6, worst_g


def _grad_report(model, g):
    """(worst cosine, its name, whole-gradient relative L2 error) of the engine's flat gradient against the oracle's."""
    worst = (1.0, None)
    num = den = 0.0
    for name, gr in g.items():
        if gr is None or np.abs(gr.numpy()).max() < 1e-12:
            continue
        got = model.params.view(name, model.params.g).cpu().numpy().reshape(gr.shape).astype(np.float64).ravel()
        ref = gr.numpy().astype(np.float64).ravel()
        cos = float(got @ ref / max(np.linalg.norm(got) * np.linalg.norm(ref), 1e-300))
        num += float(((got - ref) ** 2).sum())
        den += float((ref ** 2).sum())
        if cos < worst[0]:
            worst = (cos, name)
    return worst[0], worst[1], float(np.sqrt(num / den))


@pytest.mark.parametrize('exp_name,size,B,mode,tol_loss,tol_l2', [
    # group norm (every image independent, well conditioned): tight bounds = the implementation check at full size
    ('phiseg_7_5_gn', 128, 4, 'fast', 1e-2, 0.2),
    ('phiseg_7_5_256+gn', 256, 2, 'fast', 1e-2, 0.25),      # configs[4]: 256x256, 4 classes
    ('probunet+gn', 128, 4, 'fast', 1e-2, 0.25),
    ('phiseg_7_5_256+gn', 256, 1, 'parity', 1e-4, 1e-2),
    # batch norm in training mode at random init amplifies ANY rounding difference ~1.2x per layer through ~60 layers
    # (two correct fp32 implementations already differ by > 1e-3): statistical bounds, measured values in
    # the test output (round 2: 2.2e-2 / 9.6e-2 relative loss error for the first two)
    ('phiseg_7_5', 128, 8, 'fast', 6e-2, None),              # the BENCH configuration (batch norm, 128x128), smaller batch
    ('phiseg_7_5_256', 256, 4, 'fast', 0.25, None),
    ('probunet', 128, 8, 'fast', 6e-2, None),
    ('phiseg_7_5', 128, 4, 'parity', 2e-3, None)])
def test_training_step_full_size(pkg, oracle, exp_name, size, B, mode, tol_loss, tol_l2):
    """Both compute modes against the fp32 CPU oracle at BASELINE.json's image sizes (128x128, and 256x256 with 4 classes):
    every loss term and the whole gradient.  Under group norm the bounds are tight (fast mode: bf16 activations); under
    batch norm only the losses are bounded and the flat gradient must point the same way (global cosine > 0.2)."""
    model, orc, x, s, eps = _setup(pkg, oracle, exp_name, B, mode=mode, graph=False, size=size, fp64=False)
    loss = model.training_step(x, s, lr=1e-3, eps=eps)
    ref_loss, out, g = orc.train_step(torch.tensor(x), torch.tensor(s), [torch.tensor(e) for e in eps], 1e-3)
    rel = abs(loss - ref_loss) / max(1.0, abs(ref_loss))
    cos, name, l2 = _grad_report(model, g)
    names = [n for n, gr in g.items() if gr is not None]
    flat_ref = np.concatenate([g[n].numpy().ravel() for n in names]).astype(np.float64)
    flat_got = np.concatenate([model.params.view(n, model.params.g).cpu().numpy().ravel() for n in names]).astype(np.float64)
    gcos = float(flat_got @ flat_ref / (np.linalg.norm(flat_got) * np.linalg.norm(flat_ref)))
    print('full-size %s %s %d^2 B=%d: loss %.4f oracle %.4f rel %.2e | gradient: global cosine %.4f, rel L2 %.3f, worst '
          'tensor cosine %.3f (%s)' % (mode, exp_name, size, B, loss, ref_loss, rel, gcos, l2, cos, name))
    for k, v in out.loss_dict.items():
        v = float(v)
        assert abs(model.loss_dict[k] - v) <= 5 * tol_loss * max(1.0, abs(v)) + tol_loss * abs(ref_loss), (k, model.loss_dict[k], v)
    assert rel <= tol_loss
    if tol_l2 is not None:
        assert l2 <= tol_l2
        assert cos > 0.8, (cos, name)
    else:
        # a statistical bound: one realisation of a chaotic map (any last-bit change of the forward arithmetic re-rolls it;
        # measured 0.28 ... 0.6 over the round-2 kernel revisions for the same inputs).  The implementation check at these
        # sizes is the group-norm rows above (relative L2 <= 0.25, every tensor cosine > 0.8).
        assert gcos > 0.2, gcos


@pytest.mark.parametrize('exp_name,size', [('phiseg_7_5_gn', 64), ('phiseg_7_5', 128), ('probunet', 128), ('detunet', 128)])
def test_sampling_fast_mode(pkg, oracle, exp_name, size):
    """What bf16 costs on the sampling path (SURVEY.md D6 asks for the measured tolerance next to the fast number): the
    summed logits of one prior sample against the fp64 oracle.  The north-star contract (1e-3 per logit, exact argmax) is
    met by mode='parity'; fast mode is asserted to stay within 3% of the logit range and to give the same mask on
    >= 99% of the pixels whose margin exceeds twice the measured error."""
    B = 2
    model, orc, x, s, eps = _setup(pkg, oracle, exp_name, B, mode='fast', size=size)
    ref = orc.forward_sample(torch.tensor(x), [torch.tensor(e) for e in eps], training=False)
    seg = model.predict_segmentation_sample(x, eps=eps)
    logits = model._program('sample', B).s_out.cpu().numpy()
    ref_logits = ref.s_out_eval.numpy()
    err = float(np.abs(logits - ref_logits).max())
    rng = float(ref_logits.max() - ref_logits.min())
    srt = np.sort(ref_logits, axis=-1)
    margin = srt[..., -1] - srt[..., -2]
    safe = margin > 2 * err
    agree_all = float((seg == ref_logits.argmax(-1)).mean())
    agree_safe = float((seg[safe] == ref_logits.argmax(-1)[safe]).mean()) if safe.any() else 1.0
    print('fast sampling %s %d^2: max|dlogit| %.3e (logit range %.3f), argmax agreement %.4f overall, %.4f on the %.0f%% '
          'pixels with margin > 2*err' % (exp_name, size, err, rng, agree_all, agree_safe, 100 * safe.mean()))
    assert err <= 0.03 * max(rng, 1.0)
    assert agree_safe >= 0.99 and agree_all >= 0.97


def test_adam_update_parity_gn(pkg, oracle):
    """weights after one and two optimizer steps (TF-form Adam, phiseg_model.py:136-141)"""
    model, orc, x, s, eps = _setup(pkg, oracle, 'phiseg_7_5_gn', 2)
    xt, st, et = torch.tensor(x), torch.tensor(s), [torch.tensor(e) for e in eps]
    for it in range(2):
        model.training_step(x, s, lr=1e-3, eps=eps)
        orc.train_step(xt, st, et, 1e-3)
    # Adam normalises the step to ~lr per weight, so compare the *update* against lr
    w0 = oracle.Oracle(model.cfg.arch, image_size=(SIZE, SIZE, 1), norm='group_norm').init_params(seed=7)
    bad = 0
    tot = 0
    for name in ('likelihood/post_c_0_2/W', 'posterior/z0_pre_1/W', 'prior/z3_input_1/W', 'likelihood/y_lvl0/W',
                 'likelihood/post_c_0_2/group_norm/gamma'):
        got = model.params.view(name).cpu().numpy().astype(np.float64)
        ref = orc.P[name].numpy()
        d_ref = ref - w0[name].numpy()
        d_got = got - w0[name].numpy()
        # elements whose gradient is far from zero move by ~2*lr; sign-stable there
        big = np.abs(d_ref) > 1.5e-3
        tot += big.sum()
        bad += (np.abs(d_got - d_ref)[big] > 2e-4).sum()
    assert tot > 0 and bad / tot < 0.01, (bad, tot)


@pytest.mark.parametrize('exp_name,mode', [('phiseg_7_5_gn', 'parity'), ('phiseg_7_5', 'parity'), ('probunet', 'parity'),
                                           ('phiseg_7_5_gn', 'parity_tc'), ('phiseg_7_5', 'parity_tc'),
                                           ('probunet', 'parity_tc')])
def test_sampling_parity(pkg, oracle, exp_name, mode):
    """prior(generation_mode=True) -> likelihood -> sum of levels, training=False (phiseg_model.py:61-109):
    logits within 1e-3 per pixel, argmax bit-exact outside near-ties - the north-star contract, met by the CUDA-core mode
    and by the tensor-core mode with split operands ('parity_tc')."""
    B = 2
    model, orc, x, s, eps = _setup(pkg, oracle, exp_name, B, mode=mode)
    ref = orc.forward_sample(torch.tensor(x), [torch.tensor(e) for e in eps], training=False)
    sm = model.predict_segmentation_sample(x, return_softmax=True, eps=eps)
    seg = model.predict_segmentation_sample(x, eps=eps)
    sp = model._program('sample', B)
    logits = sp.s_out.cpu().numpy()
    ref_logits = ref.s_out_eval.numpy()
    err = np.abs(logits - ref_logits).max()
    print('max |logit diff| = %.3e (%s, %s)' % (err, exp_name, mode))
    assert err < 1e-3
    assert np.abs(sm - ref.s_out_eval_sm.numpy()).max() < 1e-3
    srt = np.sort(ref_logits, axis=-1)
    margin = srt[..., -1] - srt[..., -2]
    safe = margin > 2e-3
    assert safe.mean() > 0.5
    assert np.array_equal(seg[safe], ref_logits.argmax(-1)[safe])
    # per-level outputs and latents
    lv = model.predict_segmentation_sample_levels(x, eps=eps)
    for a, b in zip(lv, ref.s_out_eval_list):
        assert np.abs(a - b.numpy()).max() < 1e-3
    z, mu, sg = model.generate_prior_samples(x, return_params=True, eps=eps)
    for a, b in zip(z, ref.prior_z):
        assert np.abs(a - b.numpy().reshape(a.shape)).max() < 1e-3
    # decoding the same latents through generate_samples_from_z reproduces the sample
    s2 = model.generate_samples_from_z(z, x)
    assert np.abs(s2 - logits).max() < 1e-4


@pytest.mark.parametrize('exp_name,mode,tol', [('phiseg_7_5', 'parity', 1e-6), ('phiseg_7_5_gn', 'parity', 1e-5),
                                               ('probunet', 'parity', 1e-6), ('phiseg_7_5', 'fast', 1e-6),
                                               ('phiseg_7_5_gn', 'fast', 5e-2), ('probunet', 'fast', 1e-6)])
def test_batched_sampling_matches_single_samples(pkg, oracle, exp_name, mode, tol):
    """predict() evaluates rep samples of every image in one pass (rows s*B + b) and runs the x-only part of the graph
    once per image (phiseg_model.py:337-353 re-runs the whole graph per sample).  Row s*B + b of the batched pass must
    equal what the one-sample program gives for image b with the same noise; replaying only the noise-dependent part
    with new noise must do the same; the running softmax sum must be the sum over an image's samples."""
    B, rep = 2, 3
    model, orc, x, s, _ = _setup(pkg, oracle, exp_name, B, mode=mode)
    big = model._program('sample', B, rep)
    one = model._program('sample', B)
    assert 0 < big.n_enc < len(big.prog.steps)
    enc, rest = big.prog.steps[:big.n_enc], big.prog.steps[:big.n_fills] + big.prog.steps[big.n_enc:]
    model._stage_x(big, x)
    big.sm_accum.zero_()
    model._launch(big, enc, 'enc')
    want_acc = 0
    for draw in range(2):
        eps_big = oracle.synthetic_eps(model.cfg.latent_shapes(B * rep), seed=11 + draw)
        model._draw_eps(big, eps_big)
        model._launch(big, rest, 'rest')
        got = big.s_out.cpu().numpy().reshape(rep, B, *big.s_out.shape[1:])
        got_sm = big.s_out_sm.cpu().numpy().reshape(got.shape)
        for k in range(rep):
            eps_k = [e[k * B:(k + 1) * B] for e in eps_big]
            model.predict_segmentation_sample(x, eps=eps_k)
            ref = one.s_out.cpu().numpy()
            err = np.abs(got[k] - ref).max() / max(1.0, np.abs(ref).max())
            assert err <= tol, (draw, k, err)
        want_acc = want_acc + got_sm.sum(axis=0)
    assert np.abs(big.sm_accum.cpu().numpy() - want_acc).max() < 1e-5
    # the public calls built on it: 5 samples with 4 rows per pass = passes of 2 + 2 + 1 samples per image
    model.sample_rows = 4
    assert model._sample_plan(B, 5) == [(2, 2), (1, 1)]
    seg, sm = model.predict(x, num_samples=5, return_softmax=True)
    assert seg.shape == (B, SIZE, SIZE) and np.allclose(sm.sum(-1), 1.0, atol=1e-5) and np.array_equal(seg, sm.argmax(-1))
    smp = model.generate_samples(x, 5)
    assert smp.shape == (5, B, SIZE, SIZE, model.cfg.nlabels) and np.all(np.isfinite(smp))
    assert np.abs(smp[0] - smp[1]).max() > 0          # different noise per sample


def test_posterior_samples_and_eval_losses(pkg, oracle):
    B = 2
    model, orc, x, s, eps = _setup(pkg, oracle, 'phiseg_7_5', B)
    z, mu, sg = model.generate_posterior_samples(x, s, return_params=True, eps=eps)
    xt, st, et = torch.tensor(x), torch.tensor(s), [torch.tensor(e) for e in eps]
    zr, mr, sr = orc.posterior(xt.double(), orc.one_hot(st), et, training=False)
    for a, b in zip(z + mu + sg, zr + mr + sr):
        assert np.abs(a - b.numpy()).max() < 1e-3
    ld = model.evaluate_losses(x, s, eps=eps)
    ref = orc.forward_train(xt, st, et, training=False).loss_dict
    for k, v in ref.items():
        assert abs(ld[k] - float(v)) <= 1e-3 * max(1.0, abs(float(v))), (k, ld[k], float(v))


def test_predict_api_and_cuda_graph_replay(pkg, oracle):
    """predict / generate_samples shapes and dtypes; a captured CUDA graph replays to the same result as eager launches"""
    B = 2
    model_e, orc, x, s, eps = _setup(pkg, oracle, 'phiseg_7_5_gn', B, graph=False)
    model_g, _, _, _, _ = _setup(pkg, oracle, 'phiseg_7_5_gn', B, graph=True)
    for it in range(4):           # eager, capture, replay, replay
        le = model_e.training_step(x, s, lr=1e-3, eps=eps)
        lg = model_g.training_step(x, s, lr=1e-3, eps=eps)
        # same weights at it=0; afterwards atomics reorder fp32 sums and Adam's normalised step amplifies the difference
        assert abs(le - lg) <= (1e-5, 2e-3, 2e-2, 2e-2)[it] * max(1.0, abs(le)), (it, le, lg)
    assert model_g._program('train', B).graphs, 'the training step was not captured into a CUDA graph'
    seg, sm = model_g.predict(x, num_samples=3, return_softmax=True)
    assert seg.shape == (B, SIZE, SIZE) and seg.dtype == np.int64
    assert sm.shape == (B, SIZE, SIZE, 2) and np.allclose(sm.sum(-1), 1.0, atol=1e-5)
    assert np.array_equal(seg, sm.argmax(-1))
    smp = model_g.generate_samples(x, 2)
    assert smp.shape == (2, B, SIZE, SIZE, 2)
    lv = model_g.generate_samples_from_prior(x, output_all_levels=True)
    assert len(lv) == 5 and lv[4].shape == (B, SIZE, SIZE, 2)
    # the rest of the reference's class surface (phiseg_model.py:160,378-430,498-502)
    model_g.checks()
    lv2 = model_g.generate_all_output_levels(x)
    assert len(lv2) == 5 and all(a.shape == (B, SIZE, SIZE, 2) for a in lv2)
    var = model_g.predict_segmentation_sample_variance_sm_cov(x[:1], 4)
    assert var.shape == (SIZE, SIZE) and np.all(np.isfinite(var))
    det = model_g.predict_segmentation_sample_variance_sm_cov_bf(x[:1], 4)
    assert det.shape == (SIZE, SIZE) and np.all(np.isfinite(det)) and np.all(det > -1e-6)
    with pytest.raises(ValueError):
        model_g.training_step(x[:, :32], s, lr=1e-3)
    with pytest.raises(ValueError):
        model_g.training_step(x, s + 7, lr=1e-3)
    assert model_g.gpu_launches > 0


def test_loss_decreases_full_size(pkg, oracle):
    """size-independent property at BASELINE.json's full resolution (128x128): a few Adam steps on one fixed batch
    reduce the ELBO; all losses stay finite."""
    model, orc, x, s, eps = _setup(pkg, oracle, 'phiseg_7_5', 4, graph=True, size=128)
    losses = [model.training_step(x, s, lr=1e-3, eps=eps) for _ in range(8)]
    assert all(np.isfinite(l) for l in losses)
    assert losses[-1] < losses[0], losses


def test_checkpoint_roundtrip(pkg, oracle, tmp_path):
    model, orc, x, s, eps = _setup(pkg, oracle, 'phiseg_7_5', 2)
    model.training_step(x, s, lr=1e-3, eps=eps)
    p = model.save_weights(str(tmp_path), 'model.ckpt-1')
    ref = model.predict_segmentation_sample(x, return_softmax=True, eps=eps)
    model.training_step(x, s, lr=1e-3, eps=eps)
    model.load_weights(str(tmp_path), 'latest')
    assert model.params.step == 1
    again = model.predict_segmentation_sample(x, return_softmax=True, eps=eps)
    assert np.array_equal(ref, again)
    with pytest.raises(ValueError):
        model.load_weights(str(tmp_path), 'nonsense')
    # a truncated newest checkpoint (crash mid-write of an older, non-atomic writer) falls back to the previous one
    (tmp_path / 'model.ckpt-7.npz').write_bytes(b'PK\x03\x04 truncated')
    assert model.load_weights(str(tmp_path), 'latest').endswith('model.ckpt-1.npz')
    assert model.params.step == 1


def test_weight_decay_end_to_end(pkg, oracle):
    """add_weight_decay (phiseg_model.py:124-128,290-300): the loss term, its gradient wd*W on every filter (dead branches
    included: they are in the 'weight_variables' collection) and the validation total_loss, against the fp64 oracle."""
    pm, ex = _mods(pkg)
    exp = ex.load_experiment(ex.experiment_path('phiseg_7_5_gn'))
    exp.image_size = (SIZE, SIZE, 1)
    exp.weight_decay_weight = 1e-4
    B = 2
    model = pm.phiseg(exp, mode='parity', use_cuda_graph=False)
    cfg = model.cfg
    orc = oracle.Oracle(cfg.arch, image_size=(SIZE, SIZE, 1), norm=cfg.norm, weight_decay=1e-4, dtype=torch.float64)
    P = orc.init_params(seed=7)
    model.set_weights({k: v.numpy() for k, v in P.items()})
    x, s = oracle.synthetic_batch(B, SIZE, SIZE, cfg.nlabels, seed=3)
    eps = oracle.synthetic_eps(orc.latent_shapes(B), seed=5)
    xt, st, et = torch.tensor(x), torch.tensor(s), [torch.tensor(e) for e in eps]
    ld = model.evaluate_losses(x, s, eps=eps)
    ref = orc.forward_train(xt, st, et, training=False).loss_dict
    assert abs(ld['weight_decay'] - float(ref['weight_decay'])) <= 1e-5 * float(ref['weight_decay'])
    assert abs(ld['total_loss'] - float(ref['total_loss'])) <= 1e-4 * abs(float(ref['total_loss']))
    loss = model.training_step(x, s, lr=1e-3, eps=eps)
    ref_loss, out, g = orc.train_step(xt, st, et, 1e-3)
    assert abs(loss - ref_loss) <= 1e-4 * abs(ref_loss)
    assert abs(model.loss_dict['weight_decay'] - float(out.loss_dict['weight_decay'])) <= 1e-5 * float(out.loss_dict['weight_decay'])
    for name in ('likelihood/post_c_0_2/W', 'posterior/z0_pre_1/W', 'posterior/z4_ups_to_3_c_1/W'):
        got = model.params.view(name, model.params.g).cpu().numpy()
        want = g[name].numpy() if g[name] is not None else 1e-4 * P[name].numpy()
        assert _rel(got, want) < 5e-3, name


def test_session_attribute_api(pkg, oracle):
    """The de-facto attribute API of the reference's evaluation scripts (phiseg_test_quantitative.py:49-54,
    phiseg_makegif_samples.py:96-100): model.sess.run(model.s_out_eval_sm, feed_dict={model.training_pl: False,
    model.x_inp: x}).  All fetches of one call belong to one prior draw and equal what the method API returns for the same
    noise."""
    model, orc, x, s, eps = _setup(pkg, oracle, 'phiseg_7_5', 2, mode='fast', graph=True)
    fd = {model.training_pl: False, model.x_inp: x}
    model._gen.manual_seed(5)
    sm, lg, lv, z = model.sess.run([model.s_out_eval_sm, model.s_out_eval, model.s_out_eval_list, model.prior_z_list_gen], feed_dict=fd)
    model._gen.manual_seed(5)
    ref = model.predict_segmentation_sample(x, return_softmax=True)
    assert sm.shape == (2, SIZE, SIZE, 2) and np.array_equal(sm, ref)
    assert len(lv) == 5 and all(a.shape == (2, SIZE, SIZE, 2) for a in lv)
    assert np.allclose(np.sum(lv, axis=0), lg, atol=1e-4)                    # _aggregate_output_list, phiseg_model.py:210-226
    assert len(z) == 5 and z[0].shape[0] == 2
    # the quantitative script's pattern: one image tiled n_samples times, one fetch, not a list
    xt = np.tile(x[:1], [6, 1, 1, 1])
    arr = model.sess.run(model.s_out_eval_sm, feed_dict={model.training_pl: False, model.x_inp: xt})
    assert arr.shape == (6, SIZE, SIZE, 2) and np.allclose(arr.sum(-1), 1.0, atol=1e-5)
    assert not np.array_equal(arr[0], arr[1])                                # six different prior draws of the same image
    tot = model.sess.run('loss_tot', feed_dict={model.training_pl: False, model.x_inp: x, model.s_inp: s})
    assert np.isfinite(tot)
    with pytest.raises(NotImplementedError):
        model.sess.run(model.s_out_eval_sm, feed_dict={model.training_pl: True, model.x_inp: x})
    with pytest.raises(KeyError):
        model.sess.run(model.s_out_eval_sm, feed_dict={model.training_pl: False})
