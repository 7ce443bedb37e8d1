"""Validation metrics (phiseg_model.py:558-640, utils.py:103-118,270-362).

CPU: the oracle's numpy restatement against hand-computed cases, and the host-side finishing arithmetic
(phiseg-code_b200/metrics.py, fed with counts / moments computed by numpy) against the oracle on random masks.
GPU (-m gpu): the kernels of csrc/metrics.cu through the C-ABI, and phiseg.validation_metrics end to end."""
import importlib

import numpy as np
import pytest
import torch


def _counts(a, b, nl):
    inter = np.array([[[np.sum((x == l) & (y == l)) for l in range(nl)] for y in b] for x in a])
    ca = np.array([[np.sum(x == l) for l in range(nl)] for x in a])
    cb = np.array([[np.sum(y == l) for l in range(nl)] for y in b])
    return inter, ca, cb


def test_oracle_metrics_known_answers(oracle):
    # two samples, one annotation, one foreground label: d(a0,g)=1-2/3, d(a1,g)=1, d(a0,a1)=1-1/3 => GED = 4/3 - 1/3 = 1
    a = np.array([[[0, 1], [1, 1]], [[0, 0], [0, 1]]])
    g = np.array([[[0, 1], [1, 0]]])
    assert abs(oracle.generalised_energy_distance(a, g, range(1, 2)) - 1.0) < 1e-12
    # identical sets: every cross distance equals the within distances => 0; both-empty label counts as IoU 1
    assert abs(oracle.generalised_energy_distance(g, g, range(1, 3))) < 1e-12
    # exactly one mask empty -> IoU 0 -> distance 1: GED = 2*1 - 0 - 0
    z = np.zeros((1, 2, 2), int)
    assert abs(oracle.generalised_energy_distance(z, g, range(1, 2)) - 2.0) < 1e-12
    # ncc: correlation of a map with an affine image of itself is +-1
    m = np.arange(12.0).reshape(3, 4)
    assert abs(oracle.ncc(m, 3 * m + 2) - 1.0) < 1e-12 and abs(oracle.ncc(m, -m) + 1.0) < 1e-12
    # dice: |A|=3, |B|=2, |A&B|=2 -> 4/5; label absent in both -> 1; absent in one -> 0
    d = oracle.per_label_dice(a[0], g[0], 3)
    assert abs(d[1] - 0.8) < 1e-12 and d[2] == 1.0
    assert oracle.per_label_dice(z[0], g[0], 2)[1] == 0.0


def test_host_finishing_matches_oracle(oracle, pkg):
    M = importlib.import_module('phiseg_code_b200.metrics')
    rng = np.random.default_rng(3)
    for nl, N, A in ((2, 5, 3), (4, 3, 4)):
        s = rng.integers(0, nl, size=(N, 12, 10))
        y = rng.integers(0, nl, size=(A, 12, 10))
        s[0] = 0                       # an empty sample: exercises the both-empty / one-empty conventions
        i_sy, c_s, c_y = _counts(s, y, nl)
        i_ss, _, _ = _counts(s, s, nl)
        i_yy, _, _ = _counts(y, y, nl)
        got = M.ged_from_counts(i_sy, i_ss, i_yy, c_s, c_y, range(1, nl))
        assert abs(got - oracle.generalised_energy_distance(s, y, range(1, nl))) < 1e-12
        i_d, c_p, c_g = _counts(s[1:2], y[0:1], nl)
        assert np.allclose(M.dice_from_counts(i_d[0, 0], c_p[0], c_g[0]), oracle.per_label_dice(s[1], y[0], nl))
        sm = rng.random((N, 12, 10, nl))
        sm /= sm.sum(-1, keepdims=True)
        oh = np.eye(nl)[y]
        logs = np.log(sm + 1e-8)
        e_ss = np.mean(-np.sum(sm.mean(0)[None] * logs, -1), 0).ravel()
        sums = []
        for j in range(A):
            e_sy = np.mean(-np.sum(oh[j][None] * logs, -1), 0).ravel()
            sums.append([e_ss.sum(), (e_ss ** 2).sum(), e_sy.sum(), (e_sy ** 2).sum(), (e_ss * e_sy).sum()])
        assert abs(M.ncc_from_sums(np.array(sums), e_ss.size) - oracle.variance_ncc_dist(sm, oh)) < 1e-9


@pytest.mark.gpu
def test_metric_kernels(lib, oracle, pkg):
    M = importlib.import_module('phiseg_code_b200.metrics')
    h = lib.load()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(5)
    for nl, N, A, H in ((2, 6, 4, 32), (4, 3, 2, 24)):
        s = torch.randint(0, nl, (N, H, H), generator=g)
        y = torch.randint(0, nl, (A, H, H), generator=g).to(torch.uint8)
        s[0] = 0
        sd, yd = s.cuda(), y.cuda()            # int64 samples (argmax output), uint8 annotations
        P = H * H
        i_sy = torch.empty(N, A, nl, dtype=torch.int32, device='cuda')
        i_ss = torch.empty(N, N, nl, dtype=torch.int32, device='cuda')
        i_yy = torch.empty(A, A, nl, dtype=torch.int32, device='cuda')
        c_s = torch.empty(N, nl, dtype=torch.int32, device='cuda')
        c_y = torch.empty(A, nl, dtype=torch.int32, device='cuda')
        lib.check(h.phs_pairwise_label_stats(sd.data_ptr(), 8, N, yd.data_ptr(), 1, A, P, nl, i_sy.data_ptr(), c_s.data_ptr(), c_y.data_ptr(), st))
        lib.check(h.phs_pairwise_label_stats(sd.data_ptr(), 8, N, sd.data_ptr(), 8, N, P, nl, i_ss.data_ptr(), None, None, st))
        lib.check(h.phs_pairwise_label_stats(yd.data_ptr(), 1, A, yd.data_ptr(), 1, A, P, nl, i_yy.data_ptr(), None, None, st))
        ref_sy, ref_cs, ref_cy = _counts(s.numpy(), y.numpy(), nl)
        assert np.array_equal(i_sy.cpu().numpy(), ref_sy) and np.array_equal(c_s.cpu().numpy(), ref_cs)
        assert np.array_equal(c_y.cpu().numpy(), ref_cy)
        ged = M.ged_from_counts(i_sy.cpu().numpy(), i_ss.cpu().numpy(), i_yy.cpu().numpy(), c_s.cpu().numpy(), c_y.cpu().numpy(), range(1, nl))
        assert abs(ged - oracle.generalised_energy_distance(s.numpy(), y.numpy(), range(1, nl))) < 1e-12
        sm = torch.rand(N, H, H, nl, generator=g)
        sm = sm / sm.sum(-1, keepdim=True)
        e_ss = torch.empty(P, device='cuda')
        e_sy = torch.empty(A, P, device='cuda')
        sums = torch.empty(A, 5, dtype=torch.float64, device='cuda')
        lib.check(h.phs_ncc_maps(sm.cuda().contiguous().data_ptr(), yd.data_ptr(), N, A, P, nl, e_ss.data_ptr(), e_sy.data_ptr(), sums.data_ptr(), st))
        torch.cuda.synchronize()
        ref = oracle.variance_ncc_dist(sm.numpy(), np.eye(nl)[y.numpy()])
        assert abs(M.ncc_from_sums(sums.cpu().numpy(), P) - ref) < 2e-5


def test_oracle_uncertainty_maps_known_answers(oracle):
    """hand cases for the numpy restatement of phiseg_model.py:378-475"""
    # two samples, one pixel, two classes: logits (0, 0) and (ln 3, 0) -> softmax (.5, .5) and (.75, .25)
    z = np.array([[[[0.0, 0.0]]], [[[np.log(3.0), 0.0]]]])
    arg, sd, err = oracle.mean_variance_and_error_maps(z, np.array([[1]]))
    assert arg[0, 0] == 0 and abs(sd[0, 0] - 0.125) < 1e-12            # np.std of (.5, .75) = .125 for both classes
    assert abs(err[0, 0] - 0.5 * (np.log(2.0) + np.log(4.0))) < 1e-12   # -log .5 and -log .25
    # covariance of (p, 1-p) is singular: determinant 0; unbiased variance of (.5, .75) is 1/32
    assert abs(oracle.sample_variance_sm_cov_bf(z)[0, 0]) < 1e-12
    # clipped logits of class 0: (1e-5, 1 - 1e-5) -> population variance ((1 - 2e-5) / 2)^2
    assert abs(oracle.sample_variance_sm_cov(z)[0, 0] - ((1 - 2e-5) / 2) ** 2) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('nl,B', [(2, 1), (4, 2)])
def test_sample_moment_kernels(lib, oracle, nl, B):
    """phs_sample_moments / phs_sample_maps against the oracle's stacked-sample numpy (phiseg_model.py:378-475); the
    samples arrive in two passes of different size like a sampling plan with a remainder."""
    h = lib.load()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(11 + nl)
    H, W, S1, S2 = 12, 10, 5, 2
    P = H * W
    z = torch.randn(S1 + S2, B, H, W, nl, generator=g) * 2.0
    gt = torch.randint(0, nl, (B, H, W), generator=g).to(torch.uint8)
    na = nl + nl * (nl + 1) // 2 + 1
    for kind in (0, 1):
        acc = torch.zeros(B * P * na, dtype=torch.float64, device='cuda')
        for part in (z[:S1], z[S1:]):
            pd = part.contiguous().cuda()
            lib.check(h.phs_sample_moments(pd.data_ptr(), gt.cuda().data_ptr(), part.shape[0], B, P, nl, kind, 1e-5, 1 - 1e-5,
                                           acc.data_ptr(), st))
        arg = torch.empty(B, H, W, dtype=torch.int64, device='cuda')
        sd, vs, det, err = (torch.empty(B, H, W, device='cuda') for _ in range(4))
        lib.check(h.phs_sample_maps(acc.data_ptr(), B * P, nl, S1 + S2, 1, arg.data_ptr(), sd.data_ptr(), vs.data_ptr(),
                                    det.data_ptr(), err.data_ptr(), st))
        torch.cuda.synchronize()
        for b in range(B):
            zb, gb = z[:, b].numpy(), gt[b].numpy()
            if kind == 0:
                r_arg, r_sd, r_err = oracle.mean_variance_and_error_maps(zb, gb)
                assert np.array_equal(arg[b].cpu().numpy(), r_arg)
                assert np.abs(sd[b].cpu().numpy() - r_sd).max() < 2e-6
                assert np.abs(err[b].cpu().numpy() - r_err).max() < 1e-5
                r_det = oracle.sample_variance_sm_cov_bf(zb)
                assert np.abs(det[b].cpu().numpy() - r_det).max() < 1e-7 + 1e-4 * np.abs(r_det).max()
            else:
                assert np.abs(vs[b].cpu().numpy() - oracle.sample_variance_sm_cov(zb)).max() < 2e-6
                # softmax rows sum to one (singular covariance, det ~ 0 above); the clipped logits give a regular matrix
                zc = np.clip(zb.astype(np.float64), 1e-5, 1 - 1e-5).transpose((1, 2, 3, 0))
                r_det = np.array([[np.linalg.det(np.cov(zc[i, j])) for j in range(W)] for i in range(H)])
                assert np.abs(det[b].cpu().numpy() - r_det).max() < 1e-9 + 1e-4 * np.abs(r_det).max()


@pytest.mark.gpu
def test_uncertainty_maps_end_to_end(pkg, oracle):
    """The four map methods of the class (phiseg_model.py:378-475) on device moments == the oracle's numpy on the SAME
    samples (generate_samples with the same generator seed draws the same noise through the same sampling plan)."""
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
    D = importlib.import_module('phiseg_code_b200.data')
    exp = ex.load_experiment(ex.experiment_path('phiseg_7_5'))
    exp.image_size = (64, 64, 1)
    model = pm.phiseg(exp, mode='fast', use_cuda_graph=False)
    model.sample_rows = 4                      # 6 samples = one pass of 4 + a remainder pass of 2
    x, s = D.synthetic_batch(1, 64, 64, 2, seed=3)
    S = 6

    def seeded(fn):
        model._gen.manual_seed(77)
        return fn()
    smp = seeded(lambda: model.generate_samples(x, S))[:, 0]                     # [S,H,W,nl] summed-level logits
    arg, sd, err = seeded(lambda: model.predict_mean_variance_and_error_maps(s, x, S))
    r_arg, r_sd, r_err = oracle.mean_variance_and_error_maps(smp, s[0])
    assert arg.shape == (64, 64) and np.mean(arg == r_arg) > 0.999              # ties of the mean softmax may flip
    assert np.abs(sd - r_sd).max() < 1e-5 and np.abs(err - r_err).max() < 1e-4 * max(1.0, r_err.max())
    e2 = seeded(lambda: model.get_crossentropy_error_map(s, x, S))
    assert e2.shape == (1, 64, 64) and np.array_equal(e2[0], err)
    var = seeded(lambda: model.predict_segmentation_sample_variance_sm_cov(x, S))
    assert np.abs(var - oracle.sample_variance_sm_cov(smp)).max() < 1e-5
    det = seeded(lambda: model.predict_segmentation_sample_variance_sm_cov_bf(x, S))
    r_det = oracle.sample_variance_sm_cov_bf(smp)
    assert np.abs(det - r_det).max() < 1e-7 + 1e-3 * np.abs(r_det).max()


@pytest.mark.gpu
def test_validation_metrics_end_to_end(pkg, oracle, tmp_path):
    """phiseg.validation_metrics against the oracle's metrics evaluated on the SAME samples (copied back for the check), and
    the full validation loop with its four best-model savers."""
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    ex = importlib.import_module('phiseg_code_b200.phiseg.experiments')
    D = importlib.import_module('phiseg_code_b200.data')
    exp = ex.load_experiment(ex.experiment_path('phiseg_7_5'))
    exp.image_size = (64, 64, 1)
    exp.validation_samples, exp.num_validation_images = 6, 3
    exp.log_root = str(tmp_path)
    model = pm.phiseg(exp, mode='fast', use_cuda_graph=False)
    data = D.SyntheticLIDC(num_train=8, num_val=3, size=64, annotators=4, seed=2)
    x, labs = data.validation.images[1], data.validation.labels[1]
    m = model.validation_metrics(x[None, ..., None], labs, 6, annotator=2)
    sp = model._program('sample', 1, 6)
    sm = sp.s_out_sm.cpu().numpy()
    arg = sp.argmax.cpu().numpy()
    gts = np.moveaxis(labs, -1, 0)
    assert abs(m['ged'] - oracle.generalised_energy_distance(arg, gts, range(1, 2))) < 1e-9
    ncc_ref = oracle.variance_ncc_dist(sm, np.eye(2)[gts])
    assert (np.isnan(m['ncc']) and np.isnan(ncc_ref)) or abs(m['ncc'] - ncc_ref) < 1e-4
    assert np.allclose(m['dice'], oracle.per_label_dice(sm.mean(0).argmax(-1), gts[2], 2))
    assert np.isfinite(m['elbo'])
    model._setup_log_dir_and_continue_mode()
    model.best_loss = np.inf
    out = model._do_validation(data, 7)
    assert out['images'] == 3 and 0.0 <= out['dice'] <= 1.0 and np.isfinite(out['ged'])
    import os
    names = sorted(os.listdir(model.log_dir))
    for prefix in ('model.ckpt-7', 'model_best_dice.ckpt-7', 'model_best_loss.ckpt-7', 'model_best_ged.ckpt-7'):
        assert prefix + '.npz' in names, names
