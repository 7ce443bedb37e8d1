"""-m gpu: every kernel of libphiseg_sm100.so against the CPU oracle (oracle/phiseg_oracle.py) through the C-ABI.

Tolerances: fp32 kernels vs the fp32/fp64 oracle 1e-4 relative (sums over up to 9*192 products); index / argmax
outputs bit-exact."""
import ctypes
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_util import Caller, cu

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-4, atol=1e-5, what=''):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    err = (a - b).abs().max().item()
    scale = b.abs().max().item()
    assert err <= atol + rtol * scale, '%s: max|diff| %.3e vs scale %.3e' % (what, err, scale)


@pytest.fixture()
def call(lib):
    return Caller(lib)


# ---------------------------------------------------------------------------------------------------------
# convolution (tfwrapper/layers.py:123,132) -- CUDA-core path
# ---------------------------------------------------------------------------------------------------------
CONV_CASES = [  # N, H, W, Cin, Cout, k
    (2, 16, 16, 3, 32, 3), (2, 8, 8, 32, 32, 3), (1, 4, 4, 192, 192, 3), (3, 2, 2, 192, 2, 3), (2, 16, 16, 64, 2, 1),
    (2, 32, 32, 1, 32, 3), (1, 8, 8, 70, 32, 1), (2, 16, 8, 5, 17, 3),
    # the z inputs (zdim_0 = 2 channels, 3x3): the unrolled small_cin specialisation, incl. 24 output vectors and tiny images
    (2, 16, 16, 2, 64, 3), (3, 8, 8, 2, 192, 3), (2, 2, 2, 2, 32, 3), (1, 32, 16, 2, 128, 3)]


@pytest.mark.parametrize('N,H,W,Cin,Cout,k', CONV_CASES)
def test_conv_fwd_dgrad_wgrad_simt(call, lib, oracle, N, H, W, Cin, Cout, k):
    g = torch.Generator().manual_seed(N * 1000 + Cin * 10 + Cout)
    x = torch.randn(N, H, W, Cin, generator=g, dtype=torch.float64)
    w = torch.randn(k, k, Cin, Cout, generator=g, dtype=torch.float64) * 0.1
    b = torch.randn(Cout, generator=g, dtype=torch.float64)
    x.requires_grad_(True); w.requires_grad_(True); b.requires_grad_(True)
    y = oracle.conv2d_same(x, w, b)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(gy)
    xd, wd, bd = cu(x, torch.float32), cu(w, torch.float32), cu(b, torch.float32)
    yd = torch.empty(N, H, W, Cout, device='cuda')
    call('phs_conv2d', call.T(xd), wd, bd, call.T(yd), k, 0, 0, lib.IMPL_SIMT)
    close(yd, y, what='conv fwd')
    # accumulate=1 adds on top
    call('phs_conv2d', call.T(xd), wd, None, call.T(yd), k, 0, 1, lib.IMPL_SIMT)
    close(yd, 2 * y - b, what='conv fwd accumulate')
    gyd = cu(gy, torch.float32)
    gxd = torch.empty_like(xd)
    call('phs_conv2d', call.T(gyd), wd, None, call.T(gxd), k, 1, 0, lib.IMPL_SIMT)
    close(gxd, x.grad, what='conv dgrad')
    gwd = torch.full_like(wd, 7.0)
    gbd = torch.full_like(bd, 7.0)
    call('phs_conv2d_wgrad', call.T(xd), call.T(gyd), gwd, gbd, k, 0, lib.IMPL_SIMT)
    close(gwd, w.grad, what='conv wgrad')
    close(gbd, b.grad, what='conv bias grad')
    call('phs_conv2d_wgrad', call.T(xd), call.T(gyd), gwd, gbd, k, 1, lib.IMPL_SIMT)
    close(gwd, 2 * w.grad, what='conv wgrad accumulate')


def test_conv_channel_slices(call, lib, oracle):
    """zero-copy tf.concat: inputs and outputs addressed as channel slices of wider buffers"""
    g = torch.Generator().manual_seed(5)
    xb = torch.randn(2, 8, 8, 48, generator=g)
    w = torch.randn(3, 3, 32, 16, generator=g) * 0.1
    yb = torch.zeros(2, 8, 8, 40)
    xd, wd, yd = cu(xb), cu(w), cu(yb)
    call('phs_conv2d', call.T(xd, 16, 32), wd, None, call.T(yd, 8, 16), 3, 0, 0, lib.IMPL_SIMT)
    ref = oracle.conv2d_same(xb[..., 16:48].double(), w.double())
    close(yd[..., 8:24], ref, what='slice conv')
    assert float(yd[..., :8].abs().max()) == 0 and float(yd[..., 24:].abs().max()) == 0


def test_conv_mixed_dtypes(call, lib, oracle):
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 8, 8, 32, generator=g)
    w = torch.randn(3, 3, 32, 2, generator=g) * 0.1
    xd = cu(x, torch.bfloat16)
    yd = torch.empty(2, 8, 8, 2, device='cuda')
    call('phs_conv2d', call.T(xd), cu(w), None, call.T(yd), 3, 0, 0, lib.IMPL_SIMT)
    ref = oracle.conv2d_same(xd.float().cpu().double(), w.double())
    close(yd, ref, what='bf16 in, f32 out')
    # f32 in (a latent sample z), bf16 out: the first convolution behind every z in fast mode
    z = torch.randn(3, 16, 16, 2, generator=g)
    wz = torch.randn(3, 3, 2, 64, generator=g) * 0.3
    bz = torch.randn(64, generator=g)
    yz = torch.empty(3, 16, 16, 64, device='cuda', dtype=torch.bfloat16)
    call('phs_conv2d', call.T(cu(z)), cu(wz), cu(bz), call.T(yz), 3, 0, 0, lib.IMPL_SIMT)
    refz = oracle.conv2d_same(z.double(), wz.double(), bz.double())
    close(yz, refz, rtol=2 ** -8, what='f32 in, bf16 out (z input)')
    # ... and its filter gradient: f32 z, bf16 dy (all nine taps per thread, wgrad_small3_kernel), against the same sum in fp64
    gy = torch.randn(3, 16, 16, 64, generator=g).to(torch.bfloat16)
    zp = torch.nn.functional.pad(z.double(), (0, 0, 1, 1, 1, 1))
    gw_ref = torch.stack([torch.einsum('nhwc,nhwo->co', zp[:, kh:kh + 16, kw:kw + 16], gy.double())
                          for kh in range(3) for kw in range(3)]).reshape(3, 3, 2, 64)
    gwz = torch.zeros(3, 3, 2, 64, device='cuda')
    call('phs_conv2d_wgrad', call.T(cu(z)), call.T(gy.cuda()), gwz, None, 3, 0, lib.IMPL_SIMT)
    close(gwz, gw_ref, rtol=5e-5, what='wgrad of the z-input convolution (f32 x, bf16 dy)')


def test_conv_argument_errors(call, lib):
    x = torch.zeros(1, 4, 4, 8, device='cuda')
    y = torch.zeros(1, 4, 4, 8, device='cuda')
    w = torch.zeros(5, 5, 8, 8, device='cuda')
    assert call.rc('phs_conv2d', call.T(x), w, None, call.T(y), 5, 0, 0, lib.IMPL_SIMT) < 0
    assert b'ksize' in lib.load().phs_last_error()
    y2 = torch.zeros(1, 2, 2, 8, device='cuda')
    assert call.rc('phs_conv2d', call.T(x), w, None, call.T(y2), 3, 0, 0, lib.IMPL_SIMT) < 0
    assert call.rc('phs_conv2d', None, w, None, call.T(y), 3, 0, 0, lib.IMPL_SIMT) < 0


# ---------------------------------------------------------------------------------------------------------
# normalisation (tfwrapper/normalisation.py:17-36,145-163)
# ---------------------------------------------------------------------------------------------------------
def _norm_ref(oracle, mode, y, gamma, beta, mm, mv):
    if mode == 'gn':
        P = {'s/gamma': gamma.reshape(1, 1, 1, -1), 's/beta': beta.reshape(1, 1, 1, -1)}
        return F.relu(oracle.group_norm2d(y, P, 's')), None
    P = {'s/BatchNorm/gamma': gamma, 's/BatchNorm/beta': beta, 's/BatchNorm/moving_mean': mm,
         's/BatchNorm/moving_variance': mv}
    ns = {}
    out = F.relu(oracle.batch_norm(y, P, 's', mode == 'bn_train', ns))
    return out, ns


@pytest.mark.parametrize('mode', ['bn_train', 'bn_infer', 'gn'])
@pytest.mark.parametrize('shape', [(3, 8, 8, 32), (2, 16, 16, 192), (4, 2, 2, 64)])
def test_norm_fwd_bwd(call, lib, oracle, mode, shape):
    N, H, W, C = shape
    g = torch.Generator().manual_seed(C + N)
    y = (torch.randn(shape, generator=g, dtype=torch.float64) * 2 + 0.5).requires_grad_(True)
    gamma = (torch.rand(C, generator=g, dtype=torch.float64) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, generator=g, dtype=torch.float64) * 0.3).requires_grad_(True)
    mm = torch.randn(C, generator=g, dtype=torch.float64) * 0.1 + 0.5
    mv = torch.rand(C, generator=g, dtype=torch.float64) + 3.5
    ref, ns = _norm_ref(oracle, mode, y, gamma, beta, mm, mv)
    ga = torch.randn(shape, generator=g, dtype=torch.float64)
    ref.backward(ga)
    lmode = {'bn_train': lib.NORM_BN_TRAIN, 'bn_infer': lib.NORM_BN_INFER, 'gn': lib.NORM_GN}[mode]
    eps = 1e-5 if mode == 'gn' else 1e-3
    yd, gd, bd = cu(y, torch.float32), cu(gamma, torch.float32), cu(beta, torch.float32)
    mmd, mvd = cu(mm, torch.float32), cu(mv, torch.float32)
    mmd2, mvd2 = mmd.clone(), mvd.clone()
    stats = torch.empty(N * C * 2, device='cuda', dtype=torch.float64)   # statistics buffers are fp64 (reproducible atomics)
    mean = torch.empty(N * C, device='cuda')
    rstd = torch.empty(N * C, device='cuda')
    call('phs_chan_stats', call.T(yd), stats)
    close(stats.view(N, C, 2)[..., 0], y.sum(dim=(1, 2)), what='chan sums')
    close(stats.view(N, C, 2)[..., 1], (y * y).sum(dim=(1, 2)), what='chan sumsq')
    call('phs_norm_finalize', stats, N, H * W, C, lmode, eps, 0.99, mmd, mvd, mean, rstd)
    ad = torch.empty_like(yd)
    call('phs_norm_act_fwd', call.T(yd), mean, rstd, gd, bd, 1, call.T(ad))
    close(ad, ref, what='norm+relu fwd')
    if mode == 'bn_train':
        close(mmd, ns['s/BatchNorm/moving_mean'], what='moving mean')
        close(mvd, ns['s/BatchNorm/moving_variance'], what='moving variance (Bessel corrected)')
    if mode == 'bn_infer':
        return
    # finalize folded into the activation kernel: per-sample sums followed by the batch totals (phs_conv2d_stats_acc layout)
    stats2 = torch.cat([stats, stats.view(N, C, 2).sum(dim=0).reshape(-1)]).contiguous()
    mean2, rstd2, ad2 = torch.empty_like(mean), torch.empty_like(rstd), torch.empty_like(yd)
    call('phs_norm_act_fwd_stats', call.T(yd), stats2, lmode, eps, 0.99, mmd2 if mode == 'bn_train' else None,
         mvd2 if mode == 'bn_train' else None, mean2, rstd2, gd, bd, 1, call.T(ad2))
    close(ad2, ref, what='fused finalize + norm + relu fwd')
    close(mean2, mean, rtol=1e-6, what='fused mean')
    close(rstd2, rstd, rtol=1e-6, what='fused rstd')
    if mode == 'bn_train':
        close(mmd2, mmd, rtol=1e-6, what='fused moving mean')
        close(mvd2, mvd, rtol=1e-6, what='fused moving variance')
    gad = cu(ga, torch.float32)
    sums = torch.empty(N * C * 2, device='cuda', dtype=torch.float64)
    coef = torch.empty(N * C * 2, device='cuda')
    dgam = torch.zeros(C, device='cuda')
    dbet = torch.zeros(C, device='cuda')
    dbias = torch.zeros(C, device='cuda')
    dyd = torch.empty_like(yd)
    call('phs_norm_bwd_reduce', call.T(gad), call.T(yd), mean, rstd, gd, bd, 1, sums)
    call('phs_norm_bwd_finalize', sums, stats, mean, rstd, gd, N, H * W, C, lmode, coef, dgam, dbet, dbias, 1)
    call('phs_norm_bwd_apply', call.T(gad), call.T(yd), mean, rstd, gd, bd, 1, coef, call.T(dyd))
    close(dyd, y.grad, rtol=2e-4, what='norm bwd dx')
    close(dgam, gamma.grad, rtol=2e-4, what='dgamma')
    close(dbet, beta.grad, rtol=2e-4, what='dbeta')
    # gradient of a conv bias added before the norm = sum over pixels of dy
    close(dbias, y.grad.sum(dim=(0, 1, 2)), rtol=1e-3, atol=1e-3 * float(y.grad.abs().sum(dim=(0, 1, 2)).max()) + 1e-5,
          what='dbias')
    if mode == 'bn_train':
        # two-launch batch-norm variant: totals-only reduction onto a cleared buffer, finalize folded into the apply
        tot = torch.zeros(C * 2, device='cuda', dtype=torch.float64)
        dgam2 = torch.full((C,), 0.5, device='cuda')
        dbet2 = torch.full((C,), -0.25, device='cuda')
        dyd2 = torch.empty_like(yd)
        call('phs_norm_bwd_reduce_bn', call.T(gad), call.T(yd), mean, rstd, gd, bd, 1, tot)
        close(tot, sums.view(N, C, 2).sum(dim=0).reshape(-1).cpu(), rtol=1e-6, what='batch totals')
        call('phs_norm_bwd_apply_bn', call.T(gad), call.T(yd), mean, rstd, gd, bd, 1, tot, call.T(dyd2), dgam2, dbet2, 1)
        assert torch.equal(dyd2, dyd), 'two-launch batch-norm backward differs from the three-launch one'
        close(dgam2 - 0.5, gamma.grad, rtol=2e-4, what='dgamma (accumulated onto 0.5)')
        close(dbet2 + 0.25, beta.grad, rtol=2e-4, what='dbeta (accumulated onto -0.25)')


# ---------------------------------------------------------------------------------------------------------
# resampling (tfwrapper/layers.py:44-54,336-345)
# ---------------------------------------------------------------------------------------------------------
def test_legacy_bilinear_known_answer(call):
    """hand-computed: TF1 legacy bilinear x2, align_corners=False: out[2k]=in[k], out[2k+1]=(in[k]+in[min(k+1,n-1)])/2"""
    x = torch.tensor([[1.0, 3.0], [5.0, 11.0]]).reshape(1, 2, 2, 1)
    want = torch.tensor([[1, 2, 3, 3], [3, 5, 7, 7], [5, 8, 11, 11], [5, 8, 11, 11]], dtype=torch.float32)
    xd = cu(x)
    yd = torch.empty(1, 4, 4, 1, device='cuda')
    call('phs_upsample2_fwd', call.T(xd), call.T(yd))
    assert torch.equal(yd.cpu().reshape(4, 4), want)


@pytest.mark.parametrize('shape', [(2, 4, 4, 8), (1, 8, 16, 2), (2, 2, 2, 192), (3, 16, 16, 3)])
def test_pool_upsample_fwd_bwd(call, oracle, shape):
    g = torch.Generator().manual_seed(sum(shape))
    N, H, W, C = shape
    x = torch.randn(shape, generator=g, dtype=torch.float64).requires_grad_(True)
    up = oracle.bilinear_upsample2d(x)
    gu = torch.randn(up.shape, generator=g, dtype=torch.float64)
    up.backward(gu)
    xd = cu(x, torch.float32)
    ud = torch.empty(N, 2 * H, 2 * W, C, device='cuda')
    call('phs_upsample2_fwd', call.T(xd), call.T(ud))
    close(ud, up, what='upsample fwd')
    gxd = torch.ones_like(xd)
    call('phs_upsample2_bwd', call.T(cu(gu, torch.float32)), call.T(gxd), 1)
    close(gxd, x.grad + 1, what='upsample bwd (accumulate)')
    call('phs_upsample2_bwd', call.T(cu(gu, torch.float32)), call.T(gxd), 0)
    close(gxd, x.grad, what='upsample bwd')
    x.grad = None
    po = oracle.averagepool2d(x)
    gp = torch.randn(po.shape, generator=g, dtype=torch.float64)
    po.backward(gp)
    pd = torch.empty(N, H // 2, W // 2, C, device='cuda')
    call('phs_avgpool2_fwd', call.T(xd), call.T(pd))
    close(pd, po, what='avgpool fwd')
    call('phs_avgpool2_bwd', call.T(cu(gp, torch.float32)), call.T(gxd), 0)
    close(gxd, x.grad, what='avgpool bwd')
    call('phs_avgpool2_bwd', call.T(cu(gp, torch.float32)), call.T(gxd), 1)
    close(gxd, 2 * x.grad, what='avgpool bwd (accumulate)')


# ---------------------------------------------------------------------------------------------------------
# latent heads + KL (posteriors.py:105-108; phiseg_model.py:210-226)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('N,hw,zd,gap', [(3, 64, 2, 0), (2, 4, 2, 0), (5, 1024, 2, 0), (3, 4, 6, 1)])
def test_latent_fwd_bwd(call, oracle, N, hw, zd, gap):
    g = torch.Generator().manual_seed(hw + zd)
    mk = lambda: torch.randn(N, hw, zd, generator=g, dtype=torch.float64).requires_grad_(True)
    mu_q, sp_q, mu_p, sp_p = mk(), mk(), mk(), mk()
    if gap:
        eps = torch.randn(N, zd, generator=g, dtype=torch.float64)
        mq, sq = mu_q.mean(1), F.softplus(sp_q).mean(1)
        mp, sg = mu_p.mean(1), F.softplus(sp_p).mean(1)
    else:
        eps = torch.randn(N, hw, zd, generator=g, dtype=torch.float64)
        mq, sq, mp, sg = mu_q, F.softplus(sp_q), mu_p, F.softplus(sp_p)
    z = mq + sq * eps
    kl = oracle.Oracle.KL_two_gauss_with_diag_cov(mq, sq, mp, sg)
    w = 4.0
    gz = torch.randn(z.shape, generator=g, dtype=torch.float64)
    (w * kl + (z * gz).sum()).backward()
    d = lambda t: cu(t, torch.float32)
    n_lat = N * zd if gap else N * hw * zd
    sig_q, sig_p, zd_ = (torch.empty(n_lat, device='cuda') for _ in range(3))
    mqo, mpo = torch.empty(n_lat, device='cuda'), torch.empty(n_lat, device='cuda')
    klo = torch.zeros(1, device='cuda')
    a = [d(mu_q), d(sp_q), d(mu_p), d(sp_p), d(eps)]
    call('phs_latent_fwd', a[0], a[1], a[2], a[3], a[4], N, hw, zd, gap, 0, mqo, sig_q, mpo, sig_p, zd_, klo, w / N)
    close(zd_, z.reshape(-1), what='z')
    close(sig_q, sq.reshape(-1), what='sigma_q')
    close(klo, (w * kl).reshape(1), rtol=2e-4, what='KL')
    # prior sample (generation mode, priors.py:100)
    call('phs_latent_fwd', None, None, a[2], a[3], a[4], N, hw, zd, gap, 1, None, None, mpo, sig_p, zd_, None, 0.0)
    close(zd_, (mp + sg * eps).reshape(-1), what='prior z')
    outs = [torch.empty(N * hw * zd, device='cuda') for _ in range(4)]
    call('phs_latent_bwd', d(gz), mqo if gap else a[0], a[1], sig_q, mpo if gap else a[2], a[3], sig_p, a[4], N, hw, zd,
         gap, w / N, *outs)
    for o, ref, nm in zip(outs, (mu_q, sp_q, mu_p, sp_p), ('dmu_q', 'dsp_q', 'dmu_p', 'dsp_p')):
        close(o, ref.grad.reshape(-1), rtol=2e-4, what=nm)


@pytest.mark.parametrize('shape', [(2, 8, 24, 64), (1, 16, 16, 192), (3, 4, 6, 32)])
def test_pool_upsample_bf16_vectors(call, oracle, shape):
    """The engine's fast mode runs these kernels on bf16 tensors with 8-channel (16-byte) vectors and multiply-high index
    arithmetic: non-square sizes, three vectors per pixel, several samples."""
    g = torch.Generator().manual_seed(sum(shape) + 1)
    N, H, W, C = shape
    xb = torch.randn(shape, generator=g).to(torch.bfloat16)
    x = xb.double().requires_grad_(True)
    up = oracle.bilinear_upsample2d(x)
    gub = torch.randn(up.shape, generator=g).to(torch.bfloat16)
    up.backward(gub.double())
    ud = torch.empty(N, 2 * H, 2 * W, C, device='cuda', dtype=torch.bfloat16)
    call('phs_upsample2_fwd', call.T(xb.cuda()), call.T(ud))
    close(ud, up, rtol=2 ** -8, atol=2 ** -8, what='upsample fwd bf16')
    gxd = torch.empty(shape, device='cuda', dtype=torch.bfloat16)
    call('phs_upsample2_bwd', call.T(gub.cuda()), call.T(gxd), 0)
    close(gxd, x.grad, rtol=2 ** -7, atol=2 ** -6, what='upsample bwd bf16')
    x.grad = None
    po = oracle.averagepool2d(x)
    gpb = torch.randn(po.shape, generator=g).to(torch.bfloat16)
    po.backward(gpb.double())
    pd = torch.empty(N, H // 2, W // 2, C, device='cuda', dtype=torch.bfloat16)
    call('phs_avgpool2_fwd', call.T(xb.cuda()), call.T(pd))
    close(pd, po, rtol=2 ** -8, atol=2 ** -8, what='avgpool fwd bf16')
    call('phs_avgpool2_bwd', call.T(gpb.cuda()), call.T(gxd), 0)
    close(gxd, x.grad, rtol=2 ** -8, atol=2 ** -8, what='avgpool bwd bf16')


# ---------------------------------------------------------------------------------------------------------
# multi-scale residual cross entropy (phiseg_model.py:229-262) and aggregation (:304-311)
# ---------------------------------------------------------------------------------------------------------
def _ptrs(ts):
    return (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


@pytest.mark.parametrize('N,H,nl,Lv', [(2, 32, 2, 5), (3, 16, 4, 3), (2, 16, 2, 1), (2, 8, 6, 2)])
def test_xent_multiscale(call, oracle, N, H, nl, Lv):
    g = torch.Generator().manual_seed(H + nl)
    logits = [torch.randn(N, H >> l, H >> l, nl, generator=g, dtype=torch.float64).requires_grad_(True) for l in range(Lv)]
    s = torch.randint(0, nl, (N, H, H), generator=g, dtype=torch.uint8)
    full = [oracle.nearest_upsample(lg, 1 << l) for l, lg in enumerate(logits)]
    acc, losses = None, [None] * Lv
    for l in reversed(range(Lv)):
        acc = full[l] if acc is None else acc + full[l]
        xe = F.cross_entropy(acc.reshape(-1, nl), s.reshape(-1).long(), reduction='none').reshape(N, -1)
        losses[l] = xe.sum(1).mean()
    sum(losses).backward()
    ld = [cu(t, torch.float32) for t in logits]
    gd = [torch.zeros_like(t) for t in ld]
    out = torch.zeros(Lv, device='cuda')
    lp, gp = _ptrs(ld), _ptrs(gd)
    call('phs_xent_multiscale', lp, gp, cu(s), N, H, H, nl, Lv, 1.0 / N, out)
    close(out, torch.stack(losses), rtol=2e-4, what='xent losses')
    for l in range(Lv):
        close(gd[l], logits[l].grad, rtol=2e-4, atol=1e-6, what='dlogits[%d]' % l)
    # forward only
    out.zero_()
    call('phs_xent_multiscale', lp, None, cu(s), N, H, H, nl, Lv, 1.0 / N, out)
    close(out, torch.stack(losses), rtol=2e-4, what='xent losses (fwd only)')
    # aggregation: s_out = sum of NN-upsampled levels; softmax; running softmax sum; argmax (bit-exact)
    s_out = torch.empty(N, H, H, nl, device='cuda')
    sm = torch.empty_like(s_out)
    smacc = torch.ones_like(s_out)
    am = torch.empty(N, H, H, dtype=torch.int64, device='cuda')
    call('phs_aggregate_logits', lp, N, H, H, nl, Lv, 1, s_out, sm, smacc, am)
    tot = sum(f.detach() for f in full)
    close(s_out, tot, what='s_out')
    close(sm, torch.softmax(tot, -1), what='softmax')
    close(smacc, torch.softmax(tot, -1) + 1, what='softmax accumulate')
    assert torch.equal(am.cpu(), s_out.cpu().argmax(-1)), 'argmax must be bit-exact w.r.t. the emitted logits'
    if N % 2 == 0:
        # rows = 2 samples of N/2 images (sample-major): the softmax sum over an image's samples lands on the image
        smacc2 = torch.zeros(N // 2, H, H, nl, device='cuda')
        call('phs_aggregate_logits', lp, N, H, H, nl, Lv, 2, s_out, sm, smacc2, am)
        close(s_out, tot, what='s_out (rep 2)')
        smr = torch.softmax(tot, -1)
        close(smacc2, smr[:N // 2] + smr[N // 2:], what='softmax summed over the samples of an image')
        assert torch.equal(am.cpu(), s_out.cpu().argmax(-1))
    am2 = torch.empty_like(am)
    call('phs_argmax_f32', smacc, N * H * H, nl, am2)
    assert torch.equal(am2.cpu(), smacc.cpu().argmax(-1))


# ---------------------------------------------------------------------------------------------------------
# optimizers (phiseg_model.py:134-141)
# ---------------------------------------------------------------------------------------------------------
def test_adam_matches_tf_form(call):
    n = 100003
    g0 = torch.Generator().manual_seed(1)
    p = torch.randn(n, generator=g0, dtype=torch.float64)
    m = torch.zeros(n, dtype=torch.float64)
    v = torch.zeros(n, dtype=torch.float64)
    pd, md, vd = cu(p, torch.float32), cu(m, torch.float32), cu(v, torch.float32)
    lr, b1, b2, eps = 1e-3, 0.9, 0.999, 1e-8
    lr_dev = torch.zeros(1, device='cuda')
    for t in range(1, 4):
        gr = torch.randn(n, generator=g0, dtype=torch.float64)
        lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        m = b1 * m + (1 - b1) * gr
        v = b2 * v + (1 - b2) * gr * gr
        p = p - lr_t * m / (v.sqrt() + eps)
        if t == 2:      # step size from device memory (CUDA-graph replay path)
            lr_dev.fill_(lr_t)
            call('phs_adam_step', pd, cu(gr * 2, torch.float32), md, vd, n, 0.0, lr_dev, b1, b2, eps, 0.5)
        else:
            call('phs_adam_step', pd, cu(gr, torch.float32), md, vd, n, lr_t, None, b1, b2, eps, 1.0)
    close(pd, p, rtol=1e-5, what='adam params')
    close(md, m, rtol=1e-5, what='adam m')
    close(vd, v, rtol=1e-5, what='adam v')


def test_momentum_nesterov(call):
    n = 5000
    g0 = torch.Generator().manual_seed(2)
    p = torch.randn(n, generator=g0, dtype=torch.float64)
    acc = torch.zeros(n, dtype=torch.float64)
    pd, ad = cu(p, torch.float32), cu(acc, torch.float32)
    for t in range(3):
        gr = torch.randn(n, generator=g0, dtype=torch.float64)
        acc = 0.9 * acc + gr                       # tf ApplyMomentum(use_nesterov=True)
        p = p - 0.01 * (gr + 0.9 * acc)
        call('phs_momentum_step', pd, cu(gr, torch.float32), ad, n, 0.01, None, 0.9, 1.0)
    close(pd, p, rtol=1e-5, what='momentum params')


def test_small_helpers(call, lib):
    x = torch.rand(2, 4, 4, 1)
    s = torch.randint(0, 3, (2, 4, 4), dtype=torch.uint8)
    out = torch.empty(2, 4, 4, 4, device='cuda')
    call('phs_posterior_input', cu(x), cu(s), 2, 4, 4, 1, 3, call.T(out))
    ref = torch.cat([x, F.one_hot(s.long(), 3).float() - 0.5], -1)
    assert torch.equal(out.cpu(), ref)
    z = torch.randn(2, 6)
    buf = torch.zeros(2, 4, 4, 10, device='cuda')
    call('phs_broadcast_z', cu(z), call.T(buf, 4, 6))
    assert torch.equal(buf[..., 4:].cpu(), z.reshape(2, 1, 1, 6).expand(2, 4, 4, 6))
    gz = torch.zeros(12, device='cuda')
    gb = torch.randn(2, 4, 4, 10).cuda()
    call('phs_broadcast_z_bwd', call.T(gb, 4, 6), gz, 0)
    close(gz.view(2, 6), gb[..., 4:].sum(dim=(1, 2)), what='broadcast_z bwd')
    a = torch.zeros(100, device='cuda')
    call('phs_fill_f32', a, 100, 2.5)
    call('phs_axpy_f32', a, a.clone(), 100, 2.0)
    assert float(a.min()) == 7.5 and float(a.max()) == 7.5
    acc = torch.zeros(1, device='cuda')
    # add_weight_decay over a segment table: loss += 0.5*wd*sum W^2 and g += wd*W on the listed ranges only
    pbuf = torch.arange(40, device='cuda', dtype=torch.float32) * 0.1
    gbuf = torch.ones(40, device='cuda')
    segs = torch.tensor([[4, 8], [20, 12]], device='cuda', dtype=torch.int64)
    lossw = torch.zeros(1, device='cuda')
    call('phs_weight_decay', pbuf, gbuf, segs, 2, 0.25, lossw)
    m = torch.zeros(40, dtype=torch.bool)
    m[4:12] = True
    m[20:32] = True
    pc = pbuf.cpu()
    close(lossw, (0.5 * 0.25 * (pc[m] ** 2).sum()).reshape(1), what='weight decay loss')
    close(gbuf, torch.where(m, 1.0 + 0.25 * pc, torch.ones(40)), what='weight decay gradient')
    call('phs_weight_decay', pbuf, None, segs, 2, 0.25, lossw)          # validation form: loss only
    close(lossw, (2 * 0.5 * 0.25 * (pc[m] ** 2).sum()).reshape(1), what='weight decay loss (no gradient)')
    call('phs_sumsq_f32', a, 100, 0.5, acc)
    close(acc, torch.tensor([0.5 * 100 * 7.5 ** 2]), what='sumsq')
    src = torch.randn(2, 4, 4, 8).cuda()
    dst = torch.zeros(2, 4, 4, 16, dtype=torch.bfloat16, device='cuda')
    call('phs_copy_cast', call.T(src), call.T(dst, 8, 8))
    assert torch.equal(dst[..., 8:].float(), src.to(torch.bfloat16).float())


@pytest.mark.parametrize('Cin,Co,dt', [(3, 32, torch.bfloat16), (1, 32, torch.float32), (5, 64, torch.bfloat16)])
def test_im2col3x3(call, lib, oracle, Cin, Co, dt):
    """phs_im2col3x3 followed by a 1x1 convolution with the HWIO filter flattened to [9*Cin, Cout] equals the 3x3 SAME
    convolution (tfwrapper/layers.py:123) of the network input"""
    g = torch.Generator().manual_seed(Cin + Co)
    x = torch.randn(2, 16, 8, Cin, generator=g).to(dt)
    xd = x.cuda()
    col = torch.full((2, 16, 8, Co), 7.0, device='cuda', dtype=torch.bfloat16)
    call('phs_im2col3x3', call.T(xd), call.T(col))
    w = torch.randn(3, 3, Cin, 4, generator=g, dtype=torch.float64)
    ref = oracle.conv2d_same(x.double(), w)
    wf = torch.zeros(Co, 4, dtype=torch.float64)
    wf[:9 * Cin] = w.reshape(9 * Cin, 4)
    got = col.double().cpu() @ wf
    close(got, ref, rtol=1e-2 if dt == torch.bfloat16 else 1e-2, what='im2col conv')
    assert float(col[..., 9 * Cin:].float().abs().max()) == 0.0
    # exactness of the gather itself: centre tap reproduces x
    assert torch.equal(col[..., 4 * Cin:5 * Cin].float().cpu(), x.to(torch.bfloat16).float())


@pytest.mark.parametrize('N,H,W,Cin,Cout', [(2, 32, 32, 128, 2), (3, 16, 16, 192, 2), (1, 16, 24, 256, 2), (2, 16, 16, 64, 4),
                                            (2, 8, 8, 32, 6), (5, 7, 9, 128, 2), (2, 16, 16, 96, 4), (2, 128, 128, 32, 2)])
@pytest.mark.parametrize('xdt', [torch.bfloat16, torch.float32])
def test_head_1x1_streaming_kernels(call, lib, oracle, N, H, W, Cin, Cout, xdt):
    """The 1x1 heads (z*_mu / z*_sigma / y_lvl* / prediction, posteriors.py:105-128, likelihoods.py:218) and their filter
    gradient: register-resident filters, several pixel groups per thread in flight, one-wave grids - every unroll / tail
    combination (pixel counts that do not divide the grid stride, 1..4 vectors per thread, 2 / 4 / 8 output lanes)."""
    g = torch.Generator().manual_seed(N * 100 + Cin + Cout + H)
    x = torch.randn(N, H, W, Cin, generator=g).to(xdt)
    w = torch.randn(1, 1, Cin, Cout, generator=g) * 0.1
    b = torch.randn(Cout, generator=g)
    xr = x.double()
    ref = oracle.conv2d_same(xr, w.double(), b.double())
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    yd = torch.full((N, H, W, Cout), 3.0, device='cuda')
    call('phs_conv2d', call.T(xd), wd, bd, call.T(yd), 1, 0, 0, lib.IMPL_SIMT)
    close(yd, ref, what='1x1 head fwd', rtol=2e-5)
    call('phs_conv2d', call.T(xd), wd, None, call.T(yd), 1, 0, 1, lib.IMPL_SIMT)
    close(yd, 2 * ref - b.double(), what='1x1 head fwd accumulate', rtol=2e-5)
    gy = torch.randn(N, H, W, Cout, generator=g)
    gw_ref = torch.einsum('nhwc,nhwo->co', xr, gy.double()).reshape(1, 1, Cin, Cout)
    gwd = torch.zeros_like(wd)
    call('phs_conv2d_wgrad', call.T(xd), call.T(gy.cuda()), gwd, None, 1, 0, lib.IMPL_SIMT)
    close(gwd, gw_ref, what='1x1 head wgrad', rtol=5e-5)
