"""-m gpu: the tcgen05 implicit-GEMM convolutions (conv_tc.cu) against the CPU oracle through the C-ABI.

The tensor-core path multiplies bf16 operands and accumulates in fp32.  The oracle is evaluated in fp64 on the SAME
bf16-rounded inputs / filters, so the only differences left are fp32 accumulation order (and, for bf16 outputs, one final
rounding): tolerance 2e-5 relative to the output scale for fp32 outputs, 2^-8 for bf16 outputs."""
import numpy as np
import pytest
import torch

from gpu_util import Caller, cu

pytestmark = pytest.mark.gpu


@pytest.fixture()
def call(lib):
    return Caller(lib)


def relerr(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def shadows(w):
    """bf16 K-major filter shadows in the two layouts phs_weight_prep writes (include/phiseg_sm100.h):
    fwd [cout][tap*cin + ci], dgrad [cin][(taps-1-tap)*cout + co]."""
    k, _, cin, cout = w.shape
    taps = k * k
    wf = w.reshape(taps, cin, cout)
    fwd = wf.permute(2, 0, 1).reshape(cout, taps * cin)
    dg = wf.flip(0).permute(1, 0, 2).reshape(cin, taps * cout)
    return fwd.contiguous().to(torch.bfloat16).cuda(), dg.contiguous().to(torch.bfloat16).cuda()


TC_CASES = [  # N, H, W, Cin, Cout, k
    (2, 16, 16, 64, 64, 3),      # one 64-channel chunk, two bricks per image
    (1, 128, 128, 32, 32, 3),    # BK=32 (64B swizzle), one image row per brick
    (3, 8, 8, 192, 192, 3),      # bricks spanning two images, ragged batch (3 images, TN=2)
    (5, 2, 2, 192, 192, 3),      # 32 images per brick, mostly out of range
    (2, 4, 4, 256, 192, 3),      # 4 chunks
    (2, 32, 32, 128, 128, 3),
    (2, 64, 64, 192, 64, 3),
    (1, 16, 16, 384, 192, 3),
    (2, 32, 32, 160, 64, 3),     # Cin multiple of 32 only (ProbUNet decoder/conv_5_1)
    (2, 16, 16, 96, 32, 1),      # 1x1
    (1, 24, 40, 64, 96, 3),      # sizes that are not powers of two: bricks overhang the image
    (1, 256, 256, 32, 32, 3),    # two bricks per image row
    (2, 32, 32, 192, 32, 3),     # wide input, 32 outputs: wgrad takes all three filter rows per CTA (two channel-slab boxes)
    (2, 32, 32, 32, 192, 3),     # BK=32 activations, 192 outputs: three 64-channel TMA-store groups per sub-tile
    (2, 16, 64, 128, 32, 3),
]


@pytest.mark.parametrize('N,H,W,Cin,Cout,k', TC_CASES)
def test_conv_tc_fwd_dgrad_wgrad(call, lib, oracle, N, H, W, Cin, Cout, k):
    g = torch.Generator().manual_seed(N * 1000 + Cin * 10 + Cout + H)
    x = torch.randn(N, H, W, Cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(k, k, Cin, Cout, generator=g) * (1.0 / np.sqrt(k * k * Cin))).to(torch.bfloat16)
    b = torch.randn(Cout, generator=g)
    gy = torch.randn(N, H, W, Cout, generator=g).to(torch.bfloat16)
    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    y = oracle.conv2d_same(xr, wr, b.double())
    y.backward(gy.double())
    xd, gyd, bd = x.cuda(), gy.cuda(), b.cuda()
    wf, wdg = shadows(w.float())
    # forward, fp32 output
    yd = torch.zeros(N, H, W, Cout, device='cuda')
    call('phs_conv2d', call.T(xd), wf, bd, call.T(yd), k, 0, 0, lib.IMPL_TC)
    torch.cuda.synchronize()
    e = relerr(yd, y)
    assert e < 2e-5, 'conv fwd f32: %.3e' % e
    # forward, bf16 output, then accumulate on top
    yb = torch.zeros(N, H, W, Cout, device='cuda', dtype=torch.bfloat16)
    call('phs_conv2d', call.T(xd), wf, bd, call.T(yb), k, 0, 0, lib.IMPL_TC)
    e = relerr(yb, y)
    assert e < 2 ** -8, 'conv fwd bf16: %.3e' % e
    call('phs_conv2d', call.T(xd), wf, None, call.T(yd), k, 0, 1, lib.IMPL_TC)
    e = relerr(yd, 2 * y - b.double())
    assert e < 4e-5, 'conv fwd accumulate: %.3e' % e
    # input gradient
    gxd = torch.zeros(N, H, W, Cin, device='cuda')
    call('phs_conv2d', call.T(gyd), wdg, None, call.T(gxd), k, 1, 0, lib.IMPL_TC)
    e = relerr(gxd, xr.grad)
    assert e < 2e-5, 'conv dgrad: %.3e' % e
    # filter gradient (+ bias gradient)
    gwd = torch.full((k, k, Cin, Cout), 7.0, device='cuda')
    gbd = torch.full((Cout,), 7.0, device='cuda')
    call('phs_conv2d_wgrad', call.T(xd), call.T(gyd), gwd, gbd, k, 0, lib.IMPL_TC)
    e = relerr(gwd, wr.grad)
    assert e < 2e-5, 'conv wgrad: %.3e' % e
    e = relerr(gbd, gy.double().sum(dim=(0, 1, 2)))
    assert e < 2e-5, 'bias grad: %.3e' % e
    call('phs_conv2d_wgrad', call.T(xd), call.T(gyd), gwd, None, k, 1, lib.IMPL_TC)
    e = relerr(gwd, 2 * wr.grad)
    assert e < 4e-5, 'conv wgrad accumulate: %.3e' % e


PAIR_CASES = [  # N, H, W, Cin, Cout: shapes the halo kernel takes (H % 16 == 0, W % 8 == 0)
    (2, 32, 32, 128, 128),       # streamed filter, two 64-channel chunks
    (3, 16, 16, 192, 192),       # odd number of tiles: rank 1 of the last pair owns an empty tile
    (2, 64, 64, 64, 128),
    (1, 16, 16, 384, 192),       # six chunks
    (2, 32, 32, 32, 192),        # BK = 32
    (2, 128, 128, 32, 32),       # filter resident in both halves
    (1, 16, 8, 64, 64),          # a single tile: one CTA of the pair idles
    (5, 16, 16, 128, 256),
]


@pytest.mark.parametrize('N,H,W,Cin,Cout', PAIR_CASES)
def test_conv_halo_cta_pairs(call, lib, oracle, monkeypatch, N, H, W, Cin, Cout):
    """conv_halo_kernel<BK, PAIR=true> (cta_group::2: clusters of two CTAs, each staging half of every filter tile, one
    M=256 tcgen05.mma per tap issued by the leader): forward with bf16 / fp32 outputs, fused statistics, accumulation and
    the input gradient against the fp64 oracle on the same bf16-rounded operands - and bit-identical to the single-CTA
    kernel for fp32 outputs (same products, same accumulation order inside the tensor core)."""
    g = torch.Generator().manual_seed(N * 100 + Cin + Cout + H)
    x = torch.randn(N, H, W, Cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(3, 3, Cin, Cout, generator=g) * (1.0 / np.sqrt(9 * Cin))).to(torch.bfloat16)
    b = torch.randn(Cout, generator=g)
    gy = torch.randn(N, H, W, Cout, generator=g).to(torch.bfloat16)
    xr = x.double().requires_grad_(True)
    y = oracle.conv2d_same(xr, w.double(), b.double())
    y.backward(gy.double())
    xd, gyd, bd = x.cuda(), gy.cuda(), b.cuda()
    wf, wdg = shadows(w.float())
    y1 = torch.zeros(N, H, W, Cout, device='cuda')
    monkeypatch.setenv('PHS_HALO_PAIR', '0')
    call('phs_conv2d', call.T(xd), wf, bd, call.T(y1), 3, 0, 0, lib.IMPL_TC)
    monkeypatch.setenv('PHS_HALO_PAIR', '1')
    y2 = torch.zeros(N, H, W, Cout, device='cuda')
    call('phs_conv2d', call.T(xd), wf, bd, call.T(y2), 3, 0, 0, lib.IMPL_TC)
    torch.cuda.synchronize()
    assert relerr(y2, y) < 2e-5, 'pair fwd f32: %.3e' % relerr(y2, y)
    assert torch.equal(y1, y2), 'pair and single-CTA kernels differ'
    call('phs_conv2d', call.T(xd), wf, None, call.T(y2), 3, 0, 1, lib.IMPL_TC)
    assert relerr(y2, 2 * y - b.double()) < 4e-5
    yb = torch.zeros(N, H, W, Cout, device='cuda', dtype=torch.bfloat16)
    acc = torch.zeros(N + 1, Cout, 2, device='cuda', dtype=torch.float64)
    call('phs_conv2d_stats_acc', call.T(xd), wf, bd, call.T(yb), 3, acc)
    assert relerr(yb, y) < 2 ** -8
    assert relerr(acc[:N, :, 0], y.sum(dim=(1, 2))) < 5e-3 and relerr(acc[:N, :, 1], (y * y).sum(dim=(1, 2))) < 5e-3
    assert relerr(acc[N], acc[:N].sum(dim=0)) < 1e-6
    gxd = torch.zeros(N, H, W, Cin, device='cuda')
    call('phs_conv2d', call.T(gyd), wdg, None, call.T(gxd), 3, 1, 0, lib.IMPL_TC)
    assert relerr(gxd, xr.grad) < 2e-5, 'pair dgrad: %.3e' % relerr(gxd, xr.grad)


def test_conv_tc_channel_slices(call, lib, oracle):
    """zero-copy tf.concat: operands addressed as channel slices of wider buffers (ld > C)"""
    g = torch.Generator().manual_seed(11)
    xb = torch.randn(2, 16, 16, 160, generator=g).to(torch.bfloat16)
    w = (torch.randn(3, 3, 64, 32, generator=g) * 0.05).to(torch.bfloat16)
    wf, wdg = shadows(w.float())
    xd = xb.cuda()
    yb = torch.zeros(2, 16, 16, 96, device='cuda', dtype=torch.bfloat16)
    call('phs_conv2d', call.T(xd, 32, 64), wf, None, call.T(yb, 64, 32), 3, 0, 0, lib.IMPL_TC)
    ref = oracle.conv2d_same(xb[..., 32:96].double(), w.double())
    assert relerr(yb[..., 64:96], ref) < 2 ** -8
    assert float(yb[..., :64].float().abs().max()) == 0
    # wgrad with both operands sliced
    gyb = torch.randn(2, 16, 16, 96, generator=g).to(torch.bfloat16)
    xr = xb[..., 32:96].double()
    wr = w.double().requires_grad_(True)
    oracle.conv2d_same(xr, wr).backward(gyb[..., 64:96].double())
    gwd = torch.zeros(3, 3, 64, 32, device='cuda')
    call('phs_conv2d_wgrad', call.T(xd, 32, 64), call.T(gyb.cuda(), 64, 32), gwd, None, 3, 1, lib.IMPL_TC)
    assert relerr(gwd, wr.grad) < 2e-5


def test_conv_tc_large_linearity(call, lib):
    """BASELINE.json's full size (B=64, 128x128, 128->128): size-independent property instead of a CPU reference:
    conv(x1 + x2) == conv(x1) + conv(x2) with the fp32 output, and a checksum against a strided sub-problem."""
    g = torch.Generator(device='cuda').manual_seed(3)
    N, H, W, C = 64, 128, 128, 128
    x1 = torch.randint(-4, 5, (N, H, W, C), generator=g, device='cuda').to(torch.bfloat16)
    x2 = torch.randint(-4, 5, (N, H, W, C), generator=g, device='cuda').to(torch.bfloat16)
    w = torch.randint(-2, 3, (3, 3, C, C), generator=g, device='cuda').float()
    wf = w.reshape(9, C, C).permute(2, 0, 1).reshape(C, 9 * C).contiguous().to(torch.bfloat16)
    outs = []
    for x in (x1, x2, x1 + x2):
        y = torch.empty(N, H, W, C, device='cuda')
        call('phs_conv2d', call.T(x), wf, None, call.T(y), 3, 0, 0, lib.IMPL_TC)
        outs.append(y)
    # small integers: every product and partial sum is exact in fp32, so linearity must hold bit for bit
    assert torch.equal(outs[0] + outs[1], outs[2])
    # one image against cuDNN-free torch on the GPU (exact integer arithmetic again)
    ref = torch.nn.functional.conv2d(x1[:1].float().permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=1)
    assert torch.equal(ref.permute(0, 2, 3, 1), outs[0][:1])


@pytest.mark.parametrize('N,H,W,Cin,Cout', [(3, 16, 16, 64, 64), (2, 32, 32, 128, 128), (2, 128, 128, 32, 32),
                                            (2, 64, 64, 192, 64), (3, 8, 8, 192, 192), (2, 16, 32, 256, 192)])
def test_conv_tc_fused_statistics(call, lib, oracle, N, H, W, Cin, Cout):
    """phs_conv2d_stats: convolution + the per-(sample, channel) sum / sum of squares batch_norm / group_norm2D need
    (tfwrapper/normalisation.py:27-34,156), taken from the fp32 accumulators in the epilogue (halo kernel) or by a
    separate pass (other shapes)."""
    g = torch.Generator().manual_seed(N + H + Cin + Cout)
    x = torch.randn(N, H, W, Cin, generator=g).to(torch.bfloat16)
    w = (torch.randn(3, 3, Cin, Cout, generator=g) * (1.0 / np.sqrt(9 * Cin))).to(torch.bfloat16)
    b = torch.randn(Cout, generator=g)
    y = oracle.conv2d_same(x.double(), w.double(), b.double())
    wf, _ = shadows(w.float())
    yb = torch.zeros(N, H, W, Cout, device='cuda', dtype=torch.bfloat16)
    stats = torch.full((N, Cout, 2), 5.0, device='cuda', dtype=torch.float64)
    call('phs_conv2d_stats', call.T(x.cuda()), wf, b.cuda(), call.T(yb), 3, stats)
    assert relerr(yb, y) < 2 ** -8
    s_ref = y.sum(dim=(1, 2))
    q_ref = (y * y).sum(dim=(1, 2))
    # the separate pass sees the bf16-rounded y, the fused epilogue the fp32 accumulators
    assert relerr(stats[..., 0], s_ref) < 5e-3
    assert relerr(stats[..., 1], q_ref) < 5e-3
    # accumulating variant (the engine clears one arena for all layers): adds onto what the caller left in stats
    acc = torch.zeros(N + 1, Cout, 2, device='cuda', dtype=torch.float64)    # per-sample sums + the [C][2] batch totals
    call('phs_conv2d_stats_acc', call.T(x.cuda()), wf, b.cuda(), call.T(yb), 3, acc)
    # few-tile layers (fewer than PHS_STATS_MIN_HW = 2048 pixels per image): phs_conv2d_stats_acc takes its statistics from a
    # separate pass over the bf16-rounded output, phs_conv2d_stats from the fp32 accumulators - bf16 rounding apart
    tol = 5e-3 if H * W < 2048 else 1e-4
    assert relerr(acc[:N], stats) < tol
    assert relerr(acc[N], stats.sum(dim=0)) < tol
    call('phs_conv2d_stats_acc', call.T(x.cuda()), wf, b.cuda(), call.T(yb), 3, acc)
    assert relerr(acc[:N], 2 * stats) < tol and relerr(acc[N], 2 * stats.sum(dim=0)) < tol
    assert relerr(acc[N], acc[:N].sum(dim=0)) < 1e-9          # the batch totals are the sum of the per-sample sums
