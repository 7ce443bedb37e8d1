"""CPU check of the host-side launch planning of the halo-tile tcgen05 convolution (csrc/conv_halo.cu, reached through
phs_conv_halo_plan: no device work): for every 3x3 layer shape of phiseg_7_5 (128x128 B=64, 256x256 B=32) and the
Probabilistic U-Net the chosen geometry must fit the SM (shared memory with two / one CTA per SM, TMEM columns), keep a
usable filter ring and cover the image.  These are the invariants a bad geometry would break as a launch failure or a
hang on the GPU box."""
import ctypes

import pytest

SM_SMEM = 228 * 1024            # bytes of shared memory per SM
STATIC_PLUS_RESERVED = 1760 + 1024
PAIRS = [(32, 32), (32, 64), (64, 64), (64, 128), (128, 128), (128, 192), (192, 192), (256, 192), (192, 256), (192, 64),
         (64, 192), (192, 32), (32, 192), (192, 128), (128, 64), (64, 32), (160, 64), (64, 160), (96, 32), (320, 128),
         (128, 256), (224, 64), (384, 192)]


def _plan(lib, N, H, W, cin, cout, stats, acc_flags=0):
    h = lib.load()
    x = lib.phs_tensor(None, N, H, W, cin, cin, lib.PHS_BF16)
    y = lib.phs_tensor(None, N, H, W, cout, cout, lib.PHS_BF16)
    out = (ctypes.c_int * 12)()
    rc = h.phs_conv_halo_plan(ctypes.byref(x), ctypes.byref(y), acc_flags, int(stats), out)
    return rc, list(out)


@pytest.mark.parametrize('N,res', [(64, 128), (64, 64), (64, 32), (64, 16), (32, 256), (3, 64), (1, 16), (512, 128)])
def test_halo_geometry_fits_the_sm(lib, N, res):
    taken = 0
    for cin, cout in PAIRS:
        for stats in (False, True):
            for flags in (0, 2):
                rc, p = _plan(lib, N, res, res, cin, cout, stats, flags)
                assert rc in (0, 1), (cin, cout, rc)
                if rc == 0:
                    assert cout > 256 or res % 16 != 0, 'the halo kernel should take %d->%d at %d' % (cin, cout, res)
                    continue
                taken += 1
                ctas, S, na, nb, resident, G, acc_stages, tmem, smem, grid, tiles, BK = p
                tag = (N, res, cin, cout, stats, p)
                assert ctas in (1, 2) and S in (1, 2, 4, 8) and (res // 8) % S == 0, tag
                assert BK == (64 if cin % 64 == 0 else 32), tag
                assert smem + STATIC_PLUS_RESERVED <= SM_SMEM // ctas, tag
                assert tmem in (32, 64, 128, 256, 512) and tmem * ctas <= 512, tag
                assert S * cout <= tmem // acc_stages and acc_stages in (1, 2), tag
                assert na >= 1 and (nb >= 2 or resident), tag
                if resident:
                    assert nb == (cin // BK) * 9, tag
                assert G in (0, 32, 64) and (G != 64 or cout % 64 == 0) and (G == 0 or cout % 32 == 0), tag
                assert tiles == N * (res // 16) * (res // (8 * S)), tag
                assert 1 <= grid <= min(tiles, ctas * 148), tag
    assert taken > 0


def test_halo_plan_declines_what_it_cannot_take(lib):
    assert _plan(lib, 4, 8, 8, 64, 64, False)[0] == 0          # image smaller than one 16x8 tile: shifted-box kernel
    assert _plan(lib, 4, 32, 32, 64, 8, False)[0] == 0         # 8 output channels: small-channel kernel
    assert _plan(lib, 4, 32, 32, 48, 64, False)[0] == 0        # Cin not a multiple of 32


@pytest.mark.parametrize('N,res', [(64, 128), (64, 64), (64, 16), (32, 256), (3, 32), (512, 128)])
def test_wgrad_halo_geometry_is_one_resident_wave(lib, N, res):
    h = lib.load()
    for cin, cout in PAIRS:
        if cout > 256:
            continue
        x = lib.phs_tensor(None, N, res, res, cin, cin, lib.PHS_BF16)
        dy = lib.phs_tensor(None, N, res, res, cout, cout, lib.PHS_BF16)
        out = (ctypes.c_int * 12)()
        rc = h.phs_wgrad_halo_plan(ctypes.byref(x), ctypes.byref(dy), out)
        assert rc == 1, (cin, cout, rc)
        resident, nkh, n_acc, nb, cib, cob, stages, tmem, smem, items, splits, bps = list(out)
        tag = (N, res, cin, cout, list(out))
        bricks = N * (res // 16) * (res // 8)
        assert resident in (1, 2) and nkh in (1, 3) and stages >= 2, tag
        assert smem + 208 + 1024 <= SM_SMEM // resident, tag
        assert tmem in (32, 64, 128, 256, 512) and nkh * n_acc * nb <= tmem and tmem * resident <= 512, tag
        assert nb % 32 == 0 and nb * cob == cout and cib == (cin + 127) // 128 or cin in (32, 64), tag
        assert items == (1 if nkh == 3 else 3) * cib * cob, tag
        assert splits * bps >= bricks and (splits - 1) * bps < bricks, tag      # every brick exactly once
        # one wave of resident CTAs (the point of sizing the grid by residency), unless there are more items than slots
        assert items * splits <= max(resident * 148, items), tag
