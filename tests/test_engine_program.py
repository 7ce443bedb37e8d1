"""CPU checks of the launch-program builder (engine.build_program): programs are built on the CPU device (nothing is
launched) and their structure is checked - lane discipline (every launch on a lane that has been forked or chained,
every side lane joined before the program ends: what CUDA-graph capture requires), the single arena fill in front, one
filter-gradient launch per live convolution writing into its own slot of the flat gradient buffer, and the algorithmic
FLOP count of SURVEY.md section 8d."""
import collections
import importlib

import pytest
import torch


@pytest.fixture(scope='module')
def E(pkg):
    return importlib.import_module('phiseg_code_b200.engine')


def _build(E, arch, kind, size=64, B=2, norm='batch_norm', nlabels=2):
    kw = dict(arch=arch, image_size=(size, size, 1), mode='fast', norm=norm, nlabels=nlabels)
    if arch == 'probunet':
        kw.update(zdim0=6, latent_levels=1)
    if arch == 'det_unet':
        kw.update(zdim0=6, latent_levels=1, KL_weight=None)
    cfg = E.NetConfig(**kw)
    P = E.Params(cfg, torch.device('cpu'))
    return cfg, P, E.build_program(cfg, P, B, kind, torch.device('cpu'))


def _check_lanes(steps):
    live = {0}
    dirty = set()
    for st in steps:
        fn, args, name = st
        if fn is None:
            if name == 'fork':
                for ln in args[0]:
                    live.add(ln)
                    dirty.add(ln)
            elif name == 'join':
                for ln in args[0]:
                    assert ln in live, 'join of a lane that was never forked: %s' % (ln,)
                    dirty.discard(ln)
            elif name == 'after':
                src, dst = args[0]
                assert src in live, 'dependency on an unknown lane %d' % src
                live.add(dst)
                dirty.add(dst)
            else:
                raise AssertionError('unknown sync step %r' % name)
            continue
        lane = getattr(st, 'lane', 0)
        assert lane in live, '%s launched on lane %d outside a fork / after' % (name, lane)
        if lane != 0:
            dirty.add(lane)
    assert not dirty, 'side lanes never joined: %s' % sorted(dirty)


@pytest.mark.parametrize('arch,kind,norm', [('phiseg', 'train', 'batch_norm'), ('phiseg', 'train', 'group_norm'),
                                            ('phiseg', 'eval', 'batch_norm'), ('phiseg', 'sample', 'batch_norm'),
                                            ('phiseg', 'posterior', 'group_norm'), ('phiseg', 'from_z', 'batch_norm'),
                                            ('probunet', 'train', 'batch_norm'), ('probunet', 'sample', 'group_norm'),
                                            ('det_unet', 'train', 'batch_norm'), ('det_unet', 'sample', 'batch_norm'),
                                            ('det_unet', 'eval', 'group_norm')])
def test_programs_keep_lane_discipline(E, arch, kind, norm):
    cfg, P, sp = _build(E, arch, kind, norm=norm)
    _check_lanes(sp.prog.steps)
    # the forward part alone and the backward part alone are captured separately under data parallelism
    _check_lanes(sp.prog.steps[:sp.n_fwd])
    _check_lanes(sp.prog.steps[sp.n_fwd:])
    assert sp.prog.launches() > 50


@pytest.mark.parametrize('arch,norm,rep', [('phiseg', 'batch_norm', 1), ('phiseg', 'group_norm', 4),
                                           ('probunet', 'batch_norm', 3), ('probunet', 'group_norm', 1)])
def test_sampling_program_splits_per_image_and_per_sample_parts(E, arch, norm, rep, monkeypatch):
    """predict() replays steps[:n_enc] once per batch of images and fills + steps[n_enc:] once per noise draw
    (phiseg_model.py:337-353 re-runs everything per sample): both parts must be capturable on their own (lane
    discipline), the per-image part must hold the whole x-only encoder (and, for the probabilistic U-Net, the U-Net) and
    nothing that reads eps, and with rep samples per image the rows downstream of the latents are rep * B."""
    for lanes in (True, False):
        if not lanes:
            monkeypatch.setenv('PHS_NO_LANES', '1')
        B = 2
        cfg = E.NetConfig(arch=arch, image_size=(64, 64, 1), mode='fast', norm=norm,
                          **(dict(zdim0=6, latent_levels=1) if arch == 'probunet' else {}))
        P = E.Params(cfg, torch.device('cpu'))
        sp = E.build_program(cfg, P, B, 'sample', torch.device('cpu'), rep=rep)
        steps = sp.prog.steps
        assert 0 <= sp.n_fills < sp.n_enc < len(steps)        # (batch norm at inference has no statistics arena)
        enc, rest = steps[:sp.n_enc], steps[:sp.n_fills] + steps[sp.n_enc:]
        _check_lanes(enc)
        _check_lanes(rest)
        names_enc = [s[2] for s in enc if s[0] is not None]
        names_rest = [s[2] for s in rest if s[0] is not None]
        assert 'phs_latent_fwd' not in names_enc and 'phs_aggregate_logits' not in names_enc
        assert names_rest.count('phs_latent_fwd') == cfg.L and names_rest.count('phs_aggregate_logits') == 1
        n_enc_convs = sum(1 for n in names_enc if n.startswith('phs_conv2d'))
        assert n_enc_convs >= 3 * cfg.R + (40 if arch == 'probunet' else 0)
        if rep > 1:
            assert 'phs_copy_cast' in names_enc           # per-image features tiled over the samples
        assert sp.x.N == B and sp.Bs == rep * B
        assert [tuple(e.shape)[0] for e in sp.eps] == [rep * B] * cfg.L
        assert sp.s_out.shape[0] == rep * B and sp.sm_accum.shape[0] == B
        for a in sp.logits:
            assert a.N == rep * B


@pytest.mark.parametrize('arch', ['phiseg', 'probunet'])
def test_data_parallel_program_orders_allreduce_after_its_writers(E, arch):
    """parallel.insert_gradient_allreduce: every bucket's all-reduce sits behind the last launch that writes into the
    bucket, on the communication lane, which first waits for every lane that wrote into it; the resulting backward list
    keeps the lane discipline CUDA-graph capture needs."""
    par = importlib.import_module('phiseg_code_b200.parallel')
    cfg, P, sp = _build(E, arch, 'train')
    bwd = sp.prog.steps[sp.n_fwd:]
    out = par.insert_gradient_allreduce(bwd, P, 2, allreduce=lambda t: (lambda stream: 0))
    _check_lanes(sp.prog.steps[:sp.n_fwd] + out)
    base = P.g.data_ptr()
    buckets = par.gradient_buckets(P)
    done = {}
    waits = set()
    for st in out:
        fn, args, name = st
        if fn is None:
            if name == 'after' and args[0][1] == par.COMM_LANE:
                waits.add(args[0][0])
            continue
        if name.startswith('allreduce'):
            assert st.lane == par.COMM_LANE
            g = name[len('allreduce['):-1]
            done[g] = set(waits)
            continue
        for a in args:
            if isinstance(a, int) and base <= a < base + 4 * P.n:
                off = (a - base) // 4
                grp = [g for g, lo, hi in buckets if lo <= off < hi][0]
                assert grp not in done, '%s writes into %s after its all-reduce' % (name, grp)
                # (the lane of a later writer must have been waited for; checked when the bucket is reduced)
    written = {g for g, lo, hi in buckets}
    assert set(done) <= written and len(done) >= 5
    for g in done:
        assert done[g], 'all-reduce of %s waits for no lane' % g


def test_training_program_structure(E):
    cfg, P, sp = _build(E, 'phiseg', 'train')
    steps = sp.prog.steps
    names = [s[2] for s in steps]
    # one fill per arena chunk in front of everything clears all fused-statistics buffers
    assert names[0] == 'phs_fill_f32' and sp.n_fwd > 1
    # (phs_conv2d_pre = the same with the producer's normalisation applied to the operand, PHS_FUSE_NORM; stats is argument 5
    # of both entry points)
    fused = [s for s in steps if s[2] == 'phs_conv2d_stats_acc' or (s[2] == 'phs_conv2d_pre' and s[1][5] is not None)]
    assert len(fused) > 80 and 'phs_conv2d_stats' not in names
    fills = [s for s in steps[:4] if s[2] == 'phs_fill_f32']
    lo = min(s[1][0] for s in fills)
    hi = max(s[1][0] + 4 * s[1][1] for s in fills)
    for s in fused:                     # every statistics buffer lies inside the cleared arena
        assert lo <= s[1][5] < hi
    # one filter gradient per live convolution, each into its own slot of the flat gradient buffer, on a wgrad lane
    wg = [s for s in steps if s[2] == 'phs_conv2d_wgrad']
    live_w = [n for n, (off, shape, kind) in P.table.items() if kind == 'W' and '_ups_to_' not in n or
              (kind == 'W' and '_ups_to_' in n and n.split('_ups_to_')[0][-1] == n.split('_ups_to_')[1][0])]
    assert len(wg) == len(live_w) == 131
    g0, g1 = P.g.data_ptr(), P.g.data_ptr() + 4 * P.n
    direct = [s[1][2] for s in wg if g0 <= s[1][2] < g1]
    assert len(set(direct)) == len(direct) >= 129        # the two im2col'ed input convs go through a scratch + axpy
    want = {P.ptr(n, 'g') for n in live_w}
    assert set(direct) <= want
    assert all(getattr(s, 'lane', 0) >= 3 for s in wg)
    # forward convolutions: one per live conv (+ none for the dead z*_ups_to_* branches)
    fwd_convs = [s for s in steps[:sp.n_fwd] if s[2] in ('phs_conv2d', 'phs_conv2d_stats_acc', 'phs_conv2d_pre')]
    assert len(fwd_convs) == 131


@pytest.mark.parametrize('arch,kind,norm,size', [('phiseg', 'train', 'batch_norm', 128), ('phiseg', 'train', 'group_norm', 64),
                                                 ('probunet', 'train', 'batch_norm', 64), ('phiseg', 'sample', 'batch_norm', 128),
                                                 ('phiseg', 'eval', 'group_norm', 64)])
def test_fused_norm_program(E, monkeypatch, arch, kind, norm, size):
    """conv -> norm -> ReLU -> conv fusion (PHS_FUSE_NORM=1): a fused pair drops the stand-alone normalisation launch of
    the producer; in training programs its adjoint re-materialises the activation (phs_norm_bwd_reduce_remat) BEFORE the
    consumer's filter gradient reads it, and that filter gradient still exists exactly once.  Everything else about the
    program (lanes, launch counts of the other entry points) is unchanged."""
    monkeypatch.setenv('PHS_BN_FOLD', '0')        # (inference-mode batch norm normally folds into the PRODUCER's epilogue)
    monkeypatch.setenv('PHS_FUSE_NORM', '0')
    cfg, P, sp0 = _build(E, arch, kind, size=size, norm=norm)
    monkeypatch.setenv('PHS_FUSE_NORM', '1')
    cfg, P, sp1 = _build(E, arch, kind, size=size, norm=norm)
    c0 = collections.Counter(s[2] for s in sp0.prog.steps if s[0] is not None)
    c1 = collections.Counter(s[2] for s in sp1.prog.steps if s[0] is not None)
    nf = c1['phs_conv2d_pre']
    assert c0['phs_conv2d_pre'] == 0 and nf >= 8
    _check_lanes(sp1.prog.steps)
    _check_lanes(sp1.prog.steps[:sp1.n_fwd])
    _check_lanes(sp1.prog.steps[sp1.n_fwd:])
    training = kind == 'train'
    infer_bn = norm == 'batch_norm' and not training
    gone = 'phs_norm_act_fwd' if infer_bn else 'phs_norm_act_fwd_stats'
    assert c0[gone] - c1[gone] == nf
    if infer_bn:
        assert c0['phs_norm_finalize'] - c1['phs_norm_finalize'] == nf
    assert c1['phs_norm_bwd_reduce_remat'] == (nf if training else 0)
    assert c0['phs_norm_bwd_reduce'] == c1['phs_norm_bwd_reduce'] + c1['phs_norm_bwd_reduce_remat']
    assert c0['phs_conv2d'] + c0['phs_conv2d_stats_acc'] == c1['phs_conv2d'] + c1['phs_conv2d_stats_acc'] + nf
    for name in ('phs_conv2d_wgrad', 'phs_norm_bwd_apply', 'phs_norm_bwd_finalize', 'phs_avgpool2_fwd', 'phs_upsample2_fwd'):
        assert c0[name] == c1[name], name
    assert sp0.conv_flop_fwd == sp1.conv_flop_fwd
    # a fused convolution reads a RAW conv output (written by an earlier forward conv), never a normalised activation
    ptr = lambda a: a._obj.ptr
    steps = sp1.prog.steps
    conv_out = set()
    norm_out = set()
    remat_at = {}
    for i, st in enumerate(steps):
        fn, args, name = st
        if name in ('phs_conv2d_stats_acc', 'phs_conv2d'):
            conv_out.add(ptr(args[3]))
        elif name == 'phs_conv2d_pre':
            assert ptr(args[0]) in conv_out and ptr(args[0]) not in norm_out
            conv_out.add(ptr(args[4]))
        elif name in ('phs_norm_act_fwd_stats', 'phs_norm_act_fwd'):
            norm_out.add(ptr(args[-1]))
        elif name == 'phs_norm_bwd_reduce_remat':
            remat_at[ptr(args[-1])] = i
    if training:
        # every activation that was never written in the forward pass is read by exactly one filter gradient, after its remat
        readers = collections.Counter()
        for i, st in enumerate(steps):
            if st[2] == 'phs_conv2d_wgrad' and ptr(st[1][0]) in remat_at:
                assert i > remat_at[ptr(st[1][0])]
                readers[ptr(st[1][0])] += 1
        assert len(readers) == nf and set(readers.values()) == {1}


def test_algorithmic_flops_match_the_survey(E):
    """25.03 GFLOP / image forward for phiseg_7_5 at 128x128, 22.04 for the Probabilistic U-Net."""
    cfg, P, sp = _build(E, 'phiseg', 'eval', size=128, B=1)
    assert abs(sp.conv_flop_fwd / 1e9 - 25.03) < 0.05
    cfg, P, sp = _build(E, 'probunet', 'eval', size=128, B=1)
    assert abs(sp.conv_flop_fwd / 1e9 - 22.04) < 0.05


@pytest.mark.parametrize('arch,norm', [('phiseg', 'batch_norm'), ('phiseg', 'group_norm'), ('probunet', 'batch_norm'),
                                       ('probunet', 'group_norm'), ('det_unet', 'batch_norm'), ('det_unet', 'group_norm')])
def test_variable_set_matches_the_oracle(E, oracle, arch, norm):
    """Same TF variable names and shapes in the engine's flat parameter buffer and in the oracle's parameter dict
    (checkpoints and set_weights / get_weights rely on it), and the SURVEY's parameter count for phiseg_7_5."""
    kw = dict(arch=arch, image_size=(128, 128, 1), mode='parity', norm=norm)
    if arch in ('probunet', 'det_unet'):
        kw.update(zdim0=6, latent_levels=1)
    cfg = E.NetConfig(**kw)
    spec = {n: tuple(shape) for n, shape, kind in E.build_spec(cfg)}
    okw = dict(zdim0=6, latent_levels=1) if arch in ('probunet', 'det_unet') else {}
    orc = oracle.Oracle(arch, image_size=(128, 128, 1), norm=norm, **okw)
    P = orc.init_params(seed=1)
    assert set(spec) == set(P), sorted(set(spec) ^ set(P))[:8]
    for n, shape in spec.items():
        assert tuple(P[n].shape) == shape, (n, shape, tuple(P[n].shape))
    if arch == 'phiseg' and norm == 'batch_norm':
        n_w = sum(int(torch.tensor(s).prod()) for n, s in spec.items() if n.endswith('/W'))
        assert abs(n_w / 1e6 - 18.68) < 0.1          # conv weights incl. the dead z*_ups_to_* branches


def test_latest_checkpoint_resolution(pkg, tmp_path):
    """Counterpart of tfwrapper/utils.py:189-210 (get_latest_model_checkpoint_path): the highest iteration with the
    given prefix wins; other prefixes and files are ignored; no checkpoint -> None."""
    pm = importlib.import_module('phiseg_code_b200.phiseg.phiseg_model')
    assert pm._latest_checkpoint(str(tmp_path), 'model.ckpt') is None
    assert pm._latest_checkpoint(str(tmp_path / 'missing'), 'model.ckpt') is None
    for f in ('model.ckpt-500.npz', 'model.ckpt-12000.npz', 'model.ckpt-9000.npz', 'model_best_loss.ckpt-99999.npz',
              'model.ckpt-13000.txt', 'notes.npz'):
        (tmp_path / f).write_bytes(b'')
    assert pm._latest_checkpoint(str(tmp_path), 'model.ckpt').endswith('model.ckpt-12000.npz')
    assert pm._latest_checkpoint(str(tmp_path), 'model_best_loss.ckpt').endswith('model_best_loss.ckpt-99999.npz')
    assert pm._latest_checkpoint(str(tmp_path), 'model_best_ged.ckpt') is None
    # tf.train.Saver(max_to_keep=...) (phiseg_model.py:144-148): all but the newest `keep` files of one prefix go
    pm._prune_checkpoints(str(tmp_path), 'model.ckpt', keep=1)
    left = sorted(f.name for f in tmp_path.iterdir())
    assert left == ['model.ckpt-12000.npz', 'model.ckpt-13000.txt', 'model_best_loss.ckpt-99999.npz', 'notes.npz']
    (tmp_path / 'model.ckpt-12000.npz.tmp77.npz').write_bytes(b'')        # a half-written file is never 'latest'
    assert pm._latest_checkpoint(str(tmp_path), 'model.ckpt').endswith('model.ckpt-12000.npz')
    # learning-rate schedule lookup (phiseg_model.py:189-190, utils.find_floor_in_list)
    assert pm.find_floor_in_list([0, 1000, 5000], 999) == 0
    assert pm.find_floor_in_list([0, 1000, 5000], 1000) == 1000
    assert pm.find_floor_in_list([5000, 0, 1000], 7000) == 5000


@pytest.mark.parametrize('arch,kind', [('phiseg', 'sample'), ('phiseg', 'eval'), ('probunet', 'sample'), ('phiseg', 'from_z')])
def test_inference_batch_norm_folds_into_the_convolution(E, monkeypatch, arch, kind):
    """Sampling / validation programs under batch norm: every tensor-core layer is ONE launch (phs_conv2d_post: convolution +
    moving-statistics normalisation + ReLU in the epilogue) instead of conv -> phs_norm_finalize -> phs_norm_act_fwd; only
    the layers the CUDA-core small-channel kernels take keep the separate passes.  Training programs and group norm are
    untouched."""
    monkeypatch.setenv('PHS_BN_FOLD', '0')
    cfg, P, sp0 = _build(E, arch, kind, size=128)
    monkeypatch.setenv('PHS_BN_FOLD', '1')
    cfg, P, sp1 = _build(E, arch, kind, size=128)
    c0 = collections.Counter(s[2] for s in sp0.prog.steps if s[0] is not None)
    c1 = collections.Counter(s[2] for s in sp1.prog.steps if s[0] is not None)
    nf = c1['phs_conv2d_post']
    assert c0['phs_conv2d_post'] == 0 and nf >= 25
    assert c0['phs_norm_act_fwd'] - c1['phs_norm_act_fwd'] == nf
    assert c0['phs_norm_finalize'] - c1['phs_norm_finalize'] == nf
    assert c0['phs_conv2d'] - c1['phs_conv2d'] == nf
    assert c1['phs_norm_act_fwd'] <= 13               # the network-input and z-input layers (1 - 3 input channels)
    assert sp1.prog.launches() == sp0.prog.launches() - 2 * nf
    assert sp0.conv_flop_fwd == sp1.conv_flop_fwd
    assert sp1.prog.bytes < sp0.prog.bytes                # the raw convolution outputs are not even allocated
    _check_lanes(sp1.prog.steps)
    for a, k, n in (('phiseg', 'train', 'batch_norm'), ('phiseg', 'sample', 'group_norm')):
        cfg, P, sp = _build(E, a, k, norm=n)
        assert not any(s[2] == 'phs_conv2d_post' for s in sp.prog.steps)


def test_det_unet_program(E):
    """likelihoods.det_unet2D (likelihoods.py:10-79) with the dummy posterior / prior (experiments/detunet.py): the
    probabilistic U-Net's U-Net without encoders and without z - 3 * 7 encoder + 3 * 6 decoder + 3 recombination convs +
    the prediction head, one cross-entropy level, no latent kernels, every convolution with its filter gradient."""
    cfg, P, sp = _build(E, 'det_unet', 'train', size=128)
    c = collections.Counter(s[2] for s in sp.prog.steps if s[0] is not None)
    n_conv = 3 * 7 + 3 * 6 + 3 + 1
    assert c['phs_conv2d_wgrad'] == n_conv
    assert c['phs_latent_fwd'] == 0 and c['phs_latent_bwd'] == 0 and c['phs_xent_multiscale'] == 1
    assert not any(n.startswith(('posterior/', 'prior/')) for n in P.table)
    assert len(sp.eps) == 0 and len(sp.logits) == 1 and sp.logits[0].C == 2
    assert cfg.latent_shapes(3) == []
    # FLOPs: the probabilistic U-Net's likelihood minus the z channels of recomb_0
    cfg2, P2, sp2 = _build(E, 'probunet', 'eval', size=128, B=1)
    cfg1, P1, sp1 = _build(E, 'det_unet', 'eval', size=128, B=1)
    assert 0.65 < sp1.conv_flop_fwd / sp2.conv_flop_fwd < 0.85       # (the other 25 % are the posterior and prior encoders)
