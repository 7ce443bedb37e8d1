"""Host-side engine of the B200-native PHiSeg path.

It turns an architecture description (the graph phiseg/phiseg_model.py:37-141 builds out of
phiseg/model_zoo/{posteriors,priors,likelihoods}.py) into *static programs*: flat lists of C-ABI kernel launches
with pre-bound arguments over statically allocated NHWC device buffers.  A program is built once per
(kind, batch size), replayed every step, and can be captured into a CUDA graph because nothing in it allocates,
synchronises or depends on host values (the Adam step size is read from device memory).

Backward is hand-written: every forward op pushes an emitter on a tape, the tape is walked in reverse to lay down
the adjoint launches (SURVEY.md section 8a "backward obligations").  PyTorch is used for device memory, streams,
the RNG for eps and torch.distributed; every arithmetic step of the path is a kernel of libphiseg_sm100.so.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import lib as L

BN_EPS, BN_DECAY, GN_EPS = 1e-3, 0.99, 1e-5   # tfwrapper/normalisation.py:145,157 and :17
FUSE_NORM_DEFAULT = '0'     # PHS_FUSE_NORM: conv -> norm -> ReLU -> conv fusion (Builder.conv, phs_conv2d_pre)


def num_channels(n0):
    """posteriors.py:59, priors.py:54, likelihoods.py:168"""
    return [n0, 2 * n0, 4 * n0, 6 * n0, 6 * n0, 6 * n0, 6 * n0]


class NetConfig:
    """Plain description of the network + loss + optimizer, derived from an experiment module by phiseg_model."""

    def __init__(self, arch='phiseg', image_size=(128, 128, 1), nlabels=2, zdim0=2, n0=32, resolution_levels=7,
                 latent_levels=5, norm='batch_norm', KL_weight=1.0, xent_weight=1.0, exponential_weighting=True,
                 weight_decay=None, optimizer='adam', mode='parity'):
        assert arch in ('phiseg', 'probunet', 'det_unet'), arch
        assert norm in ('batch_norm', 'group_norm'), norm
        assert mode in ('parity', 'fast', 'parity_tc'), mode
        self.arch = arch
        self.H, self.W, self.Cx = image_size
        self.nlabels, self.zdim0, self.n0 = nlabels, zdim0, n0
        self.R, self.L = resolution_levels, latent_levels
        self.norm = norm
        self.KL_weight, self.xent_weight = KL_weight, xent_weight
        self.exponential_weighting = exponential_weighting
        self.weight_decay = weight_decay
        self.optimizer = optimizer
        self.mode = mode
        self.nc = num_channels(n0)
        if arch in ('probunet', 'det_unet'):
            assert latent_levels == 1
        d = 1 << (resolution_levels - 1)
        if self.H % d or self.W % d:
            raise ValueError('image size %dx%d is not divisible by 2^(resolution_levels-1)=%d' % (self.H, self.W, d))

    def latent_shapes(self, B):
        if self.arch == 'det_unet':
            return []                   # posteriors.dummy / priors.dummy: no latent variables (likelihoods.py:10-79)
        if self.arch == 'probunet':
            return [(B, self.zdim0)]
        d = self.R - self.L
        return [(B, self.H >> (i + d), self.W >> (i + d), self.zdim0) for i in range(self.L)]


# ------------------------------------------------------------------------------------------------------------
# parameters
# ------------------------------------------------------------------------------------------------------------
def build_spec(cfg):
    """Variables in TF graph-construction order with the names the reference's variable scopes produce
    (layers.py:119-131, normalisation.py:24-25,156, phiseg_model.py:37-98).  Returns [(name, shape, kind)],
    kind in W, b, gamma, beta, moving_mean, moving_variance.  Includes the dead z*_ups_to_* branches of
    posteriors.py:112-118 (they own variables but feed nothing; they never receive a gradient)."""
    ent = []
    nc, R, Lv, z0, n0, nrm = cfg.nc, cfg.R, cfg.L, cfg.zdim0, cfg.n0, cfg.norm

    def conv(scope, k, cin, cout, normed, bias=None):
        ent.append((scope + '/W', (k, k, cin, cout), 'W'))
        has_bias = not (normed and nrm == 'batch_norm') if bias is None else bias
        if has_bias:
            ent.append((scope + '/b', (cout,), 'b'))
        if normed:
            if nrm == 'batch_norm':
                pre = scope + '/batch_norm/BatchNorm/'
                ent.extend([(pre + 'beta', (cout,), 'beta'), (pre + 'gamma', (cout,), 'gamma'),
                            (pre + 'moving_mean', (cout,), 'moving_mean'),
                            (pre + 'moving_variance', (cout,), 'moving_variance')])
            else:
                ent.extend([(scope + '/group_norm/gamma', (1, 1, 1, cout), 'gamma'),
                            (scope + '/group_norm/beta', (1, 1, 1, cout), 'beta')])

    if cfg.arch == 'phiseg':
        d = R - Lv
        for net, cin0 in (('posterior', cfg.Cx + cfg.nlabels), ('prior', cfg.Cx)):
            for i in range(R):
                conv('%s/z%d_pre_1' % (net, i), 3, cin0 if i == 0 else nc[i - 1], nc[i], True)
                conv('%s/z%d_pre_2' % (net, i), 3, nc[i], nc[i], True)
                conv('%s/z%d_pre_3' % (net, i), 3, nc[i], nc[i], True)
            for i in reversed(range(Lv)):
                if i == Lv - 1:
                    conv('%s/z%d_mu' % (net, i), 3, nc[i + d], z0, False)
                    conv('%s/z%d_sigma' % (net, i), 1, nc[i + d], z0, False)
                else:
                    for j in reversed(range(i + 1)):
                        conv('%s/z%d_ups_to_%d_c_1' % (net, i + 1, j + 1), 3, z0 if j == i else z0 * n0, z0 * n0, True)
                        conv('%s/z%d_ups_to_%d_c_2' % (net, i + 1, j + 1), 3, z0 * n0, z0 * n0, True)
                    conv('%s/z%d_input_1' % (net, i), 3, nc[i + d] + z0 * n0, nc[i], True)
                    conv('%s/z%d_input_2' % (net, i), 3, nc[i], nc[i], True)
                    conv('%s/z%d_mu' % (net, i), 1, nc[i], z0, False)
                    conv('%s/z%d_sigma' % (net, i), 1, nc[i], z0, False)
        net = 'likelihood'
        for i in range(Lv):
            conv('%s/z%d_post_1' % (net, i), 3, z0, nc[i], True)
            conv('%s/z%d_post_2' % (net, i), 3, nc[i], nc[i], True)
            for t in range(d):
                conv('%s/preups_%d/z%d_post' % (net, i, t), 3, nc[i], nc[i], True)
        for i in reversed(range(Lv - 1)):
            below = nc[Lv - 1] if i + 1 == Lv - 1 else nc[i + 1 + d]
            conv('%s/post_z%d_ups_c' % (net, i + 1), 3, below, nc[i], True)
            conv('%s/post_c_%d_1' % (net, i), 3, 2 * nc[i], nc[i + d], True)
            conv('%s/post_c_%d_2' % (net, i), 3, nc[i + d], nc[i + d], True)
        for i in range(Lv):
            conv('%s/y_lvl%d' % (net, i), 1, nc[Lv - 1] if i == Lv - 1 else nc[i + d], cfg.nlabels, False)
    else:
        # prob_unet2D passes add_bias explicitly (posteriors.py:25): off under batch_norm, on otherwise
        # det_unet2D (likelihoods.py:10-79): the same U-Net without the encoders and without z behind the decoder
        det = cfg.arch == 'det_unet'
        for net, cin0 in (() if det else (('posterior', cfg.Cx + cfg.nlabels), ('prior', cfg.Cx))):
            for i in range(R):
                for t in (1, 2, 3):
                    conv('%s/conv_%d_%d' % (net, i, t), 3, (cin0 if i == 0 else nc[i - 1]) if t == 1 else nc[i], nc[i], True)
            conv('%s/pre_mu' % net, 1, nc[R - 1], z0, False)
            conv('%s/pre_sigma' % net, 1, nc[R - 1], z0, False)
        net = 'likelihood'
        for i in range(R):
            for t in (1, 2, 3):
                conv('%s/encoder/conv_%d_%d' % (net, i, t), 3, (cfg.Cx if i == 0 else nc[i - 1]) if t == 1 else nc[i], nc[i], True)
        prev = nc[R - 1]
        for jj in range(R - 1):
            ii = R - jj - 1
            conv('%s/decoder/conv_%d_1' % (net, jj), 3, prev + nc[ii - 1], nc[ii], True)
            conv('%s/decoder/conv_%d_2' % (net, jj), 3, nc[ii], nc[ii], True)
            conv('%s/decoder/conv_%d_3' % (net, jj), 3, nc[ii], nc[ii], True)
            prev = nc[ii]
        conv('%s/recomb_0' % net, 1, prev + (0 if det else z0), nc[0], True)
        conv('%s/recomb_1' % net, 1, nc[0], nc[0], True)
        conv('%s/recomb_2' % net, 1, nc[0], nc[0], True)
        conv('%s/prediction' % net, 1, nc[0], cfg.nlabels, False)
    return ent


def he_normal(gen, shape):
    """tfwrapper/utils.py:225-226: variance_scaling_initializer(factor=2, mode=FAN_IN, uniform=False), i.e. a
    normal truncated at +-2 sigma (resampled) with sigma = sqrt(1.3 * 2 / fan_in), fan_in = kh*kw*Cin."""
    fan_in = shape[0] * shape[1] * shape[2]
    std = math.sqrt(1.3 * 2.0 / fan_in)
    w = torch.randn(shape, generator=gen, dtype=torch.float32)
    bad = w.abs() > 2.0
    while bool(bad.any()):
        w[bad] = torch.randn(int(bad.sum()), generator=gen, dtype=torch.float32)
        bad = w.abs() > 2.0
    return w * std


def tc_eligible(k, cin, cout):
    """Shapes the tcgen05 implicit-GEMM kernels take (conv_tc.cu)."""
    return cin % 32 == 0 and cout % 32 == 0 and cout <= 256


def pad_eligible(k, cin, cout):
    """3x3 convs of a 1..7-channel tensor that can run as im2col + 1x1 tensor-core conv (no input gradient)."""
    return k == 3 and cin <= 7 and cout % 32 == 0 and cout <= 256


class Params:
    """Flat fp32 master / gradient / optimizer-slot buffers addressed through a name table, plus the BN moving
    statistics and (fast mode) the bf16 filter shadows the tensor-core kernels read."""

    def __init__(self, cfg, device):
        self.cfg, self.device = cfg, device
        self.spec = build_spec(cfg)
        self.table, self.state_table = {}, {}
        off = soff = 0
        for name, shape, kind in self.spec:
            n = int(np.prod(shape))
            if kind in ('moving_mean', 'moving_variance'):
                self.state_table[name] = (soff, shape, kind)
                soff += n
            else:
                self.table[name] = (off, shape, kind)
                off += (n + 3) // 4 * 4          # keep every tensor 16-byte aligned
        self.n = off
        self.p = torch.zeros(off, dtype=torch.float32, device=device)
        self.g = torch.zeros(off, dtype=torch.float32, device=device)
        self.slots = None                        # (m, v) or (acc,) created at the first optimizer step
        self.state = torch.zeros(max(soff, 1), dtype=torch.float32, device=device)
        self.step = 0
        # tensor-core shadows
        self.shadow = None
        self.shadow_table = {}
        self.prep_table = None
        self.pad_table = {}
        self.shadow_lo = None
        if cfg.mode in ('fast', 'parity_tc'):
            rows, soff = [], 0
            for name, shape, kind in self.spec:
                if kind == 'W' and tc_eligible(shape[0], shape[2], shape[3]):
                    n = int(np.prod(shape))
                    taps = shape[0] * shape[1]
                    self.shadow_table[name] = (soff, soff + n)
                    rows.append([self.table[name][0], soff, soff + n, taps, shape[2], shape[3], 0])
                    soff += 2 * n
                elif (kind == 'W' and cfg.mode == 'fast' and pad_eligible(shape[0], shape[2], shape[3])
                      and not os.environ.get('PHS_NO_PAD')):
                    # network-input convs: im2col'ed 1x1 form, K = 9*cin zero padded to kp (phs_im2col3x3)
                    k9 = 9 * shape[2]
                    kp = 32 if k9 <= 32 else 64
                    self.pad_table[name] = (soff, kp)
                    rows.append([self.table[name][0], soff, -1, 1, k9, shape[3], kp])
                    soff += kp * shape[3]
            self.shadow = torch.zeros(max(soff, 8), dtype=torch.bfloat16, device=device)
            if cfg.mode == 'parity_tc':
                # low halves of the filters: w = bf16(w) + bf16(w - bf16(w)) to 16 mantissa bits (phs_weight_prep_lo)
                self.shadow_lo = torch.zeros_like(self.shadow)
            self.prep_table = torch.tensor(rows, dtype=torch.int64, device=device).reshape(-1, 7)
        self.init()

    # -- views ------------------------------------------------------------------------------------------
    def view(self, name, buf=None):
        if name in self.table:
            off, shape, _ = self.table[name]
            return (self.p if buf is None else buf)[off:off + int(np.prod(shape))].view(shape)
        off, shape, _ = self.state_table[name]
        return self.state[off:off + int(np.prod(shape))].view(shape)

    def ptr(self, name, which='p'):
        if name in self.table:
            base = {'p': self.p, 'g': self.g}[which]
            return base.data_ptr() + 4 * self.table[name][0]
        return self.state.data_ptr() + 4 * self.state_table[name][0]

    def has(self, name):
        return name in self.table or name in self.state_table

    def shadow_ptr(self, name, dgrad, lo=False):
        off = self.shadow_table[name][1 if dgrad else 0]
        return (self.shadow_lo if lo else self.shadow).data_ptr() + 2 * off

    def pad_shadow_ptr(self, name):
        return self.shadow.data_ptr() + 2 * self.pad_table[name][0]

    def names(self):
        return [n for n, _, _ in self.spec]

    # -- init / io --------------------------------------------------------------------------------------
    def init(self, seed=1234):
        """he_normal filters, zero biases, gamma=1, beta=0, moving_mean=0, moving_variance=1
        (tfwrapper/utils.py:214-271; tf.contrib.layers.batch_norm defaults)."""
        gen = torch.Generator().manual_seed(seed)
        for name, shape, kind in self.spec:
            if kind == 'W':
                v = he_normal(gen, shape)
            elif kind in ('gamma', 'moving_variance'):
                v = torch.ones(shape)
            else:
                v = torch.zeros(shape)
            self.view(name).copy_(v.to(self.device))
        self.slots = None
        self.step = 0
        self.refresh_shadow()

    def refresh_shadow(self, stream=None):
        if self.shadow is None or self.prep_table.numel() == 0 or self.p.device.type != 'cuda':
            return
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        L.check(L.load().phs_weight_prep(self.p.data_ptr(), self.shadow.data_ptr(), self.prep_table.data_ptr(),
                                         self.prep_table.shape[0], st), 'phs_weight_prep')
        if self.shadow_lo is not None:
            L.check(L.load().phs_weight_prep_lo(self.p.data_ptr(), self.shadow_lo.data_ptr(), self.prep_table.data_ptr(),
                                                self.prep_table.shape[0], st), 'phs_weight_prep_lo')

    def state_dict(self):
        return {n: self.view(n).detach().cpu().clone() for n in self.names()}

    def load_state_dict(self, sd, strict=True):
        missing = [n for n in self.names() if n not in sd]
        if strict and missing:
            raise KeyError('missing variables: %s' % missing[:5])
        for n in self.names():
            if n in sd:
                v = torch.as_tensor(np.asarray(sd[n]) if not torch.is_tensor(sd[n]) else sd[n])
                self.view(n).copy_(v.to(torch.float32).reshape(self.view(n).shape).to(self.device))
        self.refresh_shadow()

    def ensure_slots(self):
        if self.slots is None:
            k = 2 if self.cfg.optimizer == 'adam' else 1
            self.slots = tuple(torch.zeros_like(self.p) for _ in range(k))


# ------------------------------------------------------------------------------------------------------------
# device buffers
# ------------------------------------------------------------------------------------------------------------
_TORCH_DT = {L.PHS_F32: torch.float32, L.PHS_BF16: torch.bfloat16}
_ES = {L.PHS_F32: 4, L.PHS_BF16: 2}


class Buf:
    """One NHWC allocation [N,H,W,ld]; its gradient twin is created on demand."""

    def __init__(self, prog, N, H, W, ld, dtype, zero=False):
        self.prog, self.N, self.H, self.W, self.ld, self.dtype = prog, N, H, W, ld, dtype
        f = torch.zeros if zero else torch.empty
        self.t = f((N, H, W, ld), dtype=_TORCH_DT[dtype], device=prog.device)
        prog.bytes += self.t.numel() * _ES[dtype]
        prog.keep.append(self.t)     # launches hold raw pointers: the program owns every buffer it addresses
        self.gbuf = None
        self.gw = []          # channel ranges of the gradient already written by an adjoint launch

    def act(self, c_off=0, C=None):
        return Act(self, c_off, self.ld - c_off if C is None else C)


class Act:
    """A channel slice [c_off, c_off+C) of a Buf: the unit every kernel addresses (pitch = buf.ld)."""

    def __init__(self, buf, c_off, C):
        self.buf, self.c_off, self.C = buf, c_off, C
        self._desc = None
        self.pending = None      # conv -> norm -> ReLU -> conv fusion: this activation has not been written yet (Builder.conv)
        self.deferred = []       # ... and these emitters (the consumer's filter gradient) wait for its re-materialisation

    N = property(lambda s: s.buf.N)
    H = property(lambda s: s.buf.H)
    W = property(lambda s: s.buf.W)
    dtype = property(lambda s: s.buf.dtype)

    @property
    def ptr(self):
        return self.buf.t.data_ptr() + self.c_off * _ES[self.buf.dtype]

    def desc(self):
        if self._desc is None:
            self._desc = L.phs_tensor(self.ptr, self.N, self.H, self.W, self.C, self.buf.ld, self.buf.dtype)
        return ctypes.byref(self._desc)

    def tensor(self):
        return self.buf.t[..., self.c_off:self.c_off + self.C]

    # gradient bookkeeping (static, at program-build time)
    def grad(self):
        b = self.buf
        if b.gbuf is None:
            b.gbuf = Buf(b.prog, b.N, b.H, b.W, b.ld, b.dtype)
        return Act(b.gbuf, self.c_off, self.C)

    def grad_written(self):
        lo, hi = self.c_off, self.c_off + self.C
        for a, b in self.buf.gw:
            if a <= lo and hi <= b:
                return True
            if not (hi <= a or b <= lo):
                raise AssertionError('partially written gradient range (%d,%d) vs (%d,%d)' % (lo, hi, a, b))
        return False

    def mark_grad_written(self):
        lo, hi = self.c_off, self.c_off + self.C
        self.buf.gw = [(a, b) for a, b in self.buf.gw if not (lo <= a and b <= hi)] + [(lo, hi)]


class Step(tuple):
    """(fn, args, name) of one C-ABI launch, plus the lane (stream) it is enqueued on."""
    lane = 0


class Program:
    """A replayable list of kernel launches."""

    def __init__(self, device):
        self.device = device
        self.steps = []
        self.keep = []
        self.bytes = 0
        self.lib = L.load()
        self.graph = None

    def emit(self, name, *args, lane=0):
        st = Step((getattr(self.lib, name), args, name))
        st.lane = lane
        self.steps.append(st)

    def emit_sync(self, kind, lanes):
        """'fork': the side lanes wait for everything enqueued so far on lane 0; 'join': lane 0 waits for the side lanes."""
        st = Step((None, (tuple(lanes),), kind))
        st.lane = 0
        self.steps.append(st)

    def emit_after(self, src, dst):
        """Lane dst waits for everything enqueued so far on lane src (a one-way dependency, no region)."""
        st = Step((None, ((src, dst),), 'after'))
        st.lane = 0
        self.steps.append(st)

    def stats_vec(self, n):
        """Fused-statistics buffer of one layer (n DOUBLES: the library accumulates statistics with fp64 atomics so that a
        step is reproducible run to run), carved out of an arena that ONE fill at the start of the program clears (a
        memset node in front of every convolution sat on the critical path ~100 times per step)."""
        n = (2 * int(n) + 63) // 64 * 64          # in float32 units; every carve-out stays 256-byte aligned
        chunk = 1 << 22
        if not hasattr(self, '_arena') or self._arena_off + n > self._arena[-1].numel():
            if not hasattr(self, '_arena'):
                self._arena = []
            t = torch.zeros(max(chunk, n), dtype=torch.float32, device=self.device)
            self._arena.append(t)
            self.keep.append(t)
            self.bytes += 4 * t.numel()
            self._arena_off = 0
        v = self._arena[-1][self._arena_off:self._arena_off + n]
        self._arena_off += n
        return v

    def arena_fills(self):
        """[(ptr, count)] of the statistics arena chunks (only the used part of the last one)."""
        ar = getattr(self, '_arena', [])
        return [(t.data_ptr(), t.numel() if i < len(ar) - 1 else self._arena_off) for i, t in enumerate(ar)]

    def vec(self, n, zero=False, dtype=torch.float32):
        t = (torch.zeros if zero else torch.empty)(max(int(n), 1), dtype=dtype, device=self.device)
        self.keep.append(t)
        self.bytes += t.element_size() * t.numel()
        return t

    def dvec(self, n, zero=False):
        """float64 vector (statistics / reduction buffers, see stats_vec)"""
        return self.vec(n, zero, torch.float64)

    def run_eager(self, steps=None):
        """Enqueue the launches.  Lane 0 is the current stream; independent sub-graphs (the posterior and prior
        encoders, the likelihood towers) are emitted on side lanes = side streams, forked from / joined to lane 0
        with events.  Under CUDA-graph capture the same calls become parallel branches of the graph, which is what lets
        the many small launches of the coarse pyramid levels overlap."""
        main = torch.cuda.current_stream()
        streams = {0: main}
        for step in (self.steps if steps is None else steps):
            fn, args, name = step
            if fn is None and name == 'after':
                src, dst = args[0]
                if dst not in streams:
                    streams[dst] = self._side_stream(dst)
                ev = torch.cuda.Event()
                ev.record(streams[src])
                streams[dst].wait_event(ev)
                continue
            if fn is None:
                for ln in args[0]:
                    if ln not in streams:
                        streams[ln] = self._side_stream(ln)
                    ev = torch.cuda.Event()
                    if name == 'fork':
                        ev.record(main)
                        streams[ln].wait_event(ev)
                    else:
                        ev.record(streams[ln])
                        main.wait_event(ev)
                continue
            lane = getattr(step, 'lane', 0)
            if lane not in streams:
                raise RuntimeError('launch on lane %d outside a fork/join region' % lane)
            rc = fn(*args, streams[lane].cuda_stream)
            if rc:
                L.check(rc, name)

    def _side_stream(self, lane):
        if not hasattr(self, '_sides'):
            self._sides = {}
        if lane not in self._sides:
            self._sides[lane] = torch.cuda.Stream(device=self.device)
        return self._sides[lane]

    def launches(self, steps=None):
        return sum(1 for s in (self.steps if steps is None else steps) if s[0] is not None)


# ------------------------------------------------------------------------------------------------------------
# graph builder
# ------------------------------------------------------------------------------------------------------------
class Builder:
    """Lays down forward launches immediately and records adjoint emitters on a tape."""

    def __init__(self, cfg, params, B, training, want_grad, device):
        self.cfg, self.P, self.B = cfg, params, B
        self.training, self.want_grad = training, want_grad
        self.prog = Program(device)
        self.fwd_steps = self.prog.steps
        self.tape = []
        self.adt = L.PHS_BF16 if cfg.mode == 'fast' else L.PHS_F32      # activation dtype
        self.n_conv_flop = 0
        self.lane = 0
        self.use_lanes = os.environ.get('PHS_NO_LANES') is None
        # filter gradients hang off the backward chain (nothing but the optimizer consumes them): they go to their own
        # lane so the tensor-bound wgrad kernels overlap the HBM-bound normalisation adjoints of the layers below
        self.wlane = 3 if (self.use_lanes and os.environ.get('PHS_NO_WLANE') is None) else None
        self.wlanes_used = set()
        self._splits = {}       # parity_tc: (buffer, channel slice) -> its (hi, lo) bf16 pair, made once per activation
        # conv -> norm -> ReLU -> conv fusion (phs_conv2d_pre): the consumer convolution normalises its operand tile in
        # shared memory, the activation between the two convolutions is not written in the forward pass
        self.fuse = cfg.mode == 'fast' and os.environ.get('PHS_FUSE_NORM', FUSE_NORM_DEFAULT) != '0'
        self.n_fused = 0
        # PHS_BN_FOLD=0: keep inference-mode batch norm as separate launches (A/B switch)
        self.fold_bn = cfg.mode == 'fast' and cfg.norm == 'batch_norm' and not training and os.environ.get('PHS_BN_FOLD', '1') != '0'
        self.n_folded = 0
        # PHS_FUSE_MINCIN / PHS_FUSE_MAXHW: fuse only consumers with at least that many input channels / at most that many
        # pixels per image (the narrow 128x128 layers are bound by the TMA load path; the transform lengthens their
        # load -> MMA chain)
        self.fuse_mincin = int(os.environ.get('PHS_FUSE_MINCIN', '0'))
        self.fuse_maxhw = int(os.environ.get('PHS_FUSE_MAXHW', str(1 << 30)))

    # -- helpers ------------------------------------------------------------------------------------------
    def new(self, N, H, W, C, dtype=None, ld=None, zero=False):
        return Buf(self.prog, N, H, W, C if ld is None else ld, self.adt if dtype is None else dtype, zero).act(0, C)

    def emit(self, name, *args):
        self.prog.emit(name, *args, lane=self.lane)

    def split(self, x, cache=True):
        """fp32 activation -> (hi, lo) bf16 pair for the three-pass tensor-core product of mode 'parity_tc'"""
        key = (x.buf.t.data_ptr(), x.c_off, x.C)      # (the tensor outlives the Buf wrapper: the program keeps it)
        if cache and key in self._splits:
            return self._splits[key]
        hi = self.new(x.N, x.H, x.W, x.C, L.PHS_BF16)
        lo = self.new(x.N, x.H, x.W, x.C, L.PHS_BF16)
        self.emit('phs_split_bf16', x.desc(), hi.desc(), lo.desc())
        if cache:
            self._splits[key] = (hi, lo)
        return hi, lo

    # -- concurrency regions: fork() ... work on several lanes ... join(); the adjoints mirror them in reverse ---
    def fork(self, lanes):
        if not self.use_lanes:
            return
        self.prog.emit_sync('fork', lanes)
        if self.want_grad:
            self.tape.append(('join', tuple(lanes)))

    def join(self, lanes):
        if not self.use_lanes:
            return
        self.lane = 0
        self.prog.emit_sync('join', lanes)
        if self.want_grad:
            self.tape.append(('fork', tuple(lanes)))

    def set_lane(self, lane):
        self.lane = lane if self.use_lanes else 0

    def push_bwd(self, fn):
        fn.lane = self.lane
        self.tape.append(fn)

    def _norm_mode(self):
        if self.cfg.norm == 'group_norm':
            return L.NORM_GN, GN_EPS
        return (L.NORM_BN_TRAIN if self.training else L.NORM_BN_INFER), BN_EPS

    # -- layers.conv2D (tfwrapper/layers.py:94-145) ----------------------------------------------------------
    def realize(self, x):
        """Write a deferred activation now (its consumer turned out not to be a fusable convolution)."""
        if x.pending is not None:
            emit_norm = x.pending['emit']
            x.pending = None
            emit_norm()
        return x

    def conv(self, x, scope, k, cout, normed=True, relu=True, out=None, need_dx=True, out_dtype=None, fuse_next=False):
        """fuse_next: the caller promises that the result feeds exactly ONE operation, issued next on the same lane; if that
        operation is a 3x3 tensor-core convolution the normalisation + ReLU of this layer move into its operand path."""
        P, cfg, pr = self.P, self.cfg, self.prog
        wname = scope + '/W'
        cin = x.C
        assert P.table[wname][1] == (k, k, cin, cout), (scope, P.table[wname][1], (k, k, cin, cout))
        bias = P.ptr(scope + '/b') if P.has(scope + '/b') else None
        tc = (cfg.mode == 'fast' and tc_eligible(k, cin, cout) and x.dtype == L.PHS_BF16
              and (out_dtype in (None, L.PHS_BF16)))
        # fp32-accurate tensor-core mode: x*w = x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, three bf16 tcgen05 passes accumulated in
        # the fp32 output (the dropped x_lo*w_lo term is 2^-17 relative); everything else is the fp32 parity graph
        tc3 = (cfg.mode == 'parity_tc' and tc_eligible(k, cin, cout) and x.dtype == L.PHS_F32
               and out_dtype in (None, L.PHS_F32))
        self.n_conv_flop += 2 * x.N * x.H * x.W * k * k * cin * cout
        # network inputs (no input gradient): im2col once, then a 1x1 tensor-core conv forward and in the filter gradient
        pad_in = (cfg.mode == 'fast' and not need_dx and wname in P.pad_table and out_dtype in (None, L.PHS_BF16))
        k_real, cin_real, x_real = k, cin, x
        if pad_in:
            kp = P.pad_table[wname][1]
            xcol = self.new(x.N, x.H, x.W, kp, L.PHS_BF16)
            self.emit('phs_im2col3x3', x.desc(), xcol.desc())
            x, k, cin, tc = xcol, 1, kp, True
        impl = L.IMPL_TC if tc else L.IMPL_SIMT
        if pad_in:
            w_f = w_d = P.pad_shadow_ptr(wname)
        else:
            w_f = P.shadow_ptr(wname, False) if tc else P.ptr(wname)
            w_d = P.shadow_ptr(wname, True) if tc else P.ptr(wname)
        ydt = self.adt if out_dtype is None else out_dtype
        # inference-mode batch norm (moving statistics: known before the convolution runs) folds into the epilogue of every
        # tensor-core layer: one launch instead of three, the raw convolution output is never written (phs_conv2d_post)
        fold = (normed and tc and self.fold_bn and ydt == L.PHS_BF16 and self._norm_mode()[0] == L.NORM_BN_INFER
                and (out is None or (out.buf.ld % 8 == 0 and out.c_off % 8 == 0)))   # (tensor-core stores: 16-byte rows)
        fpre = None             # the producer's pending normalisation when this convolution applies it itself
        if fold:
            self.realize(x_real)
        if x_real.pending is not None:
            assert not pad_in, 'a deferred activation reached an im2col layer'
            pend = x.pending
            ok = (self.fuse and tc and k == 3 and ydt == L.PHS_BF16 and cin >= self.fuse_mincin
                  and x.H * x.W <= self.fuse_maxhw)
            if ok:
                probe_y = L.phs_tensor(0, x.N, x.H, x.W, cout, cout, L.PHS_BF16)
                plan = (ctypes.c_int * 12)()
                with_stats = int(normed and self._norm_mode()[0] != L.NORM_BN_INFER)
                ok = pr.lib.phs_conv2d_pre_plan(pend['y'].desc(), ctypes.byref(probe_y), with_stats, plan) == 1
            if ok:
                fpre = pend
                x.pending = None
                self.n_fused += 1
            else:
                self.realize(x)
        x_act = x               # the activation object (x is rebound below for the im2col path)

        def conv_fwd(dst, stats_ptr):
            """the forward launch of a tensor-core layer: plain, with fused statistics, or with the operand transform"""
            if fpre is not None:
                self.emit('phs_conv2d_pre', fpre['y'].desc(), ctypes.byref(fpre['struct']), w_f, bias, dst.desc(), stats_ptr)
            elif stats_ptr is not None:
                self.emit('phs_conv2d_stats_acc', x.desc(), w_f, bias, dst.desc(), k, stats_ptr)
            else:
                self.emit('phs_conv2d', x.desc(), w_f, bias, dst.desc(), k, 0, 0, impl)

        def conv3(src, w_hi, w_lo, b, dst, dgrad, acc):
            """dst (+)= conv(src, w) through the (hi, lo) split: three tensor-core launches"""
            sh, sl = src
            self.emit('phs_conv2d', sh.desc(), w_hi, b, dst.desc(), k, dgrad, acc, L.IMPL_TC)
            self.emit('phs_conv2d', sl.desc(), w_hi, None, dst.desc(), k, dgrad, 1, L.IMPL_TC)
            self.emit('phs_conv2d', sh.desc(), w_lo, None, dst.desc(), k, dgrad, 1, L.IMPL_TC)

        xs = None
        if tc3:
            xs = self.split(x)
            w_f, w_d = P.shadow_ptr(wname, False), P.shadow_ptr(wname, True)
            w_f_lo, w_d_lo = P.shadow_ptr(wname, False, lo=True), P.shadow_ptr(wname, True, lo=True)
        if not normed:
            y = out if out is not None else self.new(x.N, x.H, x.W, cout, ydt)
            if tc3:
                conv3(xs, w_f, w_f_lo, bias, y, 0, 0)
            else:
                conv_fwd(y, None)
            a = y
            nb = None
        elif fold:
            mode, eps = self._norm_mode()
            pre = scope + '/batch_norm/BatchNorm/'
            a = out if out is not None else self.new(x.N, x.H, x.W, cout, ydt)
            st = L.phs_norm_pre(None, mode, eps, BN_DECAY, P.ptr(pre + 'moving_mean'), P.ptr(pre + 'moving_variance'), None,
                                None, P.ptr(pre + 'gamma'), P.ptr(pre + 'beta'), int(relu))
            pr.keep.append(st)
            self.emit('phs_conv2d_post', x.desc(), w_f, bias, ctypes.byref(st), a.desc(), k)
            self.n_folded += 1
            nb = None
        else:
            y = self.new(x.N, x.H, x.W, cout, ydt)
            mode, eps = self._norm_mode()
            N, HW, C = x.N, x.H * x.W, cout
            fused_stats = tc and mode != L.NORM_BN_INFER
            stats = pr.stats_vec((N + 1) * C * 2) if fused_stats else pr.dvec(N * C * 2)
            if fused_stats:
                # statistics of the following norm come out of the conv epilogue (fp32 accumulators); the arena they
                # live in is cleared by one fill at the start of the program (build_program)
                conv_fwd(y, stats.data_ptr())
            elif tc3:
                conv3(xs, w_f, w_f_lo, bias, y, 0, 0)
            else:
                conv_fwd(y, None)
            if cfg.norm == 'batch_norm':
                pre = scope + '/batch_norm/BatchNorm/'
                gamma, beta = P.ptr(pre + 'gamma'), P.ptr(pre + 'beta')
                dgamma, dbeta = P.ptr(pre + 'gamma', 'g'), P.ptr(pre + 'beta', 'g')
                mm, mv = P.ptr(pre + 'moving_mean'), P.ptr(pre + 'moving_variance')
            else:
                pre = scope + '/group_norm/'
                gamma, beta = P.ptr(pre + 'gamma'), P.ptr(pre + 'beta')
                dgamma, dbeta = P.ptr(pre + 'gamma', 'g'), P.ptr(pre + 'beta', 'g')
                mm = mv = None
            mean, rstd = pr.vec(N * C), pr.vec(N * C)
            a = out if out is not None else self.new(x.N, x.H, x.W, cout, ydt)

            def emit_norm():
                if fused_stats:
                    # finalize folded into the activation kernel (mean / rstd still land in their arrays for the backward)
                    self.emit('phs_norm_act_fwd_stats', y.desc(), stats.data_ptr(), mode, eps, BN_DECAY, mm, mv,
                              mean.data_ptr(), rstd.data_ptr(), gamma, beta, int(relu), a.desc())
                else:
                    if mode != L.NORM_BN_INFER:
                        self.emit('phs_chan_stats', y.desc(), stats.data_ptr())
                    self.emit('phs_norm_finalize', stats.data_ptr(), N, HW, C, mode, eps, BN_DECAY, mm, mv, mean.data_ptr(),
                              rstd.data_ptr())
                    self.emit('phs_norm_act_fwd', y.desc(), mean.data_ptr(), rstd.data_ptr(), gamma, beta, int(relu), a.desc())

            if (fuse_next and self.fuse and out is None and y.dtype == L.PHS_BF16
                    and (fused_stats or mode == L.NORM_BN_INFER)):
                # deferred: the next operation decides (Builder.conv fuses, everything else calls realize())
                st = L.phs_norm_pre(stats.data_ptr() if mode != L.NORM_BN_INFER else None, mode, eps, BN_DECAY, mm, mv,
                                    mean.data_ptr() if self.want_grad else None, rstd.data_ptr() if self.want_grad else None,
                                    gamma, beta, int(relu))
                pr.keep.append(st)
                a.pending = {'y': y, 'struct': st, 'emit': emit_norm}
            else:
                emit_norm()
            nb = (mode, stats, mean, rstd, gamma, beta, dgamma, dbeta)

        if self.want_grad:
            def bwd():
                ga = a.grad()
                assert a.grad_written(), 'no gradient reaches %s' % scope
                if nb is not None:
                    mode, stats, mean, rstd, gamma, beta, dgamma, dbeta = nb
                    N, HW, C = x.N, x.H * x.W, cout
                    dy = self.new(x.N, x.H, x.W, cout, y.dtype)
                    dbias = P.ptr(scope + '/b', 'g') if bias is not None else None
                    if mode == L.NORM_BN_TRAIN and dbias is None and os.environ.get('PHS_BN_BWD2') and not a.deferred:
                        # batch norm in two launches: the reduction adds into batch totals that live in the arena the
                        # program clears once, and the apply kernel derives its coefficients (and dgamma / dbeta) from
                        # them - no memset node and no finalize launch per layer.  Bit-identical results, but measured
                        # SLOWER inside the step (12.36-12.38 vs 12.30 ms, tools/step_ab.py): the backward pass is bound
                        # by the SMs' total work, not by the length of a layer's chain, and N x more blocks now contend for
                        # the same 2C atomics.  Opt-in.
                        tot = pr.stats_vec(C * 2)
                        self.emit('phs_norm_bwd_reduce_bn', ga.desc(), y.desc(), mean.data_ptr(), rstd.data_ptr(), gamma,
                                  beta, int(relu), tot.data_ptr())
                        self.emit('phs_norm_bwd_apply_bn', ga.desc(), y.desc(), mean.data_ptr(), rstd.data_ptr(), gamma,
                                  beta, int(relu), tot.data_ptr(), dy.desc(), dgamma, dbeta, 1)
                    else:
                        sums, coef = pr.dvec(N * C * 2), pr.vec(N * C * 2)
                        if a.deferred or (os.environ.get('PHS_REMAT_ALWAYS') and cfg.mode == 'fast' and out is None
                                          and y.dtype == L.PHS_BF16):
                            # the forward pass never wrote a (its consumer normalised y on the fly): the reduction pass
                            # reads y anyway and writes a for the consumer's filter gradient, which is emitted now
                            # (PHS_REMAT_ALWAYS: test switch - the unfused program runs the same kernel variant, so that a
                            # fused / unfused comparison under batch norm is not blurred by the variants' last-bit differences)
                            self.emit('phs_norm_bwd_reduce_remat', ga.desc(), y.desc(), mean.data_ptr(), rstd.data_ptr(),
                                      gamma, beta, int(relu), sums.data_ptr(), a.desc())
                            for fn in a.deferred:
                                fn()
                            a.deferred = []
                        else:
                            self.emit('phs_norm_bwd_reduce', ga.desc(), y.desc(), mean.data_ptr(), rstd.data_ptr(), gamma,
                                      beta, int(relu), sums.data_ptr())
                        self.emit('phs_norm_bwd_finalize', sums.data_ptr(), stats.data_ptr(), mean.data_ptr(),
                                  rstd.data_ptr(), gamma, N, HW, C, mode, coef.data_ptr(), dgamma, dbeta, dbias, 1)
                        self.emit('phs_norm_bwd_apply', ga.desc(), y.desc(), mean.data_ptr(), rstd.data_ptr(), gamma, beta,
                                  int(relu), coef.data_ptr(), dy.desc())
                    db = None
                else:
                    dy = ga
                    db = P.ptr(scope + '/b', 'g') if bias is not None else None
                dys = self.split(dy, cache=False) if tc3 else None   # on the chain lane: filter and input gradient read them

                def emit_wgrad():
                    lane = self.lane
                    if self.wlane is not None:
                        # one filter-gradient lane per origin lane: the two encoders' gradients do not queue behind each other
                        wl = self.wlane + (lane if (lane in (1, 2) and os.environ.get('PHS_WLANES', '3') != '1') else 0)
                        pr.emit_after(lane, wl)
                        self.lane = wl
                        self.wlanes_used.add(wl)
                    if pad_in:
                        scratch = pr.vec(cin * cout)      # [kp][cout]: the first 9*cin_real rows are dW in HWIO order
                        self.emit('phs_conv2d_wgrad', x.desc(), dy.desc(), scratch.data_ptr(), db, 1, 0, impl)
                        self.emit('phs_axpy_f32', P.ptr(wname, 'g'), scratch.data_ptr(), 9 * cin_real * cout, 1.0)
                    elif tc3:
                        dW = P.ptr(wname, 'g')
                        self.emit('phs_conv2d_wgrad', xs[0].desc(), dys[0].desc(), dW, db, k, 1, L.IMPL_TC)
                        self.emit('phs_conv2d_wgrad', xs[1].desc(), dys[0].desc(), dW, None, k, 1, L.IMPL_TC)
                        self.emit('phs_conv2d_wgrad', xs[0].desc(), dys[1].desc(), dW, db, k, 1, L.IMPL_TC)
                    else:
                        self.emit('phs_conv2d_wgrad', x.desc(), dy.desc(), P.ptr(wname, 'g'), db, k, 1, impl)
                    self.lane = lane

                if fpre is not None:
                    # the input activation does not exist yet: the producer's adjoint re-materialises it (it comes next on
                    # the tape) and emits this filter gradient right behind
                    x_act.deferred.append(emit_wgrad)
                else:
                    emit_wgrad()
                if need_dx:
                    gx = x.grad()
                    acc = int(x.grad_written())
                    if tc3:
                        conv3(dys, w_d, w_d_lo, None, gx, 1, acc)
                    else:
                        self.emit('phs_conv2d', dy.desc(), w_d, None, gx.desc(), k, 1, acc, impl)
                    x.mark_grad_written()
            self.push_bwd(bwd)
        return a

    # -- layers.averagepool2D (tfwrapper/layers.py:44-54) ---------------------------------------------------
    def pool(self, x, out=None):
        self.realize(x)
        y = out if out is not None else self.new(x.N, x.H // 2, x.W // 2, x.C, x.dtype)
        self.emit('phs_avgpool2_fwd', x.desc(), y.desc())
        if self.want_grad:
            def bwd():
                assert y.grad_written()
                self.emit('phs_avgpool2_bwd', y.grad().desc(), x.grad().desc(), int(x.grad_written()))
                x.mark_grad_written()
            self.push_bwd(bwd)
        return y

    # -- layers.bilinear_upsample2D (tfwrapper/layers.py:336-345) -------------------------------------------
    def up(self, x, out=None, need_dx=True):
        self.realize(x)
        y = out if out is not None else self.new(x.N, x.H * 2, x.W * 2, x.C, x.dtype)
        self.emit('phs_upsample2_fwd', x.desc(), y.desc())
        if self.want_grad and need_dx:
            def bwd():
                assert y.grad_written()
                self.emit('phs_upsample2_bwd', y.grad().desc(), x.grad().desc(), int(x.grad_written()))
                x.mark_grad_written()
            self.push_bwd(bwd)
        return y

    # -- identity with its own gradient buffer: lets a consumer on another lane write "its" dz without racing the
    #    other consumers of x; the adjoint folds the copy's gradient back into x's
    def copy(self, x):
        self.realize(x)
        y = self.new(x.N, x.H, x.W, x.C, x.dtype)
        self.emit('phs_copy_cast', x.desc(), y.desc())
        if self.want_grad:
            def bwd():
                assert y.grad_written()
                if x.grad_written():
                    assert x.dtype == L.PHS_F32 and x.buf.ld == x.C and x.c_off == 0
                    self.emit('phs_axpy_f32', x.grad().ptr, y.grad().ptr, x.N * x.H * x.W * x.C, 1.0)
                else:
                    self.emit('phs_copy_cast', y.grad().desc(), x.grad().desc())
                    x.mark_grad_written()
            self.push_bwd(bwd)
        return y

    def emit_backward(self):
        for f in reversed(self.tape):
            if isinstance(f, tuple):            # mirrored concurrency region
                self.lane = 0
                self.prog.emit_sync(f[0], f[1])
                continue
            self.lane = getattr(f, 'lane', 0)
            f()
        self.lane = 0
        if self.wlanes_used:
            self.prog.emit_sync('join', sorted(self.wlanes_used))
        self.tape = []


# ------------------------------------------------------------------------------------------------------------
# networks
# ------------------------------------------------------------------------------------------------------------
class StepProgram:
    """One built graph: inputs, outputs and the launch lists (forward [+ backward])."""
    pass


def _ptr_array(ptrs):
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    return arr


def build_program(cfg, params, B, kind, device, rep=1):
    """B images.  rep > 1 (kind 'sample' only): rep samples per image are evaluated in ONE pass, batched along N
    (sample-major: row s*B + b); everything that depends on x alone (the prior's encoder pyramid - 2.74 of its 3.16 GFLOP -
    and, for the probabilistic U-Net, the whole U-Net) runs once per image at batch B and is tiled over the samples.  The
    launches of that per-image part are steps[:n_enc] (sp.n_enc): predict() replays only steps[n_enc:] for further noise
    draws of the same images (the reference re-runs the whole graph per sample, phiseg_model.py:337-353).
    kind:
      'train'      posterior + prior(generation_mode=False) + likelihood(posterior z) + ELBO, backward
                   (phiseg_model.py:37-59,75-83,113-141), training=True
      'eval'       the same forward with training=False (validation losses, phiseg_model.py:537-549)
      'sample'     prior(generation_mode=True) + likelihood(prior z) + aggregate (phiseg_model.py:61-109), training=False
      'posterior'  posterior only, training=False (generate_posterior_samples, :484-495)
      'from_z'     likelihood on given z (generate_samples_from_z, :313-322), training=False
    """
    training = kind == 'train'
    want_grad = kind == 'train'
    assert rep == 1 or kind == 'sample', 'samples are batched along N only in the sampling program'
    Bi = B              # images
    B = B * rep         # rows of everything downstream of a latent sample
    b = Builder(cfg, params, B, training, want_grad, device)
    pr = b.prog
    sp = StepProgram()
    sp.kind, sp.B, sp.prog, sp.cfg, sp.rep, sp.Bs = kind, Bi, pr, cfg, rep, B
    sp.n_enc = 0

    def tiled(a, out=None):
        """per-image activation -> one copy per sample (rows s*Bi + b)"""
        if rep == 1:
            return a
        t = out if out is not None else b.new(B, a.H, a.W, a.C, a.dtype)
        pr.emit('phs_copy_cast', a.desc(), t.desc(), lane=b.lane)
        return t
    H, W, Cx, nl, Lv, R, zd = cfg.H, cfg.W, cfg.Cx, cfg.nlabels, cfg.L, cfg.R, cfg.zdim0
    nc = cfg.nc
    f32 = L.PHS_F32
    # --- inputs
    sp.x = b.new(Bi, H, W, Cx, f32)
    sp.s = torch.zeros((Bi, H, W), dtype=torch.uint8, device=device)
    shapes = cfg.latent_shapes(B)
    sp.eps = [torch.zeros(s, dtype=torch.float32, device=device) for s in shapes]
    sp.losses = torch.zeros(2 * Lv + 2, dtype=torch.float32, device=device)   # [xent_l | KL_l | pad]
    need_post = kind in ('train', 'eval', 'posterior')
    need_prior = kind in ('train', 'eval', 'sample')
    need_lik = kind in ('train', 'eval', 'sample', 'from_z')
    gen_mode = kind == 'sample'

    if kind in ('train', 'eval'):
        pr.emit('phs_fill_f32', sp.losses.data_ptr(), sp.losses.numel(), 0.0)

    nets = []
    if cfg.arch == 'det_unet':
        need_post = need_prior = False      # posteriors.dummy / priors.dummy
    if need_post:
        pin = b.new(Bi, H, W, Cx + nl)
        pr.emit('phs_posterior_input', sp.x.ptr, sp.s.data_ptr(), Bi, H, W, Cx, nl, pin.desc())
        nets.append(('posterior', pin))
    if need_prior:
        nets.append(('prior', sp.x))

    sp.z = sp.mu = sp.sigma = sp.prior_mu = sp.prior_sigma = None
    if cfg.arch == 'phiseg':
        d = R - Lv
        zbufs = {}
        if nets:
            # --- encoders (posteriors.py:84-95): pre_z[r]; levels that are concatenated later are written
            # straight into the first channels of their concat buffer (tf.concat at posteriors.py:120 is free)
            pre_z = {}
            cat = {}
            two = len(nets) == 2
            if two:
                b.fork([1])
            for li, (net, inp) in enumerate(nets):
                b.set_lane(li if two else 0)
                h = inp
                for r in range(R):
                    if r > 0:
                        h = b.pool(h)
                    h = b.conv(h, '%s/z%d_pre_1' % (net, r), 3, nc[r], need_dx=r > 0, fuse_next=True)
                    h = b.conv(h, '%s/z%d_pre_2' % (net, r), 3, nc[r], fuse_next=True)
                    out = None
                    l = r - d
                    if 0 <= l < Lv - 1:
                        cbuf = Buf(pr, B, H >> r, W >> r, nc[r] + zd * cfg.n0, b.adt)
                        cat[(net, l)] = cbuf
                        out = cbuf.act(0, nc[r])
                    if rep > 1:
                        h = b.conv(h, '%s/z%d_pre_3' % (net, r), 3, nc[r])
                        if out is not None:
                            tiled(h, out)
                    else:
                        h = b.conv(h, '%s/z%d_pre_3' % (net, r), 3, nc[r], out=out)
                    pre_z[(net, r)] = h
                pre_z[(net, R - 1)] = tiled(pre_z[(net, R - 1)])
            if two:
                b.join([1])
            if kind == 'sample':
                sp.n_enc = len(pr.steps)
            # --- latent hierarchy, posterior and prior level by level (posteriors.py:98-130, priors.py:92-126)
            mu = {n: [None] * Lv for n, _ in nets}
            spre = {n: [None] * Lv for n, _ in nets}
            sig = {n: [None] * Lv for n, _ in nets}
            zl = [None] * Lv
            early = need_lik and b.use_lanes            # likelihood tower l starts as soon as z_l exists (lane 2)
            post_z, lcat = [None] * Lv, [None] * Lv
            for l in reversed(range(Lv)):
                hl, wl = H >> (l + d), W >> (l + d)
                # the x2 up-sampling of z_{l+1} stays on lane 0: both nets' adjoints accumulate into the same dz
                ups = {net: b.up(zl[l + 1]) for net, _ in nets} if l < Lv - 1 else {}
                if two:
                    b.fork([1])
                for li, (net, _) in enumerate(nets):
                    b.set_lane(li if two else 0)
                    if l == Lv - 1:
                        src = pre_z[(net, l + d)]
                        mu[net][l] = b.conv(src, '%s/z%d_mu' % (net, l), 3, zd, normed=False, out_dtype=f32)
                        spre[net][l] = b.conv(src, '%s/z%d_sigma' % (net, l), 1, zd, normed=False, out_dtype=f32)
                    else:
                        u = b.conv(ups[net], '%s/z%d_ups_to_%d_c_1' % (net, l + 1, l + 1), 3, zd * cfg.n0, fuse_next=True)
                        cbuf = cat[(net, l)]
                        b.conv(u, '%s/z%d_ups_to_%d_c_2' % (net, l + 1, l + 1), 3, zd * cfg.n0,
                               out=cbuf.act(nc[l + d], zd * cfg.n0))
                        zin = b.conv(cbuf.act(), '%s/z%d_input_1' % (net, l), 3, nc[l], fuse_next=True)
                        zin = b.conv(zin, '%s/z%d_input_2' % (net, l), 3, nc[l])
                        mu[net][l] = b.conv(zin, '%s/z%d_mu' % (net, l), 1, zd, normed=False, out_dtype=f32)
                        spre[net][l] = b.conv(zin, '%s/z%d_sigma' % (net, l), 1, zd, normed=False, out_dtype=f32)
                    sig[net][l] = b.new(B, hl, wl, zd, f32)
                if two:
                    b.join([1])
                b.set_lane(0)
                zl[l] = b.new(B, hl, wl, zd, f32)
                _emit_latent(b, sp, l, hl * wl, mu, spre, sig, zl[l], gen_mode, need_post, need_prior, gap=0)
                if early:
                    zt = b.copy(zl[l])
                    b.fork([2])
                    b.set_lane(2)
                    post_z[l] = _phiseg_tower(b, cfg, l, zt, lcat)
                    b.set_lane(0)
            if early:
                b.join([2])
                sp.logits = _phiseg_merge(b, cfg, post_z, lcat)
            sp.z = zl
            if need_post:
                sp.mu, sp.sigma = mu['posterior'], sig['posterior']
            if need_prior:
                sp.prior_mu, sp.prior_sigma = mu['prior'], sig['prior']
        else:
            sp.z = [b.new(*shapes[l], f32) for l in range(Lv)]          # fed by the caller ('from_z')
        if need_lik and not (nets and b.use_lanes):
            sp.logits = _phiseg_likelihood(b, cfg, sp.z)
    elif cfg.arch == 'det_unet':
        # likelihoods.det_unet2D: U-Net -> three 1x1 recombination convs -> prediction; nothing is sampled
        sp.z = []
        if need_lik:
            rc, hC = _probunet_unet(b, cfg, sp.x, tiled, zd=0)
            h = rc.act(0, hC)
            for t in range(3):
                h = b.conv(h, 'likelihood/recomb_%d' % t, 1, cfg.nc[0])
            sp.logits = [b.conv(h, 'likelihood/prediction', 1, cfg.nlabels, normed=False, out_dtype=L.PHS_F32)]
            if kind == 'sample':
                sp.n_enc = len(pr.steps)     # everything depends on x alone: only the aggregation is replayed per "draw"
    else:
        if nets:
            mu = {n: [None] for n, _ in nets}
            spre = {n: [None] for n, _ in nets}
            sig = {n: [None] for n, _ in nets}
            mu_out = {n: [None] for n, _ in nets}
            unet = None
            split = kind == 'sample'       # per-image part first and complete, so that it can be replayed on its own
            if need_lik and (b.use_lanes or split):
                b.fork([2])                 # (fork / join / set_lane are no-ops in single-stream programs)
                b.set_lane(2)
                unet = _probunet_unet(b, cfg, sp.x, tiled)
                b.set_lane(0)
            two = len(nets) == 2
            if two:
                b.fork([1])
            for li, (net, inp) in enumerate(nets):
                b.set_lane(li if two else 0)
                h = inp
                for r in range(R):
                    if r > 0:
                        h = b.pool(h)
                    for t in (1, 2, 3):
                        h = b.conv(h, '%s/conv_%d_%d' % (net, r, t), 3, nc[r], need_dx=not (r == 0 and t == 1),
                                   fuse_next=t < 3)
                mu[net][0] = tiled(b.conv(h, '%s/pre_mu' % net, 1, zd, normed=False, out_dtype=f32))
                spre[net][0] = tiled(b.conv(h, '%s/pre_sigma' % net, 1, zd, normed=False, out_dtype=f32))
                sig[net][0] = b.new(B, 1, 1, zd, f32)
                mu_out[net][0] = b.new(B, 1, 1, zd, f32)
            if two:
                b.join([1])
            b.set_lane(0)
            if split:
                if unet is not None:
                    b.join([2])
                sp.n_enc = len(pr.steps)
            z = b.new(B, 1, 1, zd, f32)
            hw = (H >> (R - 1)) * (W >> (R - 1))
            _emit_latent(b, sp, 0, hw, mu, spre, sig, z, gen_mode, need_post, need_prior, gap=1, mu_out=mu_out)
            sp.z = [z]
            if need_post:
                sp.mu, sp.sigma = mu_out['posterior'], sig['posterior']
            if need_prior:
                sp.prior_mu, sp.prior_sigma = mu_out['prior'], sig['prior']
            if unet is not None:
                if not split:
                    b.join([2])
                sp.logits = _probunet_head(b, cfg, sp.z[0], *unet)
        else:
            sp.z = [b.new(B, 1, 1, zd, f32)]
        if need_lik and not (nets and (b.use_lanes or kind == 'sample')):
            sp.logits = _probunet_likelihood(b, cfg, sp.z[0], sp.x)

    # --- heads of the graph
    sp.s_out = sp.s_out_sm = sp.sm_accum = sp.argmax = None
    if need_lik:
        nlev = len(sp.logits)
        lp = _ptr_array([a.ptr for a in sp.logits])
        pr.keep.append(lp)
        if kind in ('train', 'eval') and cfg.xent_weight is not None:
            dl = None
            if want_grad:
                for a in sp.logits[1:]:
                    g = a.grad()
                    pr.emit('phs_fill_f32', g.ptr, a.N * a.H * a.W * a.C, 0.0)
                dl = _ptr_array([a.grad().ptr for a in sp.logits])
                pr.keep.append(dl)
                for a in sp.logits:
                    a.mark_grad_written()
            pr.emit('phs_xent_multiscale', lp, dl, sp.s.data_ptr(), B, H, W, nl, nlev, cfg.xent_weight / B,
                    sp.losses.data_ptr())
        if kind in ('sample', 'from_z', 'eval'):
            sp.s_out = torch.empty((B, H, W, nl), dtype=torch.float32, device=device)
            sp.s_out_sm = torch.empty((B, H, W, nl), dtype=torch.float32, device=device)
            sp.sm_accum = torch.zeros((Bi, H, W, nl), dtype=torch.float32, device=device)   # summed over the rep samples
            sp.argmax = torch.empty((B, H, W), dtype=torch.int64, device=device)
            pr.emit('phs_aggregate_logits', lp, B, H, W, nl, nlev, rep, sp.s_out.data_ptr(), sp.s_out_sm.data_ptr(),
                    sp.sm_accum.data_ptr(), sp.argmax.data_ptr())
    sp.n_fwd = len(pr.steps)
    if want_grad:
        b.emit_backward()
    # one fill per arena chunk in front of everything clears the fused statistics / reduction buffers of ALL layers
    fills = []
    for ptr, cnt in pr.arena_fills():
        st = Step((pr.lib.phs_fill_f32, (ptr, cnt, 0.0), 'phs_fill_f32'))
        st.lane = 0
        fills.append(st)
    pr.steps[0:0] = fills
    sp.n_fwd += len(fills)
    sp.n_fills = len(fills)
    if sp.n_enc:
        sp.n_enc += len(fills)
    sp.conv_flop_fwd = b.n_conv_flop
    return sp


def _emit_latent(b, sp, l, hw, mu, spre, sig, z, gen_mode, need_post, need_prior, gap, mu_out=None):
    """softplus + reparameterisation + KL for one level (posteriors.py:105-108,125-128; phiseg_model.py:210-226,265-287)."""
    cfg, pr, B, zd = b.cfg, b.prog, b.B, b.cfg.zdim0
    q = 'posterior' if need_post else None
    p = 'prior' if need_prior else None
    w_l = float(4 ** l) if cfg.exponential_weighting else 1.0
    want_kl = bool(q and p and cfg.KL_weight is not None)
    kl_ptr = sp.losses.data_ptr() + 4 * (cfg.L + l) if want_kl else None
    g = lambda d, n: d[n][l].ptr if n else None
    pr.emit('phs_latent_fwd', g(mu, q), g(spre, q), g(mu, p), g(spre, p), sp.eps[l].data_ptr(), B, hw, zd, gap,
            int(gen_mode), g(mu_out, q) if gap else None, g(sig, q), g(mu_out, p) if gap else None, g(sig, p), z.ptr,
            kl_ptr, w_l / B)
    if b.want_grad:
        def bwd():
            dz = z.grad().ptr if z.grad_written() else None
            klw = (cfg.KL_weight * w_l / B) if want_kl else 0.0
            outs = [mu['posterior'][l], spre['posterior'][l], mu['prior'][l], spre['prior'][l]]
            mq = (mu_out if gap else mu)['posterior'][l]
            mp = (mu_out if gap else mu)['prior'][l]
            pr.emit('phs_latent_bwd', dz, mq.ptr, spre['posterior'][l].ptr, sig['posterior'][l].ptr, mp.ptr,
                    spre['prior'][l].ptr, sig['prior'][l].ptr, sp.eps[l].data_ptr(), B, hw, zd, gap, klw,
                    *[o.grad().ptr for o in outs])
            for o in outs:
                o.mark_grad_written()
        b.push_bwd(bwd)


def _phiseg_tower(b, cfg, i, z_i, cat):
    """Level i of likelihoods.phiseg before the top-down merge (likelihoods.py:196-199): z_i -> 2 convs -> d x
    [up-sample -> conv]; the last conv writes into the first half of the level's concat buffer."""
    nc, Lv, R = cfg.nc, cfg.L, cfg.R
    d = R - Lv
    h = b.conv(z_i, 'likelihood/z%d_post_1' % i, 3, nc[i], fuse_next=True)
    h = b.conv(h, 'likelihood/z%d_post_2' % i, 3, nc[i])
    for t in range(d):
        h = b.up(h)
        out = None
        if t == d - 1 and i < Lv - 1:
            cat[i] = Buf(b.prog, b.B, h.H, h.W, 2 * nc[i], b.adt)
            out = cat[i].act(0, nc[i])
        h = b.conv(h, 'likelihood/preups_%d/z%d_post' % (i, t), 3, nc[i], out=out)
    return h


def _phiseg_merge(b, cfg, post_z, cat):
    """Top-down merge and per-level heads of likelihoods.phiseg (likelihoods.py:204-221)."""
    nc, Lv, R = cfg.nc, cfg.L, cfg.R
    d = R - Lv
    if d == 0:
        raise NotImplementedError('resolution_levels == latent_levels')
    post_c = [None] * Lv
    post_c[Lv - 1] = post_z[Lv - 1]
    for i in reversed(range(Lv - 1)):
        u = b.up(post_c[i + 1])
        b.conv(u, 'likelihood/post_z%d_ups_c' % (i + 1), 3, nc[i], out=cat[i].act(nc[i], nc[i]))
        h = b.conv(cat[i].act(), 'likelihood/post_c_%d_1' % i, 3, nc[i + d], fuse_next=True)
        post_c[i] = b.conv(h, 'likelihood/post_c_%d_2' % i, 3, nc[i + d])
    # the per-level heads are independent of each other (1x1 convolutions onto nlabels channels, HBM bound): levels >= 1
    # run on their own lanes next to the big level-0 head, forward and (mirrored) backward
    side = [10 + i for i in range(1, Lv)] if b.use_lanes else []
    if side:
        b.fork(side)
    outs = []
    for i in range(Lv):
        b.set_lane(10 + i if (side and i > 0) else 0)
        outs.append(b.conv(post_c[i], 'likelihood/y_lvl%d' % i, 1, cfg.nlabels, normed=False, out_dtype=L.PHS_F32))
    if side:
        b.join(side)
    b.set_lane(0)
    return outs


def _phiseg_likelihood(b, cfg, z):
    """likelihoods.phiseg (likelihoods.py:162-223) on given latents.  Returns native-resolution head outputs (the
    nearest-neighbour resize of :221 is folded into the loss / aggregation kernels).  The per-level towers only depend
    on their own z: the coarse levels (small, latency-bound launches) run on a side lane."""
    Lv = cfg.L
    post_z, cat = [None] * Lv, [None] * Lv
    towers = Lv > 2
    if towers:
        b.fork([1])
    for i in range(Lv):
        b.set_lane(1 if (towers and i >= 2) else 0)
        post_z[i] = _phiseg_tower(b, cfg, i, z[i], cat)
    if towers:
        b.join([1])
    b.set_lane(0)
    return _phiseg_merge(b, cfg, post_z, cat)


def _probunet_likelihood(b, cfg, z, x):
    """likelihoods.prob_unet2D (likelihoods.py:81-159)."""
    rc, hC = _probunet_unet(b, cfg, x)
    return _probunet_head(b, cfg, z, rc, hC)


def _probunet_unet(b, cfg, x, tiled=None, zd=None):
    """The U-Net of likelihoods.prob_unet2D up to the point where z is tiled in (likelihoods.py:104-146): it does not
    depend on z, so it can run on its own lane next to the posterior / prior encoders."""
    nc, R = cfg.nc, cfg.R
    zd = cfg.zdim0 if zd is None else zd  # channels reserved behind the decoder output for the broadcast z (0: det_unet2D)
    pr = b.prog
    B, H, W = x.N, cfg.H, cfg.W          # images (the samples of one image share the whole U-Net)
    enc = []
    cats = {}
    h = x
    for i in range(R):
        if i > 0:
            h = b.pool(h)
        h = b.conv(h, 'likelihood/encoder/conv_%d_1' % i, 3, nc[i], need_dx=i > 0, fuse_next=True)
        h = b.conv(h, 'likelihood/encoder/conv_%d_2' % i, 3, nc[i], fuse_next=True)
        out = None
        if i < R - 1:
            # decoder stage jj = R-2-i concatenates [up(prev) | enc[i]] (crop_and_concat, layers.py:586-622)
            prev_c = nc[min(i + 2, R - 1)]       # channels of up(previous decoder stage / enc[R-1])
            cats[i] = Buf(pr, B, H >> i, W >> i, prev_c + nc[i], b.adt)
            out = cats[i].act(prev_c, nc[i])
        h = b.conv(h, 'likelihood/encoder/conv_%d_3' % i, 3, nc[i], out=out)
        enc.append(h)
    for jj in range(R - 1):
        ii = R - jj - 1
        cb = cats[ii - 1]
        b.up(h, out=cb.act(0, h.C))
        h = cb.act()
        for t in (1, 2, 3):
            out = None
            if jj == R - 2 and t == 3:
                rc = Buf(pr, B, H, W, nc[ii] + zd, b.adt)
                out = rc.act(0, nc[ii])
            h = b.conv(h, 'likelihood/decoder/conv_%d_%d' % (jj, t), 3, nc[ii], out=out, fuse_next=t < 3)
    if b.B != B:
        # one copy of the U-Net features per sample drawn for the image; z is tiled in behind them (likelihoods.py:147-151)
        big = Buf(pr, b.B, H, W, rc.ld, b.adt)
        tiled(rc.act(0, h.C), big.act(0, h.C))
        rc = big
    return rc, h.C


def _probunet_head(b, cfg, z, rc, hC):
    """z broadcast + the three 1x1 recombination convs + prediction head (likelihoods.py:147-157)."""
    nc, zd = cfg.nc, cfg.zdim0
    pr = b.prog
    zs = rc.act(hC, zd)
    pr.emit('phs_broadcast_z', z.ptr, zs.desc())          # tf.tile of z over H x W (likelihoods.py:147-151)
    if b.want_grad:
        def bwd():
            pr.emit('phs_broadcast_z_bwd', zs.grad().desc(), z.grad().ptr, int(z.grad_written()))
            z.mark_grad_written()
        b.push_bwd(bwd)
    h = rc.act()
    for t in range(3):
        h = b.conv(h, 'likelihood/recomb_%d' % t, 1, nc[0])
    return [b.conv(h, 'likelihood/prediction', 1, cfg.nlabels, normed=False, out_dtype=L.PHS_F32)]
