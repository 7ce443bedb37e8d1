"""B200-native counterpart of phiseg/phiseg_model.py::phiseg (reference lines cited per method).

Same constructor argument (an experiment module), same method names, numpy in / numpy out.  Differences, all
deliberate and listed in DESIGN.md: `training_step` and `generate_samples` exist as first-class methods (the
reference only has the loop body phiseg_model.py:193-197 and the broken :478-481); the unseeded tf.random_normal
draws can be replaced by injected `eps` so runs are reproducible; the BN moving-average double update of
phiseg_model.py:135-141 is not reproduced.
"""
import logging
import math
import os
import time
import types

import numpy as np
import torch

from .. import engine as E
from .. import metrics as M
from .. import lib as L
from .. import parallel
from ..tf_compat import optimizer_kind
from ..tfwrapper.normalisation import norm_kind

logging.basicConfig(level=logging.INFO, format='%(asctime)s %(message)s')


def find_floor_in_list(keys, value):
    """utils.py:70-84: largest key <= value (learning-rate schedule lookup, phiseg_model.py:189-190)."""
    ks = sorted(keys)
    best = None
    for k in ks:
        if k <= value:
            best = k
    if best is None:
        raise ValueError('no schedule entry at or below step %d' % value)
    return best


def net_config_from_experiment(exp, mode):
    """Reads the attributes phiseg_model.py:20-141 reads from exp_config."""
    archs = {getattr(exp.posterior, 'arch', None), getattr(exp.prior, 'arch', None),
             getattr(exp.likelihood, 'arch', None)}
    if archs == {'dummy', 'det_unet'} and getattr(exp.likelihood, 'arch', None) == 'det_unet':
        archs = {'det_unet'}            # experiments/detunet.py: posteriors.dummy, priors.dummy, likelihoods.det_unet2D
    if len(archs) != 1 or None in archs:
        raise ValueError('posterior/prior/likelihood must name one architecture, got %r' % (archs,))
    arch = archs.pop()
    if arch not in ('phiseg', 'probunet', 'det_unet'):
        raise NotImplementedError('architecture %r is not built (phiseg, prob_unet2D and det_unet2D are)' % arch)

    def opt(name):
        return getattr(exp, name) if hasattr(exp, name) and getattr(exp, name) is not None else None

    return E.NetConfig(arch=arch, image_size=tuple(exp.image_size), nlabels=exp.nlabels, zdim0=exp.zdim0, n0=exp.n0,
                       resolution_levels=exp.resolution_levels, latent_levels=exp.latent_levels,
                       norm=norm_kind(exp.layer_norm), KL_weight=opt('KL_divergence_loss_weight'),
                       xent_weight=opt('residual_multinoulli_loss_weight'),
                       exponential_weighting=getattr(exp, 'exponential_weighting', True),
                       weight_decay=opt('weight_decay_weight'), optimizer=optimizer_kind(exp.optimizer), mode=mode)


class _GraphHandle:
    """Stand-in for a TensorFlow placeholder / tensor attribute of the reference model (phiseg_model.py:26-32,85-109): the
    evaluation scripts use them as keys of feed_dict and as fetches of model.sess.run."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return '<phiseg graph handle %s>' % self.name


class _Session:
    """model.sess.run(fetches, feed_dict) for the de-facto attribute API of the reference (phiseg_test_quantitative.py:49-54,
    phiseg_makegif_samples.py:96-100): inference-time fetches of ONE prior draw per call -
    s_out_eval, s_out_eval_sm (summed level outputs / their softmax, phiseg_model.py:107-109), s_out_eval_list,
    s_out_eval_sm_list (per level, resized to the image, :89-102), prior_z_list_gen - and, when s_inp is fed, loss_tot of
    the evaluation graph.  All fetches of one call come from the same launch of the sampling program (same noise), like
    one TensorFlow session run.  Training goes through training_step / train, not through the session."""

    def __init__(self, model):
        self.model = model

    def run(self, fetches, feed_dict=None):
        m = self.model
        single = not isinstance(fetches, (list, tuple))
        fl = [fetches] if single else list(fetches)
        fd = {getattr(k, 'name', k): v for k, v in (feed_dict or {}).items()}
        if bool(fd.get('training_time', False)):
            raise NotImplementedError('sess.run with training_pl=True: the training graph runs through training_step() / train()')
        if 'x_input' not in fd:
            raise KeyError('feed_dict must feed model.x_inp')
        x = np.asarray(fd['x_input'], dtype=np.float32)
        names = [getattr(f, 'name', f) for f in fl]
        known = ('s_out_eval', 's_out_eval_sm', 's_out_eval_list', 's_out_eval_sm_list', 'prior_z_list_gen', 'loss_tot')
        for n in names:
            if n not in known:
                raise KeyError('fetch %r is not part of the session surface (%s)' % (n, ', '.join(known)))
        out = {}
        if any(n != 'loss_tot' for n in names):
            B = int(x.shape[0])
            sp = m._program('sample', B)
            m._stage_x(sp, x)
            m._sample_once(sp)
            levels = None
            for n in set(names):
                if n == 's_out_eval':
                    out[n] = m._np(sp.s_out)
                elif n == 's_out_eval_sm':
                    out[n] = m._np(sp.s_out_sm)
                elif n in ('s_out_eval_list', 's_out_eval_sm_list'):
                    levels = levels if levels is not None else m._levels_full_res(sp)
                    if n == 's_out_eval_list':
                        out[n] = levels
                    else:
                        sm = []
                        for y in levels:
                            e = np.exp(y - y.max(axis=-1, keepdims=True))
                            sm.append(e / e.sum(axis=-1, keepdims=True))
                        out[n] = sm
                elif n == 'prior_z_list_gen':
                    out[n] = [m._np(a.tensor()).reshape(s) for a, s in zip(sp.z, m.cfg.latent_shapes(B))]
        if 'loss_tot' in names:
            if 's_input' not in fd:
                raise KeyError('loss_tot needs model.s_inp in feed_dict')
            out['loss_tot'] = m.evaluate_losses(x, np.asarray(fd['s_input']))['total_loss']
        res = [out[n] for n in names]
        return res[0] if single else res


class phiseg():

    def __init__(self, exp_config, mode=None, device=None, use_cuda_graph=True, seed=1234, data_parallel=True):
        """mode: 'fast' (bf16 tcgen05 tensor-core kernels, fp32 accumulate and fp32 normalisation statistics / losses /
        optimizer), 'parity_tc' (fp32 activations; every 32-channel-aligned convolution as three bf16 tcgen05 passes over a
        (hi, lo) operand split = fp32-accurate products on the tensor cores: meets the 1e-3 logit contract) or 'parity'
        (fp32 CUDA-core kernels).  Default: exp_config.compute_mode if present, else 'fast'."""
        self.exp_config = exp_config
        if not torch.cuda.is_available():
            raise RuntimeError('phiseg-code_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        self.lib = L.load()
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        mode = mode or getattr(exp_config, 'compute_mode', 'fast')
        self.cfg = net_config_from_experiment(exp_config, mode)
        self.params = E.Params(self.cfg, self.device)
        self.params.init(seed)
        self.use_cuda_graph = use_cuda_graph
        self._progs = {}
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(seed)
        self._hyper = torch.zeros(4, dtype=torch.float32, device=self.device)
        self.loss_dict = {}
        self.loss_tot = None
        # the attribute surface the reference's evaluation scripts use (phiseg_model.py:26-32,85-109)
        self.sess = _Session(self)
        self.x_inp, self.s_inp = _GraphHandle('x_input'), _GraphHandle('s_input')
        self.training_pl, self.lr_pl = _GraphHandle('training_time'), _GraphHandle('learning_rate')
        for h in ('s_out_eval', 's_out_eval_sm', 's_out_eval_list', 's_out_eval_sm_list', 'prior_z_list_gen'):
            setattr(self, h, _GraphHandle(h))
        self.log_dir = None
        self.gpu_launches = 0
        self.world = 1
        self.rank = 0
        # 'eager': grad graph -> NCCL all-reduce -> optimizer graph; 'graph': bucketed all-reduces captured inside the step
        self.dp_mode = os.environ.get('PHS_DP_MODE', 'eager')
        # data_parallel=False: a replica that trains alone although a process group exists (no collectives at all)
        if data_parallel and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size()
            self.rank = torch.distributed.get_rank()
            if self.world > 1:
                # identical replicas: rank 0's initial weights everywhere
                torch.distributed.broadcast(self.params.p, 0)
                torch.distributed.broadcast(self.params.state, 0)
                self.params.refresh_shadow()

    # ---------------------------------------------------------------------------------------------------
    # programs
    # ---------------------------------------------------------------------------------------------------
    def _program(self, kind, B, rep=1):
        """The static launch program of one graph kind at B images (rep samples per image batched along N for sampling)."""
        key = (kind, B) if rep == 1 else (kind, B, rep)
        sp = self._progs.get(key)
        if sp is None:
            sp = E.build_program(self.cfg, self.params, B, kind, self.device, rep=rep)
            sp.graph = None
            sp.runs = 0
            sp.h_x = torch.zeros((B, self.cfg.H, self.cfg.W, self.cfg.Cx), dtype=torch.float32).pin_memory()
            sp.h_s = torch.zeros((B, self.cfg.H, self.cfg.W), dtype=torch.uint8).pin_memory()
            sp.h_losses = torch.zeros(sp.losses.numel(), dtype=torch.float32).pin_memory()
            if kind == 'train':
                self._append_optimizer(sp)
            elif kind == 'eval' and self.cfg.weight_decay is not None:
                # the reference's validation total_loss carries the weight-decay term too (phiseg_model.py:124-128,537-549)
                self._emit_weight_decay(sp.prog, sp, with_grad=False)
            self._progs[key] = sp
        return sp

    def _weight_segments(self):
        """(offset, count) of every filter ('weight_variables' collection, tfwrapper/utils.py:254-255) in the flat buffer."""
        if getattr(self, '_wd_segs', None) is None:
            rows = [[off, int(np.prod(shape))] for name, (off, shape, kind) in self.params.table.items() if kind == 'W']
            self._wd_segs = torch.tensor(rows, dtype=torch.int64, device=self.device).reshape(-1, 2)
        return self._wd_segs

    def _emit_weight_decay(self, pr, sp, with_grad):
        """add_weight_decay (phiseg_model.py:290-300) as ONE launch over all filters: loss term, and its gradient wd*W."""
        P, cfg = self.params, self.cfg
        segs = self._weight_segments()
        pr.emit('phs_weight_decay', P.p.data_ptr(), P.g.data_ptr() if with_grad else None, segs.data_ptr(), segs.shape[0],
                float(cfg.weight_decay), sp.losses.data_ptr() + 4 * (2 * cfg.L))

    def _append_optimizer(self, sp):
        """Gradient zeroing goes in front of the backward launches; weight decay, the optimizer update and the bf16
        shadow refresh go behind them (phiseg_model.py:134-141,290-300)."""
        P, pr, cfg = self.params, sp.prog, self.cfg
        P.ensure_slots()
        steps = pr.steps
        fwd, bwd = steps[:sp.n_fwd], steps[sp.n_fwd:]
        pr.steps = []
        pr.emit('phs_fill_f32', P.g.data_ptr(), P.n, 0.0)
        zero = pr.steps
        pr.steps = []
        if cfg.weight_decay is not None:
            self._emit_weight_decay(pr, sp, with_grad=True)
        wd = pr.steps
        pr.steps = []
        gs = 1.0 / self.world
        if cfg.optimizer == 'adam':
            pr.emit('phs_adam_step', P.p.data_ptr(), P.g.data_ptr(), P.slots[0].data_ptr(), P.slots[1].data_ptr(), P.n,
                    0.0, self._hyper.data_ptr(), 0.9, 0.999, 1e-8, gs)
        else:
            pr.emit('phs_momentum_step', P.p.data_ptr(), P.g.data_ptr(), P.slots[0].data_ptr(), P.n, 0.0,
                    self._hyper.data_ptr(), 0.9, gs)
        if P.shadow is not None and P.prep_table.numel():
            pr.emit('phs_weight_prep', P.p.data_ptr(), P.shadow.data_ptr(), P.prep_table.data_ptr(),
                    P.prep_table.shape[0])
            if P.shadow_lo is not None:
                pr.emit('phs_weight_prep_lo', P.p.data_ptr(), P.shadow_lo.data_ptr(), P.prep_table.data_ptr(),
                        P.prep_table.shape[0])
        opt = pr.steps
        if self.world > 1:
            # weight decay is a local, identical term on every replica: it is added AFTER the all-reduce (scaled by
            # world so that the optimizer's 1/world leaves wd*W), never summed over the replicas
            if cfg.weight_decay is not None:
                pr.steps = []
                segs = self._weight_segments()
                pr.emit('phs_weight_decay', P.p.data_ptr(), P.g.data_ptr(), segs.data_ptr(), segs.shape[0],
                        float(cfg.weight_decay) * self.world, None)
                pr.emit('phs_weight_decay', P.p.data_ptr(), None, segs.data_ptr(), segs.shape[0],
                        float(cfg.weight_decay), sp.losses.data_ptr() + 4 * (2 * cfg.L))
                wd = pr.steps
            if self.dp_mode == 'graph':
                bwd = parallel.insert_gradient_allreduce(bwd, P, self.world)
        if self.world > 1 and self.dp_mode != 'graph':
            sp.grad_steps = fwd + zero + bwd            # produces losses and the LOCAL gradient
            sp.opt_steps = wd + opt                     # after the all-reduce: weight decay, optimizer, shadow refresh
        else:
            sp.grad_steps = fwd + zero + bwd + wd       # produces losses and the (all-reduced) gradient
            sp.opt_steps = opt                          # consumes it
        pr.steps = sp.grad_steps + sp.opt_steps

    def _launch(self, sp, steps, tag):
        """Replay a launch list: eagerly the first time (loads modules, validates arguments), then from a CUDA
        graph captured on the second call."""
        pr = sp.prog
        self.gpu_launches += pr.launches(steps)
        if not self.use_cuda_graph:
            pr.run_eager(steps)
            return
        graphs = sp.__dict__.setdefault('graphs', {})
        seen = sp.__dict__.setdefault('seen', {})
        if tag in graphs:
            graphs[tag].replay()
            return
        if seen.get(tag, 0) >= 1:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                pr.run_eager(steps)
            graphs[tag] = g
            g.replay()
            return
        seen[tag] = seen.get(tag, 0) + 1
        pr.run_eager(steps)

    # ---------------------------------------------------------------------------------------------------
    # input staging
    # ---------------------------------------------------------------------------------------------------
    def _stage_x(self, sp, x_in):
        x = np.asarray(x_in)
        if x.shape != tuple(sp.h_x.shape):
            raise ValueError('x has shape %s, expected %s' % (x.shape, tuple(sp.h_x.shape)))
        np.copyto(sp.h_x.numpy(), x, casting='unsafe')        # one pass: convert + copy into the pinned buffer
        sp.x.buf.t.copy_(sp.h_x, non_blocking=True)
        return sp.h_x.numel() * 4

    def _stage_s(self, sp, s_in):
        s = np.asarray(s_in)
        if s.shape != tuple(sp.h_s.shape):
            raise ValueError('s has shape %s, expected %s' % (s.shape, tuple(sp.h_s.shape)))
        self._check_labels(s)
        np.copyto(sp.h_s.numpy(), s, casting='unsafe')
        sp.s.copy_(sp.h_s, non_blocking=True)
        return sp.h_s.numel()

    def _check_labels(self, s):
        if s.size:
            # unsigned inputs cannot be negative: one max() pass instead of min() + max()
            bad = s.max() >= self.cfg.nlabels if s.dtype.kind == 'u' else (s.min() < 0 or s.max() >= self.cfg.nlabels)
            if bad:
                raise ValueError('labels must lie in [0, %d)' % self.cfg.nlabels)

    def _draw_eps(self, sp, eps=None):
        if eps is None:
            for e in sp.eps:
                e.normal_(generator=self._gen)
        else:
            if len(eps) != len(sp.eps):
                raise ValueError('eps must have one entry per latent level (%d)' % len(sp.eps))
            for e, v in zip(sp.eps, eps):
                v = torch.as_tensor(np.asarray(v, dtype=np.float32)) if not torch.is_tensor(v) else v
                e.copy_(v.reshape(e.shape).to(self.device, torch.float32))

    # ---------------------------------------------------------------------------------------------------
    # training
    # ---------------------------------------------------------------------------------------------------
    def _lr_for_step(self, step):
        sched = self.exp_config.lr_schedule_dict
        return sched[find_floor_in_list(sched.keys(), step)]

    def _device_step(self, sp, lr):
        """loss, gradients, (all-reduce,) optimizer on inputs already resident in sp.x / sp.s / sp.eps."""
        P = self.params
        t = P.step + 1
        if self.cfg.optimizer == 'adam':
            # tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
            lr_t = lr * math.sqrt(1.0 - 0.999 ** t) / (1.0 - 0.9 ** t)
        else:
            lr_t = lr
        # the step size travels as a kernel ARGUMENT (captured by value at launch) into the device word the captured
        # optimizer node reads: a pinned staging word could be overwritten by the next step's host code before its
        # async copy ran when steps are enqueued back to back without a synchronisation
        L.check(self.lib.phs_fill_f32(self._hyper.data_ptr(), 1, float(lr_t), torch.cuda.current_stream().cuda_stream),
                'phs_fill_f32')
        if self.world > 1 and self.dp_mode != 'graph':
            # default data-parallel path: gradient graph, ONE NCCL all-reduce over the flat gradient buffer (NVLink),
            # optimizer graph
            self._launch(sp, sp.grad_steps, 'grad')
            parallel.allreduce_sum_(P.g)
            self._launch(sp, sp.opt_steps, 'opt')
        else:
            # single GPU, or PHS_DP_MODE=graph: the bucketed all-reduces are part of the program (NCCL launches on a
            # communication lane, captured into the same CUDA graph, overlapping the backward tail)
            self._launch(sp, sp.prog.steps, 'step')
        P.step = t

    def _read_losses(self, sp):
        sp.h_losses.copy_(sp.losses, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._decode_losses(sp.h_losses.numpy())

    def _decode_losses(self, v):
        """loss vector [xent_l | KL_l | weight decay] -> loss_dict / loss_tot (phiseg_model.py:113-128)"""
        cfg = self.cfg
        ld = {}
        tot = 0.0
        if cfg.xent_weight is not None:
            for l in range(cfg.L):
                ld['residual_multinoulli_loss_lvl%d' % l] = float(v[l]) / cfg.xent_weight
                tot += float(v[l])
        if cfg.KL_weight is not None:
            for l in range(cfg.L):
                ld['KL_divergence_loss_lvl%d' % l] = float(v[cfg.L + l])
                tot += cfg.KL_weight * float(v[cfg.L + l])
        if cfg.weight_decay is not None:
            ld['weight_decay'] = float(v[2 * cfg.L])
            tot += float(v[2 * cfg.L])
        ld['total_loss'] = tot
        self.loss_dict = ld
        self.loss_tot = tot
        return tot

    def _pipe(self, sp):
        """Double-buffered staging of a training program: two pinned host slots and two device staging slots for the
        batch, two pinned slots for the loss vector, a copy stream and the events that order them."""
        p = getattr(sp, 'pipe', None)
        if p is None:
            p = types.SimpleNamespace()
            p.h_x = [torch.zeros_like(sp.h_x).pin_memory() for _ in range(2)]
            p.h_s = [torch.zeros_like(sp.h_s).pin_memory() for _ in range(2)]
            p.h_losses = [torch.zeros_like(sp.h_losses).pin_memory() for _ in range(2)]
            p.d_x = [torch.empty_like(sp.x.buf.t) for _ in range(2)]
            p.d_s = [torch.empty_like(sp.s) for _ in range(2)]
            p.copy_stream = torch.cuda.Stream(device=self.device)
            p.ev_h2d = [torch.cuda.Event() for _ in range(2)]     # slot's host -> device copy finished
            p.ev_free = [torch.cuda.Event() for _ in range(2)]    # slot's device staging buffer consumed by its step
            p.ev_done = [torch.cuda.Event() for _ in range(2)]    # slot's loss vector is in pinned memory
            p.k = 0
            p.pending = None
            sp.pipe = p
        return p

    def training_step(self, x_b, s_b, lr=None, eps=None, defer=False):
        """One iteration of the hot loop (phiseg_model.py:186-197): feed a batch, run forward, ELBO, backward and the
        optimizer, return loss_tot.  x_b [B,H,W,C] float32, s_b [B,H,W] uint8: host arrays, or CUDA tensors of those
        shapes / dtypes (data.BatchProvider.next_batch_device), which skip the staging below.

        Staging is double buffered: the batch is converted into a pinned slot and copied to a device staging slot on a
        copy stream while the previous step may still be running; the step itself starts with a device-to-device copy.
        defer=False returns this step's loss (one host synchronisation per step, like sess.run).  defer=True returns
        the PREVIOUS step's loss (None on the first call) and does not wait for this one, so the host prepares batch k+1
        while the device runs step k; flush() returns the last one.  Every step's loss vector is still read back."""
        B = int(np.shape(x_b)[0])
        sp = self._program('train', B)
        if lr is None:
            lr = self._lr_for_step(self.params.step)
        p = self._pipe(sp)
        j = p.k & 1
        p.k += 1
        cur = torch.cuda.current_stream()
        if torch.is_tensor(x_b) and x_b.is_cuda:
            # a batch that is already on the device (data.BatchProvider.next_batch_device): no host staging at all
            if tuple(x_b.shape) != tuple(sp.h_x.shape) or tuple(s_b.shape) != tuple(sp.h_s.shape):
                raise ValueError('x / s have shapes %s / %s, expected %s / %s'
                                 % (tuple(x_b.shape), tuple(s_b.shape), tuple(sp.h_x.shape), tuple(sp.h_s.shape)))
            sp.x.buf.t.copy_(x_b.reshape(sp.x.buf.t.shape), non_blocking=True)
            sp.s.copy_(s_b, non_blocking=True)
            self.h2d_bytes = 0
        else:
            p.ev_h2d[j].synchronize()                # the slot's previous copy (two steps ago) has left host memory
            x = np.asarray(x_b)
            s = np.asarray(s_b)
            if x.shape != tuple(sp.h_x.shape):
                raise ValueError('x has shape %s, expected %s' % (x.shape, tuple(sp.h_x.shape)))
            if s.shape != tuple(sp.h_s.shape):
                raise ValueError('s has shape %s, expected %s' % (s.shape, tuple(sp.h_s.shape)))
            self._check_labels(s)
            np.copyto(p.h_x[j].numpy(), x, casting='unsafe')
            np.copyto(p.h_s[j].numpy(), s, casting='unsafe')
            self.h2d_bytes = p.h_x[j].numel() * 4 + p.h_s[j].numel()
            with torch.cuda.stream(p.copy_stream):
                p.copy_stream.wait_event(p.ev_free[j])
                p.d_x[j].copy_(p.h_x[j], non_blocking=True)
                p.d_s[j].copy_(p.h_s[j], non_blocking=True)
                p.ev_h2d[j].record(p.copy_stream)
            cur.wait_event(p.ev_h2d[j])
            sp.x.buf.t.copy_(p.d_x[j], non_blocking=True)
            sp.s.copy_(p.d_s[j], non_blocking=True)
            p.ev_free[j].record(cur)
        self._draw_eps(sp, eps)
        self._device_step(sp, lr)
        p.h_losses[j].copy_(sp.losses, non_blocking=True)
        p.ev_done[j].record(cur)
        self.d2h_bytes = p.h_losses[j].numel() * 4
        prev, p.pending = p.pending, j
        if defer:
            return self._finish(p, prev) if prev is not None else None
        if prev is not None:
            self._finish(p, prev)
        p.pending = None
        return self._finish(p, j)

    def _finish(self, p, j):
        p.ev_done[j].synchronize()
        return self._decode_losses(p.h_losses[j].numpy())

    def flush(self):
        """Loss of the last deferred training_step (None if there is none pending)."""
        out = None
        for sp in self._progs.values():
            p = getattr(sp, 'pipe', None)
            if p is not None and p.pending is not None:
                out = self._finish(p, p.pending)
                p.pending = None
        return out

    def train(self, data):
        """phiseg_model.py:166-207 without TensorBoard: lr schedule lookup, next_batch, training_step, periodic
        validation losses and checkpoints."""
        exp = self.exp_config
        self._setup_log_dir_and_continue_mode()
        self.best_loss = np.inf
        for step in range(self.init_step, exp.num_iter):
            lr = self._lr_for_step(step)
            # a device-resident provider hands the batch over as CUDA tensors (no host round trip)
            nb = getattr(data.train, 'next_batch_device', None) or data.train.next_batch
            x_b, s_b = nb(exp.batch_size)
            # deferred: the loss that comes back belongs to the previous step, the host never waits for the device
            loss = self.training_step(x_b, s_b, lr, defer=True)
            if step % exp.tensorboard_update_frequency == 0 and loss is not None:
                logging.info('step %d  loss %.4f  lr %g' % (step - 1, loss, lr))
            if step % exp.validation_frequency == 0:
                self.flush()
                self._do_validation(data, step)
        self.flush()

    def _do_validation(self, data, step):
        """phiseg_model.py:530-660 without TensorBoard: checkpoint, batch validation losses (training=False), then the
        full validation pass - per image validation_samples prior samples and ELBO evaluations, GED / NCC / Dice computed
        on the device (validation_metrics) - and the four best-model savers (Dice, ELBO, GED, NCC; max_to_keep=2)."""
        exp = self.exp_config
        self.save_weights(self.log_dir, 'model.ckpt-%d' % step)
        _prune_checkpoints(self.log_dir, 'model.ckpt', keep=1)             # tf.train.Saver(max_to_keep=1), :144
        if not hasattr(data, 'validation'):
            return None
        x_b, s_b = data.validation.next_batch(exp.batch_size)
        val = self.evaluate_losses(x_b, s_b)
        logging.info('validation  step %d  batch loss %.4f' % (step, val['total_loss']))
        images = getattr(data.validation, 'images', None)
        if images is None:
            if val['total_loss'] < self.best_loss:
                self.best_loss = val['total_loss']
                self._save_best('model_best_loss.ckpt', step)
            return val
        n = images.shape[0] if exp.num_validation_images == 'all' else min(int(exp.num_validation_images), images.shape[0])
        rng = np.random.default_rng(step)
        dice, elbo, ged, ncc = [], [], [], []
        t0 = time.time()
        for ii in range(n):
            x = np.asarray(images[ii]).reshape((1,) + tuple(exp.image_size))
            m = self.validation_metrics(x, data.validation.labels[ii], exp.validation_samples,
                                        annotator=int(rng.choice(list(exp.annotator_range))))
            dice.append(m['dice']); elbo.append(m['elbo']); ged.append(m['ged']); ncc.append(m['ncc'])
        per_structure = np.mean(np.asarray(dice), axis=0)
        out = {'dice': float(per_structure.mean()), 'per_structure_dice': per_structure, 'elbo': float(np.mean(elbo)),
               'ged': float(np.mean(ged)), 'ncc': float(np.nanmean(ncc)), 'images': n, 'seconds': time.time() - t0}
        logging.info('FULL VALIDATION (%d images, %.2f s): dice %.4f  ELBO %.4f  GED %.4f  NCC %.4f'
                     % (n, out['seconds'], out['dice'], out['elbo'], out['ged'], out['ncc']))
        if out['dice'] >= getattr(self, 'best_dice', -np.inf):
            self.best_dice = out['dice']
            self._save_best('model_best_dice.ckpt', step)
        if out['elbo'] <= self.best_loss:
            self.best_loss = out['elbo']
            self._save_best('model_best_loss.ckpt', step)
        if out['ged'] <= getattr(self, 'best_ged', np.inf):
            self.best_ged = out['ged']
            self._save_best('model_best_ged.ckpt', step)
        if out['ncc'] >= getattr(self, 'best_ncc', -np.inf):
            self.best_ncc = out['ncc']
            self._save_best('model_best_ncc.ckpt', step)
        self.last_validation = out
        return out

    def _save_best(self, prefix, step):
        self.save_weights(self.log_dir, '%s-%d' % (prefix, step))
        _prune_checkpoints(self.log_dir, prefix, keep=2)                  # saver_best_*: max_to_keep=2, :145-148

    def validation_metrics(self, x_img, s_gt_arr, num_samples=None, annotator=0, eps=None):
        """One image of the reference's validation loop (phiseg_model.py:567-611).  x_img [1,H,W,C] (or [H,W,C]),
        s_gt_arr [H,W,A]: the A annotations of the image.  num_samples prior samples (batched along N, x-only part once)
        and num_samples ELBO evaluations against annotation `annotator`, then on the device: pairwise label
        intersections -> generalised energy distance over the foreground labels, cross-entropy maps -> variance NCC,
        and the Dice of the mean prediction.  Returns {'ged', 'ncc', 'dice' [nlabels], 'elbo'}."""
        cfg, st = self.cfg, torch.cuda.current_stream().cuda_stream
        S = int(num_samples or self.exp_config.validation_samples)
        x = np.asarray(x_img, dtype=np.float32).reshape(1, cfg.H, cfg.W, cfg.Cx)
        gts = np.ascontiguousarray(np.moveaxis(np.asarray(s_gt_arr), -1, 0)).astype(np.uint8)      # [A,H,W]
        A, P, nl = gts.shape[0], cfg.H * cfg.W, cfg.nlabels
        s = gts[annotator]
        # ELBO with training=False on S copies of (x, s) (:574-581)
        ev = self.evaluate_losses(np.tile(x, (S, 1, 1, 1)), np.tile(s[None], (S, 1, 1)))
        # S prior samples: one pass, rows = samples
        sp = self._program('sample', 1, S)
        self._stage_x(sp, x)
        L.check(self.lib.phs_fill_f32(sp.sm_accum.data_ptr(), sp.sm_accum.numel(), 0.0, st), 'phs_fill_f32')
        self._draw_eps(sp, eps)
        self._launch(sp, sp.prog.steps, 'fwd')
        dev = self.device
        gt_d = torch.as_tensor(gts).to(dev)
        mean_arg = torch.empty((1, cfg.H, cfg.W), dtype=torch.int64, device=dev)
        L.check(self.lib.phs_argmax_f32(sp.sm_accum.data_ptr(), P, nl, mean_arg.data_ptr(), st), 'phs_argmax_f32')
        i32 = lambda *shape: torch.empty(shape, dtype=torch.int32, device=dev)
        i_sy, i_ss, i_yy, i_d = i32(S, A, nl), i32(S, S, nl), i32(A, A, nl), i32(1, 1, nl)
        c_s, c_y, c_p, c_g = i32(S, nl), i32(A, nl), i32(1, nl), i32(1, nl)
        pls = self.lib.phs_pairwise_label_stats
        am, sgt = sp.argmax, gt_d[annotator:annotator + 1].contiguous()
        L.check(pls(am.data_ptr(), 8, S, gt_d.data_ptr(), 1, A, P, nl, i_sy.data_ptr(), c_s.data_ptr(), c_y.data_ptr(), st), 'phs_pairwise_label_stats')
        L.check(pls(am.data_ptr(), 8, S, am.data_ptr(), 8, S, P, nl, i_ss.data_ptr(), None, None, st), 'phs_pairwise_label_stats')
        L.check(pls(gt_d.data_ptr(), 1, A, gt_d.data_ptr(), 1, A, P, nl, i_yy.data_ptr(), None, None, st), 'phs_pairwise_label_stats')
        L.check(pls(mean_arg.data_ptr(), 8, 1, sgt.data_ptr(), 1, 1, P, nl, i_d.data_ptr(), c_p.data_ptr(), c_g.data_ptr(), st), 'phs_pairwise_label_stats')
        e_ss = torch.empty(P, dtype=torch.float32, device=dev)
        e_sy = torch.empty((A, P), dtype=torch.float32, device=dev)
        sums = torch.empty((A, 5), dtype=torch.float64, device=dev)
        L.check(self.lib.phs_ncc_maps(sp.s_out_sm.data_ptr(), gt_d.data_ptr(), S, A, P, nl, e_ss.data_ptr(), e_sy.data_ptr(),
                                      sums.data_ptr(), st), 'phs_ncc_maps')
        self.gpu_launches += 7
        h = [t.cpu().numpy() for t in (i_sy, i_ss, i_yy, c_s, c_y, i_d, c_p, c_g, sums)]
        fg = range(1, nl)                                        # label_range=range(1, nlabels) (:586-588)
        return {'ged': M.ged_from_counts(h[0], h[1], h[2], h[3], h[4], fg), 'ncc': M.ncc_from_sums(h[8], P),
                'dice': M.dice_from_counts(h[5][0, 0], h[6][0], h[7][0]), 'elbo': ev['total_loss']}

    def evaluate_losses(self, x_b, s_b, eps=None):
        """loss_dict with training=False (phiseg_model.py:537-549)."""
        B = int(np.shape(x_b)[0])
        sp = self._program('eval', B)
        self._stage_x(sp, x_b)
        self._stage_s(sp, s_b)
        self._draw_eps(sp, eps)
        self._launch(sp, sp.prog.steps, 'fwd')
        self._read_losses(sp)
        return dict(self.loss_dict)

    # ---------------------------------------------------------------------------------------------------
    # sampling / prediction (phiseg_model.py:313-502)
    # ---------------------------------------------------------------------------------------------------
    def _sample_once(self, sp, eps=None):
        self._draw_eps(sp, eps)
        self._launch(sp, sp.prog.steps, 'fwd')

    def _np(self, t):
        return t.detach().cpu().numpy()

    sample_rows = 64      # images x samples evaluated per sampling pass (rows of the batch the kernels see)

    def _sample_plan(self, B, num_samples):
        """[(samples per pass, passes)]: as many samples of every image as fit sample_rows go through the network in one
        pass, batched along N; a remainder gets its own (smaller) program."""
        rep = max(1, min(int(num_samples), self.sample_rows // max(B, 1)))
        plan = [(rep, num_samples // rep)]
        if num_samples % rep:
            plan.append((num_samples % rep, 1))
        return [p for p in plan if p[1] > 0]

    def _run_samples(self, x_in, num_samples, visit=None):
        """num_samples prior samples of every image of x_in.  The x-only part of the graph (the prior's encoder pyramid;
        for the probabilistic U-Net the whole U-Net) runs ONCE per image, only the noise-dependent part is replayed per
        draw - the reference re-runs everything (phiseg_model.py:344-348).  Returns the device tensor holding the sum of
        the samples' softmax maps [B,H,W,nlabels]; visit(sp) is called after every pass (sp.s_out etc. hold rep samples)."""
        B = int(np.shape(x_in)[0])
        st = torch.cuda.current_stream().cuda_stream
        total = None
        for rep, passes in self._sample_plan(B, num_samples):
            sp = self._program('sample', B, rep)
            self._stage_x(sp, x_in)
            L.check(self.lib.phs_fill_f32(sp.sm_accum.data_ptr(), sp.sm_accum.numel(), 0.0, st), 'phs_fill_f32')
            steps = sp.prog.steps
            if not hasattr(sp, 'rest_steps'):
                sp.enc_steps = steps[:sp.n_enc]
                sp.rest_steps = steps[:sp.n_fills] + steps[sp.n_enc:]     # arena fills + the noise-dependent launches
            self._launch(sp, sp.enc_steps, 'enc')
            for _ in range(passes):
                self._draw_eps(sp)
                self._launch(sp, sp.rest_steps, 'rest')
                if visit is not None:
                    visit(sp)
            if total is None:
                total = sp.sm_accum
            else:
                L.check(self.lib.phs_axpy_f32(total.data_ptr(), sp.sm_accum.data_ptr(), total.numel(), 1.0, st), 'phs_axpy_f32')
            self.gpu_launches += 2
        return total

    def predict(self, x_in, num_samples=50, return_softmax=False):
        """phiseg_model.py:337-353: softmax of the summed level outputs averaged over num_samples prior draws, argmax.
        Accumulation and argmax stay on the device; one device->host copy of the mask (and of the mean softmax)."""
        B = int(np.shape(x_in)[0])
        acc = self._run_samples(x_in, num_samples)
        if getattr(self, '_argmax_img', None) is None or self._argmax_img.shape[0] != B:
            self._argmax_img = torch.empty((B, self.cfg.H, self.cfg.W), dtype=torch.int64, device=self.device)
        st = torch.cuda.current_stream().cuda_stream
        npix = B * self.cfg.H * self.cfg.W
        L.check(self.lib.phs_argmax_f32(acc.data_ptr(), npix, self.cfg.nlabels, self._argmax_img.data_ptr(), st), 'phs_argmax_f32')
        self.gpu_launches += 1
        if return_softmax:
            return self._np(self._argmax_img), self._np(acc) / num_samples
        return self._np(self._argmax_img)

    def predict_segmentation_sample(self, x_in, return_softmax=False, eps=None):
        """phiseg_model.py:356-364"""
        B = int(np.shape(x_in)[0])
        sp = self._program('sample', B)
        self._stage_x(sp, x_in)
        self._sample_once(sp, eps)
        if return_softmax:
            return self._np(sp.s_out_sm)
        return self._np(sp.argmax)

    def _levels_full_res(self, sp):
        """tf.image.resize_images(..., NEAREST_NEIGHBOR) of each level's head output (likelihoods.py:221): pure
        replication, done on the host copy."""
        out = []
        for a in sp.logits:
            y = self._np(a.tensor().float())
            f = self.cfg.H // y.shape[1]
            out.append(np.repeat(np.repeat(y, f, axis=1), f, axis=2) if f > 1 else y)
        return out

    def predict_segmentation_sample_levels(self, x_in, return_softmax=False, eps=None):
        """phiseg_model.py:367-375"""
        B = int(np.shape(x_in)[0])
        sp = self._program('sample', B)
        self._stage_x(sp, x_in)
        self._sample_once(sp, eps)
        lv = self._levels_full_res(sp)
        if return_softmax:
            res = []
            for y in lv:
                e = np.exp(y - y.max(axis=-1, keepdims=True))
                res.append(e / e.sum(axis=-1, keepdims=True))
            return res
        return lv

    def generate_prior_samples(self, x_in, return_params=False, eps=None):
        """phiseg_model.py:325-334.  Unlike the reference (three sess.run calls, three different noise draws) the
        returned mu / sigma belong to the returned z."""
        B = int(np.shape(x_in)[0])
        sp = self._program('sample', B)
        self._stage_x(sp, x_in)
        self._sample_once(sp, eps)
        shp = self.cfg.latent_shapes(B)
        z = [self._np(a.tensor()).reshape(s) for a, s in zip(sp.z, shp)]
        if return_params:
            mu = [self._np(a.tensor()).reshape(s) for a, s in zip(sp.prior_mu, shp)]
            sg = [self._np(a.tensor()).reshape(s) for a, s in zip(sp.prior_sigma, shp)]
            return z, mu, sg
        return z

    def generate_posterior_samples(self, x_in, s_in, return_params=False, eps=None):
        """phiseg_model.py:484-495"""
        B = int(np.shape(x_in)[0])
        sp = self._program('posterior', B)
        self._stage_x(sp, x_in)
        self._stage_s(sp, s_in)
        self._sample_once(sp, eps)
        shp = self.cfg.latent_shapes(B)
        z = [self._np(a.tensor()).reshape(s) for a, s in zip(sp.z, shp)]
        if return_params:
            mu = [self._np(a.tensor()).reshape(s) for a, s in zip(sp.mu, shp)]
            sg = [self._np(a.tensor()).reshape(s) for a, s in zip(sp.sigma, shp)]
            return z, mu, sg
        return z

    def generate_samples_from_z(self, z_list, x_in, output_all_levels=False):
        """phiseg_model.py:313-322: decode given latents with training=False."""
        B = int(np.shape(x_in)[0])
        sp = self._program('from_z', B)
        self._stage_x(sp, x_in)
        if len(z_list) != len(sp.z):
            raise ValueError('z_list must have %d entries' % len(sp.z))
        for a, z in zip(sp.z, z_list):
            zt = torch.as_tensor(np.asarray(z, dtype=np.float32)).reshape(a.buf.t.shape)
            a.buf.t.copy_(zt.to(self.device))
        self._launch(sp, sp.prog.steps, 'fwd')
        if output_all_levels:
            return self._levels_full_res(sp)
        return self._np(sp.s_out)

    def generate_samples_from_prior(self, x_in, output_all_levels=False):
        """Intent of phiseg_model.py:478-481 (the reference passes output_all_levels as x_in by mistake)."""
        z = self.generate_prior_samples(x_in)
        return self.generate_samples_from_z(z, x_in, output_all_levels)

    def generate_samples(self, x_in, num_samples, output_all_levels=False):
        """num_samples segmentation-logit samples per image: [num_samples, B, H, W, nlabels]
        (or a list over levels of such arrays)."""
        if output_all_levels:
            outs = [self.generate_samples_from_prior(x_in, True) for _ in range(num_samples)]
            return [np.stack([o[l] for o in outs]) for l in range(len(outs[0]))]
        B = int(np.shape(x_in)[0])
        outs = []
        # rows are sample-major, so a pass's s_out [rep*B,H,W,nl] is rep stacked samples as it stands
        self._run_samples(x_in, num_samples, visit=lambda sp: outs.append(
            self._np(sp.s_out).reshape(sp.rep, B, self.cfg.H, self.cfg.W, self.cfg.nlabels)))
        return np.concatenate(outs, axis=0)

    def checks(self):
        """phiseg_model.py:160-164: a no-op in the reference too (its only check is commented out)."""
        pass

    def generate_all_output_levels(self, x_in):
        """phiseg_model.py:498-502: s_out_list of one prior sample with training=False - the per-level head outputs,
        nearest-neighbour resized to the image size (likelihoods.py:221)."""
        return self.predict_segmentation_sample_levels(x_in, return_softmax=False)

    def _sample_maps(self, x_in, num_samples, s_gt=None, kind=0, drop_last=0, want=('mean_arg',)):
        """Per-pixel moments of num_samples prior samples, accumulated on the device (phs_sample_moments after every
        batched sampling pass: class sums, class cross products, cross entropy against s_gt) and turned into the requested
        maps by phs_sample_maps.  kind 0: moments of softmax(s_out); kind 1: of s_out clipped to [1e-5, 1-1e-5]
        (phiseg_model.py:390).  Returns {name: numpy [B,H,W]}; only the maps leave the device."""
        cfg, st = self.cfg, torch.cuda.current_stream().cuda_stream
        B, P, nl = int(np.shape(x_in)[0]), cfg.H * cfg.W, cfg.nlabels
        na = nl + nl * (nl + 1) // 2 + 1
        acc = torch.zeros(B * P * na, dtype=torch.float64, device=self.device)
        gt = None
        if s_gt is not None:
            s = np.asarray(s_gt).reshape(B, cfg.H, cfg.W)
            self._check_labels(s)
            gt = torch.as_tensor(np.ascontiguousarray(s.astype(np.uint8))).to(self.device)

        def visit(sp):
            L.check(self.lib.phs_sample_moments(sp.s_out.data_ptr(), gt.data_ptr() if gt is not None else None, sp.rep, B, P,
                                                nl, kind, 1e-5, 1.0 - 1e-5, acc.data_ptr(), st), 'phs_sample_moments')
            self.gpu_launches += 1
        self._run_samples(x_in, num_samples, visit=visit)
        out = {k: torch.empty((B, cfg.H, cfg.W), dtype=torch.int64 if k == 'mean_arg' else torch.float32, device=self.device)
               for k in want}
        ptr = lambda k: out[k].data_ptr() if k in out else None
        L.check(self.lib.phs_sample_maps(acc.data_ptr(), B * P, nl, int(num_samples), drop_last, ptr('mean_arg'),
                                         ptr('std_mean'), ptr('var_sum'), ptr('cov_det'), ptr('err'), st), 'phs_sample_maps')
        self.gpu_launches += 1
        return {k: self._np(v) for k, v in out.items()}

    def predict_segmentation_sample_variance_sm_cov(self, x_in, num_samples):
        """phiseg_model.py:378-403: per-pixel sum of the eigenvalues of the (biased) sample covariance of s_out_eval over
        all classes but the last, clipped to [1e-5, 1-1e-5].  The eigenvalue sum of a symmetric matrix is its trace, so the
        map is the sum of the per-class population variances; like the reference it expects a single image."""
        return self._sample_maps(x_in, num_samples, kind=1, drop_last=1, want=('var_sum',))['var_sum'][0].astype(np.float64)

    def predict_segmentation_sample_variance_sm_cov_bf(self, x_in, num_samples):
        """phiseg_model.py:406-430: per-pixel determinant of np.cov of the softmax samples (a python loop over the pixels
        in the reference; one thread per pixel here)."""
        return self._sample_maps(x_in, num_samples, kind=0, want=('cov_det',))['cov_det'][0].astype(np.float64)

    def get_crossentropy_error_map(self, s_gt, x_in, num_samples=100):
        """phiseg_model.py:433-446: mean over samples of eval_xent, the per-pixel cross entropy of s_out_eval: [B,H,W]."""
        return self._sample_maps(x_in, num_samples, s_gt=s_gt, want=('err',))['err']

    def predict_mean_variance_and_error_maps(self, s_gt, x_in, num_samples):
        """phiseg_model.py:449-475: argmax of the mean softmax, class-mean of the per-class standard deviation, mean cross
        entropy - each [H,W] for the single image the reference squeezes to."""
        m = self._sample_maps(x_in, num_samples, s_gt=s_gt, want=('mean_arg', 'std_mean', 'err'))
        return np.squeeze(m['mean_arg']), np.squeeze(m['std_mean']), np.squeeze(m['err'])

    # ---------------------------------------------------------------------------------------------------
    # weights / checkpoints (phiseg_model.py:144-148,505-525,821-845)
    # ---------------------------------------------------------------------------------------------------
    def get_weights(self):
        """{TF variable name: numpy array} (names as produced by the reference's variable scopes)."""
        return {k: v.numpy() for k, v in self.params.state_dict().items()}

    def set_weights(self, weights, strict=True):
        self.params.load_state_dict(weights, strict)

    def save_weights(self, log_dir, name='model.ckpt'):
        os.makedirs(log_dir, exist_ok=True)
        path = os.path.join(log_dir, name + '.npz')
        extra = {'__global_step__': np.asarray(self.params.step)}
        if self.params.slots is not None:
            for i, sl in enumerate(self.params.slots):
                extra['__slot%d__' % i] = sl.detach().cpu().numpy()
        # written under a temporary name and renamed into place: a crash mid-write never leaves a truncated 'latest'
        tmp = path + '.tmp%d.npz' % os.getpid()
        np.savez(tmp, **self.get_weights(), **extra)
        os.replace(tmp, path)
        return path

    def load_weights(self, log_dir=None, type='latest', **kwargs):
        """phiseg_model.py:505-525"""
        if not log_dir:
            log_dir = self.log_dir
        prefix = {'latest': 'model.ckpt', 'best_dice': 'model_best_dice.ckpt', 'best_loss': 'model_best_loss.ckpt',
                  'best_ged': 'model_best_ged.ckpt', 'best_ncc': 'model_best_ncc.ckpt'}
        if type == 'iter':
            assert 'iteration' in kwargs, "argument 'iteration' must be provided for type='iter'"
            path = os.path.join(log_dir, 'model.ckpt-%d.npz' % kwargs['iteration'])
            candidates = [path]
        elif type in prefix:
            candidates = _checkpoints(log_dir, prefix[type])            # newest first
        else:
            raise ValueError('Argument type=%s is unknown. type can be latest/iter.' % type)
        candidates = [p for p in candidates if p is not None and os.path.exists(p)]
        if not candidates:
            raise FileNotFoundError('no checkpoint of type %s in %s' % (type, log_dir))
        data = path = None
        for cand in candidates:
            if cand.endswith('.index'):
                # a TensorFlow-1 bundle written by the reference's tf.train.Saver (model.ckpt-N.index / .data-*)
                from ..tfwrapper import checkpoint as tfck
                try:
                    sd = tfck.read_bundle(cand[:-len('.index')])
                except Exception as e:       # noqa: BLE001
                    logging.warning('checkpoint %s is unreadable (%s); trying an older one' % (cand, e))
                    continue
                self.params.load_state_dict({k: v for k, v in sd.items() if self.params.has(k)}, strict=False)
                missing = [n for n in self.params.names() if n not in sd]
                if missing:
                    logging.warning('%d variables are not in %s (first: %s)' % (len(missing), cand, missing[0]))
                if 'global_step' in sd:
                    self.params.step = int(np.asarray(sd['global_step']).reshape(-1)[0])
                else:
                    tail = cand[:-len('.index')].rsplit('-', 1)[-1]
                    self.params.step = int(tail) if tail.isdigit() else 0
                self.params.slots = None          # (the Adam slots of a TF checkpoint live under other names; start fresh)
                return cand
            # an unreadable newest file (e.g. written by a run that died before this code wrote atomically) falls back to
            # the next older one instead of breaking the resume
            try:
                data = np.load(cand)
                data.files
                path = cand
                break
            except Exception as e:       # noqa: BLE001 - zipfile / pickle / OSError, depending on where the file is cut
                logging.warning('checkpoint %s is unreadable (%s); trying an older one' % (cand, e))
        if data is None:
            raise IOError('no readable checkpoint of type %s in %s' % (type, log_dir))
        self.params.load_state_dict({k: data[k] for k in data.files if not k.startswith('__')})
        if '__global_step__' in data.files:
            self.params.step = int(data['__global_step__'])
        if '__slot0__' in data.files:
            self.params.ensure_slots()
            for i, sl in enumerate(self.params.slots):
                sl.copy_(torch.as_tensor(data['__slot%d__' % i]).to(self.device))
        return path

    def _setup_log_dir_and_continue_mode(self):
        """phiseg_model.py:821-845: resume from <log_dir> if it holds a checkpoint."""
        root = getattr(self.exp_config, 'log_root', os.environ.get('PHISEG_LOG_ROOT', './logs'))
        self.log_dir = os.path.join(root, self.exp_config.log_dir_name, self.exp_config.experiment_name)
        os.makedirs(self.log_dir, exist_ok=True)
        self.init_step = 0
        self.continue_run = False
        path = _latest_checkpoint(self.log_dir, 'model.ckpt')
        if path is not None:
            self.load_weights(self.log_dir, 'latest')
            self.init_step = self.params.step
            self.continue_run = True
            logging.info('continuing from %s (step %d)' % (path, self.init_step))
            if getattr(self.exp_config, 'continue_in_new_dir', False):
                # the reference writes the continued run next to the old one (log_dir += '_cont', phiseg_model.py:836);
                # opt-in here because a second restart would then look in '<log_dir>' again and restart from the old step
                self.log_dir += '_cont'
                os.makedirs(self.log_dir, exist_ok=True)


def _checkpoints(log_dir, prefix):
    """All checkpoints '<prefix>[-<iteration>].npz' in log_dir, highest iteration first."""
    found = []
    if not os.path.isdir(log_dir):
        return found
    for f in os.listdir(log_dir):
        for ext in ('.npz', '.index'):          # our own format, or a TensorFlow-1 bundle of the reference
            if f.startswith(prefix) and f.endswith(ext) and '.tmp' not in f:
                mid = f[len(prefix):-len(ext)]
                if mid == '':
                    found.append((0, ext == '.npz', os.path.join(log_dir, f)))
                elif mid.startswith('-') and mid[1:].isdigit():
                    found.append((int(mid[1:]), ext == '.npz', os.path.join(log_dir, f)))
    return [p for _, _, p in sorted(found, reverse=True)]


def _latest_checkpoint(log_dir, prefix):
    """tfwrapper/utils.py:189-210: highest-iteration checkpoint with the given prefix."""
    c = _checkpoints(log_dir, prefix)
    return c[0] if c else None


def _prune_checkpoints(log_dir, prefix, keep):
    """tf.train.Saver(max_to_keep=keep) (phiseg_model.py:144-148): delete all but the `keep` newest checkpoints."""
    for p in _checkpoints(log_dir, prefix)[keep:]:
        try:
            os.remove(p)
        except OSError:
            pass
