"""CPU oracle for the PHiSeg hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain PyTorch-CPU restatement (fp32 by default, fp64 switch) of the graph the
reference builds with TensorFlow 1.12.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this file;
the product package (``phiseg-code_b200/``) never does.

PARITY UNPINNED: the reference ships no tests, golden vectors, seeds or
checkpoints, and TensorFlow 1.12 cannot be installed here (Python 3.12, no
network), so nothing produced by the reference itself pins this oracle.  Its
correctness is established by the hand-computed known-answer tests in
``tests/test_oracle_kats.py`` (legacy bilinear resize, nearest resize, BN/GN,
KL closed form, softmax-xent, TF-form Adam) and fp64 cross-checks.

Every function cites the reference lines (relative to /root/reference) it follows.
Layout conventions are the reference's: activations NHWC, filters HWIO, labels
uint8 [N,H,W]; parameters are keyed by the TF variable names the reference's
``tf.variable_scope`` nesting produces (e.g. ``posterior/z0_pre_1/W``).
The unseeded ``tf.random_normal`` draws (posteriors.py:108,128) are replaced by
*injected* eps tensors so that oracle and CUDA path can be fed identical noise.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3      # tfwrapper/normalisation.py:157
BN_DECAY = 0.99    # tfwrapper/normalisation.py:145 (moving_average_decay)
GN_EPS = 1e-5      # tfwrapper/normalisation.py:17


# ----------------------------------------------------------------------------
# op layer (tfwrapper/layers.py, tfwrapper/normalisation.py)
# ----------------------------------------------------------------------------
def conv2d_same(x, W, b=None):
    """tf.nn.conv2d(x, W, [1,1,1,1], 'SAME') + bias_add  (layers.py:123,132). x NHWC, W HWIO."""
    kh, kw = W.shape[0], W.shape[1]
    y = F.conv2d(x.permute(0, 3, 1, 2), W.permute(3, 2, 0, 1), bias=b, padding=(kh // 2, kw // 2))
    return y.permute(0, 2, 3, 1)


def averagepool2d(x):
    """tf.nn.avg_pool 2x2 stride 2 SAME on even sizes = mean of 4 (layers.py:44-54)."""
    return F.avg_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)


def _up2_axis(x, axis):
    n = x.shape[axis]
    idx_next = torch.clamp(torch.arange(n) + 1, max=n - 1)
    odd = 0.5 * (x + x.index_select(axis, idx_next))
    st = torch.stack([x, odd], dim=axis + 1)
    shape = list(x.shape)
    shape[axis] = 2 * n
    return st.reshape(shape)


def bilinear_upsample2d(x):
    """tf.image.resize_images(x, [2H,2W]) -- TF1 legacy bilinear, align_corners=False, no
    half-pixel centres (layers.py:336-345): src = dst*0.5, so out[2k]=in[k],
    out[2k+1]=0.5*(in[k]+in[min(k+1,n-1)]), separable."""
    return _up2_axis(_up2_axis(x, 1), 2)


def nearest_upsample(x, factor):
    """tf.image.resize_images(..., NEAREST_NEIGHBOR) legacy: out[y,x]=in[y//f, x//f] (likelihoods.py:221)."""
    if factor == 1:
        return x
    return x.repeat_interleave(factor, dim=1).repeat_interleave(factor, dim=2)


def global_averagepool2d(x):
    """tf.reduce_mean(x, axis=(1,2)) (layers.py:70-78)."""
    return x.mean(dim=(1, 2))


def batch_norm(x, P, scope, training, new_stats=None):
    """tf.contrib.layers.batch_norm(decay=.99, epsilon=1e-3, center, scale) (normalisation.py:145-163).
    Training: per-channel batch mean / biased variance; the moving variance is updated with the
    Bessel-corrected batch variance (TF FusedBatchNorm).  Inference: moving statistics."""
    g = P[scope + '/BatchNorm/gamma']
    b = P[scope + '/BatchNorm/beta']
    if training:
        mean = x.mean(dim=(0, 1, 2))
        var = x.var(dim=(0, 1, 2), unbiased=False)
        if new_stats is not None:
            n = x.shape[0] * x.shape[1] * x.shape[2]
            unb = var * (n / max(n - 1, 1))
            mm = P[scope + '/BatchNorm/moving_mean']
            mv = P[scope + '/BatchNorm/moving_variance']
            new_stats[scope + '/BatchNorm/moving_mean'] = (BN_DECAY * mm + (1 - BN_DECAY) * mean).detach()
            new_stats[scope + '/BatchNorm/moving_variance'] = (BN_DECAY * mv + (1 - BN_DECAY) * unb).detach()
    else:
        mean = P[scope + '/BatchNorm/moving_mean']
        var = P[scope + '/BatchNorm/moving_variance']
    return (x - mean) * torch.rsqrt(var + BN_EPS) * g + b


def group_norm2d(x, P, scope):
    """normalisation.py:17-36: G = max(2, C//16) groups of contiguous channels, moments over (H,W,C/G),
    eps 1e-5, gamma/beta of shape [1,1,1,C]."""
    N, H, W, C = x.shape
    G = max(2, C // 16)
    xg = x.reshape(N, H, W, G, C // G)
    mean = xg.mean(dim=(1, 2, 4), keepdim=True)
    var = xg.var(dim=(1, 2, 4), unbiased=False, keepdim=True)
    xg = (xg - mean) / torch.sqrt(var + GN_EPS)
    return xg.reshape(N, H, W, C) * P[scope + '/gamma'].reshape(1, 1, 1, C) + P[scope + '/beta'].reshape(1, 1, 1, C)


# ----------------------------------------------------------------------------
# parameter specification / initialisation (tfwrapper/utils.py:214-271)
# ----------------------------------------------------------------------------
def _nc(n0):
    return [n0, 2 * n0, 4 * n0, 6 * n0, 6 * n0, 6 * n0, 6 * n0]


class ParamSpec:
    """Ordered list of (name, shape, kind); kind in W,b,gamma,beta,moving_mean,moving_variance."""

    def __init__(self):
        self.entries = []

    def conv(self, scope, k, cin, cout, norm, normed, force_bias=None):
        self.entries.append((scope + '/W', (k, k, cin, cout), 'W'))
        has_bias = True
        if normed and norm == 'batch_norm':
            has_bias = False          # layers.py:126-128
        if force_bias is not None:
            has_bias = force_bias
        if has_bias:
            self.entries.append((scope + '/b', (cout,), 'b'))
        if normed:
            if norm == 'batch_norm':
                for nm, kind in (('beta', 'beta'), ('gamma', 'gamma'), ('moving_mean', 'moving_mean'),
                                 ('moving_variance', 'moving_variance')):
                    self.entries.append((scope + '/batch_norm/BatchNorm/' + nm, (cout,), kind))
            elif norm == 'group_norm':
                self.entries.append((scope + '/group_norm/gamma', (1, 1, 1, cout), 'gamma'))
                self.entries.append((scope + '/group_norm/beta', (1, 1, 1, cout), 'beta'))


def he_normal(rng, shape):
    """variance_scaling_initializer(factor=2, FAN_IN, uniform=False): truncated normal (+-2 sigma,
    resampled) with sigma = sqrt(1.3*2/fan_in), fan_in = kh*kw*Cin (tfwrapper/utils.py:225-226)."""
    fan_in = shape[0] * shape[1] * shape[2]
    std = math.sqrt(1.3 * 2.0 / fan_in)
    w = rng.standard_normal(shape)
    bad = np.abs(w) > 2.0
    while bad.any():
        w[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(w) > 2.0
    return (w * std).astype(np.float64)


# ----------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------
class Oracle:
    """Restatement of phiseg/phiseg_model.py:20-141 for the 'phiseg', 'probunet' and 'det_unet' architectures."""

    def __init__(self, arch='phiseg', image_size=(128, 128, 1), nlabels=2, zdim0=2, n0=32,
                 resolution_levels=7, latent_levels=5, norm='batch_norm',
                 KL_weight=1.0, xent_weight=1.0, exponential_weighting=True, weight_decay=None,
                 dtype=torch.float32):
        self.arch = arch
        self.H, self.W, self.Cx = image_size
        self.nlabels = nlabels
        self.zdim0 = zdim0
        self.n0 = n0
        self.R = resolution_levels
        self.L = latent_levels
        self.norm = norm
        self.KL_weight = KL_weight
        self.xent_weight = xent_weight
        self.exponential_weighting = exponential_weighting
        self.weight_decay = weight_decay
        self.dtype = dtype
        self.nc = _nc(n0)
        self.spec = self._build_spec()
        self.P = None
        self.adam = None
        self.step = 0

    # ---- parameter spec in graph-construction order ------------------------------------------
    def _build_spec(self):
        s = ParamSpec()
        nrm, nc, R, L = self.norm, self.nc, self.R, self.L
        z0 = self.zdim0
        if self.arch == 'phiseg':
            for net, cin0 in (('posterior', self.Cx + self.nlabels), ('prior', self.Cx)):
                for i in range(R):
                    cin = cin0 if i == 0 else nc[i - 1]
                    s.conv('%s/z%d_pre_1' % (net, i), 3, cin, nc[i], nrm, True)
                    s.conv('%s/z%d_pre_2' % (net, i), 3, nc[i], nc[i], nrm, True)
                    s.conv('%s/z%d_pre_3' % (net, i), 3, nc[i], nc[i], nrm, True)
                for i in reversed(range(L)):
                    if i == L - 1:
                        s.conv('%s/z%d_mu' % (net, i), 3, nc[i + R - L], z0, nrm, False)
                        s.conv('%s/z%d_sigma' % (net, i), 1, nc[i + R - L], z0, nrm, False)
                    else:
                        for j in reversed(range(0, i + 1)):
                            cin = z0 if j == i else z0 * self.n0
                            s.conv('%s/z%d_ups_to_%d_c_1' % (net, i + 1, j + 1), 3, cin, z0 * self.n0, nrm, True)
                            s.conv('%s/z%d_ups_to_%d_c_2' % (net, i + 1, j + 1), 3, z0 * self.n0, z0 * self.n0, nrm, True)
                        s.conv('%s/z%d_input_1' % (net, i), 3, nc[i + R - L] + z0 * self.n0, nc[i], nrm, True)
                        s.conv('%s/z%d_input_2' % (net, i), 3, nc[i], nc[i], nrm, True)
                        s.conv('%s/z%d_mu' % (net, i), 1, nc[i], z0, nrm, False)
                        s.conv('%s/z%d_sigma' % (net, i), 1, nc[i], z0, nrm, False)
            net = 'likelihood'
            for i in range(L):
                s.conv('%s/z%d_post_1' % (net, i), 3, z0, nc[i], nrm, True)
                s.conv('%s/z%d_post_2' % (net, i), 3, nc[i], nc[i], nrm, True)
                for t in range(R - L):
                    s.conv('%s/preups_%d/z%d_post' % (net, i, t), 3, nc[i], nc[i], nrm, True)
            for i in reversed(range(L - 1)):
                s.conv('%s/post_z%d_ups_c' % (net, i + 1), 3, nc[i + 1 + R - L] if i + 1 < L - 1 else nc[L - 1], nc[i], nrm, True)
                s.conv('%s/post_c_%d_1' % (net, i), 3, 2 * nc[i], nc[i + R - L], nrm, True)
                s.conv('%s/post_c_%d_2' % (net, i), 3, nc[i + R - L], nc[i + R - L], nrm, True)
            for i in range(L):
                cin = nc[i + R - L] if i < L - 1 else nc[L - 1]
                s.conv('%s/y_lvl%d' % (net, i), 1, cin, self.nlabels, nrm, False)
        elif self.arch in ('probunet', 'det_unet'):
            # prob_unet2D passes add_bias explicitly: False under batch_norm, True otherwise
            # (posteriors.py:25, priors.py:22, likelihoods.py:103)
            # det_unet2D (likelihoods.py:10-79, with posteriors.dummy / priors.dummy): the same U-Net, no encoders, no z
            det = self.arch == 'det_unet'
            for net, cin0 in (() if det else (('posterior', self.Cx + self.nlabels), ('prior', self.Cx))):
                for i in range(R):
                    for t in (1, 2, 3):
                        cin = (cin0 if i == 0 else nc[i - 1]) if t == 1 else nc[i]
                        s.conv('%s/conv_%d_%d' % (net, i, t), 3, cin, nc[i], nrm, True)
                s.conv('%s/pre_mu' % net, 1, nc[R - 1], z0, nrm, False)
                s.conv('%s/pre_sigma' % net, 1, nc[R - 1], z0, nrm, False)
            net = 'likelihood'
            for i in range(R):
                for t in (1, 2, 3):
                    cin = (self.Cx if i == 0 else nc[i - 1]) if t == 1 else nc[i]
                    s.conv('%s/encoder/conv_%d_%d' % (net, i, t), 3, cin, nc[i], nrm, True)
            prev = nc[R - 1]
            for jj in range(R - 1):
                ii = R - jj - 1
                s.conv('%s/decoder/conv_%d_1' % (net, jj), 3, prev + nc[ii - 1], nc[ii], nrm, True)
                s.conv('%s/decoder/conv_%d_2' % (net, jj), 3, nc[ii], nc[ii], nrm, True)
                s.conv('%s/decoder/conv_%d_3' % (net, jj), 3, nc[ii], nc[ii], nrm, True)
                prev = nc[ii]
            s.conv('%s/recomb_0' % net, 1, prev + (0 if det else z0), nc[0], nrm, True)
            s.conv('%s/recomb_1' % net, 1, nc[0], nc[0], nrm, True)
            s.conv('%s/recomb_2' % net, 1, nc[0], nc[0], nrm, True)
            s.conv('%s/prediction' % net, 1, nc[0], self.nlabels, nrm, False)
        else:
            raise ValueError('unknown arch %s' % self.arch)
        return s

    def init_params(self, seed=1234):
        """he_normal W, zero b, gamma=1, beta=0, moving_mean=0, moving_variance=1 (R2, R3)."""
        rng = np.random.default_rng(seed)
        P = {}
        for name, shape, kind in self.spec.entries:
            if kind == 'W':
                a = he_normal(rng, shape)
            elif kind in ('gamma', 'moving_variance'):
                a = np.ones(shape)
            else:
                a = np.zeros(shape)
            P[name] = torch.tensor(a, dtype=self.dtype)
        self.set_params(P)
        return P

    def set_params(self, P):
        self.P = {k: v.detach().clone().to(self.dtype) for k, v in P.items()}
        self.adam = None
        self.step = 0

    def trainable_names(self):
        return [n for n, _, k in self.spec.entries if k not in ('moving_mean', 'moving_variance')]

    # ---- layers.conv2D (layers.py:94-145) ----------------------------------------------------
    def _conv(self, x, scope, training, normed=True, act='relu', new_stats=None):
        P = self.P
        W = P[scope + '/W']
        b = P.get(scope + '/b')
        y = conv2d_same(x, W, b)
        if normed:
            if self.norm == 'batch_norm':
                y = batch_norm(y, P, scope + '/batch_norm', training, new_stats)
            elif self.norm == 'group_norm':
                y = group_norm2d(y, P, scope + '/group_norm')
        if act == 'relu':
            y = F.relu(y)
        elif act == 'softplus':
            y = F.softplus(y)
        return y

    # ---- posteriors.phiseg / priors.phiseg (posteriors.py:56-132, priors.py:51-128) -----------
    def _latent_net(self, net, inp, eps, training, z_feed=None, generation_mode=True, new_stats=None):
        R, L = self.R, self.L
        conv = lambda x, name, **kw: self._conv(x, net + '/' + name, training, new_stats=new_stats, **kw)
        pre_z = [None] * R
        for i in range(R):
            h = inp if i == 0 else averagepool2d(pre_z[i - 1])
            h = conv(h, 'z%d_pre_1' % i)
            h = conv(h, 'z%d_pre_2' % i)
            h = conv(h, 'z%d_pre_3' % i)
            pre_z[i] = h
        mu, sigma, z = [None] * L, [None] * L, [None] * L
        carry = None                      # z_ups_mat[i][i]
        for i in reversed(range(L)):
            if i == L - 1:
                mu[i] = conv(pre_z[i + R - L], 'z%d_mu' % i, normed=False, act=None)
                sigma[i] = conv(pre_z[i + R - L], 'z%d_sigma' % i, normed=False, act='softplus')
            else:
                u = bilinear_upsample2d(carry)
                u = conv(u, 'z%d_ups_to_%d_c_1' % (i + 1, i + 1))
                u = conv(u, 'z%d_ups_to_%d_c_2' % (i + 1, i + 1))
                # the j<i branches of posteriors.py:112-118 feed nothing (dead); not evaluated here
                zin = torch.cat([pre_z[i + R - L], u], dim=3)
                zin = conv(zin, 'z%d_input_1' % i)
                zin = conv(zin, 'z%d_input_2' % i)
                mu[i] = conv(zin, 'z%d_mu' % i, normed=False, act=None)
                sigma[i] = conv(zin, 'z%d_sigma' % i, normed=False, act='softplus')
            z[i] = mu[i] + sigma[i] * eps[i]
            carry = z[i] if (z_feed is None or generation_mode) else z_feed[i]   # priors.py:122-126
        return z, mu, sigma

    def _probunet_latent(self, net, inp, eps, training, new_stats=None):
        """posteriors.prob_unet2D / priors.prob_unet2D (posteriors.py:9-52, priors.py:8-48)."""
        conv = lambda x, name, **kw: self._conv(x, net + '/' + name, training, new_stats=new_stats, **kw)
        h = inp
        for i in range(self.R):
            if i > 0:
                h = averagepool2d(h)
            for t in (1, 2, 3):
                h = conv(h, 'conv_%d_%d' % (i, t))
        mu = global_averagepool2d(conv(h, 'pre_mu', normed=False, act=None))
        sigma = global_averagepool2d(conv(h, 'pre_sigma', normed=False, act='softplus'))
        z = mu + sigma * eps[0]
        return [z], [mu], [sigma]

    def posterior(self, x, s_oh, eps, training, new_stats=None):
        inp = torch.cat([x, s_oh - 0.5], dim=-1)       # posteriors.py:87
        if self.arch == 'phiseg':
            return self._latent_net('posterior', inp, eps, training, new_stats=new_stats)
        return self._probunet_latent('posterior', inp, eps, training, new_stats)

    def prior(self, z_list, x, eps, generation_mode, training, new_stats=None):
        if self.arch == 'phiseg':
            return self._latent_net('prior', x, eps, training, z_feed=z_list,
                                    generation_mode=generation_mode, new_stats=new_stats)
        return self._probunet_latent('prior', x, eps, training, new_stats)

    # ---- likelihoods (likelihoods.py:81-223) -------------------------------------------------
    def likelihood(self, z_list, x, training, new_stats=None):
        net = 'likelihood'
        conv = lambda h, name, **kw: self._conv(h, net + '/' + name, training, new_stats=new_stats, **kw)
        R, L = self.R, self.L
        if self.arch == 'phiseg':
            post_z, post_c, s = [None] * L, [None] * L, [None] * L
            for i in range(L):
                h = conv(z_list[i], 'z%d_post_1' % i)
                h = conv(h, 'z%d_post_2' % i)
                for t in range(R - L):
                    h = conv(bilinear_upsample2d(h), 'preups_%d/z%d_post' % (i, t))
                post_z[i] = h
            post_c[L - 1] = post_z[L - 1]
            for i in reversed(range(L - 1)):
                u = conv(bilinear_upsample2d(post_c[i + 1]), 'post_z%d_ups_c' % (i + 1))
                h = torch.cat([post_z[i], u], dim=3)
                h = conv(h, 'post_c_%d_1' % i)
                h = conv(h, 'post_c_%d_2' % i)
                post_c[i] = h
            for i in range(L):
                y = conv(post_c[i], 'y_lvl%d' % i, normed=False, act=None)
                s[i] = nearest_upsample(y, self.H // y.shape[1])
            return s
        # prob_unet2D
        enc = []
        h = x
        for i in range(R):
            if i > 0:
                h = averagepool2d(h)
            for t in (1, 2, 3):
                h = conv(h, 'encoder/conv_%d_%d' % (i, t))
            enc.append(h)
        for jj in range(R - 1):
            ii = R - jj - 1
            h = torch.cat([bilinear_upsample2d(h), enc[ii - 1]], dim=3)   # crop_and_concat, same sizes
            for t in (1, 2, 3):
                h = conv(h, 'decoder/conv_%d_%d' % (jj, t))
        if self.arch != 'det_unet':                    # likelihoods.py:147-151 (det_unet2D: :69 feeds the decoder output)
            z = z_list[0]
            bz = z.reshape(z.shape[0], 1, 1, z.shape[1]).expand(-1, self.H, self.W, -1)
            h = torch.cat([h, bz], dim=-1)
        for t in range(3):
            # recomb_* are 1x1 convs
            h = conv(h, 'recomb_%d' % t)
        return [conv(h, 'prediction', normed=False, act=None)]

    # ---- losses (phiseg_model.py:210-311) ----------------------------------------------------
    @staticmethod
    def KL_two_gauss_with_diag_cov(mu0, sigma0, mu1, sigma1):
        B = mu0.shape[0]
        s0 = sigma0.reshape(B, -1) ** 2
        s1 = sigma1.reshape(B, -1) ** 2
        l0 = torch.log(s0 + 1e-10)
        l1 = torch.log(s1 + 1e-10)
        m0 = mu0.reshape(B, -1)
        m1 = mu1.reshape(B, -1)
        return (0.5 * ((s0 + (m1 - m0) ** 2) / (s1 + 1e-10) + l1 - l0 - 1).sum(dim=1)).mean()

    def multinoulli_loss_with_logits(self, s, logits):
        B = s.shape[0]
        lg = logits.reshape(-1, self.nlabels)
        xe = F.cross_entropy(lg, s.reshape(-1).long(), reduction='none').reshape(B, -1)
        return xe.sum(dim=1).mean()

    def losses(self, s, s_out_list, mu, sigma, pmu, psigma):
        L = self.L
        ld = {}
        tot = 0.0
        if self.xent_weight is not None:
            acc = None
            for ii in reversed(range(L)):
                acc = s_out_list[ii] if acc is None else acc + s_out_list[ii]
                ld['residual_multinoulli_loss_lvl%d' % ii] = self.multinoulli_loss_with_logits(s, acc)
                tot = tot + self.xent_weight * ld['residual_multinoulli_loss_lvl%d' % ii]
        if self.KL_weight is not None:
            for ii in reversed(range(L)):
                w = 4 ** ii if self.exponential_weighting else 1
                ld['KL_divergence_loss_lvl%d' % ii] = w * self.KL_two_gauss_with_diag_cov(
                    mu[ii], sigma[ii], pmu[ii], psigma[ii])
                tot = tot + self.KL_weight * ld['KL_divergence_loss_lvl%d' % ii]
        if self.weight_decay is not None:
            wn = sum(0.5 * (self.P[n] ** 2).sum() for n, _, k in self.spec.entries if k == 'W')
            ld['weight_decay'] = self.weight_decay * wn
            tot = tot + ld['weight_decay']
        ld['total_loss'] = tot
        return ld

    # ---- whole-graph entry points ------------------------------------------------------------
    def one_hot(self, s):
        return F.one_hot(s.long(), self.nlabels).to(self.dtype)

    def forward_train(self, x, s, eps_post, eps_prior=None, training=True, new_stats=None):
        """posterior -> prior(generation_mode=False) -> likelihood(posterior z) -> losses
        (phiseg_model.py:37-83,113-130)."""
        x = x.to(self.dtype)
        s_oh = self.one_hot(s)
        if self.arch == 'det_unet':
            # posteriors.dummy / priors.dummy (posteriors.py:135-138, priors.py:130-133): constants, no KL term
            assert self.KL_weight is None, 'det_unet2D has no latent variables: KL_divergence_loss_weight must be None'
            s_out = self.likelihood(None, x, training, new_stats)
            ld = self.losses(s, s_out, None, None, None, None)
            return SimpleNamespace(z=[], mu=[], sigma=[], prior_z=[], prior_mu=[], prior_sigma=[], s_out_list=s_out,
                                   loss_dict=ld)
        if eps_prior is None:
            eps_prior = [torch.zeros_like(e) for e in eps_post]
        z, mu, sigma = self.posterior(x, s_oh, eps_post, training, new_stats)
        pz, pmu, psigma = self.prior(z, x, eps_prior, False, training, new_stats)
        s_out = self.likelihood(z, x, training, new_stats)
        ld = self.losses(s, s_out, mu, sigma, pmu, psigma)
        return SimpleNamespace(z=z, mu=mu, sigma=sigma, prior_z=pz, prior_mu=pmu, prior_sigma=psigma,
                               s_out_list=s_out, loss_dict=ld)

    def forward_sample(self, x, eps_prior, training=False):
        """prior(generation_mode=True) -> likelihood(prior z) -> sum over levels (phiseg_model.py:61-109)."""
        x = x.to(self.dtype)
        if self.arch == 'det_unet':
            pz, pmu, psigma = [], [], []
        else:
            pz, pmu, psigma = self.prior(None, x, eps_prior, True, training)
        s_list = self.likelihood(pz, x, training)
        s_out = s_list[-1]
        for i in range(len(s_list) - 1):
            s_out = s_out + s_list[i]
        return SimpleNamespace(prior_z=pz, prior_mu=pmu, prior_sigma=psigma, s_out_eval_list=s_list,
                               s_out_eval=s_out, s_out_eval_sm=torch.softmax(s_out, dim=-1))

    def latent_shapes(self, B):
        if self.arch == 'det_unet':
            return []
        if self.arch == 'probunet':
            return [(B, self.zdim0)]
        d = self.R - self.L
        return [(B, self.H >> (i + d), self.W >> (i + d), self.zdim0) for i in range(self.L)]

    def grads(self, x, s, eps_post, eps_prior=None):
        """d loss_tot / d theta by autograd (stands for optimizer.minimize's tf.gradients, phiseg_model.py:141)."""
        names = self.trainable_names()
        for n in names:
            self.P[n].requires_grad_(True)
            self.P[n].grad = None
        new_stats = {}
        out = self.forward_train(x, s, eps_post, eps_prior, True, new_stats)
        out.loss_dict['total_loss'].backward()
        g = {n: (self.P[n].grad.detach().clone() if self.P[n].grad is not None else None) for n in names}
        for n in names:
            self.P[n].requires_grad_(False)
            self.P[n].grad = None
        return out, g, new_stats

    def train_step(self, x, s, eps_post, lr, eps_prior=None, beta1=0.9, beta2=0.999, eps_hat=1e-8):
        """One iteration of the loop body phiseg_model.py:193-197: loss, gradients, tf.train.AdamOptimizer
        update  theta -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)  and BN moving-average updates."""
        out, g, new_stats = self.grads(x, s, eps_post, eps_prior)
        if self.adam is None:
            self.adam = {n: (torch.zeros_like(self.P[n]), torch.zeros_like(self.P[n])) for n in g}
        self.step += 1
        t = self.step
        lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
        with torch.no_grad():
            for n, gn in g.items():
                if gn is None:
                    continue
                m, v = self.adam[n]
                m.mul_(beta1).add_(gn, alpha=1 - beta1)
                v.mul_(beta2).addcmul_(gn, gn, value=1 - beta2)
                self.P[n].sub_(lr_t * m / (v.sqrt() + eps_hat))
            for k, v in new_stats.items():
                self.P[k] = v
        return float(out.loss_dict['total_loss'].detach()), out, g


# ----------------------------------------------------------------------------
# validation metrics (utils.py:103-118, 270-362; phiseg_model.py:596-606) -- plain numpy on whole masks
# ----------------------------------------------------------------------------
def _mask_distance(m1, m2, label_range):
    """1 - mean_l IoU(m1 == l, m2 == l); both empty -> IoU 1, exactly one empty -> IoU 0 (utils.py:272-292; jc of medpy)"""
    tot = 0.0
    for lbl in label_range:
        a, b = (m1 == lbl), (m2 == lbl)
        na, nb = int(a.sum()), int(b.sum())
        if na == 0 and nb == 0:
            tot += 1.0
        elif na == 0 or nb == 0:
            tot += 0.0
        else:
            tot += float(np.logical_and(a, b).sum()) / float(np.logical_or(a, b).sum())
    return 1.0 - tot / len(list(label_range))


def generalised_energy_distance(sample_arr, gt_arr, label_range):
    """utils.py:270-320: sample_arr [N,X,Y], gt_arr [M,X,Y] label masks.  (The reference passes nlabels-1 as the
    divisor together with label_range = range(1, nlabels), phiseg_model.py:586-588: the mean is over the foreground labels.)"""
    N, M = sample_arr.shape[0], gt_arr.shape[0]
    lr = list(label_range)
    d_sy = sum(_mask_distance(sample_arr[i], gt_arr[j], lr) for i in range(N) for j in range(M))
    d_ss = sum(_mask_distance(sample_arr[i], sample_arr[j], lr) for i in range(N) for j in range(N))
    d_yy = sum(_mask_distance(gt_arr[i], gt_arr[j], lr) for i in range(M) for j in range(M))
    return 2.0 / (N * M) * d_sy - d_ss / N ** 2 - d_yy / M ** 2


def ncc(a, v):
    """utils.py:103-118 with zero_norm=True: correlation of the standardised maps"""
    a = np.asarray(a, np.float64).ravel()
    v = np.asarray(v, np.float64).ravel()
    a = (a - a.mean()) / (a.std() * a.size)
    v = (v - v.mean()) / v.std()
    return float(np.sum(a * v))


def variance_ncc_dist(sample_sm, gt_onehot):
    """utils.py:323-362: sample_sm [N,X,Y,L] softmax samples, gt_onehot [M,X,Y,L]"""
    s = np.asarray(sample_sm, np.float64)
    g = np.asarray(gt_onehot, np.float64)
    logs = np.log(s + 1e-8)
    mean_seg = s.mean(axis=0)
    e_ss = np.mean(-np.sum(mean_seg[None] * logs, axis=-1), axis=0)
    vals = []
    for j in range(g.shape[0]):
        e_sy = np.mean(-np.sum(g[j][None] * logs, axis=-1), axis=0)
        vals.append(ncc(e_ss, e_sy))
    return float(np.mean(vals))


def per_label_dice(pred, gt, nlabels):
    """phiseg_model.py:596-606 (dc of medpy): both empty -> 1, exactly one empty -> 0"""
    out = []
    for lbl in range(nlabels):
        a, b = (pred == lbl), (gt == lbl)
        na, nb = int(a.sum()), int(b.sum())
        if na == 0 and nb == 0:
            out.append(1.0)
        elif na == 0 or nb == 0:
            out.append(0.0)
        else:
            out.append(2.0 * float(np.logical_and(a, b).sum()) / float(na + nb))
    return np.asarray(out)


# ----------------------------------------------------------------------------
# uncertainty maps (phiseg_model.py:378-475) -- numpy on stacked samples of ONE image, as the reference computes them
# ----------------------------------------------------------------------------
def _softmax_np(logits):
    z = np.asarray(logits, np.float64)
    e = np.exp(z - z.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def _xent_np(logits, gt):
    """tf.nn.softmax_cross_entropy_with_logits against one_hot(gt) (phiseg_model.py:304-311): logsumexp - logit[gt]"""
    z = np.asarray(logits, np.float64)
    m = z.max(axis=-1, keepdims=True)
    lse = np.log(np.exp(z - m).sum(axis=-1)) + m[..., 0]
    return lse - np.take_along_axis(z, np.asarray(gt)[..., None].astype(np.int64), axis=-1)[..., 0]


def sample_variance_sm_cov(logit_samples):
    """phiseg_model.py:386-403: logit_samples [S,X,Y,L] (s_out_eval of S prior samples of one image)"""
    segm = np.asarray(logit_samples, np.float64)[..., :-1].transpose((1, 2, 3, 0))
    segm = np.clip(segm, 1e-5, 1 - 1e-5)
    n = segm.shape[-1]
    corr = np.einsum('ghij,ghkj->ghik', segm, segm) / n
    mu = segm.mean(axis=-1)
    cov = corr - np.einsum('ghi,ghj->ghij', mu, mu)
    ev, _ = np.linalg.eig(cov)
    return np.sum(ev, axis=-1).real


def sample_variance_sm_cov_bf(logit_samples):
    """phiseg_model.py:414-430: per-pixel det(np.cov) of the softmax samples"""
    sm = _softmax_np(logit_samples).transpose((1, 2, 3, 0))
    out = np.zeros(sm.shape[:2])
    for i in range(sm.shape[0]):
        for j in range(sm.shape[1]):
            out[i, j] = np.linalg.det(np.cov(sm[i, j]))
    return out


def mean_variance_and_error_maps(logit_samples, gt):
    """phiseg_model.py:458-475: (argmax of the mean softmax, class-mean of np.std over the samples, mean cross entropy)"""
    sm = _softmax_np(logit_samples)
    errs = np.stack([_xent_np(z, gt) for z in np.asarray(logit_samples)])
    return np.argmax(sm.mean(axis=0), axis=-1), np.std(sm, axis=0).mean(axis=-1), errs.mean(axis=0)


# ----------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------
def synthetic_batch(B, H=128, W=128, nlabels=2, seed=1234):
    """x: smooth random field in [-0.5,0.5] (real LIDC input is image-0.5, data/lidc_data_loader.py:92);
    s: random ellipse 'lesions' (uint8), some images empty."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, H, W)).astype(np.float32)
    k = np.exp(-0.5 * (np.arange(-6, 7) / 2.5) ** 2)
    k /= k.sum()
    for ax in (1, 2):
        x = np.apply_along_axis(lambda v: np.convolve(v, k, mode='same'), ax, x)
    x = np.clip(x / (3 * x.std() + 1e-8), -0.5, 0.5).astype(np.float32)[..., None]
    yy, xx = np.mgrid[0:H, 0:W]
    s = np.zeros((B, H, W), np.uint8)
    for b in range(B):
        if rng.random() < 0.2:
            continue
        for lab in range(1, nlabels):
            cy, cx = rng.uniform(0.3 * H, 0.7 * H), rng.uniform(0.3 * W, 0.7 * W)
            ry, rx = rng.uniform(0.04 * H, 0.16 * H) / lab, rng.uniform(0.04 * W, 0.16 * W) / lab
            s[b][((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = lab
    return x, s


def synthetic_eps(shapes, seed=1234):
    rng = np.random.default_rng(seed + 77)
    return [rng.standard_normal(sh).astype(np.float32) for sh in shapes]
