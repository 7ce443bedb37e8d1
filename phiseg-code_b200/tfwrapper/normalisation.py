"""Selectable normalisation symbols with the reference's names (tfwrapper/normalisation.py:17,145).

In the reference these are graph-building functions.  Here they play two roles: an experiment file assigns them to
`layer_norm` (the engine reads `.kind`), and called on a float32 CUDA tensor [N,H,W,C] they ARE the operator
(eager forward through csrc/elementwise.cu: phs_chan_stats, phs_norm_finalize, phs_norm_act_fwd), with their variables
in tfwrapper/utils.py under the TensorFlow names."""


class _Norm:
    def __init__(self, kind, doc):
        self.kind = kind
        self.__doc__ = doc
        self.__name__ = kind

    def __call__(self, x, *a, **k):
        import torch
        if not (torch.is_tensor(x) and x.is_cuda and x.dim() == 4 and x.dtype == torch.float32):
            raise TypeError('%s: expected a float32 CUDA tensor [N,H,W,C] (as a bare symbol it is the selector for '
                            'exp_config.layer_norm)' % self.kind)
        return _apply(self.kind, x.contiguous(), *a, **k)

    def __repr__(self):
        return '<normalisation %s>' % self.kind


def _apply(kind, x, training=None, moving_average_decay=0.99, eps=None, scope=None, **kwargs):
    """batch_norm(x, training, moving_average_decay=0.99, scope='batch_norm') (normalisation.py:145-163) /
    group_norm2D(x, eps=1e-5, scope='group_norm') (:17-36)."""
    import torch
    from .. import lib as L
    from . import utils
    N, H, W, C = x.shape
    h, st = L.load(), torch.cuda.current_stream().cuda_stream
    xd = L.phs_tensor(x.data_ptr(), N, H, W, C, C, L.PHS_F32)
    y = torch.empty_like(x)
    yd = L.phs_tensor(y.data_ptr(), N, H, W, C, C, L.PHS_F32)
    stats = torch.empty(N * C * 2, device=x.device, dtype=torch.float64)
    mean, rstd = torch.empty(N * C, device=x.device), torch.empty(N * C, device=x.device)
    if kind == 'batch_norm':
        if training is None:
            raise TypeError("batch_norm() missing required argument 'training'")
        with utils.variable_scope(scope or 'batch_norm'), utils.variable_scope('BatchNorm'):
            beta = utils.get_constant_variable([C], 'beta', 0.0)
            gamma = utils.get_constant_variable([C], 'gamma', 1.0)
            mm = utils.get_constant_variable([C], 'moving_mean', 0.0)
            mv = utils.get_constant_variable([C], 'moving_variance', 1.0)
        mode = L.NORM_BN_TRAIN if training else L.NORM_BN_INFER
        if training:
            L.check(h.phs_chan_stats(xd, stats.data_ptr(), st), 'phs_chan_stats')
        L.check(h.phs_norm_finalize(stats.data_ptr(), N, H * W, C, mode, 1e-3, float(moving_average_decay), mm.data_ptr(),
                                    mv.data_ptr(), mean.data_ptr(), rstd.data_ptr(), st), 'phs_norm_finalize')
    else:
        if 'num_groups' in kwargs and kwargs['num_groups'] != max(2, C // 16):
            raise ValueError('group_norm2D: only the default grouping max(2, C // 16) is implemented')
        with utils.variable_scope(scope or 'group_norm'):
            gamma = utils.get_constant_variable([1, 1, 1, C], 'gamma', 1.0)
            beta = utils.get_constant_variable([1, 1, 1, C], 'beta', 0.0)
        L.check(h.phs_chan_stats(xd, stats.data_ptr(), st), 'phs_chan_stats')
        L.check(h.phs_norm_finalize(stats.data_ptr(), N, H * W, C, L.NORM_GN, 1e-5 if eps is None else float(eps), 0.0, None,
                                    None, mean.data_ptr(), rstd.data_ptr(), st), 'phs_norm_finalize')
    L.check(h.phs_norm_act_fwd(xd, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 0, yd, st),
            'phs_norm_act_fwd')
    return y


def identity(x, **kwargs):
    """normalisation.py:166-171"""
    return x


batch_norm = _Norm('batch_norm', 'tf.contrib.layers.batch_norm(decay=0.99, epsilon=1e-3), normalisation.py:145-163')
group_norm2D = _Norm('group_norm', 'groups of 16 channels, eps 1e-5, normalisation.py:17-36')


def norm_kind(sym):
    kind = getattr(sym, 'kind', None)
    if kind is None:
        name = getattr(sym, '__name__', str(sym))
        kind = {'batch_norm': 'batch_norm', 'group_norm2D': 'group_norm'}.get(name)
    if kind not in ('batch_norm', 'group_norm'):
        raise ValueError('unsupported layer_norm %r: batch_norm and group_norm2D are implemented' % (sym,))
    return kind
