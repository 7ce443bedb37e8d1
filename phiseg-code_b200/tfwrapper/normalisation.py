"""Selectable normalisation symbols with the reference's names (tfwrapper/normalisation.py:17,145).

In the reference these are graph-building functions; here they are markers an experiment file assigns to
`layer_norm`; the arithmetic lives in csrc/elementwise.cu (phs_chan_stats, phs_norm_finalize, phs_norm_act_fwd and
the three backward kernels)."""


class _Norm:
    def __init__(self, kind, doc):
        self.kind = kind
        self.__doc__ = doc
        self.__name__ = kind

    def __call__(self, *a, **k):
        raise TypeError('%s is a selector for exp_config.layer_norm; the kernels are invoked by the engine' % self.kind)

    def __repr__(self):
        return '<normalisation %s>' % self.kind


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
