"""Stand-in for the few TensorFlow names a reference experiment module touches
(`tf.train.AdamOptimizer`, `tf.train.MomentumOptimizer`; phiseg/experiments/phiseg_7_5.py:2,37 and
phiseg/phiseg_model.py:137).  They are selectors only: the optimizer arithmetic is phs_adam_step / phs_momentum_step."""


class _Optimizer:
    kind = None

    def __init__(self, learning_rate=None, **kwargs):
        self.learning_rate = learning_rate
        self.kwargs = kwargs


class AdamOptimizer(_Optimizer):
    kind = 'adam'


class MomentumOptimizer(_Optimizer):
    kind = 'momentum'


class train:
    AdamOptimizer = AdamOptimizer
    MomentumOptimizer = MomentumOptimizer


def optimizer_kind(opt):
    """Map whatever an experiment file put in `optimizer` to 'adam' / 'momentum'."""
    name = getattr(opt, 'kind', None) or getattr(opt, '__name__', str(opt))
    name = str(name).lower()
    if 'momentum' in name:
        return 'momentum'
    if 'adam' in name:
        return 'adam'
    raise ValueError('unsupported optimizer %r (the reference handles AdamOptimizer and MomentumOptimizer, '
                     'phiseg_model.py:137-140)' % (opt,))
