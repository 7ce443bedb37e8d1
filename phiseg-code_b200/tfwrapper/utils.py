"""Variable helpers with the reference's names (tfwrapper/utils.py:214-271) for the eager operator layer
(tfwrapper/layers.py): variables are float32 torch CUDA tensors kept in a process-wide store under the names
TensorFlow's variable scopes would produce ('<scope>/W', '<scope>/b', '<scope>/batch_norm/BatchNorm/gamma', ...).
The training engine (engine.Params) keeps its own flat buffers with the same names; this store serves the op-level API."""
import contextlib

import numpy as np
import torch

_scope = []
_store = {}
_weight_variables = []          # the reference's 'weight_variables' collection (utils.py:254-255)
_gen = torch.Generator().manual_seed(1234)


@contextlib.contextmanager
def variable_scope(name):
    """tf.variable_scope(name): nests the names of the variables created inside."""
    _scope.append(name)
    try:
        yield
    finally:
        _scope.pop()


def _full(name):
    return '/'.join(_scope + [name])


def reset_variables(seed=1234):
    _store.clear()
    del _weight_variables[:]
    _gen.manual_seed(seed)


def get_variables():
    return _store


def set_variable(name, value, device='cuda'):
    _store[name] = torch.as_tensor(np.asarray(value, dtype=np.float32)).to(device).contiguous()
    return _store[name]


def get_weight_variable(shape, name=None, type='he_normal', regularize=True, **kwargs):
    """utils.py:214-257.  'he_normal' = variance_scaling_initializer(factor=2, FAN_IN, uniform=False): a normal truncated
    at +-2 sigma with sigma = sqrt(1.3 * 2 / fan_in); created once, then reused (tf.get_variable semantics)."""
    full = _full(name or 'W')
    if full not in _store:
        if type != 'he_normal':
            raise ValueError('Unknown initialisation requested: %s' % type)     # the PHiSeg path only uses he_normal
        from ..engine import he_normal
        _store[full] = he_normal(_gen, tuple(shape)).cuda().contiguous()
        if regularize:
            _weight_variables.append(full)
    return _store[full]


def get_bias_variable(shape, name=None, init_value=0.0):
    """utils.py:261-271: constant initialiser."""
    full = _full(name or 'b')
    if full not in _store:
        _store[full] = torch.full(tuple(shape), float(init_value), dtype=torch.float32, device='cuda')
    return _store[full]


def get_constant_variable(shape, name, value):
    full = _full(name)
    if full not in _store:
        _store[full] = torch.full(tuple(shape), float(value), dtype=torch.float32, device='cuda')
    return _store[full]
