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


def _truncated_normal(shape, std):
    w = torch.randn(tuple(shape), generator=_gen, dtype=torch.float32)
    bad = w.abs() > 2.0
    while bool(bad.any()):
        w[bad] = torch.randn(int(bad.sum()), generator=_gen, dtype=torch.float32)
        bad = w.abs() > 2.0
    return w * std


def _uniform(shape, limit):
    return (torch.rand(tuple(shape), generator=_gen, dtype=torch.float32) * 2.0 - 1.0) * limit


def get_weight_variable(shape, name=None, type='xavier_uniform', regularize=True, **kwargs):
    """utils.py:214-257.  Initialisers as TensorFlow 1.x defines them (fan_in = kh*kw*Cin, fan_out = kh*kw*Cout):
    'he_normal' = variance_scaling_initializer(factor=2, FAN_IN, uniform=False): a normal truncated at +-2 sigma with
    sigma = sqrt(1.3 * 2 / fan_in) - the one every layer of the PHiSeg path uses; init_weights=... = 'pretrained'.
    Named variables are created once and then reused (tf.get_variable semantics)."""
    if kwargs.get('init_weights') is not None:
        type = 'pretrained'
    full = _full(name or 'W')
    if name is not None and full in _store:
        return _store[full]
    shape = tuple(int(d) for d in shape)
    rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    fan_in, fan_out = rf * shape[-2], rf * shape[-1]
    if type == 'xavier_uniform':
        w = _uniform(shape, np.sqrt(6.0 / (fan_in + fan_out)))
    elif type == 'xavier_normal':
        w = _truncated_normal(shape, np.sqrt(1.3 * 2.0 / (fan_in + fan_out)))
    elif type == 'he_normal':
        w = _truncated_normal(shape, np.sqrt(1.3 * 2.0 / fan_in))
    elif type == 'he_uniform':
        w = _uniform(shape, np.sqrt(3.0 * 2.0 / fan_in))
    elif type == 'caffe_uniform':
        w = _uniform(shape, np.sqrt(3.0 * 1.0 / fan_in))
    elif type == 'simple':
        w = _truncated_normal(shape, float(kwargs.get('stddev', 0.02)))
    elif type == 'pretrained':
        w = torch.as_tensor(np.asarray(kwargs.get('init_weights'), dtype=np.float32)).reshape(shape)
    else:
        raise ValueError('Unknown initialisation requested: %s' % type)
    w = w.cuda().contiguous()
    if name is not None:
        _store[full] = w
        if regularize:
            _weight_variables.append(full)
    return w


def get_bias_variable(shape, name=None, init_value=0.0, **kwargs):
    """utils.py:261-271: constant initialiser, or init_biases=... for given values."""
    full = _full(name or 'b')
    if name is not None and full in _store:
        return _store[full]
    if kwargs.get('init_biases') is not None:
        b = torch.as_tensor(np.asarray(kwargs['init_biases'], dtype=np.float32)).reshape(tuple(shape)).cuda().contiguous()
    else:
        b = torch.full(tuple(shape), float(init_value), dtype=torch.float32, device='cuda')
    if name is not None:
        _store[full] = b
    return b


def get_constant_variable(shape, name, value):
    full = _full(name)
    if full not in _store:
        _store[full] = torch.full(tuple(shape), float(value), dtype=torch.float32, device='cuda')
    return _store[full]
