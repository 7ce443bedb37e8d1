"""Operator-level drop-in (tfwrapper/layers.py, tfwrapper/normalisation.py with the reference's names and argument
meaning) against plain PyTorch fp32 references of the same ops: conv2D (+bias rule, normalisation, activation order),
batch_norm (training statistics, moving averages with Bessel-corrected variance), group_norm2D (16 channels per group),
2x2 average pool, legacy bilinear x2, global average pool, crop_and_concat, and the reference's error behaviour."""
import importlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tfw(pkg):
    layers = importlib.import_module('phiseg_code_b200.tfwrapper.layers')
    norm = importlib.import_module('phiseg_code_b200.tfwrapper.normalisation')
    utils = importlib.import_module('phiseg_code_b200.tfwrapper.utils')
    utils.reset_variables(seed=7)
    return layers, norm, utils


def _conv_ref(x, w, b=None):
    """float64 on the CPU: cuDNN would run an fp32 convolution in TF32 and be the less accurate side"""
    x, w = x.detach().cpu().double(), w.detach().cpu().double()
    b = None if b is None else b.detach().cpu().double()
    y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=w.shape[0] // 2)
    return y.permute(0, 2, 3, 1).float().cuda()


def test_conv2d_batch_norm_relu(tfw):
    layers, norm, utils = tfw
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 16, 24, 5, generator=g).cuda()
    out = layers.conv2D(x, 'enc', num_filters=32, normalisation=norm.batch_norm, training=True)
    V = utils.get_variables()
    assert 'enc/W' in V and 'enc/b' not in V, 'batch norm switches the conv bias off (layers.py:126-128)'
    assert {'enc/batch_norm/BatchNorm/' + k for k in ('beta', 'gamma', 'moving_mean', 'moving_variance')} <= set(V)
    w = V['enc/W']
    assert tuple(w.shape) == (3, 3, 5, 32) and float(w.abs().max()) <= 2.0 * np.sqrt(1.3 * 2.0 / 45) + 1e-6
    y = _conv_ref(x, w)
    m, v = y.mean(dim=(0, 1, 2)), y.var(dim=(0, 1, 2), unbiased=False)
    ref = torch.relu((y - m) / torch.sqrt(v + 1e-3))
    assert float((out - ref).abs().max()) < 2e-4
    cnt = 3 * 16 * 24
    assert torch.allclose(V['enc/batch_norm/BatchNorm/moving_mean'], 0.01 * m, atol=1e-5)
    assert torch.allclose(V['enc/batch_norm/BatchNorm/moving_variance'], 0.99 + 0.01 * v * cnt / (cnt - 1), atol=1e-5)
    # inference mode uses the moving statistics; the same scope reuses the same variables
    out_inf = layers.conv2D(x, 'enc', num_filters=32, normalisation=norm.batch_norm, training=False)
    mm, mv = V['enc/batch_norm/BatchNorm/moving_mean'], V['enc/batch_norm/BatchNorm/moving_variance']
    assert float((out_inf - torch.relu((y - mm) / torch.sqrt(mv + 1e-3))).abs().max()) < 2e-4


def test_conv2d_group_norm_bias_and_order(tfw):
    layers, norm, utils = tfw
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 8, 8, 32, generator=g).cuda()
    utils.set_variable('z/b', torch.randn(64, generator=g).numpy())
    out = layers.conv2D(x, 'z', kernel_size=(1, 1), num_filters=64, normalisation=norm.group_norm2D, training=True)
    V = utils.get_variables()
    assert tuple(V['z/group_norm/gamma'].shape) == (1, 1, 1, 64)
    y = _conv_ref(x, V['z/W'], V['z/b'])
    yg = y.reshape(2, 8, 8, 4, 16)
    m = yg.mean(dim=(1, 2, 4), keepdim=True)
    v = yg.var(dim=(1, 2, 4), unbiased=False, keepdim=True)
    ref = torch.relu(((yg - m) / torch.sqrt(v + 1e-5)).reshape(2, 8, 8, 64))
    assert float((out - ref).abs().max()) < 2e-4
    # no normalisation, identity activation: plain convolution + bias
    lin = layers.conv2D(x, 'z', kernel_size=(1, 1), num_filters=64, activation=layers.identity)
    assert float((lin - y).abs().max()) < 2e-4
    # normalise_post_activation swaps the order
    post = layers.conv2D(x, 'z', kernel_size=(1, 1), num_filters=64, normalisation=norm.group_norm2D,
                         normalise_post_activation=True, training=True)
    ya = torch.relu(y).reshape(2, 8, 8, 4, 16)
    ma, va = ya.mean(dim=(1, 2, 4), keepdim=True), ya.var(dim=(1, 2, 4), unbiased=False, keepdim=True)
    assert float((post - ((ya - ma) / torch.sqrt(va + 1e-5)).reshape(2, 8, 8, 64)).abs().max()) < 2e-4


def test_resampling_ops_and_errors(tfw, oracle):
    layers, norm, utils = tfw
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 6, 10, 8, generator=g)
    xd = x.cuda()
    assert float((layers.averagepool2D(xd).cpu() - oracle.averagepool2d(x)).abs().max()) < 1e-6
    assert float((layers.bilinear_upsample2D(xd, 'up', 2).cpu() - oracle.bilinear_upsample2d(x)).abs().max()) < 1e-6
    assert float((layers.global_averagepool2D(xd).cpu() - x.mean(dim=(1, 2))).abs().max()) < 1e-5
    cc = layers.crop_and_concat([xd, layers.bilinear_upsample2D(xd, 'up', 2)[:, :8, :12]], axis=-1)
    assert tuple(cc.shape) == (2, 6, 10, 16)
    with pytest.raises(ValueError):
        layers.conv2D(xd, 'bad', strides=(2, 2))
    with pytest.raises(ValueError):
        layers.averagepool2D(xd[:, :5])
    with pytest.raises(TypeError):
        layers.conv2D(x.numpy(), 'bad')
    with pytest.raises(TypeError):
        norm.batch_norm(xd)                         # `training` is required, as in the reference signature
    with pytest.raises(ValueError):
        utils.get_weight_variable([3, 3, 1, 1], name='w2', type='no_such_initialiser')
    wx = utils.get_weight_variable([3, 3, 4, 8], name='wx')          # the reference's default: xavier_uniform
    assert float(wx.abs().max()) <= np.sqrt(6.0 / (36 + 72)) + 1e-6 and wx is utils.get_weight_variable([3, 3, 4, 8], name='wx')
    wp = utils.get_weight_variable([1, 1, 2, 2], name='wp', init_weights=np.arange(4.0))
    assert wp.flatten().tolist() == [0.0, 1.0, 2.0, 3.0]
