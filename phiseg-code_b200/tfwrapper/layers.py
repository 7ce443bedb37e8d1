"""Eager operator layer with the reference's names and argument meaning (tfwrapper/layers.py) for the ops on the PHiSeg
path: conv2D (:94-145), averagepool2D (:44-54), global_averagepool2D (:70-78), bilinear_upsample2D (:336-345),
crop_and_concat (:586-622).  Tensors are float32 torch CUDA tensors in NHWC; every arithmetic step is a kernel of
libphiseg_sm100.so reached through the C-ABI (no CPU fallback); variables come from tfwrapper/utils.py.

The training / sampling engine does not go through this layer (it lays the same kernels down as a static launch
program); this is the operator-level drop-in a user of the reference's layer functions would call."""
import torch

from .. import lib as L
from . import normalisation as tfnorm
from . import utils


def _check(x, what):
    if not (torch.is_tensor(x) and x.is_cuda and x.dim() == 4 and x.dtype == torch.float32):
        raise TypeError('%s: expected a float32 CUDA tensor [N,H,W,C], got %r' % (what, type(x)))
    return x.contiguous()


def _desc(t):
    N, H, W, C = t.shape
    return L.phs_tensor(t.data_ptr(), N, H, W, C, C, L.PHS_F32)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def relu(x):
    """tf.nn.relu (STANDARD_NONLINEARITY, layers.py:14)."""
    x = _check(x, 'relu')
    N, H, W, C = x.shape
    y = torch.empty_like(x)
    one, zero = torch.ones(N * C, device=x.device), torch.zeros(N * C, device=x.device)
    xd, yd = _desc(x), _desc(y)
    L.check(L.load().phs_norm_act_fwd(xd, zero.data_ptr(), one.data_ptr(), one.data_ptr(), zero.data_ptr(), 1, yd,
                                      _stream()), 'phs_norm_act_fwd')
    return y


def identity(x, **kwargs):
    return x


STANDARD_NONLINEARITY = relu


def averagepool2D(x, kernel_size=(2, 2), strides=(2, 2), padding="SAME"):
    """tf.nn.avg_pool 2x2 stride 2 (the only form the networks use; sizes must be even)."""
    x = _check(x, 'averagepool2D')
    if tuple(kernel_size) != (2, 2) or tuple(strides) != (2, 2):
        raise ValueError('averagepool2D: only kernel_size=(2,2), strides=(2,2) is implemented')
    N, H, W, C = x.shape
    if H % 2 or W % 2:
        raise ValueError('averagepool2D: spatial size %dx%d must be even' % (H, W))
    y = torch.empty((N, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
    L.check(L.load().phs_avgpool2_fwd(_desc(x), _desc(y), _stream()), 'phs_avgpool2_fwd')
    return y


def global_averagepool2D(x, name=None):
    """tf.reduce_mean over H and W: [N, C]."""
    x = _check(x, 'global_averagepool2D')
    N, H, W, C = x.shape
    stats = torch.empty(N * C * 2, device=x.device, dtype=torch.float64)
    L.check(L.load().phs_chan_stats(_desc(x), stats.data_ptr(), _stream()), 'phs_chan_stats')
    return (stats.view(N, C, 2)[..., 0] / float(H * W)).float()


def bilinear_upsample2D(x, name, factor):
    """tf.image.resize_images(x, [factor*H, factor*W]) = legacy bilinear, align_corners=False (factor 2)."""
    x = _check(x, 'bilinear_upsample2D')
    if factor != 2:
        raise ValueError('bilinear_upsample2D: only factor=2 is implemented')
    N, H, W, C = x.shape
    y = torch.empty((N, 2 * H, 2 * W, C), dtype=x.dtype, device=x.device)
    L.check(L.load().phs_upsample2_fwd(_desc(x), _desc(y), _stream()), 'phs_upsample2_fwd')
    return y


def crop_and_concat(inputs, axis=-1):
    """Centre-crop every input to the smallest spatial size, then concatenate (data movement only)."""
    hs = min(int(t.shape[1]) for t in inputs)
    ws = min(int(t.shape[2]) for t in inputs)
    out = []
    for t in inputs:
        dh, dw = (int(t.shape[1]) - hs) // 2, (int(t.shape[2]) - ws) // 2
        out.append(t[:, dh:dh + hs, dw:dw + ws, :])
    return torch.cat(out, dim=axis).contiguous()


def conv2D(x, name, kernel_size=(3, 3), num_filters=32, strides=(1, 1), activation=STANDARD_NONLINEARITY,
           normalisation=identity, normalise_post_activation=False, dropout_p=None, padding="SAME",
           weight_init='he_normal', add_bias=True, **kwargs):
    """Standard 2-D convolutional layer: conv (+bias) -> normalisation -> activation.  kwargs carries `training` (and is
    handed to the normalisation).  As in the reference the bias is dropped when the normalisation is batch_norm."""
    x = _check(x, 'conv2D')
    if tuple(strides) != (1, 1) or padding != "SAME":
        raise ValueError('conv2D: only strides=(1,1), padding="SAME" is implemented (all the PHiSeg path uses)')
    if dropout_p is not None:
        raise ValueError('conv2D: dropout is not on the PHiSeg path')
    if kernel_size[0] != kernel_size[1] or kernel_size[0] not in (1, 3):
        raise ValueError('conv2D: kernel_size must be (1,1) or (3,3)')
    N, H, W, cin = x.shape
    with utils.variable_scope(name):
        weights = utils.get_weight_variable([kernel_size[0], kernel_size[1], cin, num_filters], name='W',
                                            type=weight_init, regularize=True)
        if add_bias and normalisation is tfnorm.batch_norm:
            add_bias = False                      # layers.py:126-128
        biases = utils.get_bias_variable([num_filters], name='b') if add_bias else None
        op = torch.empty((N, H, W, num_filters), dtype=torch.float32, device=x.device)
        L.check(L.load().phs_conv2d(_desc(x), weights.data_ptr(), biases.data_ptr() if biases is not None else None,
                                    _desc(op), kernel_size[0], 0, 0, L.IMPL_SIMT, _stream()), 'phs_conv2d')
        if not normalise_post_activation:
            op = activation(normalisation(op, **kwargs))
        else:
            op = normalisation(activation(op), **kwargs)
    return op
