"""ctypes binding of libphiseg_sm100.so (the C-ABI declared in include/phiseg_sm100.h).

There is no CPU fallback: if the shared library is missing the import fails loudly.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libphiseg_sm100.so')

PHS_F32, PHS_BF16, PHS_F64 = 0, 1, 2
AUG_ROTATE, AUG_SCALE, AUG_FLIPLR, AUG_FLIPUD = 1, 2, 4, 8
NORM_BN_TRAIN, NORM_BN_INFER, NORM_GN = 0, 1, 2
IMPL_SIMT, IMPL_TC = 0, 1


class phs_tensor(ctypes.Structure):
    _fields_ = [('ptr', c_void_p), ('N', c_int32), ('H', c_int32), ('W', c_int32), ('C', c_int32),
                ('ld', c_int32), ('dtype', c_int32)]


class phs_aug_params(ctypes.Structure):
    """include/phiseg_sm100.h: one output image of phs_augment_batch (72 bytes)"""
    _fields_ = [('src', c_int32), ('annot', c_int32), ('flags', c_int32), ('crop', c_int32), ('px', c_int32),
                ('py', c_int32), ('minv', ctypes.c_double * 6)]


class phs_norm_pre(ctypes.Structure):
    """include/phiseg_sm100.h: the producer layer's normalisation handed to phs_conv2d_pre"""
    _fields_ = [('stats', c_void_p), ('mode', c_int32), ('eps', c_float), ('decay', c_float),
                ('moving_mean', c_void_p), ('moving_var', c_void_p), ('mean', c_void_p), ('rstd', c_void_p),
                ('gamma', c_void_p), ('beta', c_void_p), ('relu', c_int32)]


class PhisegError(RuntimeError):
    pass


_T = POINTER(phs_tensor)
_P = c_void_p          # any device pointer (float*, uint8_t*, bf16*)
_S = c_void_p          # cudaStream_t

# name -> argtypes (every function returns int unless noted)
SIGNATURES = {
    'phs_conv2d': [_T, _P, _P, _T, c_int, c_int, c_int, c_int, _S],
    'phs_conv2d_stats': [_T, _P, _P, _T, c_int, _P, _S],
    'phs_conv2d_stats_acc': [_T, _P, _P, _T, c_int, _P, _S],
    'phs_conv2d_pre': [_T, POINTER(phs_norm_pre), _P, _P, _T, _P, _S],
    'phs_conv2d_post': [_T, _P, _P, POINTER(phs_norm_pre), _T, c_int, _S],
    'phs_conv2d_pre_plan': [_T, _T, c_int, POINTER(c_int)],
    'phs_conv_halo_plan': [_T, _T, c_int, c_int, POINTER(c_int)],
    'phs_wgrad_halo_plan': [_T, _T, POINTER(c_int)],
    'phs_conv2d_wgrad': [_T, _T, _P, _P, c_int, c_int, c_int, _S],
    'phs_chan_stats': [_T, _P, _S],
    'phs_norm_finalize': [_P, c_int, c_int, c_int, c_int, c_float, c_float, _P, _P, _P, _P, _S],
    'phs_norm_act_fwd': [_T, _P, _P, _P, _P, c_int, _T, _S],
    'phs_norm_act_fwd_stats': [_T, _P, c_int, c_float, c_float, _P, _P, _P, _P, _P, _P, c_int, _T, _S],
    'phs_norm_bwd_reduce': [_T, _T, _P, _P, _P, _P, c_int, _P, _S],
    'phs_norm_bwd_reduce_remat': [_T, _T, _P, _P, _P, _P, c_int, _P, _T, _S],
    'phs_norm_bwd_reduce_bn': [_T, _T, _P, _P, _P, _P, c_int, _P, _S],
    'phs_norm_bwd_apply_bn': [_T, _T, _P, _P, _P, _P, c_int, _P, _T, _P, _P, c_int, _S],
    'phs_norm_bwd_finalize': [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_int, _S],
    'phs_norm_bwd_apply': [_T, _T, _P, _P, _P, _P, c_int, _P, _T, _S],
    'phs_avgpool2_fwd': [_T, _T, _S],
    'phs_avgpool2_bwd': [_T, _T, c_int, _S],
    'phs_upsample2_fwd': [_T, _T, _S],
    'phs_upsample2_bwd': [_T, _T, c_int, _S],
    'phs_latent_fwd': [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, c_float, _S],
    'phs_latent_bwd': [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P, _S],
    'phs_xent_multiscale': [POINTER(c_void_p), POINTER(c_void_p), _P, c_int, c_int, c_int, c_int, c_int, c_float, _P, _S],
    'phs_aggregate_logits': [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _S],
    'phs_adam_step': [_P, _P, _P, _P, c_int64, c_float, _P, c_float, c_float, c_float, c_float, _S],
    'phs_momentum_step': [_P, _P, _P, c_int64, c_float, _P, c_float, c_float, _S],
    'phs_weight_prep': [_P, _P, _P, c_int, _S],
    'phs_weight_prep_lo': [_P, _P, _P, c_int, _S],
    'phs_split_bf16': [_T, _T, _T, _S],
    'phs_copy_cast': [_T, _T, _S],
    'phs_im2col3x3': [_T, _T, _S],
    'phs_posterior_input': [_P, _P, c_int, c_int, c_int, c_int, c_int, _T, _S],
    'phs_broadcast_z': [_P, _T, _S],
    'phs_broadcast_z_bwd': [_T, _P, c_int, _S],
    'phs_fill_f32': [_P, c_int64, c_float, _S],
    'phs_axpy_f32': [_P, _P, c_int64, c_float, _S],
    'phs_sumsq_f32': [_P, c_int64, c_float, _P, _S],
    'phs_weight_decay': [_P, _P, _P, c_int, c_float, _P, _S],
    'phs_argmax_f32': [_P, c_int64, c_int, _P, _S],
    'phs_pairwise_label_stats': [_P, c_int, c_int, _P, c_int, c_int, c_int64, c_int, _P, _P, _P, _S],
    'phs_ncc_maps': [_P, _P, c_int, c_int, c_int64, c_int, _P, _P, _P, _S],
    'phs_sample_moments': [_P, _P, c_int, c_int, c_int64, c_int, c_int, c_float, c_float, _P, _S],
    'phs_augment_batch': [_P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, _S],
    'phs_sample_maps': [_P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, _P, _S],
}

_lib = None


def load():
    """Load the shared library once; raise if it has not been built (see __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('libphiseg_sm100.so is missing at %s: run `python __graft_entry__.py build` (or '
                          'phiseg-code_b200/build.py). There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, argt in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argt
        fn.restype = c_int
    lib.phs_version.restype = c_int
    lib.phs_arch.restype = c_int
    lib.phs_device_ok.restype = c_int
    lib.phs_last_error.restype = c_char_p
    _lib = lib
    return lib


def exported_symbols():
    return list(SIGNATURES) + ['phs_version', 'phs_arch', 'phs_device_ok', 'phs_last_error', 'phs_crc32c']


def check(rc, what=''):
    if rc != 0:
        msg = load().phs_last_error().decode('utf-8', 'replace')
        raise PhisegError('%s failed (rc=%d): %s' % (what, rc, msg))
