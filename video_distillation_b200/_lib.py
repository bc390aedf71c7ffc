"""ctypes binding of libvd_b200.so (the C ABI declared in include/vd_b200.h).

The library is loaded lazily; on a machine with a GPU a missing library is a hard error (there
is no CPU or PyTorch fallback anywhere in this package).
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libvd_b200.so')

_lib = None


class ConvGeom(Structure):
    _fields_ = [(n, c_int32) for n in ('N', 'Cin', 'T', 'H', 'W', 'Cout', 'To', 'Ho', 'Wo',
                                       'kt', 'kh', 'kw', 'st', 'sh', 'sw', 'pt', 'ph', 'pw')]


class TcPlan(Structure):
    _fields_ = ([(n, c_int32) for n in ('T', 'H', 'W', 'c1', 'c2', 'c3', 'T1', 'H1', 'W1', 'T1p', 'H1p', 'W1p',
                                        'T2', 'H2', 'W2', 'T2p', 'H2p', 'W2p', 'T3', 'H3', 'W3', 'T3p', 'H3p', 'W3p',
                                        'embed_dim')] +
                [('_pad0', c_int32)] +
                [(n, c_int64) for n in ('x0_bytes_per_video', 'a1_bytes_per_video', 'a2_bytes_per_video',
                                        'w0_bytes', 'w1_bytes', 'w2_bytes', 'tab_bytes',
                                        'wt0_bytes', 'wt1_bytes', 'wt2_bytes',
                                        'dy0_bytes_per_video', 'dy1_bytes_per_video', 'dy2_bytes_per_video',
                                        'col0_bytes_per_video', 'col1_bytes_per_video', 'col2_bytes_per_video')] +
                [('reserved', c_int32 * 8)])


def _declare(lib):
    P = c_void_p
    sig = {
        'vd_last_error': (c_char_p, []),
        'vd_abi_version': (c_int, []),
        'vd_launch_count': (c_int64, []),
        'vd_launch_count_reset': (None, []),
        'vd_conv3d_fprop_f32': (c_int, [P, P, P, P, POINTER(ConvGeom), P]),
        'vd_conv3d_dgrad_f32': (c_int, [P, P, P, POINTER(ConvGeom), P]),
        'vd_conv3d_wgrad_f32': (c_int, [P, P, P, P, POINTER(ConvGeom), P]),
        'vd_relu_maxpool_fwd_f32': (c_int, [P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, P]),
        'vd_route_scatter_f32': (c_int, [P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, P]),
        'vd_route_gather_f32': (c_int, [P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, P]),
        'vd_inorm_relu_fwd_f32': (c_int, [P, P, P, P, P, P, c_int, c_int, c_int64, P]),
        'vd_inorm_relu_bwd_f32': (c_int, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int64, P]),
        'vd_inorm_relu_avgpool_fwd_f32': (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
        'vd_avgpool2_fwd_f32': (c_int, [P, P, c_int64, c_int, c_int, c_int, P]),
        'vd_avgpool2_bwd_f32': (c_int, [P, P, c_int64, c_int, c_int, c_int, P]),
        'vd_compose_fwd_f32': (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
        'vd_compose_bwd_f32': (c_int, [P, P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
        'vd_compose_bwd_fused_f32': (c_int, [P, P, P, P, P, P, P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, P]),
        'vd_compose_fwd_ex_f32': (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int64, c_int64, P]),
        'vd_compose_bwd_fused_ex_f32': (c_int, [P, P, P, P, P, P, P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int,
                                                c_int64, c_int64, P]),
        'vd_class_mean_f32': (c_int, [P, P, c_int, c_int, c_int, P]),
        'vd_class_sum_ragged_f32': (c_int, [P, P, P, c_int, c_int, P]),
        'vd_dm_loss_f32': (c_int, [P, P, P, P, c_int, c_int, c_int, c_float, P]),
        'vd_dm_loss_ex_f32': (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_float, P]),
        'vd_sgd_momentum_f32': (c_int, [P, P, P, c_int64, c_float, c_float, c_int, P]),
        'vd_axpy_f32': (c_int, [P, P, P, c_int64, c_float, P]),
        'vd_sqdist_f32': (c_int, [P, P, P, c_int64, P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTED = None


def lib():
    """The loaded shared library; raises if it has not been built (no fallback)."""
    global _lib, EXPORTED
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: run `python -m video_distillation_b200.build` '
                '(there is no CPU / PyTorch fallback for the hot path)')
        _lib = ctypes.CDLL(LIB_PATH)
        EXPORTED = _declare(_lib)
        from . import _lib_tc
        _lib_tc.declare(_lib)
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().vd_last_error().decode()
        raise RuntimeError(f'libvd_b200 {what} failed (rc={rc}): {msg}')


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Refuses CPU tensors: the hot path is CUDA only."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('video_distillation_b200: tensors must live on a CUDA device (no CPU fallback)')
    if not t.is_contiguous():
        raise RuntimeError('video_distillation_b200: tensor must be contiguous at the C ABI')
    return c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(lib().vd_launch_count())


def launch_count_reset():
    lib().vd_launch_count_reset()
