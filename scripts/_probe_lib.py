"""Tuning-only build of the library: the product sources + csrc/tc_probe.cu compiled with -DVD_PROBE into
scripts/libvd_b200_probe.so.  The probe entry points (scripts/vd_b200_probe.h) are not part of libvd_b200.so.

    from _probe_lib import use_probe_library;  use_probe_library()    # before anything else touches video_distillation_b200._lib
"""
import ctypes
import os
import sys
from ctypes import c_int, c_uint32, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
PROBE_LIB = os.path.join(HERE, 'libvd_b200_probe.so')


def use_probe_library():
    """Build (if stale) and make video_distillation_b200._lib load the probe library instead of the product one."""
    from video_distillation_b200 import _lib, build
    build.build_probe(PROBE_LIB)
    _lib.LIB_PATH = PROBE_LIB
    _lib._lib = None
    lib = _lib.lib()
    P = c_void_p
    sig = {
        'vd_tc_probe': (c_int, [P, P, P, c_int, c_int, c_int, c_uint32, c_uint32, c_uint32, c_uint32, c_uint32, c_int, P]),
        'vd_tc_mma_rate': (c_int, [P, c_int, c_int, c_int, c_uint32, c_uint32, c_uint32, c_int, c_int, c_int, P]),
        'vd_tc_mma_rate2': (c_int, [P, c_int, c_int, c_int, c_uint32, c_uint32, c_uint32, c_int, c_uint32, c_uint32, c_uint32, c_int,
                            c_uint32, c_int, c_int, c_uint32, c_int, c_int, P]),
        'vd_tc_set_profile_buffer': (c_int, [P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
