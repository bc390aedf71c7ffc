"""The C-ABI library builds, loads and exports every symbol include/vd_b200.h declares (no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'vd_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(vd_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from video_distillation_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the Python binding declares a signature for each of them
    _lib.lib()
    unbound = [n for n in names if n not in _lib.EXPORTED]
    assert not unbound, unbound
    assert _lib.lib().vd_abi_version() == 1


def test_argument_validation_without_gpu():
    """Bad arguments are rejected on the host before any CUDA call (negative return + message)."""
    from video_distillation_b200 import _lib
    lib = _lib.lib()
    g = _lib.ConvGeom(1, 3, 4, 8, 8, 4, 4, 99, 4, 3, 7, 7, 1, 2, 2, 1, 3, 3)       # inconsistent Ho
    rc = lib.vd_conv3d_fprop_f32(ctypes.c_void_p(16), ctypes.c_void_p(16), None, ctypes.c_void_p(16), ctypes.byref(g), None)
    assert rc < 0 and b'output extent' in lib.vd_last_error()
    plan = _lib.TcPlan()
    assert lib.vd_tc_plan_make(ctypes.byref(plan), 16, 100, 100) < 0
    assert lib.vd_tc_plan_make(ctypes.byref(plan), 16, 112, 112) == 0
    assert plan.embed_dim == 2048 and plan.T3p == 4 and plan.H3p == 2
    assert lib.vd_relu_maxpool_fwd_f32(ctypes.c_void_p(16), ctypes.c_void_p(16), None, 1, 4, 4, 4, 3, 2, 2, None) < 0


def test_cpu_tensors_are_refused():
    """No CPU fallback: ops raise instead of silently computing elsewhere."""
    import torch
    from video_distillation_b200 import ops
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.conv3d(torch.zeros(1, 3, 4, 8, 8), torch.zeros(4, 3, 3, 7, 7), None, (1, 2, 2), (1, 3, 3))
