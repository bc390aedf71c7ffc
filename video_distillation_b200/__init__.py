"""video_distillation_b200 — the distillation inner loop of yuz1wan/video_distillation on B200.

Host code is Python/PyTorch (device memory, streams, autograd bookkeeping, torch.distributed);
every tensor operation of the hot path is a hand-written sm_100a kernel in libvd_b200.so reached
through the C ABI of include/vd_b200.h.  There is no CPU or ATen fallback.
"""
__version__ = '0.1.0'

from . import _lib  # noqa: F401


def library_path():
    return _lib.LIB_PATH
