"""MMA-rate probe: time ws_gemm_kernel over dummy operands for several shared-memory layouts.

    python scripts/tc_probe.py   (on the GPU box; prints cycles per MMA at the current SM clock)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _probe_lib import use_probe_library  # noqa: E402
use_probe_library()               # probe entry points live in scripts/libvd_b200_probe.so, not in the product library
from video_distillation_b200 import _lib  # noqa: E402

lib = _lib.lib()
dev = 'cuda'
pix = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
w = torch.zeros(1 << 16, dtype=torch.uint8, device=dev)
raw = torch.empty(148 * 4 * 128 * 256, dtype=torch.float32, device=dev)
V1 = 1 << 14


def hi(sbo_bytes, layout):
    return (sbo_bytes >> 4) | V1 | (layout << 29)


def run(name, ncols, a_lbo16, a_hi, b_lbo16, b_hi, b_step16, n_acc=1, n_sa=32, n_steps=64):
    def go():
        _lib.check(lib.vd_tc_probe(_lib.ptr(pix), _lib.ptr(w), _lib.ptr(raw), ncols, n_sa, n_steps, a_lbo16, a_hi, b_lbo16,
                                   b_hi, b_step16, n_acc, _lib.stream()), 'probe')
    go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        go()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    mmas = n_sa * n_steps * n_acc
    ns_per = ms * 1e6 / mmas
    tflops = 148 * mmas * 2 * 128 * ncols * 16 / (ms * 1e-3) / 1e12
    print(f'{name:58s} N={ncols:3d} acc={n_acc}  {ms:8.3f} ms  {ns_per:7.1f} ns/MMA  {tflops:7.1f} TFLOP/s', flush=True)


A_NS = (128, hi(128, 0))
A_SW128 = (1, hi(1024, 2))
A_SW32 = (1, hi(256, 6))
for N in (224, 256, 128):
    run('A no-swz | B no-swz planar LBO=4608', N, *A_NS, 288, hi(128, 0), 1)
run('A no-swz | B no-swz planar LBO=4608, 2 acc', 224, *A_NS, 288, hi(128, 0), 1, n_acc=2)
run('A no-swz | B no-swz planar LBO=4624 (odd x16)', 224, *A_NS, 289, hi(128, 0), 1)
run('A no-swz | B no-swz compact LBO=128 SBO=256', 224, *A_NS, 8, hi(256, 0), 16)
run('A sw128  | B sw128 (reference GEMM layout)', 224, *A_SW128, 1, hi(1024, 2), 8)
run('A sw128  | B sw128 (reference GEMM layout)', 256, *A_SW128, 1, hi(1024, 2), 8)
run('A sw128  | B sw128, 2 acc', 256, *A_SW128, 1, hi(1024, 2), 8, n_acc=2)
run('A no-swz | B sw128', 224, *A_NS, 1, hi(1024, 2), 8)
run('A sw128  | B no-swz planar', 224, *A_SW128, 288, hi(128, 0), 1)
run('A no-swz | B sw32 (row 32 B), shift 1 row', 224, *A_NS, 1, hi(256, 6), 2)
run('A sw32   | B sw32, shift 1 row', 224, *A_SW32, 1, hi(256, 6), 2)
run('A sw32   | B sw32, aligned', 224, *A_SW32, 1, hi(256, 6), 16)
run('A no-swz | B sw64 (row 64 B)', 224, *A_NS, 1, hi(512, 4), 4)
