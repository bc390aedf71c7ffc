"""tcgen05.mma rate in cycles AND wall time at full-chip scale (148 CTAs), long enough (~10 ms) for the
clock / power management to show: cycles per MMA (clock64), ns per MMA (CUDA events), the implied SM clock
and TFLOP/s.  Variants emulate the stage boundaries of ws_gemm_kernel (commit / fence every G MMAs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _probe_lib import use_probe_library  # noqa: E402
use_probe_library()               # probe entry points live in scripts/libvd_b200_probe.so, not in the product library
from video_distillation_b200 import _lib  # noqa: E402

lib = _lib.lib()
out = torch.zeros(148 * 2, dtype=torch.int64, device='cuda')
V1 = 1 << 14
NS = 8 | V1
grid = 148
iters = 100000
for n_acc, ncols in ((1, 224), (1, 256), (2, 256)):
    for name, vary in (('back-to-back', 1), ('commit/11', 1 | (11 << 8) | (1 << 16)), ('commit+fence/11', 1 | (11 << 8) | (3 << 16)),
                       ('commit+fence/4', 1 | (4 << 8) | (3 << 16))):
        def go():
            _lib.check(lib.vd_tc_mma_rate(_lib.ptr(out), n_acc, ncols, iters, NS, NS, 288, vary, grid, 0, _lib.stream()), 'mma_rate')
        go()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        go()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        o = out.cpu()[:grid * 2].view(grid, 2).double()
        cyc = o[:, 1].mean().item()
        mmas = iters * n_acc
        print(f'n_acc={n_acc} N={ncols:3d} {name:16s}: {cyc / mmas:7.1f} cyc/MMA (ideal {ncols / 2:5.1f})  {ms * 1e6 / mmas:7.1f} ns/MMA  '
              f'clock {cyc / (ms * 1e6):5.2f} GHz  {grid * mmas * 2 * 128 * ncols * 16 / (ms * 1e-3) / 1e12:7.1f} TFLOP/s', flush=True)
