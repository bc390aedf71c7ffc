"""Hardware floor of tcgen05.mma (cycles per MMA) and how much per-step scalar work the issue
thread can hide: each step issues n_acc MMAs back to back, then spins `delay` cycles."""
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
for n_acc, ncols in ((1, 256), (1, 224), (2, 224), (4, 128)):
    for delay in (0, 50, 100, 150, 200, 300, 400):
        iters = 4096
        _lib.check(lib.vd_tc_mma_rate(_lib.ptr(out), n_acc, ncols, iters, NS, NS, 288, 1, grid, delay, _lib.stream()), 'mma_rate')
        torch.cuda.synchronize()
        o = out.cpu()[:grid * 2].view(grid, 2).double()
        total = o[:, 1].mean().item()
        print(f'n_acc={n_acc} N={ncols:3d} delay={delay:3d}: {total / iters:7.1f} cyc/step '
              f'(ideal {n_acc * ncols / 2:5.1f}; serial {n_acc * ncols / 2 + delay:5.1f})', flush=True)
