"""Hardware floor of tcgen05.mma (cycles per MMA) for n_acc interleaved accumulators."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200 import _lib  # noqa: E402

lib = _lib.lib()
out = torch.zeros(148 * 2, dtype=torch.int64, device='cuda')
V1 = 1 << 14
NS = 8 | V1
SW128 = (1024 >> 4) | V1 | (2 << 29)
for grid in (1, 148):
    for (name, a_hi, b_hi, lbo) in (('no-swizzle', NS, NS, 288), ('sw128', SW128, SW128, 1)):
        for n_acc, ncols in ((1, 256), (1, 224), (1, 128), (1, 64), (2, 256), (2, 224), (2, 128), (4, 128), (4, 64)):
            for vary in (0, 1):
                iters = 4096
                _lib.check(lib.vd_tc_mma_rate(_lib.ptr(out), n_acc, ncols, iters, a_hi, b_hi, lbo, vary, grid, _lib.stream()), 'mma_rate')
                torch.cuda.synchronize()
                o = out.cpu()[:grid * 2].view(grid, 2).double()
                issue, total = o[:, 0].mean().item(), o[:, 1].mean().item()
                n = iters * n_acc
                print(f'grid={grid:3d} {name:10s} n_acc={n_acc} N={ncols:3d} vary={vary}: issue {issue / n:6.1f} cyc/MMA, '
                      f'complete {total / n:6.1f} cyc/MMA (ideal {ncols / 2:5.1f})', flush=True)
