"""What limits tcgen05.mma in ws_gemm_kernel?  Full-chip (148 CTAs) sustained MMA rate with moving A tiles and
moving / 128-byte-misaligned B windows (vd_tc_mma_rate2): cycles per MMA, ns per MMA, implied clock, TFLOP/s."""
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
NS = 8 | V1                   # no swizzle, SBO = 128 B
grid = 148
iters = int(os.environ.get('ITERS', 40000))


def run(name, n_acc=2, ncols=224, a_lbo=128, a_step=0, a_n=1, b_lbo=288, b_step=0, b_n=1, b_base=0, group=0, same=0, fill=0, gdelay=0):
    def go():
        _lib.check(lib.vd_tc_mma_rate2(_lib.ptr(out), n_acc, ncols, iters, NS, a_lbo, a_step, a_n, NS, b_lbo, b_step, b_n, b_base,
                                       group, same, fill, gdelay, grid, _lib.stream()), 'mma_rate2')
    go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    go()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cyc = out.cpu()[:grid * 2].view(grid, 2).double()[:, 1].mean().item()
    mmas = iters * n_acc
    print(f'{name:64s} N={ncols} acc={n_acc}: {cyc / mmas:6.1f} cyc/MMA (ideal {ncols / 2:5.1f})  {ms * 1e6 / mmas:6.1f} ns/MMA  '
          f'{cyc / (ms * 1e6):4.2f} GHz  {grid * mmas * 2 * 128 * ncols * 16 / (ms * 1e-3) / 1e12:6.0f} TFLOP/s', flush=True)


if os.environ.get('PROBE', 'delay') == 'delay':
    # how much scalar work between two stages does the MMA queue hide?  1 accumulator, 2 MMAs per loop step
    for group, label in ((6, '12 MMAs / stage'), (11, '22 MMAs / stage'), (24, '48 MMAs / stage')):
        for gdelay in (0, 100, 200, 300, 400, 600, 800, 1200):
            run(f'{label}, {gdelay} idle cycles between stages', same=1, a_step=256, a_n=11, b_step=1, b_n=7, group=group, gdelay=gdelay)
