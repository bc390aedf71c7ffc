"""Does the TMA composer touch memory beyond its tensors?  Places the memories at the very END (and START) of a dedicated
cudaMalloc region and runs the forward / backward there."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200 import ops  # noqa: E402

dev = 'cuda'
for (C, dpc, T, H, W, spc) in [(3, 2, 4, 10, 12, 2), (3, 2, 5, 9, 8, 2), (4, 2, 16, 112, 112, 2), (4, 2, 8, 64, 64, 2)]:
    for where in ('end', 'start'):
        for trial in range(3):
            torch.cuda.empty_cache()
            n_s, n_d = C * spc * 3 * H * W, C * dpc * T * H * W
            seg = torch.empty(20 * 2**20 // 4, device=dev)             # one 20 MiB cudaMalloc (large pool)
            if where == 'end':
                d_flat = seg[seg.numel() - n_d:]
                s_flat = seg[seg.numel() - n_d - n_s - 128:seg.numel() - n_d - 128]
            else:
                d_flat = seg[:n_d]
                s_flat = seg[n_d + 128:n_d + 128 + n_s]
            d_flat.normal_(); s_flat.normal_()
            dyn = d_flat.view(C, dpc, T, 1, H, W).requires_grad_(True)
            sta = s_flat.view(C * spc, 3, H, W)
            w = (torch.randn(3, 4, 3, 3, 3, device=dev) * 0.1).requires_grad_(True)
            b = torch.zeros(3, device=dev, requires_grad=True)
            label = torch.arange(C, device=dev)
            didx = torch.full((C,), dpc - 1 if where == 'end' else 0, device=dev)
            sidx = spc * label + (spc - 1 if where == 'end' else 0)
            y = ops.compose(sta, dyn, w, b, sidx, label, didx, unique_rows=True)
            y.backward(torch.randn_like(y))
            torch.cuda.synchronize()
            del seg, d_flat, s_flat, dyn, sta, y
        print(f'ok: {(C, dpc, T, H, W)} memories at the {where} of a 20 MiB region', flush=True)
print('no fault')
