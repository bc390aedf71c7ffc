"""Repeatability of the direct conv-1 dgrad (vd_tc_dgrad1) from a cold process: magnitude and location of differences."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200.tc_trio import TcTrio  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device('cuda', 0)
trio = TcTrio(16, 112, 112, dev)
p = trio.plan
torch.manual_seed(0)
w = (torch.randn(128, 64, 3, 7, 7, device=dev) * 0.02)
gy = torch.randn(B, 128, p.T2, p.H2, p.W2, device=dev)
ref = None
for trial in range(12):
    if trial % 3 == 0:
        junk = torch.randn(64 * 1024 * 1024, device=dev)      # churn the allocator between groups of calls
        del junk
    x = trio.dgrad(1, gy, w).clone()
    if ref is None:
        ref = x
        continue
    d = x != ref
    n = int(d.sum())
    print('trial', trial, 'n diff', n, 'max abs', float((x - ref).abs().max()), 'scale', float(ref.abs().max()))
    if n:
        idx = d.nonzero()
        print('  videos', idx[:, 0].unique().tolist(), 'frames', idx[:, 2].unique().tolist(), 'rows', idx[:, 3].unique().tolist(),
              'cols', idx[:, 4].unique().tolist()[:12], 'n channels', idx[:, 1].unique().numel())
        print('  sample', idx[:5].tolist(), x[d][:5].tolist(), ref[d][:5].tolist())
