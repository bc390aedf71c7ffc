"""GPU timeline of one DM+S2D iteration (torch.profiler / CUPTI): per-kernel time, launch count and the
idle gaps between kernels.  `--classes N` emulates the per-rank share of a sharded run (e.g. 7 of 50)."""
import argparse
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200.distill import DeviceDataset, DMS2DTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--classes', type=int, default=50)
ap.add_argument('--max-batch', type=int, default=640)
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--syn-mode', default='fused', choices=['fused', 'split', 'fp32'])
args = ap.parse_args()
C, T, HW, PER = args.classes, 16, 112, 72
dev = torch.device('cuda', 0)
labels = [c for c in range(C) for _ in range(PER)]
vids = torch.randn(C * PER, T, 3, HW, HW, device=dev)
ds = DeviceDataset.from_device_shard(vids, labels, C, dev, 0, 1)
torch.manual_seed(0)
tr = DMS2DTrainer(ds, num_classes=C, im_size=(HW, HW), frames=T, vpc=1, spc=2, dpc=2, batch_real=64, precision='bf16',
                  device=dev, init_on_device=True, max_batch=args.max_batch,
                  syn_on_tensor_cores={'fused': True, 'split': 'split', 'fp32': False}[args.syn_mode])
ds.prepack(tr.embedder.tc, extra_slots=len(tr.owned) * tr.vpc)
np.random.seed(0)
for i in range(3):
    tr.step(net_seed=i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(args.steps):
    tr.step(net_seed=10 + i)
e1.record()
torch.cuda.synchronize()
print(f'classes={C}: {e0.elapsed_time(e1) / args.steps:.3f} ms/step (events, no profiler)')
import time
t0 = time.perf_counter()
for i in range(args.steps):
    tr.step(net_seed=20 + i)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f'host enqueue time {1e3 * (t1 - t0) / args.steps:.3f} ms/step')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(args.steps):
        tr.step(net_seed=30 + i)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
busy = 0.0
for e in evs:
    d = e.time_range.end - e.time_range.start
    a = agg.setdefault(e.name[:70], [0, 0.0])
    a[0] += 1
    a[1] += d
    busy += d
span = evs[-1].time_range.end - evs[0].time_range.start
print(f'profiled span {span / 1e3 / args.steps:.3f} ms/step, kernel busy {busy / 1e3 / args.steps:.3f} ms/step, '
      f'idle {(span - busy) / 1e3 / args.steps:.3f} ms/step, {len(evs) / args.steps:.0f} GPU ops/step')
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f'{a[1] / 1e3 / args.steps:9.3f} ms/step {a[0] / args.steps:6.1f}x  {k}')
# largest gaps
gaps = []
for a, b in zip(evs[:-1], evs[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 20:
        gaps.append((g, a.name[:40], b.name[:40]))
gaps.sort(reverse=True)
print('largest gaps (us):')
for g, a, b in gaps[:25]:
    print(f'  {g:8.1f}  after {a}  before {b}')
