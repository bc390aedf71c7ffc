"""Kernel-level breakdown (CUPTI via torch.profiler) of the f16x3 embed backward for B synthetic videos."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200.networks import ConvNet3D  # noqa: E402
from video_distillation_b200.tc import TcConvNet3D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 50
T, HW = 16, 112
torch.manual_seed(0)
net = ConvNet3D(3, 50, 128, 3, 'relu', 'none', 'maxpooling', T, (HW, HW)).cuda()
tc = TcConvNet3D(T, HW, HW, 'cuda', max_batch=640, split=True)
f = net.features
tc.load_weights(f[0].weight, f[0].bias, f[3].weight, f[3].bias, f[6].weight, f[6].bias)
video = torch.randn(B, T, 3, HW, HW, device='cuda')
emb, codes = tc.embed(video, want_codes=True)
g = torch.randn_like(emb) * 1e-3
for _ in range(2):
    tc.embed_backward(g, codes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    tc.embed_backward(g, codes)
e1.record()
torch.cuda.synchronize()
print(f'embed_backward of {B} videos: {e0.elapsed_time(e1) / 3:.3f} ms')
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tc.embed_backward(g, codes)
    torch.cuda.synchronize()
dur = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        d = dur.setdefault(e.name[:90], [0, 0.0])
        d[0] += 1
        d[1] += (e.time_range.end - e.time_range.start) * 1e-3
tot = sum(v[1] for v in dur.values())
print(f'sum of kernel time {tot:.3f} ms')
for k, v in sorted(dur.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f'{v[1]:8.3f} ms  {v[0]:4d}x  {k}')
