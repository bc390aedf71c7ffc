"""Per-launch device times of the tensor-core backward chain (column GEMMs, col2im) for B synthetic videos."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200.networks import ConvNet3D  # noqa: E402
from video_distillation_b200.tc import TcConvNet3D  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

T, HW = 16, 112
for B in (50, 7):
    torch.manual_seed(0)
    net = ConvNet3D(3, 50, 128, 3, 'relu', 'none', 'maxpooling', T, (HW, HW)).cuda()
    tc = TcConvNet3D(T, HW, HW, 'cuda', max_batch=640)
    f = net.features
    tc.load_weights(f[0].weight, f[0].bias, f[3].weight, f[3].bias, f[6].weight, f[6].bias)
    video = torch.randn(B, T, 3, HW, HW, device='cuda', requires_grad=True)
    for _ in range(2):
        emb = tc.embed_autograd(video)
        emb.backward(torch.randn_like(emb))
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        emb = tc.embed_autograd(video)
        emb.backward(torch.randn_like(emb))
        torch.cuda.synchronize()
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    p = tc.plan
    print(f'--- B={B}: column buffers {B * p.col0_bytes_per_video / 1e9:.2f} / {B * p.col1_bytes_per_video / 1e9:.2f} / '
          f'{B * p.col2_bytes_per_video / 1e9:.3f} GB (conv0/1/2)')
    for e in evs:
        d = e.time_range.end - e.time_range.start
        if d > 5:
            print(f'{d:9.1f} us  {e.name[:90]}')
