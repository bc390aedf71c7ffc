"""Backward chain time vs videos per chunk (column buffers of 1-2 videos stay resident in the 126 MB L2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200.networks import ConvNet3D  # noqa: E402
from video_distillation_b200.tc import TcConvNet3D  # noqa: E402

T, HW, B = 16, 112, 50
torch.manual_seed(0)
net = ConvNet3D(3, 50, 128, 3, 'relu', 'none', 'maxpooling', T, (HW, HW)).cuda()
tc = TcConvNet3D(T, HW, HW, 'cuda', max_batch=640)
f = net.features
tc.load_weights(f[0].weight, f[0].bias, f[3].weight, f[3].bias, f[6].weight, f[6].bias)
video = torch.randn(B, T, 3, HW, HW, device='cuda')
emb, codes = tc.embed(video, want_codes=True)
g = torch.randn_like(emb)
for chunk in (64, 8, 4, 2, 1):
    tc.bwd_chunk = chunk
    tc._ws = {}
    for _ in range(2):
        tc.embed_backward(g, codes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        tc.embed_backward(g, codes)
    e1.record()
    torch.cuda.synchronize()
    print(f'bwd_chunk={chunk:3d}: {e0.elapsed_time(e1) / 5:.3f} ms for {B} videos', flush=True)
