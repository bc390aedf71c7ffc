"""Per-layer CUDA-event timing of the fused tensor-core embed: single-pass bf16 vs split-fp16 (f16x3).
usage: python scripts/x3_timing.py [B] [T] [HW]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200.networks import ConvNet3D  # noqa: E402
from video_distillation_b200.tc import TcConvNet3D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
HW = int(sys.argv[3]) if len(sys.argv) > 3 else 112
GF = {112: (2.832, 7.553, 0.617), 64: (0.462, 1.233, 0.077)}[HW]
GF = [g * T / (16 if HW == 112 else 8) for g in GF]
torch.manual_seed(0)
net = ConvNet3D(3, 50, 128, 3, 'relu', 'none', 'maxpooling', T, (HW, HW)).cuda()
video = torch.randn(B, T, 3, HW, HW, device='cuda')
ref = None
for split, products in ((False, 1), (True, 3), (True, 2)):
    tc = TcConvNet3D(T, HW, HW, 'cuda', max_batch=B, split=split, real_products=products if split else 3)
    f = net.features
    tc.load_weights(f[0].weight, f[0].bias, f[3].weight, f[3].bias, f[6].weight, f[6].bias)
    if products == 2:
        x0 = tc.pack_dataset(video)
        idx = torch.arange(B, device='cuda')
        run = lambda: tc.embed_resident(x0, idx)
    else:
        x0 = tc.pack_video(video)
        run = lambda: tc.embed_packed(x0, B)
    for _ in range(2):
        emb = run()
    torch.cuda.synchronize()
    tc.timing = []
    for _ in range(5):
        emb = run()
    torch.cuda.synchronize()
    per = {0: [], 1: [], 2: []}
    for layer, b, e0, e1, _ in tc.timing:
        per[layer].append(e0.elapsed_time(e1))
    name = ('f16x3' if products == 3 else 'f16x2') if split else 'bf16 '
    tot = 0.0
    for layer in range(3):
        ms = sorted(per[layer])[len(per[layer]) // 2]
        tot += ms
        print(f'{name} conv{layer}: {ms:8.3f} ms  {GF[layer] * B / ms:8.1f} useful TFLOP/s ({products * GF[layer] * B / ms:8.1f} issued-equivalent)')
    print(f'{name} embed of {B} videos: {tot:.3f} ms -> {B / tot * 1e3:.0f} videos/s')
    if ref is None:
        with torch.no_grad():
            n = min(B, 8)
            ref = net.embed(video[:n]) if hasattr(net, 'embed') else None
    if ref is not None:
        r = ((emb[:ref.shape[0]] - ref).norm() / ref.norm()).item()
        print(f'{name} embed vs fp32 CUDA-core kernels (first {ref.shape[0]} videos): relL2 {r:.2e}')
