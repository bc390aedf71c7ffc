"""Short driver for ncu: tensor-core embed of a batch of synthetic videos (conv0 -> conv1 -> conv2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200.networks import ConvNet3D  # noqa: E402
from video_distillation_b200.tc import TcConvNet3D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
SPLIT = len(sys.argv) > 2 and sys.argv[2] in ('x3', 'x2')  # f16x3 (split-fp16) pipeline instead of single-pass bf16; x2: two-product mode
X2 = len(sys.argv) > 2 and sys.argv[2] == 'x2'
T, HW = 16, 112
torch.manual_seed(0)
net = ConvNet3D(3, 50, 128, 3, 'relu', 'none', 'maxpooling', T, (HW, HW)).cuda()
tc = TcConvNet3D(T, HW, HW, 'cuda', max_batch=B, split=SPLIT, real_products=2 if X2 else 3)
f = net.features
tc.load_weights(f[0].weight, f[0].bias, f[3].weight, f[3].bias, f[6].weight, f[6].bias)
video = torch.randn(B, T, 3, HW, HW, device='cuda')
for _ in range(3):
    emb = tc.embed(video, frozen=X2)
torch.cuda.synchronize()
print('embed ok', tuple(emb.shape), float(emb.abs().mean()))
