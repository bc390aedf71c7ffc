"""Time the tensor-core embed of 640 videos for several weight-ring shapes (env overrides)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch
sys.path.insert(0, %r)
from video_distillation_b200.networks import ConvNet3D
from video_distillation_b200.tc import TcConvNet3D
B, T, HW = 592, 16, 112
torch.manual_seed(0)
net = ConvNet3D(3, 50, 128, 3, 'relu', 'none', 'maxpooling', T, (HW, HW)).cuda()
tc = TcConvNet3D(T, HW, HW, 'cuda', max_batch=B)
f = net.features
tc.load_weights(f[0].weight, f[0].bias, f[3].weight, f[3].bias, f[6].weight, f[6].bias)
video = torch.randn(64, T, 3, HW, HW, device='cuda')
idx = torch.arange(B, device='cuda') %% 64
x0 = tc.pack_dataset(video)
for _ in range(2):
    tc.embed_resident(x0, idx)
tc.timing = []
for _ in range(3):
    tc.embed_resident(x0, idx)
torch.cuda.synchronize()
ms = {0: 0.0, 1: 0.0, 2: 0.0}
for layer, b, a, e, _ in tc.timing:
    ms[layer] += a.elapsed_time(e) / 3
F = {0: 2.832e9, 1: 7.553e9, 2: 0.617e9}
print(' '.join('conv%%d %%6.2f ms %%6.0f TF/s' %% (k, ms[k], F[k] * B / ms[k] / 1e9) for k in ms))
''' % ROOT
configs = [dict(), dict(VD_TC_L1_G='5', VD_TC_L1_RW='3'), dict(VD_TC_L1_G='4', VD_TC_L1_RW='4'), dict(VD_TC_L1_G='4', VD_TC_L1_RW='3'),
           dict(VD_TC_L1_G='3', VD_TC_L1_RW='5'), dict(VD_TC_L2_G='3', VD_TC_L2_RW='4'), dict(VD_TC_L2_G='4', VD_TC_L2_RW='3'), dict()]
for cfg in configs:
    env = dict(os.environ, **cfg)
    r = subprocess.run([sys.executable, '-c', CODE], env=env, capture_output=True, text=True)
    print(cfg, '->', r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
