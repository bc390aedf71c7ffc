"""Which part of the conv kernels costs what?  Times each layer with parts of the kernel disabled
(VD_TC_DBG bitmask: 1 no pixel copies, 2 no weight copies, 4 no epilogue work).  Timing only — the
results of a run with a non-zero mask are garbage."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200.networks import ConvNet3D  # noqa: E402
from video_distillation_b200.tc import TcConvNet3D  # noqa: E402

B, T, HW = 592, 16, 112
torch.manual_seed(0)
net = ConvNet3D(3, 50, 128, 3, 'relu', 'none', 'maxpooling', T, (HW, HW)).cuda()
tc = TcConvNet3D(T, HW, HW, 'cuda', max_batch=B)
if os.environ.get('VD_EXP_ZERO'):
    for q in net.parameters():
        q.data.zero_()          # power experiment: all-zero operands toggle (almost) no datapath bits
f = net.features
tc.load_weights(f[0].weight, f[0].bias, f[3].weight, f[3].bias, f[6].weight, f[6].bias)
video = torch.randn(64, T, 3, HW, HW, device='cuda')
if os.environ.get('VD_EXP_ZERO'):
    video.zero_()
idx = torch.arange(B, device='cuda') % 64
x0 = tc.pack_dataset(video)
F = {0: 2.832e9, 1: 7.553e9, 2: 0.617e9}
ISSUED = {0: 1 / 0.626, 1: 1 / 0.875, 2: 1.0}
for dbg in [int(v) for v in os.environ.get("VD_EXP_DBG", "0,4,1,2,3,7").split(",")]:
    os.environ['VD_TC_DBG'] = str(dbg)
    for _ in range(2):
        tc.embed_resident(x0, idx)
    tc.timing = []
    reps = 3
    for _ in range(reps):
        tc.embed_resident(x0, idx)
    torch.cuda.synchronize()
    ms = {0: 0.0, 1: 0.0, 2: 0.0}
    for layer, b, a, e, _ in tc.timing:
        ms[layer] += a.elapsed_time(e) / reps
    tc.timing = None
    print(f'dbg={dbg}: ' + ' | '.join('conv%d %6.3f ms %5.0f TF/s useful %5.0f issued' % (k, ms[k], F[k] * B / ms[k] / 1e9, F[k] * ISSUED[k] * B / ms[k] / 1e9)
                                      for k in ms), flush=True)
os.environ['VD_TC_DBG'] = '0'
