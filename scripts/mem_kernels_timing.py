"""CUDA-event timing of the memory-bound kernels at the bench shape (50 synthetic videos 16x3x112x112), each against its
algorithmic bytes and, for the stencils, against the fp32 FMA peak.
usage: python scripts/mem_kernels_timing.py [n_videos]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200 import ops  # noqa: E402
from video_distillation_b200.tc import TcConvNet3D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 50
T, H, W, dpc, spc = 16, 112, 112, 2, 2
HBM, FMA = 6547.8e9, 36.0e12           # measured copy rate (MEASURED_PEAKS.json), fp32 FMA/s at 1.965 GHz (scripts/microbench/ffma_rate.cu)
dev = 'cuda'
torch.manual_seed(0)
static = torch.randn(B * spc, 3, H, W, device=dev)
dynamic = torch.randn(B, dpc, T, 1, H, W, device=dev, requires_grad=True)
w = (torch.randn(3, 4, 3, 3, 3, device=dev) * 0.1).requires_grad_(True)
b = torch.zeros(3, device=dev, requires_grad=True)
label = torch.arange(B, device=dev)
didx = torch.randint(2, (B,), device=dev)
sidx = spc * label + torch.randint(2, (B,), device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, e in ev:
        a.record(); fn(); e.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(e) for a, e in ev)
    return ts[len(ts) // 2]


def report(name, ms, nbytes, fmas=0):
    line = f'{name:34s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s = {nbytes / ms / 1e-3 / HBM:5.3f} of HBM copy rate'
    if fmas:
        line += f' | {fmas / ms / 1e9:6.2f} TFMA/s = {fmas / ms / 1e-3 / FMA:5.3f} of the fp32 FMA peak'
    print(line)


thw = T * H * W
out = ops.compose(static, dynamic, w, b, sidx, label, didx, unique_rows=True)
g = torch.randn_like(out)
fwd_bytes = B * 4 * (3 * H * W + thw + 3 * thw)
bwd_bytes = B * 4 * (3 * thw + thw + 3 * H * W + thw)
report('compose forward', timed(lambda: ops.compose(static, dynamic, w, b, sidx, label, didx, unique_rows=True)), fwd_bytes, B * thw * 81)


def bwd():
    dynamic.grad = None; w.grad = None; b.grad = None
    out.backward(g, retain_graph=True)


report('compose backward (fused) + grads', timed(bwd), bwd_bytes, B * thw * 162)

tc = TcConvNet3D(T, H, W, dev, split=True)
video = torch.randn(B, T, 3, H, W, device=dev)
report('pack_video_x3 (hi + lo)', timed(lambda: tc.pack_video(video)), B * (4 * 3 * thw + tc.x0_per))
report('pack_video_x3 (hi only)', timed(lambda: tc.pack_video(video, hi_only=True)), B * (4 * 3 * thw + tc.x0h_per))
emb = torch.randn(50, 64, 2048, device=dev)
report('class_mean (50 x 64 x 2048)', timed(lambda: ops.class_mean(emb)), emb.numel() * 4 + 50 * 2048 * 4)

# instancenorm + ReLU + avgpool between the convs of the IN / avgpool variant (networks.py:765-790): conv-0 output of 32 videos
xin = torch.randn(32, 64, 16, 56, 56, device=dev)
gam, bet = torch.ones(64, device=dev), torch.zeros(64, device=dev)
with torch.no_grad():
    nb = xin.numel() * 4 + xin.numel() // 8 * 4
    report('IN + ReLU + avgpool (one launch)', timed(lambda: ops.instancenorm_relu_avgpool(xin, gam, bet)), nb)
    report('IN + ReLU, then avgpool (pair)', timed(lambda: ops.avgpool3d_2(ops.instancenorm_relu(xin, gam, bet))), nb)
