"""Is one f16x3 DM+S2D iteration bitwise reproducible?  Runs the same iteration twice and bisects differences."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_distillation_b200.distill import DeviceDataset, DMS2DTrainer  # noqa: E402
from video_distillation_b200 import tc as tcmod  # noqa: E402

C, T, HW, PER, BR = 10, 16, 112, 66, 64
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(11)
vids = torch.empty(C * PER, T, 3, HW, HW, device=dev).normal_(generator=g)
labels = [c for c in range(C) for _ in range(PER)]
ds = DeviceDataset.from_device_shard(vids, labels, C, dev, 0, 1)
stash = {}
orig = tcmod.TcConvNet3D._embed_backward_split


def wrapped(self, g_emb, codes):
    out = orig(self, g_emb, codes)
    stash.setdefault('runs', []).append(dict(g_emb=g_emb.clone(), codes=[c.clone() for c in codes], dvideo=out.clone()))
    return out


tcmod.TcConvNet3D._embed_backward_split = wrapped
outs = []
for rep in range(2):
    torch.manual_seed(5)
    tr = DMS2DTrainer(ds, num_classes=C, im_size=(HW, HW), frames=T, vpc=1, spc=2, dpc=2, batch_real=BR, lr_dynamic=1.0, lr_hal=1e-4,
                      precision='f16x3', device='cuda', init_on_device=True, max_batch=640)
    if ds.x0 is None:
        ds.prepack(tr.embedder.tc, extra_slots=C)
    np.random.seed(3)
    torch.cuda.manual_seed(103)
    loss = tr.step(net_seed=77)
    torch.cuda.synchronize()
    outs.append(dict(loss=loss.clone(), gd=tr.dynamic_syn.grad.clone(), gw=tr.hal.encoder.weight.grad.clone(), emb=tr.last['emb_syn'].clone(),
                     mean=tr.last['mean_real'].clone(), img=tr.last['image_syn'].clone()))
a, b = outs
for k in a:
    print(k, 'equal' if torch.equal(a[k], b[k]) else f'DIFF max {float((a[k] - b[k]).abs().max()):.3e} n {int((a[k] != b[k]).sum())}')
r0, r1 = stash['runs']
print('g_emb', torch.equal(r0['g_emb'], r1['g_emb']), 'codes', [torch.equal(x, y) for x, y in zip(r0['codes'], r1['codes'])],
      'dvideo', torch.equal(r0['dvideo'], r1['dvideo']), int((r0['dvideo'] != r1['dvideo']).sum()))
# the backward alone, twice, on identical inputs
net = tr.embedder.tc
d1 = orig(net, r1['g_emb'], r1['codes'])
d2 = orig(net, r1['g_emb'], r1['codes'])
print('backward alone repeatable:', torch.equal(d1, d2), int((d1 != d2).sum()))
from video_distillation_b200 import ops  # noqa: E402
p = net.plan
gy = ops.route_scatter_raw(r1['g_emb'].view(-1, 128, p.T3p, p.H3p, p.W3p), r1['codes'][2], (C, 128, p.T3, p.H3, p.W3), (2, 2, 2))
trio = net._trio
w = net._fp32_w
for layer, shape in ((2, None),):
    x1 = trio.dgrad(2, gy, w[2]).clone()
    x2 = trio.dgrad(2, gy, w[2]).clone()
    print('dgrad2 repeatable', torch.equal(x1, x2), int((x1 != x2).sum()))
gy1 = torch.randn(C, 128, p.T2, p.H2, p.W2, device=dev)
x1 = trio.dgrad(1, gy1, w[1]).clone(); x2 = trio.dgrad(1, gy1, w[1]).clone()
print('dgrad1 repeatable', torch.equal(x1, x2), int((x1 != x2).sum()))
gy0 = torch.randn(C, 64, p.T1, p.H1, p.W1, device=dev)
x1 = trio.dgrad(0, gy0, w[0]).clone(); x2 = trio.dgrad(0, gy0, w[0]).clone()
print('dgrad0 repeatable', torch.equal(x1, x2), int((x1 != x2).sum()))
for trial in range(3):
    x1 = trio.dgrad(1, gy1, w[1]).clone(); x2 = trio.dgrad(1, gy1, w[1]).clone()
    d = (x1 != x2)
    idx = d.nonzero()
    print('trial', trial, 'n diff', int(d.sum()), 'max abs', float((x1 - x2).abs().max()), 'rel', float((x1 - x2).abs().max() / x1.abs().max()))
    if idx.numel():
        print('  videos', idx[:, 0].unique().tolist(), 'channels', idx[:, 1].unique().tolist()[:20], 'frames', idx[:, 2].unique().tolist(),
              'rows', idx[:, 3].unique().tolist(), 'cols', idx[:, 4].unique().tolist())
        print('  first', idx[:8].tolist())
        print('  vals', x1[d][:6].tolist(), x2[d][:6].tolist())
# hypothesis: some output elements of the direct conv-1 dgrad are never written (torch.empty output keeps stale values)
import ctypes  # noqa: E402
from video_distillation_b200 import _lib  # noqa: E402
lib = _lib.lib()
sz = (ctypes.c_int64 * 3)()
lib.vd_tc_dgrad1_sizes(ctypes.byref(p), sz)
w0b = torch.empty(int(sz[1]), dtype=torch.uint8, device=dev); w1b = torch.empty(int(sz[2]), dtype=torch.uint8, device=dev)
dyp = torch.empty(C * int(sz[0]), dtype=torch.uint8, device=dev)
st = _lib.stream()
_lib.check(lib.vd_tc_pack_dgrad1_weights(_lib.ptr(w[1]), _lib.ptr(w0b), _lib.ptr(w1b), ctypes.byref(p), st))
_lib.check(lib.vd_tc_pack_dyp1(_lib.ptr(gy1), _lib.ptr(dyp), ctypes.byref(p), C, st))
for trial in range(3):
    gx = torch.full((C, 64, p.T1p, p.H1p, p.W1p), float('nan'), device=dev)
    _lib.check(lib.vd_tc_dgrad1(_lib.ptr(dyp), _lib.ptr(w0b), _lib.ptr(w1b), None, _lib.ptr(gx), ctypes.byref(p), C, st))
    torch.cuda.synchronize()
    nan = torch.isnan(gx)
    idx = nan.nonzero()
    print('NaN-prefilled output: unwritten elements', int(nan.sum()), 'of', gx.numel())
    if idx.numel():
        print('  videos', idx[:, 0].unique().tolist(), 'frames', idx[:, 2].unique().tolist(), 'rows', idx[:, 3].unique().tolist()[:10], 'cols', idx[:, 4].unique().tolist()[:10], 'ch', idx[:, 1].unique().tolist()[:10])
