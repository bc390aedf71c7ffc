"""Test tool (imports the CPU oracle, so it lives under tests/): gradient / loss error of every precision mode of DMS2DTrainer
against the exact fp32 path (same draws).  python tests/parity_modes.py on a GPU box -> profiles/r01_parity_modes.log."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from oracle import synth  # noqa: E402
from video_distillation_b200.distill import DeviceDataset, DMS2DTrainer  # noqa: E402
from video_distillation_b200.networks import ConvNet3D  # noqa: E402
from video_distillation_b200.utils import Conv3DNet  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


for (C, per, T, H, batch_real) in ((5, 10, 8, 64, 8), (4, 10, 8, 112, 8)):
    videos = synth.hash_uniform((C * per, T, 3, H, H), 51)
    labels = [c for c in range(C) for _ in range(per)]
    ds = DeviceDataset(videos, labels, C, 'cuda')
    params = synth.synth_convnet3d_params(60, num_classes=C) if H == 64 else None
    res = {}
    for name, prec, syn in (('fp32', 'fp32', True), ('bf16x3', 'bf16x3', True), ('bf16+split', 'bf16', 'split'), ('bf16 fused', 'bf16', True)):
        hal = Conv3DNet()
        hal.load_state_dict(synth.synth_hallucinator(5))
        tr = DMS2DTrainer(ds, num_classes=C, im_size=(H, H), frames=T, vpc=1, spc=2, dpc=2, batch_real=batch_real,
                          lr_dynamic=10.0, lr_hal=0.01, precision=prec, hal=hal, syn_on_tensor_cores=syn,
                          static_syn=synth.hash_uniform((C * 2, 3, H, H), 52), dynamic_syn=synth.hash_uniform((C, 2, T, 1, H, H), 53))
        np.random.seed(9)
        torch.manual_seed(0)
        net = ConvNet3D(3, C, 128, 3, 'relu', 'none', 'maxpooling', T, (H, H))
        if params is not None:
            net.load_state_dict(params)
        net = net.cuda().train()
        for p in net.parameters():
            p.requires_grad = False
        cd = torch.arange(C) % 2
        label, _, didx, sidx = oracle.s2d_sample_indices(C, 1, 2, cd, 1 - cd)
        loss = tr.step(net=net, indices=(label.cuda(), didx.cuda(), sidx.cuda()))
        res[name] = (loss.item(), tr.last['mean_real'].clone(), tr.last['emb_syn'].clone(), tr.dynamic_syn.grad.clone(),
                     tr.hal.encoder.weight.grad.clone())
    ref = res['fp32']
    print(f'--- C={C} T={T} H={H} batch_real={batch_real}: loss {ref[0]:.6e}')
    for name in ('bf16x3', 'bf16+split', 'bf16 fused'):
        r = res[name]
        print(f'{name:11s} loss {abs(r[0] - ref[0]) / abs(ref[0]):.2e} | real class means {rel(r[1], ref[1]):.2e} | syn embeddings {rel(r[2], ref[2]):.2e} | '
              f'd dynamic {rel(r[3], ref[3]):.2e} | d hallucinator {rel(r[4], ref[4]):.2e}', flush=True)
