"""world_size-2 gloo test of the class-sharded DM algebra (CPU, no GPU).

Each rank computes the loss / gradients of ITS classes with the CPU oracle, the product's
allreduce_sum_ combines them, and the result must equal the single-rank oracle iteration: this is
the N>1 path of DMS2DTrainer.step minus the CUDA kernels.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from oracle import synth


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _case():
    C, per, T, H, vpc, spc, dpc, batch_real = 4, 3, 4, 64, 1, 2, 2, 2
    videos = synth.hash_uniform((C * per, T, 3, H, H), 51)
    indices_class = [list(range(c * per, (c + 1) * per)) for c in range(C)]
    static_syn = synth.hash_uniform((C * spc, 3, H, H), 52)
    dyn = synth.hash_uniform((C, dpc, T, 1, H, H), 53)
    hal = synth.synth_hallucinator(5)
    params = synth.synth_convnet3d_params(60, num_classes=C)
    gen = torch.Generator().manual_seed(1)
    cd = torch.randint(2, (C * vpc,), generator=gen)
    cs = torch.randint(2, (C * vpc,), generator=gen)
    return dict(C=C, vpc=vpc, spc=spc, batch_real=batch_real, videos=videos, indices_class=indices_class,
                static_syn=static_syn, dyn=dyn, hal=hal, params=params, cd=cd, cs=cs)


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from video_distillation_b200.distill import allreduce_sum_, owned_classes
    k = _case()
    np.random.seed(11)
    real_idx = np.stack([oracle.sample_real_indices(k['indices_class'], c, k['batch_real']) for c in range(k['C'])])
    own = owned_classes(k['C'], rank, world)
    label, idx, didx, sidx = oracle.s2d_sample_indices(k['C'], k['vpc'], k['spc'], k['cd'], k['cs'])
    dyn = k['dyn'].clone().requires_grad_(True)
    w = k['hal']['encoder.weight'].clone().requires_grad_(True)
    b = k['hal']['encoder.bias'].clone().requires_grad_(True)
    sel = torch.tensor([c * k['vpc'] + i for c in own for i in range(k['vpc'])])
    image_syn = oracle.compose(k['static_syn'][sidx[sel]], dyn[label[sel], didx[sel]], w, b)
    loss = torch.tensor(0.0)
    for j, c in enumerate(own):
        er = oracle.convnet3d_embed(k['params'], k['videos'][torch.as_tensor(real_idx[c])]).detach()
        es = oracle.convnet3d_embed(k['params'], image_syn[j * k['vpc']:(j + 1) * k['vpc']])
        loss = loss + oracle.dm_loss(er, es)
    loss.backward()
    tensors = [dyn.grad, w.grad, b.grad, loss.detach().reshape(1).clone()]
    allreduce_sum_(tensors)
    if rank == 0:
        torch.save([t.clone() for t in tensors], out)
    dist.barrier()
    dist.destroy_process_group()


def test_class_sharded_dm_equals_single_rank(tmp_path):
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    k = _case()
    np.random.seed(11)
    r = oracle.dm_s2d_iteration(k['params'], k['static_syn'], k['dyn'], k['hal'], k['videos'], k['indices_class'],
                                vpc=k['vpc'], spc=k['spc'], batch_real=k['batch_real'], coin_dynamic=k['cd'],
                                coin_static=k['cs'])

    def rel(a, b):
        return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
    assert rel(got[0], r['grad_dynamic']) < 1e-5
    assert rel(got[1], r['grad_hal_weight']) < 1e-5 and rel(got[2], r['grad_hal_bias']) < 1e-5
    assert rel(got[3], r['loss'].reshape(1)) < 1e-6
