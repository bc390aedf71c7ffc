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


# ------------------------------------------------------------------------------------------ MTT batch sharding
def _mtt_case():
    C, T, H, vpc, spc, dpc, steps = 4, 8, 64, 1, 2, 2, 2
    like = synth.synth_convnet3d_params(80, num_classes=C)
    theta0 = oracle.flatten_params(like)
    target = theta0 + synth.hash_uniform(tuple(theta0.shape), 901, 2.0 ** -10)
    static_syn = synth.hash_uniform((C * spc, 3, H, H), 71)
    dyn = synth.hash_uniform((C, dpc, T, 1, H, H), 72)
    hal = synth.synth_hallucinator(7)
    gen = torch.Generator().manual_seed(3)
    perms = [torch.randperm(C * vpc, generator=gen) for _ in range(steps)]
    cds = [torch.randint(2, (C * vpc,), generator=gen) for _ in range(steps)]
    css = [torch.randint(2, (C * vpc,), generator=gen) for _ in range(steps)]
    return dict(C=C, T=T, H=H, vpc=vpc, spc=spc, like=like, theta0=theta0, target=target, static_syn=static_syn, dyn=dyn,
                hal=hal, perms=perms, cds=cds, css=css, syn_lr=torch.tensor(0.01))


def _mtt_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    import torch.nn.functional as F
    from oracle.mtt import mtt_sample_step_indices
    from video_distillation_b200.distill import allreduce_sum_, sharded_inner_grad
    k = _mtt_case()
    dyn = k['dyn'].clone().requires_grad_(True)
    w = k['hal']['encoder.weight'].clone().requires_grad_(True)
    b = k['hal']['encoder.bias'].clone().requires_grad_(True)
    syn_lr = k['syn_lr'].clone().requires_grad_(True)
    student = [k['theta0'].clone().requires_grad_(True)]
    for perm, cd, cs in zip(k['perms'], k['cds'], k['css']):
        label, idx, didx, sidx = mtt_sample_step_indices(perm, k['vpc'], k['spc'], cd, cs)
        sl = slice(rank, None, world)
        n_step = perm.numel()

        def shard_loss(theta, label=label, didx=didx, sidx=sidx, sl=sl, n_step=n_step):
            x = oracle.compose(k['static_syn'][sidx[sl]], dyn[label[sl], didx[sl]], w, b)
            logits = oracle.convnet3d_forward(oracle.unflatten_params(theta, k['like']), x, (k['H'], k['H']),
                                              dropout_mask=torch.ones(1))
            return F.cross_entropy(logits, label[sl].long(), reduction='sum') / n_step
        grad = sharded_inner_grad(student[-1], shard_loss, world)
        student.append(student[-1] - syn_lr * grad)
    n = k['theta0'].numel()
    grand = (F.mse_loss(student[-1], k['target'], reduction='sum') / n) / (F.mse_loss(k['theta0'], k['target'], reduction='sum') / n)
    grand.backward()
    grads = [dyn.grad, w.grad, b.grad]
    allreduce_sum_(grads)
    if rank == 0:
        torch.save([g.clone() for g in grads] + [syn_lr.grad.clone(), grand.detach().clone()], out)
    dist.barrier()
    dist.destroy_process_group()


def test_batch_sharded_mtt_equals_single_rank(tmp_path):
    """The product's differentiable collectives (sharded_inner_grad: _ReplicatedInput / _SumAcrossRanks) around the
    oracle student: 2 ranks, each holding half of every inner step's batch, reproduce the single-rank unroll and
    its second-order gradients."""
    out = str(tmp_path / 'mtt_r0.pt')
    mp.spawn(_mtt_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    k = _mtt_case()
    masks = [torch.ones(1) for _ in k['perms']]
    r = oracle.mtt_s2d_iteration(k['theta0'], k['target'], k['like'], k['static_syn'], k['dyn'], k['hal'], k['syn_lr'],
                                 vpc=k['vpc'], spc=k['spc'], perms=k['perms'], coins_dynamic=k['cds'], coins_static=k['css'],
                                 dropout_masks=masks, im_size=(k['H'], k['H']))

    def rel(a, b):
        return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
    assert rel(got[4], r['grand_loss']) < 1e-6
    assert rel(got[0], r['grad_dynamic']) < 1e-4, rel(got[0], r['grad_dynamic'])
    assert rel(got[1], r['grad_hal_weight']) < 1e-4 and rel(got[2], r['grad_hal_bias']) < 1e-4
    assert rel(got[3], r['grad_syn_lr']) < 1e-4


# ------------------------------------------------------------------ driver-level replica consistency (cli.py)
def _seed_worker(rank, world, port, out):
    """Two entropy-seeded processes follow the driver's protocol (cli.main_s2d): shared base seed, broadcast of the initial
    state, rank-0-only work that consumes random numbers (the evaluation block), per-iteration re-seeding.  Everything a rank
    draws afterwards — memories, expert walk, per-iteration indices — must be identical on both ranks."""
    import random
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from video_distillation_b200 import cli
    # processes start with different entropy, like two fresh torchrun workers
    torch.manual_seed(1234 + 99 * rank)
    np.random.seed(77 + rank)
    random.seed(5 + rank)
    base = cli._shared_seed(world, torch.device('cpu'))
    cli._seed_all(base)
    static_syn = torch.randn(6, 3, 8, 8)
    perm = np.random.permutation(20)
    shuffled = list(range(10))
    random.shuffle(shuffled)
    state = [static_syn + (0.0 if rank == 0 else 0.0)]
    cli._broadcast_state(state, world)
    draws = []
    for it in range(3):
        if rank == 0 and it in (0, 2):                  # "evaluation": only rank 0 consumes its generators
            torch.randn(100 + it)
            np.random.rand(7)
            random.random()
        cli._barrier(world)
        cli._seed_all(base + 7919 * (it + 1))
        draws.append((torch.randperm(12).tolist(), int(torch.randint(2, (1,))), int(np.random.randint(0, 50)), random.random()))
    out[rank] = dict(base=base, static=static_syn, perm=perm.tolist(), shuffled=shuffled, draws=draws,
                     agree=cli._replicas_agree(state, world))
    dist.destroy_process_group()


def test_driver_replicas_share_seed_and_stay_aligned():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_seed_worker, args=(world, port, out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert a['base'] == b['base']
    assert torch.equal(a['static'], b['static']) and a['perm'] == b['perm'] and a['shuffled'] == b['shuffled']
    assert a['draws'] == b['draws']
    assert a['agree'] == 0.0 and b['agree'] == 0.0
