"""End-to-end parity of the DM / DM+S2D / MTT+S2D iterations on the GPU against the committed golden
vectors (outputs of the live reference, tests/golden/*.npz) and the CPU oracle.

fp32 path: loss <= 1e-5 rel, gradients / updated parameters <= 1e-3 relL2 (the north_star gate;
observed ~1e-6 because routing is identical on these inputs).  bf16 tensor-core path for the real
embeddings: loss within 2e-2 (bf16 operand rounding).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def rel(a, b):
    a = a.double().cpu() if torch.is_tensor(a) else torch.as_tensor(a).double()
    b = b.double().cpu() if torch.is_tensor(b) else torch.as_tensor(b).double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def check_summary(t, sums, sample, tol=1e-3):
    from oracle import synth
    s, samp = synth.summarize(t.detach().cpu())
    assert rel(samp, sample) < tol, rel(samp, sample)
    assert abs(s[1] - sums[1]) <= 2 * tol * abs(sums[1]) + 1e-12


def real_set(C, per, T, H, seed):
    from oracle import synth
    videos = synth.hash_uniform((C * per, T, 3, H, H), seed)
    labels = [c for c in range(C) for _ in range(per)]
    return videos, labels


def net_from(params, C, T, H):
    from video_distillation_b200.networks import ConvNet3D
    net = ConvNet3D(3, C, 128, 3, 'relu', 'none', 'maxpooling', T, (H, H))
    net.load_state_dict(params)
    net = net.cuda().train()
    for p in net.parameters():
        p.requires_grad = False
    return net


def test_convnet3d_golden():
    """ConvNet3D.embed / forward (eval + train with the golden dropout mask) and the IN/avgpool variant."""
    from oracle import synth
    from video_distillation_b200.networks import ConvNet3D
    gold = np.load(os.path.join(GOLD, 'convnet3d.npz'))
    T, H = 8, 64
    params = synth.synth_convnet3d_params(1, num_classes=5)
    net = net_from(params, 5, T, H)
    x = synth.hash_uniform((3, T, 3, H, H), 11).cuda()
    net.eval()
    with torch.no_grad():
        assert rel(net.embed(x), gold['embed']) < 1e-5
        assert rel(net(x), gold['logits_eval']) < 1e-5
    # train mode with the reference's dropout mask: patch the dropout module with the golden mask
    mask = torch.from_numpy(gold['dropout_mask']).cuda()
    net.dropout = type('FixedDropout', (torch.nn.Module,), {'forward': lambda self, t: t * mask / 0.5})()
    with torch.no_grad():
        assert rel(net(x), gold['logits_train']) < 1e-5
    # instancenorm / avgpooling variant (north_star; networks.py:771-790)
    net2 = ConvNet3D(3, 5, 128, 3, 'relu', 'instancenorm', 'avgpooling', 8, (H, H))
    p2 = {}
    for d in range(3):
        p2[f'features.{4 * d}.weight'] = params[f'features.{3 * d}.weight']
        p2[f'features.{4 * d}.bias'] = params[f'features.{3 * d}.bias']
        c = p2[f'features.{4 * d}.weight'].shape[0]
        p2[f'features.{4 * d + 1}.weight'] = 1.0 + synth.hash_uniform((c,), 200 + d, 0.25)
        p2[f'features.{4 * d + 1}.bias'] = synth.hash_uniform((c,), 210 + d, 0.25)
    p2['logit.weight'], p2['logit.bias'] = params['logit.weight'], params['logit.bias']
    net2.load_state_dict(p2)
    net2 = net2.cuda()
    x8 = synth.hash_uniform((2, 8, 3, H, H), 12).cuda()
    with torch.no_grad():
        assert rel(net2.embed(x8), gold['embed_instancenorm_avgpool']) < 1e-4


def test_dm_baseline_golden():
    from oracle import synth
    from video_distillation_b200.distill import DeviceDataset, DMBaselineTrainer
    gold = np.load(os.path.join(GOLD, 'dm_baseline.npz'))
    C, per, T, H, ipc, batch_real, lr_img = 3, 4, 4, 64, 2, 3, 0.5
    videos, labels = real_set(C, per, T, H, 31)
    ds = DeviceDataset(videos, labels, C, 'cuda')
    tr = DMBaselineTrainer(ds, num_classes=C, im_size=(H, H), frames=T, ipc=ipc, batch_real=batch_real, lr_img=lr_img,
                           precision='fp32', image_syn=synth.hash_uniform((C * ipc, T, 3, H, H), 32))
    np.random.seed(7)
    for it in range(2):
        net = net_from(synth.synth_convnet3d_params(40 + it, num_classes=C), C, T, H)
        loss = tr.step(net=net)                                  # draws real indices from the numpy global RNG
        assert np.array_equal(tr.last['real_idx'], gold[f'real_idx{it}'])          # bit-exact class sampling
        assert rel(loss, gold[f'loss{it}']) < 1e-5
        assert rel(tr.last['emb_syn'].reshape(C, ipc, -1), gold[f'emb_syn{it}']) < 1e-5
        assert rel(tr.last['mean_real'], gold[f'emb_real_mean{it}']) < 1e-5
        check_summary(tr.image_syn.grad, gold[f'grad_sums{it}'], gold[f'grad_sample{it}'])
        check_summary(tr.image_syn, gold[f'syn_sums{it}'], gold[f'syn_sample{it}'], tol=1e-5)


def test_dm_baseline_bf16_tensor_core_path():
    """DM baseline (distill_baseline.py:334-356) with real AND synthetic embeds on tensor cores vs the exact path."""
    from oracle import synth
    from video_distillation_b200.distill import DeviceDataset, DMBaselineTrainer
    C, per, T, H, ipc, batch_real = 3, 4, 8, 64, 2, 3
    videos, labels = real_set(C, per, T, H, 31)
    out = {}
    for prec in ('fp32', 'bf16'):
        ds = DeviceDataset(videos, labels, C, 'cuda')
        tr = DMBaselineTrainer(ds, num_classes=C, im_size=(H, H), frames=T, ipc=ipc, batch_real=batch_real, lr_img=0.5,
                               precision=prec, image_syn=synth.hash_uniform((C * ipc, T, 3, H, H), 32))
        np.random.seed(7)
        net = net_from(synth.synth_convnet3d_params(40, num_classes=C), C, T, H)
        loss = tr.step(net=net)
        out[prec] = (loss.item(), tr.image_syn.grad.clone())
    assert abs(out['bf16'][0] - out['fp32'][0]) < 2e-2 * abs(out['fp32'][0]) + 1e-6
    # bf16 activations between layers flip some ReLU / pool routings against fp32: ~1.2e-1
    assert rel(out['bf16'][1], out['fp32'][1]) < 2.5e-1, rel(out['bf16'][1], out['fp32'][1])


def test_dm_s2d_synthetic_branch_modes():
    """Synthetic branch of DM+S2D on tensor cores: fused bf16 pipeline (throughput) and the split-bf16 trio
    (fp32 activations, routing as in fp32) against the exact fp32 path, real embeddings on tensor cores in all."""
    import oracle
    from oracle import synth
    from video_distillation_b200.distill import DeviceDataset, DMS2DTrainer
    from video_distillation_b200.utils import Conv3DNet
    C, per, T, H, batch_real = 3, 4, 8, 64, 3
    videos, labels = real_set(C, per, T, H, 51)
    ds = DeviceDataset(videos, labels, C, 'cuda')
    grads = {}
    for mode in (False, 'split', True):
        hal = Conv3DNet()
        hal.load_state_dict(synth.synth_hallucinator(5))
        tr = DMS2DTrainer(ds, num_classes=C, im_size=(H, H), frames=T, vpc=1, spc=2, dpc=2, batch_real=batch_real,
                          lr_dynamic=10.0, lr_hal=0.01, precision='bf16', hal=hal, syn_on_tensor_cores=mode,
                          static_syn=synth.hash_uniform((C * 2, 3, H, H), 52),
                          dynamic_syn=synth.hash_uniform((C, 2, T, 1, H, H), 53))
        np.random.seed(9)
        net = net_from(synth.synth_convnet3d_params(60, num_classes=C), C, T, H)
        label, _, didx, sidx = oracle.s2d_sample_indices(C, 1, 2, torch.tensor([0, 1, 1]), torch.tensor([1, 0, 1]))
        loss = tr.step(net=net, indices=(label.cuda(), didx.cuda(), sidx.cuda()))
        grads[mode] = (loss.item(), tr.dynamic_syn.grad.clone(), tr.hal.encoder.weight.grad.clone())
    e_split = rel(grads['split'][1], grads[False][1]), rel(grads['split'][2], grads[False][2])
    e_fused = rel(grads[True][1], grads[False][1]), rel(grads[True][2], grads[False][2])
    print('synthetic-branch gradient error vs exact fp32 (dynamic memory, hallucinator): split', e_split, 'fused', e_fused)
    assert abs(grads['split'][0] - grads[False][0]) < 1e-4 * abs(grads[False][0]) + 1e-7
    assert e_split[0] < 2e-2 and e_split[1] < 2e-2, e_split
    assert e_fused[0] < 3e-1 and e_fused[1] < 3e-1, e_fused


@pytest.mark.parametrize('tag,vpc,spc,dpc', [('v1', 1, 2, 2), ('v2', 2, 4, 4)])
def test_dm_s2d_golden(tag, vpc, spc, dpc):
    from oracle import synth, s2d_sample_indices
    from video_distillation_b200.distill import DeviceDataset, DMS2DTrainer
    from video_distillation_b200.utils import Conv3DNet
    gold = np.load(os.path.join(GOLD, 'dm_s2d.npz'))
    C, per, T, H, batch_real = 3, 4, 4, 64, 3
    videos, labels = real_set(C, per, T, H, 51)
    ds = DeviceDataset(videos, labels, C, 'cuda')
    hal = Conv3DNet()
    hal.load_state_dict(synth.synth_hallucinator(5))
    tr = DMS2DTrainer(ds, num_classes=C, im_size=(H, H), frames=T, vpc=vpc, spc=spc, dpc=dpc, batch_real=batch_real,
                      lr_dynamic=10.0, lr_hal=0.01, precision='fp32', hal=hal,
                      static_syn=synth.hash_uniform((C * spc, 3, H, H), 52),
                      dynamic_syn=synth.hash_uniform((C, dpc, T, 1, H, H), 53))
    np.random.seed(9)
    for it in range(2):
        k = f'{tag}_{it}'
        net = net_from(synth.synth_convnet3d_params(60 + it, num_classes=C), C, T, H)
        cd, cs = torch.from_numpy(gold[f'coin_dynamic_{k}']), torch.from_numpy(gold[f'coin_static_{k}'])
        label, idx, didx, sidx = s2d_sample_indices(C, vpc, spc, cd, cs)
        assert np.array_equal(didx.numpy(), gold[f'dynamic_idx_{k}']) and np.array_equal(sidx.numpy(), gold[f'static_idx_{k}'])
        loss = tr.step(net=net, indices=(label.cuda(), didx.cuda(), sidx.cuda()))
        assert np.array_equal(tr.last['real_idx'], gold[f'real_idx_{k}'])
        assert rel(loss, gold[f'loss_{k}']) < 1e-5
        assert rel(tr.last['emb_syn'].reshape(C, vpc, -1), gold[f'emb_syn_{k}']) < 1e-5
        check_summary(tr.dynamic_syn.grad, gold[f'grad_dynamic_sums_{k}'], gold[f'grad_dynamic_sample_{k}'])
        assert rel(tr.hal.encoder.weight.grad, gold[f'grad_hal_weight_{k}']) < 1e-3
        assert rel(tr.hal.encoder.bias.grad, gold[f'grad_hal_bias_{k}']) < 1e-3
        check_summary(tr.dynamic_syn, gold[f'dynamic_sums_{k}'], gold[f'dynamic_sample_{k}'], tol=1e-4)
        assert rel(tr.hal.encoder.weight, gold[f'hal_weight_{k}']) < 1e-5


def test_s2d_device_sampling_matches_reference_formula():
    """The trainer's own draws (torch device generator) follow distill_s2d_ms.py:402-406 bit-exactly."""
    from video_distillation_b200.distill import DMS2DTrainer
    tr = DMS2DTrainer.__new__(DMS2DTrainer)
    tr.C, tr.vpc, tr.spc, tr.device = 50, 5, 10, torch.device('cuda')
    torch.manual_seed(1234)
    label, didx, sidx = tr.sample_syn_indices()
    torch.manual_seed(1234)
    n = 250
    ref_label = torch.tensor(np.stack([np.ones(5) * i for i in range(0, 50)]), dtype=torch.long, device='cuda').view(-1)
    idx = torch.arange(0, n).cuda() % 5
    ref_d = 2 * idx + torch.randint(2, (n,), device='cuda')
    ref_s = 10 * ref_label + 2 * idx + torch.randint(2, (n,), device='cuda')
    assert torch.equal(label, ref_label) and torch.equal(didx, ref_d) and torch.equal(sidx, ref_s)


def test_dm_s2d_bf16_tensor_core_real_path():
    """Same iteration with the real embeddings on tcgen05 (bf16 operands): loss within bf16 tolerance,
    and the tensor-core embeddings within 1e-2 relL2 of the fp32 ones."""
    from oracle import synth
    from video_distillation_b200.distill import DeviceDataset, DMS2DTrainer
    from video_distillation_b200.utils import Conv3DNet
    C, per, T, H, batch_real = 3, 4, 8, 64, 3
    videos, labels = real_set(C, per, T, H, 51)
    ds = DeviceDataset(videos, labels, C, 'cuda')
    losses = {}
    for prec in ('fp32', 'bf16'):
        hal = Conv3DNet()
        hal.load_state_dict(synth.synth_hallucinator(5))
        tr = DMS2DTrainer(ds, num_classes=C, im_size=(H, H), frames=T, vpc=1, spc=2, dpc=2, batch_real=batch_real,
                          lr_dynamic=10.0, lr_hal=0.01, precision=prec, hal=hal,
                          static_syn=synth.hash_uniform((C * 2, 3, H, H), 52),
                          dynamic_syn=synth.hash_uniform((C, 2, T, 1, H, H), 53))
        np.random.seed(9)
        torch.manual_seed(3)
        net = net_from(synth.synth_convnet3d_params(60, num_classes=C), C, T, H)
        losses[prec] = (tr.step(net=net).item(), tr.last['mean_real'].clone())
    assert rel(losses['bf16'][1], losses['fp32'][1]) < 1e-2
    assert abs(losses['bf16'][0] - losses['fp32'][0]) < 2e-2 * abs(losses['fp32'][0]) + 1e-6


def test_dm_s2d_bf16x3_parity_mode_on_tensor_cores():
    """precision='bf16x3': real and synthetic embeds on the unfused tensor-core trio with every primitive on hi / lo operand
    pairs and fp32 activations, against the exact fp32 path: loss / class means within 1e-4 (measured 5e-5 / 3e-5), UNconditioned
    gradients within 5e-3 (measured 4.4e-4 / 1.4e-4; with single-pass bf16 dgrads they were 3e-3 / 1e-3)."""
    import oracle
    from oracle import synth
    from video_distillation_b200.distill import DeviceDataset, DMS2DTrainer
    from video_distillation_b200.utils import Conv3DNet
    C, per, T, H, batch_real = 3, 4, 8, 64, 3
    videos, labels = real_set(C, per, T, H, 51)
    ds = DeviceDataset(videos, labels, C, 'cuda')
    res = {}
    for prec in ('fp32', 'bf16x3'):
        hal = Conv3DNet()
        hal.load_state_dict(synth.synth_hallucinator(5))
        tr = DMS2DTrainer(ds, num_classes=C, im_size=(H, H), frames=T, vpc=1, spc=2, dpc=2, batch_real=batch_real,
                          lr_dynamic=10.0, lr_hal=0.01, precision=prec, hal=hal,
                          static_syn=synth.hash_uniform((C * 2, 3, H, H), 52),
                          dynamic_syn=synth.hash_uniform((C, 2, T, 1, H, H), 53))
        np.random.seed(9)
        net = net_from(synth.synth_convnet3d_params(60, num_classes=C), C, T, H)
        label, _, didx, sidx = oracle.s2d_sample_indices(C, 1, 2, torch.tensor([0, 1, 1]), torch.tensor([1, 0, 1]))
        loss = tr.step(net=net, indices=(label.cuda(), didx.cuda(), sidx.cuda()))
        res[prec] = (loss.item(), tr.last['mean_real'].clone(), tr.dynamic_syn.grad.clone(), tr.hal.encoder.weight.grad.clone())
    a, b = res['bf16x3'], res['fp32']
    errs = (abs(a[0] - b[0]) / abs(b[0]), rel(a[1], b[1]), rel(a[2], b[2]), rel(a[3], b[3]))
    print('bf16x3 vs fp32 (loss, real class means, dynamic-memory grad, hallucinator grad):', errs)
    assert errs[0] < 1e-4 and errs[1] < 1e-4, errs
    assert errs[2] < 5e-3 and errs[3] < 5e-3, errs


@pytest.mark.parametrize('precision', ['fp32', 'bf16', 'bf16x3'])
def test_mtt_s2d_golden(precision):
    """Unrolled student with second-order autograd through our conv trio vs the reference: exact fp32 kernels,
    the tensor-core trio (split fprop, single-pass bf16 dgrad / wgrad, every first- and second-order conv term) and the
    parity-grade tensor-core trio 'bf16x3' (every primitive on hi / lo operand pairs)."""
    from oracle import synth
    from video_distillation_b200.distill import MTTS2DTrainer
    from video_distillation_b200.networks import ConvNet3D
    from video_distillation_b200.reparam_module import ReparamModule
    from video_distillation_b200.utils import Conv3DNet
    gold = np.load(os.path.join(GOLD, 'mtt_s2d.npz'))
    C, T, H, vpc, spc, dpc, syn_steps = 3, 8, 64, 1, 2, 2, 2
    hal = Conv3DNet()
    hal.load_state_dict(synth.synth_hallucinator(7))
    tr = MTTS2DTrainer(num_classes=C, im_size=(H, H), frames=T, vpc=vpc, spc=spc, dpc=dpc, syn_steps=syn_steps,
                       lr_teacher=0.01, hal=hal, precision=precision, static_syn=synth.hash_uniform((C * spc, 3, H, H), 71),
                       dynamic_syn=synth.hash_uniform((C, dpc, T, 1, H, H), 72))
    start = synth.synth_convnet3d_params(80, num_classes=C)
    target = {k: v + synth.hash_uniform(tuple(v.shape), 900 + i, 2.0 ** -10) for i, (k, v) in enumerate(start.items())}
    student = ReparamModule(ConvNet3D(3, C, 128, 3, 'relu', 'none', 'maxpooling', T, (H, H)).cuda())
    # replay the reference's draws: patch the device RNG calls with the golden values
    perms, cds, css = gold['perms'], gold['coins_dynamic'], gold['coins_static']
    masks = torch.from_numpy(gold['dropout_masks']).cuda()
    state = {'perm': 0, 'coin': 0, 'mask': 0}
    real_randperm, real_randint = torch.randperm, torch.randint

    def fake_randperm(n, device=None, **kw):
        v = torch.from_numpy(perms[state['perm']]).to(device)
        state['perm'] += 1
        return v

    def fake_randint(high, size, device=None, **kw):
        i = state['coin']
        state['coin'] += 1
        return torch.from_numpy((cds if i % 2 == 0 else css)[i // 2]).to(device)

    class FixedDropout(torch.nn.Module):
        def forward(self, t):
            m = masks[state['mask']]
            state['mask'] += 1
            return t * m / 0.5
    student.module.dropout = FixedDropout()
    torch.randperm, torch.randint = fake_randperm, fake_randint
    try:
        grand = tr.step(list(start.values()), list(target.values()), student_net=student)
    finally:
        torch.randperm, torch.randint = real_randperm, real_randint
    assert rel(tr.last['param_dist'], gold['param_dist']) < 1e-5
    if precision == 'fp32':
        assert rel(grand, gold['grand_loss']) < 1e-5
        assert rel(tr.hal.encoder.weight.grad, gold['grad_hal_weight']) < 1e-3
        assert rel(tr.hal.encoder.bias.grad, gold['grad_hal_bias']) < 1e-3
        assert rel(tr.syn_lr.grad, gold['grad_syn_lr']) < 1e-3
        check_summary(tr.dynamic_syn.grad, gold['grad_dynamic_sums'], gold['grad_dynamic_sample'])
    else:
        from oracle import synth as _synth
        errs = dict(grand=rel(grand, gold['grand_loss']), hal_w=rel(tr.hal.encoder.weight.grad, gold['grad_hal_weight']),
                    hal_b=rel(tr.hal.encoder.bias.grad, gold['grad_hal_bias']), lr=rel(tr.syn_lr.grad, gold['grad_syn_lr']),
                    dynamic_sample=rel(_synth.summarize(tr.dynamic_syn.grad.detach().cpu())[1], gold['grad_dynamic_sample']))
        print(f'mtt {precision} tensor-core trio vs reference:', errs)
        assert errs['grand'] < 1e-5, errs
        if precision == 'bf16':
            # split fprop + single-pass bf16 dgrad / wgrad: hallucinator / lr gradients within 1e-2, dynamic memory 2e-2 (observed 1.2e-2)
            assert errs['hal_w'] < 1e-2 and errs['hal_b'] < 1e-2 and errs['lr'] < 1e-2, errs
            check_summary(tr.dynamic_syn.grad, gold['grad_dynamic_sums'], gold['grad_dynamic_sample'], tol=2e-2)
        else:
            # every primitive split: the 1e-3 gate of north_star on tensor cores (observed 1.5e-5 / 1.4e-5 / 3.1e-5 / 7e-6)
            assert errs['hal_w'] < 2e-4 and errs['hal_b'] < 2e-4 and errs['lr'] < 2e-4, errs
            check_summary(tr.dynamic_syn.grad, gold['grad_dynamic_sums'], gold['grad_dynamic_sample'], tol=2e-4)


def test_streamed_real_batch_in_any_row_order():
    """step(real_batch=..., real_batch_index=perm): the uploaded rows may come in another order (merged host ranges, uint8
    frames) — loss, class means and gradients are bit-identical to the resident-dataset step."""
    from oracle import synth
    from video_distillation_b200.distill import DeviceDataset, DMS2DTrainer
    from video_distillation_b200.utils import Conv3DNet
    C, per, T, H, batch_real = 3, 5, 8, 64, 4
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    frames = torch.randint(0, 256, (C * per, T, 3, H, H), dtype=torch.uint8, generator=torch.Generator().manual_seed(4))
    videos = ((frames.float() / 255.0) - torch.tensor(mean).view(1, 1, 3, 1, 1)) / torch.tensor(std).view(1, 1, 3, 1, 1)
    labels = [c for c in range(C) for _ in range(per)]
    ds = DeviceDataset(videos, labels, C, 'cuda')
    outs = []
    for mode in ('resident', 'streamed'):
        hal = Conv3DNet()
        hal.load_state_dict(synth.synth_hallucinator(5))
        tr = DMS2DTrainer(ds, num_classes=C, im_size=(H, H), frames=T, vpc=1, spc=2, dpc=2, batch_real=batch_real, lr_dynamic=10.0,
                          lr_hal=0.01, precision='bf16', hal=hal, static_syn=synth.hash_uniform((C * 2, 3, H, H), 52),
                          dynamic_syn=synth.hash_uniform((C, 2, T, 1, H, H), 53))
        net = net_from(synth.synth_convnet3d_params(60, num_classes=C), C, T, H)
        np.random.seed(9)
        torch.manual_seed(3)
        idx = tr.sample_syn_indices()
        if mode == 'resident':
            loss = tr.step(net=net, indices=idx)
        else:
            real_idx = ds.sample_all_classes(batch_real)
            loc = real_idx.reshape(-1)
            order = np.argsort(loc, kind='stable')
            perm = np.empty_like(order)
            perm[order] = np.arange(order.size)
            tr.embedder.tc.set_normalization(mean, std)
            stage = frames[torch.from_numpy(loc[order])].cuda()              # uint8 rows in ascending host order
            loss = tr.step(net=net, indices=idx, real_idx=real_idx, real_batch=stage, real_batch_index=torch.from_numpy(perm).cuda())
        outs.append((loss.clone(), tr.last['mean_real'].clone(), tr.dynamic_syn.grad.clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


@pytest.mark.parametrize('precision', ['fp32', 'bf16x3'])
def test_mtt_baseline_golden(precision):
    """MTT on leaf synthetic videos (distill_baseline.py:196-272, MTTBaselineTrainer) against the live-reference golden:
    ipc = 2 with batch_syn = 4, so the randperm is split and the chunks are consumed last-first like the reference's pop()."""
    from oracle import synth
    from video_distillation_b200.distill import MTTBaselineTrainer
    from video_distillation_b200.networks import ConvNet3D
    from video_distillation_b200.reparam_module import ReparamModule
    gold = np.load(os.path.join(GOLD, 'mtt_baseline.npz'))
    C, T, H, ipc, syn_steps, batch_syn = 3, 8, 64, 2, 2, 4
    tr = MTTBaselineTrainer(num_classes=C, im_size=(H, H), frames=T, ipc=ipc, syn_steps=syn_steps, lr_img=1.0, lr_lr=1e-5,
                            lr_teacher=0.01, train_lr=True, batch_syn=batch_syn, image_syn=synth.hash_uniform((C * ipc, T, 3, H, H), 91),
                            precision=precision)
    start = synth.synth_convnet3d_params(81, num_classes=C)
    target = {k: v + synth.hash_uniform(tuple(v.shape), 950 + i, 2.0 ** -10) for i, (k, v) in enumerate(start.items())}
    student = ReparamModule(ConvNet3D(3, C, 128, 3, 'relu', 'none', 'maxpooling', T, (H, H)).cuda())
    masks = [torch.from_numpy(gold['dropout_mask_0']).cuda(), torch.from_numpy(gold['dropout_mask_1']).cuda()]
    state = {'perm': 0, 'mask': 0}
    real_randperm = torch.randperm

    def fake_randperm(n, **kw):
        v = torch.from_numpy(gold['perms'][state['perm']])
        state['perm'] += 1
        return v

    class FixedDropout(torch.nn.Module):
        def forward(self, t):
            m = masks[state['mask']]
            state['mask'] += 1
            return t * m / 0.5
    student.module.dropout = FixedDropout()
    torch.randperm = fake_randperm
    try:
        grand = tr.step(list(start.values()), list(target.values()), student_net=student)
    finally:
        torch.randperm = real_randperm
    assert state['perm'] == 1                                              # one permutation feeds both steps
    assert [d.tolist() for d in tr.last['draws']] == [gold['used_0'].tolist(), gold['used_1'].tolist()]
    assert rel(tr.last['param_dist'], gold['param_dist']) < 1e-5
    assert rel(grand, gold['grand_loss']) < 1e-5
    from oracle import synth as _synth
    print(f'baseline MTT [{precision}] vs reference: lr grad', rel(tr.syn_lr.grad, gold['grad_syn_lr']), 'image grad sample',
          rel(_synth.summarize(tr.image_syn.grad.detach().cpu())[1], gold['grad_image_sample']))
    assert rel(tr.syn_lr.grad, gold['grad_syn_lr']) < 1e-3
    check_summary(tr.image_syn.grad, gold['grad_image_sums'], gold['grad_image_sample'])
    rows = tr.image_syn.grad.flatten(1).abs().sum(1) > 0
    assert rows.cpu().tolist() == gold['rows_with_grad'].tolist()
