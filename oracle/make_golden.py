#!/usr/bin/env python
"""Pin the oracle against the LIVE reference and write tests/golden/*.npz.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

For every case it (1) runs the unmodified reference modules (networks.ConvNet3D via
utils.get_network, utils.Conv3DNet, reparam_module.ReparamModule, torch.optim.SGD) through a
verbatim transcription of the reference loop bodies, (2) runs the oracle restatement on the
same inputs, (3) asserts they agree (bit-exact for integer work, <=1e-5 rel for fp32), and
(4) stores the REFERENCE outputs as golden vectors.  Inputs are hash-generated
(oracle/synth.py) and therefore not stored.
"""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('VD_REFERENCE', '/root/reference')
GOLD = os.path.join(REPO, 'tests', 'golden')


def import_reference():
    """Import the reference's top-level modules without shadowing by this repo's drop-ins."""
    saved = list(sys.path)
    sys.path[:] = [REF] + [p for p in saved if os.path.abspath(p or '.') != REPO]
    for name in ('networks', 'utils', 'reparam_module', 'distill_utils'):
        sys.modules.pop(name, None)
    import networks as ref_networks          # noqa
    import utils as ref_utils                # noqa
    import reparam_module as ref_reparam     # noqa
    mods = (ref_networks, ref_utils, ref_reparam)
    for name in ('networks', 'utils', 'reparam_module', 'distill_utils'):
        sys.modules.pop(name, None)
    sys.path[:] = saved
    return mods


ref_networks, ref_utils, ref_reparam = import_reference()
sys.path.insert(0, REPO)
import oracle                                 # noqa: E402
from oracle import synth                      # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def ref_get_network(seed, channel, num_classes, im_size):
    """utils.get_network('ConvNet3D') with the wall clock patched so that utils.py:519 seeds
    torch with exactly ``seed``."""
    real_time = time.time
    ref_utils.time.time = lambda: seed / 1000.0 + 1e-7
    try:
        net = ref_utils.get_network('ConvNet3D', channel, num_classes, im_size, dist=False)
    finally:
        ref_utils.time.time = real_time
    return net


def load_params(net, params):
    sd = {k: v.clone() for k, v in params.items()}
    net.load_state_dict(sd)
    return net


def case_init():
    """Seeded init parity + parameter order/offsets."""
    seed = 4321
    net = ref_get_network(seed, 3, 50, (112, 112))
    ref_sd = net.state_dict()
    mine = oracle.init_convnet3d(seed, 3, 50)
    assert list(ref_sd.keys()) == list(mine.keys()) == oracle.convnet3d_param_names()
    for k in ref_sd:
        assert torch.equal(ref_sd[k], mine[k]), k
    rp = ref_reparam.ReparamModule(net)
    flat = oracle.flatten_params(mine)
    assert torch.equal(rp.flat_param.detach(), flat)
    offs = np.cumsum([0] + [v.numel() for v in mine.values()])
    ssum, samp = synth.summarize(flat, 9973)
    hal = ref_utils.Conv3DNet()
    torch.random.manual_seed(77)
    hal = ref_utils.Conv3DNet()
    mh = oracle.init_hallucinator(77)
    assert torch.equal(hal.encoder.weight.detach(), mh['encoder.weight'])
    assert torch.equal(hal.encoder.bias.detach(), mh['encoder.bias'])
    np.savez_compressed(os.path.join(GOLD, 'init.npz'), seed=seed, offsets=offs, flat_sums=ssum,
                        flat_sample=samp, hal_seed=77, hal_weight=mh['encoder.weight'].numpy(),
                        hal_bias=mh['encoder.bias'].numpy())
    print('init ok: P =', flat.numel(), 'offsets', offs.tolist())


def case_convnet3d():
    """embed / forward(eval) / forward(train, replayed dropout mask) and all intermediates."""
    T, H = 8, 64
    params = synth.synth_convnet3d_params(1, num_classes=5)
    net = ref_networks.ConvNet3D(3, 5, 128, 3, 'relu', 'none', 'maxpooling', frames=T, im_size=(H, H))
    load_params(net, params)
    x = synth.hash_uniform((3, T, 3, H, H), 11)
    net.eval()
    with torch.no_grad():
        e_ref = net.embed(x)
        f_ref = net(x)
    e = oracle.convnet3d_embed(params, x)
    f = oracle.convnet3d_forward(params, x, (H, H))
    assert torch.equal(e_ref, e) and torch.equal(f_ref, f)
    # train mode: replay the dropout mask drawn by nn.Dropout on the (B,128,t,1,1) tensor
    net.train()
    torch.manual_seed(5)
    with torch.no_grad():
        ft_ref = net(x)
    feat = oracle.convnet3d_features(params, x.permute(0, 2, 1, 3, 4))
    pooled = torch.nn.functional.avg_pool3d(feat, (2, 1, 1), (1, 1, 1))
    torch.manual_seed(5)
    mask = (torch.nn.functional.dropout(torch.ones_like(pooled), 0.5, True) > 0).float()
    ft = oracle.convnet3d_forward(params, x, (H, H), dropout_mask=mask)
    assert torch.equal(ft_ref, ft)
    _, inter = oracle.convnet3d_features(params, x.permute(0, 2, 1, 3, 4), return_intermediates=True)
    out = dict(embed=e_ref.numpy(), logits_eval=f_ref.numpy(), logits_train=ft_ref.numpy(),
               dropout_mask=mask.numpy())
    for kind, d, t in inter:
        s, samp = synth.summarize(t)
        out[f'{kind}{d}_sums'] = s
        out[f'{kind}{d}_sample'] = samp
    # instancenorm / avgpooling variant (north_star names it; networks.py:771-790)
    params_in = dict(params)
    net2 = ref_networks.ConvNet3D(3, 5, 128, 3, 'relu', 'instancenorm', 'avgpooling', frames=8, im_size=(H, H))
    sd2 = net2.state_dict()
    names2 = oracle.convnet3d_param_names(3, 'instancenorm')
    assert list(sd2.keys()) == names2
    p2 = {}
    for d in range(3):
        p2[f'features.{4 * d}.weight'] = params[f'features.{3 * d}.weight']
        p2[f'features.{4 * d}.bias'] = params[f'features.{3 * d}.bias']
        c = p2[f'features.{4 * d}.weight'].shape[0]
        p2[f'features.{4 * d + 1}.weight'] = 1.0 + synth.hash_uniform((c,), 200 + d, 0.25)
        p2[f'features.{4 * d + 1}.bias'] = synth.hash_uniform((c,), 210 + d, 0.25)
    p2['logit.weight'] = params['logit.weight']
    p2['logit.bias'] = params['logit.bias']
    p2 = {k: p2[k] for k in names2}
    load_params(net2, p2)
    x8 = synth.hash_uniform((2, 8, 3, H, H), 12)
    with torch.no_grad():
        e2_ref = net2.embed(x8)
    e2 = oracle.convnet3d_embed(p2, x8, net_norm='instancenorm', net_pooling='avgpooling')
    assert torch.equal(e2_ref, e2)
    out['embed_instancenorm_avgpool'] = e2_ref.numpy()
    np.savez_compressed(os.path.join(GOLD, 'convnet3d.npz'), **out)
    print('convnet3d ok: embed', tuple(e_ref.shape), 'logits', tuple(f_ref.shape))


def case_composer():
    T, H = 4, 16
    hal = synth.synth_hallucinator(3)
    ref = ref_utils.Conv3DNet()
    ref.load_state_dict({k: v.clone() for k, v in hal.items()})
    static = synth.hash_uniform((5, 3, H, H), 21)
    dynamic = synth.hash_uniform((5, T, 1, H, H), 22)
    with torch.no_grad():
        y_ref = ref(static, dynamic)
    y = oracle.compose(static, dynamic, hal['encoder.weight'], hal['encoder.bias'])
    assert torch.equal(y_ref, y)
    # backward
    s = static.clone().requires_grad_(True)
    d = dynamic.clone().requires_grad_(True)
    gy = synth.hash_uniform(tuple(y_ref.shape), 23)
    out = ref(s, d)
    out.backward(gy)
    gw, gb = ref.encoder.weight.grad.clone(), ref.encoder.bias.grad.clone()
    np.savez_compressed(os.path.join(GOLD, 'composer.npz'), y=y_ref.numpy(), grad_static=s.grad.numpy(),
                        grad_dynamic=d.grad.numpy(), grad_weight=gw.numpy(), grad_bias=gb.numpy())
    print('composer ok', tuple(y_ref.shape))


class _SynthSet:
    """dst_train stand-in: __getitem__ -> (video, label), .labels (distill_baseline.py:74-90)."""
    def __init__(self, videos, labels):
        self.videos, self.labels = videos, labels

    def __len__(self):
        return len(self.labels)

    def __getitem__(self, i):
        return self.videos[i], self.labels[i]


def _real_set(C, per, T, H, seed):
    videos = synth.hash_uniform((C * per, T, 3, H, H), seed)
    labels = [c for c in range(C) for _ in range(per)]
    indices_class = [[] for _ in range(C)]
    for i, lab in enumerate(labels):
        indices_class[lab].append(i)
    return videos, labels, indices_class


def case_dm_baseline():
    """distill_baseline.py:84-90,334-356 transcribed, two iterations (momentum 0.5)."""
    C, per, T, H, ipc, batch_real, lr_img = 3, 4, 4, 64, 2, 3, 0.5
    videos, labels, indices_class = _real_set(C, per, T, H, 31)
    dst_train = _SynthSet(videos, labels)

    def get_images(c, n):                       # distill_baseline.py:84-90
        idx_shuffle = np.random.permutation(indices_class[c])[:n]
        if n == 1:
            imgs = dst_train[idx_shuffle[0]][0].unsqueeze(0)
        else:
            imgs = torch.cat([dst_train[i][0].unsqueeze(0) for i in idx_shuffle], 0)
        return imgs

    image_syn = synth.hash_uniform((C * ipc, T, 3, H, H), 32).requires_grad_(True)
    optimizer_img = torch.optim.SGD([image_syn], lr=lr_img, momentum=0.5)
    np.random.seed(7)
    mine_syn, mine_buf = image_syn.detach().clone(), None
    out = {}
    for it in range(2):
        params = synth.synth_convnet3d_params(40 + it, num_classes=C)
        net = ref_networks.ConvNet3D(3, C, 128, 3, 'relu', 'none', 'maxpooling', frames=T, im_size=(H, H))
        load_params(net, params)
        net.train()
        for p in net.parameters():
            p.requires_grad = False
        embed = net.embed
        state = np.random.get_state()
        loss = torch.tensor(0.0)
        for c in range(C):                      # distill_baseline.py:344-351
            img_real = get_images(c, batch_real)
            img_syn = image_syn[c * ipc:(c + 1) * ipc].reshape((ipc, T, 3, H, H))
            output_real = embed(img_real).detach()
            output_syn = embed(img_syn)
            loss += torch.sum((torch.mean(output_real, dim=0) - torch.mean(output_syn, dim=0)) ** 2)
        optimizer_img.zero_grad()
        loss.backward()
        grad_ref = image_syn.grad.detach().clone()
        optimizer_img.step()
        # oracle replay from the same numpy state
        np.random.set_state(state)
        r = oracle.dm_baseline_iteration(params, mine_syn, videos, indices_class, ipc=ipc, batch_real=batch_real)
        mine_syn, mine_buf = oracle.sgd_momentum_step(mine_syn, r['grad_image_syn'], mine_buf, lr_img, 0.5)
        assert rel(r['loss'], loss.detach()) < 1e-6, (r['loss'], loss)
        assert rel(r['grad_image_syn'], grad_ref) < 1e-5
        assert rel(mine_syn, image_syn.detach()) < 1e-6
        out[f'loss{it}'] = loss.detach().numpy()
        out[f'real_idx{it}'] = np.stack(r['real_idx'])
        s, samp = synth.summarize(grad_ref)
        out[f'grad_sums{it}'], out[f'grad_sample{it}'] = s, samp
        s, samp = synth.summarize(image_syn)
        out[f'syn_sums{it}'], out[f'syn_sample{it}'] = s, samp
        out[f'emb_syn{it}'] = torch.stack(r['emb_syn']).numpy()
        out[f'emb_real_mean{it}'] = torch.stack([e.mean(0) for e in r['emb_real']]).numpy()
    np.savez_compressed(os.path.join(GOLD, 'dm_baseline.npz'), **out)
    print('dm_baseline ok: losses', out['loss0'], out['loss1'])


def case_dm_s2d():
    """distill_s2d_ms.py:393-438 transcribed (vpc=1/spc=2/dpc=2 and vpc=2/spc=4/dpc=4), 2 iterations."""
    out = {}
    for tag, (vpc, spc, dpc) in {'v1': (1, 2, 2), 'v2': (2, 4, 4)}.items():
        C, per, T, H, batch_real = 3, 4, 4, 64, 3
        lr_dynamic, lr_hal = 10.0, 0.01
        videos, labels, indices_class = _real_set(C, per, T, H, 51)
        static_syn = synth.hash_uniform((C * spc, 3, H, H), 52)
        dynamic_syn = synth.hash_uniform((C, dpc, T, 1, H, H), 53).requires_grad_(True)
        hal_p = synth.synth_hallucinator(5)
        hal = ref_utils.Conv3DNet()
        hal.load_state_dict({k: v.clone() for k, v in hal_p.items()})
        hals = torch.nn.ModuleList([hal])
        optimizer_dynamic = torch.optim.SGD([dynamic_syn], lr=lr_dynamic, momentum=0.95)
        optimizer_hals = torch.optim.SGD(hals.parameters(), lr=lr_hal, momentum=0.95)
        np.random.seed(9)
        gen = torch.Generator().manual_seed(123)
        m_dyn, m_dyn_buf = dynamic_syn.detach().clone(), None
        m_w, m_w_buf = hal_p['encoder.weight'].clone(), None
        m_b, m_b_buf = hal_p['encoder.bias'].clone(), None
        for it in range(2):
            params = synth.synth_convnet3d_params(60 + it, num_classes=C)
            net = ref_networks.ConvNet3D(3, C, 128, 3, 'relu', 'none', 'maxpooling', frames=T, im_size=(H, H))
            load_params(net, params)
            net.train()
            for p in net.parameters():
                p.requires_grad = False
            embed = net.embed
            # distill_s2d_ms.py:402-412
            label = torch.tensor(np.stack([np.ones(vpc) * i for i in range(0, C)]), dtype=torch.long).view(-1)
            ran = torch.arange(0, C * vpc)
            idx = ran % vpc
            cd = torch.randint(2, (C * vpc,), generator=gen)
            cs = torch.randint(2, (C * vpc,), generator=gen)
            dynamic_idx = 2 * idx + cd
            static_idx = spc * label + 2 * idx + cs
            static = static_syn[static_idx, :, :, :]
            dynamic = dynamic_syn[label, dynamic_idx, :, :, :, :]
            image_syn = hals[0](static, dynamic)
            state = np.random.get_state()
            loss = torch.tensor(0.0)
            for c in range(C):                  # :415-422
                idx_shuffle = np.random.permutation(indices_class[c])[:batch_real]
                img_real = torch.cat([videos[i].unsqueeze(0) for i in idx_shuffle], 0)
                img_syn = image_syn[c * vpc:(c + 1) * vpc].reshape((vpc, T, 3, H, H))
                output_real = embed(img_real).detach()
                output_syn = embed(img_syn)
                loss += torch.sum((torch.mean(output_real, dim=0) - torch.mean(output_syn, dim=0)) ** 2)
            optimizer_dynamic.zero_grad()
            optimizer_hals.zero_grad()
            loss.backward()
            g_dyn = dynamic_syn.grad.detach().clone()
            g_w = hal.encoder.weight.grad.detach().clone()
            g_b = hal.encoder.bias.grad.detach().clone()
            optimizer_dynamic.step()
            optimizer_hals.step()
            # oracle replay
            np.random.set_state(state)
            r = oracle.dm_s2d_iteration(params, static_syn, m_dyn, {'encoder.weight': m_w, 'encoder.bias': m_b},
                                        videos, indices_class, vpc=vpc, spc=spc, batch_real=batch_real,
                                        coin_dynamic=cd, coin_static=cs)
            assert torch.equal(r['label'], label) and torch.equal(r['dynamic_idx'], dynamic_idx)
            assert torch.equal(r['static_idx'], static_idx)
            assert rel(r['loss'], loss.detach()) < 1e-6
            assert rel(r['grad_dynamic'], g_dyn) < 1e-5 and rel(r['grad_hal_weight'], g_w) < 1e-5
            assert rel(r['grad_hal_bias'], g_b) < 1e-5
            m_dyn, m_dyn_buf = oracle.sgd_momentum_step(m_dyn, r['grad_dynamic'], m_dyn_buf, lr_dynamic, 0.95)
            m_w, m_w_buf = oracle.sgd_momentum_step(m_w, r['grad_hal_weight'], m_w_buf, lr_hal, 0.95)
            m_b, m_b_buf = oracle.sgd_momentum_step(m_b, r['grad_hal_bias'], m_b_buf, lr_hal, 0.95)
            assert rel(m_dyn, dynamic_syn.detach()) < 1e-6 and rel(m_w, hal.encoder.weight.detach()) < 1e-6
            k = f'{tag}_{it}'
            out[f'loss_{k}'] = loss.detach().numpy()
            out[f'coin_dynamic_{k}'], out[f'coin_static_{k}'] = cd.numpy(), cs.numpy()
            out[f'static_idx_{k}'], out[f'dynamic_idx_{k}'] = static_idx.numpy(), dynamic_idx.numpy()
            out[f'real_idx_{k}'] = np.stack(r['real_idx'])
            s, samp = synth.summarize(g_dyn)
            out[f'grad_dynamic_sums_{k}'], out[f'grad_dynamic_sample_{k}'] = s, samp
            out[f'grad_hal_weight_{k}'], out[f'grad_hal_bias_{k}'] = g_w.numpy(), g_b.numpy()
            s, samp = synth.summarize(dynamic_syn)
            out[f'dynamic_sums_{k}'], out[f'dynamic_sample_{k}'] = s, samp
            out[f'hal_weight_{k}'] = hal.encoder.weight.detach().numpy().copy()
            out[f'emb_syn_{k}'] = torch.stack(r['emb_syn']).numpy()
        print(f'dm_s2d[{tag}] ok: losses', out[f'loss_{tag}_0'], out[f'loss_{tag}_1'])
    np.savez_compressed(os.path.join(GOLD, 'dm_s2d.npz'), **out)


def case_mtt_s2d():
    """distill_s2d_ms.py:197-292 transcribed: ReparamModule student, 2 inner steps, train mode."""
    C, T, H, vpc, spc, dpc, syn_steps = 3, 8, 64, 1, 2, 2, 2
    static_syn = synth.hash_uniform((C * spc, 3, H, H), 71)
    dynamic_syn = synth.hash_uniform((C, dpc, T, 1, H, H), 72).requires_grad_(True)
    hal_p = synth.synth_hallucinator(7)
    hal = ref_utils.Conv3DNet()
    hal.load_state_dict({k: v.clone() for k, v in hal_p.items()})
    syn_lr = torch.tensor(0.01).requires_grad_(True)
    start = synth.synth_convnet3d_params(80, num_classes=C)
    target = {k: v + synth.hash_uniform(tuple(v.shape), 900 + i, 2.0 ** -10) for i, (k, v) in enumerate(start.items())}
    net = ref_networks.ConvNet3D(3, C, 128, 3, 'relu', 'none', 'maxpooling', frames=T, im_size=(H, H))
    student_net = ref_reparam.ReparamModule(net)
    student_net.train()
    num_params = sum([np.prod(p.size()) for p in (student_net.parameters())])
    target_params = torch.cat([p.data.reshape(-1) for p in target.values()], 0)
    student_params = [torch.cat([p.data.reshape(-1) for p in start.values()], 0).requires_grad_(True)]
    starting_params = torch.cat([p.data.reshape(-1) for p in start.values()], 0)
    criterion = torch.nn.CrossEntropyLoss()
    torch.manual_seed(99)
    perms, cds, css, masks = [], [], [], []
    for step in range(syn_steps):              # :238-266
        indices = torch.randperm(C * vpc)
        these_indices = list(torch.split(indices, C * vpc)).pop()
        label = these_indices // vpc
        idx = these_indices % vpc
        cd = torch.randint(2, (these_indices.shape[0],))
        cs = torch.randint(2, (these_indices.shape[0],))
        dynamic_idx = 2 * idx + cd
        static_idx = spc * label + 2 * idx + cs
        static = static_syn[static_idx, :, :, :]
        dynamic = dynamic_syn[label, dynamic_idx, :, :, :, :]
        x = hal(static, dynamic)
        this_y = label.long()
        # replay the dropout mask: same generator state, same shape as nn.Dropout sees
        st = torch.get_rng_state()
        mask = (torch.nn.functional.dropout(torch.ones(C * vpc, 128, T // 8, 1, 1), 0.5, True) > 0).float()
        torch.set_rng_state(st)
        x = student_net(x, flat_param=student_params[-1])
        loss = criterion(x, this_y)
        grad = torch.autograd.grad(loss, student_params[-1], create_graph=True)[0]
        student_params.append(student_params[-1] - syn_lr * grad)
        perms.append(these_indices)
        cds.append(cd)
        css.append(cs)
        masks.append(mask)
    param_loss = torch.nn.functional.mse_loss(student_params[-1], target_params, reduction="sum")
    param_dist = torch.nn.functional.mse_loss(starting_params, target_params, reduction="sum")
    param_loss = param_loss / num_params
    param_dist = param_dist / num_params
    grand_loss = param_loss / param_dist
    grand_loss.backward()
    r = oracle.mtt_s2d_iteration(starting_params, target_params, start, static_syn, dynamic_syn.detach(), hal_p,
                                 syn_lr.detach(), vpc=vpc, spc=spc, perms=perms, coins_dynamic=cds,
                                 coins_static=css, dropout_masks=masks, im_size=(H, H))
    assert rel(r['grand_loss'], grand_loss.detach()) < 1e-6, (r['grand_loss'], grand_loss)
    assert rel(r['grad_dynamic'], dynamic_syn.grad) < 1e-4, rel(r['grad_dynamic'], dynamic_syn.grad)
    assert rel(r['grad_hal_weight'], hal.encoder.weight.grad) < 1e-4
    assert rel(r['grad_syn_lr'], syn_lr.grad) < 1e-4
    s, samp = synth.summarize(dynamic_syn.grad)
    np.savez_compressed(
        os.path.join(GOLD, 'mtt_s2d.npz'), grand_loss=grand_loss.detach().numpy(),
        param_dist=param_dist.detach().numpy(), perms=torch.stack(perms).numpy(),
        coins_dynamic=torch.stack(cds).numpy(), coins_static=torch.stack(css).numpy(),
        dropout_masks=torch.stack(masks).numpy(), grad_dynamic_sums=s, grad_dynamic_sample=samp,
        grad_hal_weight=hal.encoder.weight.grad.numpy(), grad_hal_bias=hal.encoder.bias.grad.numpy(),
        grad_syn_lr=syn_lr.grad.numpy(), ce_losses=torch.stack(r['ce_losses']).numpy(),
        theta_final_sums=synth.summarize(student_params[-1])[0])
    print('mtt_s2d ok: grand_loss', float(grand_loss))


def case_mtt_baseline():
    """distill_baseline.py:196-272 transcribed: MTT on leaf synthetic videos, ReparamModule student, 2 inner steps, train mode,
    ipc = 2 with batch_syn = 4 < C * ipc, so that the split / pop() order of the index chunks (:233-237) is exercised."""
    C, T, H, ipc, syn_steps, batch_syn = 3, 8, 64, 2, 2, 4
    image_syn = synth.hash_uniform((C * ipc, T, 3, H, H), 91).requires_grad_(True)
    label_syn = torch.tensor(np.stack([np.ones(ipc) * i for i in range(0, C)]), dtype=torch.long).view(-1)
    syn_lr = torch.tensor(0.01).requires_grad_(True)
    start = synth.synth_convnet3d_params(81, num_classes=C)
    target = {k: v + synth.hash_uniform(tuple(v.shape), 950 + i, 2.0 ** -10) for i, (k, v) in enumerate(start.items())}
    net = ref_networks.ConvNet3D(3, C, 128, 3, 'relu', 'none', 'maxpooling', frames=T, im_size=(H, H))
    student_net = ref_reparam.ReparamModule(net)
    student_net.train()
    num_params = sum([np.prod(p.size()) for p in (student_net.parameters())])
    target_params = torch.cat([p.data.reshape(-1) for p in target.values()], 0)
    student_params = [torch.cat([p.data.reshape(-1) for p in start.values()], 0).requires_grad_(True)]
    starting_params = torch.cat([p.data.reshape(-1) for p in start.values()], 0)
    criterion = torch.nn.CrossEntropyLoss()
    torch.manual_seed(98)
    syn_images, y_hat = image_syn, label_syn
    indices_chunks, perms, used, masks = [], [], [], []
    for step in range(syn_steps):              # :231-252
        if not indices_chunks:
            indices = torch.randperm(len(syn_images))
            perms.append(indices.clone())
            indices_chunks = list(torch.split(indices, batch_syn))
        these_indices = indices_chunks.pop()
        x = syn_images[these_indices]
        this_y = y_hat[these_indices]
        st = torch.get_rng_state()
        mask = (torch.nn.functional.dropout(torch.ones(these_indices.shape[0], 128, T // 8, 1, 1), 0.5, True) > 0).float()
        torch.set_rng_state(st)
        x = student_net(x, flat_param=student_params[-1])
        ce_loss = criterion(x, this_y)
        grad = torch.autograd.grad(ce_loss, student_params[-1], create_graph=True)[0]
        student_params.append(student_params[-1] - syn_lr * grad)
        used.append(these_indices)
        masks.append(mask)
    param_loss = torch.nn.functional.mse_loss(student_params[-1], target_params, reduction="sum")
    param_dist = torch.nn.functional.mse_loss(starting_params, target_params, reduction="sum")
    param_loss = param_loss / num_params
    param_dist = param_dist / num_params
    grand_loss = param_loss / param_dist
    grand_loss.backward()
    r = oracle.mtt_baseline_iteration(starting_params, target_params, start, image_syn.detach(), label_syn, syn_lr.detach(),
                                      perms=used, dropout_masks=masks, im_size=(H, H))
    assert rel(r['grand_loss'], grand_loss.detach()) < 1e-6, (r['grand_loss'], grand_loss)
    assert rel(r['grad_image_syn'], image_syn.grad) < 1e-4, rel(r['grad_image_syn'], image_syn.grad)
    assert rel(r['grad_syn_lr'], syn_lr.grad) < 1e-4
    s, samp = synth.summarize(image_syn.grad)
    np.savez_compressed(
        os.path.join(GOLD, 'mtt_baseline.npz'), grand_loss=grand_loss.detach().numpy(), param_dist=param_dist.detach().numpy(),
        perms=torch.stack(perms).numpy(), used_0=used[0].numpy(), used_1=used[1].numpy(),
        dropout_mask_0=masks[0].numpy(), dropout_mask_1=masks[1].numpy(),
        grad_image_sums=s, grad_image_sample=samp, grad_syn_lr=syn_lr.grad.numpy(),
        rows_with_grad=(image_syn.grad.flatten(1).abs().sum(1) > 0).numpy())
    print('mtt_baseline ok: grand_loss', float(grand_loss), 'chunks', [u.tolist() for u in used])


def case_coreset():
    """distill_coreset.py:75-110: the selection statements of the reference's main(), executed verbatim (text between
    `features = embed(imgs)` and the `image_syn[...] =` assignment of each method) on hash-generated feature matrices."""
    import textwrap
    from oracle import coreset as oc
    src = open(os.path.join(REF, 'distill_coreset.py')).read()
    blocks = {}
    for method, result in (('k-center', 'idx_centers'), ('herding', 'idx_selected')):
        seg = src[src.index("args.method == '%s'" % method):]
        seg = seg[seg.index('features = embed(imgs)') + len('features = embed(imgs)'):]
        seg = seg[:seg.index('image_syn[c*args.ipc')]
        blocks[method] = (textwrap.dedent('\n'.join(seg.splitlines()[1:])), result)
    out = {}
    for tag, n, d, ipc in (('a', 7, 5, 1), ('b', 20, 16, 5), ('c', 64, 32, 10), ('d', 5, 3, 5)):
        features = synth.hash_uniform((n, d), 700 + n)
        args = type('A', (), {})()
        args.ipc = ipc
        for method, fn in (('k-center', oc.k_center), ('herding', oc.herding)):
            code, result = blocks[method]
            # the reference's k-center reduces the (n,) centre distances over the SAMPLE axis (:87 torch.min(dis_center, dim=-1) on a
            # vector), so its second centre is always sample 0, and from the third centre on :86 no longer broadcasts and raises.
            # It is pinned where it is well defined (the first centre, ipc = 1); the oracle / product implement the intended
            # greedy farthest-point rule (min over centres, argmax over samples)
            args.ipc = 1 if method == 'k-center' else ipc
            ns = {'torch': torch, 'np': np, 'features': features, 'args': args}
            exec(code, ns)
            assert ns[result] == fn(features, args.ipc), (tag, method, ns[result], fn(features, args.ipc))
            out[f'{tag}_{method}'] = np.asarray(ns[result], dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, 'coreset.npz'), **out)
    print('coreset ok:', {k: v.tolist() for k, v in out.items() if k.startswith('b_')})


def write_expert_files(directory, n_files=2, n_traj=3, n_epochs=5):
    """replay_buffer_{n}.pt in the layout of buffer.py:75-103; every tensor encodes (file, trajectory, epoch)."""
    for n in range(n_files):
        traj = [[[torch.full((2,), float(100 * n + 10 * t + e)) for _ in range(3)] for e in range(n_epochs)] for t in range(n_traj)]
        torch.save(traj, os.path.join(directory, 'replay_buffer_{}.pt'.format(n)))


def case_expert_walk():
    """distill_s2d_ms.py:114-133 (buffer discovery / shuffles) and :209-229 (per-iteration trajectory and start-epoch draw): the
    reference's own statements executed verbatim on synthetic buffer files; pins cli.ExpertBuffers."""
    import random
    import tempfile
    import textwrap
    src = open(os.path.join(REF, 'distill_s2d_ms.py')).read()
    head = src[src.index('        expert_files = []'):src.index('        best_acc = {m: 0 for m in model_eval_pool}')]
    body = src[src.index('            expert_trajectory = buffer[expert_idx]'):src.index('            target_params = torch.cat([p.data.to(args.device)')]
    with tempfile.TemporaryDirectory() as d:
        write_expert_files(d)
        args = type('A', (), {})()
        args.max_start_epoch, args.expert_epochs = 3, 1
        ns = {'os': os, 'torch': torch, 'np': np, 'random': random, 'expert_dir': d, 'args': args, 'print': lambda *a, **k: None}
        random.seed(11)
        np.random.seed(12)
        exec(textwrap.dedent(head), ns)
        seq = []
        for _ in range(9):
            exec(textwrap.dedent(body), ns)
            seq.append((float(ns['starting_params'][0][0]), float(ns['target_params'][0][0]), int(ns['start_epoch'])))
    np.savez_compressed(os.path.join(GOLD, 'expert_walk.npz'), walk=np.asarray(seq, dtype=np.float64))
    print('expert walk ok:', seq[:4])


def case_multistatic():
    """utils.MultiStaticSharedDataset.__getitem__ (:469-488) of the live reference with a recording hallucinator: which static row /
    dynamic memory each sample pairs, and in which order random.randint is consumed, for spc = 2 (vpc = 1) and spc = 10 (vpc = 5)."""
    import random
    out = {}
    for tag, C, spc, dpc in (('vpc1', 4, 2, 2), ('vpc5', 3, 10, 10)):
        static = torch.arange(C * spc).float().view(-1, 1, 1, 1).expand(C * spc, 3, 2, 2).contiguous()
        dynamic = (torch.arange(C).view(C, 1) * 100 + torch.arange(dpc).view(1, dpc)).float().view(C, dpc, 1, 1, 1, 1).expand(C, dpc, 2, 1, 2, 2).contiguous()
        seen = []

        class Rec(torch.nn.Module):
            def forward(self, st, dy):
                seen.append((int(st.flatten()[0]), int(dy.flatten()[0])))
                return st.unsqueeze(1)

        ds = ref_utils.MultiStaticSharedDataset(static, dynamic, torch.nn.ModuleList([Rec(), Rec(), Rec()]))
        random.seed(21)
        rows = []
        for index in list(range(len(ds))) * 2:
            _, label = ds[index]
            st, dy = seen[-1]
            rows.append((index, st, int(label), dy % 100))
            assert dy // 100 == int(label)
        out[tag] = np.asarray(rows, dtype=np.int64)
        out[tag + '_len'] = np.int64(len(ds))
        out[tag + '_next_random'] = np.float64(random.random())          # state of the generator after the walk
    np.savez_compressed(os.path.join(GOLD, 'multistatic.npz'), **out)
    print('multistatic ok:', out['vpc1'][:4].tolist())


def case_eval_schedule():
    """utils.evaluate_synset (:848-886) of the live reference with utils.epoch replaced by a recorder: the sequence of
    (mode, learning rate of the optimizer handed to epoch) calls for Epoch = 7 without and with test_freq."""
    out = {}
    for tag, test_freq in (('final', None), ('freq3', 3)):
        calls = []

        def recorder(mode, dataloader, net, optimizer, criterion, args):
            calls.append((0 if mode == 'train' else 1, optimizer.param_groups[0]['lr'], optimizer.param_groups[0]['momentum'],
                          optimizer.param_groups[0]['weight_decay']))
            return 0.5, 0.25, [0.25]
        real = ref_utils.epoch
        ref_utils.epoch = recorder
        try:
            args = type('A', (), {})()
            args.lr_net, args.epoch_eval_train, args.device, args.batch_train, args.eval_mode = 0.02, 7, 'cpu', 4, 'S'
            net = torch.nn.Linear(3, 2)
            _, acc_train, acc_test, acc_per = ref_utils.evaluate_synset(0, net, torch.zeros(6, 2, 3, 4, 4), torch.zeros(6).long(), None, args,
                                                                        mode='none', test_freq=test_freq)
        finally:
            ref_utils.epoch = real
        out[tag] = np.asarray(calls, dtype=np.float64)
        out[tag + '_ret'] = np.asarray([acc_train, acc_test], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, 'eval_schedule.npz'), **out)
    print('eval schedule ok:', out['final'][:, :2].tolist())


def epoch_case_inputs(C, sizes, seed):
    """Hash-generated logits / labels / per-batch losses of the epoch bookkeeping case (shared with tests/test_epoch_stats_cpu.py)."""
    batches = []
    for i, n in enumerate(sizes):
        logits = synth.hash_uniform((n, C), seed + 10 * i, 3.0)
        labels = (synth.hash_uniform((n,), seed + 10 * i + 1, 0.5) + 0.5).mul(min(C, 12)).long().clamp_(0, C - 1)
        labels[::2] = logits[::2].argmax(-1)                       # every second sample is classified correctly
        labels[1::4] = logits[1::4].topk(min(3, C), dim=-1).indices[:, -1]      # ... and some only within the top 3
        batches.append((logits, labels))
    return batches


def case_epoch():
    """utils.epoch (:752-845) of the LIVE reference driven by a stub network that returns prescribed logits: pins the accuracy /
    top-k / per-class bookkeeping that video_distillation_b200.utils.EpochStats reproduces on the device."""
    out = {}
    for tag, C, train in (('test50', 50, False), ('train50', 50, True), ('test3', 3, False), ('train7', 7, True)):
        batches = epoch_case_inputs(C, (16, 16, 5), 400 + C)

        class Stub(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.w = torch.nn.Parameter(torch.zeros(()))
                self.k = 0

            def forward(self, x):
                logits = batches[self.k % len(batches)][0]
                self.k += 1
                return logits + 0.0 * self.w

        net = Stub()
        loader = [(torch.zeros(lab.shape[0], 2, 3, 4, 4) + torch.arange(lab.shape[0]).view(-1, 1, 1, 1, 1).float(), lab) for _, lab in batches]
        args = type('A', (), {})()
        args.device, args.model, args.eval_mode = 'cpu', 'ConvNet3D', 'top5'
        opt = torch.optim.SGD(net.parameters(), lr=0.0)
        loss, accs, per = ref_utils.epoch('train' if train else 'test', loader, net, opt, torch.nn.CrossEntropyLoss(), args)
        out[tag + '_loss'] = np.float64(loss)
        out[tag + '_accs'] = np.asarray(accs, dtype=np.float64)
        out[tag + '_per_class'] = np.asarray([np.nan if v is None else v for v in per], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, 'epoch_stats.npz'), **out)
    print('epoch ok:', {k: (v.tolist() if v.size < 6 else v.shape) for k, v in out.items() if 'accs' in k})


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1:                      # regenerate selected cases only: python oracle/make_golden.py case_mtt_baseline
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    torch.set_num_threads(8)
    case_init()
    case_convnet3d()
    case_composer()
    case_dm_baseline()
    case_dm_s2d()
    case_mtt_s2d()
    case_mtt_baseline()
    case_epoch()
    case_coreset()
    case_expert_walk()
    case_multistatic()
    case_eval_schedule()
    print('golden vectors written to', GOLD)
