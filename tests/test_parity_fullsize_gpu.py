"""Parity of the DEFAULT (f16x3) mode at the bench shape against the CPU oracle (oracle/dm.py, pinned to the reference's
distill_s2d_ms.py:393-431): DM + S2D, 16x3x112x112 videos, batch_real 64, vpc 1 — two FULL classes (128 real + 2 synthetic
videos, ~1.5 TFLOP on the host cores).

Gates (north_star: 1e-3 relative): loss, real class means, synthetic embeddings <= 1e-3 (asserted at 2e-4); gradients of the
dynamic memory and the hallucinator <= 1e-3 CONDITIONED on the oracle's ReLU masks / pool indices; the unconditioned figures
are printed beside the fp32-vs-fp64 floor of the same problem and asserted loosely.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

C, T, HW, PER, BATCH_REAL = 2, 16, 112, 64, 64
S, P = (1, 2, 2), (1, 3, 3)
POOL = [(1, 2, 2), (2, 2, 2), (2, 2, 2)]


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def oracle_codes(params, video, dtype=torch.float32):
    h = video.permute(0, 2, 1, 3, 4).to(dtype)
    codes = []
    for d in range(3):
        y = F.conv3d(h, params[f'features.{3 * d}.weight'].to(dtype), params[f'features.{3 * d}.bias'].to(dtype), S, P)
        k = POOL[d]
        h, idx = F.max_pool3d(F.relu(y), k, k, return_indices=True)
        To, Ho, Wo = y.shape[2:]
        it, ih, iw = idx // (Ho * Wo), (idx // Wo) % Ho, idx % Wo
        pos = (it % k[0]) * (k[1] * k[2]) + (ih % k[1]) * k[2] + (iw % k[2])
        codes.append((pos | ((h > 0).long() << 3)).to(torch.uint8))
    return codes


_ORACLE = {}


@pytest.mark.parametrize('precision', ['f16x3r2', 'f16x3'])
def test_default_mode_matches_the_cpu_oracle_on_two_full_classes(precision):
    """f16x3r2 (default): synthetic branch on three products per MAC, frozen real branch on two (exact weights, activations
    rounded once to fp16); f16x3: three products everywhere."""
    import oracle
    from oracle import synth
    from video_distillation_b200.distill import DeviceDataset, DMS2DTrainer
    from video_distillation_b200.networks import ConvNet3D
    from video_distillation_b200.utils import Conv3DNet
    torch.set_num_threads(max(1, torch.get_num_threads()))
    gen = torch.Generator().manual_seed(21)
    videos = torch.randn(C * PER, T, 3, HW, HW, generator=gen)
    labels = [c for c in range(C) for _ in range(PER)]
    params = oracle.init_convnet3d(4242, num_classes=C)
    static = torch.randn(C * 2, 3, HW, HW, generator=gen)
    dynamic = torch.randn(C, 2, T, 1, HW, HW, generator=gen)
    halp = synth.synth_hallucinator(5)
    coin_d, coin_s = torch.tensor([1, 0]), torch.tensor([0, 1])
    real_idx = [np.random.RandomState(7 + c).permutation(np.arange(c * PER, (c + 1) * PER))[:BATCH_REAL] for c in range(C)]
    indices_class = [list(range(c * PER, (c + 1) * PER)) for c in range(C)]

    # ---- CPU oracle, fp32 (what the reference computes) and fp64 (for the floor); computed once for both precisions
    if not _ORACLE:
        _ORACLE['ref'] = oracle.dm_s2d_iteration(params, static, dynamic, halp, videos, indices_class, vpc=1, spc=2, batch_real=BATCH_REAL,
                                                 coin_dynamic=coin_d, coin_static=coin_s, real_idx=real_idx)
        p64 = {k: v.double() for k, v in params.items()}
        _ORACLE['ref64'] = oracle.dm_s2d_iteration(p64, static.double(), dynamic.double(), {k: v.double() for k, v in halp.items()},
                                                   videos.double(), indices_class, vpc=1, spc=2, batch_real=BATCH_REAL,
                                                   coin_dynamic=coin_d, coin_static=coin_s, real_idx=real_idx)
        _ORACLE['codes32'] = oracle_codes(params, _ORACLE['ref']['image_syn'])
    ref, ref64, codes32 = _ORACLE['ref'], _ORACLE['ref64'], _ORACLE['codes32']

    # ---- the CUDA path in its default precision
    ds = DeviceDataset(videos, labels, C, 'cuda')
    net = ConvNet3D(3, C, 128, 3, 'relu', 'none', 'maxpooling', T, (HW, HW))
    net.load_state_dict(params)
    net = net.cuda().train()
    for q in net.parameters():
        q.requires_grad = False
    label, _, didx, sidx = oracle.s2d_sample_indices(C, 1, 2, coin_d, coin_s)

    def run(codes=None):
        hal = Conv3DNet()
        hal.load_state_dict(halp)
        tr = DMS2DTrainer(ds, num_classes=C, im_size=(HW, HW), frames=T, vpc=1, spc=2, dpc=2, batch_real=BATCH_REAL,
                          lr_dynamic=1e4, lr_hal=1e-2, precision=precision, hal=hal, static_syn=static, dynamic_syn=dynamic)
        tr.embedder.tc.codes_override = codes
        loss = tr.step(net=net, indices=(label.cuda(), didx.cuda(), sidx.cuda()), real_idx=np.stack(real_idx))
        return loss.item(), tr.last, tr.dynamic_syn.grad.clone(), tr.hal.encoder.weight.grad.clone(), tr.hal.encoder.bias.grad.clone()
    loss, last, g_dyn, g_hw, g_hb = run()
    _, _, gc_dyn, gc_hw, gc_hb = run(tuple(c.cuda() for c in codes32))

    mean_ref = torch.stack([e.mean(0) for e in ref['emb_real']])
    es_ref = torch.stack(ref['emb_syn'])
    e = dict(loss=abs(loss - ref['loss'].item()) / abs(ref['loss'].item()),
             mean_real=rel(last['mean_real'], mean_ref), emb_syn=rel(last['emb_syn'], es_ref),
             image_syn=rel(last['image_syn'], ref['image_syn']))
    cond = dict(dyn=rel(gc_dyn, ref['grad_dynamic']), hal_w=rel(gc_hw, ref['grad_hal_weight']), hal_b=rel(gc_hb, ref['grad_hal_bias']))
    unc = dict(dyn=rel(g_dyn, ref['grad_dynamic']), hal_w=rel(g_hw, ref['grad_hal_weight']), hal_b=rel(g_hb, ref['grad_hal_bias']))
    floor = dict(dyn=rel(ref['grad_dynamic'], ref64['grad_dynamic']), hal_w=rel(ref['grad_hal_weight'], ref64['grad_hal_weight']),
                 hal_b=rel(ref['grad_hal_bias'], ref64['grad_hal_bias']))
    print(f'{precision} vs CPU oracle at the bench shape (2 full classes):', {k: f'{v:.2e}' for k, v in e.items()})
    print('  gradients conditioned on the oracle routing :', {k: f'{v:.2e}' for k, v in cond.items()})
    print('  gradients unconditioned                     :', {k: f'{v:.2e}' for k, v in unc.items()})
    print('  fp32 oracle vs fp64 oracle (the floor)      :', {k: f'{v:.2e}' for k, v in floor.items()})
    assert e['image_syn'] < 1e-5
    assert e['loss'] < 2e-4 and e['mean_real'] < 2e-4 and e['emb_syn'] < 2e-4, e
    assert all(v < 1e-3 for v in cond.values()), cond
    assert all(v < 5e-2 for v in unc.values()), unc
    assert np.array_equal(last['real_idx'], np.stack(real_idx))
