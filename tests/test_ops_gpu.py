"""GPU parity of the exact fp32 kernels against the CPU oracle / plain torch fp64 (through the C ABI).

Tolerance for fp32 kernels vs fp64 truth: 2e-5 relative L2 (summation order only); integer /
index outputs (routing codes, gathers) are bit-exact.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


GEOMS = [
    # (N, Cin, T, H, W, Cout, k, s, p)
    (2, 3, 4, 20, 20, 16, (3, 7, 7), (1, 2, 2), (1, 3, 3)),
    (2, 8, 4, 7, 7, 12, (3, 7, 7), (1, 2, 2), (1, 3, 3)),          # H=7 -> 4: non-injective output size
    (3, 4, 5, 9, 8, 3, (3, 3, 3), (1, 1, 1), (1, 1, 1)),           # composer geometry
    (2, 16, 3, 1, 1, 5, (1, 1, 1), (1, 1, 1), (0, 0, 0)),          # logit conv
]


@pytest.mark.parametrize('geom', GEOMS)
def test_conv_trio(geom):
    from video_distillation_b200 import ops
    N, Cin, T, H, W, Cout, k, s, p = geom
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, Cin, T, H, W, generator=g)
    w = torch.randn(Cout, Cin, *k, generator=g) * 0.1
    b = torch.randn(Cout, generator=g)
    xd, wd, bd = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    y_ref = F.conv3d(xd, wd, bd, stride=s, padding=p)
    gy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(gy.double())
    xc, wc, bc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    y = ops.conv3d(xc, wc, bc, s, p)
    assert rel(y, y_ref) < 2e-5
    y.backward(gy.cuda())
    assert rel(xc.grad, xd.grad) < 2e-5
    assert rel(wc.grad, wd.grad) < 2e-5
    assert rel(bc.grad, bd.grad) < 2e-5


def test_conv_double_backward():
    """grad-of-grad through the Function closure == torch's _convolution_double_backward (MTT path)."""
    from video_distillation_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 3, 9, 9, generator=g)
    w = torch.randn(6, 4, 3, 7, 7, generator=g) * 0.1
    s, p = (1, 2, 2), (1, 3, 3)

    def run(x, w, conv):
        x = x.requires_grad_(True)
        w = w.requires_grad_(True)
        y = conv(x, w)
        loss = (y ** 3).sum()
        gw, = torch.autograd.grad(loss, w, create_graph=True)
        w2 = w - 0.1 * gw
        y2 = conv(x, w2)
        out = (y2 ** 2).sum() + (gw ** 2).sum()
        gx, gw2 = torch.autograd.grad(out, [x, w])
        return out, gx, gw2

    ref = run(x.double(), w.double(), lambda a, b: F.conv3d(a, b, None, s, p))
    got = run(x.cuda(), w.cuda(), lambda a, b: ops.conv3d(a, b, None, s, p))
    for r, q in zip(ref, got):
        assert rel(q, r) < 5e-5, rel(q, r)


@pytest.mark.parametrize('k', [(1, 2, 2), (2, 2, 2), (1, 1, 1)])
def test_relu_maxpool_routing(k):
    from video_distillation_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 5, 4, 6, 10, generator=g)
    x[0, 0, :, :2, :2] = 0.0                      # all-zero window: first index wins, gradient is zero
    x[0, 1, 0, 0, :4] = torch.tensor([0.0, 3.0, 3.0, 1.0])   # tie -> first maximum
    xr = x.clone().requires_grad_(True)
    y_ref, idx = F.max_pool3d(F.relu(xr), k, k, return_indices=True)
    gy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(gy)
    xc = x.cuda().requires_grad_(True)
    y, code = ops.relu_maxpool3d(xc, k, return_code=True)
    assert torch.equal(y.cpu(), y_ref.detach())
    y.backward(gy.cuda())
    assert torch.equal(xc.grad.cpu(), xr.grad)
    # argmax position equals ATen's wherever the output is active
    T, H, W = x.shape[2:]
    it, ih, iw = idx // (H * W), (idx // W) % H, idx % W
    pos = (it % k[0]) * (k[1] * k[2]) + (ih % k[1]) * k[2] + (iw % k[2])
    act = (code.cpu() & 8) > 0
    assert torch.equal(act, y_ref > 0)
    assert torch.equal((code.cpu() & 7).long()[act], pos[act])
    # double backward: gather with the saved code
    c0 = torch.randn(x.shape, generator=g)
    cr = c0.clone().requires_grad_(True)
    z_ref = torch.where(act, cr.flatten(2).gather(2, idx.flatten(2)).view_as(idx), torch.zeros(()))
    z_ref.backward(gy)
    c = c0.cuda().requires_grad_(True)
    z = ops.route_with_code(c, code, k)
    assert torch.equal(z.cpu(), z_ref.detach())
    z.backward(gy.cuda())
    assert torch.equal(c.grad.cpu(), cr.grad)


def test_composer_golden():
    from oracle import synth
    from video_distillation_b200.utils import Conv3DNet
    gold = np.load(os.path.join(GOLD, 'composer.npz'))
    hal_p = synth.synth_hallucinator(3)
    hal = Conv3DNet()
    hal.load_state_dict(hal_p)
    hal = hal.cuda()
    static = synth.hash_uniform((5, 3, 16, 16), 21).cuda().requires_grad_(True)
    dynamic = synth.hash_uniform((5, 4, 1, 16, 16), 22).cuda().requires_grad_(True)
    y = hal(static, dynamic)
    assert rel(y, torch.from_numpy(gold['y'])) < 1e-6
    gy = synth.hash_uniform(tuple(y.shape), 23).cuda()
    y.backward(gy)
    assert rel(static.grad, torch.from_numpy(gold['grad_static'])) < 1e-5
    assert rel(dynamic.grad, torch.from_numpy(gold['grad_dynamic'])) < 1e-5
    assert rel(hal.encoder.weight.grad, torch.from_numpy(gold['grad_weight'])) < 1e-5
    assert rel(hal.encoder.bias.grad, torch.from_numpy(gold['grad_bias'])) < 1e-5


def test_composer_gather_and_accumulate():
    """Fused index gather (distill_s2d_ms.py:409-410) incl. repeated rows -> accumulated gradient."""
    import oracle
    from video_distillation_b200 import ops
    g = torch.Generator().manual_seed(3)
    C, dpc, T, H, W, spc = 3, 2, 4, 10, 12, 2
    static_syn = torch.randn(C * spc, 3, H, W, generator=g)
    dynamic_syn = torch.randn(C, dpc, T, 1, H, W, generator=g)
    hal = oracle.init_hallucinator(5)
    label = torch.tensor([0, 1, 1, 2, 1])
    didx = torch.tensor([1, 0, 0, 1, 1])          # (1,0) appears twice
    sidx = torch.tensor([1, 2, 2, 5, 3])
    s = static_syn.clone().requires_grad_(True)
    d = dynamic_syn.clone().requires_grad_(True)
    w = hal['encoder.weight'].clone().requires_grad_(True)
    b = hal['encoder.bias'].clone().requires_grad_(True)
    y_ref = oracle.compose(s[sidx], d[label, didx], w, b)
    gy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(gy)
    sc, dc = static_syn.cuda().requires_grad_(True), dynamic_syn.cuda().requires_grad_(True)
    wc, bc = hal['encoder.weight'].cuda().requires_grad_(True), hal['encoder.bias'].cuda().requires_grad_(True)
    y = ops.compose(sc, dc, wc, bc, sidx.cuda(), label.cuda(), didx.cuda())
    assert rel(y, y_ref) < 1e-6
    y.backward(gy.cuda())
    assert rel(dc.grad, d.grad) < 1e-5 and rel(sc.grad, s.grad) < 1e-5
    assert rel(wc.grad, w.grad) < 1e-5 and rel(bc.grad, b.grad) < 1e-5


@pytest.mark.parametrize('T,H,W', [(1, 8, 8), (2, 20, 24), (5, 13, 64), (4, 30, 112)])
def test_composer_tiled_paths(T, H, W):
    """Shared-memory tiled composer kernels (frozen static memory, i.e. --no_train_static): forward, gradient to
    the dynamic memory (repeated rows accumulate) and to the hallucinator, at ragged heights / T = 1."""
    import oracle
    from video_distillation_b200 import ops
    g = torch.Generator().manual_seed(11 + T)
    C, dpc, spc = 3, 2, 2
    static_syn = torch.randn(C * spc, 3, H, W, generator=g)
    dynamic_syn = torch.randn(C, dpc, T, 1, H, W, generator=g)
    hal = oracle.init_hallucinator(7)
    label = torch.tensor([0, 1, 1, 2])
    didx = torch.tensor([1, 0, 0, 1])             # (1,0) appears twice
    sidx = torch.tensor([1, 2, 3, 5])
    d = dynamic_syn.clone().requires_grad_(True)
    w = hal['encoder.weight'].clone().requires_grad_(True)
    b = hal['encoder.bias'].clone().requires_grad_(True)
    y_ref = oracle.compose(static_syn[sidx], d[label, didx], w, b)
    gy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(gy)
    dc = dynamic_syn.cuda().requires_grad_(True)
    wc, bc = hal['encoder.weight'].cuda().requires_grad_(True), hal['encoder.bias'].cuda().requires_grad_(True)
    y = ops.compose(static_syn.cuda(), dc, wc, bc, sidx.cuda(), label.cuda(), didx.cuda())
    assert rel(y, y_ref) < 1e-6
    y.backward(gy.cuda())
    assert rel(dc.grad, d.grad) < 1e-5
    assert rel(wc.grad, w.grad) < 1e-5 and rel(bc.grad, b.grad) < 1e-5


def test_dm_loss_and_class_mean():
    import oracle
    from video_distillation_b200 import ops
    g = torch.Generator().manual_seed(4)
    C, nr, ns, D = 7, 9, 3, 300
    er = torch.randn(C, nr, D, generator=g)
    es = torch.randn(C, ns, D, generator=g).requires_grad_(True)
    loss_ref = sum(oracle.dm_loss(er[c], es[c]) for c in range(C))
    loss_ref.backward()
    esc = es.detach().cuda().requires_grad_(True)
    mr = ops.class_mean(er.cuda())
    assert rel(mr, er.mean(1)) < 1e-6
    loss = ops.dm_loss(mr, esc)
    assert rel(loss, loss_ref.detach()) < 1e-6
    (2.0 * loss).backward()
    assert rel(esc.grad, 2.0 * es.grad) < 1e-6


def test_sgd_momentum_matches_torch():
    from video_distillation_b200 import ops
    g = torch.Generator().manual_seed(5)
    p0 = torch.randn(1003, generator=g)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([p_ref], lr=0.3, momentum=0.95)
    p = p0.cuda()
    buf = torch.empty_like(p)
    for it in range(3):
        grad = torch.randn(1003, generator=g)
        if it == 1:
            grad[::2] = 0.0                      # dense semantics: zero-grad entries still move
        p_ref.grad = grad.clone()
        opt.step()
        ops.sgd_momentum_(p, grad.cuda(), buf, 0.3, 0.95, it == 0)
        assert rel(p, p_ref.detach()) < 1e-6


def test_instancenorm_avgpool_variant():
    from video_distillation_b200 import ops
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 6, 4, 6, 6, generator=g)
    gam, bet = torch.rand(6, generator=g) + 0.5, torch.randn(6, generator=g)
    xd, gd, bd = (t.double().requires_grad_(True) for t in (x, gam, bet))
    y_ref = F.avg_pool3d(F.relu(F.group_norm(xd, 6, gd, bd, 1e-5)), 2, 2)
    gy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(gy.double())
    xc, gc, bc = (t.cuda().requires_grad_(True) for t in (x, gam, bet))
    y = ops.avgpool3d_2(ops.instancenorm_relu(xc, gc, bc))
    assert rel(y, y_ref) < 2e-5
    y.backward(gy.cuda())
    assert rel(xc.grad, xd.grad) < 1e-4 and rel(gc.grad, gd.grad) < 1e-4 and rel(bc.grad, bd.grad) < 1e-4


def test_fused_instancenorm_relu_avgpool_forward():
    """vd_inorm_relu_avgpool_fwd_f32 (one launch, frozen networks) against fp64 torch and against the differentiable pair."""
    from video_distillation_b200 import ops
    g = torch.Generator().manual_seed(16)
    for shape in [(2, 6, 4, 6, 6), (3, 64, 8, 28, 28), (1, 128, 4, 14, 14), (2, 5, 2, 4, 4)]:
        x = torch.randn(*shape, generator=g) * 1.7 + 0.3
        C = shape[1]
        gam, bet = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
        y_ref = F.avg_pool3d(F.relu(F.group_norm(x.double(), C, gam.double(), bet.double(), 1e-5)), 2, 2)
        with torch.no_grad():
            y = ops.instancenorm_relu_avgpool(x.cuda(), gam.cuda(), bet.cuda())
            y_pair = ops.avgpool3d_2(ops.instancenorm_relu(x.cuda(), gam.cuda(), bet.cuda()))
        assert y.shape == y_ref.shape
        assert rel(y, y_ref) < 2e-5, (shape, rel(y, y_ref))
        assert rel(y, y_pair) < 1e-6
    # with gradients enabled the op is the differentiable pair
    xc = torch.randn(2, 6, 4, 6, 6, generator=g).cuda().requires_grad_(True)
    gc, bc = torch.ones(6).cuda(), torch.zeros(6).cuda()
    ops.instancenorm_relu_avgpool(xc, gc, bc).sum().backward()
    assert xc.grad is not None and torch.isfinite(xc.grad).all()


def test_instancenorm_avgpool_net_on_tensor_cores():
    """ConvNet3D(instancenorm, avgpooling) (networks.py:765-790) at the bench video shape with the three feature convs on
    tcgen05 (conv trio, split-bf16 fprop) and the fused IN + ReLU + avgpool kernel between them, against the CPU oracle."""
    import oracle
    from oracle import synth
    from video_distillation_b200 import ops
    from video_distillation_b200.networks import ConvNet3D
    T, H = 16, 112
    params = synth.synth_convnet3d_params(3, num_classes=5)
    p2 = {}
    for d in range(3):
        p2[f'features.{4 * d}.weight'] = params[f'features.{3 * d}.weight']
        p2[f'features.{4 * d}.bias'] = params[f'features.{3 * d}.bias']
        c = p2[f'features.{4 * d}.weight'].shape[0]
        p2[f'features.{4 * d + 1}.weight'] = 1.0 + synth.hash_uniform((c,), 300 + d, 0.25)
        p2[f'features.{4 * d + 1}.bias'] = synth.hash_uniform((c,), 310 + d, 0.25)
    p2['logit.weight'], p2['logit.bias'] = params['logit.weight'], params['logit.bias']
    net = ConvNet3D(3, 5, 128, 3, 'relu', 'instancenorm', 'avgpooling', T, (H, H))
    net.load_state_dict(p2)
    net = net.cuda()
    x = synth.hash_uniform((2, T, 3, H, H), 13)
    ref = oracle.convnet3d_embed(p2, x, net_norm='instancenorm', net_pooling='avgpooling')
    with torch.no_grad():
        e32 = net.embed(x.cuda())
        prev = ops.set_conv_backend('tc')
        try:
            etc = net.embed(x.cuda())
        finally:
            ops.set_conv_backend(prev)
    print(f'IN/avgpool variant at 16x3x112x112: fp32 kernels {rel(e32, ref):.2e}, tensor-core convs + fused IN/ReLU/avgpool {rel(etc, ref):.2e}')
    assert e32.shape == ref.shape
    assert rel(e32, ref) < 1e-4
    assert rel(etc, ref) < 1e-3


def test_sqdist():
    from video_distillation_b200 import ops
    g = torch.Generator().manual_seed(7)
    a, b = torch.randn(100003, generator=g), torch.randn(100003, generator=g)
    assert rel(ops.sqdist(a.cuda(), b.cuda()), ((a.double() - b.double()) ** 2).sum()) < 1e-5


def test_dm_loss_variants_are_bitwise_identical_and_ordered():
    """vd_dm_loss_f32 (one block, scratch-free) and vd_dm_loss_ex_f32 (one block per class + ordered finish) give the same
    bits: every class sum is the same tree and the class sums are added in class order (distill_s2d_ms.py:414-422)."""
    from video_distillation_b200._lib import check, lib, ptr, stream
    C, ns, D = 37, 3, 2048
    g = torch.Generator(device='cuda').manual_seed(5)
    mean_real = torch.randn(C, D, device='cuda', generator=g)
    emb_syn = torch.randn(C, ns, D, device='cuda', generator=g)
    la, lb = torch.zeros((), device='cuda'), torch.zeros((), device='cuda')
    ga, gb, cl = torch.empty_like(emb_syn), torch.empty_like(emb_syn), torch.empty(C, device='cuda')
    check(lib().vd_dm_loss_f32(ptr(mean_real), ptr(emb_syn), ptr(la), ptr(ga), C, ns, D, 1.0, stream()), 'a')
    check(lib().vd_dm_loss_ex_f32(ptr(mean_real), ptr(emb_syn), ptr(lb), ptr(gb), ptr(cl), C, ns, D, 1.0, stream()), 'b')
    assert torch.equal(la, lb) and torch.equal(ga, gb)
    per_class = ((mean_real - emb_syn.mean(1)) ** 2).sum(1)
    assert torch.allclose(cl, per_class, rtol=1e-5)
    seq = torch.zeros((), device='cuda')
    for c in range(C):
        seq = seq + cl[c]
    assert torch.equal(seq, lb)


def test_class_sum_ragged_matches_torch():
    from video_distillation_b200 import ops
    g = torch.Generator().manual_seed(4)
    counts = [5, 0, 9, 1, 64, 3]
    emb = torch.randn(sum(counts), 2048, generator=g).cuda()
    offs = torch.tensor([0] + list(np.cumsum(counts)), dtype=torch.int32).cuda()
    out = ops.class_sum_ragged(emb, offs, len(counts))
    ref = torch.stack([emb[int(offs[c]):int(offs[c + 1])].double().sum(0) for c in range(len(counts))]).float()
    assert float((out - ref).abs().max()) < 1e-4 and float(out[1].abs().max()) == 0.0
    # rows are added in row order: equal, bit for bit, to a sequential fp32 sum
    seq = torch.zeros(2048, device='cuda')
    for r in range(5):
        seq = seq + emb[r]
    assert torch.equal(out[0], seq)


@pytest.mark.parametrize('T,H,W,vpc', [(8, 64, 64, 1), (16, 112, 112, 1), (4, 20, 24, 2)])
def test_fused_composer_backward_is_exact_and_reproducible(T, H, W, vpc):
    """vd_compose_bwd_fused_f32 (one pass, fixed-order block sums) against fp64 autograd of utils.py:1186-1197 and against
    itself: two runs give bitwise identical hallucinator gradients (no floating-point atomics)."""
    from video_distillation_b200 import ops
    g = torch.Generator().manual_seed(T + H)
    C, spc, dpc = 5, 2 * vpc, 2 * vpc
    static = torch.randn(C * spc, 3, H, W, generator=g)
    dyn = torch.randn(C, dpc, T, 1, H, W, generator=g)
    w = torch.randn(3, 4, 3, 3, 3, generator=g) * 0.2
    b = torch.randn(3, generator=g) * 0.1
    label = torch.arange(C).repeat_interleave(vpc)
    idx = torch.arange(C * vpc) % vpc
    didx = 2 * idx + torch.randint(2, (C * vpc,), generator=g)
    sidx = spc * label + 2 * idx + torch.randint(2, (C * vpc,), generator=g)
    gout = torch.randn(C * vpc, T, 3, H, W, generator=g)

    def run(unique):
        d = dyn.cuda().requires_grad_(True)
        wc, bc = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
        out = ops.compose(static.cuda(), d, wc, bc, sidx.cuda(), label.cuda(), didx.cuda(), unique_rows=unique)
        out.backward(gout.cuda())
        return out.detach(), d.grad, wc.grad, bc.grad
    o1, gd1, gw1, gb1 = run(True)
    o2, gd2, gw2, gb2 = run(True)
    o3, gd3, gw3, gb3 = run(False)
    assert torch.equal(gw1, gw2) and torch.equal(gb1, gb2) and torch.equal(gd1, gd2)
    assert torch.equal(gd1, gd3) and torch.equal(gw1, gw3)            # atomics onto zeros with distinct rows: the same bits
    # fp64 reference
    d64 = dyn.double().requires_grad_(True)
    w64, b64 = w.double().requires_grad_(True), b.double().requires_grad_(True)
    x = torch.cat([static.double()[sidx].unsqueeze(2).expand(-1, -1, T, -1, -1), d64[label, didx].permute(0, 2, 1, 3, 4)], 1)
    ref = F.conv3d(x, w64, b64, padding=1).permute(0, 2, 1, 3, 4)
    ref.backward(gout.double())
    rel = lambda a, r: float((a.double().cpu() - r).norm() / r.norm())
    assert rel(o1, ref.detach()) < 1e-6
    assert rel(gd1, d64.grad) < 1e-6 and rel(gw1, w64.grad) < 1e-5 and rel(gb1, b64.grad) < 1e-5
