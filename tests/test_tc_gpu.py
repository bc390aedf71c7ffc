"""GPU parity of the tcgen05 tensor-core path (all calls go through the C ABI).

Order = bring-up order: packers (bit-exact vs the numpy model), raw accumulators of each layer
vs fp64 conv on the same bf16-rounded operands, fused epilogues (bias+ReLU+MaxPool+code, packed
output layouts), then the whole embed vs the oracle.
Tolerances: fp32 accumulation of bf16 products -> 2e-5 relative L2 on raw accumulators; pooled
bf16 outputs -> 1 bf16 ulp (4e-3 relative per element); embeddings vs the fp32 oracle on
bf16-representable inputs/weights -> 1e-2 relL2 (bf16 activations between layers).
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import tc_emulator as em

pytestmark = pytest.mark.gpu

DUMP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def conv_ref(x, w, bias=None):
    return F.conv3d(x.double().cpu(), w.double().cpu(), None if bias is None else bias.double().cpu(),
                    stride=(1, 2, 2), padding=(1, 3, 3)).float()


def dump(name, **arrs):
    os.makedirs(DUMP, exist_ok=True)
    np.savez_compressed(os.path.join(DUMP, name), **{k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in arrs.items()})


def make_net(T, HW, seed=0):
    from video_distillation_b200.tc import TcConvNet3D
    g = torch.Generator().manual_seed(seed)
    w0 = em.bf16_round(torch.randn(64, 3, 3, 7, 7, generator=g) * 0.08)
    w1 = em.bf16_round(torch.randn(128, 64, 3, 7, 7, generator=g) * 0.02)
    w2 = em.bf16_round(torch.randn(128, 128, 3, 7, 7, generator=g) * 0.02)
    b0 = torch.randn(64, generator=g) * 0.1
    b1 = torch.randn(128, generator=g) * 0.1
    b2 = torch.randn(128, generator=g) * 0.1
    net = TcConvNet3D(T, HW, HW, 'cuda')
    net.load_weights(*(t.cuda() for t in (w0, b0, w1, b1, w2, b2)))
    return net, (w0, b0, w1, b1, w2, b2)


CASES = [(8, 64), (4, 112)]


@pytest.mark.parametrize('T,HW', CASES)
def test_packers_bit_exact(T, HW):
    net, (w0, b0, w1, b1, w2, b2) = make_net(T, HW)
    g = em.Geo(T, HW)
    video = em.bf16_round(torch.randn(3, T, 3, HW, HW, generator=torch.Generator().manual_seed(1)))
    x0 = net.pack_video(video.cuda()).cpu().numpy().view(np.uint16)
    ref = em.pack_x0(video, g).reshape(-1)
    assert np.array_equal(x0[:ref.size], ref)
    # gather by index
    idx = torch.tensor([2, 0], device='cuda')
    x0i = net.pack_video(video.cuda(), index=idx).cpu().numpy().view(np.uint16)
    refi = em.pack_x0(video[[2, 0]], g).reshape(-1)
    assert np.array_equal(x0i[:refi.size], refi)
    assert np.array_equal(net.w0.cpu().numpy().view(np.uint16), em.pack_w0(w0).reshape(-1))
    assert np.array_equal(net.w1.cpu().numpy().view(np.uint16), em.pack_w1(w1).reshape(-1))
    assert np.array_equal(net.w2.cpu().numpy().view(np.uint16), em.pack_w2(w2).reshape(-1))


@pytest.mark.parametrize('T,HW', CASES)
def test_raw_layer2(T, HW):
    """Smallest configuration first: conv 2 (N = To2*HW2, 4 accumulators, streamed weights)."""
    net, (w0, b0, w1, b1, w2, b2) = make_net(T, HW)
    g = em.Geo(T, HW)
    B = 5
    x = em.bf16_round(torch.randn(B, 128, g.T2, g.H2, g.H2, generator=torch.Generator().manual_seed(2)))
    a2 = torch.from_numpy(em.pack_a2(x, g, Bpad=8).view(np.int16)).cuda()
    p = em.Params(2, T, HW, B)
    raw = torch.zeros(p.n_tiles, p.n_acc, 128, p.ncols, device='cuda')
    net.conv_layer(2, a2, net.w2, None, raw, B, raw=True)
    torch.cuda.synchronize()
    d = raw.cpu().reshape(-1, 128, g.To2, g.Ho2, g.Wo2)[:B]
    y = conv_ref(x, w2)
    r = rel(d, y)
    if r > 2e-5:
        dump(f'raw_l2_{T}_{HW}.npz', d=raw, x=x, w=w2)
    assert r < 2e-5, r


@pytest.mark.parametrize('T,HW', CASES)
def test_raw_layer1(T, HW):
    net, (w0, b0, w1, b1, w2, b2) = make_net(T, HW)
    g = em.Geo(T, HW)
    B = 2
    x = em.bf16_round(torch.randn(B, 64, T, g.H1, g.H1, generator=torch.Generator().manual_seed(3)))
    a1 = torch.from_numpy(em.pack_a1(x, g).view(np.int16)).cuda()
    p = em.Params(1, T, HW, B)
    raw = torch.zeros(p.n_tiles, p.n_acc, 128, p.ncols, device='cuda')
    net.conv_layer(1, a1, net.w1, None, raw, B, raw=True)
    torch.cuda.synchronize()
    d = raw.cpu().reshape(B, T // 2, 2, 128, g.Ho1, g.P1)[..., :g.Wo1]          # (B,tp,a,cout,ho,wo)
    d = d.permute(0, 3, 1, 2, 4, 5).reshape(B, 128, T, g.Ho1, g.Wo1)
    y = conv_ref(x, w1)
    r = rel(d, y)
    if r > 2e-5:
        dump(f'raw_l1_{T}_{HW}.npz', d=raw, x=x, w=w1)
    assert r < 2e-5, r


@pytest.mark.parametrize('T,HW', CASES)
def test_raw_layer0(T, HW):
    net, (w0, b0, w1, b1, w2, b2) = make_net(T, HW)
    g = em.Geo(T, HW)
    B = 2
    video = em.bf16_round(torch.randn(B, T, 3, HW, HW, generator=torch.Generator().manual_seed(4)))
    x0 = net.pack_video(video.cuda())
    p = em.Params(0, T, HW, B)
    raw = torch.zeros(p.n_tiles, p.n_acc, 128, p.ncols, device='cuda')
    net.conv_layer(0, x0, net.w0, None, raw, B, raw=True)
    torch.cuda.synchronize()
    nrb = g.Ho0 // g.R0
    d = raw.cpu().reshape(B, T // 2, nrb, 2, 64, g.R0, g.Wo0)                      # (B,tp,rb,f,cout,r,wo)
    d = d.permute(0, 4, 1, 3, 2, 5, 6).reshape(B, 64, T, g.Ho0, g.Wo0)
    y = conv_ref(video.permute(0, 2, 1, 3, 4), w0)
    r = rel(d, y)
    if r > 2e-5:
        dump(f'raw_l0_{T}_{HW}.npz', d=raw, video=video, w=w0)
    assert r < 2e-5, r


def _relu_pool(y, k):
    yp, idx = F.max_pool3d(F.relu(y), k, k, return_indices=True)
    return yp


@pytest.mark.parametrize('T,HW', CASES)
def test_fused_epilogues_and_embed(T, HW):
    net, (w0, b0, w1, b1, w2, b2) = make_net(T, HW)
    g = em.Geo(T, HW)
    B = 5
    video = em.bf16_round(torch.randn(B, T, 3, HW, HW, generator=torch.Generator().manual_seed(5)))
    emb, codes = net.embed(video.cuda(), want_codes=True)
    torch.cuda.synchronize()
    # layer-by-layer reference with the same bf16 rounding of the stored activations
    y0 = conv_ref(video.permute(0, 2, 1, 3, 4), w0, b0)
    p0 = em.bf16_round(_relu_pool(y0, (1, 2, 2)))
    a1 = em.unpack_a1(net._a1.cpu().numpy().view(np.uint16)[:B * g.video1 // 2], g, B)
    assert rel(a1, p0) < 3e-3, rel(a1, p0)
    # halo of A1 must still be zero: total energy equals the energy of the unpacked interior
    y1 = conv_ref(a1, w1, b1)
    p1 = em.bf16_round(_relu_pool(y1, (2, 2, 2)))
    a2, consistent = em.unpack_a2(net._a2.cpu().numpy().view(np.uint16)[:8 * g.video2 // 2], g, B)
    assert consistent
    assert rel(a2, p1) < 3e-3, rel(a2, p1)
    y2 = conv_ref(a2, w2, b2)
    p2 = _relu_pool(y2, (2, 2, 2)).reshape(B, -1)
    assert rel(emb, p2) < 1e-4, rel(emb, p2)
    # routing codes: active bit must agree with "pooled output > 0" wherever the margin is clear
    c0, c1, c2 = (c.cpu() for c in codes)
    act2 = (c2.reshape(B, -1) & 8) > 0
    clear = p2.abs() > 1e-3
    assert bool((act2[clear] == (p2[clear] > 0)).all())
    # argmax of layer 2 against torch's indices on the same input
    _, idx = F.max_pool3d(F.relu(y2), 2, 2, return_indices=True)
    To, Ho, Wo = y2.shape[2:]
    it, ih, iw = idx // (Ho * Wo), (idx // Wo) % Ho, idx % Wo
    pos = (it % 2) * 4 + (ih % 2) * 2 + (iw % 2)
    arg2 = (c2 & 7).long()
    agree = ((arg2 == pos) | ~act2.reshape(arg2.shape)).float().mean().item()
    assert agree > 0.995, agree
    # end to end against the fp32 oracle (no intermediate rounding): bf16-grade agreement
    from oracle import convnet3d_embed
    params = {'features.0.weight': w0, 'features.0.bias': b0, 'features.3.weight': w1, 'features.3.bias': b1,
              'features.6.weight': w2, 'features.6.bias': b2}
    e_or = convnet3d_embed(params, video)
    assert rel(emb, e_or) < 1e-2, rel(emb, e_or)


@pytest.mark.parametrize('T,HW', CASES)
def test_backward_matches_routed_fp32(T, HW):
    """d embed / d video on tensor cores vs the exact fp32 kernels CONDITIONED on the same routing codes
    (SURVEY §7.3 contract iii).  bf16 rounding of dY between layers -> 2e-2 relL2."""
    from video_distillation_b200 import ops
    net, (w0, b0, w1, b1, w2, b2) = make_net(T, HW)
    B = 3
    gen = torch.Generator().manual_seed(6)
    video = em.bf16_round(torch.randn(B, T, 3, HW, HW, generator=gen)).cuda()
    v = video.clone().requires_grad_(True)
    emb = net.embed_autograd(v)
    gemb = torch.randn(emb.shape, generator=gen).cuda()
    emb.backward(gemb)
    _, (c0, c1, c2) = net.embed(video, want_codes=True)
    S, P = (1, 2, 2), (1, 3, 3)
    x = video.clone().requires_grad_(True)
    y = ops.route_with_code(ops.conv3d(x.permute(0, 2, 1, 3, 4), w0.cuda(), b0.cuda(), S, P), c0, (1, 2, 2))
    y = ops.route_with_code(ops.conv3d(y, w1.cuda(), b1.cuda(), S, P), c1, (2, 2, 2))
    y = ops.route_with_code(ops.conv3d(y, w2.cuda(), b2.cuda(), S, P), c2, (2, 2, 2))
    y.reshape(B, -1).backward(gemb)
    r = rel(v.grad, x.grad)
    if r > 2e-2:
        dump(f'bwd_{T}_{HW}.npz', got=v.grad, want=x.grad)
    assert r < 2e-2, r


def test_resident_prepacked_dataset_matches_per_step_packing():
    """item_index of conv 0 over a pre-packed resident set == packing the gathered videos each step."""
    T, HW = 8, 64
    net, _ = make_net(T, HW)
    videos = torch.randn(7, T, 3, HW, HW, generator=torch.Generator().manual_seed(8)).cuda()
    idx = torch.tensor([5, 0, 3, 3, 6], device='cuda')
    x0_all = net.pack_dataset(videos, chunk=3)
    a = net.embed_resident(x0_all, idx)
    b = net.embed(videos, index=idx)
    assert torch.equal(a, b)


def test_joint_real_and_synthetic_pass_matches_separate_passes():
    """embed_joint (real videos of the resident set + synthetic videos in the spare slots, one launch per layer,
    codes only for the synthetic tail) == separate real / synthetic embeds, bit for bit, across chunk borders."""
    T, HW = 8, 64
    net, _ = make_net(T, HW)
    net.max_batch = 8                                   # force several chunks, one of them straddling real|syn
    gen = torch.Generator().manual_seed(9)
    videos = torch.randn(9, T, 3, HW, HW, generator=gen).cuda()
    syn = torch.randn(5, T, 3, HW, HW, generator=gen).cuda()
    idx = torch.tensor([5, 0, 3, 3, 6, 8, 1, 2, 7, 7, 4], device='cuda')
    x0_all = net.pack_dataset(videos, chunk=4, extra_slots=5)
    e_real, e_syn, codes = net.embed_joint(x0_all, idx, syn, 9)
    ref_real = net.embed(videos, index=idx)
    ref_syn, ref_codes = net.embed(syn, want_codes=True)
    assert torch.equal(e_real, ref_real)
    assert torch.equal(e_syn, ref_syn)
    for a, b in zip(codes, ref_codes):
        assert torch.equal(a, b)
    # and the autograd wrapper routes the gradient to the synthetic videos only
    v = syn.clone().requires_grad_(True)
    er, es = net.embed_joint_autograd(x0_all, idx, v, 9)
    assert not er.requires_grad
    g = torch.randn(es.shape, generator=gen).cuda()
    es.backward(g)
    v2 = syn.clone().requires_grad_(True)
    net.embed_autograd(v2).backward(g)
    assert torch.equal(v.grad, v2.grad)


@pytest.mark.parametrize('T,HW', CASES)
def test_tensor_core_trio_matches_fp32_kernels(T, HW):
    """fprop / dgrad / wgrad of the three feature convolutions on tensor cores (bf16 operands, fp32 accumulate)
    against the exact fp32 kernels on bf16-rounded operands (same products, different summation order)."""
    from video_distillation_b200 import ops
    S, P = (1, 2, 2), (1, 3, 3)
    gen = torch.Generator().manual_seed(12)
    B = 3
    shapes = [(3, 64, (T, HW, HW)), (64, 128, (T, HW // 4, HW // 4)), (128, 128, (T // 2, HW // 16, HW // 16))]
    for layer, (cin, cout, ext) in enumerate(shapes):
        x = em.bf16_round(torch.randn(B, cin, *ext, generator=gen)).cuda()
        w = em.bf16_round(torch.randn(cout, cin, 3, 7, 7, generator=gen) / (cin * 147) ** 0.5).cuda()
        prev = ops.set_conv_backend('fp32')
        try:
            y_ref = ops.conv3d_fprop_raw(x, w, None, S, P)
            gy = em.bf16_round(torch.randn(y_ref.shape, generator=gen)).cuda()
            gx_ref = ops.conv3d_dgrad_raw(gy, w, tuple(x.shape), S, P)
            gw_ref = ops.conv3d_wgrad_raw(x, gy, tuple(w.shape), S, P)
            ops.set_conv_backend('tc')
            y = ops.conv3d_fprop_raw(x, w, None, S, P)                       # bf16 pairs, three launches (any operand range)
            y16 = ops.conv3d_fprop_raw(x, w, None, S, P, fp16_ok=True)       # conv 1 / 2: fp16 pairs, one launch (activation ranges)
            gx = ops.conv3d_dgrad_raw(gy, w, tuple(x.shape), S, P)
            gw = ops.conv3d_wgrad_raw(x, gy, tuple(w.shape), S, P)
            # the parity-grade trio: every primitive on bf16 hi / lo pairs, on operands that are NOT bf16-representable
            x3, w3, gy3 = x * 1.0009765625, w * 0.9990234375, gy * 1.0029296875
            ops.set_conv_backend('fp32')
            ref3 = (ops.conv3d_fprop_raw(x3, w3, None, S, P), ops.conv3d_dgrad_raw(gy3, w3, tuple(x.shape), S, P),
                    ops.conv3d_wgrad_raw(x3, gy3, tuple(w.shape), S, P))
            ops.set_conv_backend('tc_x3')
            got3 = (ops.conv3d_fprop_raw(x3, w3, None, S, P), ops.conv3d_dgrad_raw(gy3, w3, tuple(x.shape), S, P),
                    ops.conv3d_wgrad_raw(x3, gy3, tuple(w.shape), S, P))
            # cotangent-sized operands (1e-6): the bf16 pairs keep their relative accuracy
            ops.set_conv_backend('fp32')
            y_small_ref = ops.conv3d_fprop_raw(x3 * 1e-6, w3 * 1e-3, None, S, P)
            ops.set_conv_backend('tc_x3')
            y_small = ops.conv3d_fprop_raw(x3 * 1e-6, w3 * 1e-3, None, S, P)
        finally:
            ops.set_conv_backend(prev)
        assert y.shape == y_ref.shape and gx.shape == gx_ref.shape and gw.shape == gw_ref.shape
        assert rel(y, y_ref) < 1e-5, (layer, 'fprop', rel(y, y_ref))
        assert rel(y16, y_ref) < 1e-5, (layer, 'fprop fp16 pairs', rel(y16, y_ref))
        errs3 = [rel(a, b) for a, b in zip(got3, ref3)] + [rel(y_small, y_small_ref)]
        print(f'layer {layer} tc_x3 trio vs fp32 kernels (fprop, dgrad, wgrad, fprop of 1e-6-sized operands):', errs3)
        assert max(errs3) < 1e-4, (layer, "tc_x3", errs3)
        # conv 1: column-free dgrad, fp32 accumulators straight to the output; conv 0 / 2: fp32 column buffers
        assert rel(gx, gx_ref) < 1e-5, (layer, 'dgrad', rel(gx, gx_ref))
        assert rel(gw, gw_ref) < 1e-5, (layer, 'wgrad', rel(gw, gw_ref))


@pytest.mark.parametrize('T,HW', CASES)
def test_uint8_packer_fuses_the_dataset_normalisation_bit_exactly(T, HW):
    """vd_tc_pack_video_u8(frames) == vd_tc_pack_video((frames / 255 - mean) / std): the packed bf16 operands (and hence
    every embedding) are identical whether the normalised fp32 videos or the raw uint8 frames are handed over."""
    from video_distillation_b200.tc import TcConvNet3D
    tc = TcConvNet3D(T, HW, HW, 'cuda', max_batch=8)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    tc.set_normalization(mean, std)
    g = torch.Generator(device='cuda').manual_seed(T * HW)
    frames = torch.randint(0, 256, (5, T, 3, HW, HW), dtype=torch.uint8, device='cuda', generator=g)
    m = torch.tensor(mean, device='cuda').view(1, 1, 3, 1, 1)
    s = torch.tensor(std, device='cuda').view(1, 1, 3, 1, 1)
    video = ((frames.float() / 255.0) - m) / s
    idx = torch.tensor([3, 0, 4, 4, 1], device='cuda')
    n = 5 * tc.plan.x0_bytes_per_video
    a = tc.pack_video(video, idx, out=torch.empty(n, dtype=torch.uint8, device='cuda')).clone()
    b = tc.pack_video(frames, idx, out=torch.empty(n, dtype=torch.uint8, device='cuda')).clone()
    assert torch.equal(a, b)


def test_uint8_resident_dataset_matches_float_dataset():
    """DeviceDataset over uint8 frames (+ norm) prepacks to the same operand and samples the same normalised videos."""
    import numpy as np
    from video_distillation_b200.distill import DeviceDataset
    from video_distillation_b200.tc import TcConvNet3D
    T, HW, C, per = 8, 64, 2, 3
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    frames = torch.randint(0, 256, (C * per, T, 3, HW, HW), dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
    video = ((frames.float() / 255.0) - torch.tensor(mean).view(1, 1, 3, 1, 1)) / torch.tensor(std).view(1, 1, 3, 1, 1)
    labels = [c for c in range(C) for _ in range(per)]
    tc = TcConvNet3D(T, HW, HW, 'cuda', max_batch=8)
    d8 = DeviceDataset(frames, labels, C, 'cuda', norm=(mean, std)).prepack(tc)
    df = DeviceDataset(video, labels, C, 'cuda').prepack(tc)
    assert torch.equal(d8.x0, df.x0)
    np.random.seed(2)
    a = d8.get_images(1, 2)
    np.random.seed(2)
    b = df.get_images(1, 2)
    assert torch.equal(a, b)


@pytest.mark.parametrize('T,HW,B', [(8, 64, 3), (16, 64, 11), (24, 64, 2), (32, 64, 2), (4, 112, 3), (12, 112, 3), (16, 112, 5)])
def test_embed_geometry_sweep_against_the_fp32_oracle(T, HW, B):
    """Every supported clip length / frame size through the fused tensor-core pipeline (streaming conv 0, dual-issuer
    conv 1 / 2) against the CPU oracle: bf16-grade agreement of the embeddings, finite non-zero input gradients."""
    from oracle import convnet3d_embed
    net, (w0, b0, w1, b1, w2, b2) = make_net(T, HW)
    video = em.bf16_round(torch.randn(B, T, 3, HW, HW, generator=torch.Generator().manual_seed(T + HW + B)))
    params = {'features.0.weight': w0, 'features.0.bias': b0, 'features.3.weight': w1, 'features.3.bias': b1,
              'features.6.weight': w2, 'features.6.bias': b2}
    e_or = convnet3d_embed(params, video)
    v = video.cuda().requires_grad_(True)
    emb = net.embed_autograd(v)
    assert rel(emb, e_or) < 1e-2, rel(emb, e_or)
    emb.square().sum().backward()
    assert torch.isfinite(v.grad).all() and v.grad.abs().sum() > 0
    # gradient against torch autograd through the oracle on the same (bf16-rounded) video: routing differs on a few
    # near-ties only, so the bulk of the gradient must agree
    vc = video.clone().requires_grad_(True)
    convnet3d_embed(params, vc).square().sum().backward()
    assert rel(v.grad, vc.grad) < 0.4, rel(v.grad, vc.grad)


@pytest.mark.parametrize('T,HW', [(8, 64), (16, 112)])
def test_batch_size_sweep_is_bitwise_consistent(T, HW):
    """Tile counts below, at and above the grid size (148 persistent CTAs), odd batch sizes, partially filled conv-2 tiles:
    the embedding of a video is bitwise independent of the batch it is embedded with."""
    net, _ = make_net(T, HW)
    nmax = 301 if HW == 64 else 75
    video = torch.randn(nmax, T, 3, HW, HW, device='cuda', generator=torch.Generator(device='cuda').manual_seed(T))
    ref_first = net.embed(video[:1]).clone()
    for B in ([1, 2, 7, 37, 149, 300, 301] if HW == 64 else [1, 3, 10, 19, 75]):
        e = net.embed(video[:B]).clone()
        assert torch.equal(e[:1], ref_first), B
        assert torch.equal(e[B - 1:B], net.embed(video[B - 1:B])), B
