"""CPU validation of the tensor-core path's layouts and launch tables (no GPU needed).

The emulator (tests/tc_emulator.py) replays ws_gemm_kernel's data movement with the real
parameters from vd_tc_debug_params; the result must equal F.conv3d on the same bf16-rounded
operands.  This pins everything except the hardware's own reading of the UMMA descriptors.
"""
import pytest
import torch
import torch.nn.functional as F

import tc_emulator as em

torch.manual_seed(0)


def conv_ref(x, w):
    return F.conv3d(x.double(), w.double(), None, stride=(1, 2, 2), padding=(1, 3, 3)).float()


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


@pytest.mark.parametrize('T,HW,tiles', [(8, 64, [0, 3, 5, 31]), (4, 112, [0, 13, 14, 27])])
def test_layer0_tables(T, HW, tiles):
    g = em.Geo(T, HW)
    B = 2 if HW == 64 else 1
    video = em.bf16_round(torch.randn(B, T, 3, HW, HW))
    w = em.bf16_round(torch.randn(64, 3, 3, 7, 7) * 0.1)
    y = conv_ref(video.permute(0, 2, 1, 3, 4), w)                      # (B,64,T,Ho0,Wo0)
    D, p = em.emulate_layer(0, em.pack_x0(video, g), em.pack_w0(w), T, HW, B, tiles)
    assert p.ncols == g.N0 and p.n_acc == 1
    tl = list(range(p.n_tiles)) if tiles is None else tiles
    for k, tile in enumerate(tl):
        item, sub = divmod(tile, p.tiles_per_item)
        tp, rb = divmod(sub, p.v_count)
        d = torch.from_numpy(D[k, 0]).reshape(2, 64, g.R0, g.Wo0)      # (frame in pair, cout, r, wo)
        ref = y[item, :, 2 * tp:2 * tp + 2, rb * g.R0:(rb + 1) * g.R0, :].permute(1, 0, 2, 3)
        assert rel(d, ref) < 1e-5, (tile, rel(d, ref))


@pytest.mark.parametrize('T,HW,tiles', [(8, 64, [0, 3]), (4, 112, [1])])
def test_layer1_tables(T, HW, tiles):
    g = em.Geo(T, HW)
    B = 1
    x = em.bf16_round(torch.randn(B, 64, T, g.H1, g.H1))
    w = em.bf16_round(torch.randn(128, 64, 3, 7, 7) * 0.05)
    y = conv_ref(x, w)                                                 # (B,128,T,Ho1,Wo1)
    a1 = em.pack_a1(x, g)
    assert torch.equal(em.unpack_a1(a1, g, B), x)
    D, p = em.emulate_layer(1, a1, em.pack_w1(w), T, HW, B, tiles)
    assert p.ncols == g.N1 and p.n_acc == 2
    tl = list(range(p.n_tiles)) if tiles is None else tiles
    for k, tile in enumerate(tl):
        item, tp = divmod(tile, p.tiles_per_item)
        d = torch.from_numpy(D[k]).reshape(2, 128, g.Ho1, g.P1)[:, :, :, :g.Wo1]
        ref = y[item, :, 2 * tp:2 * tp + 2].permute(1, 0, 2, 3)
        assert rel(d, ref) < 1e-5, (tile, rel(d, ref))


@pytest.mark.parametrize('T,HW', [(8, 64), (8, 112)])
def test_layer2_tables(T, HW):
    g = em.Geo(T, HW)
    B = 3                                                              # partial last tile (4 videos per tile)
    x = em.bf16_round(torch.randn(B, 128, g.T2, g.H2, g.H2))
    w = em.bf16_round(torch.randn(128, 128, 3, 7, 7) * 0.05)
    y = conv_ref(x, w)                                                 # (B,128,To2,Ho2,Wo2)
    a2 = em.pack_a2(x, g, Bpad=4)
    back, ok = em.unpack_a2(a2, g, B)
    assert ok and torch.equal(back, x)
    D, p = em.emulate_layer(2, a2, em.pack_w2(w), T, HW, B)
    assert p.ncols == g.N2 and p.n_acc == 4 and p.n_tiles == 1
    d = torch.from_numpy(D[0]).reshape(4, 128, g.To2, g.Ho2, g.Wo2)
    assert rel(d[:B], y) < 1e-5, rel(d[:B], y)
    assert float(d[B:].abs().max()) == 0.0


def test_plan_sizes():
    for T, HW, emb in [(16, 112, 2048), (8, 64, 256)]:
        p = em.Params(0, T, HW, 1).plan
        g = em.Geo(T, HW)
        assert p.embed_dim == emb == g.embed_dim
        assert p.x0_bytes_per_video == g.video0 and p.a1_bytes_per_video == g.video1
        assert p.a2_bytes_per_video == g.video2
        for layer in range(3):
            q = em.Params(layer, T, HW, 8)
            assert q.smem_total <= 232448 and q.smem_total >= 120 * 1024     # one CTA per SM (512 TMEM cols each)
            assert q.n_acc * q.acc_cols * q.acc_stages <= 512
            assert q.ncols % 16 == 0 and 16 <= q.ncols <= 256


def test_python_and_library_agree_on_supported_geometries():
    import ctypes
    from video_distillation_b200 import _lib
    from video_distillation_b200.tc import tc_supported
    for HW in (32, 64, 96, 112, 128):
        for T in range(1, 41):
            plan = _lib.TcPlan()
            ok = _lib.lib().vd_tc_plan_make(ctypes.byref(plan), T, HW, HW) == 0
            assert ok == tc_supported(T, HW, HW), (T, HW, ok)


@pytest.mark.parametrize('T,HW,cols', [(8, 64, [0, 3, 5]), (4, 112, [0, 13, 14])])
def test_layer0_streaming_schedule(T, HW, cols):
    """The input-frame streaming order of conv 0 (one staged frame -> two frame-pair accumulators through the four
    Toeplitz windows) reproduces the convolution with the real launch tables."""
    g = em.Geo(T, HW)
    B = 2
    video = em.bf16_round(torch.randn(B, T, 3, HW, HW, generator=torch.Generator().manual_seed(T)))
    w = em.bf16_round(torch.randn(64, 3, 3, 7, 7, generator=torch.Generator().manual_seed(1)) * 0.1)
    y = conv_ref(video.permute(0, 2, 1, 3, 4), w)                      # (B,64,T,Ho0,Wo0)
    out, p = em.emulate_layer0_streaming(em.pack_x0(video, g), em.pack_w0(w), T, HW, B, cols)
    for (item, rb), D in out.items():
        for tp in range(T // 2):
            d = torch.from_numpy(D[tp]).reshape(2, 64, g.R0, g.Wo0)
            ref = y[item, :, 2 * tp:2 * tp + 2, rb * g.R0:(rb + 1) * g.R0, :].permute(1, 0, 2, 3)
            assert rel(d, ref) < 1e-5, (item, rb, tp, rel(d, ref))
