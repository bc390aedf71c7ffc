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


# ------------------------------------------------------------------ split-fp16 forward (SGeo): tables + layouts
@pytest.mark.parametrize('T,HW,cols', [(8, 64, [0, 5, 15]), (4, 112, [0, 13, 27, 55])])
def test_split_layer0_tables_and_schedule(T, HW, cols):
    g = em.Geo(T, HW)
    sg = em.SGeo(g)
    B = 2
    gen = torch.Generator().manual_seed(T)
    video = torch.randn(B, T, 3, HW, HW, generator=gen)
    w = torch.randn(64, 3, 3, 7, 7, generator=gen) * 0.05
    y = conv_ref(em.f16x2_round(video).permute(0, 2, 1, 3, 4), em.f16x2_round(w))     # (B,64,T,Ho0,Wo0)
    out, p = em.emulate_layer0s(em.pack_x0s(video, g), em.pack_w0s(w), T, HW, B, cols)
    assert p.ncols == sg.N0s and p.n_acc == 1 and p.acc_cols * p.acc_stages <= 512 and p.smem_total <= 232448
    for (item, rb), D in out.items():
        d = torch.from_numpy(em.unstack_l0s(D)).reshape(T, 64, sg.R0s, g.Wo0)
        ref = y[item, :, :, rb * sg.R0s:(rb + 1) * sg.R0s, :].permute(1, 0, 2, 3)
        assert rel(d, ref) < 2e-6, (item, rb, rel(d, ref))


@pytest.mark.parametrize('T,HW,tiles', [(8, 64, [0, 1]), (4, 112, [1])])
def test_split_layer1_tables(T, HW, tiles):
    g = em.Geo(T, HW)
    B = 1
    gen = torch.Generator().manual_seed(HW)
    x = torch.randn(B, 64, T, g.H1, g.H1, generator=gen).abs()
    w = torch.randn(128, 64, 3, 7, 7, generator=gen) * 0.02
    xr, wr = em.f16x2_round(x), em.f16x2_round(w)
    # the split path drops the lo*lo term: compare against hi*hi + lo*hi + hi*lo
    xh = em.from_f16_bits(em.split_f16(x)[0]); wh = em.from_f16_bits(em.split_f16(w)[0])
    y = conv_ref(xh, wh) + conv_ref(xr - xh, wh) + conv_ref(xh, wr - wh)
    a1 = em.pack_a1s(x, g)
    assert torch.equal(em.unpack_a1s(a1, g, B), xr)
    D, p = em.emulate_layer(7, a1, em.pack_w1s(w), T, HW, B, tiles, fmt='f16')
    assert p.ncols == g.N1 and p.n_steps == 75 and p.n_wtiles == 50 and p.smem_total <= 232448 and p.n_acc * p.acc_cols <= 512
    fpt = p.n_acc
    for k, tile in enumerate(tiles):
        item, tq = divmod(tile, p.tiles_per_item)
        d = torch.from_numpy(D[k]).reshape(fpt, 128, g.Ho1, g.P1)[:, :, :, :g.Wo1]
        ref = y[item, :, fpt * tq:fpt * tq + fpt].permute(1, 0, 2, 3)
        assert rel(d, ref) < 2e-6, (tile, rel(d, ref))
        assert rel(d, conv_ref(xr, wr)[item, :, fpt * tq:fpt * tq + fpt].permute(1, 0, 2, 3)) < 1e-5


@pytest.mark.parametrize('T,HW', [(8, 64), (8, 112)])
def test_split_layer2_tables(T, HW):
    g = em.Geo(T, HW)
    B = 3
    gen = torch.Generator().manual_seed(HW + 1)
    x = torch.randn(B, 128, g.T2, g.H2, g.H2, generator=gen).abs()
    w = torch.randn(128, 128, 3, 7, 7, generator=gen) * 0.02
    xr, wr = em.f16x2_round(x), em.f16x2_round(w)
    xh = em.from_f16_bits(em.split_f16(x)[0]); wh = em.from_f16_bits(em.split_f16(w)[0])
    y = conv_ref(xh, wh) + conv_ref(xr - xh, wh) + conv_ref(xh, wr - wh)
    a2 = em.pack_a2s(x, g, Bpad=4)
    back, ok = em.unpack_a2s(a2, g, B)
    assert ok and torch.equal(back, xr)
    D, p = em.emulate_layer(8, a2, em.pack_w2s(w), T, HW, B, fmt='f16')
    assert p.ncols == g.N2 and p.n_acc == 4 and p.n_tiles == 1 and p.n_steps == 18 and p.n_wtiles == 12
    d = torch.from_numpy(D[0]).reshape(4, 128, g.To2, g.Ho2, g.Wo2)
    assert rel(d[:B], y) < 2e-6, rel(d[:B], y)
    assert float(d[B:].abs().max()) == 0.0


# ------------------------------------------------------------------ two-product mode of the frozen real videos (passes = 2)
# y = xh*wh + xh*wl on the hi part of the operands only (debug layers 9 / 10 / 11): same layouts and weight images
@pytest.mark.parametrize('T,HW,cols', [(8, 64, [0, 15]), (4, 112, [0, 27, 55])])
def test_two_product_layer0_tables_and_schedule(T, HW, cols):
    g = em.Geo(T, HW)
    sg = em.SGeo(g)
    B = 2
    gen = torch.Generator().manual_seed(T + 7)
    video = torch.randn(B, T, 3, HW, HW, generator=gen)
    w = torch.randn(64, 3, 3, 7, 7, generator=gen) * 0.05
    xh = em.from_f16_bits(em.split_f16(video)[0])
    y = conv_ref(xh.permute(0, 2, 1, 3, 4), em.f16x2_round(w))
    out, p = em.emulate_layer0s(em.pack_x0h(video, g), em.pack_w0s(w), T, HW, B, cols, layer=9)
    assert p.n_sb == 1 and p.item_stride == (T + 2) * 6 * g.plane0 and p.ncols == sg.N0s and p.smem_total <= 232448
    for (item, rb), D in out.items():
        d = torch.from_numpy(em.unstack_l0s(D)).reshape(T, 64, sg.R0s, g.Wo0)
        ref = y[item, :, :, rb * sg.R0s:(rb + 1) * sg.R0s, :].permute(1, 0, 2, 3)
        assert rel(d, ref) < 2e-6, (item, rb, rel(d, ref))


@pytest.mark.parametrize('T,HW,tiles', [(8, 64, [0, 1]), (4, 112, [1])])
def test_two_product_layer1_tables(T, HW, tiles):
    g = em.Geo(T, HW)
    B = 1
    gen = torch.Generator().manual_seed(HW + 3)
    x = torch.randn(B, 64, T, g.H1, g.H1, generator=gen).abs()
    w = torch.randn(128, 64, 3, 7, 7, generator=gen) * 0.02
    xh = em.from_f16_bits(em.split_f16(x)[0])
    y = conv_ref(xh, em.f16x2_round(w))
    a1 = em.pack_a1s(x, g)
    a1[:, :, :, 1] = 0x7e00                       # the lo planes are never read (NaN if they were)
    D, p = em.emulate_layer(10, a1, em.pack_w1s(w), T, HW, B, tiles, fmt='f16')
    assert p.n_steps == 50 and p.Gt == 0 and p.smem_total <= 232448 and p.n_acc * p.acc_cols <= 512
    assert p.stage_bytes == p.n_acc * 4 * g.plane1
    fpt = p.n_acc
    for k, tile in enumerate(tiles):
        item, tq = divmod(tile, p.tiles_per_item)
        d = torch.from_numpy(D[k]).reshape(fpt, 128, g.Ho1, g.P1)[:, :, :, :g.Wo1]
        ref = y[item, :, fpt * tq:fpt * tq + fpt].permute(1, 0, 2, 3)
        assert rel(d, ref) < 2e-6, (tile, rel(d, ref))


@pytest.mark.parametrize('T,HW', [(8, 64), (8, 112)])
def test_two_product_layer2_tables(T, HW):
    g = em.Geo(T, HW)
    B = 3
    gen = torch.Generator().manual_seed(HW + 5)
    x = torch.randn(B, 128, g.T2, g.H2, g.H2, generator=gen).abs()
    w = torch.randn(128, 128, 3, 7, 7, generator=gen) * 0.02
    xh = em.from_f16_bits(em.split_f16(x)[0])
    y = conv_ref(xh, em.f16x2_round(w))
    a2 = em.pack_a2s(x, g, Bpad=4)
    a2.reshape(4, 49, 4, 2, -1)[:, :, :, 1] = 0x7e00      # lo chunks [video][khw][quarter][part][...] are never read
    D, p = em.emulate_layer(11, a2, em.pack_w2s(w), T, HW, B, fmt='f16')
    assert p.n_acc == 4 and p.n_tiles == 1 and p.n_steps == 12 and p.Gt == 0
    d = torch.from_numpy(D[0]).reshape(4, 128, g.To2, g.Ho2, g.Wo2)
    assert rel(d[:B], y) < 2e-6, rel(d[:B], y)


def test_wgrad_plan_kt_split_sizes():
    """vd_tc_wgrad_plan in the kt-split mode (default): one im2col of the 49 spatial taps (a third of the column tiles), three gy
    images / raw buffers (one per temporal tap); host-only arithmetic."""
    import ctypes
    from video_distillation_b200 import _lib
    lib = _lib.lib()
    plan = _lib.TcPlan()
    _lib.check(lib.vd_tc_plan_make(ctypes.byref(plan), 16, 112, 112), 'plan')
    cin = (3, 64, 128)
    pixels = (16 * 56 * 56, 16 * 14 * 14, 8 * 4 * 4)
    for layer in range(3):
        assert lib.vd_tc_wgrad_kt_mode(layer) == 1
        sz = (ctypes.c_int64 * 6)()
        _lib.check(lib.vd_tc_wgrad_plan(layer, ctypes.byref(plan), 50, sz), 'wgrad_plan')
        splits, sps, ntiles, xcol, gyimg, raw = (int(v) for v in sz)
        assert ntiles == -(-cin[layer] * 49 // 256)
        stages = -(-50 * pixels[layer] // 128)
        assert stages <= splits * sps < stages + splits                           # split-K slices cover every 128-pixel stage, < 1 padded stage per slice
        assert xcol == ntiles * splits * sps * 65536
        assert gyimg == 3 * splits * sps * 8 * 4096 and raw == 3 * ntiles * splits * 128 * 256 * 4
