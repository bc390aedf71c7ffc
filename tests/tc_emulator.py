"""CPU model of the tensor-core path's data movement (test infrastructure).

* numpy packers / unpackers for the X0, A1, A2 activation layouts and the three weight images
  (mirror video_distillation_b200/csrc/tc_pack.cu and the fused epilogues in tc_conv.cu);
* ``emulate_layer``: replays one launch of ws_gemm_kernel exactly as the hardware is expected
  to execute it — bulk copies into a shared-memory image, then for every step a K=16 MMA whose
  operands are fetched through UMMA K-major/no-swizzle descriptors (start, LBO, SBO) — using
  the REAL launch parameters exported by vd_tc_debug_params.  It validates the host-side
  tables and layouts without a GPU; what it cannot validate is the hardware's reading of the
  descriptors, which tests/test_tc_gpu.py covers.
"""
import ctypes

import numpy as np
import torch

from video_distillation_b200 import _lib

MAX_COPIES, MAX_STEPS = 8, 80


def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


def to_bf16_bits(x):
    """float32 torch tensor -> uint16 numpy (round-to-nearest-even bf16 bit patterns)."""
    return x.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def from_bf16_bits(u):
    return torch.from_numpy(u.astype(np.uint16).view(np.int16).copy()).view(torch.bfloat16).to(torch.float32)


class Geo:
    def __init__(self, T, HW):
        self.T, self.HW = T, HW
        self.Ho0 = self.Wo0 = HW // 2
        self.R0 = 8 if self.Wo0 * 8 <= 256 else 4
        self.RI0 = self.Ho0 + 3
        self.N0 = self.R0 * self.Wo0
        self.plane0 = self.RI0 * self.Wo0 * 16
        self.frame0 = 6 * self.plane0
        self.video0 = (T + 2) * self.frame0
        self.H1 = self.Ho0 // 2
        self.Ho1 = self.Wo1 = self.H1 // 2
        self.P1, self.RI1 = self.Wo1 + 2, self.Ho1 + 4
        self.N1 = self.Ho1 * self.P1
        self.plane1 = self.RI1 * self.P1 * 16
        self.frame1 = 8 * self.plane1
        self.slice1 = (T + 2) * self.frame1
        self.video1 = 4 * self.slice1
        self.T2, self.H2 = T // 2, self.Ho1 // 2
        self.To2 = self.T2
        self.Ho2 = self.Wo2 = (self.H2 - 1) // 2 + 1
        self.HW2 = self.Ho2 * self.Wo2
        self.N2 = self.To2 * self.HW2
        self.chunk2 = (self.To2 + 2) * self.HW2 * 16
        self.group2 = 8 * self.chunk2
        self.video2 = 98 * self.group2
        self.T3p, self.H3p = self.To2 // 2, self.Ho2 // 2
        self.embed_dim = 128 * self.T3p * self.H3p * self.H3p


def tap_par(k):
    return (k + 1) & 1


def tap_shift(k):
    return (k - 1) // 2 if k & 1 else k // 2


def coord_par(x):
    return x & 1


def coord_pos(x):
    return (x + 3) // 2 if x & 1 else x // 2 + 1


def l0_chunk_kh(idx):
    return 2 * idx + 1 if idx < 3 else 2 * (idx - 3)


# ------------------------------------------------------------------ packers (uint16 bf16 bit images)
def pack_x0(video, g):
    """video (B,T,3,H,W) float32 -> uint16 (B, T+2, 3, 2, RI0, Wo0, 8)."""
    B = video.shape[0]
    bits = to_bf16_bits(video)                                    # (B,T,3,H,W)
    out = np.zeros((B, g.T + 2, 3, 2, g.RI0, g.Wo0, 8), np.uint16)
    HW = g.HW
    for par in range(2):
        for row in range(g.RI0):
            h = 2 * row - 3 if par else 2 * row - 2
            if not (0 <= h < HW):
                continue
            for kw in range(7):
                wo = np.arange(g.Wo0)
                w = 2 * wo + kw - 3
                ok = (w >= 0) & (w < HW)
                out[:, 1:g.T + 1, :, par, row, wo[ok], kw] = bits[:, :, :, h, w[ok]]
    return out


def pack_a1(x, g):
    """pooled conv-0 output (B,64,T,H1,H1) float32 -> uint16 (B,4,T+2,2,2,2,RI1,P1,8)."""
    B = x.shape[0]
    bits = to_bf16_bits(x)
    out = np.zeros((B, 4, g.T + 2, 2, 2, 2, g.RI1, g.P1, 8), np.uint16)
    for h in range(g.H1):
        for w in range(g.H1):
            v = bits[:, :, :, h, w].reshape(B, 4, 2, 8, g.T)      # (B, slice, k, e, T)
            out[:, :, 1:g.T + 1, coord_par(h), coord_par(w), :, coord_pos(h), coord_pos(w), :] = \
                v.transpose(0, 1, 4, 2, 3)
    return out


def unpack_a1(buf, g, B):
    """inverse of pack_a1 on a uint16 image -> float32 (B,64,T,H1,H1); also returns the halo mask check."""
    a = buf.reshape(B, 4, g.T + 2, 2, 2, 2, g.RI1, g.P1, 8)
    out = np.zeros((B, 64, g.T, g.H1, g.H1), np.uint16)
    for h in range(g.H1):
        for w in range(g.H1):
            v = a[:, :, 1:g.T + 1, coord_par(h), coord_par(w), :, coord_pos(h), coord_pos(w), :]   # (B,4,T,2,8)
            out[:, :, :, h, w] = v.transpose(0, 1, 3, 4, 2).reshape(B, 64, g.T)
    return from_bf16_bits(out)


def pack_a2(x, g, Bpad=None):
    """pooled conv-1 output (B,128,T2,H2,H2) float32 -> uint16 (Bpad,49,2,8,To2+2,Ho2,Wo2,8)."""
    B = x.shape[0]
    Bpad = Bpad or B
    bits = to_bf16_bits(x)
    out = np.zeros((Bpad, 49, 2, 8, g.To2 + 2, g.Ho2, g.Wo2, 8), np.uint16)
    for kh in range(7):
        for kw in range(7):
            for ho in range(g.Ho2):
                h = 2 * ho + kh - 3
                if not (0 <= h < g.H2):
                    continue
                for wo in range(g.Wo2):
                    w = 2 * wo + kw - 3
                    if not (0 <= w < g.H2):
                        continue
                    v = bits[:, :, :, h, w].reshape(B, 2, 8, 8, g.T2)          # (B, half, k, e, T2)
                    out[:B, kh * 7 + kw, :, :, 1:g.T2 + 1, ho, wo, :] = v.transpose(0, 1, 2, 4, 3)
    return out


def unpack_a2(buf, g, B):
    """Recover (B,128,T2,H2,H2) from the tap-expanded image and check that all copies agree."""
    a = buf.reshape(-1, 49, 2, 8, g.To2 + 2, g.Ho2, g.Wo2, 8)
    out = np.zeros((B, 128, g.T2, g.H2, g.H2), np.uint16)
    seen = np.zeros((g.H2, g.H2), bool)
    consistent = True
    for kh in range(7):
        for kw in range(7):
            for ho in range(g.Ho2):
                h = 2 * ho + kh - 3
                if not (0 <= h < g.H2):
                    continue
                for wo in range(g.Wo2):
                    w = 2 * wo + kw - 3
                    if not (0 <= w < g.H2):
                        continue
                    v = a[:B, kh * 7 + kw, :, :, 1:g.T2 + 1, ho, wo, :]            # (B,2,8,T2,8)
                    v = v.transpose(0, 1, 2, 4, 3).reshape(B, 128, g.T2)
                    if seen[h, w]:
                        consistent &= bool((out[:, :, :, h, w] == v).all())
                    out[:, :, :, h, w] = v
                    seen[h, w] = True
    return from_bf16_bits(out), consistent and bool(seen.all())


def pack_w0(w):
    """(64,3,3,7,7) -> uint16 (11,2,5,64,8)."""
    bits = to_bf16_bits(w)
    out = np.zeros((11, 2, 5, 64, 8), np.uint16)
    for p in range(11):
        for k in range(2):
            ch = 2 * p + k
            if ch >= 21:
                continue
            c, kh = ch // 7, l0_chunk_kh(ch % 7)
            for blk in (1, 2, 3):
                out[p, k, blk, :, :7] = bits[:, c, 3 - blk, kh, :]
    return out


def pack_w1(w):
    """(128,64,3,7,7) -> uint16 (3,4,7,7,2,128,8)  [kt][slice][kh][kw][k][row][e]."""
    bits = to_bf16_bits(w).reshape(128, 4, 2, 8, 3, 7, 7)          # row, slice, k, e, kt, kh, kw
    return np.ascontiguousarray(bits.transpose(4, 1, 5, 6, 2, 0, 3))


def pack_w2(w):
    """(128,128,3,7,7) -> uint16 (7,7,2,3,4,2,128,8)  [kh][kw][half][kt][kc][k][row][e]."""
    bits = to_bf16_bits(w).reshape(128, 2, 4, 2, 8, 3, 7, 7)       # row, half, kc, k, e, kt, kh, kw
    return np.ascontiguousarray(bits.transpose(6, 7, 1, 5, 2, 3, 0, 4))


# ------------------------------------------------------------------ launch parameters
class Params:
    def __init__(self, layer, T, HW, B):
        lib = _lib.lib()
        plan = _lib.TcPlan()
        _lib.check(lib.vd_tc_plan_make(ctypes.byref(plan), T, HW, HW), 'tc_plan_make')
        self.plan = plan
        buf = (ctypes.c_int64 * 320)()
        _lib.check(lib.vd_tc_debug_params(layer, ctypes.byref(plan), B, buf, 320), 'tc_debug_params')
        v = list(buf)
        names = ['n_tiles', 'tiles_per_item', 'v_count', 'item_stride', 'u_stride', 'v_stride', 'n_sa', 'n_sb',
                 'sa_stride', 'sb_stride', 'n_copies', 'stage_bytes', 'stage_pitch', 'n_steps', 'a_sa_stride16',
                 'a_lbo16', 'a_sbo16', 'w_resident', 'w_bytes', 'G', 'RW', 'RP', 'n_acc', 'acc_delta16', 'ncols',
                 'acc_cols', 'acc_stages', 'idesc', 'smem_w_off', 'smem_pix_off', 'smem_total', '_']
        for i, n in enumerate(names):
            setattr(self, n, v[i])
        o = 32
        self.copy_gofs = v[o:o + MAX_COPIES]; o += MAX_COPIES
        self.copy_sofs = v[o:o + MAX_COPIES]; o += MAX_COPIES
        self.copy_bytes = v[o:o + MAX_COPIES]; o += MAX_COPIES
        self.b_off16 = v[o:o + MAX_STEPS]; o += MAX_STEPS
        self.b_lbo16 = v[o:o + MAX_STEPS]; o += MAX_STEPS
        self.a_off16 = v[o:o + MAX_STEPS]; o += MAX_STEPS
        self.w_u_stride = v[o]; o += 1
        self.Gt, self.n_wtiles = v[o], v[o + 1]; o += 2          # streamed weights with tile reuse inside a group (0: none)
        self.a_in_group = v[o:o + 16]


def _desc_gather(mem16, start_bytes, lbo_bytes, sbo_bytes, rows):
    """Elements (rows, 16) addressed by a K-major no-swizzle UMMA descriptor over a uint16 image."""
    r = np.arange(rows)
    kk = np.arange(16)
    addr = (start_bytes + (kk[None, :] // 8) * lbo_bytes + (r[:, None] // 8) * sbo_bytes +
            (r[:, None] % 8) * 16 + (kk[None, :] % 8) * 2)
    return mem16[addr // 2]


def emulate_layer(layer, pix_u16, wimg_u16, T, HW, B, tiles=None, fmt='bf16'):
    """Returns D of shape (n_tiles_run, n_acc, 128, ncols) float32 (accumulated in float64).
    fmt: element type of the 16-bit operand images ('bf16' or 'f16')."""
    p = Params(layer, T, HW, B)
    pix = pix_u16.reshape(-1)
    wimg = wimg_u16.reshape(-1)
    tiles = range(p.n_tiles) if tiles is None else tiles
    out = []
    # 16-bit patterns -> float
    def f32(u16):
        if fmt == 'f16':
            return u16.astype(np.uint16).view(np.float16).astype(np.float64)
        return (u16.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
    for tile in tiles:
        item, sub = divmod(tile, p.tiles_per_item)
        u, v = divmod(sub, p.v_count)
        gbase = item * p.item_stride + u * p.u_stride + v * p.v_stride
        D = np.zeros((p.n_acc, 128, p.ncols))
        st = 0
        for sa in range(p.n_sa):
            for sb in range(p.n_sb):
                smem = np.zeros(p.stage_pitch // 2, np.uint16)
                src = gbase + sa * p.sa_stride + sb * p.sb_stride
                for c in range(p.n_copies):
                    n = p.copy_bytes[c] // 2
                    s0 = (src + p.copy_gofs[c]) // 2
                    chunk = pix[s0:s0 + n]
                    assert chunk.size == n, 'bulk copy reads past the end of the packed input'
                    smem[p.copy_sofs[c] // 2: p.copy_sofs[c] // 2 + n] = chunk
                for j in range(p.n_steps):
                    if p.w_resident:
                        a_start = (p.a_off16[j] + sa * p.a_sa_stride16) * 16
                        A = f32(_desc_gather(wimg, a_start, p.a_lbo16 * 16, p.a_sbo16 * 16, 128))
                    else:
                        if p.Gt:      # MMA j of the stage reads tile a_in_group[j % G] of ring slot j // G (Gt tiles per slot)
                            a_start = (st * p.n_wtiles + (j // p.G) * p.Gt + p.a_in_group[j % p.G]) * 4096
                        else:
                            a_start = (st * p.n_steps + j) * 4096
                        A = f32(_desc_gather(wimg, a_start, p.a_lbo16 * 16, p.a_sbo16 * 16, 128))
                    for a in range(p.n_acc):
                        b_start = (p.b_off16[j] + a * p.acc_delta16) * 16
                        Bm = f32(_desc_gather(smem, b_start, p.b_lbo16[j] * 16, 128, p.ncols))
                        D[a] += A @ Bm.T
                st += 1
        out.append(D)
    return np.stack(out).astype(np.float32), p


def emulate_layer0_streaming(pix_u16, wimg_u16, T, HW, B, columns=None):
    """conv 0 in the kernel's input-frame streaming order (WsParams::stream_pairs, tc_conv.cu: next_stream): a tile is a
    column (video, row band); frame i is staged once and feeds pair i//2 - 1 / i//2 (even i: windows 3 and 1) or pair
    (i-1)//2 / (i+1)//2 (odd i: windows 2 and 0) of the resident Toeplitz weight image.  Uses the REAL layer-0 tables.
    Returns {(item, rb): D of shape (T/2, 128, ncols)}."""
    p = Params(0, T, HW, B)
    pix = pix_u16.reshape(-1)
    wimg = wimg_u16.reshape(-1)
    g = Geo(T, HW)
    pairs = T // 2
    n_cols = B * p.v_count
    columns = range(n_cols) if columns is None else columns

    def f32(u16):
        return (u16.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
    out = {}
    for col in columns:
        item, rb = divmod(col, p.v_count)
        gbase = item * p.item_stride + rb * p.v_stride
        D = np.zeros((pairs, 128, p.ncols))
        written = [False] * pairs
        for i in range(T):
            smem = np.zeros(p.stage_pitch // 2, np.uint16)
            src = gbase + (i + 1) * g.frame0                           # frame i of the video is t_pad = i + 1
            for c in range(p.n_copies):
                n = p.copy_bytes[c] // 2
                s0 = (src + p.copy_gofs[c]) // 2
                smem[p.copy_sofs[c] // 2: p.copy_sofs[c] // 2 + n] = pix[s0:s0 + n]
            pp, odd = divmod(i, 2)
            groups = ([(pp - 1, 3)] if pp > 0 else []) + [(pp, 1)] if not odd else [(pp, 2)] + ([(pp + 1, 0)] if pp + 1 < pairs else [])
            for pair, window in groups:
                first = not written[pair]
                acc = np.zeros((128, p.ncols))
                for j in range(p.n_steps):
                    a_start = (p.a_off16[j] + window * p.a_sa_stride16) * 16
                    A = f32(_desc_gather(wimg, a_start, p.a_lbo16 * 16, p.a_sbo16 * 16, 128))
                    Bm = f32(_desc_gather(smem, p.b_off16[j] * 16, p.b_lbo16[j] * 16, 128, p.ncols))
                    acc += A @ Bm.T
                D[pair] = acc if first else D[pair] + acc
                written[pair] = True
        out[(item, rb)] = D.astype(np.float32)
    return out, p


# ------------------------------------------------------------------ split-fp16 forward (SGeo, csrc/tc_layout.h)
def split_f16(x):
    """float32 torch tensor -> (hi, lo) uint16 numpy fp16 bit patterns with x ~= hi + lo (device: split_h)."""
    x = x.float().clamp(-65504.0, 65504.0)
    hi = x.to(torch.float16)
    lo = (x - hi.float()).to(torch.float16)
    return hi.view(torch.int16).numpy().view(np.uint16), lo.view(torch.int16).numpy().view(np.uint16)


def from_f16_bits(u):
    return torch.from_numpy(u.astype(np.uint16).view(np.float16).astype(np.float32))


def f16x2_round(x):
    """the value an fp16 pair carries: hi + lo"""
    hi, lo = split_f16(x)
    return from_f16_bits(hi) + from_f16_bits(lo)


class SGeo:
    def __init__(self, g):
        self.R0s = 4 if g.Wo0 * 4 <= 128 else 2
        self.N0s = self.R0s * g.Wo0
        self.nrb0s = g.Ho0 // self.R0s
        self.frame0s = 12 * g.plane0
        self.video0s = (g.T + 2) * self.frame0s
        self.video1s = 8 * g.slice1
        self.video2s = 196 * g.group2


def l1s_tap(idx):
    if idx < 9:
        ph, pw, r = 0, 0, idx
    elif idx < 21:
        ph, pw, r = 0, 1, idx - 9
    elif idx < 33:
        ph, pw, r = 1, 0, idx - 21
    else:
        ph, pw, r = 1, 1, idx - 33
    nw = 4 if pw else 3
    sh, sw = divmod(r, nw)
    kh = 2 * sh if ph else 2 * sh + 1
    kw = 2 * sw if pw else 2 * sw + 1
    return kh * 7 + kw


def pack_x0s(video, g):
    """video (B,T,3,H,W) float32 -> uint16 (B, T+2, 2, 3, 2, RI0, Wo0, 8)."""
    B = video.shape[0]
    out = np.zeros((B, g.T + 2, 2, 3, 2, g.RI0, g.Wo0, 8), np.uint16)
    for part, bits in enumerate(split_f16(video)):
        for par in range(2):
            for row in range(g.RI0):
                h = 2 * row - 3 if par else 2 * row - 2
                if not (0 <= h < g.HW):
                    continue
                for kw in range(7):
                    wo = np.arange(g.Wo0)
                    w = 2 * wo + kw - 3
                    ok = (w >= 0) & (w < g.HW)
                    o = out[:, 1:g.T + 1, part, :, par, row]                 # (B,T,3,Wo0,8) view
                    o[:, :, :, wo[ok], kw] = bits[:, :, :, h, w[ok]]
    return out


def pack_x0h(video, g):
    """hi-only operand of the two-product mode: uint16 (B, T+2, 3, 2, RI0, Wo0, 8) = part 0 of pack_x0s."""
    return np.ascontiguousarray(pack_x0s(video, g)[:, :, 0])


def pack_a1s(x, g):
    """pooled conv-0 output (B,64,T,H1,H1) float32 -> uint16 (B, 8, T+2, 2, 2, 2, RI1, P1, 8) [chunk][t][part][ph][pw][i][j][e]."""
    B = x.shape[0]
    out = np.zeros((B, 8, g.T + 2, 2, 2, 2, g.RI1, g.P1, 8), np.uint16)
    for part, bits in enumerate(split_f16(x)):
        for h in range(g.H1):
            for w in range(g.H1):
                v = bits[:, :, :, h, w].reshape(B, 8, 8, g.T)            # (B, chunk, e, T)
                out[:, :, 1:g.T + 1, part, coord_par(h), coord_par(w), coord_pos(h), coord_pos(w), :] = v.transpose(0, 1, 3, 2)
    return out


def unpack_a1s(buf, g, B):
    """inverse of pack_a1s -> (hi + lo) float32 (B,64,T,H1,H1)."""
    a = buf.reshape(B, 8, g.T + 2, 2, 2, 2, g.RI1, g.P1, 8)
    out = torch.zeros(B, 64, g.T, g.H1, g.H1)
    for part in range(2):
        bits = np.zeros((B, 64, g.T, g.H1, g.H1), np.uint16)
        for h in range(g.H1):
            for w in range(g.H1):
                v = a[:, :, 1:g.T + 1, part, coord_par(h), coord_par(w), coord_pos(h), coord_pos(w), :]     # (B,8,T,8)
                bits[:, :, :, h, w] = v.transpose(0, 1, 3, 2).reshape(B, 64, g.T)
        out += from_f16_bits(bits)
    return out


def pack_a2s(x, g, Bpad=None):
    """pooled conv-1 output (B,128,T2,H2,H2) float32 -> uint16 (Bpad, 49, 4, 2, 4, To2+2, Ho2, Wo2, 8) [khw][quarter][part][k][t][ho][wo][e]."""
    B = x.shape[0]
    Bpad = Bpad or B
    out = np.zeros((Bpad, 49, 4, 2, 4, g.To2 + 2, g.Ho2, g.Wo2, 8), np.uint16)
    for part, bits in enumerate(split_f16(x)):
        for kh in range(7):
            for kw in range(7):
                for ho in range(g.Ho2):
                    h = 2 * ho + kh - 3
                    if not (0 <= h < g.H2):
                        continue
                    for wo in range(g.Wo2):
                        w = 2 * wo + kw - 3
                        if not (0 <= w < g.H2):
                            continue
                        v = bits[:, :, :, h, w].reshape(B, 4, 4, 8, g.T2)          # (B, quarter, k, e, T2)
                        out[:B, kh * 7 + kw, :, part, :, 1:g.T2 + 1, ho, wo, :] = v.transpose(0, 1, 2, 4, 3)
    return out


def unpack_a2s(buf, g, B):
    a = buf.reshape(-1, 49, 4, 2, 4, g.To2 + 2, g.Ho2, g.Wo2, 8)
    out = torch.zeros(B, 128, g.T2, g.H2, g.H2)
    consistent = True
    for part in range(2):
        bits = np.zeros((B, 128, g.T2, g.H2, g.H2), np.uint16)
        seen = np.zeros((g.H2, g.H2), bool)
        for kh in range(7):
            for kw in range(7):
                for ho in range(g.Ho2):
                    h = 2 * ho + kh - 3
                    if not (0 <= h < g.H2):
                        continue
                    for wo in range(g.Wo2):
                        w = 2 * wo + kw - 3
                        if not (0 <= w < g.H2):
                            continue
                        v = a[:B, kh * 7 + kw, :, part, :, 1:g.T2 + 1, ho, wo, :]      # (B,4,4,T2,8)
                        v = v.transpose(0, 1, 2, 4, 3).reshape(B, 128, g.T2)
                        if seen[h, w]:
                            consistent &= bool((bits[:, :, :, h, w] == v).all())
                        bits[:, :, :, h, w] = v
                        seen[h, w] = True
        consistent &= bool(seen.all())
        out += from_f16_bits(bits)
    return out, consistent


def pack_w0s(w):
    """(64,3,3,7,7) -> uint16 (3, 11, 2, 128, 8) [kt][step][k][row][e]; row 32q + l: channel 16q + (l & 15), part l >> 4."""
    parts = split_f16(w)
    out = np.zeros((3, 11, 2, 128, 8), np.uint16)
    for row in range(128):
        co, part = (row >> 5) * 16 + (row & 15), (row >> 4) & 1
        for step in range(11):
            for k in range(2):
                ch = 2 * step + k
                if ch >= 21:
                    continue
                c, kh = ch // 7, l0_chunk_kh(ch % 7)
                out[:, step, k, row, :7] = parts[part][co, c, :, kh, :]
    return out


def pack_w1s(w):
    """(128,64,3,7,7) -> uint16 (3, 8, 25, 2, 2, 128, 8) [kt][chunk][pair][part hi/lo][k][row][e]; pair 0 = (tap 0, zeros), pair p = taps l1s_tap(2p-1), l1s_tap(2p)."""
    parts = [b.reshape(128, 8, 8, 3, 7, 7) for b in split_f16(w)]        # row, chunk, e, kt, kh, kw
    out = np.zeros((3, 8, 25, 2, 2, 128, 8), np.uint16)
    for pair in range(25):
        for k in range(2):
            idx = 2 * pair - 1 + k if pair else (-1 if k else 0)        # pair 0 = (tap 0, zeros), pair p = taps 2p-1, 2p
            if idx < 0:
                continue
            kh, kw = divmod(l1s_tap(idx), 7)
            for part in range(2):
                out[:, :, pair, part, k, :, :] = parts[part][:, :, :, :, kh, kw].transpose(3, 1, 0, 2)      # (kt, chunk, row, e)
    return out


def pack_w2s(w):
    """(128,128,3,7,7) -> uint16 (7, 7, 4, 3, 2, 2, 2, 128, 8) [kh][kw][quarter][kt][pair][part][k][row][e]; chunk = 2*pair + k."""
    parts = [b.reshape(128, 4, 4, 8, 3, 7, 7) for b in split_f16(w)]     # row, quarter, c, e, kt, kh, kw
    out = np.zeros((7, 7, 4, 3, 2, 2, 2, 128, 8), np.uint16)
    for kt in range(3):
        for pair in range(2):
            for k in range(2):
                for part in range(2):
                    out[:, :, :, kt, pair, part, k, :, :] = parts[part][:, :, 2 * pair + k, :, kt, :, :].transpose(3, 4, 1, 0, 2)   # (kh, kw, quarter, row, e)
    return out


def emulate_layer0s(pix_u16, wimg_u16, T, HW, B, columns=None, layer=6):
    """Split-fp16 conv 0 in the kernel's order (tc_conv.cu: next_l0s) with the REAL tables (debug layer 6; 9 = the two-product
    mode on the hi-only operand X0h): a tile is a column (video, band of R0s rows); stage (frame i, part) feeds output frames
    i + 1 - kt through weight window kt.  Returns {(item, rb): D (T, 128, ncols)} raw accumulators (rows still M-stacked)."""
    p = Params(layer, T, HW, B)
    pix = pix_u16.reshape(-1)
    wimg = wimg_u16.reshape(-1)
    columns = range(p.n_tiles) if columns is None else columns

    def f32(u16):
        return u16.astype(np.uint16).view(np.float16).astype(np.float64)
    out = {}
    for col in columns:
        item, rb = divmod(col, p.v_count)
        gbase = item * p.item_stride + rb * p.v_stride
        D = np.zeros((T, 128, p.ncols))
        for i in range(p.n_sa):
            for part in range(p.n_sb):
                smem = np.zeros(p.stage_pitch // 2, np.uint16)
                src = gbase + i * p.sa_stride + part * p.sb_stride
                for c in range(p.n_copies):
                    n = p.copy_bytes[c] // 2
                    s0 = (src + p.copy_gofs[c]) // 2
                    smem[p.copy_sofs[c] // 2: p.copy_sofs[c] // 2 + n] = pix[s0:s0 + n]
                for kt in (2, 1, 0):
                    f = i + 1 - kt
                    if not (0 <= f < T):
                        continue
                    for j in range(p.n_steps):
                        a_start = (p.a_off16[j] + kt * p.a_sa_stride16) * 16
                        A = f32(_desc_gather(wimg, a_start, p.a_lbo16 * 16, p.a_sbo16 * 16, 128))
                        Bm = f32(_desc_gather(smem, p.b_off16[j] * 16, p.b_lbo16[j] * 16, 128, p.ncols))
                        D[f] += A @ Bm.T
        out[(item, rb)] = D.astype(np.float32)
    return out, p


def unstack_l0s(D):
    """(.., 128, n) M-stacked accumulator rows -> (.., 64, n): channel 16q + j = row 32q + j (wh part) + row 32q + 16 + j (wl part)."""
    d = D.reshape(D.shape[:-2] + (4, 2, 16, D.shape[-1]))
    return (d[..., 0, :, :] + d[..., 1, :, :]).reshape(D.shape[:-2] + (64, D.shape[-1]))
