// Packed bf16 activation / weight layouts of the tensor-core path (host + device).
//
// ConvNet3D feature convs are k=(3,7,7) s=(1,2,2) p=(1,3,3) (networks.py:799).  Each layer is a
// GEMM  D[cout, q] = sum_j A_j[cout, 0:16] . B_j[q, 0:16]^T  whose B operand (pixels) is read
// straight out of the packed input through UMMA descriptors with a per-step start offset
// ("shifted window"), so the input of a tile is staged in shared memory once and reused by
// every filter tap.  All chunks are 16 bytes = 8 bf16.
//
// X0 (input of conv 0, Cin=3): "kw-expanded"
//   [t_pad T+2][c 3][par 2][row RI0][wo Wo0] chunk = x[t_pad-1][c][h][2wo-3 .. 2wo+4],
//   h = 2*row-2 (par 0, even h) or 2*row-3 (par 1, odd h); zero outside the video; the 8th
//   element (kw=7) is stored as 0.  Output pixel (ho,wo) of tap kh reads plane par=(kh+1)&1 at
//   row ho + (par ? kh/2 : (kh-1)/2): a pure row shift, so q = ho*Wo0+wo is a linear window.
//
// A1 (input of conv 1, Cin=64): "parity-planar"
//   [slice 4 (16 ch)][t_pad T+2][ph 2][pw 2][k 2 (8 ch)][i RI1][j P1] chunk = 8 channels of
//   pixel (h,w), h = 2i-2 / 2i-3 for ph = 0/1, w = 2j-2 / 2j-3 for pw = 0/1.  P1 = Wo1+2: the
//   one out-of-row read of the odd plane (j = P1) lands on the next row's j=0, which is also a
//   zero halo cell.  Window of tap (kh,kw): plane (ph,pw) = ((kh+1)&1,(kw+1)&1), start
//   (rowshift*P1 + colshift) chunks; q = ho*P1 + wo (columns wo >= Wo1 are discarded).
//
// A2 (input of conv 2, Cin=128): "tap-expanded"
//   [khw 49][half 2][k 8 (8 ch)][t_pad To2+2][ho Ho2][wo Wo2] chunk = 8 channels of input
//   pixel (t_pad-1, 2ho+kh-3, 2wo+kw-3) (zero outside).  q = to*Ho2*Wo2 + ho*Wo2 + wo; tap kt is a
//   shift of kt frames.
#pragma once
#include <stdint.h>

namespace vd {
namespace tc {

struct Geo {
    int T, HW;
    // conv 0
    int Ho0, Wo0, R0, RI0, N0;          // R0 output rows per tile, RI0 rows per parity plane
    int64_t plane0, frame0, video0;     // bytes
    int stage0;                         // bytes of one pixel stage in smem
    // after pool (1,2,2)
    int H1;                             // = W1
    // conv 1
    int Ho1, Wo1, P1, RI1, N1;
    int64_t plane1, frame1, slice1, video1;
    // after pool (2,2,2)
    int T2, H2;
    // conv 2
    int To2, Ho2, Wo2, HW2, N2;
    int64_t chunk2, group2, video2;
    int T3p, H3p, embed_dim;
};

inline Geo make_geo(int T, int HW) {
    Geo g{};
    g.T = T; g.HW = HW;
    g.Ho0 = HW / 2; g.Wo0 = HW / 2;
    g.R0 = (g.Wo0 * 8 <= 256) ? 8 : 4;
    g.RI0 = g.Ho0 + 3;
    g.N0 = g.R0 * g.Wo0;
    g.plane0 = (int64_t)g.RI0 * g.Wo0 * 16;
    g.frame0 = 6 * g.plane0;
    g.video0 = (int64_t)(T + 2) * g.frame0;
    g.stage0 = 3 * (2 * g.R0 + 5) * g.Wo0 * 16;
    g.H1 = g.Ho0 / 2;
    g.Ho1 = g.H1 / 2; g.Wo1 = g.H1 / 2;
    g.P1 = g.Wo1 + 2; g.RI1 = g.Ho1 + 4;
    g.N1 = g.Ho1 * g.P1;
    g.plane1 = (int64_t)g.RI1 * g.P1 * 16;
    g.frame1 = 8 * g.plane1;
    g.slice1 = (int64_t)(T + 2) * g.frame1;
    g.video1 = 4 * g.slice1;
    g.T2 = T / 2; g.H2 = g.Ho1 / 2;
    g.To2 = g.T2; g.Ho2 = (g.H2 - 1) / 2 + 1; g.Wo2 = g.Ho2;
    g.HW2 = g.Ho2 * g.Wo2;
    g.N2 = g.To2 * g.HW2;
    g.chunk2 = (int64_t)(g.To2 + 2) * g.HW2 * 16;
    g.group2 = 8 * g.chunk2;
    g.video2 = 98 * g.group2;
    g.T3p = g.To2 / 2; g.H3p = g.Ho2 / 2;
    g.embed_dim = 128 * g.T3p * g.H3p * g.H3p;
    return g;
}

inline bool geo_supported(int T, int HW) {
    if (!(HW == 112 || HW == 64)) return false;
    if (T < 4 || (T % 4) != 0 || T > 32) return false;
    Geo g = make_geo(T, HW);
    return g.N0 % 16 == 0 && g.N0 <= 256 && g.N1 % 16 == 0 && g.N1 <= 256 && g.N2 % 16 == 0 && g.N2 <= 128 &&
           g.Ho0 % g.R0 == 0;
}

// row / column shift of filter tap k (0..6) inside its parity plane, and the plane it reads
__host__ __device__ inline int tap_par(int k) { return (k + 1) & 1; }           // 0: even coord, 1: odd coord
__host__ __device__ inline int tap_shift(int k) { return (k & 1) ? (k - 1) / 2 : k / 2; }
// plane index of an input coordinate and its position in the plane
__host__ __device__ inline int coord_par(int x) { return x & 1; }
__host__ __device__ inline int coord_pos(int x) { return (x & 1) ? (x + 3) / 2 : x / 2 + 1; }

// conv 0 chunk order inside one channel: sorted by shared-memory offset (even-h plane first)
// so that chunk pairs (2s, 2s+1) always have a positive descriptor LBO.  chunk = c*7 + idx.
__host__ __device__ inline int l0_chunk_kh(int idx) { return idx < 3 ? 2 * idx + 1 : 2 * (idx - 3); }

// ---- backward column GEMM of conv `layer`: col[(ci,tap), pixel] = sum_co W[co,ci,tap] dY[co,pixel]
//   dY packed  : [video][ntile NT][chunk K/8][col NC][8]      (bf16, K = Cout)
//   wT image   : [mtile NU][step K/16][k 2][128 rows][8]      row = ci*147 + tap (zero beyond Cin*147)
//   col        : [video][ntile NT][mtile NU][128 rows][NC]    (fp32)
struct BwdGeo {
    int layer, Cin, K, Mrows, NU, pixels, NC, NT, n_steps;
    int To, Ho, Wo;            // output (dY) extent of this conv
    int Ti, Hi, Wi;            // input (dX) extent of this conv
    int64_t dy_video, col_video_elems, wt_bytes;
    uint32_t nc_magic;         // ceil(2^32 / NC): pix / NC == __umulhi(pix, nc_magic) for pix < 2^16
};

inline BwdGeo make_bwd_geo(const Geo& g, int layer) {
    BwdGeo b{};
    b.layer = layer;
    if (layer == 0) { b.Cin = 3; b.K = 64; b.To = g.T; b.Ho = g.Ho0; b.Wo = g.Wo0; b.Ti = g.T; b.Hi = g.HW; b.Wi = g.HW; }
    else if (layer == 1) { b.Cin = 64; b.K = 128; b.To = g.T; b.Ho = g.Ho1; b.Wo = g.Wo1; b.Ti = g.T; b.Hi = g.H1; b.Wi = g.H1; }
    else { b.Cin = 128; b.K = 128; b.To = g.To2; b.Ho = g.Ho2; b.Wo = g.Wo2; b.Ti = g.T2; b.Hi = g.H2; b.Wi = g.H2; }
    b.Mrows = b.Cin * 147;
    b.NU = (b.Mrows + 127) / 128;
    b.pixels = b.To * b.Ho * b.Wo;
    b.NC = 16;
    for (int c = 256; c >= 16; c -= 16) if (b.pixels % c == 0) { b.NC = c; break; }
    b.NT = b.pixels / b.NC;
    b.nc_magic = (uint32_t)(((1ull << 32) + b.NC - 1) / b.NC);
    b.n_steps = b.K / 16;
    b.dy_video = (int64_t)b.NT * (b.K / 8) * b.NC * 16;
    b.col_video_elems = (int64_t)b.NT * b.NU * 128 * b.NC;
    b.wt_bytes = (int64_t)b.NU * b.n_steps * 4096;
    return b;
}

// ---- direct (column-free) dgrad of conv 1 as a shifted-window GEMM, one launch per input-row parity ph:
//   dX[ci, t, 2a+ph, 2b+pw] = sum_{kt} sum_{s_h, s_w} sum_co W[co, ci, kt, ph+3-2 s_h, pw+3-2 s_w] * dY[co, t+1-kt, a+s_h, b+s_w]
//   M rows   : m = pw*64 + ci                      (128)
//   N columns: q = a*PD + b, PD = Wo1 + 2          (= N1 of the forward conv; columns b >= Wo1 are discarded)
//   K        : (kt 3) x (co half 2) = 6 stages of [s_h slots][s_w 4][co step 4] K=16 steps; s in {2,1,0,-1}
//   dYP ("padded planar" dY of conv 1): [video][t_pad T+2][chunk 16][row RD = Ho1+4][col PD] x 16 B,
//       row r <-> ho = r-1, col c <-> wo = c-1; every cell outside the image is zero, so a window shift is a plain
//       address offset ((s_h+1)*PD + (s_w+1)) and out-of-image taps read zeros.
struct Dg1Geo {
    int Ho, Wo, PD, RD, N;
    int plane16;                   // RD * PD: 16-byte units per chunk plane
    int64_t frame_bytes, video_bytes;
    int n_sh[2];                   // valid row shifts for ph = 0 (3: kh 1,3,5) and ph = 1 (4: kh 0,2,4,6)
    int n_steps[2];                // n_sh * 4 * 4
    int64_t wimg_bytes[2];         // 6 stages * n_steps * 4 KiB
};

inline Dg1Geo make_dg1_geo(const Geo& g) {
    Dg1Geo d{};
    d.Ho = g.Ho1; d.Wo = g.Wo1; d.PD = g.P1; d.RD = g.Ho1 + 4; d.N = g.N1;
    d.plane16 = d.RD * d.PD;
    d.frame_bytes = (int64_t)16 * d.plane16 * 16;
    d.video_bytes = (int64_t)(g.T + 2) * d.frame_bytes;
    for (int ph = 0; ph < 2; ++ph) {
        d.n_sh[ph] = ph ? 4 : 3;
        d.n_steps[ph] = d.n_sh[ph] * 16;
        d.wimg_bytes[ph] = (int64_t)6 * d.n_steps[ph] * 4096;
    }
    return d;
}
// shift value of slot i (0..3): 2, 1, 0, -1; ph = 0 has no s_h = 2 slot (kh would be -1): its slots are 1, 0, -1
__host__ __device__ inline int dg1_shift(int slot, int first) { return first - slot; }

// ---- direct (column-free) dgrad of conv 0 (Cin = 3): pixels on M, the 12 (ci, ph, pw) outputs on N = 16.
//   dX[ci, t, 2a+ph, 2b+pw] = sum_{kt, s_h, s_w, co} dY[co, t+1-kt, a+s_h, b+s_w] * W[co, ci, kt, ph+3-2 s_h, pw+3-2 s_w]
//   M rows   : m = (a - a0)*PD + b, RT = 128 / PD rows of PD = 64 columns per tile (columns b >= Wo0 are discarded)
//   N columns: n = ci*4 + ph*2 + pw (12 of 16)
//   K        : kt (3 stages) x [s_h 4][s_w 4][co step 4] K=16 steps; the 96 KiB weight image is resident in smem
//   dYP0 ("padded planar" dY of conv 0): [video][t_pad T+2][chunk 8][row RD = Ho0+4][col PD] x 16 B, row r <-> ho = r-1,
//       col c <-> wo = c-1, zeros outside the image: the A operand of step (s_h, s_w) is the stage shifted by
//       (s_h+1)*PD + (s_w+1) chunks.
struct Dg0Geo {
    int Ho, Wo, PD, RT, RD, RS;    // RS = RT + 4 staged rows per chunk
    int plane16;                   // RD * PD
    int stage_plane16;             // RS * PD
    int64_t frame_bytes, video_bytes;
    int64_t wimg_bytes;            // 3 * 64 * 512
};

inline Dg0Geo make_dg0_geo(const Geo& g) {
    Dg0Geo d{};
    d.Ho = g.Ho0; d.Wo = g.Wo0; d.PD = 64; d.RT = 2; d.RD = g.Ho0 + 4; d.RS = d.RT + 4;
    d.plane16 = d.RD * d.PD;
    d.stage_plane16 = d.RS * d.PD;
    d.frame_bytes = (int64_t)8 * d.plane16 * 16;
    d.video_bytes = (int64_t)(g.T + 2) * d.frame_bytes;
    d.wimg_bytes = 3 * 64 * 512;
    return d;
}

// ---- split-fp16 forward ("x3"): every operand value v is carried as an fp16 pair v = hi + lo (hi = fp16(v),
//   lo = fp16(v - hi): 22 mantissa bits) and every product x*w is evaluated as xh*wh + xl*wh + xh*wl (+ xl*wl in conv 0)
//   into the SAME fp32 TMEM accumulator, so ReLU / pool-argmax are decided on fp32-equivalent sums.
//   The hi / lo parts are extra K-chunks of the same shifted-window GEMM — only layouts and step tables change:
//
// X0s (input of conv 0): [t_pad T+2][part 2][c 3][par 2][row RI0][wo Wo0] x 16 B (8 fp16 = the kw window, as X0).
//   conv 0 is "M-stacked": weight-tile row r = 32*q + l holds channel co = 16*q + (l & 15), hi part for l < 16 and lo
//   part for l >= 16, so ONE 4 KiB tile per (kt, K-step) carries wh and wl and the accumulator rows of a channel are
//   top = sum wh*x, bottom = sum wl*x; the epilogue adds the two rows (same warp, lanes l and l ^ 16).  A tile is a
//   column (video, band of R0s output rows, N = R0s*Wo0 <= 128); its T input frames are streamed once each as two
//   stages (hi part, lo part) and frame i feeds output frames i-1, i, i+1 (kt = 2, 1, 0) held in 4 x 128 TMEM columns.
//
// A1s (input of conv 1): [chunk 8 (8 ch)][t_pad T+2][part 2][ph 2][pw 2][i RI1][j P1] x 16 B — the A1 layout with the
//   16-channel slice replaced by an 8-channel chunk carrying both parts (same bytes per stage).  Per stage (kt, chunk) the 49
//   taps (smem-offset order, l1s_tap) are paired (a, b) = (2p-1, 2p), pair 0 = (tap 0, zero weights), into the two K halves of a
//   K = 16 MMA, three MMAs a pair:
//     B = [xh_a | xh_b] (LBO = window distance)  x  A = [wh_a | wh_b]
//     B = [xl_a | xl_b] (the same + 4 planes)    x  A = [wh_a | wh_b]   (the SAME weight tile: streamed once, used twice)
//     B = [xh_a | xh_b]                          x  A = [wl_a | wl_b]
//   -> 75 MMAs and 50 streamed weight tiles per stage.
//
// A2s (input of conv 2): [khw 49][quarter 4][part 2][k 4 (8 ch)][t_pad To2+2][ho][wo] x 16 B.  Per stage (khw, quarter) and
//   kt = 0..2 the chunk pairs (0,1), (2,3) get the same three MMAs: 18 MMAs and 12 streamed weight tiles per stage.
struct SGeo {
    int R0s, N0s, nrb0s;                // conv 0: output rows per column band, accumulator columns, bands per frame
    int stage0s;                        // bytes of one (frame, part) stage
    int64_t frame0s, video0s;           // X0s bytes
    int64_t video1s, video2s;           // A1s / A2s bytes per video
    int64_t w0s_bytes, w1s_bytes, w2s_bytes;
};
constexpr int kSteps1s = 75;                // 25 tap pairs x 3 MMAs (xh.wh, xl.wh, xh.wl)
constexpr int kSteps2s = 18;                // 3 kt x 2 chunk pairs x 3 MMAs
constexpr int kWTiles1s = 50, kWTiles2s = 12;   // streamed weight tiles per stage: [hi, lo] per pair

inline SGeo make_sgeo(const Geo& g) {
    SGeo s{};
    s.R0s = (g.Wo0 * 4 <= 128) ? 4 : 2;
    s.N0s = s.R0s * g.Wo0;
    s.nrb0s = g.Ho0 / s.R0s;
    s.stage0s = 3 * (2 * s.R0s + 5) * g.Wo0 * 16;
    s.frame0s = 12 * g.plane0;
    s.video0s = (int64_t)(g.T + 2) * s.frame0s;
    s.video1s = 8 * g.slice1;
    s.video2s = 196 * g.group2;
    s.w0s_bytes = (int64_t)3 * 11 * 4096;
    s.w1s_bytes = (int64_t)3 * 8 * kWTiles1s * 4096;
    s.w2s_bytes = (int64_t)49 * 4 * kWTiles2s * 4096;
    return s;
}

// conv 1 taps in the order of their window offsets inside a stage: planes (ph,pw) = (0,0) (0,1) (1,0) (1,1), then row
// shift, then column shift.  Returns kh*7 + kw.
__host__ __device__ inline int l1s_tap(int idx) {
    int ph, pw, r;
    if (idx < 9) { ph = 0; pw = 0; r = idx; }
    else if (idx < 21) { ph = 0; pw = 1; r = idx - 9; }
    else if (idx < 33) { ph = 1; pw = 0; r = idx - 21; }
    else { ph = 1; pw = 1; r = idx - 33; }
    const int nw = pw ? 4 : 3;
    const int sh = r / nw, sw = r - sh * nw;
    const int kh = ph ? 2 * sh : 2 * sh + 1, kw = pw ? 2 * sw : 2 * sw + 1;
    return kh * 7 + kw;
}

constexpr int kWeightTileBytes = 4096;      // [k 2][128 rows][16 B]
constexpr int kVideosPerTile2 = 4;          // conv 2: accumulators (videos) per CTA tile
constexpr int kW0Steps = 11;                // conv 0: 21 (c,kh) chunks paired into K=16 steps
constexpr int kW0Bytes = kW0Steps * 2 * 5 * 1024;

}  // namespace tc
}  // namespace vd
