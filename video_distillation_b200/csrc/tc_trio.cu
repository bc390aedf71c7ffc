// Operand packers of the differentiable conv trio on tensor cores (ops.py: fprop / dgrad / wgrad of the
// three ConvNet3D feature convolutions, networks.py:799, as used by the MTT unroll and its double backward).
// Inputs are plain fp32 NCDHW tensors; outputs are the packed bf16 operands of ws_gemm_kernel (tc_layout.h).
//   fprop : x -> X0 (vd_tc_pack_video on the permuted tensor) / A1 / A2, then vd_tc_conv_layer(raw = 2)
//   dgrad : gy -> dY [video][NT][K/8][NC][8], vd_tc_bwd_gemm, vd_tc_bwd_col2im_plain
//   wgrad : x -> im2col columns, gy -> "weight image", vd_tc_wgrad_gemm (split-K), vd_tc_wgrad_reduce
// Memory-bound kernels, one 16-byte chunk per thread.
#include "tc_common.cuh"
#include "tc_layout.h"

namespace vd {
namespace tc {

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint4 o;
    o.x = f2bf(v[0]) | ((uint32_t)f2bf(v[1]) << 16); o.y = f2bf(v[2]) | ((uint32_t)f2bf(v[3]) << 16);
    o.z = f2bf(v[4]) | ((uint32_t)f2bf(v[5]) << 16); o.w = f2bf(v[6]) | ((uint32_t)f2bf(v[7]) << 16);
    return o;
}

// x (B, 64, T, H1, H1) fp32 -> A1 [item][slice 4][t_pad T+2][ph 2][pw 2][k 2][i RI1][j P1] chunks (8 channels)
__global__ void pack_a1_kernel(const float* __restrict__ x, uint4* __restrict__ a1, int64_t total, Geo g, int part) {
    const int64_t S1 = (int64_t)g.T * g.H1 * g.H1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(i % g.P1); int64_t q = i / g.P1;
        int ii = (int)(q % g.RI1); q /= g.RI1;
        int k = (int)(q % 2); q /= 2;
        int pw = (int)(q % 2); q /= 2;
        int ph = (int)(q % 2); q /= 2;
        int tp = (int)(q % (g.T + 2)); q /= (g.T + 2);
        int slice = (int)(q % 4); int64_t item = q / 4;
        const int t = tp - 1, h = ph ? 2 * ii - 3 : 2 * ii - 2, w = pw ? 2 * j - 3 : 2 * j - 2;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (t >= 0 && t < g.T && h >= 0 && h < g.H1 && w >= 0 && w < g.H1) {
            const float* p = x + (item * 64 + slice * 16 + k * 8) * S1 + ((int64_t)t * g.H1 + h) * g.H1 + w;
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float t_ = __ldg(p + e * S1); v[e] = part ? t_ - __uint_as_float((uint32_t)f2bf(t_) << 16) : t_; }
        }
        a1[i] = pack8(v);
    }
}

// x (B, 128, T2, H2, H2) fp32 -> A2 [item][khw 49][half 2][k 8][t_pad To2+2][ho Ho2][wo Wo2] chunks (8 channels)
__global__ void pack_a2_kernel(const float* __restrict__ x, uint4* __restrict__ a2, int64_t total, Geo g, int part) {
    const int64_t S2 = (int64_t)g.T2 * g.H2 * g.H2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int wo = (int)(i % g.Wo2); int64_t q = i / g.Wo2;
        int ho = (int)(q % g.Ho2); q /= g.Ho2;
        int tp = (int)(q % (g.To2 + 2)); q /= (g.To2 + 2);
        int k = (int)(q % 8); q /= 8;
        int half = (int)(q % 2); q /= 2;
        int khw = (int)(q % 49); int64_t item = q / 49;
        const int kh = khw / 7, kw = khw % 7;
        const int t = tp - 1, h = 2 * ho + kh - 3, w = 2 * wo + kw - 3;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (t >= 0 && t < g.T2 && h >= 0 && h < g.H2 && w >= 0 && w < g.H2) {
            const float* p = x + (item * 128 + half * 64 + k * 8) * S2 + ((int64_t)t * g.H2 + h) * g.H2 + w;
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float t_ = __ldg(p + e * S2); v[e] = part ? t_ - __uint_as_float((uint32_t)f2bf(t_) << 16) : t_; }
        }
        a2[i] = pack8(v);
    }
}

// part 0: the value (rounded to bf16 by pack8); part 1: its bf16 residual v - bf16(v) — the split-bf16 dgrad packs both parts of
// the fp32 gradient itself instead of materialising them as tensors
__device__ __forceinline__ float dy_part(float v, int part) {
    return part == 0 ? v : v - __uint_as_float((uint32_t)f2bf(v) << 16);
}

// ---- split-fp16 operands of the one-launch fprop (vd_tc_x3_conv_plain): the A1s / A2s layouts of tc_layout.h from plain NCDHW
// x (B, 64, T, H1, H1) fp32 -> A1s [item][chunk 8][t_pad T+2][part 2][ph 2][pw 2][i RI1][j P1] chunks of 8 fp16 (hi / lo part)
__device__ __forceinline__ uint4 pack8_h(const float (&v)[8], int part) {
    uint16_t h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { uint16_t hi, lo; split_h(v[e], hi, lo); h[e] = part ? lo : hi; }
    uint4 o;
    o.x = h[0] | ((uint32_t)h[1] << 16); o.y = h[2] | ((uint32_t)h[3] << 16);
    o.z = h[4] | ((uint32_t)h[5] << 16); o.w = h[6] | ((uint32_t)h[7] << 16);
    return o;
}

__global__ void pack_a1s_kernel(const float* __restrict__ x, uint4* __restrict__ a1, int64_t total, Geo g) {
    const int64_t S1 = (int64_t)g.T * g.H1 * g.H1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(i % g.P1); int64_t q = i / g.P1;
        int ii = (int)(q % g.RI1); q /= g.RI1;
        int pw = (int)(q % 2); q /= 2;
        int ph = (int)(q % 2); q /= 2;
        int part = (int)(q % 2); q /= 2;
        int tp = (int)(q % (g.T + 2)); q /= (g.T + 2);
        int chunk = (int)(q % 8); int64_t item = q / 8;
        const int t = tp - 1, h = ph ? 2 * ii - 3 : 2 * ii - 2, w = pw ? 2 * j - 3 : 2 * j - 2;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (t >= 0 && t < g.T && h >= 0 && h < g.H1 && w >= 0 && w < g.H1) {
            const float* p = x + (item * 64 + chunk * 8) * S1 + ((int64_t)t * g.H1 + h) * g.H1 + w;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __ldg(p + e * S1);
        }
        a1[i] = pack8_h(v, part);
    }
}

// x (B, 128, T2, H2, H2) fp32 -> A2s [item][khw 49][quarter 4][part 2][k 4][t_pad To2+2][ho Ho2][wo Wo2] chunks of 8 fp16
__global__ void pack_a2s_kernel(const float* __restrict__ x, uint4* __restrict__ a2, int64_t total, Geo g) {
    const int64_t S2 = (int64_t)g.T2 * g.H2 * g.H2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int wo = (int)(i % g.Wo2); int64_t q = i / g.Wo2;
        int ho = (int)(q % g.Ho2); q /= g.Ho2;
        int tp = (int)(q % (g.To2 + 2)); q /= (g.To2 + 2);
        int k = (int)(q % 4); q /= 4;
        int part = (int)(q % 2); q /= 2;
        int quarter = (int)(q % 4); q /= 4;
        int khw = (int)(q % 49); int64_t item = q / 49;
        const int kh = khw / 7, kw = khw % 7;
        const int t = tp - 1, h = 2 * ho + kh - 3, w = 2 * wo + kw - 3;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (t >= 0 && t < g.T2 && h >= 0 && h < g.H2 && w >= 0 && w < g.H2) {
            const float* p = x + (item * 128 + quarter * 32 + k * 8) * S2 + ((int64_t)t * g.H2 + h) * g.H2 + w;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __ldg(p + e * S2);
        }
        a2[i] = pack8_h(v, part);
    }
}

// gy (B, K, To, Ho, Wo) fp32 -> dY [video][nt NT][chunk K/8][col NC][8] bf16
__global__ void pack_dy_kernel(const float* __restrict__ gy, uint4* __restrict__ dy, int64_t total, BwdGeo b, int part) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int col = (int)(i % b.NC); int64_t q = i / b.NC;
        int chunk = (int)(q % (b.K / 8)); q /= (b.K / 8);
        int nt = (int)(q % b.NT); int64_t vid = q / b.NT;
        const float* p = gy + (vid * b.K + chunk * 8) * (int64_t)b.pixels + nt * b.NC + col;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = dy_part(__ldg(p + (int64_t)e * b.pixels), part);
        dy[i] = pack8(v);
    }
}

// ---- wgrad operands.  Pixels k = ((video*To + to)*Ho + ho)*Wo + wo are the GEMM's K dimension, padded with
// zeros to P_pad (a multiple of 128 per split-K slice); columns n = ci*147 + tap (< Cin*147, padded to 256s).
//   xcol : [ntile][stage P_pad/128][chunk 16][col 256][8 pixels]      (N operand, one 64 KiB stage per bulk copy)
//   gyimg: [kstep P_pad/16][k 2][128 rows = cout][8 pixels]            (M operand, 4 KiB UMMA tiles)
__global__ void wgrad_im2col_kernel(const float* __restrict__ x, uint4* __restrict__ xcol, int64_t total, BwdGeo b,
                                    int64_t P, int64_t n_stage, int part) {
    const int64_t Si = (int64_t)b.Ti * b.Hi * b.Wi;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int col = (int)(i % 256); int64_t q = i / 256;
        int c = (int)(q % 16); q /= 16;
        int64_t stage = q % n_stage; int64_t ntile = q / n_stage;
        const int n = (int)(ntile * 256 + col);
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (n < b.Cin * 147) {
            const int ci = n / 147, tap = n - ci * 147;
            const int kt = tap / 49, kh = (tap / 7) % 7, kw = tap % 7;
            int64_t k = stage * 128 + c * 8;
            int wo = (int)(k % b.Wo); int64_t r = k / b.Wo;
            int ho = (int)(r % b.Ho); r /= b.Ho;
            int to = (int)(r % b.To); int64_t vid = r / b.To;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (k + e < P) {
                    const int t = to + kt - 1, h = 2 * ho + kh - 3, w = 2 * wo + kw - 3;
                    if ((unsigned)t < (unsigned)b.Ti && (unsigned)h < (unsigned)b.Hi && (unsigned)w < (unsigned)b.Wi)
                        v[e] = dy_part(__ldg(x + (vid * b.Cin + ci) * Si + ((int64_t)t * b.Hi + h) * b.Wi + w), part);
                }
                if (++wo == b.Wo) { wo = 0; if (++ho == b.Ho) { ho = 0; if (++to == b.To) { to = 0; ++vid; } } }
            }
        }
        xcol[i] = pack8(v);
    }
}

// The same operand, one block per (column tile, stage) = one 64 KiB GEMM stage.  The kernel above gives a thread 8 consecutive
// pixels of ONE column: neighbouring lanes read neighbouring taps of different rows / frames / channels (5-10 sectors per
// warp load, 8 loads per 16-byte store) and it reached 0.84 TB/s — 40 % of an MTT iteration.  Here the lanes of a warp are 32
// CONSECUTIVE PIXELS of one column (stride-2 floats: 2-3 sectors per warp load, taps of the same pixel hit in L1), a thread walks
// its 128 columns incrementally (no divisions in the loop), the tile is assembled in shared memory ([chunk 16][col 256 + 1 pad]
// x 16 B: conflict-free 2-byte stores) and leaves with coalesced 16-byte stores.
// kt_split != 0 ("temporal taps on the gy side", see wgrad_gyimg_tile_kernel): columns n = ci*49 + (kh*7 + kw) of the CENTRE temporal
// tap only, and the GEMM's K index runs over the pixels (video, t, ho, wo) of the INPUT frames t — a third of the columns.
__global__ void __launch_bounds__(256) wgrad_im2col_tile_kernel(const float* __restrict__ x, uint4* __restrict__ xcol, BwdGeo b,
                                                                int64_t P, int64_t n_stage, int kt_split, int part) {
    extern __shared__ uint4 im2col_tile[];                       // [16][257]
    uint16_t* tile16 = reinterpret_cast<uint16_t*>(im2col_tile);
    const int64_t stage = blockIdx.x;
    const int ntile = blockIdx.y;
    const int p = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int64_t Si = (int64_t)b.Ti * b.Hi * b.Wi;
    const int64_t k = stage * 128 + p;
    const bool valid = k < P;
    int wo = 0, ho = 0, to = 0;
    int64_t vid = 0;
    if (valid) {
        wo = (int)(k % b.Wo); int64_t r = k / b.Wo;
        ho = (int)(r % b.Ho); r /= b.Ho;
        to = (int)(r % b.To); vid = r / b.To;
    }
    const float* xv = x + vid * b.Cin * Si;
    const int n0 = ntile * 256 + half * 128;
    const int tpc = kt_split ? 49 : 147;                         // taps per input channel
    int ci = n0 / tpc, tap = n0 - ci * tpc;
    int kt = kt_split ? 1 : tap / 49, kh = (tap / 7) % 7, kw = tap % 7;
    uint16_t* dst = tile16 + (((p >> 3) * 257 + half * 128) * 8 + (p & 7));
    const int ncols = b.Cin * tpc;
    // 16 columns per round: all 16 loads are issued before the first conversion (the loop is bound by load latency, not bandwidth)
    for (int j0 = 0; j0 < 128; j0 += 16) {
        float v[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            v[jj] = 0.f;
            if (valid && n0 + j0 + jj < ncols) {
                const int t = to + kt - 1, h = 2 * ho + kh - 3, w = 2 * wo + kw - 3;
                if ((unsigned)t < (unsigned)b.Ti && (unsigned)h < (unsigned)b.Hi && (unsigned)w < (unsigned)b.Wi)
                    v[jj] = __ldg(xv + ci * Si + ((int64_t)t * b.Hi + h) * b.Wi + w);
            }
            if (++kw == 7) { kw = 0; if (++kh == 7) { kh = 0; if (kt_split) ++ci; else if (++kt == 3) { kt = 0; ++ci; } } }
        }
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) dst[(j0 + jj) * 8] = f2bf(dy_part(v[jj], part));
    }
    __syncthreads();
    uint4* out = xcol + ((int64_t)ntile * n_stage + stage) * (16 * 256);
    for (int idx = threadIdx.x; idx < 16 * 256; idx += 256) out[idx] = im2col_tile[(idx >> 8) * 257 + (idx & 255)];
}

__global__ void wgrad_gyimg_kernel(const float* __restrict__ gy, uint4* __restrict__ img, int64_t total, BwdGeo b, int64_t P, int part) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int row = (int)(i % 128); int64_t q = i / 128;
        int k2 = (int)(q % 2); int64_t kstep = q / 2;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (row < b.K) {
            const int64_t k0 = kstep * 16 + k2 * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int64_t k = k0 + e;
                if (k < P) {
                    const int64_t vid = k / b.pixels, pix = k - vid * b.pixels;
                    v[e] = dy_part(__ldg(gy + (vid * b.K + row) * (int64_t)b.pixels + pix), part);
                }
            }
        }
        img[i] = pack8(v);
    }
}

// gy image of the kt-split wgrad, one block per 128 pixels (= 8 K-steps = 32 KiB of the image), blockIdx.y = kt:
//   img_kt[k = (video, t, ho, wo)][co] = gy[co, video, t - kt + 1, ho, wo]   (zero when that output frame does not exist)
// so that  gw[co, ci, kt, kh, kw] = sum_k img_kt[k][co] * xcol[k][ci*49 + kh*7 + kw]:  the temporal tap becomes a FRAME SHIFT OF
// THE SMALL OPERAND (gy: Cout values per pixel) instead of a third of the columns of the large one (Cin*49 per pixel).
// Lanes = consecutive pixels (coalesced reads of a gy row), tile assembled in shared memory ([chunk 16][row 128 + 1 pad] x 16 B).
__global__ void __launch_bounds__(256) wgrad_gyimg_tile_kernel(const float* __restrict__ gy, uint4* __restrict__ img, BwdGeo b,
                                                               int64_t P, int64_t n_kstep8, int64_t img_kt_u4, int part) {
    extern __shared__ uint4 gy_tile[];                           // [16][129]
    uint16_t* tile16 = reinterpret_cast<uint16_t*>(gy_tile);
    const int64_t blk = blockIdx.x;                              // 128 pixels
    const int kt = blockIdx.y;
    const int p = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int64_t k = blk * 128 + p;
    const int HoWo = b.Ho * b.Wo;
    bool valid = k < P;
    int64_t src = 0;
    if (valid) {
        const int64_t frame = k / HoWo;                          // video * To + t
        const int hw = (int)(k - frame * HoWo);
        const int64_t vid = frame / b.To;
        const int to = (int)(frame - vid * b.To) - kt + 1;
        valid = (unsigned)to < (unsigned)b.To;
        src = (vid * b.K * b.To + to) * (int64_t)HoWo + hw;      // + row * To * HoWo
    }
    const int64_t row_stride = (int64_t)b.To * HoWo;
    uint16_t* dst = tile16 + (((p >> 3) * 129 + half * 64) * 8 + (p & 7));
    for (int j0 = 0; j0 < 64; j0 += 16) {
        float v[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int row = half * 64 + j0 + jj;
            v[jj] = (valid && row < b.K) ? __ldg(gy + src + row * row_stride) : 0.f;
        }
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) dst[(j0 + jj) * 8] = f2bf(dy_part(v[jj], part));
    }
    __syncthreads();
    // image [kstep][k2][128 rows][8 px]: chunk c of the block is K-step c / 2, half c % 2
    uint4* out = img + (int64_t)kt * img_kt_u4 + blk * (16 * 128);
    for (int idx = threadIdx.x; idx < 16 * 128; idx += 256) out[idx] = gy_tile[(idx >> 7) * 129 + (idx & 127)];
    (void)n_kstep8;
}

// gw[co][n] = sum over split-K slices of raw[(split*ntiles + ntile)][co][col], n = ntile*256 + col
// kt_split: n = (ci, kt, khw) of gw comes from raw image kt (raw_kt floats apart), column ci*49 + khw
__global__ void wgrad_reduce_kernel(const float* __restrict__ raw, float* __restrict__ gw, int Cout, int Ncols, int ntiles, int splits,
                                    int kt_split, int64_t raw_kt) {
    const int64_t total = (int64_t)Cout * Ncols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int n = (int)(i % Ncols);
        const int co = (int)(i / Ncols);
        const float* r = raw;
        if (kt_split) {
            const int ci = n / 147, tap = n - ci * 147, kt = tap / 49;
            n = ci * 49 + (tap - kt * 49);
            r += kt * raw_kt;
        }
        const int ntile = n >> 8, col = n & 255;
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += r[(((int64_t)sp * ntiles + ntile) * 128 + co) * 256 + col];
        gw[i] = s;
    }
}

// ---- direct dgrad of conv 1 (Dg1Geo): weight images and the padded planar dY
// image [stage = kt*2 + half][step = (ih*4 + iw)*4 + cs][k2][128 rows m = pw*64 + ci][8]:
//   co = half*64 + cs*16 + k2*8 + e, kh = ph + 3 - 2*s_h(ih), kw = pw + 3 - 2*s_w(iw)  (zero when outside 0..6)
__global__ void pack_dg1_w_kernel(const float* __restrict__ w, uint16_t* __restrict__ img, int ph, int n_sh, int64_t total) {
    const int n_steps = n_sh * 16;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int e = (int)(i % 8); int64_t q = i / 8;
        int m = (int)(q % 128); q /= 128;
        int k2 = (int)(q % 2); q /= 2;
        int step = (int)(q % n_steps); int stage = (int)(q / n_steps);
        const int cs = step % 4, iw = (step / 4) % 4, ih = step / 16;
        const int kt = stage / 2, half = stage % 2;
        const int pw = m >> 6, ci = m & 63;
        const int co = half * 64 + cs * 16 + k2 * 8 + e;
        const int kh = ph + 3 - 2 * dg1_shift(ih, ph ? 2 : 1), kw = pw + 3 - 2 * dg1_shift(iw, 2);
        float v = 0.f;
        if (kh >= 0 && kh < 7 && kw >= 0 && kw < 7) v = w[((((int64_t)co * 64 + ci) * 3 + kt) * 7 + kh) * 7 + kw];
        img[i] = f2bf(v);
    }
}

// gy (B, 128, T, Ho1, Wo1) fp32 -> dYP [video][t_pad T+2][chunk 16][row RD][col PD] chunks; every cell is written
__global__ void pack_dyp1_kernel(const float* __restrict__ gy, uint4* __restrict__ dyp, int64_t total, Dg1Geo d, int T, int part) {
    const int64_t So = (int64_t)T * d.Ho * d.Wo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % d.PD); int64_t q = i / d.PD;
        int r = (int)(q % d.RD); q /= d.RD;
        int chunk = (int)(q % 16); q /= 16;
        int tp = (int)(q % (T + 2)); int64_t vid = q / (T + 2);
        const int t = tp - 1, ho = r - 1, wo = c - 1;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (t >= 0 && t < T && ho >= 0 && ho < d.Ho && wo >= 0 && wo < d.Wo) {
            const float* p = gy + (vid * 128 + chunk * 8) * So + ((int64_t)t * d.Ho + ho) * d.Wo + wo;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = dy_part(__ldg(p + e * So), part);
        }
        dyp[i] = pack8(v);
    }
}

// ---- direct dgrad of conv 0 (Dg0Geo): resident weight image and the padded planar dY
// image [kt 3][step = (ih*4 + iw)*4 + cs][k2][16 rows n = ci*4 + ph*2 + pw][8]:
//   co = cs*16 + k2*8 + e, kh = ph + 3 - 2*s_h(ih), kw = pw + 3 - 2*s_w(iw), s(i) = 2 - i  (zero outside 0..6, rows >= 12)
__global__ void pack_dg0_w_kernel(const float* __restrict__ w, uint16_t* __restrict__ img, int total) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int e = i % 8; int q = i / 8;
        int n = q % 16; q /= 16;
        int k2 = q % 2; q /= 2;
        int step = q % 64; int kt = q / 64;
        const int cs = step % 4, iw = (step / 4) % 4, ih = step / 16;
        const int ci = n >> 2, ph = (n >> 1) & 1, pw = n & 1;
        const int co = cs * 16 + k2 * 8 + e;
        const int kh = ph + 3 - 2 * dg1_shift(ih, 2), kw = pw + 3 - 2 * dg1_shift(iw, 2);
        float v = 0.f;
        if (n < 12 && kh >= 0 && kh < 7 && kw >= 0 && kw < 7) v = w[(((co * 3 + ci) * 3 + kt) * 7 + kh) * 7 + kw];
        img[i] = f2bf(v);
    }
}

// gy (B, 64, T, Ho0, Wo0) fp32 -> dYP0 [video][t_pad T+2][chunk 8][row RD][col PD] chunks; every cell is written
__global__ void pack_dyp0_kernel(const float* __restrict__ gy, uint4* __restrict__ dyp, int64_t total, Dg0Geo d, int T, int part) {
    const int64_t So = (int64_t)T * d.Ho * d.Wo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % d.PD); int64_t q = i / d.PD;
        int r = (int)(q % d.RD); q /= d.RD;
        int chunk = (int)(q % 8); q /= 8;
        int tp = (int)(q % (T + 2)); int64_t vid = q / (T + 2);
        const int t = tp - 1, ho = r - 1, wo = c - 1;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (t >= 0 && t < T && ho >= 0 && ho < d.Ho && wo >= 0 && wo < d.Wo) {
            const float* p = gy + (vid * 64 + chunk * 8) * So + ((int64_t)t * d.Ho + ho) * d.Wo + wo;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = dy_part(__ldg(p + e * So), part);
        }
        dyp[i] = pack8(v);
    }
}

static inline unsigned grid_of(int64_t n) {
    int64_t x = (n + 255) / 256;
    const int64_t cap = 148 * 32;
    return (unsigned)(x < 1 ? 1 : (x > cap ? cap : x));
}

}  // namespace tc
}  // namespace vd

using namespace vd;
using namespace vd::tc;

// x fp32 NCDHW (B, 64, T, H/4, W/4) -> A1 (layer 1) or (B, 128, T/2, H/16.., ..) -> A2 (layer 2); every chunk of the
// packed operand (halo included) is written.
extern "C" int vd_tc_pack_act(int layer, const float* x, void* packed, const vd_tc_plan* plan, int B, int part, void* stream) {
    VD_REQUIRE(x && packed && plan, "tc_pack_act: NULL pointer");
    VD_REQUIRE(layer == 1 || layer == 2, "tc_pack_act: layer must be 1 or 2 (layer 0 uses vd_tc_pack_video_ncdhw)");
    VD_REQUIRE(part == 0 || part == 1, "tc_pack_act: part must be 0 (value) or 1 (bf16 residual)");
    VD_REQUIRE(geo_supported(plan->T, plan->H), "tc_pack_act: unsupported geometry");
    if (B <= 0) return 0;
    const Geo g = make_geo(plan->T, plan->H);
    cudaStream_t s = (cudaStream_t)stream;
    if (layer == 1) {
        const int64_t total = (int64_t)B * (g.video1 / 16);
        pack_a1_kernel<<<grid_of(total), 256, 0, s>>>(x, (uint4*)packed, total, g, part);
    } else {
        const int64_t total = (int64_t)B * (g.video2 / 16);
        pack_a2_kernel<<<grid_of(total), 256, 0, s>>>(x, (uint4*)packed, total, g, part);
    }
    return check_launch("tc_pack_act");
}

// x fp32 NCDHW -> the split-fp16 operand A1s (layer 1) / A2s (layer 2) of vd_tc_x3_conv_plain; every chunk (halo included) written
extern "C" int vd_tc_x3_pack_act(int layer, const float* x, void* packed, const vd_tc_plan* plan, int B, void* stream) {
    VD_REQUIRE(x && packed && plan, "tc_x3_pack_act: NULL pointer");
    VD_REQUIRE(layer == 1 || layer == 2, "tc_x3_pack_act: layer must be 1 or 2");
    VD_REQUIRE(geo_supported(plan->T, plan->H), "tc_x3_pack_act: unsupported geometry");
    if (B <= 0) return 0;
    const Geo g = make_geo(plan->T, plan->H);
    const SGeo sg = make_sgeo(g);
    cudaStream_t s = (cudaStream_t)stream;
    if (layer == 1) {
        const int64_t total = (int64_t)B * (sg.video1s / 16);
        pack_a1s_kernel<<<grid_of(total), 256, 0, s>>>(x, (uint4*)packed, total, g);
    } else {
        const int64_t total = (int64_t)B * (sg.video2s / 16);
        pack_a2s_kernel<<<grid_of(total), 256, 0, s>>>(x, (uint4*)packed, total, g);
    }
    return check_launch("tc_x3_pack_act");
}

// gy fp32 NCDHW (B, Cout, To, Ho, Wo) of conv `layer` -> packed dY operand of vd_tc_bwd_gemm
static int pack_dy_impl(int layer, const float* gy, void* dy, const vd_tc_plan* plan, int B, int part, void* stream);

extern "C" int vd_tc_pack_dy(int layer, const float* gy, void* dy, const vd_tc_plan* plan, int B, void* stream) {
    return pack_dy_impl(layer, gy, dy, plan, B, 0, stream);
}

// part = 0: bf16(gy); part = 1: bf16(gy - bf16(gy)) — the two operand parts of a split-bf16 dgrad, made inside the packer
// (vd_tc_pack_dy_part / vd_tc_pack_dyp1_part / vd_tc_pack_dyp0_part)
extern "C" int vd_tc_pack_dy_part(int layer, const float* gy, void* dy, const vd_tc_plan* plan, int B, int part, void* stream) {
    VD_REQUIRE(part == 0 || part == 1, "tc_pack_dy_part: part must be 0 or 1");
    return pack_dy_impl(layer, gy, dy, plan, B, part, stream);
}

static int pack_dy_impl(int layer, const float* gy, void* dy, const vd_tc_plan* plan, int B, int part, void* stream) {
    VD_REQUIRE(gy && dy && plan, "tc_pack_dy: NULL pointer");
    VD_REQUIRE(layer >= 0 && layer <= 2 && geo_supported(plan->T, plan->H), "tc_pack_dy: bad layer / geometry");
    if (B <= 0) return 0;
    const Geo g = make_geo(plan->T, plan->H);
    const BwdGeo b = make_bwd_geo(g, layer);
    const int64_t total = (int64_t)B * b.NT * (b.K / 8) * b.NC;
    pack_dy_kernel<<<grid_of(total), 256, 0, (cudaStream_t)stream>>>(gy, (uint4*)dy, total, b, part);
    return check_launch("tc_pack_dy");
}

// Sizes of the wgrad workspace for B videos: out[0] = split-K slices, out[1] = stages per slice, out[2] = column
// tiles, out[3] = xcol bytes, out[4] = gyimg bytes, out[5] = raw (fp32 partial sums) bytes
// kt-split mode of the wgrad (default; VD_TC_WGRAD_KT=0 restores the full im2col): the three temporal taps are three GEMMs over
// ONE im2col of the 49 spatial taps, each against a frame-shifted gy image — 2.4x fewer im2col bytes per call (5.8 -> 2.4 GB at
// 50 videos of 16x3x112x112), which was a third of an MTT iteration.
extern "C" int vd_tc_wgrad_kt_mode(int layer) {
    static int mode = -1;
    if (mode < 0) { const char* v = getenv("VD_TC_WGRAD_KT"); mode = (v && *v) ? atoi(v) : 7; }
    return (mode >> layer) & 1;                                  // bit per layer
}

extern "C" int vd_tc_wgrad_plan(int layer, const vd_tc_plan* plan, int B, int64_t* out) {
    VD_REQUIRE(plan && out, "tc_wgrad_plan: NULL pointer");
    VD_REQUIRE(layer >= 0 && layer <= 2 && geo_supported(plan->T, plan->H) && B > 0, "tc_wgrad_plan: bad layer / geometry / batch");
    const Geo g = make_geo(plan->T, plan->H);
    const BwdGeo b = make_bwd_geo(g, layer);
    const int64_t P = (int64_t)B * b.pixels;
    const int kts = vd_tc_wgrad_kt_mode(layer);
    const int64_t ntiles = (b.Cin * (kts ? 49 : 147) + 255) / 256;
    const int64_t stages = (P + 127) / 128;
    int64_t splits = (2 * 148 + ntiles - 1) / ntiles;
    if (splits > stages) splits = stages;
    if (splits < 1) splits = 1;
    const int64_t sps = (stages + splits - 1) / splits;
    out[0] = splits; out[1] = sps; out[2] = ntiles;
    out[3] = ntiles * splits * sps * 65536;
    out[4] = (kts ? 3 : 1) * splits * sps * 8 * 4096;            // kt-split: three gy images / raw buffers, one per temporal tap
    out[5] = (kts ? 3 : 1) * ntiles * splits * 128 * 256 * 4;
    return 0;
}

// x / gy may be NULL: the operand already in xcol / gyimg is kept (the split wgrad packs xh once for gh and gl);
// x_part / gy_part: 0 = the value (rounded to bf16), 1 = its bf16 residual v - bf16(v), made inside the packer
extern "C" int vd_tc_wgrad_pack_parts(int layer, const float* x, int x_part, const float* gy, int gy_part, void* xcol, void* gyimg,
                                      const vd_tc_plan* plan, int B, void* stream) {
    VD_REQUIRE((x || gy) && xcol && gyimg && plan, "tc_wgrad_pack: NULL pointer");
    VD_REQUIRE((x_part == 0 || x_part == 1) && (gy_part == 0 || gy_part == 1), "tc_wgrad_pack: part must be 0 or 1");
    int64_t w[6];
    if (int rc = vd_tc_wgrad_plan(layer, plan, B, w)) return rc;
    const Geo g = make_geo(plan->T, plan->H);
    const BwdGeo b = make_bwd_geo(g, layer);
    const int64_t P = (int64_t)B * b.pixels, n_stage = w[0] * w[1];
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t total_x = w[2] * n_stage * 16 * 256;
    const int kts = vd_tc_wgrad_kt_mode(layer);
    VD_REQUIRE(!kts || (n_stage < (1ll << 31) && w[2] <= 65535), "tc_wgrad_pack: problem too large for the kt-split packer");
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(wgrad_im2col_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 257 * 16);
        cudaFuncSetAttribute(wgrad_gyimg_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 129 * 16);
        configured = true;
    }
    if (x) {
        if (n_stage < (1ll << 31) && w[2] <= 65535) {
            wgrad_im2col_tile_kernel<<<dim3((unsigned)n_stage, (unsigned)w[2], 1), 256, (size_t)16 * 257 * 16, s>>>(x, (uint4*)xcol, b, P, n_stage,
                                                                                                                  kts, x_part);
        } else {
            wgrad_im2col_kernel<<<grid_of(total_x), 256, 0, s>>>(x, (uint4*)xcol, total_x, b, P, n_stage, x_part);
        }
        if (int e = check_launch("tc_wgrad_im2col")) return e;
    }
    if (!gy) return 0;
    if (kts) {
        const int64_t img_kt_u4 = n_stage * 8 * 2 * 128;          // uint4 per image
        wgrad_gyimg_tile_kernel<<<dim3((unsigned)n_stage, 3, 1), 256, 16 * 129 * 16, s>>>(gy, (uint4*)gyimg, b, P, n_stage, img_kt_u4, gy_part);
        return check_launch("tc_wgrad_gyimg");
    }
    const int64_t total_g = n_stage * 8 * 2 * 128;
    wgrad_gyimg_kernel<<<grid_of(total_g), 256, 0, s>>>(gy, (uint4*)gyimg, total_g, b, P, gy_part);
    return check_launch("tc_wgrad_gyimg");
}

extern "C" int vd_tc_wgrad_pack(int layer, const float* x, const float* gy, void* xcol, void* gyimg, const vd_tc_plan* plan,
                                int B, void* stream) {
    VD_REQUIRE(x && gy, "tc_wgrad_pack: NULL pointer");
    return vd_tc_wgrad_pack_parts(layer, x, 0, gy, 0, xcol, gyimg, plan, B, stream);
}

extern "C" int vd_tc_wgrad_reduce(int layer, const float* raw, float* gw, const vd_tc_plan* plan, int B, void* stream) {
    VD_REQUIRE(raw && gw && plan, "tc_wgrad_reduce: NULL pointer");
    int64_t w[6];
    if (int rc = vd_tc_wgrad_plan(layer, plan, B, w)) return rc;
    const Geo g = make_geo(plan->T, plan->H);
    const BwdGeo b = make_bwd_geo(g, layer);
    const int kts = vd_tc_wgrad_kt_mode(layer);
    wgrad_reduce_kernel<<<grid_of((int64_t)b.K * b.Cin * 147), 256, 0, (cudaStream_t)stream>>>(raw, gw, b.K, b.Cin * 147, (int)w[2], (int)w[0],
                                                                                              kts, w[2] * w[0] * 128 * 256);
    return check_launch("tc_wgrad_reduce");
}

// sizes of the direct conv-1 dgrad: out[0] = dYP bytes per video, out[1], out[2] = weight image bytes (ph = 0, 1)
extern "C" int vd_tc_dgrad1_sizes(const vd_tc_plan* plan, int64_t* out) {
    VD_REQUIRE(plan && out && geo_supported(plan->T, plan->H), "tc_dgrad1_sizes: bad plan");
    const Dg1Geo d = make_dg1_geo(make_geo(plan->T, plan->H));
    out[0] = d.video_bytes; out[1] = d.wimg_bytes[0]; out[2] = d.wimg_bytes[1];
    return 0;
}

extern "C" int vd_tc_pack_dgrad1_weights(const float* w_l1, void* wimg0, void* wimg1, const vd_tc_plan* plan, void* stream) {
    VD_REQUIRE(w_l1 && wimg0 && wimg1 && plan && geo_supported(plan->T, plan->H), "tc_pack_dgrad1_weights: bad argument");
    const Dg1Geo d = make_dg1_geo(make_geo(plan->T, plan->H));
    void* img[2] = {wimg0, wimg1};
    for (int ph = 0; ph < 2; ++ph) {
        const int64_t total = d.wimg_bytes[ph] / 2;
        pack_dg1_w_kernel<<<grid_of(total), 256, 0, (cudaStream_t)stream>>>(w_l1, (uint16_t*)img[ph], ph, d.n_sh[ph], total);
        if (int e = check_launch("tc_pack_dg1_w")) return e;
    }
    return 0;
}

// gy fp32 NCDHW (B, 128, T, Ho1, Wo1) -> padded planar dY of conv 1 (every cell written, halo zeros included)
static int pack_dyp1_impl(const float* gy, void* dyp, const vd_tc_plan* plan, int B, int part, void* stream);
extern "C" int vd_tc_pack_dyp1(const float* gy, void* dyp, const vd_tc_plan* plan, int B, void* stream) {
    return pack_dyp1_impl(gy, dyp, plan, B, 0, stream);
}
extern "C" int vd_tc_pack_dyp1_part(const float* gy, void* dyp, const vd_tc_plan* plan, int B, int part, void* stream) {
    VD_REQUIRE(part == 0 || part == 1, "tc_pack_dyp1_part: part must be 0 or 1");
    return pack_dyp1_impl(gy, dyp, plan, B, part, stream);
}
static int pack_dyp1_impl(const float* gy, void* dyp, const vd_tc_plan* plan, int B, int part, void* stream) {
    VD_REQUIRE(gy && dyp && plan && geo_supported(plan->T, plan->H), "tc_pack_dyp1: bad argument");
    if (B <= 0) return 0;
    const Geo g = make_geo(plan->T, plan->H);
    const Dg1Geo d = make_dg1_geo(g);
    const int64_t total = (int64_t)B * (d.video_bytes / 16);
    pack_dyp1_kernel<<<grid_of(total), 256, 0, (cudaStream_t)stream>>>(gy, (uint4*)dyp, total, d, g.T, part);
    return check_launch("tc_pack_dyp1");
}

// sizes of the direct conv-0 dgrad: out[0] = dYP0 bytes per video, out[1] = weight image bytes
extern "C" int vd_tc_dgrad0_sizes(const vd_tc_plan* plan, int64_t* out) {
    VD_REQUIRE(plan && out && geo_supported(plan->T, plan->H), "tc_dgrad0_sizes: bad plan");
    const Dg0Geo d = make_dg0_geo(make_geo(plan->T, plan->H));
    out[0] = d.video_bytes; out[1] = d.wimg_bytes;
    return 0;
}

extern "C" int vd_tc_pack_dgrad0_weights(const float* w_l0, void* wimg, void* stream) {
    VD_REQUIRE(w_l0 && wimg, "tc_pack_dgrad0_weights: NULL pointer");
    const int total = 3 * 64 * 512 / 2;
    pack_dg0_w_kernel<<<grid_of(total), 256, 0, (cudaStream_t)stream>>>(w_l0, (uint16_t*)wimg, total);
    return check_launch("tc_pack_dg0_w");
}

// gy fp32 NCDHW (B, 64, T, Ho0, Wo0) -> padded planar dY of conv 0 (every cell written, halo zeros included)
static int pack_dyp0_impl(const float* gy, void* dyp, const vd_tc_plan* plan, int B, int part, void* stream);
extern "C" int vd_tc_pack_dyp0(const float* gy, void* dyp, const vd_tc_plan* plan, int B, void* stream) {
    return pack_dyp0_impl(gy, dyp, plan, B, 0, stream);
}
extern "C" int vd_tc_pack_dyp0_part(const float* gy, void* dyp, const vd_tc_plan* plan, int B, int part, void* stream) {
    VD_REQUIRE(part == 0 || part == 1, "tc_pack_dyp0_part: part must be 0 or 1");
    return pack_dyp0_impl(gy, dyp, plan, B, part, stream);
}
static int pack_dyp0_impl(const float* gy, void* dyp, const vd_tc_plan* plan, int B, int part, void* stream) {
    VD_REQUIRE(gy && dyp && plan && geo_supported(plan->T, plan->H), "tc_pack_dyp0: bad argument");
    if (B <= 0) return 0;
    const Geo g = make_geo(plan->T, plan->H);
    const Dg0Geo d = make_dg0_geo(g);
    const int64_t total = (int64_t)B * (d.video_bytes / 16);
    pack_dyp0_kernel<<<grid_of(total), 256, 0, (cudaStream_t)stream>>>(gy, (uint4*)dyp, total, d, g.T, part);
    return check_launch("tc_pack_dyp0");
}
