// Memory-bound kernels of the tensor-core backward (gradient of ConvNet3D.embed w.r.t. its input
// video, frozen weights): transposed weight images, embedding-gradient routing, and the col2im
// gathers that turn the column-GEMM output of conv `l` into the packed dY operand of conv `l-1`
// (applying the ReLU/MaxPool routing code in between).  Layouts: tc_layout.h (BwdGeo).
#include "tc_common.cuh"
#include "tc_layout.h"

namespace vd {
namespace tc {

// wT image [mtile][step][k 2][128][8] bf16: row = ci*147 + tap, column co = step*16 + k*8 + e
__global__ void pack_wt_kernel(const float* __restrict__ w, uint16_t* __restrict__ img, int Cin, int K, int NU) {
    const int n_steps = K / 16;
    const int64_t total = (int64_t)NU * n_steps * 2048;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int e = (int)(i % 8); int64_t q = i / 8;
        int row = (int)(q % 128); q /= 128;
        int k = (int)(q % 2); q /= 2;
        int step = (int)(q % n_steps); int u = (int)(q / n_steps);
        const int r = u * 128 + row;
        float v = 0.f;
        if (r < Cin * 147) {
            const int ci = r / 147, tap = r % 147, co = step * 16 + k * 8 + e;
            v = w[((int64_t)co * Cin + ci) * 147 + tap];
        }
        img[i] = f2bf(v);
    }
}

// g_emb (B, 128*T3p*H3p*H3p) + code2 -> dy2 [video][NT][16 chunks][NC][8]; conv-2 output pixel
// (to,ho,wo) receives the pooled gradient iff it is the recorded argmax of an active window.
__global__ void bwd_emb_kernel(const float* __restrict__ g_emb, const uint8_t* __restrict__ code, uint4* __restrict__ dy,
                               int64_t total, Geo g, BwdGeo b) {
    const int per = g.T3p * g.H3p * g.H3p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int col = (int)(i % b.NC); int64_t q = i / b.NC;
        int chunk = (int)(q % 16); q /= 16;
        int nt = (int)(q % b.NT); int64_t vid = q / b.NT;
        const int pix = nt * b.NC + col;
        const int wo = pix % b.Wo, ho = (pix / b.Wo) % b.Ho, to = pix / (b.Wo * b.Ho);
        const int tq = to >> 1, hp = ho >> 1, wp = wo >> 1;
        const int pos = (to & 1) * 4 + (ho & 1) * 2 + (wo & 1);
        uint16_t v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int co = chunk * 8 + e;
            float val = 0.f;
            if (tq < g.T3p && hp < g.H3p && wp < g.H3p) {
                const int64_t o = vid * g.embed_dim + (int64_t)co * per + (tq * g.H3p + hp) * g.H3p + wp;
                const uint8_t cd = code[o];
                if ((cd & 8) && (cd & 7) == pos) val = g_emb[o];
            }
            v[e] = f2bf(val);
        }
        uint4 o4;
        o4.x = v[0] | ((uint32_t)v[1] << 16); o4.y = v[2] | ((uint32_t)v[3] << 16);
        o4.z = v[4] | ((uint32_t)v[5] << 16); o4.w = v[6] | ((uint32_t)v[7] << 16);
        dy[i] = o4;
    }
}

// dX[ci,t,h,w] of conv `b.layer` = sum over taps of col[(ci,tap)][pixel(t+1-kt, (h+3-kh)/2, (w+3-kw)/2)]
__device__ __forceinline__ float bf2f(uint16_t v) { return __uint_as_float((uint32_t)v << 16); }

// Memory-bound col2im: the column buffer is read exactly once (conv 1/2) or ~1.4x (conv 0 row bands),
// with coalesced 4-byte loads, and summed out of shared memory.
// One block per (video, ci, t, band of `band` input rows).  For each temporal tap kt the 49 rows
// (ci, kt, kh, kw) of the column buffer restricted to the output pixels (to = t+1-kt, ho0..ho1, all wo)
// — one contiguous pixel range — are staged in shared memory [49][npix]; every thread then gathers the
// taps of its input pixels (same parity rule as the forward conv) into register accumulators.
constexpr int kC2iMaxOut = 8;            // input pixels per thread (band * Wi <= 256 * kC2iMaxOut)

struct C2iBlock {
    int vid, ci0, t, hb, h_end, ho0, npix;
};

// block -> (video, group of `cib` input channels, frame t, band of `band` input rows); the channel group is the
// fastest index so that blocks writing neighbouring channels of the same dY chunks run concurrently (L2 merge)
__device__ __forceinline__ C2iBlock c2i_block(const BwdGeo& b, int band, int cib) {
    C2iBlock k;
    const int nb = (b.Hi + band - 1) / band;
    const int ngrp = b.Cin / cib;
    int q = blockIdx.x;
    k.ci0 = (q % ngrp) * cib; q /= ngrp;
    const int ib = q % nb; q /= nb;
    k.t = q % b.Ti; k.vid = q / b.Ti;
    k.hb = ib * band;
    k.h_end = min(b.Hi, k.hb + band);
    k.ho0 = max(0, (k.hb - 2) / 2);                       // smallest ho with 2*ho + kh - 3 >= hb for some kh <= 6
    const int ho1 = min(b.Ho - 1, (k.h_end + 2) / 2);     // largest ho with 2*ho + kh - 3 <= h_end - 1 for some kh >= 0
    k.npix = (ho1 - k.ho0 + 1) * b.Wo;
    return k;
}

// column buffers are bf16 (throughput) or fp32 (accuracy modes of the conv trio)
__device__ __forceinline__ float col_ld(const uint16_t* p) { return bf2f(*p); }
__device__ __forceinline__ float col_ld(const float* p) { return *p; }

// gather the (kh,kw) taps of input pixel (h,w) from the staged rows of one kt
template <typename T>
__device__ __forceinline__ float c2i_taps(const T* __restrict__ sm16, int pitch, const BwdGeo& b, int ho0, int h, int w) {
    float acc = 0.f;
    for (int kh = (h + 1) & 1; kh < 7; kh += 2) {
        const int hh = h + 3 - kh;
        if (hh < 0) continue;
        const int ho = hh >> 1;
        if (ho >= b.Ho) continue;
        const T* row = sm16 + (kh * 7) * pitch + (ho - ho0) * b.Wo;
        for (int kw = (w + 1) & 1; kw < 7; kw += 2) {
            const int ww = w + 3 - kw;
            if (ww < 0) continue;
            const int wo = ww >> 1;
            if (wo >= b.Wo) continue;
            acc += col_ld(row + kw * pitch + wo);
        }
    }
    return acc;
}

// Generic kernel (small / odd-width frames, e.g. the 7x7 input of conv 2): one thread per input pixel of
// `cib` channels; rows staged as bf16 pairs with plain loads.
// layers 2 and 1: every thread owns POOLED elements (ci,t,h,w) of the layer below and writes the whole
// pool window (pt x 2 x 2 conv outputs) of the next dY: the routed gradient at the recorded argmax,
// zeros elsewhere -> dY below is fully overwritten, no memset needed.
// layer 0 (code == nullptr): writes d video (B, T, 3, H, W) fp32.
template <typename T>
__global__ void __launch_bounds__(256) col2im_kernel(const T* __restrict__ colbuf, const uint8_t* __restrict__ code,
                                                     void* __restrict__ out, BwdGeo b, BwdGeo bb, int pt, int band, int pitch, int cib,
                                                     int ncdhw, Dg1Geo dg, int padded_planar) {
    extern __shared__ uint32_t c2i_smem[];
    const C2iBlock k = c2i_block(b, band, cib);
    const T* colv = colbuf + (int64_t)k.vid * b.col_video_elems;
    T* smT = reinterpret_cast<T*>(c2i_smem);
    const int n_out = (k.h_end - k.hb) * b.Wi;
    const int n_all = cib * n_out;
    float acc[kC2iMaxOut];
#pragma unroll
    for (int j = 0; j < kC2iMaxOut; ++j) acc[j] = 0.f;
    for (int kt = 0; kt < 3; ++kt) {
        const int to = k.t + 1 - kt;
        if ((unsigned)to >= (unsigned)b.To) continue;                 // block-uniform
        __syncthreads();
        const int pix0 = (to * b.Ho + k.ho0) * b.Wo;
        // bf16 columns whose staged pixel range is one 16-byte aligned run inside a column tile: 8 elements per load
        const int nt0 = (int)__umulhi((uint32_t)pix0, b.nc_magic), col0 = pix0 - nt0 * b.NC;
        // columns whose staged pixel range is one 16-byte aligned run inside a column tile: VE = 16 / sizeof(T) elements per load
        // (8 bf16 or 4 fp32: the fp32 columns of the split-bf16 backward were staged one element at a time — 0.13 of the copy rate)
        constexpr int VE = 16 / (int)sizeof(T);
        if ((k.npix % VE) == 0 && (col0 % VE) == 0 && (b.NC % VE) == 0 && col0 + k.npix <= b.NC && (pitch % VE) == 0) {
            const int v8 = k.npix / VE;
            for (int i = threadIdx.x; i < cib * 49 * v8; i += blockDim.x) {
                const int row = i / v8, jv = i - row * v8;
                const int cl = row / 49, tap = row - cl * 49;
                const int r = (k.ci0 + cl) * 147 + kt * 49 + tap;
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(colv + (((int64_t)nt0 * b.NU + (r >> 7)) * 128 + (r & 127)) * b.NC + col0) + jv);
                *reinterpret_cast<uint4*>(smT + row * pitch + jv * VE) = v;
            }
        } else
        for (int i = threadIdx.x; i < cib * 49 * k.npix; i += blockDim.x) {
            const int row = i / k.npix, j = i - row * k.npix;        // row = ci_local * 49 + tap
            const int cl = row / 49, tap = row - cl * 49;
            const int r = (k.ci0 + cl) * 147 + kt * 49 + tap;
            const int pix = pix0 + j;
            const int nt = (int)__umulhi((uint32_t)pix, b.nc_magic), col = pix - nt * b.NC;
            smT[row * pitch + j] = __ldg(colv + (((int64_t)nt * b.NU + (r >> 7)) * 128 + (r & 127)) * b.NC + col);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kC2iMaxOut; ++j) {
            const int i = threadIdx.x + j * 256;
            if (i < n_all) {
                // routed bf16 output: channel fastest, so that the 8 channels of a dY chunk are written by neighbouring threads
                const int cl = code ? i % cib : i / n_out, o = code ? i / cib : i - cl * n_out;
                const int hl = o / b.Wi, w = o - hl * b.Wi;
                acc[j] += c2i_taps<T>(smT + cl * 49 * pitch, pitch, b, k.ho0, k.hb + hl, w);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kC2iMaxOut; ++j) {
        const int i = threadIdx.x + j * 256;
        if (i >= n_all) continue;
        const int cl = code ? i % cib : i / n_out, o = code ? i / cib : i - cl * n_out;
        const int hl = o / b.Wi, w = o - hl * b.Wi, h = k.hb + hl, ci = k.ci0 + cl;
        if (code == nullptr) {                            // fp32 gradient: (B,T,Cin,H,W) video layout or plain NCDHW
            float* dv = reinterpret_cast<float*>(out);
            const int64_t plane = ncdhw ? ((int64_t)k.vid * b.Cin + ci) * b.Ti + k.t : ((int64_t)k.vid * b.Ti + k.t) * b.Cin + ci;
            dv[(plane * b.Hi + h) * b.Wi + w] = acc[j];
            continue;
        }
        const uint8_t cd = code[((((int64_t)k.vid * b.Cin + ci) * b.Ti + k.t) * b.Hi + h) * b.Wi + w];
        const uint16_t gv = f2bf((cd & 8) ? acc[j] : 0.f);
        const int arg = cd & 7;
        const int chunk = ci >> 3, e = ci & 7;
        int pos = 0;
        if (padded_planar) {
            // dY of conv 1 in the padded planar layout of the direct dgrad (Dg1Geo): [t_pad][chunk][row ho+1][col wo+1]
            uint16_t* base = reinterpret_cast<uint16_t*>(out) + (int64_t)k.vid * (dg.video_bytes / 2);
            for (int dt = 0; dt < pt; ++dt)
                for (int dh = 0; dh < 2; ++dh)
                    for (int dw = 0; dw < 2; ++dw, ++pos) {
                        const int to = k.t * pt + dt, ho = 2 * h + dh, wo = 2 * w + dw;
                        base[((((int64_t)(to + 1) * 16 + chunk) * dg.RD + ho + 1) * dg.PD + wo + 1) * 8 + e] = (pos == arg) ? gv : (uint16_t)0;
                    }
            continue;
        }
        uint16_t* base = reinterpret_cast<uint16_t*>(out) + (int64_t)k.vid * (bb.dy_video / 2);
        for (int dt = 0; dt < pt; ++dt)
            for (int dh = 0; dh < 2; ++dh)
                for (int dw = 0; dw < 2; ++dw, ++pos) {
                    const int pix = ((k.t * pt + dt) * bb.Ho + (2 * h + dh)) * bb.Wo + 2 * w + dw;
                    const int nt = (int)__umulhi((uint32_t)pix, bb.nc_magic), col = pix - nt * bb.NC;
                    base[(((int64_t)nt * (bb.K / 8) + chunk) * bb.NC + col) * 8 + e] = (pos == arg) ? gv : (uint16_t)0;
                }
    }
}

// ---- fast path (conv 1 and conv 0: even widths, Wi/2 a multiple of SEG) -----------------------------------
// Thread = (input row h of the band, column parity, segment of SEG same-parity columns): SEG register
// accumulators, every staged column value is touched by exactly one LDS + shift + add.  Staging uses
// cp.async (LDGSTS) of V bf16 per request, so many requests per thread are in flight without registers.
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst_smem, const void* src) {
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int SEG, int PAR, typename T>
__device__ __forceinline__ void c2i_row_taps(float (&acc)[SEG], const T* __restrict__ sm16, int pitch, const BwdGeo& b,
                                             int ho0, int h, int seg, bool l_ok, bool r_ok) {
    for (int kh = (h + 1) & 1; kh < 7; kh += 2) {
        const int hh = h + 3 - kh;
        if (hh < 0) continue;
        const int ho = hh >> 1;
        if (ho >= b.Ho) continue;
        const T* row = sm16 + (kh * 7) * pitch + (ho - ho0) * b.Wo + seg * SEG;
#pragma unroll
        for (int kw = (PAR + 1) & 1; kw < 7; kw += 2) {
            const int s = (PAR + 3 - kw) / 2;                                   // PAR+3-kw is even: exact, in {2,1,0,-1}
            const T* src = row + kw * pitch + s;
#pragma unroll
            for (int j = 0; j < SEG; ++j) {
                const bool ok = (j + s < 0) ? l_ok : ((j + s >= SEG) ? r_ok : true);
                if (ok) acc[j] += col_ld(src + j);
            }
        }
    }
}

// Block = `lanes_ch` channel lanes of `tpc` threads; lane c walks channels ci0 + c, ci0 + c + lanes_ch, ... of the
// block's `cib` channels with private staging buffers and its own named barrier, so that the lanes' staging
// latencies and tap sums interleave (the SM always has many warps to issue from).
template <int SEG, int V, typename T>
__global__ void __launch_bounds__(512) col2im_rows_kernel(const T* __restrict__ colbuf, const uint8_t* __restrict__ code,
                                                          void* __restrict__ out, BwdGeo b, BwdGeo bb, int pt, int band, int pitch,
                                                          int warps_per_par, int cib, int lanes_ch, int nbuf, int stash_off, int ncdhw) {
    // shared memory: lanes_ch * nbuf staging buffers [49][pitch] bf16, then (route mode) the stash
    // [cib][n_out] u32 = bf16 << 16 | argmax
    extern __shared__ uint32_t c2i_smem[];
    const C2iBlock k = c2i_block(b, band, cib);
    const T* colv = colbuf + (int64_t)k.vid * b.col_video_elems;
    const T* sm16 = reinterpret_cast<const T*>(c2i_smem);
    uint32_t* stash = c2i_smem + stash_off;
    const uint32_t sm_base = smem_u32(c2i_smem);
    const int stage_elems = 49 * pitch;
    const int nseg = (b.Wi / 2) / SEG;
    const int tpc = 2 * warps_per_par * 32;
    const int chl = (int)threadIdx.x / tpc, tid = (int)threadIdx.x - chl * tpc;
    const int warp = tid >> 5, lane = tid & 31, nwarp = tpc >> 5;
    const int par = warp / warps_per_par;
    const int idx = (warp - par * warps_per_par) * 32 + lane;
    const int hl = idx / nseg, seg = idx - hl * nseg;
    const int h = k.hb + hl;
    const bool active = h < k.h_end;
    const bool l_ok = seg > 0, r_ok = seg < nseg - 1;
    const int n_out = (k.h_end - k.hb) * b.Wi;
    float acc[SEG];
#pragma unroll
    for (int j = 0; j < SEG; ++j) acc[j] = 0.f;
    const int nvec = k.npix / V;
    // valid temporal taps: 0 <= t + 1 - kt < To
    const int kt_lo = max(0, k.t + 2 - b.To), kt_hi = min(2, k.t + 1);
    const int nkt = kt_hi - kt_lo + 1;
    const int n_stage = (cib / lanes_ch) * nkt;
    const int64_t tile_pitch = (int64_t)b.NU * 128 * b.NC;
    const int bar_id = 1 + chl;

    auto lane_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(tpc) : "memory"); };
    auto issue = [&](int st) {
        // one warp per tap row; a V-element vector never straddles an NC tile (pix0 and NC are multiples of V)
        const int ci = k.ci0 + chl + (st / nkt) * lanes_ch, kt = kt_lo + st % nkt;
        const int to = k.t + 1 - kt;
        const int pix0 = (to * b.Ho + k.ho0) * b.Wo;
        constexpr uint32_t ES = (uint32_t)sizeof(T);
        const uint32_t buf = sm_base + (uint32_t)((chl * nbuf + st % nbuf) * stage_elems) * ES;
        for (int tap = warp; tap < 49; tap += nwarp) {
            const int r = ci * 147 + kt * 49 + tap;
            const T* rbase = colv + ((int64_t)(r >> 7) * 128 + (r & 127)) * b.NC;
            const uint32_t dst = buf + (uint32_t)(tap * pitch) * ES;
            for (int j = lane; j < nvec; j += 32) {
                const int pix = pix0 + j * V;
                const int nt = (int)__umulhi((uint32_t)pix, b.nc_magic), col = pix - nt * b.NC;
                cp_async<(int)ES * V>(dst + (uint32_t)j * (ES * V), rbase + nt * tile_pitch + col);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if (nbuf > 1) issue(0);
    for (int st = 0; st < n_stage; ++st) {
        if (nbuf > 1) {
            if (st + 1 < n_stage) { issue(st + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else {
            issue(st);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        lane_sync();
        if (active) {
            const T* buf16 = sm16 + (chl * nbuf + st % nbuf) * stage_elems;
            if (par == 0) c2i_row_taps<SEG, 0, T>(acc, buf16, pitch, b, k.ho0, h, seg, l_ok, r_ok);
            else c2i_row_taps<SEG, 1, T>(acc, buf16, pitch, b, k.ho0, h, seg, l_ok, r_ok);
            if (st % nkt == nkt - 1) {                                // last temporal tap of this channel
                const int cl = chl + (st / nkt) * lanes_ch, ci = k.ci0 + cl;
                if (code == nullptr) {
                    const int64_t plane = ncdhw ? ((int64_t)k.vid * b.Cin + ci) * b.Ti + k.t : ((int64_t)k.vid * b.Ti + k.t) * b.Cin + ci;
                    float* dv = reinterpret_cast<float*>(out) + (plane * b.Hi + h) * b.Wi + 2 * seg * SEG + par;
#pragma unroll
                    for (int j = 0; j < SEG; ++j) { dv[2 * j] = acc[j]; acc[j] = 0.f; }
                } else {
                    const uint8_t* cd_row = code + ((((int64_t)k.vid * b.Cin + ci) * b.Ti + k.t) * b.Hi + h) * b.Wi + 2 * seg * SEG + par;
                    uint32_t* srow = stash + cl * n_out + hl * b.Wi + 2 * seg * SEG + par;
#pragma unroll
                    for (int j = 0; j < SEG; ++j) {
                        const uint32_t cd = cd_row[2 * j];
                        srow[2 * j] = ((uint32_t)f2bf((cd & 8) ? acc[j] : 0.f) << 16) | (cd & 7);
                        acc[j] = 0.f;
                    }
                }
            }
        }
        lane_sync();                                                  // buffer st % nbuf may be refilled now
    }
    if (code == nullptr) return;
    // route mode: the stash holds `cib` (= 8) channels = one 16-byte chunk of dY below.  One thread per output
    // pixel of the pool windows: the routed gradient where the pixel is the recorded argmax, zero elsewhere.
    __syncthreads();
    uint16_t* base = reinterpret_cast<uint16_t*>(out) + (int64_t)k.vid * (bb.dy_video / 2);
    const int rows2 = 2 * (k.h_end - k.hb), cols2 = 2 * b.Wi;
    const int chunk = k.ci0 >> 3;
    for (int pidx = threadIdx.x; pidx < pt * rows2 * cols2; pidx += blockDim.x) {
        const int ww = pidx % cols2; int q = pidx / cols2;
        const int hh = q % rows2, dt = q / rows2;
        const uint32_t pos = (uint32_t)(dt * 4 + (hh & 1) * 2 + (ww & 1));
        const uint32_t* sp = stash + (hh >> 1) * b.Wi + (ww >> 1);
        uint32_t v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t x = (e < cib) ? sp[e * n_out] : 0u;
            v[e] = ((x & 7u) == pos) ? (x >> 16) : 0u;
        }
        const int pix = ((k.t * pt + dt) * bb.Ho + (2 * k.hb + hh)) * bb.Wo + ww;
        const int nt = (int)__umulhi((uint32_t)pix, bb.nc_magic), col = pix - nt * bb.NC;
        *reinterpret_cast<uint4*>(base + (((int64_t)nt * (bb.K / 8) + chunk) * bb.NC + col) * 8) =
            make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
    }
}

static inline unsigned blocks_for(int64_t n) {
    int64_t x = (n + 255) / 256;
    const int64_t cap = 148 * 32;
    return (unsigned)(x < 1 ? 1 : (x > cap ? cap : x));
}

}  // namespace tc
}  // namespace vd

using namespace vd;
using namespace vd::tc;

extern "C" int vd_tc_pack_weights_bwd(const float* w_l0, const float* w_l1, const float* w_l2, void* wt0, void* wt1,
                                      void* wt2, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const float* w[3] = {w_l0, w_l1, w_l2};
    void* o[3] = {wt0, wt1, wt2};
    const int cin[3] = {3, 64, 128}, kk[3] = {64, 128, 128};
    for (int l = 0; l < 3; ++l) {
        if (!w[l] || !o[l]) continue;
        const int NU = (cin[l] * 147 + 127) / 128;
        pack_wt_kernel<<<blocks_for((int64_t)NU * (kk[l] / 16) * 2048), 256, 0, s>>>(w[l], (uint16_t*)o[l], cin[l], kk[l], NU);
        if (int e = check_launch("tc_pack_wt")) return e;
    }
    return 0;
}

extern "C" int vd_tc_bwd_emb(const float* g_emb, const uint8_t* code2, void* dy2, const vd_tc_plan* plan, int B,
                             void* stream) {
    VD_REQUIRE(g_emb && code2 && dy2 && plan, "tc_bwd_emb: NULL pointer");
    VD_REQUIRE(geo_supported(plan->T, plan->H), "tc_bwd_emb: unsupported geometry");
    if (B <= 0) return 0;
    const Geo g = make_geo(plan->T, plan->H);
    const BwdGeo b = make_bwd_geo(g, 2);
    const int64_t total = (int64_t)B * b.NT * 16 * b.NC;
    bwd_emb_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(g_emb, code2, (uint4*)dy2, total, g, b);
    return check_launch("tc_bwd_emb");
}

// mode 0: route with code_below into the packed dY of the layer below (layers 1,2) / d video (B,T,3,H,W) (layer 0)
// mode 1: plain fp32 NCDHW gradient (B, Cin, Ti, Hi, Wi) for any layer (dgrad of the differentiable conv trio)
static int col2im_launch(int layer, const void* col, const uint8_t* code_below, void* out, const vd_tc_plan* plan, int B,
                         void* stream, int ncdhw, int padded_planar = 0, int col_fp32 = 0) {
    VD_REQUIRE(!col_fp32 || ncdhw, "tc_bwd_col2im: fp32 column buffers exist for the plain (NCDHW fp32) output only");
    const int ES = col_fp32 ? 4 : 2;
    VD_REQUIRE(!padded_planar || (layer == 2 && !ncdhw), "tc_bwd_col2im: the padded planar output exists for layer 2 (dY of conv 1) only");
    VD_REQUIRE(col && out && plan, "tc_bwd_col2im: NULL pointer");
    VD_REQUIRE(layer >= 0 && layer <= 2, "tc_bwd_col2im: bad layer");
    VD_REQUIRE(ncdhw || (layer == 0) == (code_below == nullptr), "tc_bwd_col2im: code_below is required for layers 1,2 and must be NULL for layer 0");
    VD_REQUIRE(geo_supported(plan->T, plan->H), "tc_bwd_col2im: unsupported geometry");
    if (B <= 0) return 0;
    const Geo g = make_geo(plan->T, plan->H);
    const BwdGeo b = make_bwd_geo(g, layer);
    cudaStream_t s = (cudaStream_t)stream;
    VD_REQUIRE(b.Wo % 2 == 0 && b.NC % 2 == 0, "tc_bwd_col2im: odd output width");
    // band of input rows per block: the whole frame when it fits kC2iMaxOut pixels per thread, else halved
    int band = b.Hi;
    while ((int64_t)band * b.Wi > 256 * kC2iMaxOut) band = (band + 1) / 2;
    const int n_ho = (band + 2) / 2 + 2;
    const int npix_max = (n_ho < b.Ho ? n_ho : b.Ho) * b.Wo;
    const int pitch = ((npix_max + 7) / 8 * 8) | 8;       // bf16 elements per staged row: 16-byte aligned rows, odd multiple of 16 B
    const size_t stage_bytes = (size_t)49 * pitch * ES;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(col2im_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(col2im_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
#define VD_C2I_ATTR(S_, V_, T_) cudaFuncSetAttribute(col2im_rows_kernel<S_, V_, T_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
        VD_C2I_ATTR(7, 8, uint16_t); VD_C2I_ATTR(7, 4, uint16_t); VD_C2I_ATTR(7, 2, uint16_t);
        VD_C2I_ATTR(8, 8, uint16_t); VD_C2I_ATTR(8, 4, uint16_t); VD_C2I_ATTR(8, 2, uint16_t);
        VD_C2I_ATTR(7, 4, float); VD_C2I_ATTR(7, 2, float); VD_C2I_ATTR(8, 4, float); VD_C2I_ATTR(8, 2, float);
#undef VD_C2I_ATTR
        configured = true;
    }
    const int nb = (b.Hi + band - 1) / band;
    BwdGeo bb = b;
    int pt = 1;
    const bool to_f32 = ncdhw || layer == 0;                         // fp32 output (no routing)
    if (!to_f32) {
        bb = make_bwd_geo(g, layer - 1);
        pt = (layer == 2) ? 2 : 1;             // pool window of the layer below in T: conv1 -> (2,2,2), conv0 -> (1,2,2)
    }
    const uint8_t* cd = to_f32 ? nullptr : code_below;
    // fast path: SEG same-parity columns per thread, V-element vector staging; route mode: 8 channels (one dY
    // chunk) per block on 4 concurrent channel lanes
    const int half = b.Wi / 2;
    const int SEG = (b.Wi % 2 == 0 && half % 7 == 0) ? 7 : ((b.Wi % 2 == 0 && half % 8 == 0) ? 8 : 0);
    int V = col_fp32 ? 4 : 8;                                        // elements per 16-byte cp.async
    const int align_unit = (nb == 1) ? b.Ho * b.Wo : b.Wo;           // pix0 and npix are multiples of this
    while (V > 1 && (align_unit % V != 0 || b.NC % V != 0)) V >>= 1;
    const int per_par = SEG ? (band * (half / SEG) + 31) / 32 : 0;   // warps per column parity
    const int tpc = 2 * per_par * 32;                                // threads per channel lane
    if (SEG != 0 && V >= 2 && tpc <= 512 && (to_f32 || (b.Cin % 8 == 0 && bb.NC % 2 == 0))) {
        const int cib = (b.Cin % 8 == 0) ? 8 : 1;
        int lanes_ch = (b.Cin % 8 == 0) ? 4 : 1;
        while (lanes_ch > 1 && (lanes_ch * tpc > 512 || lanes_ch * stage_bytes > 100 * 1024)) lanes_ch >>= 1;
        const size_t stash_bytes = to_f32 ? 0 : (size_t)cib * band * b.Wi * 4;
        const int nbuf = (lanes_ch == 1 && 2 * stage_bytes + stash_bytes <= 76 * 1024) ? 2 : 1;
        const size_t smem = (size_t)lanes_ch * nbuf * stage_bytes + stash_bytes;
        VD_REQUIRE(smem <= 200 * 1024, "tc_bwd_col2im: shared memory budget exceeded (%zu bytes)", smem);
        const int64_t blocks = (int64_t)B * (b.Cin / cib) * b.Ti * nb;
        VD_REQUIRE(blocks < (1ll << 31), "tc_bwd_col2im: grid too large");
        const int stash_off = (int)(lanes_ch * nbuf * stage_bytes / 4);
        const int threads = lanes_ch * tpc;
#define VD_C2I(S_, V_, T_) col2im_rows_kernel<S_, V_, T_><<<(unsigned)blocks, threads, smem, s>>>((const T_*)col, cd, out, b, bb, pt, band, pitch, per_par, cib, lanes_ch, nbuf, stash_off, ncdhw)
        if (col_fp32) {
            if (SEG == 7) { if (V == 4) VD_C2I(7, 4, float); else VD_C2I(7, 2, float); }
            else { if (V == 4) VD_C2I(8, 4, float); else VD_C2I(8, 2, float); }
        } else if (SEG == 7) { if (V == 8) VD_C2I(7, 8, uint16_t); else if (V == 4) VD_C2I(7, 4, uint16_t); else VD_C2I(7, 2, uint16_t); }
        else { if (V == 8) VD_C2I(8, 8, uint16_t); else if (V == 4) VD_C2I(8, 4, uint16_t); else VD_C2I(8, 2, uint16_t); }
#undef VD_C2I
        return check_launch("tc_bwd_col2im_rows");
    }
    // generic path: as many channels per block as fit 256 * kC2iMaxOut outputs and 96 KiB of staging
    int cib = 8;
    while (cib > 1 && (b.Cin % cib != 0 || (int64_t)cib * band * b.Wi > 256 * kC2iMaxOut || cib * stage_bytes > 96 * 1024)) cib >>= 1;
    const size_t smem = cib * stage_bytes;
    VD_REQUIRE(smem <= 96 * 1024, "tc_bwd_col2im: staging buffer too large (%zu bytes)", smem);
    const int64_t blocks = (int64_t)B * (b.Cin / cib) * b.Ti * nb;
    VD_REQUIRE(blocks < (1ll << 31), "tc_bwd_col2im: grid too large");
    if (col_fp32) col2im_kernel<float><<<(unsigned)blocks, 256, smem, s>>>((const float*)col, cd, out, b, bb, pt, band, pitch, cib, ncdhw,
                                                                            make_dg1_geo(g), padded_planar);
    else col2im_kernel<uint16_t><<<(unsigned)blocks, 256, smem, s>>>((const uint16_t*)col, cd, out, b, bb, pt, band, pitch, cib, ncdhw,
                                                                    make_dg1_geo(g), padded_planar);
    return check_launch("tc_bwd_col2im");
}

extern "C" int vd_tc_bwd_col2im(int layer, const void* col, const uint8_t* code_below, void* out,
                                const vd_tc_plan* plan, int B, void* stream) {
    return col2im_launch(layer, col, code_below, out, plan, B, stream, 0);
}

// layer 2 with out_layout = 1: the routed dY of conv 1 goes to the padded planar layout consumed by vd_tc_dgrad1
// (the buffer's halo cells must have been zeroed once by the caller; data cells are fully overwritten)
extern "C" int vd_tc_bwd_col2im_ex(int layer, const void* col, const uint8_t* code_below, void* out,
                                   const vd_tc_plan* plan, int B, int out_layout, void* stream) {
    return col2im_launch(layer, col, code_below, out, plan, B, stream, 0, out_layout);
}

extern "C" int vd_tc_bwd_col2im_plain(int layer, const void* col, float* gx, const vd_tc_plan* plan, int B, int col_fp32,
                                      void* stream) {
    return col2im_launch(layer, col, nullptr, gx, plan, B, stream, 1, 0, col_fp32);
}
