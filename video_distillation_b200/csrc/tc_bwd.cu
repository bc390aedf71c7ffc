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
    int vid, ci, t, hb, h_end, ho0, npix;
};

__device__ __forceinline__ C2iBlock c2i_block(const BwdGeo& b, int band) {
    C2iBlock k;
    const int nb = (b.Hi + band - 1) / band;
    int q = blockIdx.x;                                   // ci fastest: the 8 channels of a 16-byte dY chunk are written
    k.ci = q % b.Cin; q /= b.Cin;                         // by concurrently running blocks and merge in L2
    const int ib = q % nb; q /= nb;
    k.t = q % b.Ti; k.vid = q / b.Ti;
    k.hb = ib * band;
    k.h_end = min(b.Hi, k.hb + band);
    k.ho0 = max(0, (k.hb - 2) / 2);                       // smallest ho with 2*ho + kh - 3 >= hb for some kh <= 6
    const int ho1 = min(b.Ho - 1, (k.h_end + 2) / 2);     // largest ho with 2*ho + kh - 3 <= h_end - 1 for some kh >= 0
    k.npix = (ho1 - k.ho0 + 1) * b.Wo;
    return k;
}

// stage rows (ci, kt, 0..48) x pixels [pix0, pix0 + npix) of one video's column buffer: smem[tap][npix] (bf16 pairs)
__device__ __forceinline__ void c2i_stage(const uint16_t* __restrict__ colv, const BwdGeo& b, int ci, int kt, int pix0, int npix,
                                          uint32_t* __restrict__ sm, int pitch2) {
    const int np2 = npix >> 1;
    for (int i = threadIdx.x; i < 49 * np2; i += blockDim.x) {
        const int tap = i / np2, j = i - tap * np2;
        const int r = ci * 147 + kt * 49 + tap;
        const int pix = pix0 + 2 * j;
        const int nt = (int)__umulhi((uint32_t)pix, b.nc_magic), col = pix - nt * b.NC;
        sm[tap * pitch2 + j] = __ldg(reinterpret_cast<const uint32_t*>(
            colv + (((int64_t)nt * b.NU + (r >> 7)) * 128 + (r & 127)) * b.NC + col));
    }
}

// gather the (kh,kw) taps of input pixel (h,w) from the staged rows of one kt
__device__ __forceinline__ float c2i_taps(const uint16_t* __restrict__ sm16, int pitch, const BwdGeo& b, int ho0, int h, int w) {
    float acc = 0.f;
    for (int kh = (h + 1) & 1; kh < 7; kh += 2) {
        const int hh = h + 3 - kh;
        if (hh < 0) continue;
        const int ho = hh >> 1;
        if (ho >= b.Ho) continue;
        const uint16_t* row = sm16 + (kh * 7) * pitch + (ho - ho0) * b.Wo;
        for (int kw = (w + 1) & 1; kw < 7; kw += 2) {
            const int ww = w + 3 - kw;
            if (ww < 0) continue;
            const int wo = ww >> 1;
            if (wo >= b.Wo) continue;
            acc += bf2f(row[kw * pitch + wo]);
        }
    }
    return acc;
}

// layers 2 and 1: every thread owns POOLED elements (ci,t,h,w) of the layer below and writes the whole
// pool window (pt x 2 x 2 conv outputs) of the next dY: the routed gradient at the recorded argmax,
// zeros elsewhere -> dY below is fully overwritten, no memset needed.
// layer 0 (code == nullptr): writes d video (B, T, 3, H, W) fp32.
__global__ void __launch_bounds__(256) col2im_kernel(const uint16_t* __restrict__ colbuf, const uint8_t* __restrict__ code,
                                                     void* __restrict__ out, BwdGeo b, BwdGeo bb, int pt, int band, int pitch) {
    extern __shared__ uint32_t c2i_smem[];
    const C2iBlock k = c2i_block(b, band);
    const uint16_t* colv = colbuf + (int64_t)k.vid * b.col_video_elems;
    const int n_out = (k.h_end - k.hb) * b.Wi;
    float acc[kC2iMaxOut];
#pragma unroll
    for (int j = 0; j < kC2iMaxOut; ++j) acc[j] = 0.f;
    for (int kt = 0; kt < 3; ++kt) {
        const int to = k.t + 1 - kt;
        if ((unsigned)to >= (unsigned)b.To) continue;                 // block-uniform
        __syncthreads();
        c2i_stage(colv, b, k.ci, kt, (to * b.Ho + k.ho0) * b.Wo, k.npix, c2i_smem, pitch >> 1);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kC2iMaxOut; ++j) {
            const int i = threadIdx.x + j * 256;
            if (i < n_out) {
                const int hl = i / b.Wi, w = i - hl * b.Wi;
                acc[j] += c2i_taps(reinterpret_cast<const uint16_t*>(c2i_smem), pitch, b, k.ho0, k.hb + hl, w);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kC2iMaxOut; ++j) {
        const int i = threadIdx.x + j * 256;
        if (i >= n_out) continue;
        const int hl = i / b.Wi, w = i - hl * b.Wi, h = k.hb + hl;
        if (code == nullptr) {
            float* dv = reinterpret_cast<float*>(out);
            dv[((((int64_t)k.vid * b.Ti + k.t) * 3 + k.ci) * b.Hi + h) * b.Wi + w] = acc[j];
            continue;
        }
        const uint8_t cd = code[((((int64_t)k.vid * b.Cin + k.ci) * b.Ti + k.t) * b.Hi + h) * b.Wi + w];
        const uint16_t gv = f2bf((cd & 8) ? acc[j] : 0.f);
        const int arg = cd & 7;
        uint16_t* base = reinterpret_cast<uint16_t*>(out) + (int64_t)k.vid * (bb.dy_video / 2);
        const int chunk = k.ci >> 3, e = k.ci & 7;
        int pos = 0;
        for (int dt = 0; dt < pt; ++dt)
            for (int dh = 0; dh < 2; ++dh)
                for (int dw = 0; dw < 2; ++dw, ++pos) {
                    const int pix = ((k.t * pt + dt) * bb.Ho + (2 * h + dh)) * bb.Wo + 2 * w + dw;
                    const int nt = (int)__umulhi((uint32_t)pix, bb.nc_magic), col = pix - nt * bb.NC;
                    base[(((int64_t)nt * (bb.K / 8) + chunk) * bb.NC + col) * 8 + e] = (pos == arg) ? gv : (uint16_t)0;
                }
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

extern "C" int vd_tc_bwd_col2im(int layer, const void* col, const uint8_t* code_below, void* out,
                                const vd_tc_plan* plan, int B, void* stream) {
    VD_REQUIRE(col && out && plan, "tc_bwd_col2im: NULL pointer");
    VD_REQUIRE(layer >= 0 && layer <= 2, "tc_bwd_col2im: bad layer");
    VD_REQUIRE((layer == 0) == (code_below == nullptr), "tc_bwd_col2im: code_below is required for layers 1,2 and must be NULL for layer 0");
    VD_REQUIRE(geo_supported(plan->T, plan->H), "tc_bwd_col2im: unsupported geometry");
    if (B <= 0) return 0;
    const Geo g = make_geo(plan->T, plan->H);
    const BwdGeo b = make_bwd_geo(g, layer);
    cudaStream_t s = (cudaStream_t)stream;
    VD_REQUIRE(b.Wo % 2 == 0 && b.NC % 2 == 0, "tc_bwd_col2im: odd output width");
    // band of input rows per block: the whole frame when it fits kC2iMaxOut pixels per thread, else 16 rows
    int band = b.Hi;
    while ((int64_t)band * b.Wi > 256 * kC2iMaxOut) band = (band + 1) / 2;
    const int n_ho = (band + 2) / 2 + 2;
    const int npix_max = (n_ho < b.Ho ? n_ho : b.Ho) * b.Wo;
    const int pitch = (npix_max + 2) | 2;                 // bf16 elements per staged row (even, odd number of words)
    const size_t smem = (size_t)49 * pitch * 2;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(col2im_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        configured = true;
    }
    VD_REQUIRE(smem <= 96 * 1024, "tc_bwd_col2im: staging buffer too large (%zu bytes)", smem);
    const int nb = (b.Hi + band - 1) / band;
    const int64_t blocks = (int64_t)B * b.Cin * b.Ti * nb;
    VD_REQUIRE(blocks < (1ll << 31), "tc_bwd_col2im: grid too large");
    BwdGeo bb = b;
    int pt = 1;
    if (layer > 0) {
        bb = make_bwd_geo(g, layer - 1);
        pt = (layer == 2) ? 2 : 1;             // pool window of the layer below in T: conv1 -> (2,2,2), conv0 -> (1,2,2)
    }
    col2im_kernel<<<(unsigned)blocks, 256, smem, s>>>((const uint16_t*)col, layer == 0 ? nullptr : code_below, out, b, bb, pt, band, pitch);
    return check_launch("tc_bwd_col2im");
}
