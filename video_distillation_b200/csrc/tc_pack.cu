// Operand packers of the tensor-core path: fp32 videos -> X0 (kw-expanded bf16), fp32 OIDHW
// weights -> UMMA weight images.  Layouts: tc_layout.h.  Memory-bound, one 16-byte chunk (or
// one bf16 element for weights) per thread.
#include "tc_common.cuh"
#include "tc_layout.h"

namespace vd {
namespace tc {

// video (Bsrc, T, 3, H, W) fp32 -> x0 (B, T+2, 3, 2, RI0, Wo0) chunks; source video of item b is
// index ? index[b] : b (the device-resident get_images gather, distill_s2d_ms.py:81-87).
// part: 0 = the value (rounded to bf16), 1 = its bf16 residual v - bf16(v) (split-bf16 operands of the conv trio)
__device__ __forceinline__ float bf16_part(float v, int part) {
    return part == 0 ? v : v - __uint_as_float((uint32_t)f2bf(v) << 16);
}

// one block per (video, padded frame, channel, row parity) plane of RI0 x Wo0 chunks: the plane indices come from the block
// index (no 64-bit divisions per chunk), the threads walk the plane with coalesced 16-byte stores
__global__ void __launch_bounds__(256) pack_video_kernel(const float* __restrict__ video, const int64_t* __restrict__ index,
                                                         uint4* __restrict__ x0, int T, int HW, int RI0, int Wo0, int part, int ncdhw) {
    int q = blockIdx.x;
    const int par = q & 1; q >>= 1;
    const int c = q % 3; q /= 3;
    const int tp = q % (T + 2);
    const int64_t b = q / (T + 2);
    const int t = tp - 1;
    uint4* dst = x0 + (int64_t)blockIdx.x * RI0 * Wo0;
    const int n = RI0 * Wo0;
    if (t < 0 || t >= T) {                                   // temporal halo frame: zeros
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const int64_t src = index ? index[b] : b;
    const int64_t plane = ncdhw ? (src * 3 + c) * T + t : (src * T + t) * 3 + c;     // (B,3,T,H,W) or (B,T,3,H,W)
    const float* pl = video + plane * HW * (int64_t)HW;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int row = i / Wo0, wo = i - row * Wo0;
        const int h = par ? 2 * row - 3 : 2 * row - 2;
        uint16_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0;
        if (h >= 0 && h < HW) {
            const float* p = pl + (int64_t)h * HW;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const int w = 2 * wo + k - 3;
                if (w >= 0 && w < HW) v[k] = f2bf(bf16_part(__ldg(p + w), part));
            }
        }
        uint4 o;
        o.x = v[0] | ((uint32_t)v[1] << 16); o.y = v[2] | ((uint32_t)v[3] << 16);
        o.z = v[4] | ((uint32_t)v[5] << 16); o.w = v[6] | ((uint32_t)v[7] << 16);
        dst[i] = o;
    }
}

// uint8 frames (B,T,3,H,W) -> X0 with the dataset normalisation fused in: v = (u / 255 - mean[c]) / std[c] in fp32 (the same three
// IEEE operations ToTensor + Normalize perform on the host, utils.py:214-230 of the reference), rounded to bf16.  A host-resident
// real set then crosses PCIe as one byte per element instead of four.
struct NormU8 { float mean[3]; float stdv[3]; };

__global__ void __launch_bounds__(256) pack_video_u8_kernel(const uint8_t* __restrict__ video, const int64_t* __restrict__ index,
                                                            uint4* __restrict__ x0, int T, int HW, int RI0, int Wo0, NormU8 nm) {
    int q = blockIdx.x;
    const int par = q & 1; q >>= 1;
    const int c = q % 3; q /= 3;
    const int tp = q % (T + 2);
    const int64_t b = q / (T + 2);
    const int t = tp - 1;
    uint4* dst = x0 + (int64_t)blockIdx.x * RI0 * Wo0;
    const int n = RI0 * Wo0;
    if (t < 0 || t >= T) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const int64_t src = index ? index[b] : b;
    const uint8_t* pl = video + ((src * T + t) * 3 + c) * HW * (int64_t)HW;
    const float mean = nm.mean[c], stdv = nm.stdv[c];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int row = i / Wo0, wo = i - row * Wo0;
        const int h = par ? 2 * row - 3 : 2 * row - 2;
        uint16_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0;
        if (h >= 0 && h < HW) {
            const uint8_t* p = pl + (int64_t)h * HW;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const int w = 2 * wo + k - 3;
                if (w >= 0 && w < HW) v[k] = f2bf(__fdiv_rn(__fsub_rn(__fdiv_rn((float)__ldg(p + w), 255.f), mean), stdv));
            }
        }
        uint4 o;
        o.x = v[0] | ((uint32_t)v[1] << 16); o.y = v[2] | ((uint32_t)v[3] << 16);
        o.z = v[4] | ((uint32_t)v[5] << 16); o.w = v[6] | ((uint32_t)v[7] << 16);
        dst[i] = o;
    }
}

// conv 0 image: [p 11][k 2][blk 5][64][8] bf16; blk 0,4 = zero, blk b = W[kt = 3-b]
__global__ void pack_w0_kernel(const float* __restrict__ w, uint16_t* __restrict__ img, int part) {
    const int total = kW0Bytes / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int e = i % 8; int q = i / 8;
        int row = q % 64; q /= 64;
        int blk = q % 5; q /= 5;
        int k = q % 2; int p = q / 2;
        const int ch = 2 * p + k;
        float v = 0.f;
        if (ch < 21 && e < 7 && blk >= 1 && blk <= 3) {
            const int c = ch / 7, kh = l0_chunk_kh(ch % 7), kt = 3 - blk;
            v = bf16_part(w[(((row * 3 + c) * 3 + kt) * 7 + kh) * 7 + e], part);
        }
        img[i] = f2bf(v);
    }
}

// conv 1 image: [kt 3][slice 4][kh 7][kw 7][k 2][128][8]
__global__ void pack_w1_kernel(const float* __restrict__ w, uint16_t* __restrict__ img, int part) {
    const int total = 588 * kWeightTileBytes / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int e = i % 8; int q = i / 8;
        int row = q % 128; q /= 128;
        int k = q % 2; q /= 2;
        int kw = q % 7; q /= 7;
        int kh = q % 7; q /= 7;
        int slice = q % 4; int kt = q / 4;
        const int ci = slice * 16 + k * 8 + e;
        img[i] = f2bf(bf16_part(w[(((row * 64 + ci) * 3 + kt) * 7 + kh) * 7 + kw], part));
    }
}

// conv 2 image: [kh 7][kw 7][half 2][kt 3][kc 4][k 2][128][8]
__global__ void pack_w2_kernel(const float* __restrict__ w, uint16_t* __restrict__ img, int part) {
    const int total = 1176 * kWeightTileBytes / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int e = i % 8; int q = i / 8;
        int row = q % 128; q /= 128;
        int k = q % 2; q /= 2;
        int kc = q % 4; q /= 4;
        int kt = q % 3; q /= 3;
        int half = q % 2; q /= 2;
        int kw = q % 7; int kh = q / 7;
        const int ci = half * 64 + kc * 16 + k * 8 + e;
        img[i] = f2bf(bf16_part(w[(((row * 128 + ci) * 3 + kt) * 7 + kh) * 7 + kw], part));
    }
}

// ---- split-fp16 forward operands (SGeo, tc_layout.h) ----
// X0s (B, T+2, part 2, 3, 2, RI0, Wo0) chunks of 8 fp16: part 0 = fp16(v), part 1 = fp16(v - part 0); parts = 1: the hi-only layout
// X0h (B, T+2, 3, 2, RI0, Wo0) of the two-product mode (operand of vd_tc_x3_conv_layer_ex, passes = 2).  U8: uint8 frames with the
// dataset normalisation fused in.
// One block per (video, padded frame, channel, row parity) plane pair (both parts).  A thread owns two adjacent chunks (wo = 2p,
// 2p + 1) of a row: their 7-wide windows cover source columns 4p-3 .. 4p+5, i.e. the three ALIGNED quads at 4p-4, 4p, 4p+4 — three
// 128-bit loads (three 32-bit loads of uint8 frames) feed four 16-byte stores, instead of seven scalar loads per store.
constexpr int kPackPairs = 32;           // chunk pairs per block row (Wo0 / 2 <= 32)
constexpr int kPackRows = 8;             // rows per pass

template <bool U8>
__global__ void __launch_bounds__(kPackPairs * kPackRows) pack_video_x3_kernel(
        const void* __restrict__ video, const int64_t* __restrict__ index, uint4* __restrict__ x0, int T, int HW, int RI0, int Wo0,
        NormU8 nm, int parts) {
    int q = blockIdx.x;
    const int par = q & 1; q >>= 1;
    const int c = q % 3; q /= 3;
    const int tp = q % (T + 2);
    const int64_t b = q / (T + 2);
    const int t = tp - 1;
    const int n = RI0 * Wo0;                                                  // chunks per plane
    // planes [b][t_pad][part][c][par]
    uint4* dst_hi = x0 + ((((b * (T + 2) + tp) * parts + 0) * 3 + c) * 2 + par) * (int64_t)n;
    uint4* dst_lo = parts == 2 ? x0 + ((((b * (T + 2) + tp) * parts + 1) * 3 + c) * 2 + par) * (int64_t)n : nullptr;
    const int tid = threadIdx.y * kPackPairs + threadIdx.x;
    if (t < 0 || t >= T) {                                                    // temporal halo frame: zeros
        for (int i = tid; i < n; i += kPackPairs * kPackRows) {
            dst_hi[i] = make_uint4(0u, 0u, 0u, 0u);
            if (dst_lo) dst_lo[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        return;
    }
    const int64_t src = index ? index[b] : b;
    const int64_t plane = ((src * T + t) * 3 + c) * HW * (int64_t)HW;
    const float mean = nm.mean[c], stdv = nm.stdv[c];
    const int p = threadIdx.x;
    if (2 * p >= Wo0) return;
    const bool second = 2 * p + 1 < Wo0;
    for (int row = threadIdx.y; row < RI0; row += kPackRows) {
        const int h = par ? 2 * row - 3 : 2 * row - 2;
        float v[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) v[k] = 0.f;
        if (h >= 0 && h < HW) {
#pragma unroll
            for (int qd = 0; qd < 3; ++qd) {
                const int w = 4 * p - 4 + 4 * qd;
                if (w < 0 || w >= HW) continue;
                if (U8) {
                    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>((const uint8_t*)video + plane + (int64_t)h * HW + w));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        v[4 * qd + k] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)((u >> (8 * k)) & 255u), 255.f), mean), stdv);
                } else {
                    const float4 f = __ldg(reinterpret_cast<const float4*>((const float*)video + plane + (int64_t)h * HW + w));
                    v[4 * qd] = f.x; v[4 * qd + 1] = f.y; v[4 * qd + 2] = f.z; v[4 * qd + 3] = f.w;
                }
            }
        }
        uint16_t hi[12], lo[12];
#pragma unroll
        for (int k = 1; k < 10; ++k) split_h(v[k], hi[k], lo[k]);
        // chunk wo = 2p: columns 4p-3 .. 4p+3 = v[1..7]; chunk wo = 2p+1: columns 4p-1 .. 4p+5 = v[3..9]; 8th element (kw = 7) = 0
        const int o = row * Wo0 + 2 * p;
        dst_hi[o] = make_uint4(hi[1] | ((uint32_t)hi[2] << 16), hi[3] | ((uint32_t)hi[4] << 16), hi[5] | ((uint32_t)hi[6] << 16), hi[7]);
        if (second) dst_hi[o + 1] = make_uint4(hi[3] | ((uint32_t)hi[4] << 16), hi[5] | ((uint32_t)hi[6] << 16), hi[7] | ((uint32_t)hi[8] << 16), hi[9]);
        if (dst_lo) {
            dst_lo[o] = make_uint4(lo[1] | ((uint32_t)lo[2] << 16), lo[3] | ((uint32_t)lo[4] << 16), lo[5] | ((uint32_t)lo[6] << 16), lo[7]);
            if (second) dst_lo[o + 1] = make_uint4(lo[3] | ((uint32_t)lo[4] << 16), lo[5] | ((uint32_t)lo[6] << 16), lo[7] | ((uint32_t)lo[8] << 16), lo[9]);
        }
    }
}

__device__ __forceinline__ uint16_t h_part(float v, int part) {
    uint16_t hi, lo;
    split_h(v, hi, lo);
    return part ? lo : hi;
}

// conv 0 image (M-stacked): [kt 3][step 11][k 2][row 128][8]; row r = 32*q + l -> channel 16*q + (l & 15), part l >> 4
__global__ void pack_w0s_kernel(const float* __restrict__ w, uint16_t* __restrict__ img) {
    const int total = 3 * kW0Steps * kWeightTileBytes / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int e = i % 8; int q = i / 8;
        int row = q % 128; q /= 128;
        int k = q % 2; q /= 2;
        int step = q % kW0Steps; int kt = q / kW0Steps;
        const int ch = 2 * step + k;
        const int co = (row >> 5) * 16 + (row & 15), part = (row >> 4) & 1;
        uint16_t v = 0;
        if (ch < 21 && e < 7) {
            const int c = ch / 7, kh = l0_chunk_kh(ch % 7);
            v = h_part(w[(((co * 3 + c) * 3 + kt) * 7 + kh) * 7 + e], part);
        }
        img[i] = v;
    }
}

// conv 1 image: [kt 3][chunk 8][pair 25][part 2 (hi, lo)][k 2][128][8]; pair 0 = (tap 0, zeros), pair p = taps l1s_tap(2p-1), l1s_tap(2p)
__global__ void pack_w1s_kernel(const float* __restrict__ w, uint16_t* __restrict__ img) {
    const int total = 3 * 8 * kWTiles1s * kWeightTileBytes / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int e = i % 8; int q = i / 8;
        int row = q % 128; q /= 128;
        int k = q % 2; q /= 2;
        int part = q % 2; q /= 2;
        int pair = q % 25; q /= 25;
        int chunk = q % 8; int kt = q / 8;
        const int idx = pair ? 2 * pair - 1 + k : (k ? -1 : 0);
        uint16_t v = 0;
        if (idx >= 0) {
            const int tap = l1s_tap(idx), kh = tap / 7, kw = tap % 7;
            v = h_part(w[(((row * 64 + chunk * 8 + e) * 3 + kt) * 7 + kh) * 7 + kw], part);
        }
        img[i] = v;
    }
}

// conv 2 image: [kh 7][kw 7][quarter 4][kt 3][pair 2][part 2 (hi, lo)][k 2][128][8]; K half k of pair p = chunk 2p + k
__global__ void pack_w2s_kernel(const float* __restrict__ w, uint16_t* __restrict__ img) {
    const int total = 49 * 4 * kWTiles2s * kWeightTileBytes / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int e = i % 8; int q = i / 8;
        int row = q % 128; q /= 128;
        int k = q % 2; q /= 2;
        int part = q % 2; q /= 2;
        int pair = q % 2; q /= 2;
        int kt = q % 3; q /= 3;
        int quarter = q % 4; q /= 4;
        int kw = q % 7; int kh = q / 7;
        const int ci = quarter * 32 + (2 * pair + k) * 8 + e;
        img[i] = h_part(w[(((row * 128 + ci) * 3 + kt) * 7 + kh) * 7 + kw], part);
    }
}

}  // namespace tc
}  // namespace vd

using namespace vd;
using namespace vd::tc;

static int pack_video_impl(const float* video, const int64_t* index, void* x0, const vd_tc_plan* plan, int B, int part,
                           int ncdhw, void* stream) {
    VD_REQUIRE(video && x0 && plan, "tc_pack_video: NULL pointer");
    VD_REQUIRE(geo_supported(plan->T, plan->H), "tc_pack_video: unsupported geometry");
    if (B <= 0) return 0;
    const Geo g = make_geo(plan->T, plan->H);
    const int64_t blocks = (int64_t)B * (g.T + 2) * 6;                       // planes [b][t_pad][c][par]
    VD_REQUIRE(blocks < (1ll << 31), "tc_pack_video: grid too large");
    pack_video_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(video, index, (uint4*)x0, g.T, g.HW, g.RI0, g.Wo0, part, ncdhw);
    return check_launch("tc_pack_video");
}

extern "C" int vd_tc_pack_video_u8(const uint8_t* video, const int64_t* index, void* x0, const vd_tc_plan* plan, int B,
                                   const float* mean3, const float* std3, void* stream) {
    VD_REQUIRE(video && x0 && plan && B >= 0 && mean3 && std3, "tc_pack_video_u8: bad argument");
    VD_REQUIRE(geo_supported(plan->T, plan->H), "tc_pack_video_u8: unsupported geometry");
    VD_REQUIRE(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "tc_pack_video_u8: std must be non-zero");
    if (B == 0) return 0;
    const Geo g = make_geo(plan->T, plan->H);
    NormU8 nm;
    for (int c = 0; c < 3; ++c) { nm.mean[c] = mean3[c]; nm.stdv[c] = std3[c]; }
    const int64_t blocks = (int64_t)B * (g.T + 2) * 6;
    VD_REQUIRE(blocks < (1ll << 31), "tc_pack_video_u8: grid too large");
    pack_video_u8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(video, index, (uint4*)x0, g.T, g.HW, g.RI0, g.Wo0, nm);
    return check_launch("tc_pack_video_u8");
}

extern "C" int vd_tc_pack_video(const float* video, const int64_t* index, void* x0, const vd_tc_plan* plan,
                                int B, void* stream) {
    return pack_video_impl(video, index, x0, plan, B, 0, 0, stream);
}

// conv-trio variant: x is fp32 NCDHW (B,3,T,H,W); part = 0 value / 1 bf16 residual
extern "C" int vd_tc_pack_video_ncdhw(const float* x, void* x0, const vd_tc_plan* plan, int B, int part, void* stream) {
    VD_REQUIRE(part == 0 || part == 1, "tc_pack_video_ncdhw: part must be 0 or 1");
    return pack_video_impl(x, nullptr, x0, plan, B, part, 1, stream);
}

static int pack_weights_impl(const float* w_l0, const float* w_l1, const float* w_l2, void* w0, void* w1, void* w2, int part,
                             void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (w_l0 && w0) {
        pack_w0_kernel<<<148, 256, 0, s>>>(w_l0, (uint16_t*)w0, part);
        if (int e = check_launch("tc_pack_w0")) return e;
    }
    if (w_l1 && w1) {
        pack_w1_kernel<<<148 * 4, 256, 0, s>>>(w_l1, (uint16_t*)w1, part);
        if (int e = check_launch("tc_pack_w1")) return e;
    }
    if (w_l2 && w2) {
        pack_w2_kernel<<<148 * 4, 256, 0, s>>>(w_l2, (uint16_t*)w2, part);
        if (int e = check_launch("tc_pack_w2")) return e;
    }
    return 0;
}

extern "C" int vd_tc_pack_weights(const float* w_l0, const float* w_l1, const float* w_l2, void* w0, void* w1,
                                  void* w2, void* stream) {
    return pack_weights_impl(w_l0, w_l1, w_l2, w0, w1, w2, 0, stream);
}

extern "C" int vd_tc_pack_weights_part(const float* w_l0, const float* w_l1, const float* w_l2, void* w0, void* w1,
                                       void* w2, int part, void* stream) {
    VD_REQUIRE(part == 0 || part == 1, "tc_pack_weights_part: part must be 0 or 1");
    return pack_weights_impl(w_l0, w_l1, w_l2, w0, w1, w2, part, stream);
}

// ---- split-fp16 forward operands (vd_tc_x3_*) ----
static int pack_video_x3_impl(const void* video, bool u8, const int64_t* index, void* x0s, const vd_tc_plan* plan, int B,
                              const float* mean3, const float* std3, void* stream, int parts = 2) {
    VD_REQUIRE(video && x0s && plan && B >= 0, "tc_x3_pack_video: bad argument");
    VD_REQUIRE(geo_supported(plan->T, plan->H), "tc_x3_pack_video: unsupported geometry");
    if (B == 0) return 0;
    const Geo g = make_geo(plan->T, plan->H);
    NormU8 nm;
    for (int c = 0; c < 3; ++c) { nm.mean[c] = mean3 ? mean3[c] : 0.f; nm.stdv[c] = std3 ? std3[c] : 1.f; }
    const int64_t blocks = (int64_t)B * (g.T + 2) * 6;                       // (b, t_pad, c, par): both parts of a plane in one block
    VD_REQUIRE(blocks < (1ll << 31), "tc_x3_pack_video: grid too large");
    VD_REQUIRE(g.Wo0 <= 2 * kPackPairs && g.HW % 4 == 0 && ((uintptr_t)video & 15) == 0, "tc_x3_pack_video: unsupported width / unaligned videos");
    const dim3 block(kPackPairs, kPackRows, 1);
    if (u8) pack_video_x3_kernel<true><<<(unsigned)blocks, block, 0, (cudaStream_t)stream>>>(video, index, (uint4*)x0s, g.T, g.HW, g.RI0, g.Wo0, nm, parts);
    else pack_video_x3_kernel<false><<<(unsigned)blocks, block, 0, (cudaStream_t)stream>>>(video, index, (uint4*)x0s, g.T, g.HW, g.RI0, g.Wo0, nm, parts);
    return check_launch("tc_x3_pack_video");
}

// hi-only operand X0h of the two-product mode (plan->x0_bytes_per_video bytes per video: the X0 layout with fp16 values)
extern "C" int vd_tc_x3_pack_video_hi(const float* video, const int64_t* index, void* x0h, const vd_tc_plan* plan, int B, void* stream) {
    return pack_video_x3_impl(video, false, index, x0h, plan, B, nullptr, nullptr, stream, 1);
}

extern "C" int vd_tc_x3_pack_video_hi_u8(const uint8_t* video, const int64_t* index, void* x0h, const vd_tc_plan* plan, int B,
                                         const float* mean3, const float* std3, void* stream) {
    VD_REQUIRE(mean3 && std3 && std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "tc_x3_pack_video_hi_u8: mean / non-zero std required");
    return pack_video_x3_impl(video, true, index, x0h, plan, B, mean3, std3, stream, 1);
}

extern "C" int vd_tc_x3_pack_video(const float* video, const int64_t* index, void* x0s, const vd_tc_plan* plan, int B, void* stream) {
    return pack_video_x3_impl(video, false, index, x0s, plan, B, nullptr, nullptr, stream);
}

extern "C" int vd_tc_x3_pack_video_u8(const uint8_t* video, const int64_t* index, void* x0s, const vd_tc_plan* plan, int B,
                                      const float* mean3, const float* std3, void* stream) {
    VD_REQUIRE(mean3 && std3 && std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "tc_x3_pack_video_u8: mean / non-zero std required");
    return pack_video_x3_impl(video, true, index, x0s, plan, B, mean3, std3, stream);
}

extern "C" int vd_tc_x3_pack_weights(const float* w_l0, const float* w_l1, const float* w_l2, void* w0s, void* w1s, void* w2s,
                                     void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (w_l0 && w0s) {
        pack_w0s_kernel<<<148, 256, 0, s>>>(w_l0, (uint16_t*)w0s);
        if (int e = check_launch("tc_x3_pack_w0")) return e;
    }
    if (w_l1 && w1s) {
        pack_w1s_kernel<<<148 * 8, 256, 0, s>>>(w_l1, (uint16_t*)w1s);
        if (int e = check_launch("tc_x3_pack_w1")) return e;
    }
    if (w_l2 && w2s) {
        pack_w2s_kernel<<<148 * 8, 256, 0, s>>>(w_l2, (uint16_t*)w2s);
        if (int e = check_launch("tc_x3_pack_w2")) return e;
    }
    return 0;
}
