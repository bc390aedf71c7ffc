// TMA-fed composer kernels (utils.py:1178-1197 forward and its backward): the 3x3x3 stencils of the static-dynamic composer with
// their inputs staged by tensor-map TMA (cp.async.bulk.tensor -> UTMALDG) instead of per-thread cp.async.
//
// A block owns (video b, band of kTH image rows, full width) and walks the T frames.  Every staged plane is ONE box of a tensor
// map — rows h0-1 .. h0+kTH, columns -4 .. WP-5 — issued by one elected thread and completing on an mbarrier; everything outside
// the tensor (the halo rows / columns of the image, the frames -1 and T of the clip) is zero-filled by the TMA unit, so there is
// no per-thread address arithmetic, no halo clearing and no zero-frame special case.  Frames stream through a ring of kSlots
// slots (three in use, two in flight).
//
// The stencils themselves run on the fp32 FMA pipe and are bound by it, not by HBM (forward: 81 FMA per output pixel against
// 16.7 B; backward: 162 FMA against 20.7 B; the B200 ridge is 5.5 FMA per byte): the dynamic-channel weights live in registers
// (81 per thread) so that the inner loops are FFMA + one LDS per 12 FMAs.
#include <cuda.h>

#include "tc_common.cuh"

namespace vd {

using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_try_wait;
using tc::smem_u32;

namespace {

constexpr int kTH = 8;                 // image rows per block
constexpr int kCT = 256;               // threads per block
constexpr int kSlots = 5;              // frame ring

// ---- tensor maps (driver entry point through the runtime: the library does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        (void)cudaGetLastError();
    }
    return fn;
}

// fp32 tensor (outer.., H, W) with `rank` dimensions (innermost first in dims / box); out-of-range elements read as zero
int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("composer: cuTensorMapEncodeTiled is not available from this driver"); return -1; }
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    uint64_t stride = 4;
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1;
        stride *= dims[i];
        if (i + 1 < rank) gs[i] = stride;                       // byte stride of dimension i + 1
    }
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("composer: cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
    return 0;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// the 6 columns w0-1 .. w0+4 of one staged row (row pointer at column index 0 of the plane; image column w at index w + 4)
__device__ __forceinline__ void load_row6(const float* __restrict__ row, int w0, float (&x)[6]) {
    const float4 m = *reinterpret_cast<const float4*>(row + 4 + w0);
    x[0] = row[3 + w0]; x[1] = m.x; x[2] = m.y; x[3] = m.z; x[4] = m.w; x[5] = row[8 + w0];
}

// (Measured and rejected: taking the two edge columns from the neighbouring lanes by shuffle instead of the two scalar LDS —
// whose stride of 4 words is a 4-way bank conflict — is 1.5x slower: the shuffle latency lands on the FMA critical path and the
// extra live registers cost a resident block per SM.)

// ------------------------------------------------------------------------------------------ forward
//   out[b,t,o,h,w] = bias[o] + sum_taps ( sum_{i<3} Wt[o,i,tap] S[b,i,.,.] [t' valid] + Wt[o,3,tap] D[b,t',.,.] )
// The static term only depends on which temporal taps are valid (first / interior / last frame): it is evaluated when that
// mask changes (<= 3 times per block), not per frame.
__global__ void __launch_bounds__(kCT) compose_fwd_tma_kernel(
        const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmD,
        const int64_t* __restrict__ static_idx, const int64_t* __restrict__ label, const int64_t* __restrict__ dynamic_idx,
        const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ out, int T, int H, int W, int dpc, int WP,
        int TC) {
    extern __shared__ __align__(128) uint8_t cmp_smem[];
    const uint32_t base = (smem_u32(cmp_smem) + 127u) & ~127u;
    float* smem = reinterpret_cast<float*>(cmp_smem + (base - smem_u32(cmp_smem)));
    const int plane = (kTH + 2) * WP;
    float* Sp = smem;                                    // [3][kTH+2][WP]
    float* Dr = Sp + 3 * plane;                          // ring [kSlots][kTH+2][WP]
    float4* ws = reinterpret_cast<float4*>(Dr + kSlots * plane);      // [3 i][9 (kh,kw)] static weights summed over the valid kt
    uint64_t* bars = reinterpret_cast<uint64_t*>(ws + 27);            // [0]: static image, [1 + slot]: frame ring
    __shared__ float sw[324];
    __shared__ float sb[3];
    const int b = blockIdx.y, h0 = blockIdx.x * kTH;
    const int tb = blockIdx.z * TC, te = min(T, tb + TC);       // this block's frames [tb, te): the clip is cut into chunks of TC frames
    const int64_t HW = (int64_t)H * W;                          // so that the grid fills whole waves (host: pick_chunk)
    const int srow = (int)static_idx[b], drow = (int)(label[b] * dpc + dynamic_idx[b]);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t plane_bytes = (uint32_t)plane * 4u;
    if (threadIdx.x == 0) {
        for (int i = 0; i <= kSlots; ++i) mbar_init(bar0 + 8u * i, 1);
        tc::fence_mbar_init();
    }
    for (int i = threadIdx.x; i < 324; i += kCT) sw[i] = weight[i];
    if (threadIdx.x < 3) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    auto issue_frame = [&](int f) {                      // frame f in tb-1 .. te (out-of-range frames arrive as zeros)
        const int n = f + 1 - tb, slot = n % kSlots;
        mbar_expect_tx(bar0 + 8u * (1 + slot), plane_bytes);
        tma_load_4d(smem_u32(Dr + slot * plane), &tmD, -4, h0 - 1, f, drow, bar0 + 8u * (1 + slot));
    };
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar0, 3 * plane_bytes);
        tma_load_4d(smem_u32(Sp), &tmS, -4, h0 - 1, 0, srow, bar0);
        for (int f = tb - 1; f <= tb + 2 && f <= te; ++f) issue_frame(f);
    }
    // dynamic-channel weights in registers: wd[tap][o]
    float wd[27][3];
#pragma unroll
    for (int tap = 0; tap < 27; ++tap)
#pragma unroll
        for (int o = 0; o < 3; ++o) wd[tap][o] = sw[(o * 4 + 3) * 27 + tap];
    const int vpr = W >> 2;
    const int r = threadIdx.x / vpr, w0 = (threadIdx.x - r * vpr) * 4;
    const bool active = r < kTH && h0 + r < H;
    const int rr = r;
    float stat[3][4];
    int cur_mask = -1;
    bar_wait(bar0, 0);
    bar_wait(bar0 + 8u * 1, 0);                          // frame -1
    bar_wait(bar0 + 8u * 2, 0);                          // frame 0
    for (int t = tb; t < te; ++t) {
        { const int n = t - tb + 2; bar_wait(bar0 + 8u * (1 + n % kSlots), (uint32_t)(n / kSlots) & 1u); }     // frame t+1
        __syncthreads();                                 // everyone finished frame t-1: the slot of frame t-2 is free
        if (threadIdx.x == 0 && t + 3 <= te) issue_frame(t + 3);
        const int mask = (t >= 1 ? 1 : 0) | 2 | (t + 1 < T ? 4 : 0);          // bit kt: frame t+kt-1 exists
        if (mask != cur_mask) {                          // block-uniform: first / interior / last frame
            cur_mask = mask;
            if (threadIdx.x < 27) {
                const int i = threadIdx.x / 9, k9 = threadIdx.x - i * 9;
                float s[3] = {0.f, 0.f, 0.f};
                for (int kt = 0; kt < 3; ++kt)
                    if (mask & (1 << kt))
                        for (int o = 0; o < 3; ++o) s[o] += sw[(o * 4 + i) * 27 + kt * 9 + k9];
                ws[threadIdx.x] = make_float4(s[0], s[1], s[2], 0.f);
            }
            __syncthreads();
            if (active) {
#pragma unroll
                for (int o = 0; o < 3; ++o)
#pragma unroll
                    for (int j = 0; j < 4; ++j) stat[o][j] = sb[o];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        float x[6];
                        load_row6(Sp + i * plane + (r + kh) * WP, w0, x);
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const float4 wv = ws[i * 9 + kh * 3 + kw];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                stat[0][j] = fmaf(wv.x, x[j + kw], stat[0][j]);
                                stat[1][j] = fmaf(wv.y, x[j + kw], stat[1][j]);
                                stat[2][j] = fmaf(wv.z, x[j + kw], stat[2][j]);
                            }
                        }
                    }
            }
            __syncthreads();                             // ws is rewritten at the next mask change
        }
        if (!active) continue;
        float acc[3][4];
#pragma unroll
        for (int o = 0; o < 3; ++o)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[o][j] = stat[o][j];
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            const float* P = Dr + ((t - tb + kt) % kSlots) * plane;           // frame t+kt-1 -> n = t-tb+kt
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                float x[6];
                load_row6(P + (rr + kh) * WP, w0, x);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int tap = (kt * 3 + kh) * 3 + kw;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[0][j] = fmaf(wd[tap][0], x[j + kw], acc[0][j]);
                        acc[1][j] = fmaf(wd[tap][1], x[j + kw], acc[1][j]);
                        acc[2][j] = fmaf(wd[tap][2], x[j + kw], acc[2][j]);
                    }
                }
            }
        }
        if (!active) continue;
        float* ob = out + (((int64_t)b * T + t) * 3) * HW + (int64_t)(h0 + r) * W + w0;
#pragma unroll
        for (int o = 0; o < 3; ++o)
            *reinterpret_cast<float4*>(ob + o * HW) = make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
    }
}

// ------------------------------------------------------------------------------------------ backward: fused, deterministic
// ONE pass over the video gradient g (B,T,3,H,W): per frame a block produces
//   * d dynamic[row(b), t, band]   (81-tap transposed stencil of the 3 g channels; plain stores when every video selects its own
//                                   memory row — distill_s2d_ms.py:405 guarantees that — else atomicAdd),
//   * the 81 dynamic-channel weight sums and the 3 bias sums (registers),
//   * the running frame sums of g, from which the 243 static-channel weight sums follow AFTER the loop (the static image is
//     t-invariant: sum_t g[t] * S collapses to three frame sums: all frames, all but the first, all but the last).
// The 327 sums leave the block as ONE row of `partial` (no floating-point atomics); compose_bwd_finish_kernel (compose_tiled.cu)
// adds the rows in block order: the hallucinator gradient is bitwise reproducible.
// Block reductions: every thread parks its partial sums in shared memory (the frame ring is free by then), value-major, and
// each warp adds the 256 partials of its values in a fixed order (8 per lane, then 5 shuffles) — 26 instructions per value
// instead of the 88 of a warp_sum per warp and value.
constexpr int kPartialStride = 328;

__global__ void __launch_bounds__(kCT) compose_bwd_tma_kernel(
        const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmD,
        const int64_t* __restrict__ static_idx, const int64_t* __restrict__ label, const int64_t* __restrict__ dynamic_idx,
        const float* __restrict__ weight, float* __restrict__ grad_dynamic, float* __restrict__ partial,
        int T, int H, int W, int dpc, int WP, int unique_rows) {
    extern __shared__ __align__(128) uint8_t cmp_smem[];
    const uint32_t base = (smem_u32(cmp_smem) + 127u) & ~127u;
    float* smem = reinterpret_cast<float*>(cmp_smem + (base - smem_u32(cmp_smem)));
    const int plane = (kTH + 2) * WP;
    float* Sp = smem;                                            // [3 i][kTH+2][WP] static image of the video
    float* Ring = Sp + 3 * plane;                                // [kSlots][4][kTH+2][WP]: g channels 0..2, then D, of one frame
    float4* wf = reinterpret_cast<float4*>(Ring + kSlots * 4 * plane);      // [27 (a,bb,cc)] flipped dynamic-channel weights {o0,o1,o2,-}
    uint64_t* bars = reinterpret_cast<uint64_t*>(wf + 27);
    const int b = blockIdx.y, h0 = blockIdx.x * kTH;
    const int64_t HW = (int64_t)H * W;
    const int64_t drow = label[b] * dpc + dynamic_idx[b];
    const int srow = (int)static_idx[b];
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t plane_bytes = (uint32_t)plane * 4u;
    if (threadIdx.x == 0) {
        for (int i = 0; i <= kSlots; ++i) mbar_init(bar0 + 8u * i, 1);
        tc::fence_mbar_init();
    }
    if (threadIdx.x < 27) {
        // input offset (a,bb,cc) in {0,1,2}^3 <-> tap (kt,kh,kw) = (2-a, 2-bb, 2-cc)
        const int a = threadIdx.x / 9, bb = (threadIdx.x / 3) % 3, cc = threadIdx.x % 3;
        const int tap = ((2 - a) * 3 + (2 - bb)) * 3 + (2 - cc);
        wf[threadIdx.x] = make_float4(weight[(0 * 4 + 3) * 27 + tap], weight[(1 * 4 + 3) * 27 + tap], weight[(2 * 4 + 3) * 27 + tap], 0.f);
    }
    __syncthreads();
    auto issue_frame = [&](int f) {                              // frame f in -1 .. T (out-of-range frames arrive as zeros)
        const int slot = (f + 1) % kSlots;
        const uint32_t bar = bar0 + 8u * (1 + slot);
        mbar_expect_tx(bar, 4 * plane_bytes);
        tma_load_5d(smem_u32(Ring + slot * 4 * plane), &tmG, -4, h0 - 1, 0, f, b, bar);
        tma_load_4d(smem_u32(Ring + (slot * 4 + 3) * plane), &tmD, -4, h0 - 1, f, (int)drow, bar);
    };
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar0, 3 * plane_bytes);
        tma_load_4d(smem_u32(Sp), &tmS, -4, h0 - 1, 0, srow, bar0);
        for (int f = -1; f <= 2 && f <= T; ++f) issue_frame(f);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vpr = W >> 2;
    const int r = threadIdx.x / vpr, w0 = (threadIdx.x - r * vpr) * 4;      // (row, 4 columns)
    const bool active = r < kTH && h0 + r < H;
    const int rr = r;
    float acc[84];                                               // [o][tap] dynamic-channel weight sums, then the 3 bias sums
#pragma unroll
    for (int k = 0; k < 84; ++k) acc[k] = 0.f;
    float gs[3][4], gf[3][4], gl[3][4];                          // frame sums of g, first frame, last frame
#pragma unroll
    for (int o = 0; o < 3; ++o)
#pragma unroll
        for (int j = 0; j < 4; ++j) { gs[o][j] = 0.f; gf[o][j] = 0.f; gl[o][j] = 0.f; }
    float* gd = grad_dynamic + drow * (int64_t)T * HW;
    bar_wait(bar0 + 8u * 1, 0);                                  // frame -1
    bar_wait(bar0 + 8u * 2, 0);                                  // frame 0
    for (int t = 0; t < T; ++t) {
        { const int n = t + 2; bar_wait(bar0 + 8u * (1 + n % kSlots), (uint32_t)(n / kSlots) & 1u); }     // frame t+1
        __syncthreads();                                         // everyone finished frame t-1: the slot of frame t-2 is free
        if (threadIdx.x == 0 && t + 3 <= T) issue_frame(t + 3);
        if (!active) continue;
        float g[3][4];
        const float* Gc = Ring + ((t + 1) % kSlots) * 4 * plane;  // frame t
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            float4 v = *reinterpret_cast<const float4*>(Gc + o * plane + (rr + 1) * WP + 4 + w0);
            g[o][0] = v.x; g[o][1] = v.y; g[o][2] = v.z; g[o][3] = v.w;
            acc[81 + o] += (v.x + v.y) + (v.z + v.w);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                gs[o][j] += g[o][j];
                if (t == 0) gf[o][j] = g[o][j];
                if (t == T - 1) gl[o][j] = g[o][j];
            }
        }
        // dynamic-channel weights: g[t] x D[t + kt - 1]
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            const float* P = Ring + (((t + kt) % kSlots) * 4 + 3) * plane;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                float x[6];
                load_row6(P + (rr + kh) * WP, w0, x);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int tap = (kt * 3 + kh) * 3 + kw;
#pragma unroll
                    for (int o = 0; o < 3; ++o) {
                        float sacc = acc[o * 27 + tap];
#pragma unroll
                        for (int j = 0; j < 4; ++j) sacc = fmaf(g[o][j], x[j + kw], sacc);
                        acc[o * 27 + tap] = sacc;
                    }
                }
            }
        }
        // d dynamic[t]: transposed stencil over g[t-1 .. t+1]
        float d3[3][4];                                          // one accumulator set per g channel: 12 independent FMA chains
#pragma unroll
        for (int o = 0; o < 3; ++o)
#pragma unroll
            for (int j = 0; j < 4; ++j) d3[o][j] = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float* F = Ring + ((t + a) % kSlots) * 4 * plane;           // frame t-1+a
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) {
                float x0[6], x1[6], x2[6];
                load_row6(F + 0 * plane + (rr + bb) * WP, w0, x0);
                load_row6(F + 1 * plane + (rr + bb) * WP, w0, x1);
                load_row6(F + 2 * plane + (rr + bb) * WP, w0, x2);
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    const float4 wv = wf[(a * 3 + bb) * 3 + cc];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        d3[0][j] = fmaf(wv.x, x0[j + cc], d3[0][j]);
                        d3[1][j] = fmaf(wv.y, x1[j + cc], d3[1][j]);
                        d3[2][j] = fmaf(wv.z, x2[j + cc], d3[2][j]);
                    }
                }
            }
        }
        if (!active) continue;
        float dd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) dd[j] = (d3[0][j] + d3[1][j]) + d3[2][j];
        float* dst = gd + (int64_t)t * HW + (int64_t)(h0 + r) * W + w0;
        if (unique_rows) *reinterpret_cast<float4*>(dst) = make_float4(dd[0], dd[1], dd[2], dd[3]);
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(dst + j, dd[j]);
        }
    }
    // ---- block sums.  park[k][tid] (value-major, reusing the frame ring); warp w adds values k = w, w + 8, ...
    float* park = Ring;                                          // >= 84 * 256 floats (the ring holds 20 planes of >= 960)
    float* prow = partial + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * kPartialStride;
    bar_wait(bar0, 0);                                           // static image (needed below; landed long ago)
    __syncthreads();                                             // everyone left the frame loop: the ring is free
#pragma unroll
    for (int k = 0; k < 84; ++k) park[k * kCT + threadIdx.x] = acc[k];
    __syncthreads();
    for (int k = warp; k < 84; k += kCT / 32) {
        float sv = 0.f;
#pragma unroll
        for (int j = 0; j < kCT / 32; ++j) sv += park[k * kCT + j * 32 + lane];
        sv = warp_sum(sv);
        if (lane == 0) {
            if (k < 81) prow[((k / 27) * 4 + 3) * 27 + k % 27] = sv;
            else prow[324 + (k - 81)] = sv;
        }
    }
    // ---- static-channel weights: tap kt reads frame t + kt - 1, so kt = 0 needs t >= 1 and kt = 2 needs t <= T - 2
#pragma unroll 1
    for (int okt = 0; okt < 9; ++okt) {
        const int o = okt / 3, kt = okt - o * 3;
        float gk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float all = o == 0 ? gs[0][j] : o == 1 ? gs[1][j] : gs[2][j];
            const float fst = o == 0 ? gf[0][j] : o == 1 ? gf[1][j] : gf[2][j];
            const float lst = o == 0 ? gl[0][j] : o == 1 ? gl[1][j] : gl[2][j];
            gk[j] = active ? (kt == 0 ? all - fst : kt == 1 ? all : all - lst) : 0.f;
        }
        __syncthreads();                                         // the previous round's partials have been read
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                float x[6];
                load_row6(Sp + i * plane + ((active ? r : 0) + kh) * WP, active ? w0 : 0, x);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    float sv = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) sv = fmaf(gk[j], x[j + kw], sv);
                    park[((i * 3 + kh) * 3 + kw) * kCT + threadIdx.x] = sv;
                }
            }
        __syncthreads();
        for (int k = warp; k < 27; k += kCT / 32) {
            float sv = 0.f;
#pragma unroll
            for (int j = 0; j < kCT / 32; ++j) sv += park[k * kCT + j * 32 + lane];
            sv = warp_sum(sv);
            if (lane == 0) {
                const int i = k / 9, kh = (k / 3) % 3, kw = k % 3;
                prow[(o * 4 + i) * 27 + (kt * 3 + kh) * 3 + kw] = sv;
            }
        }
    }
    if (threadIdx.x == 0) prow[327] = 0.f;
}

int set_smem(const void* fn, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(compose tma): %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

bool tma_enabled() {                                     // VD_COMPOSE_TMA=0: the cp.async kernels of compose_tiled.cu
    static int on = -1;
    if (on < 0) { const char* v = getenv("VD_COMPOSE_TMA"); on = (v && *v) ? atoi(v) : 1; }
    return on != 0;
}

int tma_wp(int W) { return (W + 8 + 31) / 32 * 32; }       // 128-byte rows: every staged plane starts 128-byte aligned
bool tma_ok(int H, int W) { return W % 4 == 0 && W >= 8 && (W / 4) * kTH <= kCT && tma_wp(W) <= 256 && H >= 1; }

// frames per block: the (band, video) tiles rarely fill whole waves of `slots` co-resident blocks (50 videos x 14 bands = 700
// tiles on 296 slots = 2.4 waves -> 3), so the clip may be cut into 2 or 4 chunks (each re-reads two halo frames)
int pick_chunk(int tiles, int T, int slots) {
    int best = T;
    double best_cost = 1e30;
    for (int parts = 1; parts <= 4; parts *= 2) {
        if (T % parts || T / parts < 4) continue;
        const int tc_ = T / parts;
        const double waves = (double)((int64_t)tiles * parts + slots - 1) / slots;
        const double cost = (double)(int64_t)waves * (tc_ + 1.0);              // + ~1 frame of prologue per block
        if (cost < best_cost) { best_cost = cost; best = tc_; }
    }
    return best;
}

int sm_count() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

}  // namespace

// returns 1 when the geometry is not covered or tensor maps are unavailable (the caller falls back to the cp.async kernels)
// n_static / n_dynamic: the number of (3,H,W) static images and (T,H,W) dynamic memories behind the pointers.  The tensor maps
// carry the TRUE extents: with the row dimension left open (1 << 22 rows) the kernel faulted with "illegal memory access" in
// some allocator states although every coordinate it requests is in range (reproduced, and gone with exact extents).
int compose_fwd_tma(const float* static_syn, const float* dynamic_syn, const int64_t* static_idx, const int64_t* label,
                    const int64_t* dynamic_idx, const float* weight, const float* bias, float* out, int B, int T, int H,
                    int W, int dpc, int64_t n_static, int64_t n_dynamic, cudaStream_t stream) {
    if (n_static <= 0 || n_dynamic <= 0) return 1;
    if (!tma_ok(H, W) || !encode_fn() || !tma_enabled()) return 1;
    if ((((uintptr_t)static_syn) | ((uintptr_t)dynamic_syn)) & 15) return 1;
    const int WP = tma_wp(W);
    CUtensorMap tmS, tmD;
    const uint64_t dS[4] = {(uint64_t)W, (uint64_t)H, 3, (uint64_t)n_static}, dD[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)T, (uint64_t)n_dynamic};
    const uint32_t bS[4] = {(uint32_t)WP, kTH + 2, 3, 1}, bD[4] = {(uint32_t)WP, kTH + 2, 1, 1};
    if (make_map(&tmS, static_syn, 4, dS, bS) || make_map(&tmD, dynamic_syn, 4, dD, bD)) return -1;
    const size_t smem = (size_t)(3 + kSlots) * (kTH + 2) * WP * 4 + 27 * 16 + 8 * (kSlots + 1) + 128;
    static size_t configured = 0;
    if (smem > configured) { if (int e = set_smem((const void*)compose_fwd_tma_kernel, smem)) return e; configured = smem; }
    const int nb = (int)ceil_div(H, kTH);
    const int TC = pick_chunk(nb * B, T, 2 * sm_count());              // 128 registers x 256 threads: two blocks per SM
    dim3 grid((unsigned)nb, (unsigned)B, (unsigned)ceil_div(T, TC));
    compose_fwd_tma_kernel<<<grid, kCT, smem, stream>>>(tmS, tmD, static_idx, label, dynamic_idx, weight, bias, out, T, H, W, dpc, WP, TC);
    if (int e = check_launch("compose_fwd_tma")) {
        set_error("compose_fwd_tma failed (%d): B=%d T=%d H=%d W=%d dpc=%d WP=%d TC=%d smem=%zu static=%p dynamic=%p out=%p idx=%p/%p/%p",
                  e, B, T, H, W, dpc, WP, TC, smem, (const void*)static_syn, (const void*)dynamic_syn, (void*)out,
                  (const void*)static_idx, (const void*)label, (const void*)dynamic_idx);
        return e;
    }
    return 0;
}

}  // namespace vd

namespace vd {

void compose_bwd_finish(const float* scratch, int n_rows, float* grad_weight, float* grad_bias, cudaStream_t stream);   // compose_tiled.cu

// returns 1 when the geometry is not covered or tensor maps are unavailable.  scratch: >= B * ceil(H / kTH) * 328 floats.
int compose_bwd_tma(const float* gout, const float* static_syn, const float* dynamic_syn, const int64_t* static_idx,
                    const int64_t* label, const int64_t* dynamic_idx, const float* weight, float* grad_dynamic, float* grad_weight,
                    float* grad_bias, float* scratch, int64_t scratch_floats, int unique_rows, int B, int T, int H, int W, int dpc,
                    int64_t n_static, int64_t n_dynamic, cudaStream_t stream) {
    if (!tma_ok(H, W) || !encode_fn() || !tma_enabled() || n_static <= 0 || n_dynamic <= 0) return 1;
    if ((((uintptr_t)static_syn) | ((uintptr_t)dynamic_syn) | ((uintptr_t)gout)) & 15) return 1;
    const int WP = tma_wp(W);
    if ((size_t)kSlots * 4 * (kTH + 2) * WP < (size_t)84 * kCT) return 1;            // the parked partial sums reuse the frame ring
    const int nb = (int)ceil_div(H, kTH);
    if (scratch_floats < (int64_t)B * nb * kPartialStride) { set_error("compose_bwd_fused: scratch too small"); return -1; }
    CUtensorMap tmG, tmS, tmD;
    const uint64_t dG[5] = {(uint64_t)W, (uint64_t)H, 3, (uint64_t)T, (uint64_t)B};
    const uint64_t dS[4] = {(uint64_t)W, (uint64_t)H, 3, (uint64_t)n_static}, dD[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)T, (uint64_t)n_dynamic};
    const uint32_t bG[5] = {(uint32_t)WP, kTH + 2, 3, 1, 1}, bS[4] = {(uint32_t)WP, kTH + 2, 3, 1}, bD[4] = {(uint32_t)WP, kTH + 2, 1, 1};
    if (make_map(&tmG, gout, 5, dG, bG) || make_map(&tmS, static_syn, 4, dS, bS) || make_map(&tmD, dynamic_syn, 4, dD, bD)) return -1;
    const size_t smem = (size_t)(3 + 4 * kSlots) * (kTH + 2) * WP * 4 + 27 * 16 + 8 * (kSlots + 1) + 128;
    static size_t configured = 0;
    if (smem > configured) { if (int e = set_smem((const void*)compose_bwd_tma_kernel, smem)) return e; configured = smem; }
    dim3 grid((unsigned)nb, (unsigned)B, 1);
    compose_bwd_tma_kernel<<<grid, kCT, smem, stream>>>(tmG, tmS, tmD, static_idx, label, dynamic_idx, weight, grad_dynamic, scratch,
                                                       T, H, W, dpc, WP, unique_rows);
    if (int e = check_launch("compose_bwd_tma")) return e;
    compose_bwd_finish(scratch, B * nb, grad_weight, grad_bias, stream);
    return check_launch("compose_bwd_finish");
}

}  // namespace vd
