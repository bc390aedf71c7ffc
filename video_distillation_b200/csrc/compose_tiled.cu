// Shared-memory tiled composer kernels (utils.py:1178-1197 forward, and its backward): the three 3x3x3
// stencils of the static-dynamic composer as HBM-streaming kernels.
//
// One block owns (video b, band of TH image rows, full width) and walks the T frames with a ring of four
// frame slots in shared memory: while frame t is computed, frame t+2 streams in with cp.async (16 bytes per
// request, zero rows outside the image).  Every input element is therefore read from HBM/L2 once per
// band (+2 halo rows) and all stencil reads are LDS.  Thread = (row, 4 consecutive columns): one aligned
// LDS.128 + two scalar LDS give the 6 input columns of a row.
//
//   fwd   : out[b,t,o,h,w] = bias[o] + sum_taps ( sum_{i<3} Wt[o,i,tap] S[b,i,.,.] [t' valid] + Wt[o,3,tap] D[b,t',.,.] )
//           the static term only depends on which temporal taps are valid (first / interior / last frame):
//           it is evaluated when that mask changes (<= 3 times per block), not per frame.
//   bwd d : d_dynamic[row(b),t',h',w'] += sum_{o,taps} Wt[o,3,tap] g[b,t'-kt+1,o,h'-kh+1,w'-kw+1]
//   bwd w : d_Wt[o,3,tap] += sum_{b,t,h,w} g[b,t,o,h,w] D[b,t+kt-1,h+kh-1,w+kw-1];  d_bias[o] += sum g
#include "common.cuh"

namespace vd {

constexpr int kTH = 8;                 // image rows per block
constexpr int kCT = 256;               // threads per block

__device__ __forceinline__ void cpa16(float* dst_smem, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cpa_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// rows [h0-1, h0+kTH+1) x cols [0,W) of one (H,W) fp32 plane -> smem plane [kTH+2][WP], data at column 4;
// rows outside the image (or the whole plane when src == nullptr) are zero filled.
__device__ __forceinline__ void stage_plane(float* __restrict__ dst, const float* __restrict__ src, int h0, int H, int W, int WP) {
    const int vpr = W >> 2;
    for (int i = threadIdx.x; i < (kTH + 2) * vpr; i += kCT) {
        const int r = i / vpr, v = i - r * vpr, h = h0 - 1 + r;
        float* d = dst + r * WP + 4 + 4 * v;
        if (src != nullptr && (unsigned)h < (unsigned)H) cpa16(d, src + (int64_t)h * W + 4 * v);
        else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// the 6 columns w0-1 .. w0+4 of one staged row (row pointer at column index 0 of the plane)
__device__ __forceinline__ void load_row6(const float* __restrict__ row, int w0, float (&x)[6]) {
    const float4 m = *reinterpret_cast<const float4*>(row + 4 + w0);
    x[0] = row[3 + w0]; x[1] = m.x; x[2] = m.y; x[3] = m.z; x[4] = m.w; x[5] = row[8 + w0];
}

__device__ __forceinline__ void zero_halo_columns(float* planes, int n_planes, int W, int WP) {
    for (int i = threadIdx.x; i < n_planes * (kTH + 2); i += kCT) {
        planes[i * WP + 3] = 0.f;
        planes[i * WP + 4 + W] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(kCT) compose_fwd_tiled_kernel(
        const float* __restrict__ static_syn, const float* __restrict__ dynamic_syn,
        const int64_t* __restrict__ static_idx, const int64_t* __restrict__ label,
        const int64_t* __restrict__ dynamic_idx, const float* __restrict__ weight,
        const float* __restrict__ bias, float* __restrict__ out, int T, int H, int W, int dpc, int WP) {
    extern __shared__ float4 cmp_smem4[];
    float* smem = reinterpret_cast<float*>(cmp_smem4);
    const int plane = (kTH + 2) * WP;
    float* Sp = smem;                         // [3][kTH+2][WP]
    float* Dr = Sp + 3 * plane;               // ring [4][kTH+2][WP]
    float4* wd = reinterpret_cast<float4*>(Dr + 4 * plane);      // [27] dynamic weights {o0,o1,o2,-}
    float4* ws = wd + 27;                     // [3 i][9 (kh,kw)] static weights summed over the valid kt
    __shared__ float sw[324];
    __shared__ float sb[3];
    const int b = blockIdx.y, h0 = blockIdx.x * kTH;
    const int64_t HW = (int64_t)H * W;
    const float* S = static_syn + static_idx[b] * 3 * HW;
    const float* D = dynamic_syn + (label[b] * dpc + dynamic_idx[b]) * (int64_t)T * HW;
    for (int i = threadIdx.x; i < 324; i += kCT) sw[i] = weight[i];
    if (threadIdx.x < 3) sb[threadIdx.x] = bias[threadIdx.x];
    zero_halo_columns(smem, 7, W, WP);
    for (int i = 0; i < 3; ++i) stage_plane(Sp + i * plane, S + i * HW, h0, H, W, WP);
    stage_plane(Dr + 3 * plane, nullptr, h0, H, W, WP);                       // frame -1
    stage_plane(Dr + 0 * plane, D, h0, H, W, WP);                             // frame 0
    stage_plane(Dr + 1 * plane, T > 1 ? D + HW : nullptr, h0, H, W, WP);      // frame 1
    cpa_commit();
    __syncthreads();
    if (threadIdx.x < 27) {
        const int tap = threadIdx.x;
        wd[tap] = make_float4(sw[(0 * 4 + 3) * 27 + tap], sw[(1 * 4 + 3) * 27 + tap], sw[(2 * 4 + 3) * 27 + tap], 0.f);
    }
    const int vpr = W >> 2;
    const int r = threadIdx.x / vpr, w0 = (threadIdx.x - r * vpr) * 4;
    const bool active = r < kTH && h0 + r < H;
    float stat[3][4];
    int cur_mask = -1;
    for (int t = 0; t < T; ++t) {
        cpa_wait_all();
        __syncthreads();                                   // frames <= t+1 staged; everyone finished frame t-1
        if (t + 2 < T + 1) {                               // stage frame t+2 (zeros beyond the clip) into the slot of t-2
            stage_plane(Dr + ((t + 2) & 3) * plane, t + 2 < T ? D + (int64_t)(t + 2) * HW : nullptr, h0, H, W, WP);
            cpa_commit();
        }
        const int mask = (t >= 1 ? 1 : 0) | 2 | (t + 1 < T ? 4 : 0);          // bit kt: frame t+kt-1 exists
        if (mask != cur_mask) {                            // block-uniform: first / interior / last frame
            cur_mask = mask;
            __syncthreads();
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
        }
        if (!active) continue;
        float acc[3][4];
#pragma unroll
        for (int o = 0; o < 3; ++o)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[o][j] = stat[o][j];
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            const float* P = Dr + ((t + kt + 3) & 3) * plane;                 // frame t+kt-1
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                float x[6];
                load_row6(P + (r + kh) * WP, w0, x);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4 wv = wd[(kt * 3 + kh) * 3 + kw];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[0][j] = fmaf(wv.x, x[j + kw], acc[0][j]);
                        acc[1][j] = fmaf(wv.y, x[j + kw], acc[1][j]);
                        acc[2][j] = fmaf(wv.z, x[j + kw], acc[2][j]);
                    }
                }
            }
        }
        float* ob = out + (((int64_t)b * T + t) * 3) * HW + (int64_t)(h0 + r) * W + w0;
#pragma unroll
        for (int o = 0; o < 3; ++o)
            *reinterpret_cast<float4*>(ob + o * HW) = make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
    }
}

// ------------------------------------------------------------------------------------------ backward: data
__global__ void __launch_bounds__(kCT) compose_bwd_data_tiled_kernel(
        const float* __restrict__ gout, const int64_t* __restrict__ label, const int64_t* __restrict__ dynamic_idx,
        const float* __restrict__ weight, float* __restrict__ grad_dynamic, int T, int H, int W, int dpc, int WP) {
    extern __shared__ float4 cmp_smem4[];
    float* smem = reinterpret_cast<float*>(cmp_smem4);
    const int plane = (kTH + 2) * WP;
    float* Gr = smem;                                            // ring [4 frames][3 o][kTH+2][WP]
    float4* wf = reinterpret_cast<float4*>(Gr + 12 * plane);     // [27 (a,bb,cc)] flipped weights {o0,o1,o2,-}
    const int b = blockIdx.y, h0 = blockIdx.x * kTH;
    const int64_t HW = (int64_t)H * W;
    const float* G = gout + (int64_t)b * T * 3 * HW;
    zero_halo_columns(smem, 12, W, WP);
    if (threadIdx.x < 27) {
        // input offset (a,bb,cc) in {0,1,2}^3 <-> tap (kt,kh,kw) = (2-a, 2-bb, 2-cc)
        const int a = threadIdx.x / 9, bb = (threadIdx.x / 3) % 3, cc = threadIdx.x % 3;
        const int tap = ((2 - a) * 3 + (2 - bb)) * 3 + (2 - cc);
        wf[threadIdx.x] = make_float4(weight[(0 * 4 + 3) * 27 + tap], weight[(1 * 4 + 3) * 27 + tap], weight[(2 * 4 + 3) * 27 + tap], 0.f);
    }
    auto stage_frame = [&](int f) {                              // f may be -1 or >= T: zeros
        float* dst = Gr + ((f + 4) & 3) * 3 * plane;
        const bool ok = (unsigned)f < (unsigned)T;
        for (int o = 0; o < 3; ++o) stage_plane(dst + o * plane, ok ? G + ((int64_t)f * 3 + o) * HW : nullptr, h0, H, W, WP);
    };
    stage_frame(-1); stage_frame(0); stage_frame(1);
    cpa_commit();
    const int vpr = W >> 2;
    const int r = threadIdx.x / vpr, w0 = (threadIdx.x - r * vpr) * 4;
    const bool active = r < kTH && h0 + r < H;
    float* gd = grad_dynamic + (label[b] * dpc + dynamic_idx[b]) * (int64_t)T * HW;
    for (int t = 0; t < T; ++t) {
        cpa_wait_all();
        __syncthreads();
        if (t + 2 <= T) { stage_frame(t + 2); cpa_commit(); }
        if (!active) continue;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float* F = Gr + ((t + a + 3) & 3) * 3 * plane;              // frame t-1+a
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) {
                float x0[6], x1[6], x2[6];
                load_row6(F + 0 * plane + (r + bb) * WP, w0, x0);
                load_row6(F + 1 * plane + (r + bb) * WP, w0, x1);
                load_row6(F + 2 * plane + (r + bb) * WP, w0, x2);
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    const float4 wv = wf[(a * 3 + bb) * 3 + cc];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        acc[j] = fmaf(wv.x, x0[j + cc], fmaf(wv.y, x1[j + cc], fmaf(wv.z, x2[j + cc], acc[j])));
                }
            }
        }
        float* dst = gd + (int64_t)t * HW + (int64_t)(h0 + r) * W + w0;
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(dst + j, acc[j]);               // rows may be shared between videos (vpc > 1)
    }
}

// ------------------------------------------------------------------------------------------ backward: dynamic-channel weights
__global__ void __launch_bounds__(kCT) compose_bwd_wdyn_tiled_kernel(
        const float* __restrict__ gout, const float* __restrict__ dynamic_syn, const int64_t* __restrict__ label,
        const int64_t* __restrict__ dynamic_idx, float* __restrict__ grad_weight, float* __restrict__ grad_bias,
        int T, int H, int W, int dpc, int WP) {
    extern __shared__ float4 cmp_smem4[];
    float* smem = reinterpret_cast<float*>(cmp_smem4);
    const int plane = (kTH + 2) * WP;
    float* Dr = smem;                          // ring [4][kTH+2][WP] of D frames
    float* Gb = Dr + 4 * plane;                // double buffer [2][3 o][kTH+2][WP] of g frames (rows 1..kTH used)
    __shared__ float red[8 * 84];
    const int b = blockIdx.y, h0 = blockIdx.x * kTH;
    const int64_t HW = (int64_t)H * W;
    const float* G = gout + (int64_t)b * T * 3 * HW;
    const float* D = dynamic_syn + (label[b] * dpc + dynamic_idx[b]) * (int64_t)T * HW;
    zero_halo_columns(smem, 4, W, WP);
    auto stage_d = [&](int f) { stage_plane(Dr + ((f + 4) & 3) * plane, (unsigned)f < (unsigned)T ? D + (int64_t)f * HW : nullptr, h0, H, W, WP); };
    auto stage_g = [&](int f) {
        if (f >= T) return;
        for (int o = 0; o < 3; ++o) stage_plane(Gb + ((f & 1) * 3 + o) * plane, G + ((int64_t)f * 3 + o) * HW, h0, H, W, WP);
    };
    stage_d(-1); stage_d(0); stage_d(1); stage_g(0);
    cpa_commit();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = warp, w0 = lane * 4;                     // one warp per image row, 4 columns per lane
    const bool active = h0 + r < H && w0 < W;
    float acc[84];                                         // [o][tap] then the 3 bias sums
#pragma unroll
    for (int k = 0; k < 84; ++k) acc[k] = 0.f;
    for (int t = 0; t < T; ++t) {
        cpa_wait_all();
        __syncthreads();
        if (t + 2 <= T) stage_d(t + 2);
        stage_g(t + 1);
        cpa_commit();
        if (!active) continue;
        float g[3][4];
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            const float4 v = *reinterpret_cast<const float4*>(Gb + ((t & 1) * 3 + o) * plane + (r + 1) * WP + 4 + w0);
            g[o][0] = v.x; g[o][1] = v.y; g[o][2] = v.z; g[o][3] = v.w;
            acc[81 + o] += (v.x + v.y) + (v.z + v.w);
        }
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            const float* P = Dr + ((t + kt + 3) & 3) * plane;                 // frame t+kt-1 (zeros outside the clip)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                float x[6];
                load_row6(P + (r + kh) * WP, w0, x);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int tap = (kt * 3 + kh) * 3 + kw;
#pragma unroll
                    for (int o = 0; o < 3; ++o) {
                        float s = acc[o * 27 + tap];
#pragma unroll
                        for (int j = 0; j < 4; ++j) s = fmaf(g[o][j], x[j + kw], s);
                        acc[o * 27 + tap] = s;
                    }
                }
            }
        }
    }
    constexpr int n = 84;
#pragma unroll
    for (int k = 0; k < n; ++k) {
        const float s = warp_sum(acc[k]);
        if (lane == 0) red[warp * n + k] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += kCT) {
        float s = 0.f;
        for (int w = 0; w < kCT / 32; ++w) s += red[w * n + k];
        if (k < 81) atomicAdd(grad_weight + ((k / 27) * 4 + 3) * 27 + k % 27, s);
        else if (grad_bias) atomicAdd(grad_bias + (k - 81), s);
    }
}

// ------------------------------------------------------------------------------------------ backward: fused, deterministic
// ONE pass over the video gradient g (B,T,3,H,W): a block owns (video, band of kTH rows) and streams the T frames of g and of the
// video's dynamic memory D through cp.async rings; per frame it produces
//   * d dynamic[row(b), t, band]   (81-tap transposed stencil of the 3 g channels, plain stores when every video selects its own
//                                   memory row — distill_s2d_ms.py:405 guarantees that — else atomicAdd),
//   * the 81 dynamic-channel weight sums and the 3 bias sums (registers),
//   * the running frame sums of g, from which the 243 static-channel weight sums follow AFTER the loop (the static image is
//     t-invariant: sum_t g[t] * S collapses to three frame sums: all frames, all but the first, all but the last).
// The 327 sums leave the block as ONE row of `partial` (no floating-point atomics); compose_bwd_finish_kernel adds the rows in
// block order: the hallucinator gradient is bitwise reproducible.
constexpr int kPartialStride = 328;
constexpr int kFR = 5;                 // frame slots of the fused backward's rings

__global__ void __launch_bounds__(kCT) compose_bwd_fused_kernel(
        const float* __restrict__ gout, const float* __restrict__ static_syn, const float* __restrict__ dynamic_syn,
        const int64_t* __restrict__ static_idx, const int64_t* __restrict__ label, const int64_t* __restrict__ dynamic_idx,
        const float* __restrict__ weight, float* __restrict__ grad_dynamic, float* __restrict__ partial,
        int T, int H, int W, int dpc, int WP, int unique_rows) {
    extern __shared__ float4 cmp_smem4[];
    float* smem = reinterpret_cast<float*>(cmp_smem4);
    const int plane = (kTH + 2) * WP;
    float* Gr = smem;                                            // ring [kFR frames][3 o][kTH+2][WP]
    float* Dr = Gr + 3 * kFR * plane;                            // ring [kFR frames][kTH+2][WP]
    float* Sp = Dr + kFR * plane;                                // [3 i][kTH+2][WP] static image of the video
    float4* wf = reinterpret_cast<float4*>(Sp + 3 * plane);      // [27 (a,bb,cc)] flipped dynamic-channel weights {o0,o1,o2,-}
    __shared__ float red[8 * 84];
    const int b = blockIdx.y, h0 = blockIdx.x * kTH;
    const int64_t HW = (int64_t)H * W;
    const float* G = gout + (int64_t)b * T * 3 * HW;
    const int64_t drow = label[b] * dpc + dynamic_idx[b];
    const float* D = dynamic_syn + drow * (int64_t)T * HW;
    const float* S = static_syn + static_idx[b] * 3 * HW;
    zero_halo_columns(smem, 4 * kFR + 3, W, WP);
    if (threadIdx.x < 27) {
        const int a = threadIdx.x / 9, bb = (threadIdx.x / 3) % 3, cc = threadIdx.x % 3;
        const int tap = ((2 - a) * 3 + (2 - bb)) * 3 + (2 - cc);
        wf[threadIdx.x] = make_float4(weight[(0 * 4 + 3) * 27 + tap], weight[(1 * 4 + 3) * 27 + tap], weight[(2 * 4 + 3) * 27 + tap], 0.f);
    }
    auto stage_frame = [&](int f) {                              // f may be -1 or >= T: zeros
        const bool ok = (unsigned)f < (unsigned)T;
        float* dst = Gr + ((f + kFR) % kFR) * 3 * plane;
        for (int o = 0; o < 3; ++o) stage_plane(dst + o * plane, ok ? G + ((int64_t)f * 3 + o) * HW : nullptr, h0, H, W, WP);
        stage_plane(Dr + ((f + kFR) % kFR) * plane, ok ? D + (int64_t)f * HW : nullptr, h0, H, W, WP);
    };
    // ring of kFR = 5 frame slots: frames t-1 .. t+1 are read while t+2 (waited for at the NEXT iteration) and t+3 (just issued)
    // stream in — two frames of prefetch; with four slots and cp.async.wait_all every iteration waited for the copy it had
    // issued one iteration earlier, i.e. for a full global-memory round trip per frame
    for (int i = 0; i < 3; ++i) stage_plane(Sp + i * plane, S + i * HW, h0, H, W, WP);
    stage_frame(-1); stage_frame(0); stage_frame(1);
    cpa_commit();
    stage_frame(2);
    cpa_commit();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vpr = W >> 2;
    const int r = threadIdx.x / vpr, w0 = (threadIdx.x - r * vpr) * 4;      // (row, 4 columns): W / 4 threads per row, all lanes of a warp busy
    const bool active = r < kTH && h0 + r < H;
    float acc[84];                                               // [o][tap] dynamic-channel weight sums, then the 3 bias sums
#pragma unroll
    for (int k = 0; k < 84; ++k) acc[k] = 0.f;
    float gs[3][4], gf[3][4], gl[3][4];                          // frame sums of g, first frame, last frame
#pragma unroll
    for (int o = 0; o < 3; ++o)
#pragma unroll
        for (int j = 0; j < 4; ++j) { gs[o][j] = 0.f; gf[o][j] = 0.f; gl[o][j] = 0.f; }
    float* gd = grad_dynamic + drow * (int64_t)T * HW;
    for (int t = 0; t < T; ++t) {
        cpa_wait_but_one();                                      // frames <= t+1 have landed (t+2 may still be in flight)
        __syncthreads();
        if (t + 3 <= T) stage_frame(t + 3);                      // into the slot of frame t-2, which nobody reads any more
        cpa_commit();                                            // (one group per iteration, possibly empty, keeps the count)
        if (!active) continue;
        float g[3][4];
        const float* Gc = Gr + (t % kFR) * 3 * plane;            // frame t
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            const float4 v = *reinterpret_cast<const float4*>(Gc + o * plane + (r + 1) * WP + 4 + w0);
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
            const float* P = Dr + ((t + kt - 1 + kFR) % kFR) * plane;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                float x[6];
                load_row6(P + (r + kh) * WP, w0, x);
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
        float dd[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float* F = Gr + ((t + a - 1 + kFR) % kFR) * 3 * plane;      // frame t-1+a
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) {
                float x0[6], x1[6], x2[6];
                load_row6(F + 0 * plane + (r + bb) * WP, w0, x0);
                load_row6(F + 1 * plane + (r + bb) * WP, w0, x1);
                load_row6(F + 2 * plane + (r + bb) * WP, w0, x2);
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    const float4 wv = wf[(a * 3 + bb) * 3 + cc];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        dd[j] = fmaf(wv.x, x0[j + cc], fmaf(wv.y, x1[j + cc], fmaf(wv.z, x2[j + cc], dd[j])));
                }
            }
        }
        float* dst = gd + (int64_t)t * HW + (int64_t)(h0 + r) * W + w0;
        if (unique_rows) *reinterpret_cast<float4*>(dst) = make_float4(dd[0], dd[1], dd[2], dd[3]);
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(dst + j, dd[j]);
        }
    }
    // ---- block sums: dynamic-channel weights + bias
    float* prow = partial + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * kPartialStride;
#pragma unroll
    for (int k = 0; k < 84; ++k) {
        const float sv = warp_sum(acc[k]);
        if (lane == 0) red[warp * 84 + k] = sv;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 84; k += kCT) {
        float sv = 0.f;
        for (int wv = 0; wv < kCT / 32; ++wv) sv += red[wv * 84 + k];
        if (k < 81) prow[((k / 27) * 4 + 3) * 27 + k % 27] = sv;
        else prow[324 + (k - 81)] = sv;
    }
    // ---- static-channel weights: tap kt reads frame t + kt - 1, so kt = 0 needs t >= 1 and kt = 2 needs t <= T - 2
#pragma unroll
    for (int o = 0; o < 3; ++o)
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            float gk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) gk[j] = active ? (kt == 0 ? gs[o][j] - gf[o][j] : kt == 1 ? gs[o][j] : gs[o][j] - gl[o][j]) : 0.f;
            float a27[27];
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
                        a27[(i * 3 + kh) * 3 + kw] = sv;
                    }
                }
            __syncthreads();                                     // `red` is reused
#pragma unroll
            for (int k = 0; k < 27; ++k) {
                const float sv = warp_sum(a27[k]);
                if (lane == 0) red[warp * 27 + k] = sv;
            }
            __syncthreads();
            if (threadIdx.x < 27) {
                float sv = 0.f;
                for (int wv = 0; wv < kCT / 32; ++wv) sv += red[wv * 27 + threadIdx.x];
                const int i = threadIdx.x / 9, kh = (threadIdx.x / 3) % 3, kw = threadIdx.x % 3;
                prow[(o * 4 + i) * 27 + (kt * 3 + kh) * 3 + kw] = sv;
            }
        }
    if (threadIdx.x == 0) prow[327] = 0.f;
}

// grad_weight (324) / grad_bias (3) += sum over the blocks' partial rows, in block order
__global__ void compose_bwd_finish_kernel(const float* __restrict__ partial, int n_blocks, float* __restrict__ grad_weight,
                                          float* __restrict__ grad_bias) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 327) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;               // four interleaved chains (fixed order), then one fixed combine
    int bi = 0;
    for (; bi + 4 <= n_blocks; bi += 4) {
        s0 += partial[(int64_t)(bi + 0) * kPartialStride + k];
        s1 += partial[(int64_t)(bi + 1) * kPartialStride + k];
        s2 += partial[(int64_t)(bi + 2) * kPartialStride + k];
        s3 += partial[(int64_t)(bi + 3) * kPartialStride + k];
    }
    for (; bi < n_blocks; ++bi) s0 += partial[(int64_t)bi * kPartialStride + k];
    const float s = (s0 + s1) + (s2 + s3);
    if (k < 324) grad_weight[k] += s;
    else if (grad_bias) grad_bias[k - 324] += s;
}

// ------------------------------------------------------------------------------------------ host side
static inline int tiled_wp(int W) { return W + 8; }
static bool tiled_ok(int H, int W) { return W % 4 == 0 && W >= 8 && (W / 4) * kTH <= kCT && W <= 128 && H >= 1; }

static int set_smem(const void* fn, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(compose tiled): %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

// returns 1 when the geometry is not covered (caller falls back to the generic kernels)
int compose_fwd_tiled(const float* static_syn, const float* dynamic_syn, const int64_t* static_idx, const int64_t* label,
                      const int64_t* dynamic_idx, const float* weight, const float* bias, float* out, int B, int T, int H,
                      int W, int dpc, cudaStream_t stream) {
    if (!tiled_ok(H, W)) return 1;
    const int WP = tiled_wp(W);
    const size_t smem = (size_t)7 * (kTH + 2) * WP * 4 + 54 * 16;
    static size_t configured = 0;
    if (smem > configured) { if (int e = set_smem((const void*)compose_fwd_tiled_kernel, smem)) return e; configured = smem; }
    dim3 grid((unsigned)ceil_div(H, kTH), (unsigned)B, 1);
    compose_fwd_tiled_kernel<<<grid, kCT, smem, stream>>>(static_syn, dynamic_syn, static_idx, label, dynamic_idx, weight, bias, out,
                                                         T, H, W, dpc, WP);
    return check_launch("compose_fwd_tiled");
}

int compose_bwd_data_tiled(const float* gout, const int64_t* label, const int64_t* dynamic_idx, const float* weight,
                           float* grad_dynamic, int B, int T, int H, int W, int dpc, cudaStream_t stream) {
    if (!tiled_ok(H, W)) return 1;
    const int WP = tiled_wp(W);
    const size_t smem = (size_t)12 * (kTH + 2) * WP * 4 + 27 * 16;
    static size_t configured = 0;
    if (smem > configured) { if (int e = set_smem((const void*)compose_bwd_data_tiled_kernel, smem)) return e; configured = smem; }
    dim3 grid((unsigned)ceil_div(H, kTH), (unsigned)B, 1);
    compose_bwd_data_tiled_kernel<<<grid, kCT, smem, stream>>>(gout, label, dynamic_idx, weight, grad_dynamic, T, H, W, dpc, WP);
    return check_launch("compose_bwd_data_tiled");
}

int compose_bwd_wdyn_tiled(const float* gout, const float* dynamic_syn, const int64_t* label, const int64_t* dynamic_idx,
                           float* grad_weight, float* grad_bias, int B, int T, int H, int W, int dpc, cudaStream_t stream) {
    if (!tiled_ok(H, W) || kCT / 32 != kTH) return 1;
    const int WP = tiled_wp(W);
    const size_t smem = (size_t)10 * (kTH + 2) * WP * 4;
    static size_t configured = 0;
    if (smem > configured) { if (int e = set_smem((const void*)compose_bwd_wdyn_tiled_kernel, smem)) return e; configured = smem; }
    dim3 grid((unsigned)ceil_div(H, kTH), (unsigned)B, 1);
    compose_bwd_wdyn_tiled_kernel<<<grid, kCT, smem, stream>>>(gout, dynamic_syn, label, dynamic_idx, grad_weight, grad_bias,
                                                              T, H, W, dpc, WP);
    return check_launch("compose_bwd_wdyn_tiled");
}

// returns 1 when the geometry is not covered.  scratch: >= B * ceil(H / kTH) * 328 floats.
int compose_bwd_fused(const float* gout, const float* static_syn, const float* dynamic_syn, const int64_t* static_idx,
                      const int64_t* label, const int64_t* dynamic_idx, const float* weight, float* grad_dynamic, float* grad_weight,
                      float* grad_bias, float* scratch, int64_t scratch_floats, int unique_rows, int B, int T, int H, int W, int dpc,
                      cudaStream_t stream) {
    if (!tiled_ok(H, W) || kCT / 32 != kTH) return 1;
    const int WP = tiled_wp(W);
    const int nb = (int)ceil_div(H, kTH);
    if (scratch_floats < (int64_t)B * nb * kPartialStride) { set_error("compose_bwd_fused: scratch too small"); return -1; }
    const size_t smem = (size_t)(4 * kFR + 3) * (kTH + 2) * WP * 4 + 27 * 16;
    static size_t configured = 0;
    if (smem > configured) { if (int e = set_smem((const void*)compose_bwd_fused_kernel, smem)) return e; configured = smem; }
    dim3 grid((unsigned)nb, (unsigned)B, 1);
    compose_bwd_fused_kernel<<<grid, kCT, smem, stream>>>(gout, static_syn, dynamic_syn, static_idx, label, dynamic_idx, weight,
                                                         grad_dynamic, scratch, T, H, W, dpc, WP, unique_rows);
    if (int e = check_launch("compose_bwd_fused")) return e;
    compose_bwd_finish_kernel<<<3, 128, 0, stream>>>(scratch, B * nb, grad_weight, grad_bias);
    return check_launch("compose_bwd_finish");
}

void compose_bwd_finish(const float* scratch, int n_rows, float* grad_weight, float* grad_bias, cudaStream_t stream) {
    compose_bwd_finish_kernel<<<3, 128, 0, stream>>>(scratch, n_rows, grad_weight, grad_bias);
}

}  // namespace vd
