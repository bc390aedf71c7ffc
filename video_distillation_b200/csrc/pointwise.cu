// Memory-bound kernels of the distillation inner loop: ReLU+MaxPool routing, instancenorm,
// avgpool, the static-dynamic composer (fwd/bwd with the index gathers folded in), the
// per-class mean / distribution-matching loss, momentum SGD, axpy and squared distance.
#include "common.cuh"

namespace vd {

// =============================================================== ReLU + MaxPool3d routing
// window (pt,ph,pw) = stride; first maximum wins in (t,h,w) scan order (ATen: `>` from -inf,
// SURVEY App. A); code bit 3 = max > 0.
__global__ void relu_maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                        uint8_t* __restrict__ code, int64_t total, int T, int H, int W,
                                        int To, int Ho, int Wo, int pt, int ph, int pw) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        int wo = (int)(o % Wo); int64_t q = o / Wo;
        int ho = (int)(q % Ho); q /= Ho;
        int to = (int)(q % To); int64_t nc = q / To;
        const float* xb = x + nc * T * H * W + ((int64_t)(to * pt) * H + ho * ph) * W + wo * pw;
        float best = -INFINITY; int arg = 0, pos = 0;
        for (int a = 0; a < pt; ++a)
            for (int b = 0; b < ph; ++b)
                for (int c = 0; c < pw; ++c, ++pos) {
                    float v = xb[((int64_t)a * H + b) * W + c];
                    if (v > best || v != v) { best = v; arg = pos; }   // NaN propagates like ATen
                }
        const bool act = best > 0.f;
        y[o] = act ? best : (best != best ? best : 0.f);
        if (code) code[o] = (uint8_t)(arg | (act ? 8 : 0));
    }
}

__global__ void route_scatter_kernel(const float* __restrict__ gy, const uint8_t* __restrict__ code,
                                     float* __restrict__ gx, int64_t total, int T, int H, int W,
                                     int To, int Ho, int Wo, int pt, int ph, int pw) {
    // one thread per pooled output; writes its whole window (so gx needs no memset) when the
    // windows tile the input, tails (odd extents) are zeroed by the host-side memset.
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        int wo = (int)(o % Wo); int64_t q = o / Wo;
        int ho = (int)(q % Ho); q /= Ho;
        int to = (int)(q % To); int64_t nc = q / To;
        float* xb = gx + nc * T * H * W + ((int64_t)(to * pt) * H + ho * ph) * W + wo * pw;
        const uint8_t cd = code[o];
        const float g = (cd & 8) ? gy[o] : 0.f;
        int pos = 0;
        for (int a = 0; a < pt; ++a)
            for (int b = 0; b < ph; ++b)
                for (int c = 0; c < pw; ++c, ++pos) xb[((int64_t)a * H + b) * W + c] = (pos == (cd & 7)) ? g : 0.f;
    }
}

__global__ void route_gather_kernel(const float* __restrict__ x, const uint8_t* __restrict__ code,
                                    float* __restrict__ y, int64_t total, int T, int H, int W,
                                    int To, int Ho, int Wo, int pt, int ph, int pw) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        int wo = (int)(o % Wo); int64_t q = o / Wo;
        int ho = (int)(q % Ho); q /= Ho;
        int to = (int)(q % To); int64_t nc = q / To;
        const uint8_t cd = code[o];
        const int pos = cd & 7;
        const int c = pos % pw, b = (pos / pw) % ph, a = pos / (pw * ph);
        const float* xb = x + nc * T * H * W + ((int64_t)(to * pt + a) * H + ho * ph + b) * W + wo * pw + c;
        y[o] = (cd & 8) ? *xb : 0.f;
    }
}

// =============================================================== instancenorm(+ReLU), avgpool
__global__ void inorm_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float* __restrict__ y,
                                      float* __restrict__ mean, float* __restrict__ rstd, int C, int64_t S) {
    __shared__ float red[32];
    const int64_t nc = blockIdx.x;
    const int c = (int)(nc % C);
    const float* xp = x + nc * S;
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < S; i += blockDim.x) s += xp[i];
    const float mu = block_sum(s, red) / (float)S;
    float v = 0.f;
    for (int64_t i = threadIdx.x; i < S; i += blockDim.x) { float d = xp[i] - mu; v += d * d; }
    const float var = block_sum(v, red) / (float)S;
    const float rs = rsqrtf(var + 1e-5f);
    if (threadIdx.x == 0) { mean[nc] = mu; rstd[nc] = rs; }
    const float g = gamma[c] * rs, b = beta[c] - mu * g;
    float* yp = y + nc * S;
    for (int64_t i = threadIdx.x; i < S; i += blockDim.x) yp[i] = fmaxf(fmaf(xp[i], g, b), 0.f);
}

__global__ void inorm_relu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                      const float* __restrict__ gy, const float* __restrict__ gamma,
                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                      float* __restrict__ gx, float* __restrict__ ggamma,
                                      float* __restrict__ gbeta, int C, int64_t S) {
    __shared__ float red[32];
    const int64_t nc = blockIdx.x;
    const int c = (int)(nc % C);
    const float mu = mean[nc], rs = rstd[nc], g = gamma[c];
    const float* xp = x + nc * S; const float* yp = y + nc * S; const float* gp = gy + nc * S;
    float s1 = 0.f, s2 = 0.f;   // sum dz, sum dz*xhat   (dz = gy masked by relu)
    for (int64_t i = threadIdx.x; i < S; i += blockDim.x) {
        float dz = yp[i] > 0.f ? gp[i] : 0.f;
        s1 += dz; s2 += dz * (xp[i] - mu) * rs;
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) { atomicAdd(ggamma + c, s2); atomicAdd(gbeta + c, s1); }
    const float m1 = s1 / (float)S, m2 = s2 / (float)S;
    float* gxp = gx + nc * S;
    for (int64_t i = threadIdx.x; i < S; i += blockDim.x) {
        float dz = yp[i] > 0.f ? gp[i] : 0.f;
        float xh = (xp[i] - mu) * rs;
        gxp[i] = g * rs * (dz - m1 - xh * m2);
    }
}

__global__ void avgpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total,
                                    int T, int H, int W, int To, int Ho, int Wo) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        int wo = (int)(o % Wo); int64_t q = o / Wo;
        int ho = (int)(q % Ho); q /= Ho;
        int to = (int)(q % To); int64_t nc = q / To;
        const float* xb = x + nc * T * H * W + ((int64_t)(2 * to) * H + 2 * ho) * W + 2 * wo;
        float s = 0.f;
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) s += xb[((int64_t)a * H + b) * W + c];
        y[o] = s * 0.125f;
    }
}

// instancenorm + ReLU + AvgPool3d(2) in ONE launch (networks.py:784,757,772 as Features applies them): a block owns an (n, c)
// instance of S = T*H*W values.  Pass 1 (HBM, 128-bit loads): shifted first and second moments -> mean, rstd; pass 2 (the
// instance — 200 KB at most — comes back from L2): every thread normalises the 2x2x2 inputs of its pooled outputs (float2
// loads), applies ReLU and writes the mean: the normalised activation never exists in memory.  Algorithmic HBM bytes: 4*S read
// + S/2 written per instance (the unfused pair: 4*S + 4*S + 4*S + S/2).
__global__ void __launch_bounds__(512) inorm_relu_avgpool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, float* __restrict__ y,
                                                                     float* __restrict__ mean, float* __restrict__ rstd,
                                                                     int C, int T, int H, int W) {
    __shared__ float red[32];
    const int64_t nc = blockIdx.x;
    const int c = (int)(nc % C);
    const int64_t S = (int64_t)T * H * W;
    const float* xp = x + nc * S;
    const int64_t S4 = S >> 2;                                   // S % 4 == 0 (checked by the host)
    const float4* x4 = reinterpret_cast<const float4*>(xp);
    // ONE statistics pass: sums of (x - k) and (x - k)^2 around a sample k = x[0] of the instance (|k - mean| ~ sigma, so the
    // subtraction below cancels a factor of ~2, not the mean^2 / var of the raw-moment formula)
    const float k = __ldg(xp);
    float s = 0.f, q = 0.f;
#pragma unroll 4
    for (int64_t i = threadIdx.x; i < S4; i += blockDim.x) {
        const float4 v = __ldg(x4 + i);
        const float a0 = v.x - k, a1 = v.y - k, a2 = v.z - k, a3 = v.w - k;
        s += (a0 + a1) + (a2 + a3);
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    const float m1 = block_sum(s, red) / (float)S;
    const float m2 = block_sum(q, red) / (float)S;
    const float mu = k + m1;
    const float var = fmaxf(m2 - m1 * m1, 0.f);
    const float rs = rsqrtf(var + 1e-5f);
    if (threadIdx.x == 0 && mean) { mean[nc] = mu; rstd[nc] = rs; }
    const float g = gamma[c] * rs, b = beta[c] - mu * g;
    const int To = T / 2, Ho = H / 2, Wo = W / 2;
    float* yp = y + nc * (int64_t)To * Ho * Wo;
    const int HoWo = Ho * Wo;
    for (int o = threadIdx.x; o < To * HoWo; o += blockDim.x) {
        const int to = o / HoWo, rem = o - to * HoWo, ho = rem / Wo, wo = rem - ho * Wo;
        const float* p0 = xp + ((int64_t)(2 * to) * H + 2 * ho) * W + 2 * wo;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(p0 + ((int64_t)a * H + bb) * W));
                acc += fmaxf(fmaf(v.x, g, b), 0.f) + fmaxf(fmaf(v.y, g, b), 0.f);
            }
        yp[o] = acc * 0.125f;
    }
}

__global__ void avgpool2_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, int64_t total,
                                    int T, int H, int W, int To, int Ho, int Wo) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int w = (int)(i % W); int64_t q = i / W;
        int h = (int)(q % H); q /= H;
        int t = (int)(q % T); int64_t nc = q / T;
        int to = t >> 1, ho = h >> 1, wo = w >> 1;
        gx[i] = (to < To && ho < Ho && wo < Wo) ? 0.125f * gy[((nc * To + to) * Ho + ho) * Wo + wo] : 0.f;
    }
}

// =============================================================== composer
// out[b,t,o,h,w] = bias[o] + sum_{kt,kh,kw} ( sum_{i<3} Wt[o,i,kt,kh,kw] S[b,i,h',w'] [0<=t'<T]
//                                             + Wt[o,3,kt,kh,kw] D[b,t',h',w'] )
// with t'=t+kt-1, h'=h+kh-1, w'=w+kw-1, zero padded (utils.py:1186-1197).
// One thread per (b,t,h, 4 consecutive w): 128-bit stores, weights in shared memory.
__global__ void __launch_bounds__(256) compose_fwd_kernel(
        const float* __restrict__ static_syn, const float* __restrict__ dynamic_syn,
        const int64_t* __restrict__ static_idx, const int64_t* __restrict__ label,
        const int64_t* __restrict__ dynamic_idx, const float* __restrict__ weight,
        const float* __restrict__ bias, float* __restrict__ out, int T, int H, int W, int dpc) {
    __shared__ float sw[3 * 4 * 27];
    __shared__ float sb[3];
    for (int i = threadIdx.x; i < 324; i += blockDim.x) sw[i] = weight[i];
    if (threadIdx.x < 3) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int b = blockIdx.z, t = blockIdx.y;
    const int W4 = (W + 3) / 4;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * W4) return;
    const int h = idx / W4, w0 = (idx % W4) * 4;
    const float* S = static_syn + static_idx[b] * 3 * H * W;
    const float* D = dynamic_syn + (label[b] * dpc + dynamic_idx[b]) * (int64_t)T * H * W;
    float acc[3][4];
#pragma unroll
    for (int o = 0; o < 3; ++o)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[o][j] = sb[o];
    for (int kh = 0; kh < 3; ++kh) {
        const int hh = h + kh - 1;
        if ((unsigned)hh >= (unsigned)H) continue;
        // the 6 input columns w0-1 .. w0+4 of this row
        float sv[3][6], dv[3][6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int ww = w0 + j - 1;
            const bool ok = (unsigned)ww < (unsigned)W;
#pragma unroll
            for (int i = 0; i < 3; ++i) sv[i][j] = ok ? __ldg(S + ((int64_t)i * H + hh) * W + ww) : 0.f;
#pragma unroll
            for (int kt = 0; kt < 3; ++kt) {
                const int tt = t + kt - 1;
                dv[kt][j] = (ok && (unsigned)tt < (unsigned)T) ? __ldg(D + ((int64_t)tt * H + hh) * W + ww) : 0.f;
            }
        }
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            const int tt = t + kt - 1;
            const bool tok = (unsigned)tt < (unsigned)T;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int tap = (kt * 3 + kh) * 3 + kw;
#pragma unroll
                for (int o = 0; o < 3; ++o) {
                    const float wd = sw[(o * 4 + 3) * 27 + tap];
                    const float w0s = tok ? sw[(o * 4 + 0) * 27 + tap] : 0.f;
                    const float w1s = tok ? sw[(o * 4 + 1) * 27 + tap] : 0.f;
                    const float w2s = tok ? sw[(o * 4 + 2) * 27 + tap] : 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float a = acc[o][j];
                        a = fmaf(w0s, sv[0][j + kw], a);
                        a = fmaf(w1s, sv[1][j + kw], a);
                        a = fmaf(w2s, sv[2][j + kw], a);
                        a = fmaf(wd, dv[kt][j + kw], a);
                        acc[o][j] = a;
                    }
                }
            }
        }
    }
    float* ob = out + (((int64_t)b * T + t) * 3) * H * W + (int64_t)h * W + w0;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        float* p = ob + (int64_t)o * H * W;
        if (w0 + 3 < W && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
            *reinterpret_cast<float4*>(p) = make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
        } else {
            for (int j = 0; j < 4 && w0 + j < W; ++j) p[j] = acc[o][j];
        }
    }
}

// d_dynamic[row(b), t', h', w'] += sum_{o,kt,kh,kw} Wt[o,3,kt,kh,kw] g[b, t'-kt+1, o, h'-kh+1, w'-kw+1]
// (and optionally d_static[row, i, h', w'] += sum over t and taps with t' = t+kt-1 valid)
__global__ void __launch_bounds__(256) compose_bwd_data_kernel(
        const float* __restrict__ gout, const int64_t* __restrict__ static_idx,
        const int64_t* __restrict__ label, const int64_t* __restrict__ dynamic_idx,
        const float* __restrict__ weight, float* __restrict__ grad_dynamic,
        float* __restrict__ grad_static, int T, int H, int W, int dpc) {
    __shared__ float sw[324];
    for (int i = threadIdx.x; i < 324; i += blockDim.x) sw[i] = weight[i];
    __syncthreads();
    const int b = blockIdx.z, t = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * W) return;
    const int h = idx / W, w = idx % W;
    const float* G = gout + (int64_t)b * T * 3 * H * W;
    float accd = 0.f, accs[3] = {0.f, 0.f, 0.f};
    for (int kt = 0; kt < 3; ++kt) {
        const int to = t - kt + 1;               // output frame that read input frame t via tap kt
        if ((unsigned)to >= (unsigned)T) continue;
        for (int kh = 0; kh < 3; ++kh) {
            const int ho = h - kh + 1;
            if ((unsigned)ho >= (unsigned)H) continue;
            for (int kw = 0; kw < 3; ++kw) {
                const int wo = w - kw + 1;
                if ((unsigned)wo >= (unsigned)W) continue;
                const int tap = (kt * 3 + kh) * 3 + kw;
#pragma unroll
                for (int o = 0; o < 3; ++o) {
                    const float g = __ldg(G + (((int64_t)to * 3 + o) * H + ho) * W + wo);
                    accd = fmaf(sw[(o * 4 + 3) * 27 + tap], g, accd);
                    if (grad_static) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) accs[i] = fmaf(sw[(o * 4 + i) * 27 + tap], g, accs[i]);
                    }
                }
            }
        }
    }
    float* gd = grad_dynamic + (label[b] * dpc + dynamic_idx[b]) * (int64_t)T * H * W;
    atomicAdd(gd + ((int64_t)t * H + h) * W + w, accd);
    if (grad_static) {
        // the static image sits at every frame t' in [0,T): this thread's (t) contribution
        float* gs = grad_static + static_idx[b] * 3 * H * W;
#pragma unroll
        for (int i = 0; i < 3; ++i) atomicAdd(gs + ((int64_t)i * H + h) * W + w, accs[i]);
    }
}

// d_weight[o,i,tap] = sum_{b,t,h,w} g[b,t,o,h,w] X[b,i,t+kt-1,h+kh-1,w+kw-1]; d_bias[o] = sum g.
// Two kernels, both: one warp per image row, lanes along w, 81 register accumulators per lane, block
// reduction in shared memory, one atomicAdd per (block, weight).
//   dynamic channel (i = 3): acc[o][tap] over rows (b,t,h);  27 + 3 loads per 81 FMAs.
//   static channels (i < 3): X does not depend on t, so the sum over t collapses to three frame sums of g
//     (all t | t >= 1 | t <= T-2 for kt = 1 | 0 | 2): rows (b,h), one block column per o.
__global__ void __launch_bounds__(256) compose_bwd_weight_dyn_kernel(
        const float* __restrict__ gout, const float* __restrict__ dynamic_syn, const int64_t* __restrict__ label,
        const int64_t* __restrict__ dynamic_idx, float* __restrict__ grad_weight, float* __restrict__ grad_bias,
        int B, int T, int H, int W, int dpc, int rows_per_block) {
    __shared__ float red[8 * 84];
    __shared__ int dst[84];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    float acc[84];                               // [o][tap] then the 3 bias sums
#pragma unroll
    for (int k = 0; k < 84; ++k) acc[k] = 0.f;
    const int64_t total_rows = (int64_t)B * T * H;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(total_rows, r0 + rows_per_block);
    const int64_t HW = (int64_t)H * W;
    for (int64_t r = r0 + warp; r < r1; r += nwarp) {
        const int h = (int)(r % H); const int64_t q = r / H;
        const int t = (int)(q % T); const int b = (int)(q / T);
        const float* G = gout + (((int64_t)b * T + t) * 3) * HW + (int64_t)h * W;
        const float* D = dynamic_syn + (label[b] * dpc + dynamic_idx[b]) * (int64_t)T * HW;
        for (int w = lane; w < W; w += 32) {
            float g[3];
#pragma unroll
            for (int o = 0; o < 3; ++o) { g[o] = __ldg(G + o * HW + w); acc[81 + o] += g[o]; }
#pragma unroll
            for (int kt = 0; kt < 3; ++kt) {
                const int tt = t + kt - 1;
                if ((unsigned)tt >= (unsigned)T) continue;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int hh = h + kh - 1;
                    if ((unsigned)hh >= (unsigned)H) continue;
                    const float* row = D + (int64_t)tt * HW + (int64_t)hh * W;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int ww = w + kw - 1;
                        const float xv = ((unsigned)ww < (unsigned)W) ? __ldg(row + ww) : 0.f;
                        const int tap = (kt * 3 + kh) * 3 + kw;
#pragma unroll
                        for (int o = 0; o < 3; ++o) acc[o * 27 + tap] = fmaf(g[o], xv, acc[o * 27 + tap]);
                    }
                }
            }
        }
    }
    if (threadIdx.x < 81) dst[threadIdx.x] = ((threadIdx.x / 27) * 4 + 3) * 27 + threadIdx.x % 27;
    else if (threadIdx.x < 84) dst[threadIdx.x] = 324 + (threadIdx.x - 81);     // bias follows the 324 weights in `scratch`
    __syncthreads();
    // weights and bias live in different tensors: reduce into a 327-float view [weight | bias]
    constexpr int n = 84;
#pragma unroll
    for (int k = 0; k < n; ++k) {
        const float s = warp_sum(acc[k]);
        if (lane == 0) red[warp * n + k] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w) s += red[w * n + k];
        if (k < 81) atomicAdd(grad_weight + dst[k], s);
        else if (grad_bias) atomicAdd(grad_bias + (k - 81), s);
    }
}

__global__ void __launch_bounds__(256) compose_bwd_weight_static_kernel(
        const float* __restrict__ gout, const float* __restrict__ static_syn, const int64_t* __restrict__ static_idx,
        float* __restrict__ grad_weight, int B, int T, int H, int W, int rows_per_block) {
    __shared__ float red[8 * 81];
    const int o = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    float acc[81];                               // [i][kt][kh][kw]
#pragma unroll
    for (int k = 0; k < 81; ++k) acc[k] = 0.f;
    const int64_t total_rows = (int64_t)B * H;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(total_rows, r0 + rows_per_block);
    const int64_t HW = (int64_t)H * W;
    for (int64_t r = r0 + warp; r < r1; r += nwarp) {
        const int h = (int)(r % H); const int b = (int)(r / H);
        const float* G = gout + ((int64_t)b * T * 3 + o) * HW + (int64_t)h * W;
        const float* S = static_syn + static_idx[b] * 3 * HW;
        for (int w = lane; w < W; w += 32) {
            float gall = 0.f, gfirst = 0.f, glast = 0.f;
            for (int t = 0; t < T; ++t) {
                const float g = __ldg(G + (int64_t)t * 3 * HW + w);
                gall += g;
                if (t == 0) gfirst = g;
                if (t == T - 1) glast = g;
            }
            // tap kt reads frame t+kt-1: kt=0 needs t>=1, kt=2 needs t<=T-2
            const float gk[3] = {gall - gfirst, gall, gall - glast};
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int hh = h + kh - 1;
                    if ((unsigned)hh >= (unsigned)H) continue;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int ww = w + kw - 1;
                        const float xv = ((unsigned)ww < (unsigned)W) ? __ldg(S + i * HW + (int64_t)hh * W + ww) : 0.f;
#pragma unroll
                        for (int kt = 0; kt < 3; ++kt)
                            acc[i * 27 + (kt * 3 + kh) * 3 + kw] = fmaf(gk[kt], xv, acc[i * 27 + (kt * 3 + kh) * 3 + kw]);
                    }
                }
        }
    }
    constexpr int n = 81;
#pragma unroll
    for (int k = 0; k < n; ++k) {
        const float s = warp_sum(acc[k]);
        if (lane == 0) red[warp * n + k] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w) s += red[w * n + k];
        atomicAdd(grad_weight + (o * 4 + k / 27) * 27 + k % 27, s);
    }
}

// =============================================================== class mean + DM loss
// mean[c,d] = (1/n) sum_j emb[c,j,d].  A block owns (class, 256 columns): 64 threads x float4 along d, 4 row groups; group g adds
// rows g, g+4, g+8, ... in order and the four partial sums are combined in a fixed order ((g0 + g1) + (g2 + g3)) through shared
// memory — the same bits on every launch and for every world size, with 4x the loads in flight of a one-thread-per-column loop.
__global__ void __launch_bounds__(256) class_mean_kernel(const float* __restrict__ emb, float* __restrict__ mean, int n, int D) {
    __shared__ float4 part[4][64];
    const int c = blockIdx.y;
    const int lane4 = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int d = (blockIdx.x * 64 + lane4) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d < D) {
        const float* p = emb + (int64_t)c * n * D + d;
#pragma unroll 16                 // batch_real = 64: all 16 loads of a thread in flight (the kernel lasts a few microseconds)
        for (int j = grp; j < n; j += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p + (int64_t)j * D));
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    part[grp][lane4] = s;
    __syncthreads();
    if (grp == 0 && d < D) {
        const float4 a = part[0][lane4], b = part[1][lane4], e = part[2][lane4], f = part[3][lane4];
        const float inv = (float)n;
        float4 o;
        o.x = ((a.x + b.x) + (e.x + f.x)) / inv; o.y = ((a.y + b.y) + (e.y + f.y)) / inv;
        o.z = ((a.z + b.z) + (e.z + f.z)) / inv; o.w = ((a.w + b.w) + (e.w + f.w)) / inv;
        *reinterpret_cast<float4*>(mean + (int64_t)c * D + d) = o;
    }
}

// scalar fallback (D not a multiple of 4 or unaligned pointers): one thread per column, rows added in order
__global__ void class_mean_scalar_kernel(const float* __restrict__ emb, float* __restrict__ mean, int n, int D) {
    const int c = blockIdx.y;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const float* p = emb + (int64_t)c * n * D + d;
    float s = 0.f;
    for (int j = 0; j < n; ++j) s += __ldg(p + (int64_t)j * D);
    mean[(int64_t)c * D + d] = s / (float)n;
}

// Ragged class sums: rows offsets[c] .. offsets[c+1]-1 of emb (class-major order) are added IN ROW ORDER into sum[c,:] (an empty
// segment gives zeros).  The multi-GPU DM path shards the sampled real videos of every class over the ranks; each rank reduces
// the embeddings of ITS videos with this kernel and the (C, D) partial sums are all-reduced (distill.py).  float4 along d.
__global__ void __launch_bounds__(128) class_sum_ragged_kernel(const float* __restrict__ emb, const int32_t* __restrict__ offsets,
                                                               float* __restrict__ sum, int D) {
    const int c = blockIdx.y;
    const int d = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (d >= D) return;
    const int r0 = __ldg(offsets + c), r1 = __ldg(offsets + c + 1);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* p = emb + (int64_t)r0 * D + d;
#pragma unroll 8
    for (int r = r0; r < r1; ++r, p += D) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(sum + (int64_t)c * D + d) = s;
}

// per class: diff = mean_real - mean(emb_syn); loss += sum diff^2; grad_syn = -(2/ns) diff * scale.
// ONE block walks the classes in order: every class sum is a fixed-shape tree and the per-class sums are added in class order
// (the order of the reference's `loss += torch.sum(...)` loop, distill_s2d_ms.py:414-422), so the scalar is bitwise
// reproducible — no floating-point atomics.  C x D is ~1e5 elements: a ~10 us kernel either way.
__global__ void __launch_bounds__(1024) dm_loss_kernel(const float* __restrict__ mean_real, const float* __restrict__ emb_syn,
                                                       float* __restrict__ loss, float* __restrict__ grad_syn, int C, int ns, int D,
                                                       float scale) {
    __shared__ float red[32];
    float total = 0.f;
    for (int c = 0; c < C; ++c) {
        float part = 0.f;
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            float ms = 0.f;
            for (int j = 0; j < ns; ++j) ms += emb_syn[((int64_t)c * ns + j) * D + d];
            ms /= (float)ns;
            const float diff = mean_real[(int64_t)c * D + d] - ms;
            part += diff * diff;
            if (grad_syn) {
                const float g = -(2.0f / (float)ns) * diff * scale;
                for (int j = 0; j < ns; ++j) grad_syn[((int64_t)c * ns + j) * D + d] = g;
            }
        }
        part = block_sum(part, red);
        if (threadIdx.x == 0) total += part;
        __syncthreads();                                  // `red` is reused by the next class
    }
    if (threadIdx.x == 0) *loss += total;
}

// Parallel variant with caller-provided scratch: one block per class writes its sum to class_loss[c] (and the gradient); a second
// one-block launch adds the C class sums in class order — the same fixed summation order as dm_loss_kernel, C-way parallel.
__global__ void __launch_bounds__(256) dm_loss_class_kernel(const float* __restrict__ mean_real, const float* __restrict__ emb_syn,
                                                            float* __restrict__ class_loss, float* __restrict__ grad_syn, int ns, int D,
                                                            float scale) {
    __shared__ float red[32];
    const int c = blockIdx.x;
    float part = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float ms = 0.f;
        for (int j = 0; j < ns; ++j) ms += emb_syn[((int64_t)c * ns + j) * D + d];
        ms /= (float)ns;
        const float diff = mean_real[(int64_t)c * D + d] - ms;
        part += diff * diff;
        if (grad_syn) {
            const float g = -(2.0f / (float)ns) * diff * scale;
            for (int j = 0; j < ns; ++j) grad_syn[((int64_t)c * ns + j) * D + d] = g;
        }
    }
    part = block_sum(part, red);
    if (threadIdx.x == 0) class_loss[c] = part;
}

__global__ void dm_loss_finish_kernel(const float* __restrict__ class_loss, float* __restrict__ loss, int C) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float total = 0.f;
        for (int c = 0; c < C; ++c) total += class_loss[c];
        *loss += total;
    }
}

// =============================================================== optimiser / flat-parameter kernels
__global__ void sgd_momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                    int64_t n, float lr, float momentum, int first) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float4* p4 = reinterpret_cast<float4*>(p); const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* b4 = reinterpret_cast<float4*>(buf);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 gv = g4[i], pv = p4[i], bv;
        if (first) bv = gv;
        else { bv = b4[i]; bv.x = fmaf(momentum, bv.x, gv.x); bv.y = fmaf(momentum, bv.y, gv.y);
               bv.z = fmaf(momentum, bv.z, gv.z); bv.w = fmaf(momentum, bv.w, gv.w); }
        pv.x -= lr * bv.x; pv.y -= lr * bv.y; pv.z -= lr * bv.z; pv.w -= lr * bv.w;
        b4[i] = bv; p4[i] = pv;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float bv = first ? g[i] : fmaf(momentum, buf[i], g[i]);
        buf[i] = bv; p[i] -= lr * bv;
    }
}

__global__ void axpy_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                            int64_t n, float a) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = fmaf(a, x[i], y[i]);
}

__global__ void sqdist_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n) {
    __shared__ float red[32];
    float s = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float d = a[i] - b[i]; s = fmaf(d, d, s);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(out, s);
}

static inline unsigned grid_for(int64_t n, int threads) {
    int64_t b = ceil_div(n, threads);
    const int64_t cap = 148 * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

static int pool_args(int T, int H, int W, int pt, int ph, int pw) {
    VD_REQUIRE(pt >= 1 && pt <= 2 && ph >= 1 && ph <= 2 && pw >= 1 && pw <= 2, "pool window must be 1 or 2 per axis");
    VD_REQUIRE(T >= pt && H >= ph && W >= pw, "pool: input smaller than window");
    return 0;
}

}  // namespace vd

using namespace vd;

extern "C" int vd_relu_maxpool_fwd_f32(const float* x, float* y, uint8_t* code, int64_t NC, int T, int H, int W,
                                       int pt, int ph, int pw, void* stream) {
    if (int e = pool_args(T, H, W, pt, ph, pw)) return e;
    VD_REQUIRE(x && y, "relu_maxpool_fwd: NULL pointer");
    const int To = T / pt, Ho = H / ph, Wo = W / pw;
    const int64_t total = NC * To * Ho * Wo;
    if (total == 0) return 0;
    relu_maxpool_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, code, total, T, H, W, To, Ho, Wo, pt, ph, pw);
    return check_launch("relu_maxpool_fwd_f32");
}

extern "C" int vd_route_scatter_f32(const float* gy, const uint8_t* code, float* gx, int64_t NC, int T, int H, int W,
                                    int pt, int ph, int pw, void* stream) {
    if (int e = pool_args(T, H, W, pt, ph, pw)) return e;
    VD_REQUIRE(gy && code && gx, "route_scatter: NULL pointer");
    const int To = T / pt, Ho = H / ph, Wo = W / pw;
    const int64_t total = NC * To * Ho * Wo;
    if (total == 0) return 0;
    if (To * pt != T || Ho * ph != H || Wo * pw != W) {
        cudaError_t e = cudaMemsetAsync(gx, 0, sizeof(float) * NC * T * H * W, (cudaStream_t)stream);
        if (e != cudaSuccess) { set_error("route_scatter memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    route_scatter_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(gy, code, gx, total, T, H, W, To, Ho, Wo, pt, ph, pw);
    return check_launch("route_scatter_f32");
}

extern "C" int vd_route_gather_f32(const float* x, const uint8_t* code, float* y, int64_t NC, int T, int H, int W,
                                   int pt, int ph, int pw, void* stream) {
    if (int e = pool_args(T, H, W, pt, ph, pw)) return e;
    VD_REQUIRE(x && code && y, "route_gather: NULL pointer");
    const int To = T / pt, Ho = H / ph, Wo = W / pw;
    const int64_t total = NC * To * Ho * Wo;
    if (total == 0) return 0;
    route_gather_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, code, y, total, T, H, W, To, Ho, Wo, pt, ph, pw);
    return check_launch("route_gather_f32");
}

extern "C" int vd_inorm_relu_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                     float* rstd, int N, int C, int64_t S, void* stream) {
    VD_REQUIRE(x && gamma && beta && y && mean && rstd && N > 0 && C > 0 && S > 0, "inorm_relu_fwd: bad argument");
    inorm_relu_fwd_kernel<<<N * C, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, mean, rstd, C, S);
    return check_launch("inorm_relu_fwd_f32");
}

extern "C" int vd_inorm_relu_bwd_f32(const float* x, const float* y, const float* gy, const float* gamma,
                                     const float* mean, const float* rstd, float* gx, float* ggamma, float* gbeta,
                                     int N, int C, int64_t S, void* stream) {
    VD_REQUIRE(x && y && gy && gamma && mean && rstd && gx && ggamma && gbeta && N > 0 && C > 0 && S > 0, "inorm_relu_bwd: bad argument");
    inorm_relu_bwd_kernel<<<N * C, 256, 0, (cudaStream_t)stream>>>(x, y, gy, gamma, mean, rstd, gx, ggamma, gbeta, C, S);
    return check_launch("inorm_relu_bwd_f32");
}

extern "C" int vd_inorm_relu_avgpool_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                             float* rstd, int N, int C, int T, int H, int W, void* stream) {
    VD_REQUIRE(x && gamma && beta && y && N > 0 && C > 0 && T >= 2 && H >= 2 && W >= 2, "inorm_relu_avgpool_fwd: bad argument");
    VD_REQUIRE(W % 2 == 0 && ((int64_t)T * H * W) % 4 == 0 && ((uintptr_t)x & 15) == 0,
               "inorm_relu_avgpool_fwd: W must be even, T*H*W a multiple of 4 and x 16-byte aligned");
    VD_REQUIRE((mean == nullptr) == (rstd == nullptr), "inorm_relu_avgpool_fwd: mean and rstd go together");
    inorm_relu_avgpool_fwd_kernel<<<N * C, 512, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, mean, rstd, C, T, H, W);
    return check_launch("inorm_relu_avgpool_fwd_f32");
}

extern "C" int vd_avgpool2_fwd_f32(const float* x, float* y, int64_t NC, int T, int H, int W, void* stream) {
    VD_REQUIRE(x && y && T >= 2 && H >= 2 && W >= 2, "avgpool2_fwd: bad argument");
    const int To = T / 2, Ho = H / 2, Wo = W / 2;
    const int64_t total = NC * To * Ho * Wo;
    if (total == 0) return 0;
    avgpool2_fwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, total, T, H, W, To, Ho, Wo);
    return check_launch("avgpool2_fwd_f32");
}

extern "C" int vd_avgpool2_bwd_f32(const float* gy, float* gx, int64_t NC, int T, int H, int W, void* stream) {
    VD_REQUIRE(gy && gx && T >= 2 && H >= 2 && W >= 2, "avgpool2_bwd: bad argument");
    const int64_t total = NC * T * H * W;
    if (total == 0) return 0;
    avgpool2_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(gy, gx, total, T, H, W, T / 2, H / 2, W / 2);
    return check_launch("avgpool2_bwd_f32");
}

namespace vd {
// shared-memory tiled versions (compose_tiled.cu); return 1 when the geometry is not covered
int compose_fwd_tiled(const float* static_syn, const float* dynamic_syn, const int64_t* static_idx, const int64_t* label,
                      const int64_t* dynamic_idx, const float* weight, const float* bias, float* out, int B, int T, int H,
                      int W, int dpc, cudaStream_t stream);
// TMA-fed versions (compose_tma.cu); return 1 when the geometry is not covered or tensor maps are unavailable
int compose_fwd_tma(const float* static_syn, const float* dynamic_syn, const int64_t* static_idx, const int64_t* label,
                    const int64_t* dynamic_idx, const float* weight, const float* bias, float* out, int B, int T, int H,
                    int W, int dpc, int64_t n_static, int64_t n_dynamic, cudaStream_t stream);
int compose_bwd_tma(const float* gout, const float* static_syn, const float* dynamic_syn, const int64_t* static_idx,
                    const int64_t* label, const int64_t* dynamic_idx, const float* weight, float* grad_dynamic, float* grad_weight,
                    float* grad_bias, float* scratch, int64_t scratch_floats, int unique_rows, int B, int T, int H, int W, int dpc,
                    int64_t n_static, int64_t n_dynamic, cudaStream_t stream);
int compose_bwd_data_tiled(const float* gout, const int64_t* label, const int64_t* dynamic_idx, const float* weight,
                           float* grad_dynamic, int B, int T, int H, int W, int dpc, cudaStream_t stream);
int compose_bwd_wdyn_tiled(const float* gout, const float* dynamic_syn, const int64_t* label, const int64_t* dynamic_idx,
                           float* grad_weight, float* grad_bias, int B, int T, int H, int W, int dpc, cudaStream_t stream);
int compose_bwd_fused(const float* gout, const float* static_syn, const float* dynamic_syn, const int64_t* static_idx,
                      const int64_t* label, const int64_t* dynamic_idx, const float* weight, float* grad_dynamic, float* grad_weight,
                      float* grad_bias, float* scratch, int64_t scratch_floats, int unique_rows, int B, int T, int H, int W, int dpc,
                      cudaStream_t stream);
}  // namespace vd

static int compose_fwd_impl(const float* static_syn, const float* dynamic_syn, const int64_t* static_idx,
                            const int64_t* label, const int64_t* dynamic_idx, const float* weight,
                            const float* bias, float* out, int B, int T, int H, int W, int dpc, int64_t n_static, int64_t n_dynamic,
                            void* stream);

extern "C" int vd_compose_fwd_f32(const float* static_syn, const float* dynamic_syn, const int64_t* static_idx,
                                  const int64_t* label, const int64_t* dynamic_idx, const float* weight,
                                  const float* bias, float* out, int B, int T, int H, int W, int dpc, void* stream) {
    return compose_fwd_impl(static_syn, dynamic_syn, static_idx, label, dynamic_idx, weight, bias, out, B, T, H, W, dpc, 0, 0, stream);
}

// The same with the extents of the memories (n_static images of (3,H,W), n_dynamic memories of (T,H,W)): enables the tensor-map
// (TMA) kernels, whose maps must describe the true allocations.
extern "C" int vd_compose_fwd_ex_f32(const float* static_syn, const float* dynamic_syn, const int64_t* static_idx,
                                     const int64_t* label, const int64_t* dynamic_idx, const float* weight,
                                     const float* bias, float* out, int B, int T, int H, int W, int dpc,
                                     int64_t n_static, int64_t n_dynamic, void* stream) {
    VD_REQUIRE(n_static > 0 && n_dynamic > 0, "compose_fwd_ex: the memory extents must be positive");
    return compose_fwd_impl(static_syn, dynamic_syn, static_idx, label, dynamic_idx, weight, bias, out, B, T, H, W, dpc, n_static, n_dynamic, stream);
}

static int compose_fwd_impl(const float* static_syn, const float* dynamic_syn, const int64_t* static_idx,
                            const int64_t* label, const int64_t* dynamic_idx, const float* weight,
                            const float* bias, float* out, int B, int T, int H, int W, int dpc, int64_t n_static, int64_t n_dynamic,
                            void* stream) {
    VD_REQUIRE(static_syn && dynamic_syn && static_idx && label && dynamic_idx && weight && bias && out, "compose_fwd: NULL pointer");
    VD_REQUIRE(B >= 0 && T > 0 && H > 0 && W > 0 && dpc > 0 && T <= 65535 && B <= 65535, "compose_fwd: bad extent");
    if (B == 0) return 0;
    {
        int rc = 1;
        if (n_static > 0 && n_dynamic > 0)
            rc = compose_fwd_tma(static_syn, dynamic_syn, static_idx, label, dynamic_idx, weight, bias, out, B, T, H, W, dpc,
                                 n_static, n_dynamic, (cudaStream_t)stream);
        if (rc != 1) return rc;                 // 1 = not covered: the cp.async tiled kernel, then the generic one
        rc = compose_fwd_tiled(static_syn, dynamic_syn, static_idx, label, dynamic_idx, weight, bias, out, B, T, H, W, dpc,
                               (cudaStream_t)stream);
        if (rc != 1) return rc;
    }
    const int W4 = (W + 3) / 4;
    dim3 grid((unsigned)ceil_div((int64_t)H * W4, 256), T, B);
    compose_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(static_syn, dynamic_syn, static_idx, label, dynamic_idx,
                                                              weight, bias, out, T, H, W, dpc);
    return check_launch("compose_fwd_f32");
}

extern "C" int vd_compose_bwd_f32(const float* gout, const float* static_syn, const float* dynamic_syn,
                                  const int64_t* static_idx, const int64_t* label, const int64_t* dynamic_idx,
                                  const float* weight, float* grad_dynamic, float* grad_weight, float* grad_bias,
                                  float* grad_static, int B, int T, int H, int W, int dpc, void* stream) {
    VD_REQUIRE(gout && static_syn && dynamic_syn && static_idx && label && dynamic_idx && weight, "compose_bwd: NULL pointer");
    VD_REQUIRE(B >= 0 && T > 0 && H > 0 && W > 0 && dpc > 0 && T <= 65535 && B <= 65535, "compose_bwd: bad extent");
    if (B == 0) return 0;
    int rc_data = 1;
    if (grad_dynamic && !grad_static) {
        rc_data = compose_bwd_data_tiled(gout, label, dynamic_idx, weight, grad_dynamic, B, T, H, W, dpc, (cudaStream_t)stream);
        if (rc_data != 0 && rc_data != 1) return rc_data;
    }
    if (grad_dynamic && rc_data == 1) {
        dim3 grid((unsigned)ceil_div((int64_t)H * W, 256), T, B);
        compose_bwd_data_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gout, static_idx, label, dynamic_idx, weight,
                                                                       grad_dynamic, grad_static, T, H, W, dpc);
        if (int e = check_launch("compose_bwd_data_f32")) return e;
    }
    if (grad_weight) {
        cudaStream_t s = (cudaStream_t)stream;
        const int rc_w = compose_bwd_wdyn_tiled(gout, dynamic_syn, label, dynamic_idx, grad_weight, grad_bias, B, T, H, W, dpc, s);
        if (rc_w != 0 && rc_w != 1) return rc_w;
        if (rc_w == 1) {
            const int64_t rows = (int64_t)B * T * H;
            int rpb = (int)ceil_div(rows, 148 * 4);
            if (rpb < 8) rpb = 8;
            compose_bwd_weight_dyn_kernel<<<(unsigned)ceil_div(rows, rpb), 256, 0, s>>>(gout, dynamic_syn, label, dynamic_idx, grad_weight,
                                                                                      grad_bias, B, T, H, W, dpc, rpb);
            if (int e = check_launch("compose_bwd_weight_dyn_f32")) return e;
        }
        {
            const int64_t rows = (int64_t)B * H;
            int rpb = (int)ceil_div(rows, 148);
            if (rpb < 8) rpb = 8;
            dim3 grid((unsigned)ceil_div(rows, rpb), 3, 1);
            compose_bwd_weight_static_kernel<<<grid, 256, 0, s>>>(gout, static_syn, static_idx, grad_weight, B, T, H, W, rpb);
            if (int e = check_launch("compose_bwd_weight_static_f32")) return e;
        }
    }
    return 0;
}

// One-pass deterministic backward of the composer (compose_tiled.cu: compose_bwd_fused_kernel): grad_dynamic rows written (or
// accumulated with atomics when unique_rows == 0), grad_weight / grad_bias += block sums added in a fixed order.  Falls back to
// vd_compose_bwd_f32 when the geometry is not covered by the tiled kernels.
static int compose_bwd_fused_impl(const float* gout, const float* static_syn, const float* dynamic_syn,
                                  const int64_t* static_idx, const int64_t* label, const int64_t* dynamic_idx,
                                  const float* weight, float* grad_dynamic, float* grad_weight, float* grad_bias,
                                  float* scratch, int64_t scratch_floats, int unique_rows,
                                  int B, int T, int H, int W, int dpc, int64_t n_static, int64_t n_dynamic, void* stream);

extern "C" int vd_compose_bwd_fused_f32(const float* gout, const float* static_syn, const float* dynamic_syn,
                                        const int64_t* static_idx, const int64_t* label, const int64_t* dynamic_idx,
                                        const float* weight, float* grad_dynamic, float* grad_weight, float* grad_bias,
                                        float* scratch, int64_t scratch_floats, int unique_rows,
                                        int B, int T, int H, int W, int dpc, void* stream) {
    return compose_bwd_fused_impl(gout, static_syn, dynamic_syn, static_idx, label, dynamic_idx, weight, grad_dynamic, grad_weight,
                                  grad_bias, scratch, scratch_floats, unique_rows, B, T, H, W, dpc, 0, 0, stream);
}

// with the extents of the memories (see vd_compose_fwd_ex_f32): enables the tensor-map (TMA) kernel
extern "C" int vd_compose_bwd_fused_ex_f32(const float* gout, const float* static_syn, const float* dynamic_syn,
                                           const int64_t* static_idx, const int64_t* label, const int64_t* dynamic_idx,
                                           const float* weight, float* grad_dynamic, float* grad_weight, float* grad_bias,
                                           float* scratch, int64_t scratch_floats, int unique_rows,
                                           int B, int T, int H, int W, int dpc, int64_t n_static, int64_t n_dynamic, void* stream) {
    VD_REQUIRE(n_static > 0 && n_dynamic > 0, "compose_bwd_fused_ex: the memory extents must be positive");
    return compose_bwd_fused_impl(gout, static_syn, dynamic_syn, static_idx, label, dynamic_idx, weight, grad_dynamic, grad_weight,
                                  grad_bias, scratch, scratch_floats, unique_rows, B, T, H, W, dpc, n_static, n_dynamic, stream);
}

static int compose_bwd_fused_impl(const float* gout, const float* static_syn, const float* dynamic_syn,
                                  const int64_t* static_idx, const int64_t* label, const int64_t* dynamic_idx,
                                  const float* weight, float* grad_dynamic, float* grad_weight, float* grad_bias,
                                  float* scratch, int64_t scratch_floats, int unique_rows,
                                  int B, int T, int H, int W, int dpc, int64_t n_static, int64_t n_dynamic, void* stream) {
    VD_REQUIRE(gout && static_syn && dynamic_syn && static_idx && label && dynamic_idx && weight && grad_dynamic && grad_weight && scratch,
               "compose_bwd_fused: NULL pointer");
    VD_REQUIRE(B >= 0 && T > 0 && H > 0 && W > 0 && dpc > 0 && T <= 65535 && B <= 65535, "compose_bwd_fused: bad extent");
    if (B == 0) return 0;
    int rc = 1;
    if (n_static > 0 && n_dynamic > 0)
        rc = compose_bwd_tma(gout, static_syn, dynamic_syn, static_idx, label, dynamic_idx, weight, grad_dynamic, grad_weight,
                             grad_bias, scratch, scratch_floats, unique_rows, B, T, H, W, dpc, n_static, n_dynamic, (cudaStream_t)stream);
    if (rc != 1) return rc;
    rc = compose_bwd_fused(gout, static_syn, dynamic_syn, static_idx, label, dynamic_idx, weight, grad_dynamic, grad_weight,
                           grad_bias, scratch, scratch_floats, unique_rows, B, T, H, W, dpc, (cudaStream_t)stream);
    if (rc != 1) return rc;
    return vd_compose_bwd_f32(gout, static_syn, dynamic_syn, static_idx, label, dynamic_idx, weight, grad_dynamic, grad_weight,
                              grad_bias, nullptr, B, T, H, W, dpc, stream);
}

extern "C" int vd_class_mean_f32(const float* emb, float* mean, int C, int n, int D, void* stream) {
    VD_REQUIRE(emb && mean && C >= 0 && n > 0 && D > 0 && C <= 65535, "class_mean: bad argument");
    if (C == 0) return 0;
    if (D % 4 == 0 && ((uintptr_t)emb & 15) == 0 && ((uintptr_t)mean & 15) == 0) {
        dim3 grid((unsigned)ceil_div(D, 256), C, 1);
        class_mean_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(emb, mean, n, D);
    } else {
        dim3 grid((unsigned)ceil_div(D, 128), C, 1);
        class_mean_scalar_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(emb, mean, n, D);
    }
    return check_launch("class_mean_f32");
}

extern "C" int vd_class_sum_ragged_f32(const float* emb, const int32_t* offsets, float* sum, int C, int D, void* stream) {
    VD_REQUIRE(emb && offsets && sum && C >= 0 && D > 0 && C <= 65535, "class_sum_ragged: bad argument");
    VD_REQUIRE(D % 4 == 0 && ((uintptr_t)emb & 15) == 0 && ((uintptr_t)sum & 15) == 0, "class_sum_ragged: D must be a multiple of 4 and the pointers 16-byte aligned");
    if (C == 0) return 0;
    dim3 grid((unsigned)ceil_div(D, 512), C, 1);
    class_sum_ragged_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(emb, offsets, sum, D);
    return check_launch("class_sum_ragged_f32");
}

extern "C" int vd_dm_loss_f32(const float* mean_real, const float* emb_syn, float* loss, float* grad_syn,
                              int C, int ns, int D, float loss_scale, void* stream) {
    VD_REQUIRE(mean_real && emb_syn && loss && C >= 0 && ns > 0 && D > 0, "dm_loss: bad argument");
    if (C == 0) return 0;
    dm_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mean_real, emb_syn, loss, grad_syn, C, ns, D, loss_scale);
    return check_launch("dm_loss_f32");
}

extern "C" int vd_dm_loss_ex_f32(const float* mean_real, const float* emb_syn, float* loss, float* grad_syn, float* class_loss,
                                 int C, int ns, int D, float loss_scale, void* stream) {
    VD_REQUIRE(mean_real && emb_syn && loss && class_loss && C >= 0 && ns > 0 && D > 0, "dm_loss_ex: bad argument");
    if (C == 0) return 0;
    dm_loss_class_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(mean_real, emb_syn, class_loss, grad_syn, ns, D, loss_scale);
    if (int e = check_launch("dm_loss_class_f32")) return e;
    dm_loss_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(class_loss, loss, C);
    return check_launch("dm_loss_finish_f32");
}

extern "C" int vd_sgd_momentum_f32(float* p, const float* g, float* buf, int64_t n, float lr, float momentum,
                                   int first_step, void* stream) {
    VD_REQUIRE(p && g && buf && n >= 0, "sgd_momentum: bad argument");
    if (n == 0) return 0;
    VD_REQUIRE(((uintptr_t)p & 15) == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)buf & 15) == 0, "sgd_momentum: pointers must be 16-byte aligned");
    sgd_momentum_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(p, g, buf, n, lr, momentum, first_step);
    return check_launch("sgd_momentum_f32");
}

extern "C" int vd_axpy_f32(const float* x, const float* y_in, float* y_out, int64_t n, float a, void* stream) {
    VD_REQUIRE(x && y_in && y_out && n >= 0, "axpy: bad argument");
    if (n == 0) return 0;
    axpy_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y_in, y_out, n, a);
    return check_launch("axpy_f32");
}

extern "C" int vd_sqdist_f32(const float* a, const float* b, float* out, int64_t n, void* stream) {
    VD_REQUIRE(a && b && out && n >= 0, "sqdist: bad argument");
    if (n == 0) return 0;
    sqdist_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
    return check_launch("sqdist_f32");
}
