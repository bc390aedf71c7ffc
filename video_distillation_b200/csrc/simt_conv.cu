// Exact fp32 Conv3d trio on CUDA cores: fprop / dgrad / wgrad as tiled implicit GEMMs.
//
// This is the fp32 "parity mode" of the convolution (single-pass bf16 cannot meet
// the 1e-3 gate on embeddings) and the closure the MTT double backward is built from.  The
// throughput path for the ConvNet3D feature convolutions is the tcgen05 kernel in tc_conv.cu.
//
// One kernel template: C[M x N] = A[M x K] * B[K x N] with A/B gathered on the fly by a problem
// functor; 64x64x16 tiles, 256 threads, 4x4 outputs per thread, optional split-K (atomics).
#include "common.cuh"

namespace vd {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct Geo {
    int N, Cin, T, H, W, Cout, To, Ho, Wo, kt, kh, kw, st, sh, sw, pt, ph, pw;
    int KV;            // kt*kh*kw
    int64_t Si, So;    // T*H*W, To*Ho*Wo
};

static Geo make_geo(const vd_conv_geom* g) {
    Geo q;
    q.N = g->N; q.Cin = g->Cin; q.T = g->T; q.H = g->H; q.W = g->W;
    q.Cout = g->Cout; q.To = g->To; q.Ho = g->Ho; q.Wo = g->Wo;
    q.kt = g->kt; q.kh = g->kh; q.kw = g->kw; q.st = g->st; q.sh = g->sh; q.sw = g->sw;
    q.pt = g->pt; q.ph = g->ph; q.pw = g->pw;
    q.KV = q.kt * q.kh * q.kw;
    q.Si = (int64_t)q.T * q.H * q.W;
    q.So = (int64_t)q.To * q.Ho * q.Wo;
    return q;
}

// ---------------------------------------------------------------- fprop: rows = output pixels
struct FpropProblem {
    Geo g; const float* x; const float* w; const float* bias; float* y;
    int64_t M, N, K;
    struct Row { const float* xb; int t0, h0, w0; bool ok; };
    struct Col { const float* wb; bool ok; };
    __device__ Row row(int64_t m) const {
        Row r; r.ok = m < M;
        if (!r.ok) { r.xb = x; r.t0 = r.h0 = r.w0 = 0; return r; }
        int wo = (int)(m % g.Wo); int64_t q = m / g.Wo;
        int ho = (int)(q % g.Ho); q /= g.Ho;
        int to = (int)(q % g.To); int n = (int)(q / g.To);
        r.xb = x + (int64_t)n * g.Cin * g.Si;
        r.t0 = to * g.st - g.pt; r.h0 = ho * g.sh - g.ph; r.w0 = wo * g.sw - g.pw;
        return r;
    }
    __device__ Col col(int64_t n) const { Col c; c.ok = n < N; c.wb = w + (c.ok ? n : 0) * K; return c; }
    __device__ float loadA(const Row& r, int64_t k) const {
        if (!r.ok || k >= K) return 0.f;
        int c = (int)(k % g.kw); int64_t q = k / g.kw;
        int b = (int)(q % g.kh); q /= g.kh;
        int a = (int)(q % g.kt); int ci = (int)(q / g.kt);
        int t = r.t0 + a, h = r.h0 + b, ww = r.w0 + c;
        if ((unsigned)t >= (unsigned)g.T || (unsigned)h >= (unsigned)g.H || (unsigned)ww >= (unsigned)g.W) return 0.f;
        return __ldg(r.xb + (int64_t)ci * g.Si + ((int64_t)t * g.H + h) * g.W + ww);
    }
    __device__ float loadB(int64_t k, const Col& c) const { return (c.ok && k < K) ? __ldg(c.wb + k) : 0.f; }
    __device__ void store(int64_t m, int64_t n, float v, bool) const {
        if (m >= M || n >= N) return;
        int64_t pix = m % g.So; int64_t nb = m / g.So;
        y[(nb * g.Cout + n) * g.So + pix] = v + (bias ? __ldg(bias + n) : 0.f);
    }
};

// ---------------------------------------------------------------- dgrad: rows = input pixels
struct DgradProblem {
    Geo g; const float* gy; const float* w; float* gx;
    int64_t M, N, K;
    struct Row { const float* gb; int t, h, w; bool ok; };
    struct Col { int ci; bool ok; };
    __device__ Row row(int64_t m) const {
        Row r; r.ok = m < M;
        if (!r.ok) { r.gb = gy; r.t = r.h = r.w = 0; return r; }
        int ww = (int)(m % g.W); int64_t q = m / g.W;
        int h = (int)(q % g.H); q /= g.H;
        int t = (int)(q % g.T); int n = (int)(q / g.T);
        r.gb = gy + (int64_t)n * g.Cout * g.So;
        r.t = t + g.pt; r.h = h + g.ph; r.w = ww + g.pw;
        return r;
    }
    __device__ Col col(int64_t n) const { Col c; c.ok = n < N; c.ci = (int)n; return c; }
    __device__ float loadA(const Row& r, int64_t k) const {
        if (!r.ok || k >= K) return 0.f;
        int c = (int)(k % g.kw); int64_t q = k / g.kw;
        int b = (int)(q % g.kh); q /= g.kh;
        int a = (int)(q % g.kt); int co = (int)(q / g.kt);
        int tt = r.t - a, hh = r.h - b, wv = r.w - c;
        if (tt < 0 || hh < 0 || wv < 0) return 0.f;
        if (tt % g.st || hh % g.sh || wv % g.sw) return 0.f;
        int to = tt / g.st, ho = hh / g.sh, wo = wv / g.sw;
        if (to >= g.To || ho >= g.Ho || wo >= g.Wo) return 0.f;
        return __ldg(r.gb + (int64_t)co * g.So + ((int64_t)to * g.Ho + ho) * g.Wo + wo);
    }
    __device__ float loadB(int64_t k, const Col& c) const {
        if (!c.ok || k >= K) return 0.f;
        int64_t co = k / g.KV; int tap = (int)(k % g.KV);
        return __ldg(w + (co * g.Cin + c.ci) * g.KV + tap);
    }
    __device__ void store(int64_t m, int64_t n, float v, bool) const {
        if (m >= M || n >= N) return;
        int64_t pix = m % g.Si; int64_t nb = m / g.Si;
        gx[(nb * g.Cin + n) * g.Si + pix] = v;
    }
};

// ---------------------------------------------------------------- wgrad: rows = cout, cols = (ci,tap)
struct WgradProblem {
    Geo g; const float* x; const float* gy; float* gw;
    int64_t M, N, K;
    struct Row { int co; bool ok; };
    struct Col { int ci, a, b, c; bool ok; };
    __device__ Row row(int64_t m) const { Row r; r.ok = m < M; r.co = (int)m; return r; }
    __device__ Col col(int64_t n) const {
        Col c; c.ok = n < N;
        int64_t q = c.ok ? n : 0;
        c.c = (int)(q % g.kw); q /= g.kw;
        c.b = (int)(q % g.kh); q /= g.kh;
        c.a = (int)(q % g.kt); c.ci = (int)(q / g.kt);
        return c;
    }
    __device__ float loadA(const Row& r, int64_t k) const {
        if (!r.ok || k >= K) return 0.f;
        int64_t pix = k % g.So; int64_t n = k / g.So;
        return __ldg(gy + (n * g.Cout + r.co) * g.So + pix);
    }
    __device__ float loadB(int64_t k, const Col& c) const {
        if (!c.ok || k >= K) return 0.f;
        int wo = (int)(k % g.Wo); int64_t q = k / g.Wo;
        int ho = (int)(q % g.Ho); q /= g.Ho;
        int to = (int)(q % g.To); int64_t n = q / g.To;
        int t = to * g.st - g.pt + c.a, h = ho * g.sh - g.ph + c.b, ww = wo * g.sw - g.pw + c.c;
        if ((unsigned)t >= (unsigned)g.T || (unsigned)h >= (unsigned)g.H || (unsigned)ww >= (unsigned)g.W) return 0.f;
        return __ldg(x + (n * g.Cin + c.ci) * g.Si + ((int64_t)t * g.H + h) * g.W + ww);
    }
    __device__ void store(int64_t m, int64_t n, float v, bool atomic) const {
        if (m >= M || n >= N) return;
        if (atomic) atomicAdd(gw + m * N + n, v); else gw[m * N + n] += v;
    }
};

template <class P>
__global__ void __launch_bounds__(NT) igemm_kernel(P p, int64_t k_per_split) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
    const int64_t kend = min(p.K, kbeg + k_per_split);

    const int am = tid % BM, ak = tid / BM;      // A loads: (am, ak + 4j)
    const int bk = tid % BK, bn = tid / BK;      // B loads: (bk, bn + 16j)
    const typename P::Row arow = p.row(m0 + am);
    typename P::Col bcol[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bcol[j] = p.col(n0 + bn + 16 * j);

    const int ty = tid / 16, tx = tid % 16;      // outputs: rows ty*4.., cols tx*4..
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        float ra[4], rb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t ka = k0 + ak + 4 * j;
            ra[j] = (ka < kend) ? p.loadA(arow, ka) : 0.f;
            int64_t kb = k0 + bk;
            rb[j] = (kb < kend) ? p.loadB(kb, bcol[j]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[ak + 4 * j][am] = ra[j];
            Bs[bk][bn + 16 * j] = rb[j];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    const bool atomic = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) p.store(m0 + ty * 4 + i, n0 + tx * 4 + j, acc[i][j], atomic);
}

__global__ void bias_grad_kernel(const float* __restrict__ gy, float* __restrict__ gb, int N, int Cout, int64_t So) {
    __shared__ float red[32];
    const int co = blockIdx.x;
    float s = 0.f;
    for (int n = 0; n < N; ++n) {
        const float* p = gy + ((int64_t)n * Cout + co) * So;
        for (int64_t i = threadIdx.x; i < So; i += blockDim.x) s += p[i];
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) gb[co] += s;
}

static int check_geom(const vd_conv_geom* g) {
    VD_REQUIRE(g != nullptr, "conv geometry is NULL");
    VD_REQUIRE(g->N > 0 && g->Cin > 0 && g->Cout > 0 && g->T > 0 && g->H > 0 && g->W > 0, "conv: empty extent");
    VD_REQUIRE(g->st > 0 && g->sh > 0 && g->sw > 0 && g->kt > 0 && g->kh > 0 && g->kw > 0, "conv: bad filter/stride");
    VD_REQUIRE(g->To == (g->T + 2 * g->pt - g->kt) / g->st + 1 && g->Ho == (g->H + 2 * g->ph - g->kh) / g->sh + 1 &&
               g->Wo == (g->W + 2 * g->pw - g->kw) / g->sw + 1, "conv: output extent inconsistent with input/stride/pad");
    return 0;
}

}  // namespace vd

using namespace vd;

extern "C" int vd_conv3d_fprop_f32(const float* x, const float* w, const float* bias, float* y,
                                   const vd_conv_geom* g, void* stream) {
    if (int e = check_geom(g)) return e;
    VD_REQUIRE(x && w && y, "conv3d_fprop: NULL pointer");
    FpropProblem p; p.g = make_geo(g); p.x = x; p.w = w; p.bias = bias; p.y = y;
    p.M = (int64_t)g->N * p.g.So; p.N = g->Cout; p.K = (int64_t)g->Cin * p.g.KV;
    dim3 grid((unsigned)ceil_div(p.M, BM), (unsigned)ceil_div(p.N, BN), 1);
    igemm_kernel<FpropProblem><<<grid, NT, 0, (cudaStream_t)stream>>>(p, p.K);
    return check_launch("conv3d_fprop_f32");
}

extern "C" int vd_conv3d_dgrad_f32(const float* gy, const float* w, float* gx,
                                   const vd_conv_geom* g, void* stream) {
    if (int e = check_geom(g)) return e;
    VD_REQUIRE(gy && w && gx, "conv3d_dgrad: NULL pointer");
    DgradProblem p; p.g = make_geo(g); p.gy = gy; p.w = w; p.gx = gx;
    p.M = (int64_t)g->N * p.g.Si; p.N = g->Cin; p.K = (int64_t)g->Cout * p.g.KV;
    dim3 grid((unsigned)ceil_div(p.M, BM), (unsigned)ceil_div(p.N, BN), 1);
    igemm_kernel<DgradProblem><<<grid, NT, 0, (cudaStream_t)stream>>>(p, p.K);
    return check_launch("conv3d_dgrad_f32");
}

extern "C" int vd_conv3d_wgrad_f32(const float* x, const float* gy, float* gw, float* gb,
                                   const vd_conv_geom* g, void* stream) {
    if (int e = check_geom(g)) return e;
    VD_REQUIRE(x && gy && gw, "conv3d_wgrad: NULL pointer");
    WgradProblem p; p.g = make_geo(g); p.x = x; p.gy = gy; p.gw = gw;
    p.M = g->Cout; p.N = (int64_t)g->Cin * p.g.KV; p.K = (int64_t)g->N * p.g.So;
    const int64_t tiles = ceil_div(p.M, BM) * ceil_div(p.N, BN);
    int64_t splits = 1;
    // split the pixel reduction so that the grid fills 148 SMs a few times over
    while (tiles * splits < 148 * 4 && p.K / (splits * 2) >= 4 * BK && splits < 4096) splits *= 2;
    int64_t kps = ceil_div(ceil_div(p.K, splits), BK) * BK;
    splits = ceil_div(p.K, kps);
    dim3 grid((unsigned)ceil_div(p.M, BM), (unsigned)ceil_div(p.N, BN), (unsigned)splits);
    igemm_kernel<WgradProblem><<<grid, NT, 0, (cudaStream_t)stream>>>(p, kps);
    if (int e = check_launch("conv3d_wgrad_f32")) return e;
    if (gb) {
        bias_grad_kernel<<<g->Cout, 256, 0, (cudaStream_t)stream>>>(gy, gb, g->N, g->Cout, p.g.So);
        return check_launch("conv3d_bias_grad_f32");
    }
    return 0;
}
