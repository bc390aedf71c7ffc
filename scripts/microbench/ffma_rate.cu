// FFMA vs FFMA2 (fma.rn.f32x2) issue rate on sm_100a: 8 independent chains per thread, 1024 threads per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma_rate scripts/microbench/ffma_rate.cu && /tmp/ffma_rate
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b, int iters) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_ffma2(float* out, float a, float b, int iters) {
    unsigned long long x[8], aa, bb;
    float2 av = make_float2(a, a), bv = make_float2(b, b);
    aa = *reinterpret_cast<unsigned long long*>(&av); bb = *reinterpret_cast<unsigned long long*>(&bv);
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 v = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i); x[i] = *reinterpret_cast<unsigned long long*>(&v); }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&x[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, sms * 8 * 256 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k_ffma<<<sms * 8, 256>>>(out, 1.0001f, 0.5f, iters); else k_ffma2<<<sms * 8, 256>>>(out, 1.0001f, 0.5f, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double fma = (double)sms * 8 * 256 * 16 * iters;
            if (rep) printf("%s: %.3f ms, %.2f TFMA/s = %.1f FMA/clk/SM at 1.9 GHz\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms / 1e9, fma / ms / 1e9 * 1e12 / sms / 1.9e9 / 1e0 / 1e0 / 1e0);
        }
    }
    return 0;
}
