// Hardware-floor probe for tcgen05.mma issue/execution rate (bring-up tool, not on the product path).
// One CTA per SM; warp 0 issues `iters` x `n_acc` MMAs (bf16, M=128, N=ncols, K=16) from constant
// descriptors over zeroed shared memory and reports clock64() cycles per MMA.
#include "tc_common.cuh"

namespace vd {
namespace tc {

template <int NACC>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int ncols, int iters, uint32_t a_hi, uint32_t b_hi,
                                                          uint32_t lbo16, int vary, int delay) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bar = reinterpret_cast<uint64_t*>(bp);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bp + 16);
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(bp + 1024)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(base, 1); fence_mbar_init(); }
    if (warp == 1) { tmem_alloc(base + 16, 512); tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, ncols);
        const uint32_t a16 = (base + 1024) >> 4, b16 = (base + 1024 + 16384) >> 4;
        const uint32_t acc_cols = 512 / NACC;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t v = vary ? (uint32_t)(it & 3) : 0u;
            if (elect_one()) {
                const uint64_t a_desc = ((uint64_t)a_hi << 32) | (a16 & 0x3FFFu) | (128u << 16);
#pragma unroll
                for (int a = 0; a < NACC; ++a) {
                    const uint64_t b_desc = ((uint64_t)b_hi << 32) | ((b16 + v + (uint32_t)a * 64u) & 0x3FFFu) | (lbo16 << 16);
                    umma_bf16(tmem_base + (uint32_t)a * acc_cols, a_desc, b_desc, idesc, it > 0);
                }
            }
            if (delay > 0) {                       // emulate `delay` cycles of scalar work per step
                const long long d0 = clock64();
                while (clock64() - d0 < delay) { }
            }
        }
        long long t1 = clock64();
        if (elect_one()) umma_commit(base);
        mbar_wait(base, 0);
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace tc
}  // namespace vd

using namespace vd;
using namespace vd::tc;

extern "C" int vd_tc_mma_rate(long long* out, int n_acc, int ncols, int iters, uint32_t a_hi, uint32_t b_hi, uint32_t lbo16,
                              int vary, int grid, int delay, void* stream) {
    VD_REQUIRE(out && (n_acc == 1 || n_acc == 2 || n_acc == 4) && ncols % 16 == 0 && ncols * n_acc <= 512, "mma_rate: bad argument");
    const int smem = 120 * 1024;
    cudaStream_t s = (cudaStream_t)stream;
#define GO(N) do { cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
                   mma_rate_kernel<N><<<grid, 128, smem, s>>>(out, ncols, iters, a_hi, b_hi, lbo16, vary, delay); } while (0)
    if (n_acc == 1) GO(1); else if (n_acc == 2) GO(2); else GO(4);
#undef GO
    return check_launch("tc_mma_rate");
}
