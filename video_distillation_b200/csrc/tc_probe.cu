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
    if (threadIdx.x == 0) { mbar_init(base, 1); mbar_init(base + 8, 1); fence_mbar_init(); }
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
            const uint32_t v = (vary & 1) ? (uint32_t)(it & 3) : 0u;
            if (elect_one()) {
                const uint64_t a_desc = ((uint64_t)a_hi << 32) | (a16 & 0x3FFFu) | (128u << 16);
#pragma unroll
                for (int a = 0; a < NACC; ++a) {
                    const uint64_t b_desc = ((uint64_t)b_hi << 32) | ((b16 + v + (uint32_t)a * 64u) & 0x3FFFu) | (lbo16 << 16);
                    umma_bf16(tmem_base + (uint32_t)a * acc_cols, a_desc, b_desc, idesc, it > 0);
                }
            }
            // vary bits 8..15 = group size G, bit 16 = tcgen05.commit every G steps, bit 17 = tcgen05.fence::after_thread_sync
            // every G steps (emulates the stage boundaries of ws_gemm_kernel)
            const int grp = (vary >> 8) & 0xFF;
            if (grp && (it % grp) == grp - 1) {
                if ((vary >> 16) & 1) { if (elect_one()) umma_commit(base + 8); __syncwarp(); }
                if ((vary >> 17) & 1) tc_fence_after();
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


// Operand-fetch probe: like mma_rate_kernel, but the A tile and the B window move per MMA the way they do in
// ws_gemm_kernel (a_step16 / b_step16 in 16-byte units, cycling over a_n / b_n positions), so that the cost of
// non-repeating and 128-byte-misaligned operand reads shows.  Shared memory is filled with a bf16 pattern
// (fill = 0: zeros).  `same_acc` != 0: all MMAs of a step accumulate into accumulator 0 (dependency chain).
template <int NACC>
__global__ void __launch_bounds__(128, 1) mma_rate2_kernel(long long* out, int ncols, int iters, uint32_t a_hi, uint32_t a_lbo16,
                                                           uint32_t a_step16, int a_n, uint32_t b_hi, uint32_t b_lbo16,
                                                           uint32_t b_step16, int b_n, uint32_t b_base16, int group, int same_acc,
                                                           uint32_t fill, int group_delay) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bp + 16);
    for (int i = threadIdx.x; i < (200 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(bp + 1024)[i] = fill ? fill * (uint32_t)(i * 2654435761u) : 0u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(base, 1); mbar_init(base + 8, 1); fence_mbar_init(); }
    if (warp == 1) { tmem_alloc(base + 16, 512); tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, ncols);
        const uint32_t a16 = (base + 1024) >> 4, b16 = ((base + 1024 + 65536) >> 4) + b_base16;
        const uint32_t acc_cols = 512 / NACC;
        int ai = 0, bi = 0, gi = 0;
        uint32_t ao = 0, bo = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
                const uint64_t a_desc = ((uint64_t)a_hi << 32) | ((a16 + ao) & 0x3FFFu) | (a_lbo16 << 16);
#pragma unroll
                for (int a = 0; a < NACC; ++a) {
                    const uint64_t b_desc = ((uint64_t)b_hi << 32) | ((b16 + bo + (uint32_t)a * 448u) & 0x3FFFu) | (b_lbo16 << 16);
                    umma_bf16(tmem_base + (same_acc ? 0u : (uint32_t)a * acc_cols), a_desc, b_desc, idesc, it > 0);
                }
            }
            __syncwarp();
            if (++ai == a_n) { ai = 0; ao = 0; } else ao += a_step16;
            if (++bi == b_n) { bi = 0; bo = 0; } else bo += b_step16;
            if (group && ++gi == group) {
                gi = 0;
                if (elect_one()) umma_commit(base + 8);
                __syncwarp();
                tc_fence_after();
                if (group_delay > 0) {              // emulate the scalar work between two stages of ws_gemm_kernel
                    const long long d0 = clock64();
                    while (clock64() - d0 < group_delay) { }
                }
            }
        }
        if (elect_one()) umma_commit(base);
        mbar_wait(base, 0);
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[blockIdx.x * 2] = t2 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
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

extern "C" int vd_tc_mma_rate2(long long* out, int n_acc, int ncols, int iters, uint32_t a_hi, uint32_t a_lbo16, uint32_t a_step16,
                               int a_n, uint32_t b_hi, uint32_t b_lbo16, uint32_t b_step16, int b_n, uint32_t b_base16, int group,
                               int same_acc, uint32_t fill, int group_delay, int grid, void* stream) {
    VD_REQUIRE(out && (n_acc == 1 || n_acc == 2) && ncols % 16 == 0 && ncols * n_acc <= 512 && a_n >= 1 && b_n >= 1, "mma_rate2: bad argument");
    const int smem = 210 * 1024;
    cudaStream_t s = (cudaStream_t)stream;
#define GO(N) do { cudaFuncSetAttribute(mma_rate2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
                   mma_rate2_kernel<N><<<grid, 128, smem, s>>>(out, ncols, iters, a_hi, a_lbo16, a_step16, a_n, b_hi, b_lbo16, b_step16, \
                                                                b_n, b_base16, group, same_acc, fill, group_delay); } while (0)
    if (n_acc == 1) GO(1); else GO(2);
#undef GO
    return check_launch("tc_mma_rate2");
}
