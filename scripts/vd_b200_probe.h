/* Bring-up / tuning probes of the tensor-core path.  NOT part of the product library: these entry points exist only in
 * scripts/libvd_b200_probe.so, which scripts/_probe_lib.py builds from the same sources with -DVD_PROBE. */
#pragma once
#include "../include/vd_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Tuning probe (tests/bring-up only): issues 148 x n_sa x n_steps x n_acc MMAs of N=ncols with the
 * given descriptor words over dummy operands (pix >= 64 KiB, wimg >= 16 KiB, raw >= 148*n_acc*128*ncols
 * floats) so that the MMA rate of a shared-memory layout can be timed with CUDA events. */
int vd_tc_probe(const void* pix, const void* wimg, float* raw, int ncols, int n_sa, int n_steps,
                uint32_t a_lbo16, uint32_t a_hi, uint32_t b_lbo16, uint32_t b_hi, uint32_t b_step16,
                int n_acc, void* stream);

/* Hardware-floor probe: `grid` CTAs each issue iters x n_acc MMAs (M=128, N=ncols, K=16) from constant
 * descriptors; out[2*cta] = issue cycles, out[2*cta+1] = cycles until all MMAs completed. */
int vd_tc_mma_rate(long long* out, int n_acc, int ncols, int iters, uint32_t a_hi, uint32_t b_hi, uint32_t lbo16,
                   int vary, int grid, int delay, void* stream);
/* Bring-up probe: tcgen05.mma rate with moving A tiles / B windows (operand-fetch cost), not on the product path. */
int vd_tc_mma_rate2(long long* out, int n_acc, int ncols, int iters, uint32_t a_hi, uint32_t a_lbo16, uint32_t a_step16,
                    int a_n, uint32_t b_hi, uint32_t b_lbo16, uint32_t b_step16, int b_n, uint32_t b_base16, int group,
                    int same_acc, uint32_t fill, int group_delay, int grid, void* stream);   /* delay: emulated scalar cycles per step */

/* Tuning aid: device buffer (148*8 int64) receiving per-CTA cycle counters of the MMA warp of the forward
 * conv launches: [total, wait acc_empty, wait pix_full, wait w_full, issue]; NULL disables. */
int vd_tc_set_profile_buffer(long long* buf);


#ifdef __cplusplus
}
#endif
