// Shifted-window implicit-GEMM Conv3d on tcgen05 tensor cores (sm_100a).
//
// One persistent, warp-specialised kernel serves the three ConvNet3D feature convolutions
// (networks.py:799).  Per CTA tile:
//     D[m, q] = sum_{stage} sum_{step} A_step[m, 0:16] . B_step[q, 0:16]^T        (bf16 x bf16 -> fp32)
//   m  = output channel (M = 128 lanes of TMEM; conv 0 stacks two output frames x 64 channels)
//   q  = output pixel of the tile (N = 16..256 TMEM columns), a LINEAR window of the packed input
//   A  = 128x16 weight tile, UMMA canonical K-major (no swizzle), streamed through a ring of
//        bulk-copied slots (conv 1/2) or resident in shared memory for the whole kernel (conv 0)
//   B  = 16-byte channel chunks of the packed activation staged ONCE per (tile, stage) by bulk
//        async copies; every filter tap is just a different descriptor start address.
// Warp roles: 0 pixel loader, 1 and 2 MMA issuers (one thread each, alternating MMA groups; 2 also allocates TMEM), 3 weight loader,
// 4..7 epilogue (TMEM -> registers -> bias/ReLU/MaxPool/argmax -> packed bf16 input of the next
// layer, or fp32 embeddings after conv 2).  Layouts: tc_layout.h.
#include <stdlib.h>

#include "tc_common.cuh"
#include "tc_layout.h"

namespace vd {
namespace tc {

constexpr int kMaxSteps = 80;
constexpr int kMaxCopies = 8;
constexpr int kThreads = 256;
constexpr int kThreadsL0S = 384;         // split-fp16 conv 0: a second group of four epilogue warps (see ws_gemm_kernel)

enum EpiMode { EPI_RAW = 0, EPI_L0 = 1, EPI_L1 = 2, EPI_L2 = 3, EPI_PLAIN = 4, EPI_DG1 = 5, EPI_DG0 = 6, EPI_L0S = 7, EPI_L1S = 8 };

struct EpiParams {
    float* raw;            // EPI_RAW: [tile][u][acc][128][ncols]
    int raw_bf16;          // EPI_RAW: store bf16 instead of fp32 (backward column buffers)
    const float* bias;     // per output channel (64 for conv 0, 128 otherwise)
    uint8_t* out;          // packed bf16 input of the next layer / fp32 embeddings
    uint8_t* code;         // optional ReLU/argmax routing codes, NCDHW order of the pooled tensor
    int code_first;        // items below this index record no codes (frozen real videos ahead of synthetic ones)
    int layer;             // EPI_PLAIN: which conv (tile -> NCDHW mapping)
    int accum;             // EPI_PLAIN, EPI_DG0, EPI_DG1 (plain output): out += result (split-bf16 passes accumulate into the same fp32 tensor)
    int ph;                // EPI_DG1: input-row parity of this launch
    int dy0_planar;        // EPI_DG1 (route mode): write the padded planar dY of conv 0 (Dg0Geo) instead of the column-GEMM operand
    int ncdhw;             // EPI_DG0: output (B,3,T,H,W) instead of the video layout (B,T,3,H,W)
    BwdGeo bb;             // EPI_DG1 (route mode): geometry of the packed dY operand of conv 0's column GEMM
    int T;                 // frames of the video
    int n_items;           // videos (conv 2: valid videos, tiles may be partially filled)
    int r0s;               // EPI_L0S: output rows per column band (SGeo::R0s)
    int hi_only;           // EPI_L0S / EPI_L1S: write only the hi part of the next layer's operand (two-product mode of the
                           //   frozen real videos: the next layer reads xh only)
    Geo g;
};

struct WsParams {
    const uint8_t* pix;
    const uint8_t* wimg;
    const int64_t* item_index;          // optional: item -> slot of `pix` (resident dataset)
    int32_t n_tiles, tiles_per_item, v_count;
    int64_t item_stride, u_stride, v_stride;
    int32_t n_u;                        // accumulator groups per tile sharing ONE pixel stage (column GEMM: M tiles)
    int32_t nu_total, ug_count;         // column GEMM: the nu_total M tiles of one pixel stage are split over ug_count CTA tiles
    int64_t w_u_stride;                 // bytes between the weight sequences of consecutive groups
    int64_t w_item_stride;              // bytes between the weight sequences of consecutive items (wgrad: split-K slices)
    int32_t n_sa, n_sb;
    int64_t sa_stride, sb_stride;
    int32_t n_copies;
    uint32_t stage_bytes;               // bytes landing per stage (sum of copy_bytes)
    uint32_t stage_pitch;               // smem distance between ring slots
    int64_t copy_gofs[kMaxCopies];
    uint32_t copy_sofs[kMaxCopies];
    uint32_t copy_bytes[kMaxCopies];
    int32_t n_steps;
    uint32_t b_off16[kMaxSteps];
    uint32_t b_lbo16[kMaxSteps];
    uint32_t a_off16[kMaxSteps];        // resident weights: offset of the tile of (sa=0, step)
    uint2 step_tab[kMaxSteps];          // issue table: x = A offset (16 B units) inside the slot / resident image,
                                        //              y = B offset | LBO << 16   (filled by finalize_smem)
    int32_t a_sa_stride16;              // resident weights: offset added per stage index
    uint32_t a_lbo16, a_sbo16;
    uint32_t a_hi, b_hi;                // upper descriptor words (SBO | version | layout type); 0 = default no-swizzle
    int32_t w_resident;
    uint32_t w_bytes;                   // resident image bytes
    int32_t G, RW, RP;
    // streamed weights with tile REUSE inside a group (split-fp16 conv 1 / 2: one hi-weight tile feeds the xh and the xl MMA):
    // a ring slot holds Gt tiles (0: = G) and a stage n_wtiles tiles (0: = n_steps); MMA jj of a group reads tile a_in_group[jj]
    // of its slot.  A partial last group uses a prefix of both the MMA and the tile sequence.
    int32_t Gt, n_wtiles;
    uint8_t a_in_group[16];
    int32_t n_acc;
    uint32_t acc_delta16;               // B start offset between accumulators
    uint32_t ncols, acc_cols, acc_stages;
    uint32_t idesc;
    uint32_t tab_bytes;                 // bytes of the descriptor tables behind the barrier block
    uint32_t smem_w_off, smem_pix_off;  // from the 1024-aligned dynamic smem base
    uint32_t smem_epi_off;              // != 0: 4 x 32 x 33 float transpose scratch for coalesced raw stores
    uint32_t smem_stash_off;            // != 0: conv-1 epilogue stash [H2*H2][kStashPitch] bf16 (quick accumulator drain)
    uint32_t stash_group_bytes;         // EPI_L0S: stash bytes of ONE epilogue group (two groups of four warps drain alternate frames)
    int32_t swap_ab;                    // the staged pixels are the M operand and the weight tiles the N operand (conv-0 dgrad)
    int32_t stream_pairs;               // conv 0 fused forward, != 0: "input-frame streaming" order (T/2 frame pairs per column):
                                        //   a tile is a COLUMN (video, row band); its T input frames are staged once each, in
                                        //   order, and every frame feeds the accumulators of the (up to) two frame pairs it
                                        //   touches through the 4 Toeplitz weight windows — half the L2->SM pixel traffic of the
                                        //   pair-by-pair order and no MMAs on the all-zero temporal halo frames
    int32_t stream_mode;                // 0 / 1: stream_pairs as described above; 2: split-fp16 conv 0 (SGeo): stream_pairs = T
                                        //   accumulators per column, stages (frame i, part), groups kt = 2, 1, 0 -> frame i + 1 - kt
    int32_t l0s_parts;                  // stream_mode 2: stages per input frame (2: hi and lo part; 1: hi part only = the
                                        //   two-product mode xh*wh + xh*wl of the frozen real videos)
    int32_t n_wsets;                    // resident weights: number of weight windows in the descriptor table (0: n_sa)
    int32_t dbg;                        // tuning experiments (VD_TC_DBG bitmask; results are garbage when set):
                                        //   1 no pixel copies, 2 no weight copies, 4 epilogue does no work,
                                        //   16 loaders back off with nanosleep (old behaviour), 32 epilogue spins without nanosleep
    long long* prof;                    // optional [grid][8] cycle counters of MMA issuer 0 (tuning only)
    EpiParams epi;
};

struct __align__(8) Barriers {
    uint64_t pix_full[4], pix_empty[4];
    uint64_t w_full[8], w_empty[8];
    uint64_t acc_full[4], acc_empty[4];
    uint64_t w_res;
    uint64_t baton[2];                  // MMA issuer hand-over (issuer r arrives on baton[r] after each of its groups)
    uint32_t tmem_base;
    uint32_t pad;
};
constexpr uint32_t kBarBlock = 512;      // barrier block; the MMA descriptor tables follow (WsParams::tab_bytes)
static_assert(sizeof(Barriers) <= kBarBlock, "barrier block too large");

// ------------------------------------------------------------------------------------------
// epilogues.  Thread `m` (0..127) owns TMEM lane m.  taddr = tmem base of the accumulator stage
// with the lane quarter already folded in.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) { return (uint32_t)f2bf(lo) | ((uint32_t)f2bf(hi) << 16); }

// slot = tile_pix * nu_total + u_abs: index of the 128-row block in the raw output
__device__ __forceinline__ void epi_raw(const WsParams& p, int64_t slot, uint32_t taddr, int m, float* scratch) {
    for (int a = 0; a < p.n_acc; ++a) {
        const int64_t blk = (slot * p.n_acc + a) * 128;
        if (p.epi.raw_bf16) {
            // backward column buffers (bf16).  A thread holds 32 columns (64 B) of its own row; lane pairs swap
            // 16-byte pieces so that every store instruction writes whole 32-byte sectors (even lane: first half,
            // odd lane: second half of the same sector) — half-filled sectors cost full L2 write slots.
            uint16_t* row_own = reinterpret_cast<uint16_t*>(p.epi.raw) + (blk + m) * (int64_t)p.ncols;
            uint16_t* row_peer = reinterpret_cast<uint16_t*>(p.epi.raw) + (blk + (m ^ 1)) * (int64_t)p.ncols;
            const bool odd = m & 1;
            uint16_t* row_a = odd ? row_peer : row_own;          // instructions 0,1 write the even lane's row
            uint16_t* row_b = odd ? row_own : row_peer;          // instructions 2,3 write the odd lane's row
            uint32_t c = 0;
            for (; c + 32 <= p.ncols; c += 32) {
                float v[32];
                tmem_ld32(taddr + a * p.acc_cols + c, v);
                tmem_ld_wait();
                uint4 pc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    pc[i] = make_uint4(pack_bf2(v[8 * i], v[8 * i + 1]), pack_bf2(v[8 * i + 2], v[8 * i + 3]),
                                       pack_bf2(v[8 * i + 4], v[8 * i + 5]), pack_bf2(v[8 * i + 6], v[8 * i + 7]));
                // even lane keeps pieces 0,2 of its row and needs pieces 0,2 of the odd row; odd lane keeps 1,3 and
                // needs 1,3 of the even row: send what the peer needs, receive what this lane stores
                const uint4 s0 = odd ? pc[0] : pc[1], s1 = odd ? pc[2] : pc[3];
                uint4 r0, r1;
                r0.x = __shfl_xor_sync(0xffffffffu, s0.x, 1); r0.y = __shfl_xor_sync(0xffffffffu, s0.y, 1);
                r0.z = __shfl_xor_sync(0xffffffffu, s0.z, 1); r0.w = __shfl_xor_sync(0xffffffffu, s0.w, 1);
                r1.x = __shfl_xor_sync(0xffffffffu, s1.x, 1); r1.y = __shfl_xor_sync(0xffffffffu, s1.y, 1);
                r1.z = __shfl_xor_sync(0xffffffffu, s1.z, 1); r1.w = __shfl_xor_sync(0xffffffffu, s1.w, 1);
                const int o = odd ? 8 : 0;                        // bf16 elements: odd lane writes the upper 16 bytes
                *reinterpret_cast<uint4*>(row_a + c + o) = odd ? r0 : pc[0];
                *reinterpret_cast<uint4*>(row_a + c + 16 + o) = odd ? r1 : pc[2];
                *reinterpret_cast<uint4*>(row_b + c + o) = odd ? pc[1] : r0;
                *reinterpret_cast<uint4*>(row_b + c + 16 + o) = odd ? pc[3] : r1;
            }
            for (; c < p.ncols; c += 16) {
                float v[16];
                tmem_ld16(taddr + a * p.acc_cols + c, v);
                tmem_ld_wait();
                uint4* d4 = reinterpret_cast<uint4*>(row_own + c);
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    d4[i] = make_uint4(pack_bf2(v[8 * i], v[8 * i + 1]), pack_bf2(v[8 * i + 2], v[8 * i + 3]),
                                       pack_bf2(v[8 * i + 4], v[8 * i + 5]), pack_bf2(v[8 * i + 6], v[8 * i + 7]));
            }
            continue;
        }
        float* tile_base = p.epi.raw + blk * (int64_t)p.ncols;
        if (scratch == nullptr) {                      // bring-up path (forward layers in raw mode): row per thread
            float* dst = tile_base + (int64_t)m * p.ncols;
            for (uint32_t c = 0; c < p.ncols; c += 8) {
                float v[8];
                tmem_ld8(taddr + a * p.acc_cols + c, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) dst[c + i] = v[i];
            }
        } else {                                       // coalesced fp32: transpose 32x32 blocks through shared memory
            const int lane = m & 31;
            float* s = scratch + (m >> 5) * (32 * 33);
            float* rows = tile_base + (int64_t)(m & ~31) * p.ncols;
            for (uint32_t c0 = 0; c0 < p.ncols; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + a * p.acc_cols + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) s[lane * 33 + i] = v[i];
                __syncwarp();
                if (c0 + lane < p.ncols) {
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) rows[(int64_t)r * p.ncols + c0 + lane] = s[r * 33 + lane];
                }
                __syncwarp();
            }
        }
    }
}

// conv 0: lanes 0..63 = frame 2*tp, lanes 64..127 = frame 2*tp+1, columns q = r*Wo0 + wo.
// bias + ReLU + MaxPool(1,2,2) -> A1 chunks (bf16) [+ code (B,64,T,H1,H1)]
__device__ __forceinline__ void epi_l0(const WsParams& p, int item, int tp, int rb, uint32_t taddr, int m, float bias_reg) {
    const Geo& g = p.epi.g;
    const int f = 2 * tp + (m >> 6), co = m & 63;
    const float bias = bias_reg;              // loaded once per kernel (ws_gemm_kernel): a global load per accumulator costs ~600 exposed cycles
    const int slice = co >> 4, k = (co >> 3) & 1, e = co & 7;
    uint8_t* vbase = p.epi.out + (int64_t)item * g.video1 + (int64_t)slice * g.slice1 + (int64_t)(f + 1) * g.frame1;
    uint8_t* cbase = (p.epi.code && item >= p.epi.code_first)
                         ? p.epi.code + (((int64_t)(item - p.epi.code_first) * 64 + co) * g.T + f) * g.H1 * g.H1 : nullptr;
    for (int pr = 0; pr < g.R0 / 2; ++pr) {
        const int hp = (rb * g.R0) / 2 + pr;
        const int ph = coord_par(hp), pi = coord_pos(hp);
        for (int wb = 0; wb < g.Wo0; wb += 8) {
            float r0[8], r1[8];
            tmem_ld8(taddr + (2 * pr) * g.Wo0 + wb, r0);
            tmem_ld8(taddr + (2 * pr + 1) * g.Wo0 + wb, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // scan order (h,w): (0,0) (0,1) (1,0) (1,1); first maximum wins
                float best = r0[2 * j]; int arg = 0;
                if (r0[2 * j + 1] > best) { best = r0[2 * j + 1]; arg = 1; }
                if (r1[2 * j] > best) { best = r1[2 * j]; arg = 2; }
                if (r1[2 * j + 1] > best) { best = r1[2 * j + 1]; arg = 3; }
                best += bias;
                const bool act = best > 0.f;
                const int wp = wb / 2 + j;
                const int pw = coord_par(wp), pj = coord_pos(wp);
                uint8_t* dst = vbase + (int64_t)(((ph * 2 + pw) * 2 + k)) * g.plane1 + ((int64_t)pi * g.P1 + pj) * 16 + e * 2;
                *reinterpret_cast<uint16_t*>(dst) = f2bf(act ? best : 0.f);
                if (cbase) cbase[hp * g.H1 + wp] = (uint8_t)(arg | (act ? 8 : 0));
            }
        }
    }
}

// conv 0 in two phases (as conv 1 below), so that the accumulator is released after a short drain and the packed
// output leaves as whole 16-byte chunks instead of 2-byte pieces:
//   drain: TMEM -> bias / ReLU / MaxPool(1,2,2) -> bf16 stash [frame half][pooled position][channel]  (+ routing codes)
//   store: every lane takes (chunk of 8 channels, position) items of its warp's 32 channels -> one 16-byte store each.
constexpr int kStash0Pitch = 72;          // bf16 elements per position row (64 + 8: conflict-free 16-byte reads)

__device__ __forceinline__ void epi_l0_drain(const WsParams& p, int item, int tp, int rb, uint32_t taddr, int m, uint16_t* stash, float bias_reg) {
    const Geo& g = p.epi.g;
    const int half = m >> 6, f = 2 * tp + half, co = m & 63;
    const float bias = bias_reg;              // loaded once per kernel (ws_gemm_kernel): a global load per accumulator costs ~600 exposed cycles
    const int Wp = g.Wo0 / 2, npos = (g.R0 / 2) * Wp;
    uint16_t* srow = stash + (half * npos) * kStash0Pitch + co;
    if (!(p.epi.code && item >= p.epi.code_first)) {
        // frozen real videos (no routing codes): max / add / max, ~1/3 of the instructions of the argmax path
        for (int pr = 0; pr < g.R0 / 2; ++pr) {
            uint16_t* sp = srow + (pr * Wp) * kStash0Pitch;
            for (int wb = 0; wb < g.Wo0; wb += 8, sp += 4 * kStash0Pitch) {
                float r0[8], r1[8];
                tmem_ld8(taddr + (2 * pr) * g.Wo0 + wb, r0);
                tmem_ld8(taddr + (2 * pr + 1) * g.Wo0 + wb, r1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    const float b0 = fmaxf(fmaxf(fmaxf(r0[2 * j], r0[2 * j + 1]), fmaxf(r1[2 * j], r1[2 * j + 1])) + bias, 0.f);
                    const float b1 = fmaxf(fmaxf(fmaxf(r0[2 * j + 2], r0[2 * j + 3]), fmaxf(r1[2 * j + 2], r1[2 * j + 3])) + bias, 0.f);
                    const uint32_t pk = pack_bf2(b0, b1);
                    sp[j * kStash0Pitch] = (uint16_t)pk;
                    sp[(j + 1) * kStash0Pitch] = (uint16_t)(pk >> 16);
                }
            }
        }
        return;
    }
    uint8_t* cbase = p.epi.code + (((int64_t)(item - p.epi.code_first) * 64 + co) * g.T + f) * g.H1 * g.H1;
    for (int pr = 0; pr < g.R0 / 2; ++pr) {
        const int hp = (rb * g.R0) / 2 + pr;
        for (int wb = 0; wb < g.Wo0; wb += 8) {
            float r0[8], r1[8];
            tmem_ld8(taddr + (2 * pr) * g.Wo0 + wb, r0);
            tmem_ld8(taddr + (2 * pr + 1) * g.Wo0 + wb, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // scan order (h,w): (0,0) (0,1) (1,0) (1,1); first maximum wins
                float best = r0[2 * j]; int arg = 0;
                if (r0[2 * j + 1] > best) { best = r0[2 * j + 1]; arg = 1; }
                if (r1[2 * j] > best) { best = r1[2 * j]; arg = 2; }
                if (r1[2 * j + 1] > best) { best = r1[2 * j + 1]; arg = 3; }
                best += bias;
                const bool act = best > 0.f;
                const int wp = wb / 2 + j;
                srow[(pr * Wp + wp) * kStash0Pitch] = f2bf(act ? best : 0.f);
                cbase[hp * g.H1 + wp] = (uint8_t)(arg | (act ? 8 : 0));
            }
        }
    }
}

// warp q owns lanes 32q..32q+31 = frame half q >> 1, channels (q & 1) * 32 .. + 31 (4 chunks of 8 channels)
__device__ __forceinline__ void epi_l0_store(const WsParams& p, int item, int tp, int rb, int q, int lane, const uint16_t* stash) {
    const Geo& g = p.epi.g;
    const int half = q >> 1, f = 2 * tp + half;
    const int Wp = g.Wo0 / 2, npos = (g.R0 / 2) * Wp;
    uint8_t* fbase = p.epi.out + (int64_t)item * g.video1 + (int64_t)(f + 1) * g.frame1;
    const uint16_t* sbase = stash + (half * npos) * kStash0Pitch + (q & 1) * 32;
    for (int pr = 0; pr < g.R0 / 2; ++pr) {
        const int hp = (rb * g.R0) / 2 + pr;
        const int64_t row_off = (int64_t)(coord_par(hp) * 4) * g.plane1 + (int64_t)coord_pos(hp) * g.P1 * 16;
        for (int wp = lane; wp < Wp; wp += 32) {
            const uint16_t* sp = sbase + (pr * Wp + wp) * kStash0Pitch;
            uint8_t* dpos = fbase + row_off + (int64_t)(coord_par(wp) * 2) * g.plane1 + coord_pos(wp) * 16;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int c8 = (q & 1) * 4 + kk;              // chunk of 8 channels: slice = c8 >> 1, k = c8 & 1
                const uint4 v = *reinterpret_cast<const uint4*>(sp + kk * 8);
                *reinterpret_cast<uint4*>(dpos + (int64_t)(c8 >> 1) * g.slice1 + (int64_t)(c8 & 1) * g.plane1) = v;
            }
        }
    }
}

// conv 0, split-fp16 operands (SGeo): accumulator = output frame f of the column (item, rb), columns q = r*Wo0 + wo for
// the R0s rows of the band.  TMEM lane m = 32*q + l holds channel co = 16*q + (l & 15); lanes l < 16 accumulated wh*x,
// lanes l >= 16 wl*x (x = xh + xl over the two stages of a frame): y = top + bottom, exchanged with shuffles inside the warp
// (lane l keeps the columns it pools and sends the partner's).  bias + ReLU + MaxPool(1,2,2) on the fp32 sum -> fp16 hi / lo
// stash [part][pooled position][channel] (+ routing codes for the synthetic videos).
constexpr int kStash0sPitch = 72;         // fp16 elements per position row (64 + 8: conflict-free 16-byte reads)

// [c0, c1): the accumulator columns (multiples of 8 inside a row) this epilogue group drains
__device__ __forceinline__ void epi_l0s_drain(const WsParams& p, int item, int f, int rb, uint32_t taddr, int m, uint16_t* stash,
                                              int c0, int c1, float bias_reg) {
    const Geo& g = p.epi.g;
    const int l = m & 31, hsel = l >> 4, co = (m >> 5) * 16 + (l & 15);
    const float bias = bias_reg;              // loaded once per kernel (ws_gemm_kernel): a global load per accumulator costs ~600 exposed cycles
    const int R = p.epi.r0s, Wp = g.Wo0 / 2, npos = (R / 2) * Wp;
    uint16_t* s_hi = stash + co;
    uint16_t* s_lo = stash + npos * kStash0sPitch + co;
    uint8_t* cbase = (p.epi.code && item >= p.epi.code_first)
                         ? p.epi.code + (((int64_t)(item - p.epi.code_first) * 64 + co) * g.T + f) * g.H1 * g.H1 : nullptr;
    if (cbase == nullptr) {
        // frozen real videos (no routing codes): max / add / max instead of the argmax scan.  The epilogue warps run one
        // instruction stream each, so every dependent-instruction latency is exposed: no divisions in here.
        const bool hi_only = p.epi.hi_only != 0;
        const int Wo = g.Wo0;
        // All TMEM loads of a row pair are issued before the single wait, and the (up to) four blocks of 8 columns are pooled
        // in ONE basic block so that their dependent chains (SEL -> SHFL -> FADD -> FMNMX.. -> F2F -> STS, ~190 cycles each)
        // interleave: a group with fewer than four blocks re-reads its last block and only the stores are predicated.
        const int nblk = (c1 - c0) >> 3;                            // 1..4 blocks of 8 columns per group (Wo0 <= 64)
        for (int pr = 0; pr < R / 2; ++pr) {
            const uint32_t t0 = taddr + (2 * pr) * Wo + c0, t1 = t0 + Wo;
            const int posr = pr * Wp + 2 * hsel + (c0 >> 1);
            float r0[32], r1[32];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int bb = min(b, nblk - 1);
                tmem_ld8(t0 + 8 * bb, r0 + 8 * b); tmem_ld8(t1 + 8 * bb, r1 + 8 * b);
            }
            tmem_ld_wait();
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                float a0[4], a1[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float s0 = hsel ? r0[8 * b + i] : r0[8 * b + 4 + i], s1 = hsel ? r1[8 * b + i] : r1[8 * b + 4 + i];
                    const float k0 = hsel ? r0[8 * b + 4 + i] : r0[8 * b + i], k1 = hsel ? r1[8 * b + 4 + i] : r1[8 * b + i];
                    a0[i] = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
                    a1[i] = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
                }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const float v = fmaxf(fmaxf(fmaxf(a0[2 * jj], a0[2 * jj + 1]), fmaxf(a1[2 * jj], a1[2 * jj + 1])) + bias, 0.f);
                    uint16_t hi, lo;
                    split_h(v, hi, lo);
                    if (b < nblk) {
                        s_hi[(posr + 4 * b + jj) * kStash0sPitch] = hi;
                        if (!hi_only) s_lo[(posr + 4 * b + jj) * kStash0sPitch] = lo;
                    }
                }
            }
        }
        return;
    }
    for (int pr = 0; pr < R / 2; ++pr) {
        const int hp = rb * (R / 2) + pr;
        for (int wb = c0; wb < c1; wb += 8) {
            float r0[8], r1[8];
            tmem_ld8(taddr + (2 * pr) * g.Wo0 + wb, r0);
            tmem_ld8(taddr + (2 * pr + 1) * g.Wo0 + wb, r1);
            tmem_ld_wait();
            float a0[4], a1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float s0 = hsel ? r0[i] : r0[4 + i], s1 = hsel ? r1[i] : r1[4 + i];       // columns the partner pools
                const float k0 = hsel ? r0[4 + i] : r0[i], k1 = hsel ? r1[4 + i] : r1[i];       // columns this lane pools
                a0[i] = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
                a1[i] = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                // scan order (h,w): (0,0) (0,1) (1,0) (1,1); first maximum wins
                float best = a0[2 * jj]; int arg = 0;
                if (a0[2 * jj + 1] > best) { best = a0[2 * jj + 1]; arg = 1; }
                if (a1[2 * jj] > best) { best = a1[2 * jj]; arg = 2; }
                if (a1[2 * jj + 1] > best) { best = a1[2 * jj + 1]; arg = 3; }
                best += bias;
                const bool act = best > 0.f;
                const int wp = wb / 2 + 2 * hsel + jj;
                uint16_t hi, lo;
                split_h(act ? best : 0.f, hi, lo);
                s_hi[(pr * Wp + wp) * kStash0sPitch] = hi;
                if (!p.epi.hi_only) s_lo[(pr * Wp + wp) * kStash0sPitch] = lo;
                if (cbase) cbase[hp * g.H1 + wp] = (uint8_t)(arg | (act ? 8 : 0));
            }
        }
    }
}

// warp q owns channels 16q .. 16q+15 = chunks 2q, 2q+1 of A1s: (position, part, chunk) items -> one 16-byte store each
// The store of a drained tile: every lane owns up to kL0sItems (position, part, chunk) items of its group's columns -> one
// 16-byte store each.  The stash offset and the offset inside the frame depend only on (lane, tile): they are computed once
// per tile (epi_l0s_store_plan) so that the per-frame path is LDS.128 -> STG.128.
constexpr int kL0sItems = 4;
struct L0sStorePlan { int n; int soff[kL0sItems]; int doff[kL0sItems]; };

__device__ __forceinline__ void epi_l0s_store_plan(const WsParams& p, int rb, int q, int lane, int c0, int c1, L0sStorePlan& sp) {
    const Geo& g = p.epi.g;
    const int R = p.epi.r0s, Wp = g.Wo0 / 2, npos = (R / 2) * Wp;
    const int w0 = c0 / 2, nw = (c1 - c0) / 2, npg = (R / 2) * nw;     // this group's pooled columns [w0, w0 + nw) of every row
    const int n_it = (p.epi.hi_only ? 2 : 4) * npg;
    sp.n = 0;
#pragma unroll
    for (int k = 0; k < kL0sItems; ++k) {
        const int it = lane + 32 * k;
        sp.soff[k] = 0; sp.doff[k] = 0;
        if (it < n_it) {
            // sel = it / npg (< 4) and pr = posl / nw (< R / 2 <= 2) by comparisons
            const int sel = (it >= npg) + (it >= 2 * npg) + (it >= 3 * npg), posl = it - sel * npg;
            const int part = sel >> 1, kk = sel & 1;
            const int pr = posl >= nw ? 1 : 0, wp = w0 + posl - pr * nw;
            const int hp = rb * (R / 2) + pr;
            sp.soff[k] = (part * npos + pr * Wp + wp) * kStash0sPitch + q * 16 + kk * 8;
            sp.doff[k] = (int)((int64_t)(2 * q + kk) * g.slice1 + (int64_t)((part * 2 + coord_par(hp)) * 2 + coord_par(wp)) * g.plane1 +
                               ((int64_t)coord_pos(hp) * g.P1 + coord_pos(wp)) * 16);
            sp.n = k + 1;
        }
    }
}

__device__ __forceinline__ void epi_l0s_store(uint8_t* fbase, const uint16_t* stash, const L0sStorePlan& sp) {
#pragma unroll
    for (int k = 0; k < kL0sItems; ++k)
        if (k < sp.n) *reinterpret_cast<uint4*>(fbase + sp.doff[k]) = *reinterpret_cast<const uint4*>(stash + sp.soff[k]);
}

// conv 1: accumulator a = frame 2*tp + a, lane = cout, columns q = ho*P1 + wo.
// bias + ReLU + MaxPool(2,2,2) -> A2 chunks (bf16, every tap copy) [+ code (B,128,T2,H2,H2)]
// Two phases so that the (single-buffered, 2 x 256 column) accumulator is released early:
//   drain: TMEM -> pooled bf16 values in a shared-memory stash [pos][channel]   (then acc_empty arrives)
//   store: every lane takes (16-byte chunk of 8 channels, position) items of its warp's 32 channels and
//          writes the chunk to all tap copies of A2 — 16-byte stores, overlapped with the next tile's MMAs.
constexpr int kStashPitch = 136;          // bf16 elements per position row (128 + 8: conflict-free 16-byte reads)

__device__ __forceinline__ void epi_l1_drain(const WsParams& p, int tile, uint32_t taddr, int m, uint16_t* stash, float bias_reg) {
    const Geo& g = p.epi.g;
    const int item = tile / p.tiles_per_item, tq = tile % p.tiles_per_item;
    const int npair = p.n_acc >> 1;                     // accumulator pairs of the tile = pooled frames (2, or 4 accumulators at 64x64)
    const float bias = bias_reg;              // loaded once per kernel (ws_gemm_kernel): a global load per accumulator costs ~600 exposed cycles
    for (int pp = 0; pp < npair; ++pp) {
        const int tp = tq * npair + pp;                 // pooled frame index
        const uint32_t ta = taddr + (uint32_t)(2 * pp) * p.acc_cols;
        uint16_t* st = stash + pp * (g.H2 * g.H2) * kStashPitch;
        uint8_t* cbase = (p.epi.code && item >= p.epi.code_first)
                             ? p.epi.code + (((int64_t)(item - p.epi.code_first) * 128 + m) * g.T2 + tp) * g.H2 * g.H2 : nullptr;
        for (int hp = 0; hp < g.H2; ++hp) {
            float a0[16], a1[16], b0[16], b1[16];
            tmem_ld16(ta + (2 * hp) * g.P1, a0);
            tmem_ld16(ta + (2 * hp + 1) * g.P1, a1);
            tmem_ld16(ta + p.acc_cols + (2 * hp) * g.P1, b0);
            tmem_ld16(ta + p.acc_cols + (2 * hp + 1) * g.P1, b1);
            tmem_ld_wait();
#pragma unroll
            for (int wp = 0; wp < 8; ++wp) {
                if (wp < g.H2) {
                    // scan order (t,h,w)
                    float best = a0[2 * wp]; int arg = 0;
                    if (a0[2 * wp + 1] > best) { best = a0[2 * wp + 1]; arg = 1; }
                    if (a1[2 * wp] > best) { best = a1[2 * wp]; arg = 2; }
                    if (a1[2 * wp + 1] > best) { best = a1[2 * wp + 1]; arg = 3; }
                    if (b0[2 * wp] > best) { best = b0[2 * wp]; arg = 4; }
                    if (b0[2 * wp + 1] > best) { best = b0[2 * wp + 1]; arg = 5; }
                    if (b1[2 * wp] > best) { best = b1[2 * wp]; arg = 6; }
                    if (b1[2 * wp + 1] > best) { best = b1[2 * wp + 1]; arg = 7; }
                    best += bias;
                    const bool act = best > 0.f;
                    st[(hp * g.H2 + wp) * kStashPitch + m] = f2bf(act ? best : 0.f);
                    if (cbase) cbase[hp * g.H2 + wp] = (uint8_t)(arg | (act ? 8 : 0));
                }
            }
        }
    }
}

__device__ __forceinline__ void epi_l1_store(const WsParams& p, int tile, int q, int lane, const uint16_t* stash) {
    const Geo& g = p.epi.g;
    const int item = tile / p.tiles_per_item, tq = tile % p.tiles_per_item;
    const int npair = p.n_acc >> 1;
    const int half = q >> 1;
    const int npos = g.H2 * g.H2;
    for (int pp = 0; pp < npair; ++pp) {
        const int tp = tq * npair + pp;
        const uint16_t* st = stash + pp * npos * kStashPitch;
        uint8_t* vbase = p.epi.out + (int64_t)item * g.video2 + (int64_t)half * g.group2 + (int64_t)(tp + 1) * g.HW2 * 16;
        for (int i = lane; i < 4 * npos; i += 32) {
            const int kk = i / npos, pos = i - kk * npos;
            const int hp = pos / g.H2, wp = pos - hp * g.H2;
            const uint4 v = *reinterpret_cast<const uint4*>(st + pos * kStashPitch + q * 32 + kk * 8);
            uint8_t* cb = vbase + (int64_t)((q & 1) * 4 + kk) * g.chunk2;
            // input pixel (t=tp, h=hp, w=wp) of conv 2 feeds output (ho,wo) through tap (kh,kw) iff
            // hp = 2ho+kh-3, wp = 2wo+kw-3
            for (int kh = (hp + 1) & 1; kh < 7; kh += 2) {
                const int ho = (hp + 3 - kh) / 2;
                if (hp + 3 - kh < 0 || ho >= g.Ho2) continue;
                for (int kw = (wp + 1) & 1; kw < 7; kw += 2) {
                    const int wo = (wp + 3 - kw) / 2;
                    if (wp + 3 - kw < 0 || wo >= g.Wo2) continue;
                    *reinterpret_cast<uint4*>(cb + (int64_t)((kh * 7 + kw) * 2) * g.group2 + (ho * g.Wo2 + wo) * 16) = v;
                }
            }
        }
    }
}

// conv 1, split-fp16 operands: same tile as epi_l1_*; the pooled fp32 value is split into fp16 hi / lo (stash
// [part][pooled frame][position][channel]) and both parts go to every tap copy of A2s
// ([khw 49][quarter 4][part 2][k 4][t_pad][ho][wo] x 16 B).
__device__ __forceinline__ void epi_l1s_drain(const WsParams& p, int tile, uint32_t taddr, int m, uint16_t* stash, float bias_reg) {
    const Geo& g = p.epi.g;
    const int item = tile / p.tiles_per_item, tq = tile % p.tiles_per_item;
    const int npair = p.n_acc >> 1;
    const int npos = g.H2 * g.H2;
    const float bias = bias_reg;              // loaded once per kernel (ws_gemm_kernel): a global load per accumulator costs ~600 exposed cycles
    for (int pp = 0; pp < npair; ++pp) {
        const int tp = tq * npair + pp;
        const uint32_t ta = taddr + (uint32_t)(2 * pp) * p.acc_cols;
        uint16_t* st_hi = stash + pp * npos * kStashPitch;
        uint16_t* st_lo = stash + (npair + pp) * npos * kStashPitch;
        uint8_t* cbase = (p.epi.code && item >= p.epi.code_first)
                             ? p.epi.code + (((int64_t)(item - p.epi.code_first) * 128 + m) * g.T2 + tp) * g.H2 * g.H2 : nullptr;
        for (int hp = 0; hp < g.H2; ++hp) {
            float a0[16], a1[16], b0[16], b1[16];
            tmem_ld16(ta + (2 * hp) * g.P1, a0);
            tmem_ld16(ta + (2 * hp + 1) * g.P1, a1);
            tmem_ld16(ta + p.acc_cols + (2 * hp) * g.P1, b0);
            tmem_ld16(ta + p.acc_cols + (2 * hp + 1) * g.P1, b1);
            tmem_ld_wait();
#pragma unroll
            for (int wp = 0; wp < 8; ++wp) {
                if (wp < g.H2) {
                    // scan order (t,h,w)
                    float best = a0[2 * wp]; int arg = 0;
                    if (a0[2 * wp + 1] > best) { best = a0[2 * wp + 1]; arg = 1; }
                    if (a1[2 * wp] > best) { best = a1[2 * wp]; arg = 2; }
                    if (a1[2 * wp + 1] > best) { best = a1[2 * wp + 1]; arg = 3; }
                    if (b0[2 * wp] > best) { best = b0[2 * wp]; arg = 4; }
                    if (b0[2 * wp + 1] > best) { best = b0[2 * wp + 1]; arg = 5; }
                    if (b1[2 * wp] > best) { best = b1[2 * wp]; arg = 6; }
                    if (b1[2 * wp + 1] > best) { best = b1[2 * wp + 1]; arg = 7; }
                    best += bias;
                    const bool act = best > 0.f;
                    uint16_t hi, lo;
                    split_h(act ? best : 0.f, hi, lo);
                    st_hi[(hp * g.H2 + wp) * kStashPitch + m] = hi;
                    if (!p.epi.hi_only) st_lo[(hp * g.H2 + wp) * kStashPitch + m] = lo;
                    if (cbase) cbase[hp * g.H2 + wp] = (uint8_t)(arg | (act ? 8 : 0));
                }
            }
        }
    }
}

__device__ __forceinline__ void epi_l1s_store(const WsParams& p, int tile, int q, int lane, const uint16_t* stash) {
    const Geo& g = p.epi.g;
    const int item = tile / p.tiles_per_item, tq = tile % p.tiles_per_item;
    const int npair = p.n_acc >> 1;
    const int npos = g.H2 * g.H2;
    const int n_it = (p.epi.hi_only ? 4 : 8) * npos;
    for (int pp = 0; pp < npair; ++pp) {
        const int tp = tq * npair + pp;
        uint8_t* vbase = p.epi.out + (int64_t)item * (196 * g.group2) + (int64_t)q * g.group2 + (int64_t)(tp + 1) * g.HW2 * 16;
        for (int i = lane; i < n_it; i += 32) {
            const int sel = i / npos, pos = i - sel * npos;
            const int part = sel >> 2, kk = sel & 3;                 // warp q = quarter q, chunk kk of the quarter
            const int hp = pos / g.H2, wp = pos - hp * g.H2;
            const uint4 v = *reinterpret_cast<const uint4*>(stash + ((part * npair + pp) * npos + pos) * kStashPitch + q * 32 + kk * 8);
            uint8_t* cb = vbase + (int64_t)(part * 4 + kk) * g.chunk2;
            for (int kh = (hp + 1) & 1; kh < 7; kh += 2) {
                const int ho = (hp + 3 - kh) / 2;
                if (hp + 3 - kh < 0 || ho >= g.Ho2) continue;
                for (int kw = (wp + 1) & 1; kw < 7; kw += 2) {
                    const int wo = (wp + 3 - kw) / 2;
                    if (wp + 3 - kw < 0 || wo >= g.Wo2) continue;
                    *reinterpret_cast<uint4*>(cb + (int64_t)((kh * 7 + kw) * 4) * g.group2 + (ho * g.Wo2 + wo) * 16) = v;
                }
            }
        }
    }
}

// conv 2: accumulator a = video item*4 + a, lane = cout, columns q = to*HW2 + ho*Wo2 + wo.
// bias + ReLU + MaxPool(2,2,2) -> fp32 embeddings (B, 128*T3p*H3p*H3p), NCDHW flatten order
// (networks.py:750) [+ code (B,128,T3p,H3p,H3p)]
__device__ __forceinline__ void epi_l2(const WsParams& p, int tile, uint32_t taddr, int m, float bias_reg) {
    const Geo& g = p.epi.g;
    const float bias = bias_reg;              // loaded once per kernel (ws_gemm_kernel): a global load per accumulator costs ~600 exposed cycles
    const int per = g.T3p * g.H3p * g.H3p;
    for (int a = 0; a < p.n_acc; ++a) {
        const int video = tile * p.n_acc + a;
        if (video >= p.epi.n_items) break;
        float* dst = reinterpret_cast<float*>(p.epi.out) + (int64_t)video * g.embed_dim + (int64_t)m * per;
        uint8_t* cdst = (p.epi.code && video >= p.epi.code_first)
                            ? p.epi.code + (int64_t)(video - p.epi.code_first) * g.embed_dim + (int64_t)m * per : nullptr;
        for (int tq = 0; tq < g.T3p; ++tq) {
            // two frames = 2*HW2 consecutive columns
            const uint32_t c0 = taddr + a * p.acc_cols + (2 * tq) * g.HW2;
            if (g.HW2 == 16) {
                float f0[16], f1[16];
                tmem_ld16(c0, f0);
                tmem_ld16(c0 + 16, f1);
                tmem_ld_wait();
#pragma unroll
                for (int hp = 0; hp < 2; ++hp)
#pragma unroll
                    for (int wp = 0; wp < 2; ++wp) {
                        float best = -INFINITY; int arg = 0, pos = 0;
#pragma unroll
                        for (int dt = 0; dt < 2; ++dt)
#pragma unroll
                            for (int dh = 0; dh < 2; ++dh)
#pragma unroll
                                for (int dw = 0; dw < 2; ++dw, ++pos) {
                                    const float v = (dt ? f1 : f0)[(2 * hp + dh) * 4 + 2 * wp + dw];
                                    if (v > best) { best = v; arg = pos; }
                                }
                        best += bias;
                        const bool act = best > 0.f;
                        dst[(tq * 2 + hp) * 2 + wp] = act ? best : 0.f;
                        if (cdst) cdst[(tq * 2 + hp) * 2 + wp] = (uint8_t)(arg | (act ? 8 : 0));
                    }
            } else {   // HW2 == 4 (64x64 videos): one pooled output per frame pair
                float f[8];
                tmem_ld8(c0, f);
                tmem_ld_wait();
                float best = -INFINITY; int arg = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) if (f[i] > best) { best = f[i]; arg = i; }
                best += bias;
                const bool act = best > 0.f;
                dst[tq] = act ? best : 0.f;
                if (cdst) cdst[tq] = (uint8_t)(arg | (act ? 8 : 0));
            }
        }
    }
}

// plain conv output: fp32 NCDHW (B, Cout, To, Ho, Wo), pre-activation (+ bias when given) — the fprop of the
// differentiable conv trio (ops.py) on tensor cores.  Tile -> output mapping as in epi_l0 / epi_l1 / epi_l2.
__device__ __forceinline__ void epi_plain(const WsParams& p, int tile, uint32_t taddr, int m) {
    const Geo& g = p.epi.g;
    float* out = reinterpret_cast<float*>(p.epi.out);
    if (p.epi.layer == 0) {
        const int item = tile / p.tiles_per_item, sub = tile % p.tiles_per_item;
        const int tp = sub / p.v_count, rb = sub % p.v_count;
        const int f = 2 * tp + (m >> 6), co = m & 63;
        const float bias = p.epi.bias ? __ldg(p.epi.bias + co) : 0.f;
        // 32 columns per round: all TMEM loads and (when accumulating) all global loads of the round are in flight together —
        // one lane per scheduler cannot hide a chain of load -> wait -> load -> add -> store per 8 columns
        for (int r = 0; r < g.R0; ++r) {
            float* dst = out + ((((int64_t)item * 64 + co) * g.T + f) * g.Ho0 + rb * g.R0 + r) * g.Wo0;
            for (int wb0 = 0; wb0 < g.Wo0; wb0 += 32) {
                float v[32];
                float4 o[8];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (wb0 + 8 * c < g.Wo0) tmem_ld8(taddr + r * g.Wo0 + wb0 + 8 * c, v + 8 * c);
                if (p.epi.accum) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (wb0 + 4 * c < g.Wo0) o[c] = *reinterpret_cast<const float4*>(dst + wb0 + 4 * c);
                }
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (wb0 + 4 * c < g.Wo0) {
                        float4 q = make_float4(v[4 * c] + bias, v[4 * c + 1] + bias, v[4 * c + 2] + bias, v[4 * c + 3] + bias);
                        if (p.epi.accum) { q.x += o[c].x; q.y += o[c].y; q.z += o[c].z; q.w += o[c].w; }
                        *reinterpret_cast<float4*>(dst + wb0 + 4 * c) = q;
                    }
            }
        }
    } else if (p.epi.layer == 1) {
        const int item = tile / p.tiles_per_item, tp = tile % p.tiles_per_item;
        const float bias = p.epi.bias ? __ldg(p.epi.bias + m) : 0.f;
        const int fpt = p.n_acc;                          // output frames per tile (2, or 4 for the split tables at 64x64)
        for (int a = 0; a < fpt; ++a) {
            float* dst = out + (((int64_t)item * 128 + m) * g.T + fpt * tp + a) * g.Ho1 * g.Wo1;
            for (int ho = 0; ho < g.Ho1; ho += 2) {                // two rows per round (Ho1 is even): loads of both in flight together
                float v[32];
                float2 o[16];
                tmem_ld16(taddr + a * p.acc_cols + ho * g.P1, v);
                tmem_ld16(taddr + a * p.acc_cols + (ho + 1) * g.P1, v + 16);
                if (p.epi.accum) {
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                        for (int j = 0; j < 16; j += 2)
                            if (j < g.Wo1) o[rr * 8 + j / 2] = *reinterpret_cast<const float2*>(dst + (ho + rr) * g.Wo1 + j);
                }
                tmem_ld_wait();
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                    for (int j = 0; j < 16; j += 2)
                        if (j < g.Wo1) {
                            float2 q = make_float2(v[rr * 16 + j] + bias, v[rr * 16 + j + 1] + bias);
                            if (p.epi.accum) { q.x += o[rr * 8 + j / 2].x; q.y += o[rr * 8 + j / 2].y; }
                            *reinterpret_cast<float2*>(dst + (ho + rr) * g.Wo1 + j) = q;
                        }
            }
        }
    } else {
        const float bias = p.epi.bias ? __ldg(p.epi.bias + m) : 0.f;
        for (int a = 0; a < p.n_acc; ++a) {
            const int video = tile * p.n_acc + a;
            if (video >= p.epi.n_items) break;
            float* dst = out + ((int64_t)video * 128 + m) * g.N2;
            for (int c = 0; c < g.N2; c += 8) {
                float v[8];
                tmem_ld8(taddr + a * p.acc_cols + c, v);
                tmem_ld_wait();
                float4 lo = make_float4(v[0] + bias, v[1] + bias, v[2] + bias, v[3] + bias);
                float4 hi = make_float4(v[4] + bias, v[5] + bias, v[6] + bias, v[7] + bias);
                if (p.epi.accum) {
                    const float4 a4 = *reinterpret_cast<const float4*>(dst + c), b4 = *reinterpret_cast<const float4*>(dst + c + 4);
                    lo.x += a4.x; lo.y += a4.y; lo.z += a4.z; lo.w += a4.w; hi.x += b4.x; hi.y += b4.y; hi.z += b4.z; hi.w += b4.w;
                }
                *reinterpret_cast<float4*>(dst + c) = lo;
                *reinterpret_cast<float4*>(dst + c + 4) = hi;
            }
        }
    }
}

// direct dgrad of conv 1 (Dg1Geo, tc_layout.h): lane m = pw*64 + ci, columns q = a*PD + b hold
// dX1[ci, t, 2a+ph, 2b+pw].  code != NULL: apply the ReLU / MaxPool(1,2,2) routing code of conv 0's output and write
// the 2x2 window of the packed dY of conv 0 (routed gradient at the recorded argmax, zeros elsewhere); code == NULL:
// plain fp32 NCDHW gradient (B, 64, T, H1, H1).
__device__ __forceinline__ void epi_dg1(const WsParams& p, int tile, uint32_t taddr, int m, float* scratch) {
    const Geo& g = p.epi.g;
    const int item = tile / p.tiles_per_item, t = tile % p.tiles_per_item;
    const int pw = m >> 6, ci = m & 63, ph = p.epi.ph;
    if (p.epi.code == nullptr && scratch != nullptr) {
        // Plain fp32 NCDHW gradient.  A lane owns one channel and one column parity, so stores straight from the registers put 32
        // planes under every store instruction and fill half of every sector (measured: 0.59 ms per launch against 0.23 ms of the
        // routed epilogue).  The two warps that hold pw = 0 / 1 of the same 32 channels (q and q + 2) assemble whole rows
        // [32 ci][28 w] in shared memory behind a named barrier of the pair and write them as 8-byte pieces of 112 contiguous bytes.
        const int lane = m & 31, cgrp = (m >> 5) & 1;
        float* sw = scratch + cgrp * (32 * 34);
        const int pair_bar = 2 + cgrp;
        const int tid2 = pw * 32 + lane;
        const int64_t cstride = (int64_t)g.T * g.H1 * g.H1;
        float* gx0 = reinterpret_cast<float*>(p.epi.out) + (((int64_t)item * 64 + cgrp * 32) * g.T + t) * g.H1 * g.H1;
        const int n2 = 32 * g.Wo1;
        for (int a = 0; a < g.Ho1; ++a) {
            float v[16];
            tmem_ld16(taddr + a * g.P1, v);
            tmem_ld_wait();
#pragma unroll
            for (int b = 0; b < 16; ++b)
                if (b < g.Wo1) sw[lane * 34 + 2 * b + pw] = v[b];
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            const int h = 2 * a + ph;
            for (int e = tid2; e < n2; e += 64) {
                const int c = e / g.Wo1, b = e - c * g.Wo1;
                float2 o = *reinterpret_cast<const float2*>(sw + c * 34 + 2 * b);
                float2* d2 = reinterpret_cast<float2*>(gx0 + c * cstride + h * g.H1 + 2 * b);
                if (p.epi.accum) { const float2 x = *d2; o.x += x.x; o.y += x.y; }
                *d2 = o;
            }
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        }
        return;
    }
    const int64_t plane = (((int64_t)item * 64 + ci) * g.T + t) * g.H1 * g.H1;
    const uint8_t* code = p.epi.code ? p.epi.code + plane : nullptr;
    float* gx = reinterpret_cast<float*>(p.epi.out) + plane;
    const BwdGeo& bb = p.epi.bb;
    uint16_t* dy0 = reinterpret_cast<uint16_t*>(p.epi.out) + (int64_t)item * (bb.dy_video / 2);
    const int chunk = ci >> 3, e = ci & 7;
    for (int a = 0; a < g.Ho1; ++a) {
        float v[16];
        tmem_ld16(taddr + a * g.P1, v);
        tmem_ld_wait();
        const int h = 2 * a + ph;
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            if (b >= g.Wo1) continue;
            const int w = 2 * b + pw;
            if (code == nullptr) { gx[h * g.H1 + w] = p.epi.accum ? gx[h * g.H1 + w] + v[b] : v[b]; continue; }
            const uint8_t cd = code[h * g.H1 + w];
            const uint16_t gv = f2bf((cd & 8) ? v[b] : 0.f);
            const int arg = cd & 7;
            if (p.epi.dy0_planar) {
                // padded planar dY of conv 0 (Dg0Geo): [t_pad][chunk 8][row ho+1][col wo+1][8 channels]
                constexpr int PD0 = 64;                              // Dg0Geo::PD
                const int RD0 = g.Ho0 + 4;                           // Dg0Geo::RD
                const int64_t video0 = (int64_t)(g.T + 2) * 8 * RD0 * PD0 * 8;      // bf16 elements per video
                uint16_t* pl = reinterpret_cast<uint16_t*>(p.epi.out) + (int64_t)item * video0 +
                               ((((int64_t)(t + 1) * 8 + chunk) * RD0 + 2 * h + 1) * PD0 + 2 * w + 1) * 8 + e;
                pl[0] = (arg == 0) ? gv : (uint16_t)0;
                pl[8] = (arg == 1) ? gv : (uint16_t)0;
                pl[PD0 * 8] = (arg == 2) ? gv : (uint16_t)0;
                pl[PD0 * 8 + 8] = (arg == 3) ? gv : (uint16_t)0;
                continue;
            }
#pragma unroll
            for (int dh = 0; dh < 2; ++dh) {
                const int pix = (t * bb.Ho + 2 * h + dh) * bb.Wo + 2 * w;           // even: pix and pix+1 share a tile
                const int nt = (int)__umulhi((uint32_t)pix, bb.nc_magic), col = pix - nt * bb.NC;
                uint16_t* d = dy0 + (((int64_t)nt * (bb.K / 8) + chunk) * bb.NC + col) * 8 + e;
                d[0] = (arg == 2 * dh) ? gv : (uint16_t)0;
                d[8] = (arg == 2 * dh + 1) ? gv : (uint16_t)0;
            }
        }
    }
}

// direct dgrad of conv 0 (Dg0Geo): lane m = pixel ((a - a0)*PD + b), accumulator columns n = ci*4 + ph*2 + pw hold
// dX0[ci, t, 2a+ph, 2b+pw]: the gradient w.r.t. the input video, fp32, (B,T,3,H,W) or (B,3,T,H,W).
__device__ __forceinline__ void epi_dg0(const WsParams& p, int tile, uint32_t taddr, int m) {
    const Geo& g = p.epi.g;
    const int item = tile / p.tiles_per_item, sub = tile % p.tiles_per_item;
    const int t = sub / p.v_count, rb = sub % p.v_count;
    const int a = rb * 2 + (m >> 6), b = m & 63;
    float v[16];
    tmem_ld16(taddr, v);
    tmem_ld_wait();
    if (b >= g.Wo0 || a >= g.Ho0) return;
    float* out = reinterpret_cast<float*>(p.epi.out);
    const int64_t HW = (int64_t)g.HW * g.HW;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        const int64_t plane = p.epi.ncdhw ? ((int64_t)item * 3 + ci) * g.T + t : ((int64_t)item * g.T + t) * 3 + ci;
        float* dst = out + plane * HW + (int64_t)(2 * a) * g.HW + 2 * b;
        float2 r0 = make_float2(v[ci * 4 + 0], v[ci * 4 + 1]), r1 = make_float2(v[ci * 4 + 2], v[ci * 4 + 3]);
        if (p.epi.accum) {                                   // out += result: the passes of a split-bf16 dgrad share one fp32 tensor
            const float2 o0 = *reinterpret_cast<const float2*>(dst), o1 = *reinterpret_cast<const float2*>(dst + g.HW);
            r0.x += o0.x; r0.y += o0.y; r1.x += o1.x; r1.y += o1.y;
        }
        *reinterpret_cast<float2*>(dst) = r0;
        *reinterpret_cast<float2*>(dst + g.HW) = r1;
    }
}

// ------------------------------------------------------------------------------------------
template <int EPI, int NACC>
__global__ void __launch_bounds__(EPI == EPI_L0S ? kThreadsL0S : kThreads, 1) ws_gemm_kernel(const __grid_constant__ WsParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte aligned carve-up: [barriers 256 B][weights][pixel ring]
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    Barriers* bars = reinterpret_cast<Barriers*>(base_ptr);
    const uint32_t bar0 = base;
#define BAR(field, i) (bar0 + (uint32_t)offsetof(Barriers, field) + 8u * (uint32_t)(i))
    const uint32_t smem_w = base + p.smem_w_off;
    const uint32_t smem_pix = base + p.smem_pix_off;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(BAR(pix_full, i), 1); mbar_init(BAR(pix_empty, i), 2); }     // empty: one commit per issuer
        for (int i = 0; i < 8; ++i) { mbar_init(BAR(w_full, i), 1); mbar_init(BAR(w_empty, i), 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(BAR(acc_full, i), 2); mbar_init(BAR(acc_empty, i), EPI == EPI_L0S ? 8 : 4); }      // one arrival per epilogue warp
        for (int i = 0; i < 2; ++i) mbar_init(BAR(baton, i), 1);
        mbar_init(BAR(w_res, 0), 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(bar0 + (uint32_t)offsetof(Barriers, tmem_base), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    const int stages_per_tile = p.n_sa * p.n_sb;

    if (warp == 0) {
        // ===================== pixel loader =====================
        if (lane == 0) {
            uint32_t slot = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int tile_pix = tile / p.ug_count;
                const int item = tile_pix / p.tiles_per_item, sub = tile_pix % p.tiles_per_item;
                const int u = sub / p.v_count, v = sub % p.v_count;
                const int64_t slot_item = p.item_index ? __ldg(p.item_index + item) : (int64_t)item;
                const uint8_t* gbase = p.pix + slot_item * p.item_stride + (int64_t)u * p.u_stride + (int64_t)v * p.v_stride;
                for (int sa = 0; sa < p.n_sa; ++sa)
                    for (int sb = 0; sb < p.n_sb; ++sb) {
                        if (p.dbg & 16) mbar_wait<true>(BAR(pix_empty, slot), phase ^ 1); else mbar_wait(BAR(pix_empty, slot), phase ^ 1);
                        if (p.dbg & 1) { mbar_arrive(BAR(pix_full, slot)); }
                        else {
                            mbar_expect_tx(BAR(pix_full, slot), p.stage_bytes);
                            const uint8_t* src = gbase + (int64_t)sa * p.sa_stride + (int64_t)sb * p.sb_stride;
                            const uint32_t dst = smem_pix + slot * p.stage_pitch;
                            for (int c = 0; c < p.n_copies; ++c)
                                bulk_g2s(dst + p.copy_sofs[c], src + p.copy_gofs[c], p.copy_bytes[c], BAR(pix_full, slot));
                        }
                        if (++slot == (uint32_t)p.RP) { slot = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 3) {
        // ===================== weight loader =====================
        if (lane == 0) {
            if (p.w_resident) {
                mbar_expect_tx(BAR(w_res, 0), p.w_bytes);
                for (uint32_t o = 0; o < p.w_bytes; o += 16384) {
                    const uint32_t n = min(16384u, p.w_bytes - o);
                    bulk_g2s(smem_w + o, p.wimg + o, n, BAR(w_res, 0));
                }
            } else {
                uint32_t slot = 0, phase = 0;
                const int slots_per_stage = (p.n_steps + p.G - 1) / p.G;
                const int Gt = p.Gt ? p.Gt : p.G, n_wtiles = p.n_wtiles ? p.n_wtiles : p.n_steps;
                for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                    const int u0 = (tile % p.ug_count) * p.n_u, u1 = min(p.nu_total, u0 + p.n_u);
                    const uint8_t* wbase = p.wimg + (int64_t)((tile / p.ug_count) / p.tiles_per_item) * p.w_item_stride;
                    for (int u = u0; u < u1; ++u)
                        for (int st = 0; st < stages_per_tile; ++st)
                            for (int gi = 0; gi < slots_per_stage; ++gi) {
                                const int s0 = gi * Gt;
                                const uint32_t nb = (uint32_t)min(Gt, n_wtiles - s0) * kWeightTileBytes;
                                if (p.dbg & 16) mbar_wait<true>(BAR(w_empty, slot), phase ^ 1); else mbar_wait(BAR(w_empty, slot), phase ^ 1);
                                if (p.dbg & 2) { mbar_arrive(BAR(w_full, slot)); }
                                else {
                                    mbar_expect_tx(BAR(w_full, slot), nb);
                                    bulk_g2s(smem_w + slot * (uint32_t)Gt * kWeightTileBytes,
                                             wbase + (int64_t)u * p.w_u_stride + ((int64_t)st * n_wtiles + s0) * kWeightTileBytes,
                                             nb, BAR(w_full, slot));
                                }
                                if (++slot == (uint32_t)p.RW) { slot = 0; phase ^= 1; }
                            }
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ===================== MMA issuers (two threads, alternating MMA groups) =====================
        // Measured on B200 (scripts/mma_operand_probe.py, profiles/r01_mma_operand_probe.log): the tensor pipe sustains
        // exactly N/2 cycles per MMA at full-chip scale with moving / misaligned operand windows, but it accepts only about
        // one MMA ahead of the one executing — every cycle the issuing thread spends between two MMAs beyond that slack is
        // a lost pipe cycle (an idle gap of X cycles between stages costs ~X), and a lone thread needs several hundred
        // cycles for the bookkeeping of a stage boundary (barrier waits, commits, ring / loop counters).  Hence:
        //   * every descriptor is precomputed ONCE into shared-memory tables (per pixel-ring slot / weight-ring slot);
        //   * the loop nest (tile, accumulator group, stage, weight slot) is flattened into a generator of MMA groups;
        //   * TWO threads (lane 0 of warps 1 and 2) walk the same group sequence and issue alternate groups.  While one
        //     issues, the other does its bookkeeping and waits for its next group's operands; a baton mbarrier passes
        //     the right to issue, so the MMAs still enter the pipe in exactly the sequential order (bitwise
        //     reproducible accumulation) and the hand-over hides behind the MMAs already queued.
        //   * tcgen05.commit only tracks the MMAs of the executing thread: pix_empty / acc_full expect one commit from
        //     EACH issuer (the one that did not issue the closing group commits as it walks past it).
        const int role = warp - 1;
        const bool RESIDENT = p.w_resident != 0;
        const int n_steps = p.n_steps;
        const int G = RESIDENT ? n_steps : p.G;
        const uint32_t a_hi = p.a_hi ? p.a_hi : ((p.a_sbo16 & 0x3FFFu) | (1u << 14));
        const uint32_t b_hi = p.b_hi ? p.b_hi : (8u | (1u << 14));
        const uint32_t a_lbo_bits = (p.a_lbo16 & 0x3FFFu) << 16;
        const uint32_t slot_bytes = (uint32_t)(p.Gt ? p.Gt : G) * kWeightTileBytes;
        uint64_t* tabB = reinterpret_cast<uint64_t*>(base_ptr + kBarBlock);                   // [RP][n_steps][NACC]
        uint64_t* tabA = tabB + p.RP * n_steps * NACC;                                  // [RW][G] or [n_sa][n_steps]
        const int t64 = role * 32 + lane;
        for (int i = t64; i < p.RP * n_steps * NACC; i += 64) {
            const int a = i % NACC, s = (i / NACC) % n_steps, slot = i / (NACC * n_steps);
            const uint32_t pix16 = (smem_pix + (uint32_t)slot * p.stage_pitch) >> 4;
            tabB[i] = ((uint64_t)b_hi << 32) | (p.step_tab[s].y + pix16 + (uint32_t)a * p.acc_delta16);
        }
        const int nA = RESIDENT ? (p.n_wsets ? p.n_wsets : p.n_sa) * n_steps : p.RW * G;
        for (int i = t64; i < nA; i += 64) {
            uint32_t a16;
            if (RESIDENT) a16 = (smem_w >> 4) + p.step_tab[i % n_steps].x + (uint32_t)((i / n_steps) * p.a_sa_stride16);
            else a16 = ((smem_w + (uint32_t)(i / G) * slot_bytes) >> 4) + (uint32_t)(p.Gt ? p.a_in_group[i % G] : (i % G)) * (kWeightTileBytes >> 4);
            tabA[i] = ((uint64_t)a_hi << 32) | ((a16 & 0x3FFFu) | a_lbo_bits);
        }
        asm volatile("bar.sync 1, 64;" ::: "memory");                                  // tables visible to both issuers
        // elect.sync (not `lane == 0`): the compiler then knows that exactly one thread runs the block, so descriptors move
        // to uniform registers with plain R2UR instead of a per-MMA ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop
        if (elect_one()) {
            struct Group {
                uint32_t w_acc, p_acc, w_pix, p_pix, w_w, p_w;      // barriers to wait for (0 = none) and their parities
                uint32_t c_w, c_pix, c_acc;                           // barriers to commit after the MMAs (0 = none)
                const uint64_t* ta;
                const uint64_t* tb;
                int nst;
                uint32_t d_base, acc0;
                // split-fp16 conv 0 only: up to two more (weight window, accumulator) segments issued in the same group over the
                // same pixel stage, a second accumulator to wait for and a second accumulator to commit
                // (scalar fields on purpose: arrays indexed at run time would move the whole record to local memory)
                int nseg;
                const uint64_t* ta_b;
                const uint64_t* ta_c;
                uint32_t d_b, d_c, acc0_b, acc0_c;
                uint32_t w_acc2, p_acc2, c_acc2;
            };
            const int slots_per_stage = (n_steps + G - 1) / G;
            const int n_tiles = p.n_tiles, n_sb = p.n_sb;
            const uint32_t acc_cols = p.acc_cols, idesc = p.idesc, RP = (uint32_t)p.RP, RW = (uint32_t)p.RW, acc_stages = p.acc_stages;
            const bool swap_ab = p.swap_ab != 0;
            // generator state (identical in both issuers)
            int tile = blockIdx.x, u = 0, st = 0, sa = 0, sb = 0, g = 0, j = 0;
            int n_u_eff = tile < n_tiles ? min(p.n_u, p.nu_total - (tile % p.ug_count) * p.n_u) : 0;
            uint32_t pslot = 0, pphase = 0, wslot = 0, wphase = 0, as = 0, aphase = 0;
            auto next = [&](Group& r) {
                const bool first = (st == 0 && g == 0);                  // first group of this (tile, u): fresh accumulator
                const bool last_g = (g == slots_per_stage - 1);
                const bool last_u = (u == n_u_eff - 1);
                r.w_acc = BAR(acc_empty, as);                            r.p_acc = aphase ^ 1u;
                r.w_pix = BAR(pix_full, pslot);                          r.p_pix = pphase;
                r.w_w = RESIDENT ? 0u : BAR(w_full, wslot);              r.p_w = wphase;
                r.nst = min(G, n_steps - j);
                r.ta = RESIDENT ? tabA + sa * n_steps + j : tabA + wslot * G;
                r.tb = tabB + ((int)pslot * n_steps + j) * NACC;
                r.d_base = tmem_base + as * (acc_cols * (uint32_t)NACC);
                r.acc0 = first ? 0u : 1u;
                r.c_w = RESIDENT ? 0u : BAR(w_empty, wslot);
                r.c_pix = (last_g && last_u) ? BAR(pix_empty, pslot) : 0u;
                r.c_acc = (last_g && st == stages_per_tile - 1) ? BAR(acc_full, as) : 0u;
                // advance
                if (!RESIDENT) { if (++wslot == RW) { wslot = 0; wphase ^= 1; } }
                if (!last_g) { ++g; j += r.nst; return; }
                g = 0; j = 0;
                if (last_u) { if (++pslot == RP) { pslot = 0; pphase ^= 1; } }
                if (++sb == n_sb) { sb = 0; ++sa; }
                if (++st < stages_per_tile) return;
                st = 0; sa = 0; sb = 0;
                if (++as == acc_stages) { as = 0; aphase ^= 1; }
                if (++u < n_u_eff) return;
                u = 0; tile += gridDim.x;
                n_u_eff = tile < n_tiles ? min(p.n_u, p.nu_total - (tile % p.ug_count) * p.n_u) : 0;
            };
            // conv 0, input-frame streaming: frame i of the column (pair pp = i / 2) feeds
            //   even i: X  pair pp-1 (frame 2pp-1, kt=2) through window 3 = [0 ; W2]   (absent for pp = 0; closes pair pp-1)
            //           Y  pair pp   (kt = 1, 0)         through window 1 = [W1 ; W0]   (first write of pair 0)
            //   odd  i: Z  pair pp   (kt = 2, 1)         through window 2 = [W2 ; W1]   (closes the last pair)
            //           V  pair pp+1 (frame 2pp+2, kt=0) through window 0 = [W0 ; 0]    (absent for the last pair; first write)
            // Pair q (running count over the columns of this CTA) lives in accumulator stage q & 1.
            const int pairs = p.stream_pairs;
            int fi = 0, part = 0;
            uint32_t qbase = 0;
            auto next_stream = [&](Group& r) {
                const int pp = fi >> 1;
                const bool odd = fi & 1;
                const int kind = odd ? (part == 0 ? 2 : 3) : ((part == 0 && pp > 0) ? 0 : 1);       // 0 X, 1 Y, 2 Z, 3 V
                const uint32_t q = qbase + (uint32_t)(pp + (kind == 0 ? -1 : kind == 3 ? 1 : 0));
                const uint32_t buf = q & 1u;
                const bool first_write = (kind == 3) || (kind == 1 && pp == 0);
                const bool final_write = (kind == 0) || (kind == 2 && pp == pairs - 1);
                const bool last_of_frame = odd ? (kind == 3 || pp + 1 >= pairs) : (kind == 1);
                r.w_acc = first_write ? BAR(acc_empty, buf) : 0u;        r.p_acc = ((q >> 1) & 1u) ^ 1u;
                r.w_pix = BAR(pix_full, pslot);                          r.p_pix = pphase;
                r.w_w = 0u;                                              r.p_w = 0u;
                r.nst = n_steps;
                r.ta = tabA + (kind == 0 ? 3 : kind == 1 ? 1 : kind == 2 ? 2 : 0) * n_steps;
                r.tb = tabB + (int)pslot * n_steps * NACC;
                r.d_base = tmem_base + buf * (acc_cols * (uint32_t)NACC);
                r.acc0 = first_write ? 0u : 1u;
                r.c_w = 0u;
                r.c_pix = last_of_frame ? BAR(pix_empty, pslot) : 0u;
                r.c_acc = final_write ? BAR(acc_full, buf) : 0u;
                if (!last_of_frame) { part = 1; return; }
                part = 0;
                if (++pslot == RP) { pslot = 0; pphase ^= 1; }
                if (++fi < 2 * pairs) return;
                fi = 0; qbase += (uint32_t)pairs; tile += gridDim.x;
            };
            // conv 0, split-fp16 operands (SGeo, tc_layout.h): stage s = (input frame i = s >> 1, part = s & 1); the stage feeds
            // output frames f = i + 1 - kt through the weight window kt = 2, 1, 0 (where f exists).  Frame f (running count
            // qbase + f over the columns of this CTA) lives in accumulator buffer (qbase + f) & 3; its first write is
            // (i = f - 1, hi part, kt = 0) — for f = 0: (i = 0, hi, kt = 1) — and its last (i = f + 1, lo part, kt = 2) — for
            // f = T - 1: (i = T - 1, lo, kt = 1).
            const int nparts = p.l0s_parts == 1 ? 1 : 2;
            int ss = 0;
            auto next_l0s = [&](Group& r) {
                // ONE group per stage (up to 3 x n_steps MMAs): a hand-over between the two issuers costs ~300 cycles, more than the
                // two queued N = 112 MMAs (2 x 56 cycles) hide, so groups of 11 short MMAs left the pipe 38 % idle
                const int T = pairs;
                const int i = nparts == 2 ? ss >> 1 : ss, part = nparts == 2 ? (ss & 1) : 0;
                r.nseg = 0;
                r.w_acc = 0u; r.w_acc2 = 0u; r.c_acc = 0u; r.c_acc2 = 0u; r.p_acc = 0u; r.p_acc2 = 0u;
                for (int kt = 2; kt >= 0; --kt) {
                    const int f = i + 1 - kt;
                    if (f < 0 || f >= T) continue;
                    const uint32_t q = qbase + (uint32_t)f;
                    const uint32_t buf = q & 3u;
                    const bool first_write = part == 0 && (kt == 0 || (f == 0 && kt == 1));
                    const bool final_write = part == nparts - 1 && (kt == 2 || (f == T - 1 && kt == 1));
                    const uint64_t* ta = tabA + kt * n_steps;
                    const uint32_t d = tmem_base + buf * (acc_cols * (uint32_t)NACC);
                    if (r.nseg == 0) { r.ta = ta; r.d_base = d; r.acc0 = first_write ? 0u : 1u; }
                    else if (r.nseg == 1) { r.ta_b = ta; r.d_b = d; r.acc0_b = first_write ? 0u : 1u; }
                    else { r.ta_c = ta; r.d_c = d; r.acc0_c = first_write ? 0u : 1u; }
                    ++r.nseg;
                    if (first_write) {
                        if (!r.w_acc) { r.w_acc = BAR(acc_empty, buf); r.p_acc = ((q >> 2) & 1u) ^ 1u; }
                        else { r.w_acc2 = BAR(acc_empty, buf); r.p_acc2 = ((q >> 2) & 1u) ^ 1u; }
                    }
                    if (final_write) {
                        if (!r.c_acc) r.c_acc = BAR(acc_full, buf); else r.c_acc2 = BAR(acc_full, buf);
                    }
                }
                r.w_pix = BAR(pix_full, pslot);                          r.p_pix = pphase;
                r.w_w = 0u;                                              r.p_w = 0u;
                r.nst = n_steps;
                r.tb = tabB + (int)pslot * n_steps * NACC;
                r.c_w = 0u;
                r.c_pix = BAR(pix_empty, pslot);
                if (++pslot == RP) { pslot = 0; pphase ^= 1; }
                if (++ss < nparts * T) return;
                ss = 0; qbase += (uint32_t)T; tile += gridDim.x;
            };
            long long c_acc = 0, c_pix = 0, c_w = 0, c_baton = 0, c_issue = 0;
            const bool prof = p.prof != nullptr;
            if (RESIDENT) { mbar_wait(BAR(w_res, 0), 0); tc_fence_after(); }
            const long long c_begin = clock64();
            Group r;
            r.nseg = 1; r.w_acc2 = 0u; r.p_acc2 = 0u; r.c_acc2 = 0u;
            for (uint32_t k = 0; tile < n_tiles; ++k) {
                if (EPI == EPI_L0S) next_l0s(r); else if (pairs) next_stream(r); else next(r);
                if ((k & 1u) == (uint32_t)role) {
                    // ---- my group: operands (usually long there), then the baton of the previous group's issuer
                    long long t0 = prof ? clock64() : 0;
                    // (split-fp16 conv 0: deferring the wait for the accumulator that the LAST segment initialises to just before that
                    // segment was measured and is slower — a try_wait + fence in the middle of a group idles the pipe for longer
                    // than the wait it saves)
                    if (r.w_acc) mbar_wait(r.w_acc, r.p_acc);   // (an already completed phase returns at once)
                    if (EPI == EPI_L0S) { if (r.w_acc2) mbar_wait(r.w_acc2, r.p_acc2); }
                    if (prof) { const long long t1 = clock64(); c_acc += t1 - t0; t0 = t1; }
                    mbar_wait(r.w_pix, r.p_pix);
                    if (prof) { const long long t1 = clock64(); c_pix += t1 - t0; t0 = t1; }
                    if (r.w_w) mbar_wait(r.w_w, r.p_w);
                    if (prof) { const long long t1 = clock64(); c_w += t1 - t0; t0 = t1; }
                    if (k > 0) mbar_wait(BAR(baton, role ^ 1), ((k - 1u) >> 1) & 1u);
                    if (prof) { const long long t1 = clock64(); c_baton += t1 - t0; t0 = t1; }
                    tc_fence_after();
                    const uint64_t* ta = r.ta;
                    const uint64_t* tb = r.tb;
                    const uint32_t d_base = r.d_base;
                    const int nst = r.nst;
                    uint32_t accumulate = r.acc0;
                    if (swap_ab) {                              // pixels = M operand, weight tile = N operand (NACC == 1)
#pragma unroll 4
                        for (int jj = 0; jj < nst; ++jj) {
                            umma_bf16(d_base, tb[jj * NACC], ta[jj], idesc, accumulate);
                            accumulate = 1;
                        }
                    } else {
#pragma unroll 4
                        for (int jj = 0; jj < nst; ++jj) {
                            const uint64_t a_desc = ta[jj];
#pragma unroll
                            for (int a = 0; a < NACC; ++a)
                                umma_bf16(d_base + (uint32_t)a * acc_cols, a_desc, tb[jj * NACC + a], idesc, accumulate);
                            accumulate = 1;
                        }
                        if (EPI == EPI_L0S) {                       // split-fp16 conv 0: further weight windows / accumulators, same stage
                            if (r.nseg > 1) {
                                const uint64_t* ta_s = r.ta_b;
                                const uint32_t d_s = r.d_b;
                                accumulate = r.acc0_b;
#pragma unroll 4
                                for (int jj = 0; jj < nst; ++jj) {
                                    umma_bf16(d_s, ta_s[jj], tb[jj * NACC], idesc, accumulate);
                                    accumulate = 1;
                                }
                            }
                            if (r.nseg > 2) {
                                const uint64_t* ta_s = r.ta_c;
                                const uint32_t d_s = r.d_c;
                                accumulate = r.acc0_c;
#pragma unroll 4
                                for (int jj = 0; jj < nst; ++jj) {
                                    umma_bf16(d_s, ta_s[jj], tb[jj * NACC], idesc, accumulate);
                                    accumulate = 1;
                                }
                            }
                        }
                    }
                    mbar_arrive(BAR(baton, role));            // the other issuer may enter the pipe behind these MMAs
                    if (r.c_w) umma_commit(r.c_w);
                    if (prof) c_issue += clock64() - t0;
                }
                if (r.c_pix) umma_commit(r.c_pix);            // both issuers: "MY MMAs that read this stage / wrote this
                if (r.c_acc) umma_commit(r.c_acc);            //  accumulator have completed"
                if (EPI == EPI_L0S) { if (r.c_acc2) umma_commit(r.c_acc2); }
            }
            if (prof && role == 0) {
                long long* o = p.prof + (int64_t)blockIdx.x * 8;
                o[0] = clock64() - c_begin; o[1] = c_acc; o[2] = c_pix; o[3] = c_w; o[4] = c_issue + c_baton;      // o[5..7]: epilogue warp 0
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (128 threads, lane quarter = warp % 4) =====================
        // EPI_L0S runs TWO such groups (warps 4..7 and 8..11), each draining half of the columns of EVERY accumulator: a drain is a
        // chain of dependent TMEM loads / shuffles / conversions whose latencies one warp per scheduler cannot hide (measured:
        // 2700 + 900 cycles per accumulator against 1850 cycles of MMAs per stage in the two-product mode), and a drained
        // buffer is needed again one stage later (4 buffers, 3 frames in flight) — so both the throughput (two warps per
        // scheduler) and the latency (half the columns) of a drain must drop below a stage.
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int egroup = (warp - 4) >> 2;
        uint8_t* stash_ptr = base_ptr + p.smem_stash_off + (EPI == EPI_L0S ? (uint32_t)egroup * p.stash_group_bytes : 0u);
        float bias_reg = 0.f;                                            // per-lane bias of the fused epilogues
        if (EPI == EPI_L0S) bias_reg = __ldg(p.epi.bias + (m >> 5) * 16 + (lane & 15));
        else if (EPI == EPI_L0) bias_reg = __ldg(p.epi.bias + (m & 63));
        else if (EPI == EPI_L1 || EPI == EPI_L1S || EPI == EPI_L2) bias_reg = __ldg(p.epi.bias + m);
        const int l0s_split = ((p.epi.g.Wo0 / 8 + 1) / 2) * 8;           // group 0: columns [0, split), group 1: [split, Wo0)
        const int l0s_c0 = egroup ? l0s_split : 0, l0s_c1 = egroup ? p.epi.g.Wo0 : l0s_split;
        uint32_t as = 0, aphase = 0;
        const bool eprof = p.prof != nullptr;
        long long e_wait = 0, e_drain = 0, e_store = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int tile_pix = tile / p.ug_count, u0 = (tile % p.ug_count) * p.n_u;
            // conv 0 streaming: a tile is a column (video, row band) and yields stream_pairs accumulators, pair by pair
            const int n_u_eff = p.stream_pairs ? p.stream_pairs : min(p.n_u, p.nu_total - u0);
            // split-fp16 conv 0: everything that depends on the tile only (two integer divisions and the store offsets cost a
            // lone warp several hundred exposed cycles — per tile, not per frame)
            int l0s_item = 0, l0s_rb = 0;
            L0sStorePlan l0s_plan;
            uint8_t* l0s_vbase = nullptr;
            if (EPI == EPI_L0S) {
                l0s_item = tile / p.tiles_per_item; l0s_rb = tile - l0s_item * p.tiles_per_item;
                epi_l0s_store_plan(p, l0s_rb, q, lane, l0s_c0, l0s_c1, l0s_plan);
                l0s_vbase = p.epi.out + (int64_t)l0s_item * (8 * p.epi.g.slice1) + p.epi.g.frame1;
            }
            for (int u = 0; u < n_u_eff; ++u) {
                long long e0 = eprof ? clock64() : 0;
                if (p.dbg & 32) mbar_wait(BAR(acc_full, as), aphase);
                else if (EPI == EPI_L0S && !(p.dbg & 64)) mbar_wait_sleep<32>(BAR(acc_full, as), aphase);     // a drained buffer is needed again one stage later
                else mbar_wait<true>(BAR(acc_full, as), aphase);
                tc_fence_after();
                if (eprof) { const long long e1 = clock64(); e_wait += e1 - e0; e0 = e1; }
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * (p.acc_cols * (uint32_t)p.n_acc);
                const bool work = !(p.dbg & 4);
                if (work) {
                    if (EPI == EPI_RAW) epi_raw(p, (int64_t)tile_pix * p.nu_total + u0 + u, taddr, m, p.smem_epi_off ? reinterpret_cast<float*>(base_ptr + p.smem_epi_off) : nullptr);
                    else if (EPI == EPI_L0) {
                        const int item = tile / p.tiles_per_item, sub = tile % p.tiles_per_item;
                        const int tp = p.stream_pairs ? u : sub / p.v_count, rb = p.stream_pairs ? sub : sub % p.v_count;
                        if (p.smem_stash_off) epi_l0_drain(p, item, tp, rb, taddr, m, reinterpret_cast<uint16_t*>(base_ptr + p.smem_stash_off), bias_reg);
                        else epi_l0(p, item, tp, rb, taddr, m, bias_reg);
                    }
                    else if (EPI == EPI_L0S) epi_l0s_drain(p, l0s_item, u, l0s_rb, taddr, m, reinterpret_cast<uint16_t*>(stash_ptr), l0s_c0, l0s_c1, bias_reg);
                    else if (EPI == EPI_L1S) epi_l1s_drain(p, tile, taddr, m, reinterpret_cast<uint16_t*>(base_ptr + p.smem_stash_off), bias_reg);
                    else if (EPI == EPI_L1) epi_l1_drain(p, tile, taddr, m, reinterpret_cast<uint16_t*>(base_ptr + p.smem_stash_off), bias_reg);
                    else if (EPI == EPI_PLAIN) epi_plain(p, tile, taddr, m);
                    else if (EPI == EPI_DG1) epi_dg1(p, tile, taddr, m, p.smem_epi_off ? reinterpret_cast<float*>(base_ptr + p.smem_epi_off) : nullptr);
                    else if (EPI == EPI_DG0) epi_dg0(p, tile, taddr, m);
                    else epi_l2(p, tile, taddr, m, bias_reg);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(acc_empty, as));
                if (eprof) { const long long e1 = clock64(); e_drain += e1 - e0; e0 = e1; }
                if (EPI == EPI_L1 && work) {
                    // the accumulator is free again: scatter the stashed tile while the next one is computed
                    epi_l1_store(p, tile, q, lane, reinterpret_cast<const uint16_t*>(base_ptr + p.smem_stash_off));
                    __syncwarp();       // the stash rows of this warp are rewritten by the next drain
                }
                if (EPI == EPI_L1S && work) {
                    epi_l1s_store(p, tile, q, lane, reinterpret_cast<const uint16_t*>(base_ptr + p.smem_stash_off));
                    __syncwarp();
                }
                if (EPI == EPI_L0S && work) {
                    epi_l0s_store(l0s_vbase + (int64_t)u * p.epi.g.frame1, reinterpret_cast<const uint16_t*>(stash_ptr), l0s_plan);
                    __syncwarp();
                }
                if (EPI == EPI_L0 && work && p.smem_stash_off) {
                    const int item = tile / p.tiles_per_item, sub = tile % p.tiles_per_item;
                    const int tp = p.stream_pairs ? u : sub / p.v_count, rb = p.stream_pairs ? sub : sub % p.v_count;
                    epi_l0_store(p, item, tp, rb, q, lane, reinterpret_cast<const uint16_t*>(base_ptr + p.smem_stash_off));
                    __syncwarp();
                }
                if (eprof) e_store += clock64() - e0;
                if (++as == p.acc_stages) { as = 0; aphase ^= 1; }
            }
        }
        if (eprof && warp == 4 && lane == 0) {          // epilogue warp 0: cycles waiting for acc_full / draining / storing
            long long* o = p.prof + (int64_t)blockIdx.x * 8;
            o[5] = e_wait; o[6] = e_drain; o[7] = e_store;
        }
    }
    // ===================== teardown =====================
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
#undef BAR
}

// ------------------------------------------------------------------------------------------
// host side: per-layer parameter construction
// ------------------------------------------------------------------------------------------
static long long* g_prof = nullptr;      // set by vd_tc_set_profile_buffer (tuning only)

static uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

// tuning override (integer environment variable), used by the ring-depth experiments in profiles/
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

static int finalize_smem(WsParams& p, uint32_t w_region, uint32_t* smem_total, bool epi_scratch = false, uint32_t stash_bytes = 0) {
    {
        const int G = p.w_resident ? p.n_steps : p.G;
        for (int j = 0; j < p.n_steps; ++j) {
            p.step_tab[j].x = p.w_resident ? p.a_off16[j] : (uint32_t)(j % G) * (kWeightTileBytes >> 4);
            p.step_tab[j].y = p.b_off16[j] | ((p.b_lbo16[j] & 0x3FFFu) << 16);
        }
    }
    {
        const int G = p.w_resident ? p.n_steps : p.G;
        const uint32_t nB = (uint32_t)p.RP * p.n_steps * p.n_acc;
        const uint32_t nA = p.w_resident ? (uint32_t)(p.n_wsets ? p.n_wsets : p.n_sa) * p.n_steps : (uint32_t)p.RW * G;
        p.tab_bytes = align_up((nB + nA) * 8, 128);
        if (p.tab_bytes < 3840) p.tab_bytes = 3840;     // (the r01 carve-up: 4352 bytes ahead of the weights)
    }
    p.smem_w_off = kBarBlock + p.tab_bytes;
    p.smem_pix_off = align_up(p.smem_w_off + w_region, 128);
    p.stage_pitch = align_up(p.stage_bytes + 16, 128);
    uint32_t total = p.smem_pix_off + (uint32_t)p.RP * p.stage_pitch;
    p.smem_epi_off = 0;
    if (epi_scratch) { p.smem_epi_off = align_up(total, 128); total = p.smem_epi_off + 4 * 32 * 33 * 4; }
    p.smem_stash_off = 0;
    if (stash_bytes) { p.smem_stash_off = align_up(total, 128); total = p.smem_stash_off + stash_bytes; }
    total += 1024;                                  // manual 1024-byte alignment slack
    if (total < 120 * 1024) total = 120 * 1024;     // one CTA per SM: every CTA allocates all 512 TMEM columns
    VD_REQUIRE(total <= 232448, "tc conv: shared memory budget exceeded (%u bytes)", total);
    *smem_total = total;
    return 0;
}

static int setup_l0(WsParams& p, const Geo& g, int B, uint32_t* smem, bool fused = false) {
    p.n_u = 1; p.nu_total = 1; p.ug_count = 1; p.w_u_stride = 0;
    p.n_tiles = B * (g.T / 2) * (g.Ho0 / g.R0);
    p.tiles_per_item = (g.T / 2) * (g.Ho0 / g.R0);
    p.v_count = g.Ho0 / g.R0;
    p.item_stride = g.video0; p.u_stride = 2 * g.frame0; p.v_stride = (int64_t)g.R0 * g.Wo0 * 16;
    p.n_sa = 4; p.n_sb = 1; p.sa_stride = g.frame0; p.sb_stride = 0;
    p.n_copies = 6;
    uint32_t sofs = 0;
    uint32_t blk[3][2];
    for (int c = 0; c < 3; ++c)
        for (int par = 0; par < 2; ++par) {
            const int i = c * 2 + par;
            p.copy_gofs[i] = (int64_t)i * g.plane0;
            p.copy_sofs[i] = sofs;
            p.copy_bytes[i] = (uint32_t)(g.R0 + 2 + par) * g.Wo0 * 16;
            blk[c][par] = sofs;
            sofs += p.copy_bytes[i];
        }
    p.stage_bytes = sofs;
    // 21 chunks (c,kh) paired into 11 K=16 steps; chunk order = c*7 + idx, kh = l0_chunk_kh(idx)
    p.n_steps = kW0Steps;
    for (int s = 0; s < kW0Steps; ++s) {
        const int c0 = 2 * s, c1 = (2 * s + 1 < 21) ? 2 * s + 1 : 2 * s;    // last step: 2nd half has zero weights
        auto off = [&](int ch) { int c = ch / 7, kh = l0_chunk_kh(ch % 7); return blk[c][tap_par(kh)] + (uint32_t)tap_shift(kh) * g.Wo0 * 16; };
        const uint32_t o0 = off(c0), o1 = off(c1);
        if (o1 >= o0) { p.b_off16[s] = o0 >> 4; p.b_lbo16[s] = (o1 - o0) >> 4; }
        else { return -2; }   // pairing must be monotone (host packer uses the same order)
        p.a_off16[s] = (uint32_t)(s * 10240 + 3 * 1024) >> 4;
    }
    p.a_sa_stride16 = -(1024 >> 4);
    p.a_lbo16 = 5120 >> 4; p.a_sbo16 = 8;
    p.w_resident = 1; p.w_bytes = kW0Bytes;
    p.G = 1; p.RW = 1;
    // fused forward: two-phase epilogue (bf16 stash of the pooled tile, 2 x npos x kStash0Pitch) and a 2-slot pixel ring —
    // in the streaming order a stage lasts 22 MMAs, so one slot of prefetch is enough
    const bool stash = fused && env_int("VD_TC_L0_STASH", 1);
    p.RP = env_int("VD_TC_L0_RP", stash ? 2 : 3);
    p.n_acc = 1; p.acc_delta16 = 0;
    p.ncols = g.N0; p.acc_cols = 256; p.acc_stages = 2;
    p.idesc = umma_idesc_bf16(128, g.N0);
    return finalize_smem(p, kW0Bytes, smem, false, stash ? (uint32_t)(2 * (g.R0 / 2) * (g.Wo0 / 2)) * kStash0Pitch * 2 : 0u);
}

static int setup_l1(WsParams& p, const Geo& g, int B, uint32_t* smem) {
    p.n_u = 1; p.nu_total = 1; p.ug_count = 1; p.w_u_stride = 0;
    p.n_tiles = B * (g.T / 2);
    p.tiles_per_item = g.T / 2; p.v_count = 1;
    p.item_stride = g.video1; p.u_stride = 2 * g.frame1; p.v_stride = 0;
    p.n_sa = 3; p.n_sb = 4; p.sa_stride = g.frame1; p.sb_stride = g.slice1;
    p.n_copies = 1; p.copy_gofs[0] = 0; p.copy_sofs[0] = 0; p.copy_bytes[0] = (uint32_t)(2 * g.frame1);
    p.stage_bytes = (uint32_t)(2 * g.frame1);
    p.n_steps = 49;
    for (int kh = 0; kh < 7; ++kh)
        for (int kw = 0; kw < 7; ++kw) {
            const int plane = tap_par(kh) * 2 + tap_par(kw);
            const uint32_t o = (uint32_t)(plane * 2) * (uint32_t)g.plane1 + (uint32_t)(tap_shift(kh) * g.P1 + tap_shift(kw)) * 16;
            p.b_off16[kh * 7 + kw] = o >> 4;
            p.b_lbo16[kh * 7 + kw] = (uint32_t)g.plane1 >> 4;
        }
    p.a_lbo16 = 2048 >> 4; p.a_sbo16 = 8;
    p.w_resident = 0; p.w_bytes = 0;
    // weight ring: 3 slots of 5 K-steps (20 KB each).  With the dual-issuer pipe a slot must be refilled within the time the other
    // slots last; 2 x 7 steps left the MMA threads waiting 28 % of the time, 3 x 5 measures 5.5 % faster (scripts/tune_rings.py)
    p.G = env_int("VD_TC_L1_G", 5); p.RW = env_int("VD_TC_L1_RW", 3); p.RP = 2;
    p.n_acc = 2; p.acc_delta16 = (uint32_t)g.frame1 >> 4;
    p.ncols = g.N1; p.acc_cols = 256; p.acc_stages = 1;
    p.idesc = umma_idesc_bf16(128, g.N1);
    return finalize_smem(p, (uint32_t)p.G * p.RW * kWeightTileBytes, smem, false, (uint32_t)(g.H2 * g.H2) * kStashPitch * 2);
}

static int setup_l2(WsParams& p, const Geo& g, int B, uint32_t* smem) {
    p.n_u = 1; p.nu_total = 1; p.ug_count = 1; p.w_u_stride = 0;
    const int VPT = kVideosPerTile2;
    p.n_tiles = (B + VPT - 1) / VPT;
    p.tiles_per_item = 1; p.v_count = 1;
    p.item_stride = VPT * g.video2; p.u_stride = 0; p.v_stride = 0;
    p.n_sa = 49; p.n_sb = 2; p.sa_stride = 2 * g.group2; p.sb_stride = g.group2;
    p.n_copies = VPT;
    for (int v = 0; v < VPT; ++v) {
        p.copy_gofs[v] = (int64_t)v * g.video2;
        p.copy_sofs[v] = (uint32_t)(v * g.group2);
        p.copy_bytes[v] = (uint32_t)g.group2;
    }
    p.stage_bytes = (uint32_t)(VPT * g.group2);
    p.n_steps = 12;
    for (int kt = 0; kt < 3; ++kt)
        for (int kc = 0; kc < 4; ++kc) {
            p.b_off16[kt * 4 + kc] = (uint32_t)((2 * kc) * g.chunk2 + (int64_t)kt * g.HW2 * 16) >> 4;
            p.b_lbo16[kt * 4 + kc] = (uint32_t)g.chunk2 >> 4;
        }
    p.a_lbo16 = 2048 >> 4; p.a_sbo16 = 8;
    p.w_resident = 0; p.w_bytes = 0;
    p.G = env_int("VD_TC_L2_G", 6); p.RW = env_int("VD_TC_L2_RW", 2); p.RP = 2;
    p.n_acc = VPT; p.acc_delta16 = (uint32_t)g.group2 >> 4;
    p.ncols = g.N2; p.acc_cols = 128; p.acc_stages = 1;
    p.idesc = umma_idesc_bf16(128, g.N2);
    return finalize_smem(p, (uint32_t)p.G * p.RW * kWeightTileBytes, smem);
}

// ---- split-fp16 forward (SGeo, tc_layout.h) ----
// conv 0: column tiles (video, band of R0s rows); stages (frame i, part) of the X0s operand; resident M-stacked weight
// image [kt 3][step 11] x 4 KiB; 4 accumulators of 128 columns (output frames f & 3).
//
// passes = 3 (default): all products of the hi / lo pairs.  passes = 2 ("two-product" mode of the frozen real videos):
// the activations are carried as ONE fp16 value (xh) and only the weights as a pair, y = xh*wh + xh*wl — the weights are
// exact, and the single rounding of each activation is independent from video to video, so it averages out of the class
// means the DM loss reads (distill_s2d_ms.py:419-422).  The operand is then the hi-only layout X0h = X0 of tc_layout.h
// holding fp16 values (6 planes per frame), a frame is ONE stage, and the epilogues write the hi part only.
static int setup_l0s(WsParams& p, const Geo& g, int B, uint32_t* smem, int passes = 3) {
    const SGeo sg = make_sgeo(g);
    const bool two = passes == 2;
    p.n_u = 1; p.nu_total = 1; p.ug_count = 1; p.w_u_stride = 0;
    p.n_tiles = B * sg.nrb0s; p.tiles_per_item = sg.nrb0s; p.v_count = sg.nrb0s;
    p.item_stride = two ? g.video0 : sg.video0s; p.u_stride = 0; p.v_stride = (int64_t)sg.R0s * g.Wo0 * 16;
    p.n_sa = g.T; p.n_sb = two ? 1 : 2; p.sa_stride = two ? g.frame0 : sg.frame0s; p.sb_stride = 6 * g.plane0;
    p.l0s_parts = two ? 1 : 2;
    p.n_copies = 6;
    uint32_t sofs = 0;
    uint32_t blk[3][2];
    for (int c = 0; c < 3; ++c)
        for (int par = 0; par < 2; ++par) {
            const int i = c * 2 + par;
            p.copy_gofs[i] = (two ? g.frame0 : sg.frame0s) + (int64_t)i * g.plane0;        // frame i of the video is t_pad = i + 1
            p.copy_sofs[i] = sofs;
            p.copy_bytes[i] = (uint32_t)(sg.R0s + 2 + par) * g.Wo0 * 16;
            blk[c][par] = sofs;
            sofs += p.copy_bytes[i];
        }
    p.stage_bytes = sofs;
    p.n_steps = kW0Steps;
    for (int s = 0; s < kW0Steps; ++s) {
        const int c0 = 2 * s, c1 = (2 * s + 1 < 21) ? 2 * s + 1 : 2 * s;    // last step: 2nd half has zero weights
        auto off = [&](int ch) { int c = ch / 7, kh = l0_chunk_kh(ch % 7); return blk[c][tap_par(kh)] + (uint32_t)tap_shift(kh) * g.Wo0 * 16; };
        const uint32_t o0 = off(c0), o1 = off(c1);
        if (o1 < o0) return -2;
        p.b_off16[s] = o0 >> 4; p.b_lbo16[s] = (o1 - o0) >> 4;
        p.a_off16[s] = (uint32_t)(s * kWeightTileBytes) >> 4;
    }
    p.a_sa_stride16 = (kW0Steps * kWeightTileBytes) >> 4;
    p.a_lbo16 = 2048 >> 4; p.a_sbo16 = 8;
    p.w_resident = 1; p.w_bytes = (uint32_t)sg.w0s_bytes;
    p.n_wsets = 3;
    p.G = 1; p.RW = 1; p.RP = env_int("VD_TC_L0S_RP", 3);
    p.n_acc = 1; p.acc_delta16 = 0;
    p.ncols = sg.N0s; p.acc_cols = 128; p.acc_stages = 4;
    p.idesc = umma_idesc_f16(128, sg.N0s);
    p.stream_pairs = g.T; p.stream_mode = 2;
    const uint32_t npos = (uint32_t)(sg.R0s / 2) * (g.Wo0 / 2);
    p.stash_group_bytes = align_up(2 * npos * kStash0sPitch * 2, 128);
    return finalize_smem(p, (uint32_t)sg.w0s_bytes, smem, false, 2 * p.stash_group_bytes);
}

// conv 1: the tile / stage geometry of setup_l1 with 8-channel chunks carrying both parts; 74 steps per stage
// passes = 2: the stage copies only the hi planes of each frame and a tap pair gets two MMAs, xh.wh and xh.wl (50 per stage,
// every streamed weight tile used once)
static int setup_l1s(WsParams& p, const Geo& g, int B, uint32_t* smem, int passes = 3) {
    const bool two = passes == 2;
    p.n_u = 1; p.nu_total = 1; p.ug_count = 1; p.w_u_stride = 0;
    const bool nacc4 = g.N1 <= 128 && g.T % 4 == 0;
    const int fpt = nacc4 ? 4 : 2;                                      // output frames per tile
    p.n_tiles = B * (g.T / fpt);
    p.tiles_per_item = g.T / fpt; p.v_count = 1;
    p.item_stride = 8 * g.slice1; p.u_stride = (int64_t)fpt * g.frame1; p.v_stride = 0;
    p.n_sa = 3; p.n_sb = 8; p.sa_stride = g.frame1; p.sb_stride = g.slice1;
    p.n_copies = 1; p.copy_gofs[0] = 0; p.copy_sofs[0] = 0; p.copy_bytes[0] = (uint32_t)(fpt * g.frame1);
    p.stage_bytes = (uint32_t)(fpt * g.frame1);
    const uint32_t fstride = two ? 4u * (uint32_t)g.plane1 : (uint32_t)g.frame1;          // staged bytes per frame
    if (two) {
        p.n_copies = fpt;
        for (int f = 0; f < fpt; ++f) { p.copy_gofs[f] = (int64_t)f * g.frame1; p.copy_sofs[f] = (uint32_t)f * fstride; p.copy_bytes[f] = fstride; }
        p.stage_bytes = (uint32_t)fpt * fstride;
    }
    p.n_steps = two ? 50 : kSteps1s;
    uint32_t off[49];
    for (int i = 0; i < 49; ++i) {
        const int tap = l1s_tap(i), kh = tap / 7, kw = tap % 7;
        const int plane = tap_par(kh) * 2 + tap_par(kw);                // hi part: planes 0..3, lo part: 4..7
        off[i] = (uint32_t)plane * (uint32_t)g.plane1 + (uint32_t)(tap_shift(kh) * g.P1 + tap_shift(kw)) * 16;
        if (i && off[i] <= off[i - 1]) return -2;
    }
    // tap pairs in the two K halves: pair 0 = (tap 0, -) and pair p = (tap 2p-1, tap 2p); three MMAs per pair: xh.wh, xl.wh (same
    // weight tile), xh.wl.  The empty second half of pair 0 has zero weights and addresses tap 1's window (valid staged data).
    for (int pr = 0; pr < 25; ++pr) {
        const int a = pr ? 2 * pr - 1 : 0, b = pr ? 2 * pr : 1;
        if (two) {
            for (int m = 0; m < 2; ++m) { p.b_off16[2 * pr + m] = off[a] >> 4; p.b_lbo16[2 * pr + m] = (off[b] - off[a]) >> 4; }
            continue;
        }
        for (int m = 0; m < 3; ++m) {
            p.b_off16[3 * pr + m] = (off[a] + (m == 1 ? 4u * (uint32_t)g.plane1 : 0u)) >> 4;
            p.b_lbo16[3 * pr + m] = (off[b] - off[a]) >> 4;
        }
    }
    p.a_lbo16 = 2048 >> 4; p.a_sbo16 = 8;
    p.w_resident = 0; p.w_bytes = 0;
    if (two) {
        // tiles [wh_p, wl_p] in image order, one MMA each: the ring of the single-pass conv 1 (3 slots x 5 steps)
        p.G = env_int("VD_TC_L1W_G", 5); p.Gt = 0; p.n_wtiles = 0; p.RW = env_int("VD_TC_L1W_RW", 3); p.RP = env_int("VD_TC_L1W_RP", 2);
    } else {
        // a group = two tap pairs: 6 MMAs per accumulator over 4 weight tiles [wh_p, wl_p, wh_p+1, wl_p+1]
        p.G = 6; p.Gt = 4; p.n_wtiles = 50; p.RW = env_int("VD_TC_L1S_RW", 3); p.RP = 2;
        const uint8_t pat[6] = {0, 0, 1, 2, 2, 3}; for (int i = 0; i < 6; ++i) p.a_in_group[i] = pat[i];
    }
    p.n_acc = fpt; p.acc_delta16 = fstride >> 4;
    p.ncols = g.N1; p.acc_cols = nacc4 ? 128 : 256; p.acc_stages = 1;
    p.idesc = umma_idesc_f16(128, g.N1);
    return finalize_smem(p, (uint32_t)(p.Gt ? p.Gt : p.G) * p.RW * kWeightTileBytes, smem, false, (uint32_t)(2 * (fpt / 2) * g.H2 * g.H2) * kStashPitch * 2);
}

// conv 2: the tile / stage geometry of setup_l2 with 32-channel quarters carrying both parts; 18 steps per stage
// passes = 2: hi chunks only, two MMAs per chunk pair (xh.wh, xh.wl): 12 per stage
static int setup_l2s(WsParams& p, const Geo& g, int B, uint32_t* smem, int passes = 3) {
    const bool two = passes == 2;
    p.n_u = 1; p.nu_total = 1; p.ug_count = 1; p.w_u_stride = 0;
    const int VPT = kVideosPerTile2;
    const SGeo sg = make_sgeo(g);
    p.n_tiles = (B + VPT - 1) / VPT;
    p.tiles_per_item = 1; p.v_count = 1;
    p.item_stride = VPT * sg.video2s; p.u_stride = 0; p.v_stride = 0;
    p.n_sa = 49; p.n_sb = 4; p.sa_stride = 4 * g.group2; p.sb_stride = g.group2;
    p.n_copies = VPT;
    const uint32_t vbytes = (uint32_t)(two ? g.group2 / 2 : g.group2);       // staged bytes per video: [part][k 4] chunks, hi part first
    for (int v = 0; v < VPT; ++v) {
        p.copy_gofs[v] = (int64_t)v * sg.video2s;
        p.copy_sofs[v] = (uint32_t)v * vbytes;
        p.copy_bytes[v] = vbytes;
    }
    p.stage_bytes = (uint32_t)VPT * vbytes;
    p.n_steps = two ? 12 : kSteps2s;
    for (int kt = 0; kt < 3; ++kt) {
        const uint32_t tofs = (uint32_t)((int64_t)kt * g.HW2 * 16);
        for (int pr = 0; pr < 2; ++pr) {                                // chunk pair (2pr, 2pr+1) in the two K halves
            if (two) {
                for (int m = 0; m < 2; ++m) {                           // xh.wh, xh.wl
                    p.b_off16[kt * 4 + pr * 2 + m] = (uint32_t)((2 * pr) * g.chunk2 + tofs) >> 4;
                    p.b_lbo16[kt * 4 + pr * 2 + m] = (uint32_t)g.chunk2 >> 4;
                }
                continue;
            }
            for (int m = 0; m < 3; ++m) {                               // xh.wh, xl.wh (same weight tile), xh.wl
                p.b_off16[kt * 6 + pr * 3 + m] = (uint32_t)((2 * pr + (m == 1 ? 4 : 0)) * g.chunk2 + tofs) >> 4;
                p.b_lbo16[kt * 6 + pr * 3 + m] = (uint32_t)g.chunk2 >> 4;
            }
        }
    }
    p.a_lbo16 = 2048 >> 4; p.a_sbo16 = 8;
    p.w_resident = 0; p.w_bytes = 0;
    if (two) {
        p.G = env_int("VD_TC_L2W_G", 6); p.Gt = 0; p.n_wtiles = 0; p.RW = env_int("VD_TC_L2W_RW", 3); p.RP = env_int("VD_TC_L2W_RP", 2);
    } else {
        p.G = 6; p.Gt = 4; p.n_wtiles = 12; p.RW = env_int("VD_TC_L2S_RW", 3); p.RP = 2;
        const uint8_t pat[6] = {0, 0, 1, 2, 2, 3}; for (int i = 0; i < 6; ++i) p.a_in_group[i] = pat[i];
    }
    p.n_acc = VPT; p.acc_delta16 = vbytes >> 4;
    p.ncols = g.N2; p.acc_cols = 128; p.acc_stages = 1;
    p.idesc = umma_idesc_f16(128, g.N2);
    return finalize_smem(p, (uint32_t)(p.Gt ? p.Gt : p.G) * p.RW * kWeightTileBytes, smem);
}

static int setup_bwd(WsParams& p, const Geo& g, int layer, int B, uint32_t* smem, bool fp32_out = false) {
    const BwdGeo b = make_bwd_geo(g, layer);
    VD_REQUIRE(b.pixels % 16 == 0, "tc bwd: pixel count must be a multiple of 16");
    // enough CTA tiles for two waves even at small B: split the NU row tiles of a pixel stage over ug_count CTAs
    int ug = (2 * 148 + B * b.NT - 1) / (B * b.NT);
    if (ug < 1) ug = 1;
    if (ug > b.NU) ug = b.NU;
    p.nu_total = b.NU; p.n_u = (b.NU + ug - 1) / ug; p.ug_count = (b.NU + p.n_u - 1) / p.n_u;
    p.n_tiles = B * b.NT * p.ug_count; p.tiles_per_item = b.NT; p.v_count = b.NT;
    p.item_stride = b.dy_video; p.u_stride = 0; p.v_stride = (int64_t)(b.K / 8) * b.NC * 16;
    p.w_u_stride = (int64_t)b.n_steps * kWeightTileBytes;
    p.n_sa = 1; p.n_sb = 1; p.sa_stride = 0; p.sb_stride = 0;
    p.n_copies = 1; p.copy_gofs[0] = 0; p.copy_sofs[0] = 0; p.copy_bytes[0] = (uint32_t)((b.K / 8) * b.NC * 16);
    p.stage_bytes = p.copy_bytes[0];
    p.n_steps = b.n_steps;
    for (int j = 0; j < b.n_steps; ++j) { p.b_off16[j] = (uint32_t)(2 * j * b.NC); p.b_lbo16[j] = (uint32_t)b.NC; }
    p.a_lbo16 = 2048 >> 4; p.a_sbo16 = 8;
    p.w_resident = 0; p.w_bytes = 0;
    p.G = b.n_steps; p.RW = 2; p.RP = 2;
    p.n_acc = 1; p.acc_delta16 = 0;
    p.ncols = b.NC; p.acc_cols = 256; p.acc_stages = 2;
    p.idesc = umma_idesc_bf16(128, b.NC);
    return finalize_smem(p, (uint32_t)p.G * p.RW * kWeightTileBytes, smem, fp32_out);     // fp32 columns: transposed coalesced stores
}

template <int EPI, int NACC>
static int launch_n(const WsParams& p, uint32_t smem, cudaStream_t stream) {
    static bool configured = false;
    static int sm_count = 148;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ws_gemm_kernel<EPI, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(ws_gemm): %s", cudaGetErrorString(e)); return (int)e; }
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    if (p.n_tiles <= 0) return 0;
    const int grid = p.n_tiles < sm_count ? p.n_tiles : sm_count;
    ws_gemm_kernel<EPI, NACC><<<grid, EPI == EPI_L0S ? kThreadsL0S : kThreads, smem, stream>>>(p);
    return check_launch("tc ws_gemm");
}

template <int EPI>
static int launch(const WsParams& p, uint32_t smem, cudaStream_t stream) {
    if (EPI == EPI_RAW) {
        if (p.n_acc == 1) return launch_n<EPI_RAW, 1>(p, smem, stream);
        if (p.n_acc == 2) return launch_n<EPI_RAW, 2>(p, smem, stream);
        if (p.n_acc == 4) return launch_n<EPI_RAW, 4>(p, smem, stream);
    } else if (EPI == EPI_L0 && p.n_acc == 1) {
        return launch_n<EPI_L0, 1>(p, smem, stream);
    } else if (EPI == EPI_L1 && p.n_acc == 2) {
        return launch_n<EPI_L1, 2>(p, smem, stream);
    } else if (EPI == EPI_L1 && p.n_acc == 4) {
        return launch_n<EPI_L1, 4>(p, smem, stream);
    } else if (EPI == EPI_L0S && p.n_acc == 1) {
        return launch_n<EPI_L0S, 1>(p, smem, stream);
    } else if (EPI == EPI_L1S && p.n_acc == 2) {
        return launch_n<EPI_L1S, 2>(p, smem, stream);
    } else if (EPI == EPI_L1S && p.n_acc == 4) {
        return launch_n<EPI_L1S, 4>(p, smem, stream);
    } else if (EPI == EPI_L2 && p.n_acc == 4) {
        return launch_n<EPI_L2, 4>(p, smem, stream);
    } else if (EPI == EPI_DG1 && p.n_acc == 1) {
        return launch_n<EPI_DG1, 1>(p, smem, stream);
    } else if (EPI == EPI_DG0 && p.n_acc == 1) {
        return launch_n<EPI_DG0, 1>(p, smem, stream);
    } else if (EPI == EPI_PLAIN) {
        if (p.n_acc == 1) return launch_n<EPI_PLAIN, 1>(p, smem, stream);
        if (p.n_acc == 2) return launch_n<EPI_PLAIN, 2>(p, smem, stream);
        if (p.n_acc == 4) return launch_n<EPI_PLAIN, 4>(p, smem, stream);
    }
    set_error("tc ws_gemm: no kernel instance for epilogue %d with %d accumulators", EPI, p.n_acc);
    return -1;
}

}  // namespace tc
}  // namespace vd

using namespace vd;
using namespace vd::tc;

extern "C" int vd_tc_plan_make(vd_tc_plan* plan, int T, int H, int W) {
    VD_REQUIRE(plan != nullptr, "tc_plan_make: NULL plan");
    VD_REQUIRE(H == W && geo_supported(T, H), "tc path supports square 112x112 videos with T in {4,8,12,16} or 64x64 videos with T in {8,16,24,32} (got T=%d H=%d W=%d)", T, H, W);
    const Geo g = make_geo(T, H);
    memset(plan, 0, sizeof(*plan));
    plan->T = T; plan->H = H; plan->W = W;
    plan->c1 = 64; plan->c2 = 128; plan->c3 = 128;
    plan->T1 = T; plan->H1 = g.Ho0; plan->W1 = g.Wo0; plan->T1p = T; plan->H1p = g.H1; plan->W1p = g.H1;
    plan->T2 = T; plan->H2 = g.Ho1; plan->W2 = g.Wo1; plan->T2p = g.T2; plan->H2p = g.H2; plan->W2p = g.H2;
    plan->T3 = g.To2; plan->H3 = g.Ho2; plan->W3 = g.Wo2; plan->T3p = g.T3p; plan->H3p = g.H3p; plan->W3p = g.H3p;
    plan->embed_dim = g.embed_dim;
    plan->x0_bytes_per_video = g.video0;
    plan->a1_bytes_per_video = g.video1;
    plan->a2_bytes_per_video = g.video2;
    plan->w0_bytes = kW0Bytes;
    plan->w1_bytes = (int64_t)588 * kWeightTileBytes;
    plan->w2_bytes = (int64_t)1176 * kWeightTileBytes;
    plan->tab_bytes = 0;
    const BwdGeo b0 = make_bwd_geo(g, 0), b1 = make_bwd_geo(g, 1), b2 = make_bwd_geo(g, 2);
    plan->wt0_bytes = b0.wt_bytes; plan->wt1_bytes = b1.wt_bytes; plan->wt2_bytes = b2.wt_bytes;
    plan->dy0_bytes_per_video = b0.dy_video; plan->dy1_bytes_per_video = b1.dy_video; plan->dy2_bytes_per_video = b2.dy_video;
    plan->col0_bytes_per_video = b0.col_video_elems * 2; plan->col1_bytes_per_video = b1.col_video_elems * 2;
    plan->col2_bytes_per_video = b2.col_video_elems * 2;      // bf16 column buffers
    return 0;
}

extern "C" int vd_tc_conv_layer(int layer, const void* in, const void* wimg, const float* bias, void* out,
                                uint8_t* code, int code_first_item, const vd_tc_plan* plan, const int64_t* item_index,
                                int B, int raw, void* stream) {
    VD_REQUIRE(plan && in && wimg && out, "tc_conv_layer: NULL pointer");
    VD_REQUIRE(layer >= 0 && layer <= 2, "tc_conv_layer: layer must be 0, 1 or 2");
    VD_REQUIRE(B >= 0, "tc_conv_layer: negative batch");
    VD_REQUIRE(code_first_item >= 0, "tc_conv_layer: negative code_first_item");
    VD_REQUIRE(geo_supported(plan->T, plan->H), "tc_conv_layer: unsupported geometry");
    VD_REQUIRE(raw || bias, "tc_conv_layer: bias is NULL");
    VD_REQUIRE(raw >= 0 && raw <= 3, "tc_conv_layer: raw must be 0 (fused), 1 (accumulator dump), 2 (plain NCDHW fp32) or 3 (plain, out += result)");
    if (B == 0) return 0;
    WsParams p;
    memset(&p, 0, sizeof(p));
    const Geo g = make_geo(plan->T, plan->H);
    uint32_t smem = 0;
    int rc = layer == 0 ? setup_l0(p, g, B, &smem, raw == 0) : layer == 1 ? setup_l1(p, g, B, &smem) : setup_l2(p, g, B, &smem);
    if (rc) { if (rc == -2) set_error("tc conv 0: non-monotone chunk pairing"); return rc; }
    VD_REQUIRE(item_index == nullptr || layer == 0, "tc_conv_layer: item_index is only valid for layer 0");
    p.pix = (const uint8_t*)in; p.wimg = (const uint8_t*)wimg; p.item_index = item_index;
    p.prof = g_prof; p.dbg = env_int("VD_TC_DBG", 0);
    p.epi.bias = bias; p.epi.out = (uint8_t*)out; p.epi.code = code; p.epi.code_first = code_first_item; p.epi.raw = (float*)out;
    p.epi.T = plan->T; p.epi.n_items = B; p.epi.g = g;
    p.epi.layer = layer;
    cudaStream_t s = (cudaStream_t)stream;
    p.epi.accum = raw == 3;
    if (raw >= 2) return launch<EPI_PLAIN>(p, smem, s);
    if (raw) return launch<EPI_RAW>(p, smem, s);
    if (layer == 0) {
        if (env_int("VD_TC_L0_STREAM", 1)) {
            // input-frame streaming (WsParams::stream_pairs): same tables, columns instead of (pair, row band) tiles
            p.stream_pairs = g.T / 2;
            p.n_wsets = 4;
            p.n_tiles = B * p.v_count;
            p.tiles_per_item = p.v_count;
            p.n_sa = g.T; p.n_sb = 1;
            for (int c = 0; c < p.n_copies; ++c) p.copy_gofs[c] += g.frame0;        // frame i of the video is t_pad = i + 1
        }
        return launch<EPI_L0>(p, smem, s);
    }
    if (layer == 1) {
        if (g.N1 <= 128 && g.T % 4 == 0 && env_int("VD_TC_L1_NACC4", 1)) {
            // small frames (64x64: N1 = 80 columns): four consecutive output frames per tile.  With two, a 4 KB weight tile
            // feeds only 2 x 40 MMA cycles and the weight stream (51 B/clk per SM) exceeds what L2 can deliver to 148 SMs.
            p.n_acc = 4; p.acc_cols = 128;
            p.n_tiles = B * (g.T / 4);
            p.tiles_per_item = g.T / 4;
            p.u_stride = 4 * g.frame1;
            p.copy_bytes[0] = (uint32_t)(4 * g.frame1);
            p.stage_bytes = (uint32_t)(4 * g.frame1);
            if (int rc2 = finalize_smem(p, (uint32_t)p.G * p.RW * kWeightTileBytes, &smem, false, (uint32_t)(2 * g.H2 * g.H2) * kStashPitch * 2))
                return rc2;
        }
        return launch<EPI_L1>(p, smem, s);
    }
    return launch<EPI_L2>(p, smem, s);
}

// ---- split-fp16 forward (SGeo): the same three fused layers on fp16 hi / lo operand pairs ----
extern "C" int vd_tc_x3_sizes(const vd_tc_plan* plan, int64_t* out) {
    VD_REQUIRE(plan && out && geo_supported(plan->T, plan->H), "tc_x3_sizes: bad argument");
    const Geo g = make_geo(plan->T, plan->H);
    const SGeo sg = make_sgeo(g);
    out[0] = sg.video0s; out[1] = sg.video1s; out[2] = sg.video2s;
    out[3] = sg.w0s_bytes; out[4] = sg.w1s_bytes; out[5] = sg.w2s_bytes;
    return 0;
}

static int x3_conv_layer_impl(int layer, const void* in, const void* wimg, const float* bias, void* out,
                              uint8_t* code, int code_first_item, const vd_tc_plan* plan, const int64_t* item_index,
                              int B, int passes, void* stream);

extern "C" int vd_tc_x3_conv_layer(int layer, const void* in, const void* wimg, const float* bias, void* out,
                                   uint8_t* code, int code_first_item, const vd_tc_plan* plan, const int64_t* item_index,
                                   int B, void* stream) {
    return x3_conv_layer_impl(layer, in, wimg, bias, out, code, code_first_item, plan, item_index, B, 3, stream);
}

// passes = 3: vd_tc_x3_conv_layer.  passes = 2: the two-product mode of the frozen real videos (setup_l0s): `in` holds the
// hi part only (layer 0: the X0h layout = X0 with fp16 values, vd_tc_x3_pack_video_hi; layers 1 / 2: the A1s / A2s layouts
// with their lo planes unused), the same weight images, and layers 0 / 1 write the hi part of their output only.
extern "C" int vd_tc_x3_conv_layer_ex(int layer, const void* in, const void* wimg, const float* bias, void* out,
                                      uint8_t* code, int code_first_item, const vd_tc_plan* plan, const int64_t* item_index,
                                      int B, int passes, void* stream) {
    VD_REQUIRE(passes == 2 || passes == 3, "tc_x3_conv_layer_ex: passes must be 2 or 3");
    return x3_conv_layer_impl(layer, in, wimg, bias, out, code, code_first_item, plan, item_index, B, passes, stream);
}

// Plain convolution of layer 1 / 2 on the split-fp16 tables: out = fp32 NCDHW (B, Cout, To, Ho, Wo), pre-activation (+ bias when
// non-NULL) with all three products xh*wh + xl*wh + xh*wl accumulated in TMEM by ONE launch — the fprop of the differentiable
// conv trio (MTT), which evaluated the same sum as three single-pass launches with twice the MMAs.  in = vd_tc_x3_pack_act,
// wimg = vd_tc_x3_pack_weights.
extern "C" int vd_tc_x3_conv_plain(int layer, const void* in, const void* wimg, const float* bias, float* out,
                                   const vd_tc_plan* plan, int B, void* stream) {
    VD_REQUIRE(plan && in && wimg && out, "tc_x3_conv_plain: NULL pointer");
    VD_REQUIRE(layer == 1 || layer == 2, "tc_x3_conv_plain: layer must be 1 or 2");
    VD_REQUIRE(B >= 0 && geo_supported(plan->T, plan->H), "tc_x3_conv_plain: bad batch / geometry");
    if (B == 0) return 0;
    WsParams p;
    memset(&p, 0, sizeof(p));
    const Geo g = make_geo(plan->T, plan->H);
    uint32_t smem = 0;
    int rc = layer == 1 ? setup_l1s(p, g, B, &smem, 3) : setup_l2s(p, g, B, &smem, 3);
    if (rc) { if (rc == -2) set_error("tc x3 conv plain %d: non-monotone window offsets", layer); return rc; }
    p.pix = (const uint8_t*)in; p.wimg = (const uint8_t*)wimg; p.item_index = nullptr;
    p.prof = g_prof; p.dbg = env_int("VD_TC_DBG", 0);
    p.epi.bias = bias; p.epi.out = (uint8_t*)out; p.epi.raw = out;
    p.epi.T = plan->T; p.epi.n_items = B; p.epi.g = g; p.epi.layer = layer; p.epi.accum = 0;
    return launch<EPI_PLAIN>(p, smem, (cudaStream_t)stream);
}

static int x3_conv_layer_impl(int layer, const void* in, const void* wimg, const float* bias, void* out,
                              uint8_t* code, int code_first_item, const vd_tc_plan* plan, const int64_t* item_index,
                              int B, int passes, void* stream) {
    VD_REQUIRE(plan && in && wimg && out && bias, "tc_x3_conv_layer: NULL pointer");
    VD_REQUIRE(layer >= 0 && layer <= 2, "tc_x3_conv_layer: layer must be 0, 1 or 2");
    VD_REQUIRE(B >= 0 && code_first_item >= 0, "tc_x3_conv_layer: negative batch / code_first_item");
    VD_REQUIRE(geo_supported(plan->T, plan->H), "tc_x3_conv_layer: unsupported geometry");
    VD_REQUIRE(item_index == nullptr || layer == 0, "tc_x3_conv_layer: item_index is only valid for layer 0");
    if (B == 0) return 0;
    WsParams p;
    memset(&p, 0, sizeof(p));
    const Geo g = make_geo(plan->T, plan->H);
    uint32_t smem = 0;
    int rc = layer == 0 ? setup_l0s(p, g, B, &smem, passes) : layer == 1 ? setup_l1s(p, g, B, &smem, passes) : setup_l2s(p, g, B, &smem, passes);
    if (rc) { if (rc == -2) set_error("tc x3 conv %d: non-monotone window offsets", layer); return rc; }
    p.pix = (const uint8_t*)in; p.wimg = (const uint8_t*)wimg; p.item_index = item_index;
    p.prof = g_prof; p.dbg = env_int("VD_TC_DBG", 0);
    p.epi.bias = bias; p.epi.out = (uint8_t*)out; p.epi.code = code; p.epi.code_first = code_first_item; p.epi.raw = (float*)out;
    p.epi.T = plan->T; p.epi.n_items = B; p.epi.g = g; p.epi.layer = layer; p.epi.r0s = make_sgeo(g).R0s;
    p.epi.hi_only = passes == 2;
    cudaStream_t s = (cudaStream_t)stream;
    if (layer == 0) return launch<EPI_L0S>(p, smem, s);
    if (layer == 1) return launch<EPI_L1S>(p, smem, s);
    return launch<EPI_L2>(p, smem, s);
}

// Host-side introspection for the CPU emulator in tests/ (no GPU work): dumps the exact kernel
// parameters that vd_tc_conv_layer would launch with.
extern "C" int vd_tc_debug_params(int layer, const vd_tc_plan* plan, int B, int64_t* out, int cap) {
    VD_REQUIRE(plan && out && cap >= 33 + 3 * kMaxCopies + 3 * kMaxSteps + 18, "tc_debug_params: buffer too small");
    VD_REQUIRE(layer >= 0 && layer <= 11 && geo_supported(plan->T, plan->H), "tc_debug_params: bad layer / geometry");
    WsParams p;
    memset(&p, 0, sizeof(p));
    const Geo g = make_geo(plan->T, plan->H);
    uint32_t smem = 0;
    int rc = layer == 0 ? setup_l0(p, g, B, &smem) : layer == 1 ? setup_l1(p, g, B, &smem) : layer == 2 ? setup_l2(p, g, B, &smem)
                        : layer == 6 ? setup_l0s(p, g, B, &smem) : layer == 7 ? setup_l1s(p, g, B, &smem)
                        : layer == 8 ? setup_l2s(p, g, B, &smem)
                        : layer == 9 ? setup_l0s(p, g, B, &smem, 2) : layer == 10 ? setup_l1s(p, g, B, &smem, 2)
                        : layer == 11 ? setup_l2s(p, g, B, &smem, 2) : setup_bwd(p, g, layer - 3, B, &smem);
    if (rc) return rc;
    int i = 0;
    out[i++] = p.n_tiles; out[i++] = p.tiles_per_item; out[i++] = p.v_count;
    out[i++] = p.item_stride; out[i++] = p.u_stride; out[i++] = p.v_stride;
    out[i++] = p.n_sa; out[i++] = p.n_sb; out[i++] = p.sa_stride; out[i++] = p.sb_stride;
    out[i++] = p.n_copies; out[i++] = p.stage_bytes; out[i++] = p.stage_pitch; out[i++] = p.n_steps;
    out[i++] = p.a_sa_stride16; out[i++] = p.a_lbo16; out[i++] = p.a_sbo16; out[i++] = p.w_resident;
    out[i++] = p.w_bytes; out[i++] = p.G; out[i++] = p.RW; out[i++] = p.RP; out[i++] = p.n_acc;
    out[i++] = p.acc_delta16; out[i++] = p.ncols; out[i++] = p.acc_cols; out[i++] = p.acc_stages;
    out[i++] = p.idesc; out[i++] = p.smem_w_off; out[i++] = p.smem_pix_off; out[i++] = smem; out[i++] = p.n_u;
    for (int k = 0; k < kMaxCopies; ++k) out[i++] = p.copy_gofs[k];
    for (int k = 0; k < kMaxCopies; ++k) out[i++] = p.copy_sofs[k];
    for (int k = 0; k < kMaxCopies; ++k) out[i++] = p.copy_bytes[k];
    for (int k = 0; k < kMaxSteps; ++k) out[i++] = p.b_off16[k];
    for (int k = 0; k < kMaxSteps; ++k) out[i++] = p.b_lbo16[k];
    for (int k = 0; k < kMaxSteps; ++k) out[i++] = p.a_off16[k];
    out[i++] = p.w_u_stride;
    out[i++] = p.Gt; out[i++] = p.n_wtiles;
    for (int k = 0; k < 16; ++k) out[i++] = p.a_in_group[k];
    return 0;
}

static int bwd_gemm_impl(int layer, const void* dy, const void* wt, void* col, const vd_tc_plan* plan, int B, int col_fp32,
                         void* stream);

extern "C" int vd_tc_bwd_gemm(int layer, const void* dy, const void* wt, void* col, const vd_tc_plan* plan,
                              int B, void* stream) {
    return bwd_gemm_impl(layer, dy, wt, col, plan, B, 0, stream);
}

// col_fp32 != 0: fp32 column buffer (twice the traffic; the accuracy modes of the conv trio)
extern "C" int vd_tc_bwd_gemm_ex(int layer, const void* dy, const void* wt, void* col, const vd_tc_plan* plan,
                                 int B, int col_fp32, void* stream) {
    return bwd_gemm_impl(layer, dy, wt, col, plan, B, col_fp32, stream);
}

static int bwd_gemm_impl(int layer, const void* dy, const void* wt, void* col, const vd_tc_plan* plan, int B, int col_fp32,
                         void* stream) {
    VD_REQUIRE(plan && dy && wt && col, "tc_bwd_gemm: NULL pointer");
    VD_REQUIRE(layer >= 0 && layer <= 2 && B >= 0, "tc_bwd_gemm: bad layer / batch");
    VD_REQUIRE(geo_supported(plan->T, plan->H), "tc_bwd_gemm: unsupported geometry");
    if (B == 0) return 0;
    WsParams p;
    memset(&p, 0, sizeof(p));
    const Geo g = make_geo(plan->T, plan->H);
    uint32_t smem = 0;
    if (int rc = setup_bwd(p, g, layer, B, &smem, col_fp32 != 0)) return rc;
    p.pix = (const uint8_t*)dy; p.wimg = (const uint8_t*)wt; p.item_index = nullptr;
    p.epi.raw = (float*)col; p.epi.raw_bf16 = col_fp32 ? 0 : 1; p.epi.T = plan->T; p.epi.n_items = B; p.epi.g = g;
    return launch<EPI_RAW>(p, smem, (cudaStream_t)stream);
}

// Direct dgrad of conv 1 (no column buffer): two launches (input-row parity ph = 0, 1) of the shifted-window GEMM
// described at Dg1Geo.  dyp = padded planar dY of conv 1 (vd_tc_pack_dyp1 / vd_tc_bwd_col2im_ex), wimg0/1 = weight
// images of vd_tc_pack_dgrad1_weights.  code0 != NULL: out = packed dY of conv 0's column GEMM (routing applied);
// code0 == NULL: out = fp32 NCDHW (B, 64, T, H1, H1).
static int dgrad1_impl(const void* dyp, const void* wimg0, const void* wimg1, const uint8_t* code0, void* out,
                       const vd_tc_plan* plan, int B, int dy0_planar, void* stream, int accumulate = 0);

extern "C" int vd_tc_dgrad1(const void* dyp, const void* wimg0, const void* wimg1, const uint8_t* code0, void* out,
                            const vd_tc_plan* plan, int B, void* stream) {
    return dgrad1_impl(dyp, wimg0, wimg1, code0, out, plan, B, 0, stream);
}

// plain fp32 NCDHW output (B, 64, T, H1, H1); accumulate != 0: out += result (the passes of a split-bf16 dgrad)
extern "C" int vd_tc_dgrad1_plain(const void* dyp, const void* wimg0, const void* wimg1, float* out, const vd_tc_plan* plan, int B,
                                  int accumulate, void* stream) {
    return dgrad1_impl(dyp, wimg0, wimg1, nullptr, out, plan, B, 0, stream, accumulate);
}

// out_planar != 0 (needs code0): the routed gradient goes to the padded planar dY of conv 0 consumed by vd_tc_dgrad0
// (halo cells zeroed once by the caller) instead of the column-GEMM operand
extern "C" int vd_tc_dgrad1_ex(const void* dyp, const void* wimg0, const void* wimg1, const uint8_t* code0, void* out,
                               const vd_tc_plan* plan, int B, int out_planar, void* stream) {
    VD_REQUIRE(!out_planar || code0, "tc_dgrad1_ex: the planar dY0 output needs the routing code of conv 0");
    return dgrad1_impl(dyp, wimg0, wimg1, code0, out, plan, B, out_planar, stream);
}

static int dgrad1_impl(const void* dyp, const void* wimg0, const void* wimg1, const uint8_t* code0, void* out,
                       const vd_tc_plan* plan, int B, int dy0_planar, void* stream, int accumulate) {
    VD_REQUIRE(dyp && wimg0 && wimg1 && out && plan, "tc_dgrad1: NULL pointer");
    VD_REQUIRE(!accumulate || !code0, "tc_dgrad1: accumulation is for the plain fp32 output");
    VD_REQUIRE(geo_supported(plan->T, plan->H) && B >= 0, "tc_dgrad1: unsupported geometry / batch");
    if (B == 0) return 0;
    const Geo g = make_geo(plan->T, plan->H);
    const Dg1Geo d = make_dg1_geo(g);
    VD_REQUIRE(d.PD <= 16 && d.N % 16 == 0 && d.N <= 256, "tc_dgrad1: tile does not fit one accumulator");
    for (int ph = 0; ph < 2; ++ph) {
        WsParams p;
        memset(&p, 0, sizeof(p));
        p.n_u = 1; p.nu_total = 1; p.ug_count = 1; p.w_u_stride = 0; p.w_item_stride = 0;
        p.n_tiles = B * g.T; p.tiles_per_item = g.T; p.v_count = 1;
        // tile (item, t): stage sa = kt reads the padded frame t + 2 - kt, sb = co half (8 chunks)
        p.item_stride = d.video_bytes; p.u_stride = d.frame_bytes; p.v_stride = 0;
        p.n_sa = 3; p.n_sb = 2; p.sa_stride = -d.frame_bytes; p.sb_stride = d.frame_bytes / 2;
        p.n_copies = 1; p.copy_gofs[0] = 2 * d.frame_bytes; p.copy_sofs[0] = 0; p.copy_bytes[0] = (uint32_t)(d.frame_bytes / 2);
        p.stage_bytes = p.copy_bytes[0];
        p.n_steps = d.n_steps[ph];
        const int first = ph ? 2 : 1;
        int j = 0;
        for (int ih = 0; ih < d.n_sh[ph]; ++ih)
            for (int iw = 0; iw < 4; ++iw)
                for (int cs = 0; cs < 4; ++cs, ++j) {
                    const int sh = dg1_shift(ih, first), sw = dg1_shift(iw, 2);
                    p.b_off16[j] = (uint32_t)(2 * cs * d.plane16 + (sh + 1) * d.PD + (sw + 1));
                    p.b_lbo16[j] = (uint32_t)d.plane16;
                }
        p.a_lbo16 = 2048 >> 4; p.a_sbo16 = 8;
        p.w_resident = 0; p.w_bytes = 0;
        p.G = 8; p.RW = 3; p.RP = 3;
        p.n_acc = 1; p.acc_delta16 = 0;
        p.ncols = (uint32_t)d.N; p.acc_cols = 256; p.acc_stages = 2;
        p.idesc = umma_idesc_bf16(128, (uint32_t)d.N);
        uint32_t smem = 0;
        // plain (unrouted) output: row-assembly scratch of the epilogue warp pairs
        if (int rc = finalize_smem(p, (uint32_t)p.G * p.RW * kWeightTileBytes, &smem, /*epi_scratch=*/code0 == nullptr)) return rc;
        p.pix = (const uint8_t*)dyp; p.wimg = (const uint8_t*)(ph ? wimg1 : wimg0); p.item_index = nullptr;
        p.epi.out = (uint8_t*)out; p.epi.code = const_cast<uint8_t*>(code0); p.epi.T = plan->T; p.epi.n_items = B; p.epi.g = g;
        p.epi.ph = ph; p.epi.bb = make_bwd_geo(g, 0); p.epi.dy0_planar = dy0_planar; p.epi.accum = accumulate;
        if (int rc = launch<EPI_DG1>(p, smem, (cudaStream_t)stream)) return rc;
    }
    return 0;
}

// Direct dgrad of conv 0 (no column buffer, Dg0Geo): dyp0 = padded planar dY of conv 0 (vd_tc_dgrad1_ex /
// vd_tc_pack_dyp0), wimg = resident weight image (vd_tc_pack_dgrad0_weights), out = fp32 gradient w.r.t. the input
// video in the (B,T,3,H,W) layout (ncdhw = 0) or (B,3,T,H,W) (ncdhw = 1).
static int dgrad0_impl(const void* dyp0, const void* wimg, float* out, const vd_tc_plan* plan, int B, int ncdhw, int accumulate,
                       void* stream);

extern "C" int vd_tc_dgrad0(const void* dyp0, const void* wimg, float* out, const vd_tc_plan* plan, int B, int ncdhw,
                            void* stream) {
    return dgrad0_impl(dyp0, wimg, out, plan, B, ncdhw, 0, stream);
}

// accumulate != 0: out += result (the passes of a split-bf16 dgrad share one fp32 tensor)
extern "C" int vd_tc_dgrad0_ex(const void* dyp0, const void* wimg, float* out, const vd_tc_plan* plan, int B, int ncdhw,
                               int accumulate, void* stream) {
    return dgrad0_impl(dyp0, wimg, out, plan, B, ncdhw, accumulate, stream);
}

static int dgrad0_impl(const void* dyp0, const void* wimg, float* out, const vd_tc_plan* plan, int B, int ncdhw, int accumulate,
                       void* stream) {
    VD_REQUIRE(dyp0 && wimg && out && plan, "tc_dgrad0: NULL pointer");
    VD_REQUIRE(geo_supported(plan->T, plan->H) && B >= 0, "tc_dgrad0: unsupported geometry / batch");
    if (B == 0) return 0;
    const Geo g = make_geo(plan->T, plan->H);
    const Dg0Geo d = make_dg0_geo(g);
    VD_REQUIRE(g.Wo0 + 3 <= d.PD && g.Ho0 % d.RT == 0, "tc_dgrad0: frame does not fit the pixel tile");
    WsParams p;
    memset(&p, 0, sizeof(p));
    p.n_u = 1; p.nu_total = 1; p.ug_count = 1; p.w_u_stride = 0; p.w_item_stride = 0;
    const int nrb = g.Ho0 / d.RT;
    p.n_tiles = B * g.T * nrb; p.tiles_per_item = g.T * nrb; p.v_count = nrb;
    // tile (item, t, row block): stage sa = kt reads RS rows of the padded frame t + 2 - kt, all 8 chunks
    p.item_stride = d.video_bytes; p.u_stride = d.frame_bytes; p.v_stride = (int64_t)d.RT * d.PD * 16;
    p.n_sa = 3; p.n_sb = 1; p.sa_stride = -d.frame_bytes; p.sb_stride = 0;
    p.n_copies = 8;
    for (int c = 0; c < 8; ++c) {
        p.copy_gofs[c] = 2 * d.frame_bytes + (int64_t)c * d.plane16 * 16;
        p.copy_sofs[c] = (uint32_t)(c * d.stage_plane16 * 16);
        p.copy_bytes[c] = (uint32_t)(d.stage_plane16 * 16);
    }
    p.stage_bytes = (uint32_t)(8 * d.stage_plane16 * 16);
    p.n_steps = 64;
    int j = 0;
    for (int ih = 0; ih < 4; ++ih)
        for (int iw = 0; iw < 4; ++iw)
            for (int cs = 0; cs < 4; ++cs, ++j) {
                const int sh = dg1_shift(ih, 2), sw = dg1_shift(iw, 2);
                p.b_off16[j] = (uint32_t)(2 * cs * d.stage_plane16 + (sh + 1) * d.PD + (sw + 1));
                p.b_lbo16[j] = (uint32_t)d.stage_plane16;
                p.a_off16[j] = (uint32_t)(j * 32);                 // 512-byte weight tiles [k 2][16 rows][8]
            }
    p.a_sa_stride16 = 64 * 32;
    p.a_lbo16 = 256 >> 4; p.a_sbo16 = 8;
    p.w_resident = 1; p.w_bytes = (uint32_t)d.wimg_bytes;
    p.G = 1; p.RW = 1; p.RP = 2;
    p.n_acc = 1; p.acc_delta16 = 0;
    p.ncols = 16; p.acc_cols = 16; p.acc_stages = 2;
    p.idesc = umma_idesc_bf16(128, 16);
    p.swap_ab = 1;
    uint32_t smem = 0;
    if (int rc = finalize_smem(p, (uint32_t)d.wimg_bytes, &smem)) return rc;
    p.pix = (const uint8_t*)dyp0; p.wimg = (const uint8_t*)wimg; p.item_index = nullptr;
    p.epi.out = (uint8_t*)out; p.epi.T = plan->T; p.epi.n_items = B; p.epi.g = g; p.epi.ncdhw = ncdhw; p.epi.accum = accumulate;
    return launch<EPI_DG0>(p, smem, (cudaStream_t)stream);
}

// wgrad of conv `layer` as a split-K GEMM: raw[(split, ntile)][cout 128][col 256] = sum over the slice's pixels of
// gy[cout][pixel] * xcol[(ci,tap)][pixel].  Operands from vd_tc_wgrad_pack (tc_trio.cu): the gy image is the M
// operand (4 KiB tiles streamed through the weight ring, one K = 16 pixel step each), the im2col columns are the
// N operand (one 64 KiB stage = 128 pixels x 256 columns per bulk copy).
extern "C" int vd_tc_wgrad_plan(int layer, const vd_tc_plan* plan, int B, int64_t* out);
extern "C" int vd_tc_wgrad_kt_mode(int layer);

extern "C" int vd_tc_wgrad_gemm(int layer, const void* xcol, const void* gyimg, float* raw, const vd_tc_plan* plan, int B,
                                void* stream) {
    VD_REQUIRE(xcol && gyimg && raw && plan, "tc_wgrad_gemm: NULL pointer");
    int64_t w[6];
    if (int rc = vd_tc_wgrad_plan(layer, plan, B, w)) return rc;
    const int splits = (int)w[0], sps = (int)w[1], ntiles = (int)w[2];
    WsParams p;
    memset(&p, 0, sizeof(p));
    p.n_u = 1; p.nu_total = 1; p.ug_count = 1; p.w_u_stride = 0;
    p.n_tiles = splits * ntiles; p.tiles_per_item = ntiles; p.v_count = ntiles;      // item = split-K slice, v = column tile
    p.item_stride = (int64_t)sps * 65536; p.u_stride = 0; p.v_stride = (int64_t)splits * sps * 65536;
    p.w_item_stride = (int64_t)sps * 8 * kWeightTileBytes;
    p.n_sa = sps; p.n_sb = 1; p.sa_stride = 65536; p.sb_stride = 0;
    p.n_copies = 1; p.copy_gofs[0] = 0; p.copy_sofs[0] = 0; p.copy_bytes[0] = 65536; p.stage_bytes = 65536;
    p.n_steps = 8;
    for (int j = 0; j < 8; ++j) { p.b_off16[j] = (uint32_t)(2 * j * 256); p.b_lbo16[j] = 256; }
    p.a_lbo16 = 2048 >> 4; p.a_sbo16 = 8;
    p.w_resident = 0; p.w_bytes = 0;
    p.G = 8; p.RW = 2; p.RP = 2;
    p.n_acc = 1; p.acc_delta16 = 0;
    p.ncols = 256; p.acc_cols = 256; p.acc_stages = 2;
    p.idesc = umma_idesc_bf16(128, 256);
    uint32_t smem = 0;
    if (int rc = finalize_smem(p, (uint32_t)p.G * p.RW * kWeightTileBytes, &smem, true)) return rc;
    const Geo g = make_geo(plan->T, plan->H);
    p.pix = (const uint8_t*)xcol; p.item_index = nullptr;
    p.epi.raw_bf16 = 0; p.epi.T = plan->T; p.epi.n_items = B; p.epi.g = g;
    // kt-split mode (tc_trio.cu): one GEMM per temporal tap over the same columns, against the frame-shifted gy image of that tap
    const int n_kt = vd_tc_wgrad_kt_mode(layer) ? 3 : 1;
    const int64_t img_bytes = (int64_t)splits * sps * 8 * kWeightTileBytes, raw_floats = (int64_t)ntiles * splits * 128 * 256;
    for (int kt = 0; kt < n_kt; ++kt) {
        p.wimg = (const uint8_t*)gyimg + kt * img_bytes;
        p.epi.raw = raw + kt * raw_floats;
        if (int rc = launch<EPI_RAW>(p, smem, (cudaStream_t)stream)) return rc;
    }
    return 0;
}

#ifdef VD_PROBE   // bring-up / tuning entry points: only in scripts/libvd_b200_probe.so (scripts/_probe_lib.py), never in the product library
// Bring-up / tuning probe: MMA issue rate of one layout variant, data content irrelevant.
//   hi words: SBO>>4 | 1<<14 | layout_type<<29 (0 none, 2 = 128B, 4 = 64B, 6 = 32B swizzle)
// 148 tiles x n_sa stages x n_steps MMAs of N = ncols; raw accumulators go to `raw`.
extern "C" int vd_tc_probe(const void* pix, const void* wimg, float* raw, int ncols, int n_sa, int n_steps,
                           uint32_t a_lbo16, uint32_t a_hi, uint32_t b_lbo16, uint32_t b_hi, uint32_t b_step16,
                           int n_acc, void* stream) {
    VD_REQUIRE(pix && wimg && raw && ncols % 16 == 0 && ncols >= 16 && ncols <= 256 && n_steps <= kMaxSteps, "tc_probe: bad argument");
    WsParams p;
    memset(&p, 0, sizeof(p));
    p.n_tiles = 148; p.tiles_per_item = 148; p.v_count = 148; p.item_stride = 0; p.u_stride = 0; p.v_stride = 0;
    p.n_u = 1; p.nu_total = 1; p.ug_count = 1; p.w_u_stride = 0;
    p.n_sa = n_sa; p.n_sb = 1; p.sa_stride = 0; p.sb_stride = 0;
    p.n_copies = 1; p.copy_gofs[0] = 0; p.copy_sofs[0] = 0; p.copy_bytes[0] = 65536; p.stage_bytes = 65536;
    p.n_steps = n_steps;
    for (int j = 0; j < n_steps; ++j) { p.b_off16[j] = (uint32_t)(j % 4) * b_step16; p.b_lbo16[j] = b_lbo16; p.a_off16[j] = 0; }
    p.a_sa_stride16 = 0; p.a_lbo16 = a_lbo16; p.a_sbo16 = 8; p.a_hi = a_hi; p.b_hi = b_hi;
    p.w_resident = 1; p.w_bytes = 16384; p.G = 1; p.RW = 1; p.RP = 2;
    p.n_acc = n_acc; p.acc_delta16 = 0; p.ncols = ncols; p.acc_cols = 256; p.acc_stages = (n_acc == 1) ? 2 : 1;
    p.idesc = umma_idesc_bf16(128, ncols);
    uint32_t smem = 0;
    if (int rc = finalize_smem(p, 16384, &smem)) return rc;
    p.pix = (const uint8_t*)pix; p.wimg = (const uint8_t*)wimg; p.epi.raw = raw;
    return launch<EPI_RAW>(p, smem, (cudaStream_t)stream);
}

// Tuning aid: when set (device buffer of grid*8 int64, may be NULL to disable), MMA issuer 0 of every forward conv
// launch records [total, wait acc_empty, wait pix_full, wait w_full, issue, wait baton] cycles per CTA.
extern "C" int vd_tc_set_profile_buffer(long long* buf) { g_prof = buf; return 0; }
#endif  // VD_PROBE
