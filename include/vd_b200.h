/*
 * vd_b200.h — C ABI of libvd_b200.so, the sm_100a implementation of the distillation inner
 * loop of yuz1wan/video_distillation.
 *
 * The reference is pure Python/PyTorch and has no FFI; each entry point below replaces the
 * ATen/cuDNN call the reference makes at the cited file:line (paths are relative to the
 * reference checkout).  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (the library never allocates or
 *     frees device memory and keeps no per-call state);
 *   - `stream` is a cudaStream_t passed as void*; every call is an asynchronous enqueue;
 *   - return value: 0 = ok, negative = bad argument / unsupported shape, positive = cudaError_t;
 *     vd_last_error() returns a thread-local message for the last non-zero return;
 *   - "f32 NCDHW" tensors are contiguous float32 in PyTorch's (N, C, T, H, W) order;
 *     videos at the Python API are (B, T, C, H, W) as in the reference (networks.py:739).
 */
#ifndef VD_B200_H
#define VD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- misc */
const char* vd_last_error(void);
int vd_abi_version(void);
/* number of kernels this library has enqueued in this process (bench.py `gpu_launches`) */
int64_t vd_launch_count(void);
void vd_launch_count_reset(void);

/* Convolution geometry shared by the three conv entry points. */
typedef struct {
    int32_t N, Cin, T, H, W;      /* input  (N, Cin, T, H, W)      */
    int32_t Cout, To, Ho, Wo;     /* output (N, Cout, To, Ho, Wo)  */
    int32_t kt, kh, kw;           /* filter extent                 */
    int32_t st, sh, sw;           /* stride                        */
    int32_t pt, ph, pw;           /* zero padding                  */
} vd_conv_geom;

/* ------------------------------------------------- exact fp32 conv trio (CUDA cores)
 * Replaces nn.Conv3d forward / convolution_backward at networks.py:799 (feature convs),
 * networks.py:736 (1x1x1 logit conv) and utils.py:1184 (composer conv), and — because the
 * three are mutually each other's derivatives — ATen's
 * _convolution_double_backward on the MTT path (distill_s2d_ms.py:264,292).
 *   fprop: y = conv(x, w) + bias           (bias may be NULL)
 *   dgrad: gx = conv_transpose(gy, w)      (explicit input extent in geom: stride-2 is not injective)
 *   wgrad: gw (+)= x (*) gy ; gb (+)= sum gy   (gw/gb must be zeroed by the caller; gb may be NULL)
 */
int vd_conv3d_fprop_f32(const float* x, const float* w, const float* bias, float* y,
                        const vd_conv_geom* g, void* stream);
int vd_conv3d_dgrad_f32(const float* gy, const float* w, float* gx,
                        const vd_conv_geom* g, void* stream);
int vd_conv3d_wgrad_f32(const float* x, const float* gy, float* gw, float* gb,
                        const vd_conv_geom* g, void* stream);

/* ------------------------------------------------- ReLU + MaxPool3d routing
 * Replaces nn.ReLU(inplace) networks.py:757 + nn.MaxPool3d networks.py:766-770 and their
 * backward / double backward.  kernel == stride == (pt, ph, pw) in {1,2}.
 * code[o] (uint8, one per pooled output): bits 0..2 = argmax position inside the window in
 * (t,h,w) scan order, first maximum wins (ATen's rule); bit 3 = "max > 0" (ReLU passes).
 *   fwd     : y = maxpool(relu(x)), writes code
 *   scatter : gx[src(o)] = code.active ? gy[o] : 0, all other gx = 0     (backward)
 *   gather  : y[o] = code.active ? x[src(o)] : 0                         (double backward)
 */
int vd_relu_maxpool_fwd_f32(const float* x, float* y, uint8_t* code, int64_t NC,
                            int T, int H, int W, int pt, int ph, int pw, void* stream);
int vd_route_scatter_f32(const float* gy, const uint8_t* code, float* gx, int64_t NC,
                         int T, int H, int W, int pt, int ph, int pw, void* stream);
int vd_route_gather_f32(const float* x, const uint8_t* code, float* y, int64_t NC,
                        int T, int H, int W, int pt, int ph, int pw, void* stream);

/* ------------------------------------------------- instancenorm / avgpool variant
 * GroupNorm(C, C, affine) networks.py:784 (+ReLU) and AvgPool3d(2,2) networks.py:772.
 *   inorm_relu_fwd: per (n,c) mean / rstd over T*H*W (eps 1e-5), y = relu(gamma*xhat+beta);
 *                   saves mean,rstd (N*C each).
 *   inorm_relu_bwd: gx from gy (gradient wrt y), plus ggamma/gbeta (+)= (caller zeroes).
 *   avgpool2_fwd/bwd: 2x2x2 mean, floor semantics (odd tails dropped).
 */
int vd_inorm_relu_fwd_f32(const float* x, const float* gamma, const float* beta, float* y,
                          float* mean, float* rstd, int N, int C, int64_t S, void* stream);
int vd_inorm_relu_bwd_f32(const float* x, const float* y, const float* gy, const float* gamma,
                          const float* mean, const float* rstd, float* gx, float* ggamma,
                          float* gbeta, int N, int C, int64_t S, void* stream);
/* instancenorm + ReLU + AvgPool3d(2) of networks.py:784,757,772 in one launch for frozen networks (forward only): x (N,C,T,H,W)
 * -> y (N,C,T/2,H/2,W/2); the normalised activation is never written.  mean / rstd (N*C each) may both be NULL. */
int vd_inorm_relu_avgpool_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                                  int N, int C, int T, int H, int W, void* stream);
int vd_avgpool2_fwd_f32(const float* x, float* y, int64_t NC, int T, int H, int W, void* stream);
int vd_avgpool2_bwd_f32(const float* gy, float* gx, int64_t NC, int T, int H, int W, void* stream);

/* ------------------------------------------------- static-dynamic composer
 * Replaces index gather + Conv3DNet.forward (utils.py:1186-1197; distill_s2d_ms.py:409-412,
 * 249-253) and its backward (index_put accumulate + convolution_backward) with one fused
 * kernel each.  static_syn (S, 3, H, W); dynamic_syn (C, dpc, T, 1, H, W); weight (3,4,3,3,3);
 * bias (3); per output video b: static row static_idx[b], dynamic row (label[b], dynamic_idx[b]).
 * out / gout: (B, T, 3, H, W) contiguous.
 *   bwd: grad_dynamic (dense, same shape as dynamic_syn, caller zeroes; rows are accumulated
 *        with atomics so repeated indices are legal), grad_weight (3*4*27), grad_bias (3)
 *        (caller zeroes), grad_static (S,3,H,W) optional (NULL to skip; caller zeroes).
 */
int vd_compose_fwd_f32(const float* static_syn, const float* dynamic_syn,
                       const int64_t* static_idx, const int64_t* label, const int64_t* dynamic_idx,
                       const float* weight, const float* bias, float* out,
                       int B, int T, int H, int W, int dpc, void* stream);
int vd_compose_bwd_f32(const float* gout, const float* static_syn, const float* dynamic_syn,
                       const int64_t* static_idx, const int64_t* label, const int64_t* dynamic_idx,
                       const float* weight, float* grad_dynamic, float* grad_weight,
                       float* grad_bias, float* grad_static,
                       int B, int T, int H, int W, int dpc, void* stream);
/* The same backward in ONE pass over the video gradient, without floating-point atomics on the 327 hallucinator sums: every
 * (video, 8-row band) block writes its sums to a row of `scratch` (>= B * ceil(H/8) * 328 floats) and a finishing launch adds
 * the rows in block order, so the hallucinator gradient is bitwise reproducible; grad_weight / grad_bias are accumulated
 * (+=), grad_dynamic rows are written with plain stores when unique_rows != 0 (no two videos select the same dynamic memory,
 * distill_s2d_ms.py:405) and with atomics otherwise.  grad_static is not produced (--no_train_static path). */
/* vd_compose_fwd_ex_f32 / vd_compose_bwd_fused_ex_f32: the same operations with the extents of the memories (n_static images of
 * (3,H,W), n_dynamic memories of (T,H,W) behind static_syn / dynamic_syn).  With them the kernels are fed by tensor-map TMA
 * (cp.async.bulk.tensor: halo rows / columns and the frames outside the clip zero-filled by the TMA unit); a tensor map must
 * describe the true allocation, so the entry points without extents keep to the cp.async kernels. */
int vd_compose_fwd_ex_f32(const float* static_syn, const float* dynamic_syn, const int64_t* static_idx, const int64_t* label,
                          const int64_t* dynamic_idx, const float* weight, const float* bias, float* out, int B, int T, int H,
                          int W, int dpc, int64_t n_static, int64_t n_dynamic, void* stream);
int vd_compose_bwd_fused_ex_f32(const float* gout, const float* static_syn, const float* dynamic_syn, const int64_t* static_idx,
                                const int64_t* label, const int64_t* dynamic_idx, const float* weight, float* grad_dynamic,
                                float* grad_weight, float* grad_bias, float* scratch, int64_t scratch_floats, int unique_rows,
                                int B, int T, int H, int W, int dpc, int64_t n_static, int64_t n_dynamic, void* stream);
int vd_compose_bwd_fused_f32(const float* gout, const float* static_syn, const float* dynamic_syn,
                             const int64_t* static_idx, const int64_t* label, const int64_t* dynamic_idx,
                             const float* weight, float* grad_dynamic, float* grad_weight, float* grad_bias,
                             float* scratch, int64_t scratch_floats, int unique_rows,
                             int B, int T, int H, int W, int dpc, void* stream);

/* ------------------------------------------------- distribution-matching loss
 * Replaces mean/sub/square/sum and their backward at distill_baseline.py:351,
 * distill_s2d_ms.py:422 for ALL classes of a shard in one launch.
 * emb_real (C, nr, D), emb_syn (C, ns, D).  loss (+)= sum_c ||mean_r - mean_s||^2 (caller
 * zeroes *loss); grad_syn (C, ns, D) = -(2/ns)(mean_r - mean_s) * loss_scale.
 * mean_real_out (C, D) optional (NULL to skip).
 */
int vd_class_mean_f32(const float* emb, float* mean, int C, int n, int D, void* stream);
/* Ragged form for the multi-GPU path (the sampled real videos of a class are spread over the ranks): sum[c,:] = sum of the rows
 * offsets[c] .. offsets[c+1]-1 of emb (n_rows, D), added in row order; offsets = device int32[C+1].  The (C, D) partial sums of
 * the ranks are all-reduced and divided by batch_real: torch.mean(output_real, dim=0) of distill_s2d_ms.py:422. */
int vd_class_sum_ragged_f32(const float* emb, const int32_t* offsets, float* sum, int C, int D, void* stream);
int vd_dm_loss_f32(const float* mean_real, const float* emb_syn, float* loss, float* grad_syn,
                   int C, int ns, int D, float loss_scale, void* stream);
/* Same result with a caller-provided scratch of C floats: one block per class writes class_loss[c] (the per-class terms of
 * distill_s2d_ms.py:414-422, also an output), a second launch adds them in class order into *loss.  Bitwise reproducible. */
int vd_dm_loss_ex_f32(const float* mean_real, const float* emb_syn, float* loss, float* grad_syn, float* class_loss,
                      int C, int ns, int D, float loss_scale, void* stream);

/* ------------------------------------------------- optimiser / flat-parameter kernels
 * sgd_momentum: torch.optim.SGD(momentum=m) dense step (distill_baseline.py:107,355;
 *   distill_s2d_ms.py:105-108,432-438): first ? buf=g : buf=m*buf+g ; p -= lr*buf.
 * axpy: y = a*x + y_in  (student update distill_baseline.py:252)
 * sqdist: out (+)= sum (a-b)^2 (mse_loss(reduction='sum') distill_baseline.py:258-259)
 */
int vd_sgd_momentum_f32(float* p, const float* g, float* buf, int64_t n, float lr, float momentum,
                        int first_step, void* stream);
int vd_axpy_f32(const float* x, const float* y_in, float* y_out, int64_t n, float a, void* stream);
int vd_sqdist_f32(const float* a, const float* b, float* out, int64_t n, void* stream);

/* =================================================================== tensor-core path
 * bf16 operands, fp32 accumulation in TMEM (tcgen05.mma), operands staged by bulk async
 * copies (cp.async.bulk -> UBLKCP) into shared memory in UMMA canonical K-major layout.
 * The three feature convolutions of ConvNet3D (networks.py:799; k=(3,7,7) s=(1,2,2)
 * p=(1,3,3)) are computed as "shifted-window" implicit GEMMs: weights are the M=128 operand,
 * output pixels the N operand; activations live in HBM in padded, parity-split,
 * channel-chunked bf16 layouts produced by the previous layer's fused epilogue
 * (bias + ReLU + MaxPool3d + argmax code).  See DESIGN.md §3 for the layouts.
 *
 * vd_tc_plan_* fill a plan struct on the host (sizes of every intermediate buffer);
 * the caller allocates, then calls the pack / run entry points.
 */
typedef struct {
    int32_t T, H, W;               /* input video extent (C=3)                               */
    int32_t c1, c2, c3;            /* channels of the three conv layers (64,128,128)         */
    int32_t T1, H1, W1;            /* L0 conv output extent  (T, H/2, W/2)                   */
    int32_t T1p, H1p, W1p;         /* after pool (1,2,2)                                     */
    int32_t T2, H2, W2, T2p, H2p, W2p;
    int32_t T3, H3, W3, T3p, H3p, W3p;
    int32_t embed_dim;             /* c3*T3p*H3p*W3p                                         */
    int64_t x0_bytes_per_video;    /* packed L0 input  (bf16, kw-expanded)                   */
    int64_t a1_bytes_per_video;    /* packed L1 input  (bf16, parity-planar)                 */
    int64_t a2_bytes_per_video;    /* packed L2 input  (bf16, tap-expanded)                  */
    int64_t w0_bytes, w1_bytes, w2_bytes;   /* UMMA weight images                            */
    int64_t tab_bytes;             /* unused (step tables travel as kernel parameters)       */
    /* backward (column-GEMM dgrad) operands, see vd_tc_bwd_* below */
    int64_t wt0_bytes, wt1_bytes, wt2_bytes;                 /* transposed weight images           */
    int64_t dy0_bytes_per_video, dy1_bytes_per_video, dy2_bytes_per_video;    /* packed output grads (bf16) */
    int64_t col0_bytes_per_video, col1_bytes_per_video, col2_bytes_per_video; /* bf16 column buffers        */
    int32_t reserved[8];
} vd_tc_plan;

int vd_tc_plan_make(vd_tc_plan* plan, int T, int H, int W);

/* fp32 (Bsrc,T,3,H,W) videos -> packed conv-0 operand X0 for B items (zero halo included; x0
 * need not be zeroed).  index (device int64[B], may be NULL = identity) selects the source
 * video of each item: the device-resident form of get_images (distill_s2d_ms.py:81-87). */
int vd_tc_pack_video(const float* video, const int64_t* index, void* x0, const vd_tc_plan* plan,
                     int B, void* stream);
/* Same packing from uint8 frames (B,T,3,H,W) with the dataset normalisation fused in: v = (u/255 - mean[c]) / std[c]
 * (reference: ToTensor + Normalize of the dataset transforms, utils.py:214-230), rounded to bf16.  mean3 / std3 are HOST
 * pointers to three floats.  Lets a host-resident real set cross PCIe as 1 byte per element (get_images, distill_s2d_ms.py:81-87). */
int vd_tc_pack_video_u8(const uint8_t* video, const int64_t* index, void* x0, const vd_tc_plan* plan, int B,
                        const float* mean3, const float* std3, void* stream);
/* fp32 OIDHW weights of features.{0,3,6} -> UMMA weight images (any pair may be NULL to skip).
 * w0 is the frame-pair-stacked form. */
int vd_tc_pack_weights(const float* w_l0, const float* w_l1, const float* w_l2,
                       void* w0, void* w1, void* w2, void* stream);

/* One layer: conv + bias + ReLU + MaxPool (+ optional argmax code), layer in {0,1,2}.
 *   layer 0: in = X0 (B items, or a bigger resident set addressed through item_index),
 *            out = A1 (B videos);   layer 1: in = A1, out = A2;
 *   layer 2: in = A2 (allocated for ceil(B/4)*4 videos), out = fp32 embeddings (B, embed_dim)
 *            in the reference's NCDHW flatten order (networks.py:750).
 * A1 / A2 must be zeroed ONCE by the caller before first use (halo cells are never written,
 * data cells are fully overwritten by every call).  code may be NULL (real videos need no
 * backward); otherwise items >= code_first_item record their routing codes at code[item - code_first_item]
 * (a batch of frozen real videos followed by differentiable synthetic ones runs as ONE launch).  item_index (device int64[B], layer 0 only, may be NULL) maps item -> video slot
 * of `in`.  raw == 1 (bring-up / tests): skip the fused epilogue and dump the fp32
 * accumulators to out as [tile][acc][128][ncols].  raw == 2: plain convolution, out = fp32 NCDHW
 * (B, Cout, To, Ho, Wo) pre-activation (+ bias when non-NULL) — F.conv3d of networks.py:799 as used by the
 * differentiable conv trio; raw == 3: the same, accumulated into out (out += result). */
int vd_tc_conv_layer(int layer, const void* in, const void* wimg, const float* bias,
                     void* out, uint8_t* code, int code_first_item, const vd_tc_plan* plan,
                     const int64_t* item_index, int B, int raw, void* stream);

/* ---- split-fp16 forward ("f16x3"): the parity mode of the fused pipeline.  Every operand value is carried as an fp16
 * pair v = hi + lo (22 significand bits) and every product as xh*wh + xl*wh + xh*wl (+ xl*wl in conv 0) accumulated in the
 * SAME fp32 TMEM accumulator, so embeddings equal an fp32 evaluation of networks.py:747-751 to ~3e-6 and ReLU / MaxPool
 * routing is decided on fp32-equivalent sums.  Same three fused layers, same epilogues, same routing codes as
 * vd_tc_conv_layer; the packed operands are twice as large (layouts X0s / A1s / A2s in csrc/tc_layout.h).
 *   vd_tc_x3_sizes         out[0..2] = X0s / A1s / A2s bytes per video, out[3..5] = weight image bytes of conv 0 / 1 / 2
 *   vd_tc_x3_pack_video    fp32 (Bsrc,T,3,H,W) videos (+ optional gather index) -> X0s          [get_images, distill_s2d_ms.py:81-87]
 *   vd_tc_x3_pack_video_u8 uint8 frames with (u/255 - mean[c]) / std[c] fused (HOST mean3 / std3) [utils.py:214-230]
 *   vd_tc_x3_pack_weights  fp32 OIDHW weights of features.{0,3,6} -> weight images (any pair may be NULL)
 *   vd_tc_x3_conv_layer    layer 0: X0s -> A1s; layer 1: A1s -> A2s; layer 2: A2s (ceil(B/4)*4 videos) -> fp32 embeddings
 *                          (A1s / A2s zeroed once by the caller; code / code_first_item / item_index as vd_tc_conv_layer) */
int vd_tc_x3_sizes(const vd_tc_plan* plan, int64_t* out);
int vd_tc_x3_pack_video(const float* video, const int64_t* index, void* x0s, const vd_tc_plan* plan, int B, void* stream);
int vd_tc_x3_pack_video_u8(const uint8_t* video, const int64_t* index, void* x0s, const vd_tc_plan* plan, int B,
                           const float* mean3, const float* std3, void* stream);
int vd_tc_x3_pack_weights(const float* w_l0, const float* w_l1, const float* w_l2, void* w0s, void* w1s, void* w2s,
                          void* stream);
int vd_tc_x3_conv_layer(int layer, const void* in, const void* wimg, const float* bias, void* out, uint8_t* code,
                        int code_first_item, const vd_tc_plan* plan, const int64_t* item_index, int B, void* stream);

/* Two-product mode of the FROZEN real videos (passes = 2; passes = 3 is vd_tc_x3_conv_layer).  The real branch of the DM loop
 * (distill_s2d_ms.py:416-419: output_real = embed(img_real).detach(), read only through torch.mean(output_real, dim=0)) needs no
 * routing codes and no gradient, and its embeddings enter the loss as a mean over batch_real videos.  Here every activation is
 * carried as ONE fp16 value and only the weights as an fp16 pair: y = xh*wh + xh*wl in the same TMEM accumulator — exact
 * weights, one rounding per activation (relative 2^-12, independent from video to video, so it averages out of the class
 * mean): 2/3 of the MMAs of conv 1 / conv 2, half of conv 0's stages, half of the activation bytes.  Same weight images.
 *   vd_tc_x3_pack_video_hi(_u8)  videos -> X0h: the X0 layout of vd_tc_pack_video (plan->x0_bytes_per_video) holding fp16 values
 *   vd_tc_x3_conv_layer_ex       layer 0: X0h -> hi planes of A1s; layer 1: hi planes of A1s -> hi chunks of A2s; layer 2: -> embeddings */
int vd_tc_x3_pack_video_hi(const float* video, const int64_t* index, void* x0h, const vd_tc_plan* plan, int B, void* stream);
int vd_tc_x3_pack_video_hi_u8(const uint8_t* video, const int64_t* index, void* x0h, const vd_tc_plan* plan, int B,
                              const float* mean3, const float* std3, void* stream);
/* One-launch split-fp16 fprop of the differentiable conv trio (layers 1 and 2): vd_tc_x3_pack_act packs an fp32 NCDHW activation
 * into A1s / A2s, vd_tc_x3_conv_plain writes the fp32 NCDHW pre-activation (+ bias when non-NULL) with all three products
 * accumulated in TMEM (weights from vd_tc_x3_pack_weights). */
int vd_tc_x3_pack_act(int layer, const float* x, void* packed, const vd_tc_plan* plan, int B, void* stream);
int vd_tc_x3_conv_plain(int layer, const void* in, const void* wimg, const float* bias, float* out, const vd_tc_plan* plan, int B,
                        void* stream);
int vd_tc_x3_conv_layer_ex(int layer, const void* in, const void* wimg, const float* bias, void* out, uint8_t* code,
                           int code_first_item, const vd_tc_plan* plan, const int64_t* item_index, int B, int passes, void* stream);

/* ---- backward of the tensor-core embed (gradient to the input video; weights are frozen in DM).
 * dgrad of each conv is a plain GEMM on tensor cores, col[(ci,tap), pixel] = sum_co W[co,ci,tap] *
 * dY[co,pixel] (same ws_gemm kernel: transposed weights = M operand, packed dY = N operand), followed
 * by a memory-bound col2im gather that also applies the ReLU/MaxPool routing code of the layer below
 * and re-packs the result as the dY operand of the next GEMM.  Replaces convolution_backward /
 * max_pool3d_with_indices_backward / threshold_backward on the synthetic branch
 * (distill_s2d_ms.py:431).
 *   pack_weights_bwd : fp32 OIDHW -> transposed UMMA images [m-tile][step][k 2][128][8]
 *   bwd_emb    : g_emb (B, embed_dim) fp32 + code2 -> dy2
 *   bwd_gemm   : layer in {2,1,0}: dy_layer x wt_layer -> col_layer (bf16 [video][ntile][mtile][128][NC])
 *   bwd_col2im : layer 2: col2 + code1 -> dy1;  layer 1: col1 + code0 -> dy0;
 *                layer 0: col0 -> d video (B, T, 3, H, W) fp32 (code must be NULL)
 */
int vd_tc_pack_weights_bwd(const float* w_l0, const float* w_l1, const float* w_l2,
                           void* wt0, void* wt1, void* wt2, void* stream);
int vd_tc_bwd_emb(const float* g_emb, const uint8_t* code2, void* dy2, const vd_tc_plan* plan,
                  int B, void* stream);
int vd_tc_bwd_gemm(int layer, const void* dy, const void* wt, void* col, const vd_tc_plan* plan,
                   int B, void* stream);
int vd_tc_bwd_col2im(int layer, const void* col, const uint8_t* code_below, void* out,
                     const vd_tc_plan* plan, int B, void* stream);

/* ---- the differentiable conv trio on tensor cores (MTT unroll / double backward, ops.py).  Plain fp32 NCDHW
 * tensors in and out; geometry = the three ConvNet3D feature convolutions of `plan`.
 *   fprop(x, w)   : layer 0: vd_tc_pack_video on x permuted to (B,T,3,H,W); layers 1/2: vd_tc_pack_act;
 *                   then vd_tc_conv_layer(..., raw = 2) -> y (B, Cout, To, Ho, Wo)      [F.conv3d, networks.py:799]
 *   dgrad(gy, w)  : vd_tc_pack_dy, vd_tc_pack_weights_bwd, vd_tc_bwd_gemm, vd_tc_bwd_col2im_plain -> gx NCDHW
 *   wgrad(x, gy)  : vd_tc_wgrad_plan (workspace sizes), vd_tc_wgrad_pack (im2col columns + gy image),
 *                   vd_tc_wgrad_gemm (split-K partial sums), vd_tc_wgrad_reduce -> gw (Cout, Cin, 3, 7, 7)   */
int vd_tc_pack_act(int layer, const float* x, void* packed, const vd_tc_plan* plan, int B, int part, void* stream);
/* part = 0: the value, 1: its bf16 residual v - bf16(v) (operands of the split-bf16 fprop) */
int vd_tc_pack_video_ncdhw(const float* x, void* x0, const vd_tc_plan* plan, int B, int part, void* stream);
int vd_tc_pack_weights_part(const float* w_l0, const float* w_l1, const float* w_l2, void* w0, void* w1, void* w2,
                            int part, void* stream);
int vd_tc_pack_dy(int layer, const float* gy, void* dy, const vd_tc_plan* plan, int B, void* stream);
/* col_fp32: the column buffer is fp32 (written by vd_tc_bwd_gemm_ex(..., col_fp32 = 1)) instead of bf16 */
int vd_tc_bwd_col2im_plain(int layer, const void* col, float* gx, const vd_tc_plan* plan, int B, int col_fp32, void* stream);
int vd_tc_bwd_gemm_ex(int layer, const void* dy, const void* wt, void* col, const vd_tc_plan* plan, int B, int col_fp32,
                      void* stream);
/* out[0] split-K slices, out[1] stages per slice, out[2] column tiles, out[3] xcol bytes, out[4] gyimg bytes, out[5] raw bytes */
int vd_tc_wgrad_plan(int layer, const vd_tc_plan* plan, int B, int64_t* out);
/* 1 when wgrad of `layer` runs kt-split (default; env VD_TC_WGRAD_KT = bit mask over the layers): the three temporal taps are
 * three GEMMs over ONE im2col of the 49 spatial taps, each against a frame-shifted image of gy; plan / pack / gemm / reduce
 * follow the mode by themselves, callers only size their buffers from vd_tc_wgrad_plan. */
int vd_tc_wgrad_kt_mode(int layer);
int vd_tc_wgrad_pack(int layer, const float* x, const float* gy, void* xcol, void* gyimg, const vd_tc_plan* plan,
                     int B, void* stream);
/* The split (parity-grade) wgrad of the MTT trio: x_part / gy_part = 0 packs the value (rounded to bf16), 1 its bf16 residual
 * v - bf16(v); x or gy NULL keeps the operand that is already in xcol / gyimg (xl*gh, xh*gh, xh*gl need two im2cols and two
 * gy images, not three).  [torch.autograd conv weight gradient of networks.py:799, reference fp32] */
int vd_tc_wgrad_pack_parts(int layer, const float* x, int x_part, const float* gy, int gy_part, void* xcol, void* gyimg,
                           const vd_tc_plan* plan, int B, void* stream);
int vd_tc_wgrad_gemm(int layer, const void* xcol, const void* gyimg, float* raw, const vd_tc_plan* plan, int B,
                     void* stream);
int vd_tc_wgrad_reduce(int layer, const float* raw, float* gw, const vd_tc_plan* plan, int B, void* stream);

/* ---- direct (column-free) dgrad of conv 1: a shifted-window GEMM per input-row parity over the "padded planar" dY
 * of conv 1 (tc_layout.h: Dg1Geo).  Replaces vd_tc_bwd_gemm(1) + vd_tc_bwd_col2im(1) (59 MB of column buffer per video).
 *   vd_tc_dgrad1_sizes        out[0] dYP bytes per video, out[1..2] weight image bytes (ph = 0, 1)
 *   vd_tc_pack_dgrad1_weights fp32 OIDHW weights of features.3 -> the two weight images
 *   vd_tc_pack_dyp1           fp32 NCDHW gy (B,128,T,Ho1,Wo1) -> dYP (every cell written)
 *   vd_tc_bwd_col2im_ex       layer 2, out_layout 1: col2im + routing of conv 2's columns straight into dYP (halo pre-zeroed)
 *   vd_tc_dgrad1              code0 != NULL: out = packed dY of conv 0's column GEMM (ReLU/MaxPool routing of conv 0 applied)
 *                             code0 == NULL: out = fp32 NCDHW (B, 64, T, H1, H1) */
int vd_tc_dgrad1_sizes(const vd_tc_plan* plan, int64_t* out);
int vd_tc_pack_dgrad1_weights(const float* w_l1, void* wimg0, void* wimg1, const vd_tc_plan* plan, void* stream);
int vd_tc_pack_dyp1(const float* gy, void* dyp, const vd_tc_plan* plan, int B, void* stream);
int vd_tc_bwd_col2im_ex(int layer, const void* col, const uint8_t* code_below, void* out, const vd_tc_plan* plan, int B,
                        int out_layout, void* stream);
int vd_tc_dgrad1(const void* dyp, const void* wimg0, const void* wimg1, const uint8_t* code0, void* out,
                 const vd_tc_plan* plan, int B, void* stream);
/* out_planar != 0: the routed gradient goes to the padded planar dY of conv 0 (input of vd_tc_dgrad0; halo pre-zeroed) */
int vd_tc_dgrad1_ex(const void* dyp, const void* wimg0, const void* wimg1, const uint8_t* code0, void* out,
                    const vd_tc_plan* plan, int B, int out_planar, void* stream);

/* ---- direct dgrad of conv 0 (Cin = 3): pixels on M (128 per tile), the 12 (ci, row parity, column parity) outputs on
 * N = 16, resident 96 KiB weight image; replaces vd_tc_bwd_gemm(0) + vd_tc_bwd_col2im(0) (51 MB of columns per video).
 *   vd_tc_dgrad0_sizes        out[0] dYP0 bytes per video, out[1] weight image bytes
 *   vd_tc_pack_dgrad0_weights fp32 OIDHW weights of features.0 -> weight image
 *   vd_tc_pack_dyp0           fp32 NCDHW gy (B,64,T,Ho0,Wo0) -> dYP0 (every cell written)
 *   vd_tc_dgrad0              out = fp32 gradient w.r.t. the input video, (B,T,3,H,W) (ncdhw = 0) or (B,3,T,H,W) */
int vd_tc_dgrad0_sizes(const vd_tc_plan* plan, int64_t* out);
int vd_tc_pack_dgrad0_weights(const float* w_l0, void* wimg, void* stream);
int vd_tc_pack_dyp0(const float* gy, void* dyp, const vd_tc_plan* plan, int B, void* stream);
int vd_tc_dgrad0(const void* dyp0, const void* wimg, float* out, const vd_tc_plan* plan, int B, int ncdhw, void* stream);

/* Split-bf16 dgrad helpers (the parity backward evaluates every dgrad as dgrad(gh, wh) + dgrad(gh, wl) + dgrad(gl, wh), g = gh + gl,
 * w = wh + wl in bf16 parts): the packers make part 0 = bf16(gy) / part 1 = bf16(gy - bf16(gy)) from the fp32 gradient itself,
 * and the column-free dgrads can accumulate into their fp32 output (accumulate != 0: out += result), so the three passes share
 * one tensor without elementwise launches in between. */
int vd_tc_pack_dy_part(int layer, const float* gy, void* dy, const vd_tc_plan* plan, int B, int part, void* stream);
int vd_tc_pack_dyp1_part(const float* gy, void* dyp, const vd_tc_plan* plan, int B, int part, void* stream);
int vd_tc_pack_dyp0_part(const float* gy, void* dyp, const vd_tc_plan* plan, int B, int part, void* stream);
int vd_tc_dgrad1_plain(const void* dyp, const void* wimg0, const void* wimg1, float* out, const vd_tc_plan* plan, int B,
                       int accumulate, void* stream);
int vd_tc_dgrad0_ex(const void* dyp0, const void* wimg, float* out, const vd_tc_plan* plan, int B, int ncdhw, int accumulate,
                    void* stream);

/* Host-only introspection (no GPU work): the launch parameters vd_tc_conv_layer would use,
 * flattened to int64 (layout documented in tests/tc_emulator.py); cap >= 248. */
int vd_tc_debug_params(int layer, const vd_tc_plan* plan, int B, int64_t* out, int cap);   /* layer 3,4,5 = bwd gemm of conv 0,1,2; 6,7,8 = split-fp16 conv 0,1,2 */

#ifdef __cplusplus
}
#endif
#endif /* VD_B200_H */
