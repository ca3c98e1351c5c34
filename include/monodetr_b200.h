/*
 * monodetr_b200.h -- C ABI of libmonodetr_b200.so (B200 / sm_100a).
 *
 * Plain pointers and sizes only; every pointer is a DEVICE pointer unless its comment says HOST.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Every entry point is
 * asynchronous on `stream`, never synchronises, never creates streams, and returns 0 on success,
 * a positive cudaError_t value if the launch failed, or a negative MDB_E* code for bad arguments.
 * (The reference only printf()s launch errors -- ms_deform_im2col_cuda.cuh:948-952,1321-1325 -- the
 * host side of this library turns a non-zero return into a Python RuntimeError instead.)
 *
 * Reference interface each group replaces (paths relative to /root/reference):
 *   mdb_msda_*          lib/models/monodetr/ops/src/vision.cpp:13-16  (pybind ms_deform_attn_forward/backward)
 *                       lib/models/monodetr/ops/src/ms_deform_attn.h:20-61 (dispatch)
 *                       lib/models/monodetr/ops/src/cuda/ms_deform_attn_cuda.cu:20-80, 83-153 (host launchers)
 *                       lib/models/monodetr/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299, 301-403 (kernels)
 */
#ifndef MONODETR_B200_H_
#define MONODETR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDB_EINVAL (-1)      /* bad size / null pointer / misaligned pointer */
#define MDB_EUNSUPPORTED (-2) /* shape outside what the kernels implement */
#define MDB_EWORKSPACE (-3)   /* scratch buffer missing / too small: see mdb_set_workspace */

/* Library/ABI version (bumped when a signature changes). */
int mdb_abi_version(void);
/* Human-readable name of the last error code returned on this thread ("ok" if 0). HOST string. */
const char* mdb_error_string(int code);

/* Reproducible accumulation (process-wide; default 0).  The reference scatters the MSDeformAttn value gradient with atomicAdd
 * (ms_deform_im2col_cuda.cuh:125-152) and so does the default path here (vector reductions), as does the split-K weight gradient
 * of the tensor-core family: results then differ in the last bits from run to run.  With 1:
 *   - mdb_msda_backward_f32 / _f64 accumulate grad_value in a FIXED order (one thread owns each element and adds its
 *     contributions in (query, point, corner) order with plain read-modify-writes; grad_loc / grad_attn never use atomics),
 *   - mdb_msda_fused_backward_f32 returns MDB_EUNSUPPORTED (callers take mdb_msda_prep_* + mdb_msda_backward_*),
 *   - mdb_conv2d_wgrad_* run without split-K (exactly one accumulation per output element),
 * so that two runs on the same inputs give the same bits.  A test / debugging mode: the ordered scatter walks every query serially per (image, head, level): 32 ms instead of 0.95 ms at B = 8, Lq = 10 200.
 * (Per-channel parameter gradients of the norm layers and the GroupNorm statistics still combine CTA partials atomically.) */
int mdb_set_deterministic(int on);
int mdb_get_deterministic(void);

/* ---- Multi-scale deformable attention (MSDeformAttn core) -----------------------------------
 * value          (B, S, M, D)          contiguous
 * spatial_shapes (L, 2) int64 (H_l, W_l), device        [ms_deform_attn_cuda.cu:28-38 asserts the same]
 * level_start    (L,)   int64, device
 * sampling_loc   (B, Lq, M, L, P, 2)   (x, y) normalised to [0,1]
 * attn_weight    (B, Lq, M, L, P)
 * out            (B, Lq, M*D)          fully written by the call
 * Any B is accepted (the reference's im2col_step chunking, ms_deform_attn_cuda.cu:50-52, is not needed).
 */
int mdb_msda_forward_f32(const float* value, const int64_t* spatial_shapes, const int64_t* level_start,
                         const float* sampling_loc, const float* attn_weight,
                         int B, int S, int M, int D, int L, int Lq, int P,
                         float* out, void* stream);
int mdb_msda_forward_f64(const double* value, const int64_t* spatial_shapes, const int64_t* level_start,
                         const double* sampling_loc, const double* attn_weight,
                         int B, int S, int M, int D, int L, int Lq, int P,
                         double* out, void* stream);

/* Backward.  grad_value (B,S,M,D) is zero-filled by the call and then accumulated with atomics
 * (ms_deform_attn_cuda.cu:121-123 + cuh:125-152); grad_loc / grad_attn are fully written. */
int mdb_msda_backward_f32(const float* value, const int64_t* spatial_shapes, const int64_t* level_start,
                          const float* sampling_loc, const float* attn_weight, const float* grad_out,
                          int B, int S, int M, int D, int L, int Lq, int P,
                          float* grad_value, float* grad_loc, float* grad_attn, void* stream);
int mdb_msda_backward_f64(const double* value, const int64_t* spatial_shapes, const int64_t* level_start,
                          const double* sampling_loc, const double* attn_weight, const double* grad_out,
                          int B, int S, int M, int D, int L, int Lq, int P,
                          double* grad_value, double* grad_loc, double* grad_attn, void* stream);

/* Fused MSDeformAttn pre-processing (ops/modules/ms_deform_attn.py:145-155): sampling_locations and
 * softmax(attention logits) from the raw projections in one pass, and its backward.
 * off (B,Lq,M,L,P,2), logits (B,Lq,M,L*P), ref (B,Lq,L,ref_dim) contiguous, ref_dim in {2,6}; L*P <= 16.
 * backward: doff / dlogits from dloc / dattn (+ saved attn); the gradient wrt ref (only decoder layer 0 needs it)
 * is a plain reduction of dloc done by the caller. */
int mdb_msda_prep_forward_f32(const float* off, const float* logits, const float* ref, const int64_t* spatial_shapes,
                              int B, int Lq, int M, int L, int P, int ref_dim, float* loc, float* attn, void* stream);
int mdb_msda_prep_backward_f32(const float* dloc, const float* dattn, const float* attn, const float* ref,
                               const int64_t* spatial_shapes, int B, int Lq, int M, int L, int P, int ref_dim,
                               float* doff, float* dlogits, void* stream);

/* The module's forward with the pre-processing INSIDE the sampling kernels (constant reference points; D = 32, L = 4, P = 4,
 * otherwise MDB_EUNSUPPORTED): offsets (B,Lq,M,L,P,2) and logits (B,Lq,M,L*P) are the raw projections, ref (B,Lq,L,ref_dim).
 * backward: grad_value zero-filled then accumulated; grad_offsets / grad_logits fully written. */
int mdb_msda_fused_forward_f32(const float* value, const int64_t* spatial_shapes, const int64_t* level_start, const float* offsets,
                               const float* logits, const float* ref, int B, int S, int M, int D, int L, int Lq, int P, int ref_dim,
                               float* out, void* stream);
int mdb_msda_fused_backward_f32(const float* value, const int64_t* spatial_shapes, const int64_t* level_start, const float* offsets,
                                const float* logits, const float* ref, const float* grad_out, int B, int S, int M, int D, int L, int Lq,
                                int P, int ref_dim, float* grad_value, float* grad_offsets, float* grad_logits, void* stream);

/* ---- Tensor-core convolution / linear family (tcgen05 + TMEM + TMA, fp32 storage, TF32 math) ----
 * Replaces the cuDNN / cuBLAS calls behind nn.Conv2d / nn.Linear on the reference path
 * (backbone.py:100-102; monodetr.py:83-91; depth_predictor/depth_predictor.py:29-47;
 *  ops/modules/ms_deform_attn.py:138-161; depthaware_transformer.py:339-343,467-473).
 * Activations are NHWC fp32: x[B][H][W][Cin], y[B][Ho][Wo][Cout]; weights are "packed"
 * [kh*kw][Cout][Cin] (mdb_pack_conv_weight_f32).  A linear layer y[M,N] = x[M,K] w[N,K]^T is the call
 * with B=1, H=1, W=M, Cin=K, Cout=N, kh=kw=1, stride=1, pad=0 (w itself is already "packed").
 * Supported: kh=kw in {1,3}, stride in {1,2}, Cin%4==0, 16-byte aligned pointers; Cout%4==0 for dgrad / wgrad (the
 * forward writes any Cout, e.g. the 3-class / 81-bin head widths of monodetr.py:102-117); outputs below 2^31 elements.
 * Results do not depend on the batch size (same image -> same bits) and are bit-reproducible run to run for
 * forward / dgrad; wgrad accumulates its split-K partial sums with fp32 atomics.  The one forward shape with
 * >= 256 k-blocks and few tiles (3x3 stride-2 2048->256, monodetr.py:83-91) reduces through the scratch buffer the caller
 * registers with mdb_set_workspace (per device; the library itself never allocates or frees device memory).
 */
/* Arithmetic of the tensor-core family (a process-wide numerical setting): 2 (default) = error-compensated BF16x3 (operands
 * split into bf16 hi + lo, A_hi*B_hi + A_lo*B_hi + A_hi*B_lo at twice the TF32 tensor rate, dropped terms ~2^-17 per product;
 * forward / dgrad take pre-split weights through mdb_pack_gemm_weights_bf16x3 and the *_bf16x3 entry points, wgrad splits both
 * activations on the fly; the _f32 forward / dgrad entry points run 3xTF32 in this mode); 1 = error-compensated 3xTF32
 * (A*B + A_lo*B + A*B_lo, ~fp32 accuracy); 0 = single-pass TF32 with round-to-nearest operands (cuDNN's allow_tf32 class). */
int mdb_set_precision(int mode);
int mdb_get_precision(void);
/* Split-K scratch of mdb_conv2d_forward_* for the CURRENT device (cudaGetDevice): the library never allocates device
 * memory.  Ask _workspace_bytes (0 = none needed; negative = MDB_E*), register a buffer at least that large that stays
 * valid while launches (or captured CUDA graphs) may use it; otherwise those shapes return MDB_EWORKSPACE. */
int mdb_set_workspace(void* buf, unsigned long long bytes);
long long mdb_conv2d_forward_workspace_bytes(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad,
                                             int flags, int has_residual, int split_weights);
/* Precision mode 2: weights split once per step into bf16 (hi, lo) pairs.  wf[tap][Cout][ceil(Cin/32)][hi 32 | lo 32] is the
 * forward operand, wd[tap][Cin][ceil(Cout/32)][hi 32 | lo 32] (the transposed weight) the dgrad operand; sizes in bf16
 * elements: taps*Cout*ceil(Cin/32)*64 and taps*Cin*ceil(Cout/32)*64.  n tensors per call, HOST arrays; scale (FrozenBN
 * fold, backbone.py:54-64) and wd may be NULL or hold NULL entries.  src_packed: 0 = OIHW sources (nn.Conv2d / nn.Linear
 * weights as stored), 1 = [tap][Cout][Cin] sources.  taps <= 9. */
int mdb_pack_gemm_weights_bf16x3(int n, const float* const* w, const float* const* scale, void* const* wf, void* const* wd,
                                 const int* O, const int* I, const int* taps, int src_packed, void* stream);
int mdb_conv2d_forward_bf16x3(const float* x, const void* w_split, const float* bias, const float* residual, float* y,
                              int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int flags,
                              void* stream);
int mdb_conv2d_dgrad_bf16x3(const float* dy, const void* w_split_t, const float* residual, const float* relu_mask,
                            float* dx, int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad,
                            int flags, void* stream);
int mdb_conv2d_forward_f32(const float* x, const float* w_packed, const float* bias /*[Cout]|NULL*/,
                           const float* residual /*like y|NULL*/, float* y,
                           int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad,
                           int flags /* bit0: ReLU, bit1: store y rounded-to-nearest TF32 */, void* stream);
/* dx = (conv_transpose(dy, w) + residual) * (relu_mask > 0); residual / relu_mask are shaped like dx or NULL. */
int mdb_conv2d_dgrad_f32(const float* dy, const float* w_packed, const float* residual, const float* relu_mask,
                         float* dx, int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad,
                         int flags /* bit1: store dx rounded-to-nearest TF32 */, void* stream);
/* dw_packed[tap][Cout][Cin] (+)= rowscale[co] * sum dy * x ; zero-filled first unless accumulate. */
int mdb_conv2d_wgrad_f32(const float* dy, const float* x, const float* rowscale /*[Cout]|NULL*/, float* dw_packed,
                         int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad,
                         int accumulate, void* stream);
/* Same, plus db[Cout] (+)= sum over output pixels of dy: the bias gradient of nn.Conv2d / nn.Linear, produced by the
 * same launch (zero-filled first unless accumulate; db may sit directly behind dw_packed to share the memset). */
int mdb_conv2d_wgrad_bias_f32(const float* dy, const float* x, const float* rowscale /*[Cout]|NULL*/, float* dw_packed,
                              float* db /*[Cout]|NULL*/, int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride,
                              int pad, int accumulate, void* stream);
/* w_packed[t][o][i] = w_oihw[o][i][t] * (scale ? scale[o] : 1), rounded to nearest TF32 in precision mode 0
 * (FrozenBatchNorm fold, backbone.py:54-64) */
int mdb_pack_conv_weight_f32(const float* w_oihw, const float* scale, float* w_packed, int O, int I, int taps,
                             void* stream);
int mdb_unpack_conv_wgrad_f32(const float* dw_packed, float* dw_oihw, int O, int I, int taps, int accumulate,
                              void* stream);
/* Multi-tensor forms of the two calls above: n tensors per call (one launch per 64); the array arguments are HOST arrays
 * (scale may be NULL, or hold NULL entries). */
int mdb_pack_conv_weights_multi_f32(int n, const float* const* w_oihw, const float* const* scale, float* const* w_packed,
                                    const int* O, const int* I, const int* taps, void* stream);
int mdb_unpack_conv_wgrads_multi_f32(int n, const float* const* dw_packed, float* const* dw_oihw, const int* O, const int* I,
                                     const int* taps, void* stream);
/* out[n] (+)= sum_m x[m][n]  (bias gradients) */
int mdb_colsum_f32(const float* x, float* out, long long M, int N, int accumulate, void* stream);

/* ---- Fused multi-head attention core, head_dim 32 (attention.cu) -------------------------------------
 * Replaces the core of torch's F.multi_head_attention_forward as called at depthaware_transformer.py:456-459,
 * :496 and depth_predictor/transformer.py:59.  q[b][i][h][32] with token stride ldq floats (batch stride
 * Lq*ldq), k/v likewise (Lk*ldk, Lk*ldv), out[b][i][h*32] with token stride ldo.  key_padding_mask [B][Lk]
 * bytes (nonzero = ignore) or NULL.  lse [B][H][Lq] is written by forward and read by backward.
 * Dropout on the probabilities: drop_p in [0,1); *seed is read on the device (CUDA-graph safe); `site`
 * decorrelates call sites.  delta_ws: B*H*Lq floats of workspace.
 */
int mdb_attention_forward_f32(const float* q, const float* k, const float* v, const unsigned char* key_padding_mask,
                              float* out, float* lse, int B, int H, int Lq, int Lk, int head_dim,
                              int ldq, int ldk, int ldv, int ldo, float drop_p,
                              const unsigned long long* seed, unsigned long long site, void* stream);
int mdb_attention_backward_f32(const float* q, const float* k, const float* v, const unsigned char* key_padding_mask,
                               const float* out, const float* lse, const float* dout, float* delta_ws,
                               float* dq, float* dk, float* dv, int B, int H, int Lq, int Lk, int head_dim,
                               int ldq, int ldk, int ldv, int ldo, int lddq, int lddk, int lddv, float drop_p,
                               const unsigned long long* seed, unsigned long long site, void* stream);

/* ---- Normalisation (norm.cu) ---------------------------------------------------------------------------
 * y = LayerNorm(x + dropout(res)) * gamma + beta, rows of C floats (C in {128,256,512}); res may be NULL.
 * Replaces the `src = norm(src + dropout(src2))` pattern of depthaware_transformer.py:341-349,461-462,502-513
 * and depth_predictor/transformer.py:60-65.  mean / rstd: [M] saved for backward.
 * backward: dx = grad wrt x (and wrt res when drop_p == 0); dres (may be NULL iff drop_p == 0) = grad wrt res.
 */
int mdb_add_layernorm_forward_f32(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                                  float* mean, float* rstd, long long M, int C, float eps, float drop_p,
                                  const unsigned long long* seed, unsigned long long site, void* stream);
int mdb_add_layernorm_backward_f32(const float* dy, const float* x, const float* res, const float* gamma,
                                   const float* mean, const float* rstd, float* dx, float* dres, float* dgamma,
                                   float* dbeta, long long M, int C, float drop_p, const unsigned long long* seed,
                                   unsigned long long site, int accumulate, void* stream);
/* GroupNorm(G, C) on NHWC x[B][HW][C] (+ optional fused ReLU) -- monodetr.py:83-91, depth_predictor.py:29-45.
 * stats_ws: B*G*2 doubles; mean / rstd: [B][G]. */
int mdb_groupnorm_forward_f32(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                              double* stats_ws, int B, int HW, int C, int G, float eps, int relu, void* stream);
int mdb_groupnorm_backward_f32(const float* dy, const float* x, const float* y, const float* gamma, const float* mean,
                               const float* rstd, float* dx, float* dgamma, float* dbeta, double* stats_ws,
                               int B, int HW, int C, int G, int relu, void* stream);

/* ---- Elementwise helpers and the frozen ResNet stem (elementwise.cu) ------------------------------------ */
int mdb_relu_backward_f32(const float* dy, const float* y, float* out, long long n, float scale, void* stream);
int mdb_dropout_f32(const float* x, float* out, long long n, float p, const unsigned long long* seed,
                    unsigned long long site, void* stream);
int mdb_round_tf32_f32(const float* x, float* out, long long n, void* stream);
/* backbone.py:100-102 conv1 + bn1 (FrozenBatchNorm2d, backbone.py:54-64) + relu; x NCHW [B][3][H][W] -> y NHWC [B][H/2][W/2][64] */
int mdb_stem_conv7x7_bn_relu_f32(const float* x, const float* w, const float* scale, const float* bias, float* y,
                                 int B, int H, int W, void* stream);
/* torchvision ResNet maxpool(3, 2, 1) on NHWC */
int mdb_maxpool3x3s2_nhwc_f32(const float* x, float* y, int B, int H, int W, int C, void* stream);

/* Depth-map lookup of the head (monodetr.py:248-253): grid_sample(bilinear, zeros, align_corners=True) of
 * depth (B,H,W) at xy (B,N,2) in [-1,1] -> out (B,N); backward wrt the map only (centres are detached). */
int mdb_depth_sample_forward_f32(const float* depth, const float* xy, float* out, int B, int H, int W, int N, void* stream);
int mdb_depth_sample_backward_f32(const float* dout, const float* xy, float* ddepth, int B, int H, int W, int N, void* stream);

/* ---- Fused elementwise chains around the heads and the depth predictor's tail (heads.cu) -----------------------------------
 * box refinement (depthaware_transformer.py:602-613): y[n][6] = sigmoid(tmp + inverse_sigmoid(ref)) on the first ref_dim (2 or 6)
 * components, inverse_sigmoid as utils/misc.py:473-477; backward: dtmp, and dref when non-NULL. */
int mdb_box_refine_forward_f32(const float* tmp, const float* ref, float* y, long long n, int ref_dim, void* stream);
int mdb_box_refine_backward_f32(const float* dy, const float* y, const float* ref, float* dtmp, float* dref /*or NULL*/, long long n,
                                int ref_dim, void* stream);
/* depth of a query (monodetr.py:230-262): out[b][q] = ((1/(sigmoid(reg0)+1e-6) - 1) + size3d0 / clamp((c4+c5)*img_h, 1) * fu +
 * grid_sample(weighted_depth, (c01 - 0.5)*2, bilinear, zeros, align_corners=True)) / 3 , reg1.  coord (B,N,6), size3d (B,N,3),
 * depth_reg (B,N,2), wdepth (B,H,W), calibs (B,3,4), img_sizes (B,2) = [W, H].  backward: dwdepth is zero-filled by the call. */
int mdb_head_depth_forward_f32(const float* coord, const float* size3d, const float* depth_reg, const float* wdepth, const float* calibs,
                               const float* img_sizes, float* out, int B, int N, int H, int W, void* stream);
int mdb_head_depth_backward_f32(const float* dout, const float* coord, const float* size3d, const float* depth_reg, const float* calibs,
                                const float* img_sizes, float* dcoord, float* dsize3d, float* dreg, float* dwdepth, int B, int N, int H,
                                int W, void* stream);
/* depth predictor tail (depth_predictor.py:74-104): per pixel softmax over nb (<= 96) bin logits, weighted depth = sum p * bins,
 * ip = lerp of the embedding rows floor / floor+1 of clamp(depth, 0, dmax).  logits (npix, nb), emb (E, C), C % 4 == 0, C <= 256.
 * backward: d_wd_ext (npix) = gradient reaching weighted_depth from elsewhere, or NULL; demb is zero-filled by the call. */
int mdb_depth_tail_forward_f32(const float* logits, const float* bins, const float* emb, float* wdepth, float* ip, long long npix, int nb,
                               int E, int C, float dmax, void* stream);
int mdb_depth_tail_backward_f32(const float* logits, const float* bins, const float* emb, const float* d_ip, const float* d_wd_ext,
                                float* dlogits, float* demb, long long npix, int nb, int E, int C, float dmax, void* stream);
/* (a + b + c) / 3 and a * s; n % 4 == 0 */
int mdb_mean3_f32(const float* a, const float* b, const float* c, float* out, long long n, void* stream);
int mdb_scale_f32(const float* a, float* out, long long n, float s, void* stream);
/* loss = sum_k mean(x_k^2) over `count` (<= 32) tensors (the surrogate loss of bench.py's step) and its gradient
 * g_k = 2 x_k / n_k * dloss; x / g / n are HOST arrays, loss / dloss device scalars. */
int mdb_sum_mean_squares_forward_f32(int count, const float* const* x, const long long* n, float* loss, void* stream);
int mdb_sum_mean_squares_backward_f32(int count, const float* const* x, float* const* g, const long long* n, const float* dloss,
                                      void* stream);

/* ---- Fused AdamW over flat buffers (optim.cu) -- lib/helpers/optimizer_helper.py:69-129 (the reference's AdamW.step) ----
 * p, g, m, v: n floats each, 16-byte aligned; elements [0, n_decay) get `weight_decay`, the rest 0 (the reference's
 * 'bias' in name -> no decay rule, optimizer_helper.py:9-16, realised by the flat ordering).  step_size =
 * lr * sqrt(1 - beta2^t) / (1 - beta1^t) is computed by the caller: as a host float, or -- step_size_dev != NULL -- read from
 * device memory at run time (a captured CUDA graph then sees the new value every replay).  one_minus_beta* are passed
 * separately so that they are the fp32 roundings of the DOUBLE expressions 1 - beta*, as in the reference. */
int mdb_adamw_step_f32(float* p, const float* g, float* m, float* v, long long n, long long n_decay, float beta1,
                       float one_minus_beta1, float beta2, float one_minus_beta2, float eps, float weight_decay,
                       float step_size, const float* step_size_dev, void* stream);

/* ---- Inference post-process on the device (decode.cu) -- SURVEY.md 8 f3 ----
 * extract: lib/helpers/decode_helper.py:57-110 (extract_dets_from_outputs).  logits (B,Q,C), boxes (B,Q,6) cx cy l r t b,
 * dim3 (B,Q,3), depth (B,Q,2) [depth, log-variance], angle (B,Q,24).  dets (B,topk,37) = label, score, xs2d, ys2d, w, h, depth,
 * heading[24], size3d[3], xs3d, ys3d, sigma; rows ordered by descending score, ties by ascending (query, class).
 * Q*C <= 4096 (else MDB_EUNSUPPORTED), topk <= Q*C.
 * decode: lib/helpers/decode_helper.py:8-54 (decode_detections) with the camera arithmetic of
 * lib/datasets/kitti/kitti_utils.py:150-155,207-208,277-282.  img_size (B,2) [W,H], P2 (B,3,4), cls_mean_size (C,3).
 * out (B,topk,14) = cls, alpha, x0, y0, x1, y1, h, w, l, X, Y, Z, ry, score*sigma for the count[b] leading rows whose score
 * reaches `threshold` (the rest zero-filled); count (B) int32.  Both run on `stream` without synchronising. */
int mdb_extract_dets_f32(const float* logits, const float* boxes, const float* dim3, const float* depth, const float* angle, int B,
                         int Q, int C, int topk, float* dets, void* stream);
int mdb_decode_dets_f32(const float* dets, const float* img_size, const float* P2, const float* cls_mean_size, int B, int topk, int C,
                        float threshold, float* out, int* count, void* stream);

/* ---- Training criterion on the device (criterion.cu) -- SURVEY.md 8 f1 ----
 * lib/models/monodetr/matcher.py:36-104 (HungarianMatcher.forward), monodetr.py:297-532 (SetCriterion), depth_predictor/ddn_loss/
 * {ddn_loss.py:43-127, balancer.py:21-81, focalloss.py:52-125}, lib/helpers/trainer_helper.py:175-186 (prepare_targets).
 * Targets are the data loader's PADDED arrays (B, Gmax, ...) plus the validity mask (B, Gmax) -- no compaction on the host:
 *   labels int32, boxes2d (cx cy w h), boxes3d (cx cy l r t b), depth, size3d (3), heading_bin int32, heading_res.
 * Predictions are passed per decoder layer (layer 0 = the final outputs, then the aux outputs), L <= MDB_CRITERION_MAX_LAYERS:
 *   logits (B,Q,C), boxes (B,Q,6), dim3 (B,Q,3), depth (B,Q,2), angle (B,Q,24), all contiguous fp32.
 * Call order: prepare -> [all-reduce `total` across ranks] -> match -> depth_map -> losses; backward: depth_map (gradient mode) and
 * losses_backward.  Nothing synchronises with the host.  Q / group <= 64, Gmax <= 64, B <= 1024 (else MDB_EUNSUPPORTED). */
#define MDB_CRITERION_MAX_LAYERS 4
#define MDB_CRITERION_NUM_LOSSES 10
#define MDB_LOSS_CE 0           /* loss slots of one layer in `losses` / `grad_losses` (L, MDB_CRITERION_NUM_LOSSES) */
#define MDB_LOSS_CLASS_ERROR 1  /* (logging only, no gradient) */
#define MDB_LOSS_BBOX 2
#define MDB_LOSS_GIOU 3
#define MDB_LOSS_CARDINALITY 4  /* (logging only, no gradient) */
#define MDB_LOSS_DEPTH 5
#define MDB_LOSS_DIM 6
#define MDB_LOSS_ANGLE 7
#define MDB_LOSS_CENTER 8
#define MDB_LOSS_DEPTH_MAP 9    /* layer 0 only */
/* tlist (B,Gmax): indices of the valid targets of each image in their original order, -1 padded; count (B); total (1) = sum */
int mdb_criterion_prepare(const unsigned char* mask, int B, int Gmax, int* tlist, int* count, float* total, void* stream);
/* match (L,B,group,Gmax): query index in [0,Q) assigned to the j-th valid target of the image by the group's assignment problem
 * (cost = w_bbox*L1(box) + w_center*L1(centre) + w_class*focal cost + w_giou*(-GIoU)), -1 where none; tclass (L,B,Q): class of
 * the target a query is matched to, C for "no object". */
int mdb_criterion_match_f32(int L, const float* const* logits, const float* const* boxes, const int* labels, const float* boxes3d,
                            const int* tlist, const int* count, int B, int Q, int C, int group, int Gmax, float w_class, float w_center,
                            float w_bbox, float w_giou, int* match, int* tclass, void* stream);
/* Depth-map loss per pixel.  logits are addressed as logits[b*stride_b + pixel*stride_pix + c*stride_c], c in [0, num_bins]
 * (NHWC: stride_pix = num_bins+1, stride_c = 1; NCHW: stride_pix = 1, stride_c = H*W).  boxes2d are scaled by (scale_x, scale_y)
 * (the reference hard-codes 80, 24).  dlogits == NULL: writes pix_loss (B*H*W), already weighted fg/bg; otherwise writes
 * dlogits (same addressing) = grad_loss[0] * d(mean-balanced loss)/d logits. */
int mdb_criterion_depth_map_f32(const float* logits, long long stride_b, long long stride_pix, long long stride_c, const float* boxes2d,
                                const float* depth, const int* tlist, const int* count, int B, int H, int W, int num_bins, int Gmax,
                                float scale_x, float scale_y, float depth_min, float depth_max, float alpha, float fg_weight,
                                float bg_weight, float* pix_loss, const float* grad_loss, float* dlogits, void* stream);
/* losses (L, MDB_CRITERION_NUM_LOSSES), un-weighted, normalised by num_boxes = max(total * group / world_size, 1) as the reference.
 * pix_loss / npix: output of the depth-map kernel (NULL / 0: slot stays 0).  aux (L): per-layer scalar kept for backward (the
 * gradient-free compensation weight of the dimension loss, monodetr.py:414-416).  Deterministic (fixed-order reductions). */
int mdb_criterion_losses_f32(int L, const float* const* logits, const float* const* boxes, const float* const* dim3,
                             const float* const* depth, const float* const* angle, const int* labels, const float* boxes3d,
                             const float* tdepth, const float* size3d, const int* hbin, const float* hres, const int* tlist,
                             const int* count, const float* total, const int* match, const int* tclass, const float* pix_loss, int npix,
                             int B, int Q, int C, int group, int Gmax, float focal_alpha, float world_size, float* losses, float* aux,
                             void* stream);
/* d(sum_k grad_losses[l][k] * losses[l][k]) / d predictions; every d* buffer is fully written (zeros for unmatched queries). */
int mdb_criterion_losses_backward_f32(int L, const float* const* logits, const float* const* boxes, const float* const* dim3,
                                      const float* const* depth, const float* const* angle, const int* labels, const float* boxes3d,
                                      const float* tdepth, const float* size3d, const int* hbin, const float* hres, const int* tlist,
                                      const int* count, const float* total, const int* match, const int* tclass, int B, int Q, int C,
                                      int group, int Gmax, float focal_alpha, float world_size, const float* grad_losses,
                                      const float* aux, float* const* dlogits, float* const* dboxes, float* const* ddim3, float* const* ddepth,
                                      float* const* dangle, void* stream);

/* ---- Input pipeline on the device (preprocess.cu) -- SURVEY.md 8 f4 ----
 * lib/datasets/kitti/kitti_dataset.py:140-161: [flip] -> PIL Image.transform(AFFINE, BILINEAR) -> float32 / 255 -> (x - mean) / std -> CHW.
 * src: DEVICE array of B device pointers to 8-bit RGB images (3 bytes per pixel), src_wh (B,2) [W,H] int32 and src_pitch (B) bytes
 * per row (device); trans_inv (B,6) fp64 device = PIL's `data` (output pixel centre -> input coordinates, the reference's
 * trans_inv); flip (B) bytes or NULL: sample the left-right mirrored image.  mean3 / std3: HOST floats.
 * out (B,3,out_h,out_w) fp32.  The 8-bit interpolation result and the float conversion are bit-identical to PIL + numpy. */
int mdb_warp_affine_normalize_u8(const unsigned char* const* src, const int* src_wh, const long long* src_pitch, const double* trans_inv,
                                 const unsigned char* flip, int B, int out_w, int out_h, const float* mean3, const float* std3, float* out,
                                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MONODETR_B200_H_ */
