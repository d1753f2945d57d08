/* b200_fusion.h -- C ABI of libb200fusion.so: hand-written sm_100a kernels for the fusion hot
 * path of nl1xx/simple-multimodal (models/fusion_layers.py + models/encoders.py:280-321).
 *
 * The reference has no FFI, plugin registry or operator API: its boundary is the
 * Python nn.Module contract consumed by models/multimodal_model.py:29-46,110-144.  This header is
 * the build-defined C boundary underneath that contract; each entry point names the reference
 * arithmetic it replaces.  The Python host (simple-multimodal_b200/) binds it with ctypes.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is DEVICE memory owned by the caller (PyTorch's
 *     caching allocator in practice) unless stated otherwise; the library allocates nothing
 *     persistent and keeps no pointer after return.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no internal sync.
 *   - return 0 (B200F_OK) or a b200f_status; never throws, never exits; b200f_last_error() gives
 *     a thread-local message for the last non-zero return.
 *   - dtype: B200F_F32 = fp32 storage and CUDA-core FFMA arithmetic (parity mode, rtol 1e-5);
 *            B200F_BF16 = bf16 storage, tcgen05 tensor-core contractions with fp32 accumulation,
 *            fp32 statistics (LayerNorm mean/rstd, softmax LSE, losses) and fp32 parameter grads.
 *   - matrices are row-major with an explicit leading dimension (elements).
 */
#ifndef B200_FUSION_H
#define B200_FUSION_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B200F_OK = 0,
  B200F_ERR_SHAPE = 1,
  B200F_ERR_DTYPE = 2,
  B200F_ERR_ALIGN = 3,
  B200F_ERR_CUDA = 4,
  B200F_ERR_UNSUPPORTED = 5
} b200f_status;

typedef enum { B200F_F32 = 0, B200F_BF16 = 1 } b200f_dtype;

/* GEMM epilogue flags */
enum {
  B200F_EPI_RELU = 1,       /* C = max(C, 0) after bias / residual                                  */
  B200F_EPI_OUT_F32 = 2,    /* C is fp32 regardless of dtype                                         */
  B200F_EPI_ACCUM = 4,      /* C (fp32) += result (atomic; used for parameter gradients, split-K)    */
  B200F_EPI_GELU_RSVD = 8   /* reserved                                                              */
};

int b200f_version(void);
const char* b200f_last_error(void);
/* kernels launched by this library so far in this process (bench.py's gpu_launches evidence). */
unsigned long long b200f_launch_count(void);
/* 0 if `device` is an sm_100 part that can run the bf16/tcgen05 kernels. */
int b200f_device_supported(int device);

/* ---------------------------------------------------------------------------------------------
 * GEMM with fused epilogue:  C[m,n] = epi( alpha * sum_k A(m,k) * B(n,k) )
 *   a_layout 0: A stored [M,K] (K contiguous)   1: A stored [K,M] (M contiguous)
 *   b_layout 0: B stored [N,K] (K contiguous)   1: B stored [K,N] (N contiguous)
 *   epi(x) = [relu]( x + bias[n] + residual[m,n] ) * (relu_mask[m,n] > 0 ? 1 : 0)      (relu_mask or its 1-bit form sign_bits)
 * Replaces every nn.Linear on the path (fusion_layers.py:21-28,55-57,124-128,195-200,238,
 * 304-327,395-412,471-476 and the packed in/out projections of nn.MultiheadAttention,
 * torch/nn/functional.py:5833-5860,6653) in forward (layouts 0/0), input-gradient (0/1) and
 * weight-gradient (1/1, B200F_EPI_ACCUM) form.  bf16: tcgen05.mma, TMA-fed, accumulators in TMEM.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t M, N, K;
  int32_t a_layout, b_layout;
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* C; int64_t ldc;
  const float* bias;                    /* [N] fp32 or NULL                                      */
  const void* residual; int64_t ldr;    /* [M,N] in `dtype`, or NULL                             */
  const void* relu_mask; int64_t ldm;   /* [M,N] in `dtype`: zero the output where mask <= 0     */
  float alpha;
  int32_t flags;                        /* B200F_EPI_*                                           */
  int32_t dtype;                        /* b200f_dtype of A, B, residual, relu_mask (and C)      */
  int32_t split_k;                      /* >1 only with B200F_EPI_ACCUM; 0/1 = none              */
  /* inverted dropout on the output after bias/residual/ReLU (nn.Dropout behind a Linear+ReLU, fusion_layers.py:197-199):
   * element (m, n) kept iff the counter-based hash of (seed, m, n) passes, kept values scaled by 1/(1-p); 0 = off */
  float dropout_p; uint32_t drop_seed_lo, drop_seed_hi;
  float* colsum;                        /* optional [N] fp32 accumulator: += column sums of the
                                           stored C (the bias gradient of the Linear whose output
                                           gradient C is, e.g. ffn.0.bias from the FFN2 input-gradient
                                           GEMM); summed in the epilogue of the tcgen05 kernels      */
  /* The ReLU' mask as ONE BIT per element instead of the stored activation (round 2): the Linear+ReLU(+Dropout) forward GEMM
   * writes `sign_bits_out` (bit set <=> the stored C element is > 0, i.e. passed the ReLU and was kept by the dropout), the
   * input-gradient GEMM of the next Linear reads it as `sign_bits` in place of `relu_mask` -- 1/16 of the bytes (the FFN2
   * input-gradient GEMM of fusion_layers.py:195-200 was half HBM-bound on re-reading the [M, 2048] hidden layer).  Layout:
   * [M, ldsb] 32-bit words, word c of a row covers columns 32c .. 32c+31 in the library's own bit order (opaque: only valid
   * between these two fields).  bf16 tcgen05 path only, N % 64 == 0, ldsb even and >= N / 32, 8-byte aligned; NULL = off.   */
  uint32_t* sign_bits_out; const uint32_t* sign_bits; int64_t ldsb;
} b200f_gemm_args;
int b200f_gemm(const b200f_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-head attention core (torch/nn/functional.py:6630-6647, need_weights path with the
 * weights discarded: fusion_layers.py:161-163,204):  O = softmax(scale * Q K^T) V  per (b, head).
 * Q/K/V/O are token-major: element (b, l, h, d) at  base + (b*L + l)*ld + h*D + d, so they can
 * alias column slices of packed projection outputs.  LSE[b,h,l] (fp32, natural log of the
 * scaled-score row sum) is saved for backward.  D (head dim) must be 64 for bf16.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, H, Lq, Lk, D;
  const void* Q; int64_t ldq;
  const void* K; int64_t ldk;
  const void* V; int64_t ldv;
  void* O; int64_t ldo;
  float* LSE;                           /* [B,H,Lq]                                              */
  float scale;
  int32_t dtype;
  /* backward only */
  const void* dO; int64_t lddo;
  void* dQ; int64_t lddq;
  void* dK; int64_t lddk;
  void* dV; int64_t lddv;
  float* delta;                         /* [B,H,Lq] workspace: rowsum(dO * O)                    */
  /* optional fp32 [H*D] accumulators (nullable): += column sums over (b, l) of dQ / dK / dV, i.e.
   * the bias gradients of the projections that produced Q / K / V (in_proj_bias of
   * nn.MultiheadAttention, torch nn/functional.py:5740-5760) -- fused into the backward epilogue */
  float* dbq; float* dbk; float* dbv;
  /* dropout on the attention probabilities (nn.MultiheadAttention(dropout=p) in training mode, torch nn/functional.py
   * 6647-6650): kept entries are scaled by 1/(1-p); the mask is regenerated from (drop_seed_lo, drop_seed_hi) by forward
   * and backward (counter-based hash of (b, head, query, key), see csrc/common.cuh) -- pass the same values to both.
   * dropout_p = 0 disables it.  LSE is the log-sum-exp of the UNdropped scores. */
  float dropout_p; uint32_t drop_seed_lo, drop_seed_hi;
  /* forward only, optional (nullable; bf16 / head dim 64 kernels): fp32 [B, parts, H*D] PARTIAL column sums of the O that is stored,
   * parts = b200f_attn_pool_parts(args): partial p of (b, head) covers a fixed group of query rows and is written (not added) by
   * exactly one warp, so the result is bit-reproducible; b200f_pool_finish sums the parts in order and scales -- the `.mean(dim=1)`
   * of the attended features (fusion_layers.py:166-168) straight out of the attention epilogue, no separate pass re-reads O. */
  float* pool_sum;
  /* backward only, optional (nullable): scratch of b200f_attn_bwd_ws_bytes(args) bytes.  With it, the tcgen05 backward of blocks
   * with >= 128 queries and keys computes the score gradient dS ONCE: the dK/dV kernel also stores its bf16 dS^T tiles here
   * (bulk tensor stores straight from the shared-memory operand tiles) and dQ = dS K becomes a memory-bound batched GEMM over
   * them, instead of a second kernel that recomputes S, P, dP and dS (exp, dropout hash and three of seven MMAs of the backward).
   * The route is off by default (b200f_attn_bwd_ws_bytes returns 0): on the power-capped B200 it measured neutral, see csrc/attn_tc.cu. */
  void* bwd_ws; int64_t bwd_ws_bytes;
} b200f_attn_args;
int64_t b200f_attn_bwd_ws_bytes(const b200f_attn_args* args); /* 0: this shape / dtype does not use the scratch */
int32_t b200f_attn_pool_parts(const b200f_attn_args* args);   /* 0: the kernel family serving this shape / dtype has no pooled output */
/* out[b, c] = scale * sum_p partial[b, p, c]  (p ascending; out bf16 or fp32 by `dtype`, leading dimension ldo) */
int b200f_pool_finish(const float* partial, void* out, int64_t ldo, int64_t B, int32_t parts, int32_t W, float scale, int32_t dtype, void* stream);
int b200f_attn_fwd(const b200f_attn_args* args, void* stream);
int b200f_attn_bwd(const b200f_attn_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Row-wise kernels over [rows, H] activations (128-bit vectorised, warp-shuffle reductions).
 * ------------------------------------------------------------------------------------------- */
/* y = LayerNorm(x) * gamma + beta (+ post1 + post2)  (fusion_layers.py:205,209; the optional
 * post-adds implement the 3-way residual of :156-158 in the same pass).  mean/rstd saved. */
int b200f_layernorm_fwd(const void* x, const float* gamma, const float* beta, const void* post1,
                        const void* post2, void* y, float* mean, float* rstd, int64_t rows, int32_t H,
                        float eps, int32_t dtype, void* stream);
/* dx = LN'(dy) (+ dres);  dgamma += sum dy*xhat;  dbeta += sum dy;  optional dxsum[H] += column sum of dx
 * (the bias gradient of the Linear whose output fed this LayerNorm, fused here)   (fp32 atomics). */
int b200f_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                        const float* gamma, const void* dres, void* dx, float* dgamma, float* dbeta,
                        float* dxsum, int64_t rows, int32_t H, int32_t dtype, void* stream);
/* out[n] += sum_m x[m,n]   (bias gradients). */
int b200f_colsum_accum(const void* x, int64_t ldx, float* out, int64_t M, int64_t N, int32_t dtype, void* stream);
/* y = a + b (+ c), elementwise over n elements (c may be NULL). */
int b200f_add(const void* a, const void* b, const void* c, void* y, int64_t n, int32_t dtype, void* stream);
/* y = x * (ref > 0)   (ReLU backward given the forward output). */
int b200f_relu_bwd(const void* dy, const void* ref, void* dx, int64_t n, int32_t dtype, void* stream);
/* fp32 -> bf16 cast with optional scale (weights -> tensor-core operands). */
int b200f_cast_f32_to_bf16(const float* src, void* dst, int64_t n, float scale, void* stream);
int b200f_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream);
/* mean over L: x[B,L,H] -> y[B, H] written with row stride ldy (fusion_layers.py:166-168);
 * backward: dx[b,l,:] = dy[b,:] / L. */
int b200f_meanpool_fwd(const void* x, void* y, int64_t ldy, int32_t B, int32_t L, int32_t H, int32_t dtype, void* stream);
/* Staging pass of the MulT engine, one read of x [B,L,H]: copy[b,l,:] = keep * x[b,l,:] and (mean != NULL) mean[b,:] = keep * mean_l x[b,l,:],
 * keep = mask[b*3 + col] (mask NULL: 1) -- ModalityDropout's multiply (models/encoders.py:317-319) and the mean-pool that feeds the
 * 2-D heads of HierarchicalFusion fused into the copy into the static buffers the captured chunk graphs read. */
int b200f_stage_pool(const void* x, void* copy, void* mean, int64_t ldmean, const float* mask, int32_t col, int32_t B, int32_t L, int32_t H,
                     int32_t dtype, void* stream);
/* y[r,:] = a[r,:] + b[r,:] + c[r,:] + s * v[r / L, :] over `rows` rows of H (v: one row per sample with leading dimension ldv,
 * broadcast over the sample's L tokens): the residual paths into a MulT input plus the gradient of its mean-pooled copy. */
int b200f_add_rowbcast(const void* a, const void* b, const void* c, const void* v, int64_t ldv, float s, void* y, int64_t rows, int32_t L, int32_t H,
                       int32_t dtype, void* stream);
int b200f_meanpool_bwd(const void* dy, int64_t lddy, void* dx, int32_t B, int32_t L, int32_t H, int32_t dtype, void* stream);
/* weighted pooling over L: y[b,:] = sum_l w[b,l] * x[b,l,:], w fp32 [B,L].  With w = mask / max(sum_l mask, 1e-9)
 * it is the attention-mask mean pooling of the text encoder (models/encoders.py:89-93) applied to per-token projected features;
 * backward: dx[b,l,:] = w[b,l] * dy[b,:]. */
int b200f_weighted_pool_fwd(const void* x, const float* w, void* y, int64_t ldy, int32_t B, int32_t L, int32_t H, int32_t dtype, void* stream);
int b200f_weighted_pool_bwd(const void* dy, int64_t lddy, const float* w, void* dx, int32_t B, int32_t L, int32_t H, int32_t dtype, void* stream);
/* cat = [t*m0 | a*m1 | v*m2]  rows of 3H (fusion_layers.py:38,350,437 + encoders.py:317-319);
 * mask is [B,3] fp32 keep-mask or NULL.  Backward splits and re-applies the mask, accumulating
 * into dt/da/dv when `accumulate`. */
int b200f_concat3_fwd(const void* t, const void* a, const void* v, const float* mask, void* cat,
                      int64_t B, int32_t H, int32_t dtype, void* stream);
int b200f_concat3_bwd(const void* dcat, const float* mask, void* dt, void* da, void* dv, int32_t accumulate,
                      int64_t B, int32_t H, int32_t dtype, void* stream);
/* x[b, l, :] *= mask[b, col]   in place (modality dropout on [B,L,H] sequences). */
int b200f_rowmask_apply(void* x, const float* mask, int32_t col, int64_t B, int64_t L, int32_t H, int32_t dtype, void* stream);
/* dst[b, l, :] = src[b, l, :] * mask[b, col]  out of place; mask NULL = plain copy (the staging pass of the chunk-graph
 * MulT engine: ModalityDropout's multiply, models/encoders.py:317-319, rides on the copy into the static input buffers). */
int b200f_rowmask_copy(const void* src, void* dst, const float* mask, int32_t col, int64_t B, int64_t L, int32_t H, int32_t dtype, void* stream);
/* z = y / max(||y||_2, eps) per row (F.normalize, fusion_layers.py:338-340); norm saved. */
int b200f_l2norm_fwd(const void* y, void* z, float* norm, int64_t rows, int32_t D, float eps, int32_t dtype, void* stream);
int b200f_l2norm_bwd(const void* dz, const void* z, const float* norm, void* dy, int64_t rows, int32_t D, float eps,
                     int32_t dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * InfoNCE (ContrastiveFusion.contrastive_loss, fusion_layers.py:361-375) without materialising
 * the similarity matrix.  x: [Bl, D] local rows, y: [Bg, D] gathered rows (Bg = Bl when single
 * GPU), s_ij = inv_tau * <x_i, y_j>.
 *   lse:   lse[i]  = log sum_j exp(s_ij);  diag[i] = s_{i, diag_off + i}
 *   grad:  dx[i,:] (+)= coef * g * sum_j ( exp(s_ij - lse_x[i]) + exp(s_ij - lse_y[j]) - 2*[j == diag_off+i] ) * y[j,:]
 * Called twice per modality pair with the roles swapped (rows, then columns of S): the only
 * collective the path needs is the all-gather that produced y (and lse_y).  dx is fp32.  The caller
 * provides the workspace that holds the fp32 similarity block between the GEMM and the row kernels
 * (b200f_infonce_workspace_bytes).
 * ------------------------------------------------------------------------------------------- */
size_t b200f_infonce_workspace_bytes(int64_t Bl, int64_t Bg, int32_t dtype, int32_t for_grad);
int b200f_infonce_lse(const void* x, const void* y, float* lse, float* diag, int64_t Bl, int64_t Bg, int32_t D,
                      int64_t diag_off, float inv_tau, int32_t dtype, void* workspace, size_t workspace_bytes,
                      void* stream);
/* coef is a host scalar, gscale_dev an optional device scalar (the upstream loss gradient) multiplied in. */
int b200f_infonce_grad(const void* x, const void* y, const float* lse_x, const float* lse_y, float coef,
                       const float* gscale_dev, float* dx, int32_t accumulate, int64_t Bl, int64_t Bg, int32_t D,
                       int64_t diag_off, float inv_tau, int32_t dtype, void* workspace, size_t workspace_bytes,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * Small per-sample heads (one warp per sample).
 * ------------------------------------------------------------------------------------------- */
/* GATConv core on dense 3-node graphs (fusion_layers.py:267-282 + PyG GATConv semantics,
 * SURVEY 8c): xp [B,3,heads*C] (the `lin` projection), att_src/att_dst [heads*C], bias [C]
 *   out[b,i,:] = relu( mean_h sum_j softmax_j(leaky_relu(a_src[j,h]+a_dst[i,h])) xp[b,j,h,:] + bias )
 * alpha [B,3(i),heads,3(j)] fp32 (the softmax output) is saved for backward.  dropout_p > 0: GATConv(dropout=p) in training
 * mode drops the attention coefficients before aggregation (kept / (1-p)); the mask is regenerated from the seed pair. */
int b200f_gat_fwd(const void* xp, const float* att_src, const float* att_dst, const float* bias, void* out,
                  float* alpha, int64_t B, int32_t heads, int32_t C, float slope, float dropout_p, uint32_t seed_lo,
                  uint32_t seed_hi, int32_t dtype, void* stream);
int b200f_gat_bwd(const void* dout, const void* out, const void* xp, const float* alpha, const float* att_src,
                  const float* att_dst, void* dxp, float* datt_src, float* datt_dst, float* dbias, int64_t B,
                  int32_t heads, int32_t C, float slope, float dropout_p, uint32_t seed_lo, uint32_t seed_hi, int32_t dtype,
                  void* stream);
/* Attention over the 3 modality tokens (AdaptiveFusion, fusion_layers.py:432-434): qkv [B,3,3H]
 * packed projections -> ctx [B,3,H], probs [B,heads,3,3] fp32 (saved), head-averaged weights
 * avgw [B,3,3] fp32 (returned to the caller, torch/nn/functional.py:6657-6659).  dropout_p > 0: the weights are dropped
 * (kept / (1-p)) before P V and before the head average, as nn.MultiheadAttention(dropout=p) does in training mode;
 * probs stays the softmax output and the mask is regenerated from the seed pair in backward. */
int b200f_tok3_attn_fwd(const void* qkv, void* ctx, float* probs, float* avgw, int64_t B, int32_t heads, int32_t H,
                        float scale, float dropout_p, uint32_t seed_lo, uint32_t seed_hi, int32_t dtype, void* stream);
int b200f_tok3_attn_bwd(const void* dctx, const float* davgw, const void* qkv, const float* probs, void* dqkv,
                        int64_t B, int32_t heads, int32_t H, float scale, float dropout_p, uint32_t seed_lo,
                        uint32_t seed_hi, int32_t dtype, void* stream);
/* Gated mix (fusion_layers.py:437-443): gate = softmax(logits[B,3]); mixed = sum_m att[b,m,:]*gate[b,m]. */
int b200f_gate_mix_fwd(const void* att, const void* logits, float* gate, void* mixed, int64_t B, int32_t H,
                       int32_t dtype, void* stream);
int b200f_gate_mix_bwd(const void* dmixed, const float* dgate_ext, const void* att, const float* gate, void* datt,
                       void* dlogits, int64_t B, int32_t H, int32_t dtype, void* stream);
/* LateFusion combine (fusion_layers.py:75-82): fused = sum_m softmax(w)[m] * logits_m. */
int b200f_late_combine_fwd(const void* lt, const void* la, const void* lv, const float* w3, float* wsoft, void* fused,
                           int64_t B, int32_t E, int32_t dtype, void* stream);
int b200f_late_combine_bwd(const void* dfused, const void* lt, const void* la, const void* lv, const float* wsoft,
                           const float* dwsoft_ext, void* dlt, void* dla, void* dlv, float* dw3, int64_t B, int32_t E,
                           int32_t dtype, void* stream);
/* Modality-dropout keep mask (encoders.py:303-314): mask[B,3] fp32 from a counter-based RNG,
 * keep iff u > rate, with the keep-one repair, no host sync. */
int b200f_modality_mask(float* mask, int64_t B, float rate, uint64_t seed, uint64_t offset, void* stream);
/* Inverted dropout with a counter-based RNG (nn.Dropout on the path): y = x * keep / (1-p);
 * the same (seed, offset) regenerates the mask in backward. */
int b200f_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, uint64_t offset, int32_t dtype, void* stream);
/* In-place inverted dropout of an [M, N] activation (row stride ldx) with the mask b200f_gemm's dropout epilogue generates:
 * element (m, n) kept iff the counter-based hash of (seed_lo, seed_hi, m, n) passes (csrc/common.cuh). */
/* Dropout epoch: a device-side word every mask-generating kernel XORs into its seed.  add != 0: epoch += value, else
 * epoch = value (stream-ordered).  A training step captured in a CUDA graph starts with b200f_dropout_epoch(1, 1, stream) so
 * that each replay draws new masks although the seeds (kernel arguments) are frozen at capture; eager code never needs it. */
int b200f_dropout_epoch(uint32_t value, int32_t add, void* stream);
int b200f_dropout_rowcol(void* x, int64_t ldx, int64_t M, int64_t N, float p, uint32_t seed_lo, uint32_t seed_hi, int32_t dtype,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * Heads directly downstream of the fusion output.
 * row_softmax: probs[b,:] = softmax(x[b,:C]) (F.softmax of the emotion / uncertainty logits, reference
 * models/multimodal_model.py:160-164); ce_ls: nn.CrossEntropyLoss(label_smoothing) with mean reduction
 * (training/advanced_trainer.py:53,139): *loss_sum += sum_b loss_b (caller divides by B), probs [B,C] fp32
 * saved for backward; *bad_target is set to 1 if a target is outside [0, C).  C <= 64.
 * ------------------------------------------------------------------------------------------- */
int b200f_row_softmax_fwd(const void* x, int64_t ldx, float* probs, int64_t B, int32_t C, int32_t dtype, void* stream);
int b200f_row_softmax_bwd(const float* probs, const float* dprobs, void* dx, int64_t lddx, int64_t B, int32_t C, int32_t dtype,
                          void* stream);
int b200f_ce_ls_fwd(const void* logits, int64_t ldx, const int64_t* target, float label_smoothing, float* probs, float* loss_sum,
                    int32_t* bad_target, int64_t B, int32_t C, int32_t dtype, void* stream);
int b200f_ce_ls_bwd(const float* probs, const int64_t* target, float label_smoothing, const float* gscale_dev, void* dlogits,
                    int64_t lddx, int64_t B, int32_t C, int32_t dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Trainer-side optimizer step over the fusion parameters: clip_grad_norm_ + AdamW
 * (reference training/advanced_trainer.py:91-94,174-180) as multi-tensor kernels.  `table` is a device buffer laid out as
 * params[T], grads[T], exp_avg[T], exp_avg_sq[T] (8-byte fp32 pointers), numel[T] (int64), chunks[n_chunks] ({int32 tensor,
 * int32 first element}, 4096 elements per chunk).  b200f_grad_clip_coef leaves {sum of squares, total norm, clip coefficient}
 * in scratch3 (device, 3 floats; nothing is read back to the host); b200f_adamw_step scales the gradients by *clip_coef
 * (NULL = 1) and applies torch.optim.AdamW's update with bias correction for `step` (1-based); the hyper-parameters are doubles, like the Python
 * scalars torch folds on the host before rounding to fp32.
 * ------------------------------------------------------------------------------------------- */
int b200f_grad_clip_coef(const void* table, int32_t n_tensors, int32_t n_chunks, float max_norm, float* scratch3, void* stream);
int b200f_adamw_step(const void* table, int32_t n_tensors, int32_t n_chunks, const float* clip_coef, double lr, double beta1,
                     double beta2, double eps, double weight_decay, int64_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_FUSION_H */
