// InfoNCE (ContrastiveFusion.contrastive_loss, fusion_layers.py:361-375) as two entry points that work on
// a [Bl x Bg] block of the similarity matrix (local rows x gathered rows), so data-parallel ranks need only
// the all-gather of the embeddings (and of the two LSE vectors in backward):
//   lse  : S = inv_tau * x y^T (GEMM into the caller's workspace), row log-sum-exp + diagonal
//   grad : W = coef*g * (exp(S - lse_x) + exp(S^T-side lse_y) - 2*I)  ->  dx (+)= W y   (GEMM)
// Round-1 implementation: the S block lives in the caller-provided workspace (fp32) between the GEMM and the
// row kernels; the contraction itself runs on tcgen05 (bf16) or FFMA (fp32 parity mode) via b200f_gemm.
#include "common.cuh"

namespace b200f {

// fused tcgen05 path (infonce_tc.cu): the similarity block stays in TMEM; this file's GEMM + row-kernel route remains for fp32
// parity mode, embedding widths the fused kernels do not take, and A/B testing (b200f_debug_set(6, 1))
bool infonce_tc_eligible(const void* x, const void* y, int64_t Bl, int64_t Bg, int32_t D, int32_t dtype);
size_t infonce_tc_workspace_bytes(int64_t Bl);
int infonce_lse_tc(const void* x, const void* y, float* lse, float* diag, int64_t Bl, int64_t Bg, int32_t D, int64_t diag_off, float inv_tau,
                   void* workspace, cudaStream_t st);
int infonce_grad_tc(const void* x, const void* y, const float* lse_x, const float* lse_y, float coef, const float* gscale_dev, float* dx, int64_t Bl,
                    int64_t Bg, int32_t D, int64_t diag_off, float inv_tau, cudaStream_t st);
int g_infonce_variant = 0;       // 0 = fused tcgen05 kernels when eligible, 1 = GEMM + row kernels

// warp per row: online (max, sum) over the row of S, float4 loads
__global__ void __launch_bounds__(256) row_lse_kernel(const float* __restrict__ S, long long ld, float* __restrict__ lse, float* __restrict__ diag,
                                                      long long Bl, long long Bg, long long diag_off) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= Bl) return;
  const float* s = S + row * ld;
  float m = -INFINITY, l = 0.f;
  const long long nvec = Bg / 4;
  for (long long v = lane; v < nvec; v += 32) {
    const float4 f = *reinterpret_cast<const float4*>(s + v * 4);
    const float mm = fmaxf(fmaxf(f.x, f.y), fmaxf(f.z, f.w));
    if (mm > m) { l *= expf(m - mm); m = mm; }
    l += expf(f.x - m) + expf(f.y - m) + expf(f.z - m) + expf(f.w - m);
  }
  for (long long j = nvec * 4 + lane; j < Bg; j += 32) {
    const float x = s[j];
    if (x > m) { l *= expf(m - x); m = x; }
    l += expf(x - m);
  }
  const float mw = warp_max(m);
  l = (m == -INFINITY) ? 0.f : l * expf(m - mw);
  l = warp_sum(l);
  if (lane == 0) {
    lse[row] = mw + logf(l);
    if (diag) diag[row] = s[diag_off + row];
  }
}

// W[i,j] = c * (exp(S - lse_x[i]) + exp(S - lse_y[j]) - 2*[j == diag_off + i]), written in T for the second GEMM
template <typename T>
__global__ void infonce_w_kernel(const float* __restrict__ S, long long ld, const float* __restrict__ lse_x, const float* __restrict__ lse_y,
                                 const float* __restrict__ gscale, float coef, T* __restrict__ W, long long ldw, long long Bl, long long Bg,
                                 long long diag_off) {
  const float c = coef * (gscale ? *gscale : 1.f);
  const long long total = Bl * Bg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Bg, j = i - r * Bg;
    const float s = S[r * ld + j];
    float w = expf(s - lse_x[r]) + expf(s - lse_y[j]);
    if (j == diag_off + r) w -= 2.f;
    W[r * ldw + j] = from_f32<T>(w * c);
  }
}

}  // namespace b200f

using namespace b200f;

extern "C" {

// workspace: fp32 S block [Bl, Bg8] (+ a bf16/fp32 W block of the same shape for grad); Bg8 = Bg rounded up to 8
size_t b200f_infonce_workspace_bytes(int64_t Bl, int64_t Bg, int32_t dtype, int32_t for_grad) {
  const size_t bg8 = (size_t)((Bg + 7) / 8 * 8);
  size_t n = (size_t)Bl * bg8 * 4;
  if (for_grad) n += (size_t)Bl * bg8 * (dtype == B200F_BF16 ? 2 : 4);
  return n + 256;
}

static int sim_block(const void* x, const void* y, float* S, long long bg8, int64_t Bl, int64_t Bg, int32_t D, float inv_tau, int32_t dtype, void* stream) {
  b200f_gemm_args g = {};
  g.M = Bl; g.N = Bg; g.K = D;
  g.a_layout = 0; g.b_layout = 0;
  g.A = x; g.lda = D; g.B = y; g.ldb = D;
  g.C = S; g.ldc = bg8;
  g.alpha = inv_tau;
  g.flags = (dtype == B200F_BF16) ? B200F_EPI_OUT_F32 : 0;
  g.dtype = dtype; g.split_k = 1;
  return b200f_gemm(&g, stream);
}

int b200f_infonce_lse(const void* x, const void* y, float* lse, float* diag, int64_t Bl, int64_t Bg, int32_t D, int64_t diag_off, float inv_tau,
                      int32_t dtype, void* workspace, size_t workspace_bytes, void* stream) {
  if (Bl == 0) return B200F_OK;
  B200F_REQUIRE(Bg > 0 && D > 0, B200F_ERR_SHAPE, "infonce: empty y");
  B200F_REQUIRE(!diag || (diag_off >= 0 && diag_off + Bl <= Bg), B200F_ERR_SHAPE, "infonce: diagonal offset %lld out of range", (long long)diag_off);
  B200F_REQUIRE(workspace && workspace_bytes >= b200f_infonce_workspace_bytes(Bl, Bg, dtype, 0) && aligned16(workspace), B200F_ERR_SHAPE,
                "infonce: workspace too small (%zu bytes)", workspace_bytes);
  if (g_infonce_variant == 0 && infonce_tc_eligible(x, y, Bl, Bg, D, dtype) && workspace_bytes >= infonce_tc_workspace_bytes(Bl))
    return infonce_lse_tc(x, y, lse, diag, Bl, Bg, D, diag_off, inv_tau, workspace, static_cast<cudaStream_t>(stream));
  const long long bg8 = (Bg + 7) / 8 * 8;
  float* S = static_cast<float*>(workspace);
  int rc = sim_block(x, y, S, bg8, Bl, Bg, D, inv_tau, dtype, stream);
  if (rc) return rc;
  row_lse_kernel<<<(unsigned)((Bl + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(S, bg8, lse, diag, Bl, Bg, diag_off);
  return check_launch("row_lse");
}

int b200f_infonce_grad(const void* x, const void* y, const float* lse_x, const float* lse_y, float coef, const float* gscale_dev, float* dx,
                       int32_t accumulate, int64_t Bl, int64_t Bg, int32_t D, int64_t diag_off, float inv_tau, int32_t dtype, void* workspace,
                       size_t workspace_bytes, void* stream) {
  if (Bl == 0) return B200F_OK;
  B200F_REQUIRE(Bg > 0 && D > 0, B200F_ERR_SHAPE, "infonce: empty y");
  B200F_REQUIRE(workspace && workspace_bytes >= b200f_infonce_workspace_bytes(Bl, Bg, dtype, 1) && aligned16(workspace), B200F_ERR_SHAPE,
                "infonce: workspace too small (%zu bytes)", workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g_infonce_variant == 0 && infonce_tc_eligible(x, y, Bl, Bg, D, dtype) && aligned16(dx)) {
    if (!accumulate) B200F_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)Bl * D * 4, st));
    return infonce_grad_tc(x, y, lse_x, lse_y, coef, gscale_dev, dx, Bl, Bg, D, diag_off, inv_tau, st);
  }
  const long long bg8 = (Bg + 7) / 8 * 8;
  float* S = static_cast<float*>(workspace);
  void* W = static_cast<char*>(workspace) + (size_t)Bl * bg8 * 4;
  int rc = sim_block(x, y, S, bg8, Bl, Bg, D, inv_tau, dtype, stream);
  if (rc) return rc;
  long long blocks = (Bl * Bg + 255) / 256;
  if (blocks > 32LL * num_sms()) blocks = 32LL * num_sms();
  if (dtype == B200F_BF16) {
    if (bg8 != Bg) B200F_CHECK_CUDA(cudaMemsetAsync(W, 0, (size_t)Bl * bg8 * 2, st));
    infonce_w_kernel<bf16><<<(unsigned)blocks, 256, 0, st>>>(S, bg8, lse_x, lse_y, gscale_dev, coef, static_cast<bf16*>(W), bg8, Bl, Bg, diag_off);
  } else {
    infonce_w_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(S, bg8, lse_x, lse_y, gscale_dev, coef, static_cast<float*>(W), bg8, Bl, Bg, diag_off);
  }
  rc = check_launch("infonce_w");
  if (rc) return rc;
  if (!accumulate) B200F_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)Bl * D * 4, st));
  b200f_gemm_args g = {};
  g.M = Bl; g.N = D; g.K = Bg;
  g.a_layout = 0; g.b_layout = 1;           // dx[i,d] = sum_j W[i,j] y[j,d]
  g.A = W; g.lda = bg8; g.B = y; g.ldb = D;
  g.C = dx; g.ldc = D;
  g.alpha = 1.f;
  g.flags = B200F_EPI_ACCUM;
  g.dtype = dtype; g.split_k = 1;
  return b200f_gemm(&g, stream);
}

}  // extern "C"
