// Heads directly downstream of the fusion output: the class-probability softmax of
// MultimodalEmotionModel.forward (reference models/multimodal_model.py:160-164: F.softmax over 7 emotion / uncertainty logits)
// and the trainer's label-smoothed cross-entropy (training/advanced_trainer.py:53,139: nn.CrossEntropyLoss(label_smoothing=0.1),
// mean reduction).  C is small (7): one thread per sample keeps the row in registers; fp32 arithmetic.
#include "common.cuh"

namespace b200f {

static constexpr int MAXC = 64;

template <typename T>
__global__ void row_softmax_fwd_kernel(const T* __restrict__ x, long long ldx, float* __restrict__ p, long long B, int C) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float v[MAXC], mx = -INFINITY, s = 0.f;
  for (int c = 0; c < C; ++c) { v[c] = to_f32(x[b * ldx + c]); mx = fmaxf(mx, v[c]); }
  for (int c = 0; c < C; ++c) { v[c] = expf(v[c] - mx); s += v[c]; }
  const float inv = 1.f / s;
  for (int c = 0; c < C; ++c) p[b * C + c] = v[c] * inv;
}

// dx = p o (dp - <p, dp>)
template <typename T>
__global__ void row_softmax_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp, T* __restrict__ dx, long long lddx, long long B, int C) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float dot = 0.f;
  for (int c = 0; c < C; ++c) dot += p[b * C + c] * dp[b * C + c];
  for (int c = 0; c < C; ++c) dx[b * lddx + c] = from_f32<T>(p[b * C + c] * (dp[b * C + c] - dot));
}

// loss_b = -(1-eps) log p[b, y_b] - eps/C sum_c log p[b, c];  *loss_sum += sum_b loss_b;  probs saved for backward
template <typename T>
__global__ void ce_ls_fwd_kernel(const T* __restrict__ x, long long ldx, const long long* __restrict__ target, float eps, float* __restrict__ probs,
                                 float* __restrict__ loss_sum, int* __restrict__ bad_target, long long B, int C) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.f;
  if (b < B) {
    float v[MAXC], mx = -INFINITY, s = 0.f, sumx = 0.f;
    for (int c = 0; c < C; ++c) { v[c] = to_f32(x[b * ldx + c]); mx = fmaxf(mx, v[c]); sumx += v[c]; }
    for (int c = 0; c < C; ++c) s += expf(v[c] - mx);
    const float lse = mx + logf(s);
    const long long y = target[b];
    if (y < 0 || y >= C) { atomicExch(bad_target, 1); }
    else loss = (1.f - eps) * (lse - v[y]) + eps * (lse - sumx / C);
    const float inv = 1.f / s;
    for (int c = 0; c < C; ++c) probs[b * C + c] = expf(v[c] - mx) * inv;
  }
  loss = warp_sum(loss);
  if ((threadIdx.x & 31) == 0 && loss != 0.f) atomicAdd(loss_sum, loss);
}

// dx[b, c] = g * (p[b, c] - (1-eps) [c == y_b] - eps/C) / B
template <typename T>
__global__ void ce_ls_bwd_kernel(const float* __restrict__ probs, const long long* __restrict__ target, float eps, const float* __restrict__ gscale,
                                 T* __restrict__ dx, long long lddx, long long B, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const long long b = i / C;
  const int c = int(i - b * C);
  const float g = (gscale ? *gscale : 1.f) / float(B);
  const float t = (c == target[b] ? 1.f - eps : 0.f) + eps / C;
  dx[b * lddx + c] = from_f32<T>(g * (probs[i] - t));
}

#define DISPATCH_DTYPE(dtype, T, ...)                                              \
  if ((dtype) == B200F_F32) { using T = float; __VA_ARGS__ }                       \
  else if ((dtype) == B200F_BF16) { using T = bf16; __VA_ARGS__ }                  \
  else return fail(B200F_ERR_DTYPE, "unknown dtype %d", int(dtype));

}  // namespace b200f

using namespace b200f;

extern "C" {

int b200f_row_softmax_fwd(const void* x, int64_t ldx, float* probs, int64_t B, int32_t C, int32_t dtype, void* stream) {
  if (B == 0) return B200F_OK;
  B200F_REQUIRE(C >= 1 && C <= MAXC, B200F_ERR_SHAPE, "row_softmax: C=%d not in [1,%d]", C, MAXC);
  DISPATCH_DTYPE(dtype, T, {
    row_softmax_fwd_kernel<T><<<(unsigned)((B + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const T*>(x), ldx, probs, B, C);
  })
  return check_launch("row_softmax_fwd");
}

int b200f_row_softmax_bwd(const float* probs, const float* dprobs, void* dx, int64_t lddx, int64_t B, int32_t C, int32_t dtype, void* stream) {
  if (B == 0) return B200F_OK;
  B200F_REQUIRE(C >= 1 && C <= MAXC, B200F_ERR_SHAPE, "row_softmax: C=%d not in [1,%d]", C, MAXC);
  DISPATCH_DTYPE(dtype, T, {
    row_softmax_bwd_kernel<T><<<(unsigned)((B + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(probs, dprobs, static_cast<T*>(dx), lddx, B, C);
  })
  return check_launch("row_softmax_bwd");
}

int b200f_ce_ls_fwd(const void* logits, int64_t ldx, const int64_t* target, float label_smoothing, float* probs, float* loss_sum, int32_t* bad_target,
                    int64_t B, int32_t C, int32_t dtype, void* stream) {
  if (B == 0) return B200F_OK;
  B200F_REQUIRE(C >= 1 && C <= MAXC, B200F_ERR_SHAPE, "cross_entropy: C=%d not in [1,%d]", C, MAXC);
  B200F_REQUIRE(label_smoothing >= 0.f && label_smoothing <= 1.f, B200F_ERR_SHAPE, "cross_entropy: label_smoothing=%f", label_smoothing);
  DISPATCH_DTYPE(dtype, T, {
    ce_ls_fwd_kernel<T><<<(unsigned)((B + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const T*>(logits), ldx, reinterpret_cast<const long long*>(target), label_smoothing, probs, loss_sum, bad_target, B, C);
  })
  return check_launch("ce_ls_fwd");
}

int b200f_ce_ls_bwd(const float* probs, const int64_t* target, float label_smoothing, const float* gscale_dev, void* dlogits, int64_t lddx, int64_t B,
                    int32_t C, int32_t dtype, void* stream) {
  if (B == 0) return B200F_OK;
  DISPATCH_DTYPE(dtype, T, {
    ce_ls_bwd_kernel<T><<<(unsigned)((B * C + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        probs, reinterpret_cast<const long long*>(target), label_smoothing, gscale_dev, static_cast<T*>(dlogits), lddx, B, C);
  })
  return check_launch("ce_ls_bwd");
}

}  // extern "C"
