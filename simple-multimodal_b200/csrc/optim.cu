// Trainer-side optimizer step over the fusion parameters: gradient-norm clipping + AdamW as two multi-tensor
// kernels over a chunk table (reference training/advanced_trainer.py:91-94 AdamW(param groups, weight_decay),
// :174-180 clip_grad_norm_(parameters, gradient_clip_norm) then optimizer.step()).  The ~130 parameter tensors of the hierarchical
// head are otherwise ~130 x (clip, several foreach ops) launches; here the clip coefficient stays on the device (no host sync).
// HBM-bound: read p, g, m, v and write p, m, v once -- 28 bytes per parameter element.
#include "common.cuh"

namespace b200f {

struct ChunkRef { int tensor; int start; };        // chunk of OPT_CHUNK elements of tensor `tensor` beginning at element `start`
static constexpr int OPT_CHUNK = 4096;             // = 256 threads x 4 vectors x 4 elements (the kernels rely on it)

// sumsq += sum over all gradients of g^2
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* const* __restrict__ grads, const long long* __restrict__ numel,
                                                         const ChunkRef* __restrict__ chunks, float* __restrict__ sumsq) {
  const ChunkRef c = chunks[blockIdx.x];
  const float* g = grads[c.tensor];
  const long long n = numel[c.tensor];
  const int len = int(n - c.start < OPT_CHUNK ? n - c.start : OPT_CHUNK);
  const float* gc = g + c.start;
  float acc = 0.f;
  int done = 0;
  if ((reinterpret_cast<uintptr_t>(gc) & 15) == 0) {               // 128-bit loads over the aligned body of the chunk
    const float4* g4 = reinterpret_cast<const float4*>(gc);
    const int nv = len >> 2;
    float4 x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = threadIdx.x + u * 256;
      x[u] = i < nv ? g4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += x[u].x * x[u].x + x[u].y * x[u].y + x[u].z * x[u].z + x[u].w * x[u].w;
    done = len & ~3;
  }
  for (int i = done + threadIdx.x; i < len; i += 256) { const float x = gc[i]; acc += x * x; }
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(sumsq, s);
  }
}

// total_norm = sqrt(sumsq); coef = min(1, max_norm / (total_norm + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ norm_out, float* __restrict__ coef_out) {
  const float nrm = sqrtf(*sumsq);
  *norm_out = nrm;
  const float c = max_norm / (nrm + 1e-6f);
  *coef_out = c < 1.f ? c : 1.f;
}

// torch.optim.AdamW (amsgrad = False, maximize = False), same operation order as its single-tensor path
__global__ void __launch_bounds__(256) adamw_kernel(float* const* __restrict__ params, const float* const* __restrict__ grads, float* const* __restrict__ exp_avg,
                                                    float* const* __restrict__ exp_avg_sq, const long long* __restrict__ numel,
                                                    const ChunkRef* __restrict__ chunks, const float* __restrict__ clip_coef, float decay, float omb1,
                                                    float beta2, float omb2, float eps, float step_size, float sqrt_bias_c2) {
  const ChunkRef c = chunks[blockIdx.x];
  float* p = params[c.tensor];
  const float* g = grads[c.tensor];
  float* m = exp_avg[c.tensor];
  float* v = exp_avg_sq[c.tensor];
  const long long n = numel[c.tensor];
  const float coef = clip_coef ? *clip_coef : 1.f;
  auto update = [&](float gi, float& pi, float& mi, float& vi) {
    gi *= coef;
    pi *= decay;
    mi = mi + (gi - mi) * omb1;                                   // exp_avg.lerp_(grad, 1 - beta1)
    vi = vi * beta2 + omb2 * gi * gi;                             // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vi) / sqrt_bias_c2 + eps;
    pi -= step_size * (mi / denom);
  };
  const int len = int(n - c.start < OPT_CHUNK ? n - c.start : OPT_CHUNK);
  p += c.start; g += c.start; m += c.start; v += c.start;
  int done = 0;
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0) {
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    // 128-bit loads/stores: 28 B of HBM traffic per element, nothing else.  A full chunk is 4 vectors per thread: all 16 loads are
    // issued before the first update so that every thread keeps 256 B in flight
    const int nv = len >> 2;
    float4 gg[4], pp[4], mm[4], vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = threadIdx.x + u * 256;
      if (i < nv) { gg[u] = g4[i]; pp[u] = p4[i]; mm[u] = m4[i]; vv[u] = v4[i]; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = threadIdx.x + u * 256;
      if (i < nv) {
        update(gg[u].x, pp[u].x, mm[u].x, vv[u].x); update(gg[u].y, pp[u].y, mm[u].y, vv[u].y);
        update(gg[u].z, pp[u].z, mm[u].z, vv[u].z); update(gg[u].w, pp[u].w, mm[u].w, vv[u].w);
        p4[i] = pp[u]; m4[i] = mm[u]; v4[i] = vv[u];
      }
    }
    done = len & ~3;
  }
  for (int i = done + threadIdx.x; i < len; i += 256) {
    float pi = p[i], mi = m[i], vi = v[i];
    update(g[i], pi, mi, vi);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

}  // namespace b200f

using namespace b200f;

extern "C" {

// `table` (device): [params[T] | grads[T] | exp_avg[T] | exp_avg_sq[T]] as 8-byte pointers, then numel[T] (int64), then
// chunks[n_chunks] (2 x int32 each) -- built by the host once per set of tensors.
int b200f_grad_clip_coef(const void* table, int32_t n_tensors, int32_t n_chunks, float max_norm, float* scratch3, void* stream) {
  if (n_chunks == 0) return B200F_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const char* t = static_cast<const char*>(table);
  const float* const* grads = reinterpret_cast<const float* const*>(t + (size_t)n_tensors * 8);
  const long long* numel = reinterpret_cast<const long long*>(t + (size_t)n_tensors * 32);
  const ChunkRef* chunks = reinterpret_cast<const ChunkRef*>(t + (size_t)n_tensors * 40);
  B200F_CHECK_CUDA(cudaMemsetAsync(scratch3, 0, sizeof(float), st));
  grad_sumsq_kernel<<<n_chunks, 256, 0, st>>>(grads, numel, chunks, scratch3);
  int rc = check_launch("grad_sumsq");
  if (rc) return rc;
  clip_coef_kernel<<<1, 1, 0, st>>>(scratch3, max_norm, scratch3 + 1, scratch3 + 2);     // scratch3 = {sum of squares, norm, coefficient}
  return check_launch("clip_coef");
}

int b200f_adamw_step(const void* table, int32_t n_tensors, int32_t n_chunks, const float* clip_coef, double lr, double beta1, double beta2, double eps,
                     double weight_decay, int64_t step, void* stream) {
  if (n_chunks == 0) return B200F_OK;
  B200F_REQUIRE(step >= 1, B200F_ERR_SHAPE, "adamw: step must be >= 1");
  const char* t = static_cast<const char*>(table);
  float* const* params = reinterpret_cast<float* const*>(t);
  const float* const* grads = reinterpret_cast<const float* const*>(t + (size_t)n_tensors * 8);
  float* const* m = reinterpret_cast<float* const*>(t + (size_t)n_tensors * 16);
  float* const* v = reinterpret_cast<float* const*>(t + (size_t)n_tensors * 24);
  const long long* numel = reinterpret_cast<const long long*>(t + (size_t)n_tensors * 32);
  const ChunkRef* chunks = reinterpret_cast<const ChunkRef*>(t + (size_t)n_tensors * 40);
  // scalar factors in double on the host, rounded to fp32 once -- the way torch's Python-side scalars reach its kernels
  const float step_size = float(lr / (1.0 - pow(beta1, double(step))));
  const float sbc2 = float(sqrt(1.0 - pow(beta2, double(step))));
  adamw_kernel<<<n_chunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, m, v, numel, chunks, clip_coef, float(1.0 - lr * weight_decay),
                                                                        float(1.0 - beta1), float(beta2), float(1.0 - beta2), float(eps), step_size, sbc2);
  return check_launch("adamw");
}

}  // extern "C"
