// fp32 GEMM on the CUDA cores (FFMA) -- the parity-mode contraction (fp32 rtol 1e-5 rules out
// TF32 tensor cores).  Same contract and epilogue as the tcgen05 kernel in gemm_tc.cu.
// 64x64 output tile, 16-deep k slices staged in shared memory, 4x4 register micro-tile per thread.
#include "common.cuh"

namespace b200f {

template <typename T>
struct GemmSimtParams {
  long long M, N, K;
  const T* A; long long lda;
  const T* B; long long ldb;
  void* C; long long ldc;
  const float* bias;
  const T* residual; long long ldr;
  const T* mask; long long ldm;
  float alpha;
  int flags;
  int split_k;
  int c_is_f32;
};

template <typename T, int A_T, int B_T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmSimtParams<T> p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.y * 64, n0 = (long long)blockIdx.x * 64;
  const long long kchunk = (p.K + p.split_k - 1) / p.split_k;
  const long long kbeg = (long long)blockIdx.z * kchunk;
  const long long kend = kbeg + kchunk < p.K ? kbeg + kchunk : p.K;
  float acc[4][4] = {};
  for (long long k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;  // 0..1023
      int mm, kk;
      if (A_T) { mm = e & 63; kk = e >> 6; } else { kk = e & 15; mm = e >> 4; }
      const long long gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < p.M && gk < kend) v = to_f32(A_T ? p.A[gk * p.lda + gm] : p.A[gm * p.lda + gk]);
      As[kk][mm] = v;
      int nn;
      if (B_T) { nn = e & 63; kk = e >> 6; } else { kk = e & 15; nn = e >> 4; }
      const long long gn = n0 + nn, gk2 = k0 + kk;
      v = 0.f;
      if (gn < p.N && gk2 < kend) v = to_f32(B_T ? p.B[gk2 * p.ldb + gn] : p.B[gn * p.ldb + gk2]);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool accum = (p.flags & B200F_EPI_ACCUM) != 0;
  const bool relu = (p.flags & B200F_EPI_RELU) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (p.bias && blockIdx.z == 0) v += p.bias[n];
      if (p.residual && blockIdx.z == 0) v += to_f32(p.residual[m * p.ldr + n]);
      if (relu) v = fmaxf(v, 0.f);
      if (p.mask) v = to_f32(p.mask[m * p.ldm + n]) > 0.f ? v : 0.f;
      if (p.c_is_f32) {
        float* c = static_cast<float*>(p.C) + m * p.ldc + n;
        if (accum) {
          if (p.split_k > 1) atomicAdd(c, v); else *c += v;
        } else {
          *c = v;
        }
      } else {
        static_cast<T*>(p.C)[m * p.ldc + n] = from_f32<T>(v);
      }
    }
  }
}

template <typename T>
static int gemm_simt_launch(const b200f_gemm_args& a, cudaStream_t st) {
  B200F_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, B200F_ERR_SHAPE, "gemm: empty shape");
  GemmSimtParams<T> p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.A = static_cast<const T*>(a.A); p.lda = a.lda;
  p.B = static_cast<const T*>(a.B); p.ldb = a.ldb;
  p.C = a.C; p.ldc = a.ldc;
  p.bias = a.bias;
  p.residual = static_cast<const T*>(a.residual); p.ldr = a.ldr;
  p.mask = static_cast<const T*>(a.relu_mask); p.ldm = a.ldm;
  p.alpha = a.alpha; p.flags = a.flags;
  p.c_is_f32 = (sizeof(T) == 4) || (a.flags & (B200F_EPI_OUT_F32 | B200F_EPI_ACCUM)) ? 1 : 0;
  p.split_k = 1;
  if ((a.flags & B200F_EPI_ACCUM) && a.split_k > 1) {
    B200F_REQUIRE(!(a.flags & B200F_EPI_RELU) && !a.relu_mask, B200F_ERR_UNSUPPORTED, "gemm(simt): split_k with a non-linear epilogue");
    p.split_k = a.split_k;
  }
  dim3 grid((unsigned)((a.N + 63) / 64), (unsigned)((a.M + 63) / 64), (unsigned)p.split_k);
  B200F_REQUIRE(grid.y <= 65535, B200F_ERR_SHAPE, "gemm(simt): M too large for the CUDA-core kernel (%lld)", (long long)a.M);
  const int key = (a.a_layout ? 2 : 0) | (a.b_layout ? 1 : 0);
  switch (key) {
    case 0: gemm_simt_kernel<T, 0, 0><<<grid, 256, 0, st>>>(p); break;
    case 1: gemm_simt_kernel<T, 0, 1><<<grid, 256, 0, st>>>(p); break;
    case 2: gemm_simt_kernel<T, 1, 0><<<grid, 256, 0, st>>>(p); break;
    default: gemm_simt_kernel<T, 1, 1><<<grid, 256, 0, st>>>(p); break;
  }
  return check_launch("gemm_simt_kernel");
}

int gemm_f32_simt(const b200f_gemm_args& a, cudaStream_t st) { return gemm_simt_launch<float>(a, st); }
// ragged bf16 problems the TMA path cannot take (leading dimensions that are not 16-byte multiples, e.g. the
// 7-class logits or the 3 gate logits): HBM-bound GEMV-like work, CUDA cores with fp32 accumulation
int gemm_bf16_simt(const b200f_gemm_args& a, cudaStream_t st) { return gemm_simt_launch<bf16>(a, st); }

}  // namespace b200f
