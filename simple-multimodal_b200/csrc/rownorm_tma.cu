// LayerNorm forward / backward for the MulT blocks with the row tiles staged through shared memory by the TMA engine.
// (reference: nn.LayerNorm inside CrossModalTransformer, models/fusion_layers.py:182-211 -- norm1 / norm2, post-LN.)
//
// Both kernels are HBM-bound (forward: read x [+ two residual post-adds], write y; backward: read dy, x [+ dres], write dx).
// The register-staged kernels in rowops.cu alternate "issue loads -> reduce -> store" inside every warp, so the memory pipe
// idles during the reductions (ncu: 3.6 TB/s forward, 3.5-4.6 TB/s backward of 6.5 measured).  Here one producer thread
// streams contiguous row tiles (TILE rows x H, one `cp.async.bulk` per input) into a ring of shared-memory stages guarded by
// full/empty mbarriers, so 128-192 KB per SM is in flight all the time, and the consumer warps only read shared memory,
// do the arithmetic and store.  One persistent CTA per SM.  Rows are contiguous (ld == H), H * sizeof(T) <= 1 KB per
// 32-lane pass x NV passes (NV <= 2: H <= 512 in bf16, H <= 256 in fp32) -- other shapes keep the register-staged kernels.
#include "common.cuh"
#include "ptx.cuh"

namespace b200f {

__device__ __forceinline__ void bulk_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int VN>
__device__ __forceinline__ void load_param4(const float* __restrict__ p, float (&f)[VN]) {
#pragma unroll
  for (int j = 0; j < VN; j += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + j));
    f[j] = t.x; f[j + 1] = t.y; f[j + 2] = t.z; f[j + 3] = t.w;
  }
}

// Producer loop shared by both kernels: tile t of this CTA goes to stage (k mod stages); NIN inputs per stage, `slab` bytes apart.
template <typename T, int NIN>
__device__ __forceinline__ void ln_produce(unsigned char* smem, uint64_t* full, uint64_t* empty, int stages, uint32_t slab, int tile_rows,
                                           const T* in0, const T* in1, const T* in2, long long rows, int H) {
  int s = 0;
  uint32_t ph = 0;
  const long long ntiles = (rows + tile_rows - 1) / tile_rows;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    mbar_wait(empty + s, ph ^ 1);
    const long long row0 = t * tile_rows;
    const long long left = rows - row0;
    const uint32_t bytes = uint32_t(left < tile_rows ? left : tile_rows) * uint32_t(H) * uint32_t(sizeof(T));
    mbar_expect_tx(full + s, bytes * NIN);
    unsigned char* dst = smem + (size_t)s * NIN * slab;
    bulk_load_1d(dst, in0 + row0 * H, bytes, full + s);
    if (NIN >= 2) bulk_load_1d(dst + slab, in1 + row0 * H, bytes, full + s);
    if (NIN >= 3) bulk_load_1d(dst + 2 * (size_t)slab, in2 + row0 * H, bytes, full + s);
    if (++s == stages) { s = 0; ph ^= 1; }
  }
}

// ------------------------------------------------------------------------------------------------ forward
// y = LayerNorm(x) * gamma + beta (+ post1 + post2); mean / rstd saved per row.  CW consumer warps x R rows per tile.
template <typename T, int NV, int R, int NIN, int CW>
__global__ void __launch_bounds__((CW + 1) * 32, 1) layernorm_fwd_tma_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                                            const float* __restrict__ beta, const T* __restrict__ post1,
                                                                            const T* __restrict__ post2, T* __restrict__ y,
                                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                            long long rows, int H, float eps, int stages) {
  constexpr int VN = Vec16<T>::N;
  constexpr int TILE = CW * R;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t slab = uint32_t(TILE) * uint32_t(H) * uint32_t(sizeof(T));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * NIN * slab);
  uint64_t* empty = full + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, CW); }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == CW) {
    if (lane == 0) ln_produce<T, NIN>(smem, full, empty, stages, slab, TILE, x, post1, post2, rows, H);
    return;
  }
  const int nvec = H / VN;
  const float inv_h = 1.f / H;
  float gm[NV][VN], bt[NV][VN];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
#pragma unroll
    for (int j = 0; j < VN; ++j) { gm[i][j] = 0.f; bt[i][j] = 0.f; }
    if (vi < nvec) { load_param4<VN>(gamma + vi * VN, gm[i]); load_param4<VN>(beta + vi * VN, bt[i]); }
  }
  int s = 0;
  uint32_t ph = 0;
  const long long ntiles = (rows + TILE - 1) / TILE;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long row0 = t * TILE + warp * R;
    mbar_wait(full + s, ph);
    const T* sx = reinterpret_cast<const T*>(smem + (size_t)s * NIN * slab) + (size_t)(warp * R) * H;
    const T* sp1 = reinterpret_cast<const T*>(reinterpret_cast<const unsigned char*>(sx) + slab);
    const T* sp2 = reinterpret_cast<const T*>(reinterpret_cast<const unsigned char*>(sx) + 2 * (size_t)slab);
    Vec16<T> raw[R][NV], p1[NIN >= 2 ? R : 1][NIN >= 2 ? NV : 1], p2[NIN >= 3 ? R : 1][NIN >= 3 ? NV : 1];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
          const int off = r * H + vi * VN;
          raw[r][i].load(sx + off);
          if (NIN >= 2) p1[NIN >= 2 ? r : 0][NIN >= 2 ? i : 0].load(sp1 + off);
          if (NIN >= 3) p2[NIN >= 3 ? r : 0][NIN >= 3 ? i : 0].load(sp2 + off);
        }
      }
    // WAR across proxies: the generic-proxy reads above must be ordered before the async-proxy (TMA) write that refills this stage.
    // An mbarrier arrive alone does not order them (round 2: one row in ~1e4 launches was read after the refill had begun, always in
    // a CTA's first tiles where the memory pipe is most congested -- tools/repro_full_stash.py); the proxy fence does.
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);           // the stage is in registers: hand it back to the producer
    if (++s == stages) { s = 0; ph ^= 1; }

    float v[R][NV][VN], mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (lane + i * 32 < nvec) {
          raw[r][i].unpack(v[r][i]);
#pragma unroll
          for (int j = 0; j < VN; ++j) sum += v[r][i][j];
        }
      mean[r] = sum;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) mean[r] = warp_sum(mean[r]) * inv_h;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (lane + i * 32 < nvec) {
#pragma unroll
          for (int j = 0; j < VN; ++j) { const float d = v[r][i][j] - mean[r]; q += d * d; }
        }
      rstd[r] = q;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(warp_sum(rstd[r]) * inv_h + eps);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r >= rows) break;                    // warp-uniform (rows past the end of the last tile hold stale bytes)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
          float o[VN];
#pragma unroll
          for (int j = 0; j < VN; ++j) o[j] = (v[r][i][j] - mean[r]) * rstd[r] * gm[i][j] + bt[i][j];
          if (NIN >= 2) {
            float f[VN]; p1[NIN >= 2 ? r : 0][NIN >= 2 ? i : 0].unpack(f);
#pragma unroll
            for (int j = 0; j < VN; ++j) o[j] += f[j];
          }
          if (NIN >= 3) {
            float f[VN]; p2[NIN >= 3 ? r : 0][NIN >= 3 ? i : 0].unpack(f);
#pragma unroll
            for (int j = 0; j < VN; ++j) o[j] += f[j];
          }
          Vec16<T> tv; tv.pack(o); tv.store(y + (row0 + r) * H + vi * VN);
        }
      }
      if (lane == 0) { mean_out[row0 + r] = mean[r]; rstd_out[row0 + r] = rstd[r]; }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ dres), g = dy * gamma; dgamma += sum dy * xhat; dbeta += sum dy;
// dxsum += column sum of the stored dx (the bias gradient of the Linear that produced x).  Same arithmetic, in the same order
// per row, as layernorm_bwd_kernel in rowops.cu.
template <typename T, int NV, int NIN, int CW>
__global__ void __launch_bounds__((CW + 1) * 32, 1) layernorm_bwd_tma_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                            const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                                            const float* __restrict__ gamma, const T* __restrict__ dres,
                                                                            T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                            float* __restrict__ dxsum, long long rows, int H, int stages) {
  constexpr int VN = Vec16<T>::N;
  constexpr int TILE = CW;                           // one row per consumer warp per tile
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t slab = uint32_t(TILE) * uint32_t(H) * uint32_t(sizeof(T));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * NIN * slab);
  uint64_t* empty = full + stages;
  float* sred = reinterpret_cast<float*>(empty + stages);   // [3][H]: dgamma, dbeta, column sum of dx
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) sred[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, CW); }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == CW) {
    if (lane == 0) ln_produce<T, NIN>(smem, full, empty, stages, slab, TILE, dy, x, dres, rows, H);
  } else {
    const int nvec = H / VN;
    const float inv_h = 1.f / H;
    float pg[NV][VN], pb[NV][VN], gm[NV][VN], px[NV][VN];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
#pragma unroll
      for (int j = 0; j < VN; ++j) { pg[i][j] = 0.f; pb[i][j] = 0.f; px[i][j] = 0.f; gm[i][j] = 0.f; }
      if (vi < nvec) load_param4<VN>(gamma + vi * VN, gm[i]);
    }
    int s = 0;
    uint32_t ph = 0;
    const long long ntiles = (rows + TILE - 1) / TILE;
    // the row statistics come straight from global memory: fetched one tile ahead, so their latency hides behind this tile's work
    long long nrow = (long long)blockIdx.x * TILE + warp;
    float nmean = nrow < rows ? __ldg(mean_in + nrow) : 0.f, nrstd = nrow < rows ? __ldg(rstd_in + nrow) : 0.f;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const long long row = nrow;
      const bool ok = row < rows;                      // warp-uniform
      const float mean = nmean, rstd = nrstd;
      nrow = (t + gridDim.x) * TILE + warp;
      if (nrow < rows) { nmean = __ldg(mean_in + nrow); nrstd = __ldg(rstd_in + nrow); }
      mbar_wait(full + s, ph);
      const T* sdy = reinterpret_cast<const T*>(smem + (size_t)s * NIN * slab) + (size_t)warp * H;
      const T* sx = reinterpret_cast<const T*>(reinterpret_cast<const unsigned char*>(sdy) + slab);
      const T* sres = reinterpret_cast<const T*>(reinterpret_cast<const unsigned char*>(sdy) + 2 * (size_t)slab);
      Vec16<T> rdy[NV], rx[NV], rres[NIN >= 3 ? NV : 1];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
          rdy[i].load(sdy + vi * VN);
          rx[i].load(sx + vi * VN);
          if (NIN >= 3) rres[NIN >= 3 ? i : 0].load(sres + vi * VN);
        }
      }
      fence_proxy_async_smem();                       // generic reads of the stage before the TMA refill (see the forward kernel)
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
      if (++s == stages) { s = 0; ph ^= 1; }
      if (!ok) continue;
      float xh[NV][VN], g[NV][VN];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (lane + i * 32 < nvec) {
          float d[VN], xv[VN]; rdy[i].unpack(d); rx[i].unpack(xv);
#pragma unroll
          for (int j = 0; j < VN; ++j) {
            xh[i][j] = (xv[j] - mean) * rstd;
            g[i][j] = d[j] * gm[i][j];
            s1 += g[i][j];
            s2 += g[i][j] * xh[i][j];
            pg[i][j] += d[j] * xh[i][j];
            pb[i][j] += d[j];
          }
        }
      const float c1 = warp_sum(s1) * inv_h, c2 = warp_sum(s2) * inv_h;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
          float o[VN];
#pragma unroll
          for (int j = 0; j < VN; ++j) o[j] = rstd * (g[i][j] - c1 - xh[i][j] * c2);
          if (NIN >= 3) {
            float f[VN]; rres[NIN >= 3 ? i : 0].unpack(f);
#pragma unroll
            for (int j = 0; j < VN; ++j) o[j] += f[j];
          }
          Vec16<T> tv; tv.pack(o); tv.store(dx + row * H + vi * VN);
          if (dxsum) {            // sum what was stored (rounded to T), so it equals a column sum over the dx tensor
            float q[VN]; tv.unpack(q);
#pragma unroll
            for (int j = 0; j < VN; ++j) px[i][j] += q[j];
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          atomicAdd(&sred[vi * VN + j], pg[i][j]);
          atomicAdd(&sred[H + vi * VN + j], pb[i][j]);
          if (dxsum) atomicAdd(&sred[2 * H + vi * VN + j], px[i][j]);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    atomicAdd(dgamma + i, sred[i]);
    atomicAdd(dbeta + i, sred[H + i]);
    if (dxsum) atomicAdd(dxsum + i, sred[2 * H + i]);
  }
}

// ------------------------------------------------------------------------------------------------ launchers
static constexpr int kSmemBudget = 200 * 1024;      // of the 227 KB a CTA may use; the rest stays with L1

bool ln_tma_shape_ok(int64_t rows, int32_t H, int32_t dtype) {
  const int vn = dtype == B200F_F32 ? 4 : 8;
  return H % vn == 0 && (H / vn + 31) / 32 <= 2 && rows >= 64 && (size_t)H * (dtype == B200F_F32 ? 4 : 2) % 16 == 0;
}

template <typename T, int NV, int R, int NIN>
static int launch_fwd(const T* x, const float* gamma, const float* beta, const T* p1, const T* p2, T* y, float* mean, float* rstd, int64_t rows,
                      int32_t H, float eps, cudaStream_t st) {
  constexpr int CW = 15;   // + the producer warp = 16 warps, 4 per scheduler: 128 registers each
  auto kern = layernorm_fwd_tma_kernel<T, NV, R, NIN, CW>;
  const size_t stage = (size_t)NIN * CW * R * H * sizeof(T);
  int stages = int(kSmemBudget / stage);
  stages = stages > 8 ? 8 : stages;
  B200F_REQUIRE(stages >= 2, B200F_ERR_SHAPE, "layernorm (TMA): H=%d does not fit two stages", H);
  const size_t bytes = stages * stage + 2 * stages * sizeof(uint64_t);
  static PerDeviceOnce configured;
  if (configured.first()) {
    B200F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 1024));
  }
  const long long tiles = (rows + CW * R - 1) / (CW * R);
  const int grid = int(tiles < num_sms() ? tiles : num_sms());
  kern<<<grid, (CW + 1) * 32, bytes, st>>>(x, gamma, beta, p1, p2, y, mean, rstd, rows, H, eps, stages);
  return check_launch("layernorm_fwd_tma");
}

template <typename T, int NV, int NIN>
static int launch_bwd(const T* dy, const T* x, const float* mean, const float* rstd, const float* gamma, const T* dres, T* dx, float* dgamma,
                      float* dbeta, float* dxsum, int64_t rows, int32_t H, cudaStream_t st) {
  constexpr int CW = 11;   // + the producer warp = 12 warps, 3 per scheduler: 168 registers each
  auto kern = layernorm_bwd_tma_kernel<T, NV, NIN, CW>;
  const size_t stage = (size_t)NIN * CW * H * sizeof(T);
  const size_t tail = 3 * (size_t)H * sizeof(float);
  int stages = int((kSmemBudget - tail) / stage);
  stages = stages > 8 ? 8 : stages;
  B200F_REQUIRE(stages >= 2, B200F_ERR_SHAPE, "layernorm backward (TMA): H=%d does not fit two stages", H);
  const size_t bytes = stages * stage + 2 * stages * sizeof(uint64_t) + tail;
  static PerDeviceOnce configured;
  if (configured.first()) {
    B200F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 1024));
  }
  const long long tiles = (rows + CW - 1) / CW;
  const int grid = int(tiles < num_sms() ? tiles : num_sms());
  kern<<<grid, (CW + 1) * 32, bytes, st>>>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, dxsum, rows, H, stages);
  return check_launch("layernorm_bwd_tma");
}

template <typename T>
static int fwd_t(const void* x, const float* gamma, const float* beta, const void* post1, const void* post2, void* y, float* mean, float* rstd,
                 int64_t rows, int32_t H, float eps, cudaStream_t st) {
  constexpr int VN = Vec16<T>::N;
  const T* p1 = static_cast<const T*>(post1 ? post1 : post2);
  const T* p2 = static_cast<const T*>(post1 ? post2 : nullptr);
  const int nin = 1 + (p1 ? 1 : 0) + (p2 ? 1 : 0);
  const int nv = (H / VN + 31) / 32;
#define B200F_LN_FWD(NV_, R_, NIN_) \
  return launch_fwd<T, NV_, R_, NIN_>(static_cast<const T*>(x), gamma, beta, p1, p2, static_cast<T*>(y), mean, rstd, rows, H, eps, st)
  if (nv == 1) {
    if (nin == 1) B200F_LN_FWD(1, 2, 1);
    if (nin == 2) B200F_LN_FWD(1, 2, 2);
    B200F_LN_FWD(1, 2, 3);
  }
  if (nin == 1) B200F_LN_FWD(2, 2, 1);
  if (nin == 2) B200F_LN_FWD(2, 1, 2);
  B200F_LN_FWD(2, 1, 3);
#undef B200F_LN_FWD
}

template <typename T>
static int bwd_t(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, const void* dres, void* dx, float* dgamma,
                 float* dbeta, float* dxsum, int64_t rows, int32_t H, cudaStream_t st) {
  constexpr int VN = Vec16<T>::N;
  const int nv = (H / VN + 31) / 32;
#define B200F_LN_BWD(NV_, NIN_)                                                                                                          \
  return launch_bwd<T, NV_, NIN_>(static_cast<const T*>(dy), static_cast<const T*>(x), mean, rstd, gamma, static_cast<const T*>(dres), \
                                  static_cast<T*>(dx), dgamma, dbeta, dxsum, rows, H, st)
  if (nv == 1) {
    if (dres) B200F_LN_BWD(1, 3);
    B200F_LN_BWD(1, 2);
  }
  if (dres) B200F_LN_BWD(2, 3);
  B200F_LN_BWD(2, 2);
#undef B200F_LN_BWD
}

int layernorm_fwd_tma(const void* x, const float* gamma, const float* beta, const void* post1, const void* post2, void* y, float* mean, float* rstd,
                      int64_t rows, int32_t H, float eps, int32_t dtype, cudaStream_t st) {
  if (dtype == B200F_F32) return fwd_t<float>(x, gamma, beta, post1, post2, y, mean, rstd, rows, H, eps, st);
  return fwd_t<bf16>(x, gamma, beta, post1, post2, y, mean, rstd, rows, H, eps, st);
}

int layernorm_bwd_tma(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, const void* dres, void* dx,
                      float* dgamma, float* dbeta, float* dxsum, int64_t rows, int32_t H, int32_t dtype, cudaStream_t st) {
  if (dtype == B200F_F32) return bwd_t<float>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, dxsum, rows, H, st);
  return bwd_t<bf16>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, dxsum, rows, H, st);
}

}  // namespace b200f
