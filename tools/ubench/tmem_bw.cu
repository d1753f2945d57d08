// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM as a function of the number of warps issuing it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../simple-multimodal_b200/csrc/ptx.cuh"
using namespace b200f;

template <int X>
__global__ void k(int iters, long long* out_cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (X == 32) {
      uint32_t r[32];
      tmem_ld32(base + ((i * 32) & 255), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
    } else {
      uint32_t r[16];
      tmem_ld16(base + ((i * 16) & 255), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= r[j];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) *out_cycles = t1 - t0;
  sink[threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

// two loads in flight before the wait
__global__ void k2(int iters, long long* out_cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i += 2) {
    uint32_t r[32], q[32];
    tmem_ld32(base + ((i * 32) & 255), r);
    tmem_ld32(base + ((i * 32 + 32) & 255), q);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= r[j] ^ q[j];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) *out_cycles = t1 - t0;
  sink[threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
  long long* d; uint32_t* s;
  cudaMalloc(&d, 8); cudaMalloc(&s, 4096);
  const int iters = 4096;
  for (int warps : {1, 2, 4, 8, 16}) {
    long long c;
    k<32><<<1, warps * 32>>>(iters, d, s); cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    k<32><<<1, warps * 32>>>(iters, d, s); cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("x32 warps=%2d  %lld cycles  %.1f B/cycle/SM  (%.1f cycles per ld per warp)\n", warps, c, double(warps) * iters * 32 * 32 * 4 / c, double(c) / iters);
    k<16><<<1, warps * 32>>>(iters, d, s); cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("x16 warps=%2d  %lld cycles  %.1f B/cycle/SM\n", warps, c, double(warps) * iters * 16 * 32 * 4 / c);
    k2<<<1, warps * 32>>>(iters, d, s); cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("x32x2 warps=%2d  %lld cycles  %.1f B/cycle/SM\n", warps, c, double(warps) * iters * 32 * 32 * 4 / c);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
