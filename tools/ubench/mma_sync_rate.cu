// Microbenchmark: issue rate of the legacy warp-level tensor path (mma.sync.m16n8k16 bf16 -> HMMA) on sm_100a, as a function of
// resident warps per SM and independent accumulator chains per warp.  Decides whether attn_narrow.cu (mma.sync) can be HMMA-bound.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mma_sync_rate tools/ubench/mma_sync_rate.cu && tools/ubench/mma_sync_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int CHAINS>
__global__ void hmma_loop(float* out, int iters) {
  float c[CHAINS][4];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 11, b0 = 13, b1 = threadIdx.x + 17;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) out[0] = s;
}

template <int CHAINS>
void run(int warps_per_sm, int sms) {
  const int iters = 20000;
  float* out;
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  hmma_loop<CHAINS><<<sms, warps_per_sm * 32>>>(out, 100);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  hmma_loop<CHAINS><<<sms, warps_per_sm * 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flop = 2.0 * 16 * 8 * 16 * double(iters) * CHAINS * warps_per_sm * sms;
  printf("warps/SM %2d  chains %2d : %8.1f TFLOP/s  (%.1f HMMA / clk / SM at 1.9 GHz)\n", warps_per_sm, CHAINS, flop / ms / 1e9,
         double(iters) * CHAINS * warps_per_sm / (ms * 1e-3 * 1.9e9));
  cudaFree(out);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs: %d\n", sms);
  for (int w : {4, 8, 16, 32}) {
    run<1>(w, sms);
    run<4>(w, sms);
    run<8>(w, sms);
    run<16>(w, sms);
  }
  return 0;
}
