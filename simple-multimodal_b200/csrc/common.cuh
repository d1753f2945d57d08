// Shared host/device helpers: error reporting across the C ABI, dtype traits, 128-bit vector IO.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200_fusion.h"

namespace b200f {

// thread-local last-error buffer (returned by b200f_last_error)
char* err_buf();
int fail(int code, const char* fmt, ...);

#define B200F_CHECK_CUDA(expr)                                                                   \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) return ::b200f::fail(B200F_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,  \
                                                cudaGetErrorString(_e), __FILE__, __LINE__);     \
  } while (0)

#define B200F_REQUIRE(cond, code, ...) \
  do {                                 \
    if (!(cond)) return ::b200f::fail(code, __VA_ARGS__); \
  } while (0)

// number of kernel launches issued by this library in this process (b200f_launch_count)
extern unsigned long long g_launches;

inline int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(B200F_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  return B200F_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int num_sms();

// cudaFuncSetAttribute is per device: "configured once" flags are kept per device ordinal (one process may drive several GPUs)
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// ---------------------------------------------------------------- dtype helpers
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float x) { return __float2bfloat16_rn(x); }

// A 16-byte vector of T: 4 floats or 8 bf16.
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float (&f)[4]) const { f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w; }
  __device__ __forceinline__ void pack(const float (&f)[4]) { raw = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const bf16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  __device__ __forceinline__ void pack(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    raw = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// counter-based RNG (splitmix64 finaliser over (seed, offset + index)) -> uniform in [0,1)
__host__ __device__ __forceinline__ float rng_uniform(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return float(z >> 40) * (1.0f / 16777216.0f);
}

// ---- dropout masks generated inside the kernels (no mask tensor in memory) -------------------------------------------
// An element is addressed by (row, col) of the matrix the dropout acts on -- attention probabilities: row = (b*H + h)*Lq + i,
// col = key j; an activation [M, N]: row m, col n -- and kept iff a 32-bit multiplicative hash of (row key, col) is >= thr32 =
// p * 2^32.  The row key is a full avalanche hash of (seed, row), computed once per row; the per-element part is 3 integer
// instructions, cheap enough for the MUFU-bound softmax loops.  Forward, backward and recomputation regenerate the same mask
// from (seed_lo, seed_hi).  tests/test_dropout.py restates this in numpy and checks rate / independence.
__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t drop_row_key(uint32_t seed_lo, uint32_t seed_hi, uint32_t row) {
  return hash32(seed_lo + hash32(seed_hi ^ row));
}
constexpr uint32_t kDropColMul = 0x9E3779B1u, kDropMix = 0x85EBCA6Bu;
__host__ __device__ __forceinline__ bool drop_keep_c(uint32_t row_key, uint32_t col_times_mul, uint32_t thr32) {
  return ((row_key ^ col_times_mul) * kDropMix) >= thr32;
}
__host__ __device__ __forceinline__ bool drop_keep(uint32_t row_key, uint32_t col, uint32_t thr32) {
  return drop_keep_c(row_key, col * kDropColMul, thr32);
}
// Attention-probability sites: the column term decomposes by XOR at 64-key granularity, term(col) = (col / 64) * MulHi ^ (col % 64) * Mul
// (== col * Mul for col < 64).  The softmax loops walk a row in 64- / 128-key tiles with the in-tile offsets known at compile time, so
// `row_key ^ tile term` is hoisted per tile and the element costs LOP3 (xor with an immediate) + IMAD + ISETP -- the integer add that
// `(tile base + offset) * Mul` needed per element is gone (1 of ~11.5 instructions per score in the issue-bound forward / dQ kernels).
constexpr uint32_t kDropColMulHi = 0xC2B2AE35u;
__host__ __device__ __forceinline__ uint32_t drop_col_attn(uint32_t col) { return ((col >> 6) * kDropColMulHi) ^ ((col & 63u) * kDropColMul); }
__host__ __device__ __forceinline__ bool drop_keep_attn(uint32_t row_key, uint32_t col, uint32_t thr32) {
  return drop_keep_c(row_key, drop_col_attn(col), thr32);
}
// Dropout epoch: a device-side word mixed into the row key by every kernel that generates a mask.  It exists for CUDA graphs:
// kernel arguments (the seeds) are frozen at capture, so a captured training step starts with a one-thread kernel that
// advances the epoch -- every replay then draws fresh masks, and forward / backward of one replay still agree.  Eager use
// leaves it at 0 (the key is then drop_row_key itself).  The epoch goes through its own avalanche hash and is ADDED to seed_lo,
// i.e. outside the inner hash that carries the row: XORing it next to the row index (the first version) made replay e's mask of
// row r equal replay 0's mask of row r ^ e -- a row permutation, not a fresh draw.  The library is built without relocatable
// device code, so each translation unit holds its own copy and b200f_dropout_epoch() updates all of them
// (B200F_DEFINE_EPOCH_HOOK).
static __device__ uint32_t g_drop_epoch = 0;
__host__ __device__ __forceinline__ uint32_t drop_epoch_mix(uint32_t epoch) {
  return epoch ? hash32(epoch * 0x9E3779B1u + 0x7F4A7C15u) : 0u;
}
__host__ __device__ __forceinline__ uint32_t drop_row_key_at(uint32_t seed_lo, uint32_t seed_hi, uint32_t row, uint32_t epoch) {
  return drop_row_key(seed_lo + drop_epoch_mix(epoch), seed_hi, row);
}
__device__ __forceinline__ uint32_t drop_row_key_e(uint32_t seed_lo, uint32_t seed_hi, uint32_t row) {
  return drop_row_key_at(seed_lo, seed_hi, row, g_drop_epoch);
}
#define B200F_DEFINE_EPOCH_HOOK(name)                                                                         \
  __global__ void name##_epoch_kernel(uint32_t v, int add) { g_drop_epoch = add ? g_drop_epoch + v : v; }      \
  void name##_epoch(uint32_t v, int add, cudaStream_t st) { name##_epoch_kernel<<<1, 1, 0, st>>>(v, add); }

__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  const double t = double(p) * 4294967296.0;
  return t <= 0.0 ? 0u : (t >= 4294967295.0 ? 4294967295u : uint32_t(t));
}

}  // namespace b200f
