// Multi-head attention on the 5th-gen tensor cores (bf16, head_dim 64): flash-style, scores never leave the SM.
//
// Forward, one CTA per (128-query tile, head, batch), 6 warps:
//   warp 0    TMA producer : Q tile once, then a 2-deep ring of {K,V} tiles (128 keys) straight out of the packed
//                            projection output (4-D tensor map: d, head, token, batch; SWIZZLE_128B)
//   warp 1    MMA issuer   : S = Q K^T  (tcgen05.mma 128x128x16, both operands K-major)   -> TMEM cols [0,128)
//                            O += P V   (128x64x16, P K-major from smem, V MN-major)      -> TMEM cols [128,192)
//   warps 2-5 softmax      : one thread per query row (= TMEM lane): tcgen05.ld the scores, online max / exp2 /
//                            row-sum in fp32, P -> bf16 into the swizzled smem tile the PV MMA reads, rescale O in
//                            TMEM when the running max moves, final O / l -> global (128 B per row), LSE saved.
// 192 TMEM columns (256 allocated) and ~113 KB smem per CTA: two CTAs per SM overlap each other's softmax and MMA.
//
// Backward (two kernels, both recompute P from Q, K and the saved LSE; no atomics):
//   dQ   : per 128-query tile, loop over key tiles:  S, dP = dO V^T, dS = P o (dP - delta), dQ += dS K
//   dKdV : per 128-key tile, loop over query tiles, transposed so every thread-written operand is K-major:
//          S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q
#include "common.cuh"
#include "ptx.cuh"
#include <string.h>

namespace b200f {

static constexpr int ATT_THREADS = 192;
static constexpr int TQ = 128;   // query rows per CTA (UMMA M)
static constexpr int TK = 128;   // keys per tile
static constexpr int HD = 64;    // head dim
static constexpr float LOG2E = 1.4426950408889634f;

int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);
int g_attn_fwd_variant = 0;      // 0 = by shape, 1 = one 128-query tile per CTA, 2 = persistent two-tile kernel (A-B)

struct AttnTcParams {
  int B, H, Lq, Lk;
  float scale;
  bf16* O; long long ldo;
  float* LSE;
  // backward
  float* delta;                          // [B,H,Lq]: written by the dQ kernel (or attn_delta_kernel), read by the dKdV kernel
  bf16* dQ; long long lddq;
  bf16* dK; long long lddk;
  bf16* dV; long long lddv;
  float* dbq; float* dbk; float* dbv;   // optional [H*64] fp32 accumulators: column sums of dQ / dK / dV (in-proj bias gradients)
  float* pool;                           // forward, optional [B, ceil(Lq/32), H*64] fp32: partial column sums of the stored O, one per 32-row tile
  // attention-probability dropout (persistent kernels, DROP = true instantiation): keep iff hash >= drop_thr, kept scaled by inv_keep
  uint32_t drop_thr, drop_seed_lo, drop_seed_hi;
  float inv_keep;
};

// Epilogue helper: this warp's 32 accumulator rows x 64 fp32 columns in TMEM -> (x mul) -> bf16 -> a private [32 x 128 B]
// swizzled smem tile -> global, 4 full 128-byte rows per store instruction (a thread owns a ROW in TMEM, so direct
// stores would touch 32 different lines per instruction).  `gbase` points at (first row of this warp, first column).
// `colsum` (optional, 64 floats for this head): += the column sums of the rows stored (the bf16-rounded values, so it
// equals a column sum over the stored tensor) -- the bias gradient of the projection that produced Q / K / V.
__device__ __forceinline__ void store_rows64(uint32_t taddr, uint8_t* stage, bf16* gbase, long long ld, int rows_valid, int lane, float mul,
                                             float* colsum = nullptr, float* colpart = nullptr) {
#pragma unroll
  for (int cc = 0; cc < HD; cc += 16) {
    uint32_t r[16];
    tmem_ld16(taddr + cc, r);
    tmem_ld_wait();
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2) pk[i >> 1] = pack_bf16(__uint_as_float(r[i]) * mul, __uint_as_float(r[i + 1]) * mul);
    *reinterpret_cast<uint4*>(stage + sw128_offset(lane, cc >> 3)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(stage + sw128_offset(lane, (cc >> 3) + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
  __syncwarp();
  const int crow = lane >> 3, cchunk = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + crow;
    if (r < rows_valid)
      *reinterpret_cast<uint4*>(gbase + (long long)r * ld + cchunk * 8) = *reinterpret_cast<const uint4*>(stage + sw128_offset(r, cchunk));
  }
  if (colsum || colpart) {               // ONES x tile on the warp-level tensor path (ptx.cuh): lanes 0..3 hold columns 8*nt + 2*lane, +1
    const int nr = rows_valid < 32 ? rows_valid : 32;
    if (nr > 0) {                        // warp-uniform
      float cs[8][2];
      colsum32x64_hmma(stage, nr, lane, cs);
      if (lane < 4) {
        if (colpart) {                   // this tile's own slot: a plain store, bit-reproducible
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<float2*>(colpart + nt * 8 + 2 * lane) = make_float2(cs[nt][0], cs[nt][1]);
        } else {
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) { atomicAdd(colsum + nt * 8 + 2 * lane, cs[nt][0]); atomicAdd(colsum + nt * 8 + 2 * lane + 1, cs[nt][1]); }
        }
      }
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct FwdSmem {
  static constexpr int Q_OFF = 0;                       // [128 x 64] bf16, K-major SW128
  static constexpr int K_OFF = Q_OFF + TQ * HD * 2;     // 2 stages x [128 x 64]
  static constexpr int V_OFF = K_OFF + 2 * TK * HD * 2; // 2 stages x [128 keys x 64 d] (MN-major B operand)
  static constexpr int P_OFF = V_OFF + 2 * TK * HD * 2; // [128 x 128] bf16 as two K-major SW128 halves of 64 keys
  static constexpr int BAR_OFF = P_OFF + TQ * TK * 2;
  static constexpr int TOTAL = BAR_OFF + 128;
};

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BAR_OFF);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* pv_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.Lk + TK - 1) / TK;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();   // SWIZZLE_128B tiles need a 1024-byte aligned base
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base, t_o = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, TQ * HD * 2);
      tma_load_4d(smem + FwdSmem::Q_OFF, &tm_q, q_full, 0, h, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * TK * HD * 2);
        tma_load_4d(smem + FwdSmem::K_OFF + st * TK * HD * 2, &tm_k, &kv_full[st], 0, h, j * TK, b);
        tma_load_4d(smem + FwdSmem::V_OFF + st * TK * HD * 2, &tm_v, &kv_full[st], 0, h, j * TK, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(TQ, TK, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(TQ, HD, 0, 1);
      const uint32_t sq = smem_u32(smem + FwdSmem::Q_OFF), sp = smem_u32(smem + FwdSmem::P_OFF);
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        const uint32_t sk = smem_u32(smem + FwdSmem::K_OFF + st * TK * HD * 2);
        const uint32_t sv = smem_u32(smem + FwdSmem::V_OFF + st * TK * HD * 2);
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss(t_s, umma_desc(sq + k * 32, 16, 1024), umma_desc(sk + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(s_full);
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        // only the 16-key groups that hold valid keys: the softmax threads leave the P columns past ceil32(valid) untouched
        const int ksteps = (min(TK, p.Lk - j * TK) + 15) >> 4;
        for (int k = 0; k < ksteps; ++k)
          umma_ss(t_o, umma_desc(sp + (k >> 2) * (TQ * 128) + (k & 3) * 32, 16, 1024), umma_desc(sv + k * 2048, TK * 128, 1024), idesc_o,
                  (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
      }
    }
    __syncwarp();
  } else {
    const int grp = warp & 3;
    const int row = grp * 32 + lane;                 // row of the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(grp * 32) << 16;
    const float c = p.scale * LOG2E;
    const uint64_t c2 = f2_pack(c, c);
    float m = -INFINITY, l = 0.f;                    // m: reference max the stored P / O are scaled against (raw score units)
    uint8_t* prow = smem + FwdSmem::P_OFF;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1);                      // also implies PV(j-1) retired: O and the P tile are ours again
      tc_fence_after();
      const int valid = min(TK, p.Lk - j * TK);      // keys beyond Lk were zero-filled by TMA: mask them
      const int cols = (valid + 31) & ~31;
      const bool ragged = valid < cols;
      float tmax = -INFINITY, tmax_b = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < cols; cc += 32) {
        uint32_t r[32];
        tmem_ld32(t_s + lane_addr + cc, r);
        tmem_ld_wait();
        if (ragged && cc + 32 > valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc + i >= valid) r[i] = 0xff800000u;   // -inf
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          tmax = fmax3(tmax, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
          tmax_b = fmax3(tmax_b, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
        }
      }
      const float m_new = fmax3(m, tmax, tmax_b);
      // lazy rescale: move the reference max only when it would grow by more than 2^8, so P <= 2^8 (exact after the final O / l)
      const bool grow = (m_new - m) * c > 8.f;       // first tile: m = -inf -> true
      const float m_use = grow ? m_new : m;
      const float alpha = grow ? ex2_approx((m - m_use) * c) : 1.f;   // first tile: 0
      const float nmc = -m_use * c;
      const uint64_t nmc2 = f2_pack(nmc, nmc);
      uint64_t lsum2 = 0ull;
#pragma unroll 1
      for (int cc = 0; cc < cols; cc += 32) {
        uint32_t r[32];
        tmem_ld32(t_s + lane_addr + cc, r);
        tmem_ld_wait();
        if (ragged && cc + 32 > valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc + i >= valid) r[i] = 0xff800000u;   // ex2(-inf) = 0
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float x0, x1;
          f2_unpack(f2_fma(f2_pack_u(r[i], r[i + 1]), c2, nmc2), x0, x1);
          const float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
          lsum2 = f2_add(lsum2, f2_pack(e0, e1));
          pk[i >> 1] = pack_bf16(e0, e1);
        }
        // 32 keys = four 16-byte chunks of this row inside the 64-key half (cc >> 6)
        uint8_t* half = prow + (cc >> 6) * (TQ * 128);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t chunk = ((cc & 63) >> 3) + q4;
          *reinterpret_cast<uint4*>(half + sw128_offset(row, chunk)) = make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
        }
      }
      float ls0, ls1;
      f2_unpack(lsum2, ls0, ls1);
      l = l * alpha + (ls0 + ls1);
      m = m_use;
      if (j > 0 && __any_sync(0xffffffffu, grow)) {    // O was accumulated against the old reference max: rescale it in TMEM
#pragma unroll
        for (int cc = 0; cc < HD; cc += 16) {
          uint32_t r[16];
          tmem_ld16(t_o + lane_addr + cc, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st16(t_o + lane_addr + cc, r);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      fence_proxy_async_smem();                        // P (generic-proxy stores) -> visible to the UMMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l;
    const int qrow = q0 + row;
    // O / l -> bf16 -> this warp's 32 x 128 B slice of the (now free) P tile -> full 128-byte lines to global
    if (qrow < p.Lq) p.LSE[((long long)b * p.H + h) * p.Lq + qrow] = m * p.scale + logf(l);
    store_rows64(t_o + lane_addr, smem + FwdSmem::P_OFF + grp * (32 * 128), p.O + ((long long)b * p.Lq + q0 + grp * 32) * p.ldo + h * HD, p.ldo,
                 p.Lq - (q0 + grp * 32), lane, inv, nullptr,
                 p.pool ? p.pool + (((long long)b * ((p.Lq + 31) / 32) + (q0 + grp * 32) / 32) * p.H + h) * HD : nullptr);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// forward v2 (default): persistent, one CTA per SM, TWO 128-query tiles in flight.
//   warp 8      TMA producer : Q tiles of the next work item, 3-deep ring of {K,V} tiles
//   warp 9      MMA issuer   : per key tile j and query tile t:  PV_t(j), then S_t(j+1) -- so while warpgroup t runs the
//                              softmax of S_t(j+1) the tensor pipe works on the other tile, and the two warpgroups keep
//                              the MUFU pipe (the real bound at head dim 64: one exp per 256 MMA FLOPs) continuously busy
//   warps 0-3   softmax warpgroup 0 (query rows 0..127 of the item),  warps 4-7  warpgroup 1 (rows 128..255); 224 registers
//               each after setmaxnreg (a whole 128-score row stays in registers: the scores are read from TMEM once)
// TMEM: S_0 [0,128) S_1 [128,256) O_0 [256,320) O_1 [320,384).  A work item = (batch, head, 256-query block).
// ------------------------------------------------------------------------------------------------
static constexpr int F2_THREADS = 384;     // warps 0-3 softmax warpgroup 0, 4-7 warpgroup 1, 8 TMA, 9 MMA, 10-11 idle (complete the third warpgroup)
static constexpr int SOFTMAX_REGS = 216, IO_REGS = 72;   // setmaxnreg: 256 x 216 + 128 x 72 = 384 x 168 (the launch allocation)
static constexpr int F2_ST = 3;      // K/V ring depth
struct Fwd2Smem {
  static constexpr int Q_OFF = 0;                              // 2 x [128 x 64] bf16, K-major SW128
  static constexpr int K_OFF = Q_OFF + 2 * TQ * HD * 2;        // F2_ST x [128 keys x 64]
  static constexpr int V_OFF = K_OFF + F2_ST * TK * HD * 2;    // F2_ST x [128 keys x 64 d] (MN-major B operand)
  static constexpr int P_OFF = V_OFF + F2_ST * TK * HD * 2;    // 2 x [128 x 128] bf16, each two K-major SW128 halves of 64 keys
  static constexpr int BAR_OFF = P_OFF + 2 * TQ * TK * 2;
  static constexpr int TOTAL = BAR_OFF + 256;
};

template <bool DROP>
__global__ void __launch_bounds__(F2_THREADS, 1)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const AttnTcParams p, const int n_qblk, const int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Fwd2Smem::BAR_OFF);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;               // [F2_ST]
  uint64_t* kv_empty = kv_full + F2_ST;       // [F2_ST]
  uint64_t* s_full = kv_empty + F2_ST;        // [2]  S_t(j) complete
  uint64_t* s_free = s_full + 2;              // [2]  S_t(j) is in the softmax warps' registers: the MMA warp may overwrite it with S_t(j+1)
  uint64_t* p_ready = s_free + 2;             // [2]  P_t(j) in smem, O_t rescaled
  uint64_t* pv_done = p_ready + 2;            // [2]  PV_t(j) retired (j < last): O_t and the P_t tile belong to the softmax warps again
  uint64_t* o_full = pv_done + 2;             // [2]  last PV_t of the item complete
  uint64_t* o_empty = o_full + 2;             // [2]  epilogue has read O_t
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.Lk + TK - 1) / TK;

  if (warp == 8 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < F2_ST; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1); mbar_init(&s_free[t], 4); mbar_init(&p_ready[t], 4); mbar_init(&pv_done[t], 1);
      mbar_init(&o_full[t], 1); mbar_init(&o_empty[t], 4);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
  setmaxnreg_dec<IO_REGS>();
  if (warp == 8) {
    if (lane == 0) {
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int qb = item % n_qblk, h = (item / n_qblk) % p.H, b = item / (n_qblk * p.H);
        const int q0 = qb * 2 * TQ;
        const bool two = q0 + TQ < p.Lq;
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_expect_tx(q_full, (two ? 2 : 1) * TQ * HD * 2);
        tma_load_4d(smem + Fwd2Smem::Q_OFF, &tm_q, q_full, 0, h, q0, b);
        if (two) tma_load_4d(smem + Fwd2Smem::Q_OFF + TQ * HD * 2, &tm_q, q_full, 0, h, q0 + TQ, b);
        for (int j = 0; j < n_tiles; ++j, ++g) {
          const int st = g % F2_ST;
          mbar_wait(&kv_empty[st], ((g / F2_ST) & 1) ^ 1);
          mbar_expect_tx(&kv_full[st], 2 * TK * HD * 2);
          tma_load_4d(smem + Fwd2Smem::K_OFF + st * TK * HD * 2, &tm_k, &kv_full[st], 0, h, j * TK, b);
          tma_load_4d(smem + Fwd2Smem::V_OFF + st * TK * HD * 2, &tm_v, &kv_full[st], 0, h, j * TK, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    {
      const bool leader = elect_one();   // warp-uniform loops; only the MMA / commit instructions are predicated
      constexpr uint32_t idesc_s = umma_idesc_bf16(TQ, TK, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(TQ, HD, 0, 1);
      uint32_t g = 0, it = 0, pr_cnt[2] = {0, 0}, o_cnt[2] = {0, 0}, s_cnt[2] = {0, 0};
      // S_t(j+1) is issued as soon as the softmax warps hold S_t(j) in registers (s_free), i.e. it runs under their exp work;
      // PV_t(j) follows when P_t(j) is in smem.  Order on the tensor pipe per key tile: S_0(j+1) S_1(j+1) PV_0(j) PV_1(j).
      // low descriptor words of the resident operands (see umma_lo): K-major tiles step 32 B (2 units) per K=16, the
      // MN-major V tile steps 16 key rows = 2048 B (128 units); stages / tiles are (bytes >> 4) apart
      const uint32_t q_lo = umma_lo(smem_u32(smem + Fwd2Smem::Q_OFF), 16), k_lo = umma_lo(smem_u32(smem + Fwd2Smem::K_OFF), 16);
      const uint32_t v_lo = umma_lo(smem_u32(smem + Fwd2Smem::V_OFF), TK * 128), p_lo = umma_lo(smem_u32(smem + Fwd2Smem::P_OFF), 16);
      constexpr uint32_t TILE16 = TQ * HD * 2 / 16, PT16 = TQ * TK * 2 / 16;
      auto issue_s = [&](int t, int st) {
        mbar_wait(&s_free[t], (s_cnt[t] & 1) ^ 1);     // first use passes immediately
        ++s_cnt[t];
        tc_fence_after();
        if (leader) umma_chain<HD / 16>(tmem_base + t * TK, q_lo + t * TILE16, 2, k_lo + st * TILE16, 2, idesc_s, 0);
        if (leader) umma_commit(&s_full[t]);
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int qb = item % n_qblk;
        const int nt = (qb * 2 * TQ + TQ < p.Lq) ? 2 : 1;
        mbar_wait(q_full, it & 1);
        {
          const int st = g % F2_ST;
          mbar_wait(&kv_full[st], (g / F2_ST) & 1);
          tc_fence_after();
          for (int t = 0; t < nt; ++t) issue_s(t, st);
          if (n_tiles == 1) if (leader) umma_commit(q_empty);
        }
        for (int j = 0; j < n_tiles; ++j) {
          const int st = (g + j) % F2_ST;
          if (j + 1 < n_tiles) {
            const int stn = (g + j + 1) % F2_ST;
            mbar_wait(&kv_full[stn], ((g + j + 1) / F2_ST) & 1);
            tc_fence_after();
            for (int t = 0; t < nt; ++t) issue_s(t, stn);
            if (j + 2 == n_tiles) if (leader) umma_commit(q_empty);
          }
          const int ksteps = (min(TK, p.Lk - j * TK) + 15) >> 4;   // P columns past ceil32(valid) are never written
          for (int t = 0; t < nt; ++t) {
            mbar_wait(&p_ready[t], pr_cnt[t] & 1);
            ++pr_cnt[t];
            if (j == 0) mbar_wait(&o_empty[t], (o_cnt[t] & 1) ^ 1);   // the previous item's epilogue has drained O_t
            tc_fence_after();
            const uint32_t t_o = tmem_base + 2 * TK + t * HD;
            const uint32_t pa = p_lo + t * PT16, vb = v_lo + st * TILE16;     // P: two K-major halves of 64 keys, (TQ * 128) B apart
            if (ksteps == 8) {
              if (leader) umma_chain<4>(t_o, pa, 2, vb, 128, idesc_o, j > 0);
              if (leader) umma_chain<4>(t_o, pa + TQ * 128 / 16, 2, vb + 4 * 128, 128, idesc_o, 1);
            } else {
              for (int k = 0; k < ksteps; ++k)
                if (leader) umma_ss(t_o, umma_desc_lo(pa + (k >> 2) * (TQ * 128 / 16) + (k & 3) * 2), umma_desc_lo(vb + k * 128), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
            }
            if (j + 1 < n_tiles) {
              if (leader) umma_commit(&pv_done[t]);
            } else {
              if (leader) umma_commit(&o_full[t]);
              ++o_cnt[t];
            }
          }
          if (leader) umma_commit(&kv_empty[st]);
        }
        g += n_tiles;
      }
    }
    __syncwarp();
  }
  } else {
    setmaxnreg_inc<SOFTMAX_REGS>();
    const int t = warp >> 2;                         // query tile / warpgroup
    const int grp = warp & 3;                        // TMEM lane group this warp may touch
    const int row = grp * 32 + lane;                 // row of the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(grp * 32) << 16;
    const uint32_t t_s = tmem_base + t * TK, t_o = tmem_base + 2 * TK + t * HD;
    const float c = p.scale * LOG2E;
    const uint64_t c2 = f2_pack(c, c);
    uint8_t* prow = smem + Fwd2Smem::P_OFF + t * TQ * TK * 2;
    uint32_t sf_cnt = 0, o_cnt = 0, pv_cnt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int qb = item % n_qblk, h = (item / n_qblk) % p.H, b = item / (n_qblk * p.H);
      const int q0 = qb * 2 * TQ + t * TQ;
      if (q0 >= p.Lq) continue;                      // this warpgroup's tile does not exist (the MMA warp skips it too)
      float m = -INFINITY, l = 0.f;                  // m: reference max the stored P / O are scaled against (raw score units)
      // dropout: P (what multiplies V) is masked, the row sum l is not; the 1/(1-p) scale rides on the final 1/l
      const uint32_t rk = DROP ? drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, uint32_t((b * p.H + h) * p.Lq + q0 + row)) : 0u;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&s_full[t], sf_cnt & 1);
        ++sf_cnt;
        tc_fence_after();
        const int valid = min(TK, p.Lk - j * TK);    // keys beyond Lk were zero-filled by TMA: mask them
        const int cols = (valid + 31) & ~31;
        const bool ragged = valid < cols;
        // ONE pass over the scores: the whole row (up to 128 fp32) is loaded into registers with all tcgen05.ld in flight
        // together (a load costs ~94 cycles per warp: serialising 8 of them per tile was a third of the softmax time)
        uint32_t r[4][32];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q * 32 < cols) tmem_ld32(t_s + lane_addr + q * 32, r[q]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);        // S_t is in registers: the next S_t may be computed under this tile's softmax
        float tmax = -INFINITY, tmax_b = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q * 32 < cols) {
            if (ragged && q * 32 + 32 > valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (q * 32 + i >= valid) r[q][i] = 0xff800000u;   // -inf: ex2(-inf) = 0
            }
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              tmax = fmax3(tmax, __uint_as_float(r[q][i]), __uint_as_float(r[q][i + 1]));
              tmax_b = fmax3(tmax_b, __uint_as_float(r[q][i + 2]), __uint_as_float(r[q][i + 3]));
            }
          }
        }
        const float m_new = fmax3(m, tmax, tmax_b);
        // lazy rescale: move the reference max only when it would grow by more than 2^8, so P <= 2^8 (exact after the final O / l)
        const bool grow = (m_new - m) * c > 8.f;     // first tile: m = -inf -> true
        const float m_use = grow ? m_new : m;
        const float alpha = grow ? ex2_approx((m - m_use) * c) : 1.f;   // first tile: 0
        const float nmc = -m_use * c;
        const uint64_t nmc2 = f2_pack(nmc, nmc);
        uint64_t lsum2 = 0ull;
        const uint32_t rkh[2] = {rk ^ (uint32_t(2 * j) * kDropColMulHi), rk ^ (uint32_t(2 * j + 1) * kDropColMulHi)};   // dropout: row key ^ 64-key group term
        if (j > 0) {                                   // PV_t(j-1) must have retired before P_t is overwritten / O_t rescaled
          mbar_wait(&pv_done[t], pv_cnt & 1);          // (holding all of P_t in registers to wait later spills at 216 registers: measured slower)
          ++pv_cnt;
          tc_fence_after();
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q * 32 < cols) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float x0, x1;
              f2_unpack(f2_fma(f2_pack_u(r[q][i], r[q][i + 1]), c2, nmc2), x0, x1);
              float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
              lsum2 = f2_add(lsum2, f2_pack(e0, e1));
              if (DROP) {                            // key = j*128 + q*32 + i: 64-key group 2j + q/2 (hoisted xor), offset (q%2)*32 + i (immediate)
                e0 = drop_keep_c(rkh[q >> 1], uint32_t((q & 1) * 32 + i) * kDropColMul, p.drop_thr) ? e0 : 0.f;
                e1 = drop_keep_c(rkh[q >> 1], uint32_t((q & 1) * 32 + i + 1) * kDropColMul, p.drop_thr) ? e1 : 0.f;
              }
              pk[i >> 1] = pack_bf16(e0, e1);
            }
            uint8_t* half = prow + (q >> 1) * (TQ * 128);   // 32 keys = four 16-byte chunks of this row inside the 64-key half
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const uint32_t chunk = (q & 1) * 4 + q4;
              *reinterpret_cast<uint4*>(half + sw128_offset(row, chunk)) = make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
            }
          }
        }
        float ls0, ls1;
        f2_unpack(lsum2, ls0, ls1);
        l = l * alpha + (ls0 + ls1);
        m = m_use;
        if (j > 0 && __any_sync(0xffffffffu, grow)) {  // O was accumulated against the old reference max: rescale it in TMEM
#pragma unroll
          for (int cc = 0; cc < HD; cc += 16) {
            uint32_t r[16];
            tmem_ld16(t_o + lane_addr + cc, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st16(t_o + lane_addr + cc, r);
          }
          tmem_st_wait();
        }
        tc_fence_before();
        fence_proxy_async_smem();                      // P (generic-proxy stores) -> visible to the UMMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t]);
      }
      mbar_wait(&o_full[t], o_cnt & 1);
      ++o_cnt;
      tc_fence_after();
      const int qrow = q0 + row;
      if (qrow < p.Lq) p.LSE[((long long)b * p.H + h) * p.Lq + qrow] = m * p.scale + logf(l);
      // O / l -> bf16 -> this warp's 32 x 128 B slice of the (now free) P_t tile -> full 128-byte lines to global
      store_rows64(t_o + lane_addr, prow + grp * (32 * 128), p.O + ((long long)b * p.Lq + q0 + grp * 32) * p.ldo + h * HD, p.ldo,
                   p.Lq - (q0 + grp * 32), lane, (DROP ? p.inv_keep : 1.f) / l, nullptr,
                   p.pool ? p.pool + (((long long)b * ((p.Lq + 31) / 32) + (q0 + grp * 32) / 32) * p.H + h) * HD : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[t]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

// 4-D view (d, head, token, batch) of a token-major [B, L, ld] tensor slice; box = 64 x 1 x rows x 1
static int make_head_tmap(CUtensorMap* m, const void* base, long long ld, int B, int H, int L, int rows) {
  uint64_t dims[4] = {uint64_t(HD), uint64_t(H), uint64_t(L), uint64_t(B)};
  uint64_t strides[3] = {uint64_t(HD) * 2, uint64_t(ld) * 2, uint64_t(L) * uint64_t(ld) * 2};
  uint32_t box[4] = {uint32_t(HD), 1, uint32_t(rows), 1};
  return make_tmap_bf16(m, base, 4, dims, strides, box);
}

static int attn_tc_check(const b200f_attn_args& a) {
  B200F_REQUIRE(a.D == HD, B200F_ERR_UNSUPPORTED, "attention(tcgen05): head dim must be 64");
  B200F_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0, B200F_ERR_ALIGN, "attention(tcgen05): leading dims must be multiples of 8");
  B200F_REQUIRE(aligned16(a.Q) && aligned16(a.K) && aligned16(a.V) && aligned16(a.O), B200F_ERR_ALIGN, "attention(tcgen05): 16-byte alignment");
  B200F_REQUIRE(a.H <= 65535 && a.B <= 65535, B200F_ERR_SHAPE, "attention(tcgen05): grid limits");
  return B200F_OK;
}

int attn_fwd_tc(const b200f_attn_args& a, cudaStream_t st) {
  int rc = attn_tc_check(a);
  if (rc) return rc;
  CUtensorMap tq, tk, tv;
  if ((rc = make_head_tmap(&tq, a.Q, a.ldq, a.B, a.H, a.Lq, TQ))) return rc;
  if ((rc = make_head_tmap(&tk, a.K, a.ldk, a.B, a.H, a.Lk, TK))) return rc;
  if ((rc = make_head_tmap(&tv, a.V, a.ldv, a.B, a.H, a.Lk, TK))) return rc;
  AttnTcParams p = {};
  p.B = a.B; p.H = a.H; p.Lq = a.Lq; p.Lk = a.Lk; p.scale = a.scale;
  p.O = static_cast<bf16*>(a.O); p.ldo = a.ldo; p.LSE = a.LSE;
  p.pool = a.pool_sum;
  static PerDeviceOnce configured;
  if (configured.first()) {
    B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
    B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd2Smem::TOTAL));
    B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd2Smem::TOTAL));
  }
  // a single 128-query tile per (batch, head) leaves the second softmax warpgroup of the two-tile kernel idle: use the
  // one-tile-per-CTA kernel there (measured Lq=30, Lk=512: 0.094 ms vs 0.122 ms).  b200f_debug_set(4, 1|2) forces either.
  p.drop_thr = drop_threshold(a.dropout_p); p.drop_seed_lo = a.drop_seed_lo; p.drop_seed_hi = a.drop_seed_hi;
  p.inv_keep = 1.f / (1.f - a.dropout_p);
  const bool drop = p.drop_thr != 0;
  B200F_REQUIRE(!drop || g_attn_fwd_variant != 1, B200F_ERR_UNSUPPORTED, "attention(tcgen05): dropout needs the persistent kernel");
  if (!drop && (g_attn_fwd_variant == 1 || (g_attn_fwd_variant == 0 && a.Lq <= TQ))) {
    dim3 grid((a.Lq + TQ - 1) / TQ, a.H, a.B);
    attn_fwd_tc_kernel<<<grid, ATT_THREADS, FwdSmem::TOTAL, st>>>(tq, tk, tv, p);
    return check_launch("attn_fwd_tc_kernel");
  }
  const int n_qblk = (a.Lq + 2 * TQ - 1) / (2 * TQ);
  const long long n_items = (long long)n_qblk * a.H * a.B;
  B200F_REQUIRE(n_items < (1ll << 31), B200F_ERR_SHAPE, "attention(tcgen05): too many work items");
  const int grid = int(n_items < num_sms() ? n_items : num_sms());
  if (drop) attn_fwd_tc2_kernel<true><<<grid, F2_THREADS, Fwd2Smem::TOTAL, st>>>(tq, tk, tv, p, n_qblk, int(n_items));
  else attn_fwd_tc2_kernel<false><<<grid, F2_THREADS, Fwd2Smem::TOTAL, st>>>(tq, tk, tv, p, n_qblk, int(n_items));
  return check_launch("attn_fwd_tc2_kernel");
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
static constexpr int BT = 64;    // inner tile (keys for the dQ kernel, queries for the dKdV kernel)

// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d].  Thread block (8, H, Z): eight lanes share one 128-byte (row, head) segment (16 bytes
// each -- a warp instruction reads 512 contiguous bytes of a token row), Z token rows per pass, U passes in flight per thread; no index
// division on the load path (the first version spent more issue slots on 64-bit div / mod than the memory system needed time).
static constexpr int DELTA_U = 4;
__global__ void attn_delta_kernel(const bf16* __restrict__ dO, long long lddo, const bf16* __restrict__ O, long long ldo, float* __restrict__ delta,
                                  int B, int H, int Lq) {
  const int part = threadIdx.x, h = threadIdx.y;
  const long long n_rows = (long long)B * Lq;
  const long long row0 = ((long long)blockIdx.x * DELTA_U) * blockDim.z + threadIdx.z;
  Vec16<bf16> a[DELTA_U], b2[DELTA_U];
#pragma unroll
  for (int u = 0; u < DELTA_U; ++u) {
    const long long bl = row0 + (long long)u * blockDim.z;
    if (bl < n_rows) {
      a[u].load(dO + bl * lddo + h * HD + part * 8);
      b2[u].load(O + bl * ldo + h * HD + part * 8);
    }
  }
#pragma unroll
  for (int u = 0; u < DELTA_U; ++u) {
    const long long bl = row0 + (long long)u * blockDim.z;     // uniform over the 8 lanes of a segment; a warp holds whole segments
    float acc = 0.f;
    if (bl < n_rows) {
      float fa[8], fb[8];
      a[u].unpack(fa); b2[u].unpack(fb);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc += fa[e] * fb[e];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (bl < n_rows && part == 0) {
      const long long b = bl / Lq, i = bl - b * Lq;
      delta[(b * H + h) * Lq + i] = acc;
    }
  }
}
static int launch_delta(const b200f_attn_args& a, cudaStream_t st) {
  const int z = a.H * 8 >= 256 ? 1 : 256 / (a.H * 8);          // H = 8: 4 token rows per pass, 256 threads
  B200F_REQUIRE(a.H * 8 <= 1024, B200F_ERR_UNSUPPORTED, "attention(tcgen05): more than 128 heads");
  const long long n_rows = (long long)a.B * a.Lq, per_block = (long long)DELTA_U * z;
  attn_delta_kernel<<<(unsigned)((n_rows + per_block - 1) / per_block), dim3(8, a.H, z), 0, st>>>(
      static_cast<const bf16*>(a.dO), a.lddo, static_cast<const bf16*>(a.O), a.ldo, a.delta, a.B, a.H, a.Lq);
  return check_launch("attn_delta_kernel");
}

struct DqSmem {
  static constexpr int Q_OFF = 0;                          // [128 x 64]
  static constexpr int DO_OFF = Q_OFF + TQ * HD * 2;       // [128 x 64]
  static constexpr int K_OFF = DO_OFF + TQ * HD * 2;       // 2 x [64 keys x 64]
  static constexpr int V_OFF = K_OFF + 2 * BT * HD * 2;    // 2 x [64 keys x 64]
  static constexpr int DS_OFF = V_OFF + 2 * BT * HD * 2;   // [128 x 64 keys] bf16, K-major SW128
  static constexpr int BAR_OFF = DS_OFF + TQ * BT * 2;
  static constexpr int TOTAL = BAR_OFF + 128;
};

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                      const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DqSmem::BAR_OFF);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* ds_ready = bars + 6;
  uint64_t* dq_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.Lk + BT - 1) / BT;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(ds_ready, 4);
    mbar_init(dq_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base, t_dp = tmem_base + 64, t_dq = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * TQ * HD * 2);
      tma_load_4d(smem + DqSmem::Q_OFF, &tm_q, q_full, 0, h, q0, b);
      tma_load_4d(smem + DqSmem::DO_OFF, &tm_do, q_full, 0, h, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * BT * HD * 2);
        tma_load_4d(smem + DqSmem::K_OFF + st * BT * HD * 2, &tm_k, &kv_full[st], 0, h, j * BT, b);
        tma_load_4d(smem + DqSmem::V_OFF + st * BT * HD * 2, &tm_v, &kv_full[st], 0, h, j * BT, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(TQ, BT, 0, 0);     // S = Q K^T, dP = dO V^T   (N = 64 keys)
      constexpr uint32_t idesc_dq = umma_idesc_bf16(TQ, HD, 0, 1);    // dQ += dS K             (K MN-major, N = d)
      const uint32_t sq = smem_u32(smem + DqSmem::Q_OFF), sdo = smem_u32(smem + DqSmem::DO_OFF), sds = smem_u32(smem + DqSmem::DS_OFF);
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        const uint32_t sk = smem_u32(smem + DqSmem::K_OFF + st * BT * HD * 2);
        const uint32_t sv = smem_u32(smem + DqSmem::V_OFF + st * BT * HD * 2);
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss(t_s, umma_desc(sq + k * 32, 16, 1024), umma_desc(sk + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss(t_dp, umma_desc(sdo + k * 32, 16, 1024), umma_desc(sv + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(sdp_full);
        mbar_wait(ds_ready, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BT / 16; ++k)
          umma_ss(t_dq, umma_desc(sds + k * 32, 16, 1024), umma_desc(sk + k * 2048, BT * 128, 1024), idesc_dq, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&kv_empty[st]);
        umma_commit(dq_done);
      }
    }
    __syncwarp();
  } else {
    const int grp = warp & 3;
    const int row = grp * 32 + lane;
    const uint32_t lane_addr = uint32_t(grp * 32) << 16;
    const int qrow = q0 + row;
    const long long ri = ((long long)b * p.H + h) * p.Lq + qrow;
    const float c = p.scale * LOG2E;
    const float nlse2 = -(qrow < p.Lq ? p.LSE[ri] : 0.f) * LOG2E;
    const float ndl = -(qrow < p.Lq ? p.delta[ri] : 0.f) * p.scale;
    const uint64_t c2 = f2_pack(c, c), nlse22 = f2_pack(nlse2, nlse2), sc2 = f2_pack(p.scale, p.scale), ndl2 = f2_pack(ndl, ndl);
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(sdp_full, j & 1);      // also implies the previous tile's dQ MMAs (which read the dS tile) retired
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < BT; cc += 32) {
        uint32_t rs[32], rp[32];
        tmem_ld32(t_s + lane_addr + cc, rs);
        tmem_ld32(t_dp + lane_addr + cc, rp);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {     // dS = P o (dP - delta) * scale, two elements per FFMA2 / FMUL2
          float x0, x1, d0, d1;
          f2_unpack(f2_fma(f2_pack_u(rs[i], rs[i + 1]), c2, nlse22), x0, x1);
          const uint64_t t2 = f2_fma(f2_pack_u(rp[i], rp[i + 1]), sc2, ndl2);
          f2_unpack(f2_mul(f2_pack(ex2_approx(x0), ex2_approx(x1)), t2), d0, d1);
          pk[i >> 1] = pack_bf16(d0, d1);
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          *reinterpret_cast<uint4*>(smem + DqSmem::DS_OFF + sw128_offset(row, (cc >> 3) + q4)) = make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready);
    }
    mbar_wait(dq_done, (n_tiles - 1) & 1);
    tc_fence_after();
    // the dS tile is free once the last dQ MMA retired: stage dQ through this warp's 4 KB slice of it
    store_rows64(t_dq + lane_addr, smem + DqSmem::DS_OFF + grp * (32 * 128), p.dQ + ((long long)b * p.Lq + q0 + grp * 32) * p.lddq + h * HD, p.lddq,
                 p.Lq - (q0 + grp * 32), lane, 1.f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

struct DkvSmem {
  static constexpr int K_OFF = 0;                          // [128 keys x 64]
  static constexpr int V_OFF = K_OFF + TK * HD * 2;        // [128 keys x 64]
  static constexpr int Q_OFF = V_OFF + TK * HD * 2;        // 2 x [64 queries x 64]
  static constexpr int DO_OFF = Q_OFF + 2 * BT * HD * 2;   // 2 x [64 queries x 64]
  static constexpr int P_OFF = DO_OFF + 2 * BT * HD * 2;   // P^T  [128 keys x 64 queries] bf16
  static constexpr int DS_OFF = P_OFF + TK * BT * 2;       // dS^T [128 keys x 64 queries] bf16
  static constexpr int STAT_OFF = DS_OFF + TK * BT * 2;    // 2 x {lse2[64], delta[64]} fp32
  static constexpr int BAR_OFF = STAT_OFF + 2 * 2 * BT * 4;
  static constexpr int TOTAL = BAR_OFF + 128;
};

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                       const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DkvSmem::BAR_OFF);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;     // [2]
  uint64_t* q_empty = bars + 3;    // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* ds_ready = bars + 6;
  uint64_t* acc_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* stats = reinterpret_cast<float*>(smem + DkvSmem::STAT_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * TK, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.Lq + BT - 1) / BT;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(ds_ready, 4);
    mbar_init(acc_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base, t_dp = tmem_base + 64, t_dv = tmem_base + 128, t_dk = tmem_base + 192;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * TK * HD * 2);
      tma_load_4d(smem + DkvSmem::K_OFF, &tm_k, kv_full, 0, h, k0, b);
      tma_load_4d(smem + DkvSmem::V_OFF, &tm_v, kv_full, 0, h, k0, b);
      for (int i = 0; i < n_tiles; ++i) {
        const int st = i & 1;
        mbar_wait(&q_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], 2 * BT * HD * 2);
        tma_load_4d(smem + DkvSmem::Q_OFF + st * BT * HD * 2, &tm_q, &q_full[st], 0, h, i * BT, b);
        tma_load_4d(smem + DkvSmem::DO_OFF + st * BT * HD * 2, &tm_do, &q_full[st], 0, h, i * BT, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(TK, BT, 0, 0);     // S^T = K Q^T, dP^T = V dO^T   (N = 64 queries)
      constexpr uint32_t idesc_acc = umma_idesc_bf16(TK, HD, 0, 1);   // dV += P^T dO, dK += dS^T Q   (B MN-major, N = d)
      const uint32_t sk = smem_u32(smem + DkvSmem::K_OFF), sv = smem_u32(smem + DkvSmem::V_OFF);
      const uint32_t sp = smem_u32(smem + DkvSmem::P_OFF), sds = smem_u32(smem + DkvSmem::DS_OFF);
      mbar_wait(kv_full, 0);
      for (int i = 0; i < n_tiles; ++i) {
        const int st = i & 1;
        const uint32_t sq = smem_u32(smem + DkvSmem::Q_OFF + st * BT * HD * 2);
        const uint32_t sdo = smem_u32(smem + DkvSmem::DO_OFF + st * BT * HD * 2);
        mbar_wait(&q_full[st], (i >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss(t_s, umma_desc(sk + k * 32, 16, 1024), umma_desc(sq + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss(t_dp, umma_desc(sv + k * 32, 16, 1024), umma_desc(sdo + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(sdp_full);
        mbar_wait(ds_ready, i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BT / 16; ++k)
          umma_ss(t_dv, umma_desc(sp + k * 32, 16, 1024), umma_desc(sdo + k * 2048, BT * 128, 1024), idesc_acc, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < BT / 16; ++k)
          umma_ss(t_dk, umma_desc(sds + k * 32, 16, 1024), umma_desc(sq + k * 2048, BT * 128, 1024), idesc_acc, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&q_empty[st]);
        umma_commit(acc_done);
      }
    }
    __syncwarp();
  } else {
    const int grp = warp & 3;
    const int row = grp * 32 + lane;                 // key row of the tile == TMEM lane
    const int t128 = threadIdx.x - 64;               // 0..127 among the softmax threads
    const uint32_t lane_addr = uint32_t(grp * 32) << 16;
    const float c = p.scale * LOG2E;
    const uint64_t c2 = f2_pack(c, c), sc2 = f2_pack(p.scale, p.scale);
    const long long stat_base = ((long long)b * p.H + h) * p.Lq;
    for (int i = 0; i < n_tiles; ++i) {
      float* sl = stats + (i & 1) * 2 * BT;          // [lse2[64] | delta[64]] for this tile's queries
      {
        const int qi = i * BT + (t128 & 63);
        float v = 0.f;                               // stored negated (and delta pre-scaled) so the inner loop is pure FFMA2
        if (qi < p.Lq) v = t128 < 64 ? -p.LSE[stat_base + qi] * LOG2E : -p.delta[stat_base + qi] * p.scale;
        sl[t128] = v;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(sdp_full, i & 1);                    // implies the previous tile's dV/dK MMAs (reading P^T/dS^T) retired
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < BT; cc += 32) {
        uint32_t rs[32], rp[32];
        tmem_ld32(t_s + lane_addr + cc, rs);
        tmem_ld32(t_dp + lane_addr + cc, rp);
        tmem_ld_wait();
        uint32_t pp[16], pd[16];
        const uint64_t* nl2 = reinterpret_cast<const uint64_t*>(sl + cc);         // -lse2 of queries cc.., pairs (warp-broadcast reads)
        const uint64_t* nd2 = reinterpret_cast<const uint64_t*>(sl + BT + cc);    // -delta*scale
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float x0, x1, d0, d1;
          f2_unpack(f2_fma(f2_pack_u(rs[e], rs[e + 1]), c2, nl2[e >> 1]), x0, x1);
          const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
          const uint64_t t2 = f2_fma(f2_pack_u(rp[e], rp[e + 1]), sc2, nd2[e >> 1]);
          f2_unpack(f2_mul(f2_pack(p0, p1), t2), d0, d1);
          pp[e >> 1] = pack_bf16(p0, p1);
          pd[e >> 1] = pack_bf16(d0, d1);
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t off = sw128_offset(row, (cc >> 3) + q4);
          *reinterpret_cast<uint4*>(smem + DkvSmem::P_OFF + off) = make_uint4(pp[q4 * 4], pp[q4 * 4 + 1], pp[q4 * 4 + 2], pp[q4 * 4 + 3]);
          *reinterpret_cast<uint4*>(smem + DkvSmem::DS_OFF + off) = make_uint4(pd[q4 * 4], pd[q4 * 4 + 1], pd[q4 * 4 + 2], pd[q4 * 4 + 3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready);
    }
    mbar_wait(acc_done, (n_tiles - 1) & 1);
    tc_fence_after();
    // P^T / dS^T tiles are free once the last accumulation MMA retired: stage dV / dK through this warp's slices of them
    store_rows64(t_dv + lane_addr, smem + DkvSmem::P_OFF + grp * (32 * 128), p.dV + ((long long)b * p.Lk + k0 + grp * 32) * p.lddv + h * HD, p.lddv,
                 p.Lk - (k0 + grp * 32), lane, 1.f);
    store_rows64(t_dk + lane_addr, smem + DkvSmem::DS_OFF + grp * (32 * 128), p.dK + ((long long)b * p.Lk + k0 + grp * 32) * p.lddk + h * HD, p.lddk,
                 p.Lk - (k0 + grp * 32), lane, 1.f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}


// ------------------------------------------------------------------------------------------------
// backward v2 (default): persistent kernels, one CTA per SM, TWO 128-row tiles in flight -- the forward v2 schedule.
// While warpgroup t turns S_t / dP_t into dS_t (MUFU + FMA bound), the tensor pipe runs the other tile's MMAs, and the
// next S / dP of a tile is issued right behind the accumulation MMA that consumed the previous dS; K/V (dQ kernel) or
// Q/dO (dKdV kernel) tiles are shared by both tiles through a 3-deep TMA ring, and the next work item's resident
// operands are fetched while the current item's last tile is still in the softmax warps.
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 warpgroup 0, warps 6-9 warpgroup 1.
// ------------------------------------------------------------------------------------------------
static constexpr int B2_THREADS = 384;     // warps 0-3 warpgroup 0, 4-7 warpgroup 1, 8 TMA, 9 MMA, 10-11 idle
static constexpr int B2_ST = 3;
static constexpr int B2_SOFTMAX_REGS = 200, B2_IO_REGS = 104;   // 256 x 200 + 128 x 104 = 384 x 168: the MMA issuer must not spill

struct Dq2Smem {
  static constexpr int Q_OFF = 0;                              // 2 x [128 x 64]
  static constexpr int DO_OFF = Q_OFF + 2 * TQ * HD * 2;       // 2 x [128 x 64]
  static constexpr int K_OFF = DO_OFF + 2 * TQ * HD * 2;       // B2_ST x [64 keys x 64]
  static constexpr int V_OFF = K_OFF + B2_ST * BT * HD * 2;    // B2_ST x [64 keys x 64]
  static constexpr int DS_OFF = V_OFF + B2_ST * BT * HD * 2;   // 2 x [128 x 64 keys] bf16, K-major SW128
  static constexpr int O_OFF = DS_OFF + 2 * TQ * BT * 2;       // 2 x [128 x 64]: the forward output, only read by the softmax warps (delta)
  static constexpr int BAR_OFF = O_OFF + 2 * TQ * HD * 2;
  static constexpr int TOTAL = BAR_OFF + 256;
};

// work item = (batch, head, 256-query block); TMEM per tile t: S_t [192t, +64) dP_t [+64, +128) dQ_t [+128, +192)
template <bool DROP>
__global__ void __launch_bounds__(B2_THREADS, 1)
attn_bwd_dq_tc2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                       const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                       const __grid_constant__ CUtensorMap tm_o, const AttnTcParams p,
                       const int n_qblk, const int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Dq2Smem::BAR_OFF);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;               // [B2_ST]
  uint64_t* kv_empty = kv_full + B2_ST;       // [B2_ST]
  uint64_t* sdp_full = kv_empty + B2_ST;      // [2]  S_t(j), dP_t(j) complete
  uint64_t* sdp_free = sdp_full + 2;          // [2]  S_t(j), dP_t(j) are in the softmax warps' registers: the MMA warp may overwrite them
  uint64_t* ds_ready = sdp_free + 2;          // [2]  dS_t(j) in smem
  uint64_t* ds_free = ds_ready + 2;           // [2]  dQ_t(j) MMAs retired: the dS_t tile may be rewritten; the last one of an item = dQ_t complete
  uint64_t* dq_empty = ds_free + 2;           // [2]  epilogue has read dQ_t
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.Lk + BT - 1) / BT;

  if (warp == 8 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1 + 8);      // the MMA warp's commit + the eight softmax warps (they read dO / O from the same buffers for delta)
    for (int i = 0; i < B2_ST; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&sdp_full[t], 1); mbar_init(&sdp_free[t], 4); mbar_init(&ds_ready[t], 4); mbar_init(&ds_free[t], 1); mbar_init(&dq_empty[t], 4);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
  setmaxnreg_dec<B2_IO_REGS>();
  if (warp == 8) {
    if (lane == 0) {
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int qb = item % n_qblk, h = (item / n_qblk) % p.H, b = item / (n_qblk * p.H);
        const int q0 = qb * 2 * TQ;
        const bool two = q0 + TQ < p.Lq;
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_expect_tx(q_full, (two ? 6 : 3) * TQ * HD * 2);
        tma_load_4d(smem + Dq2Smem::Q_OFF, &tm_q, q_full, 0, h, q0, b);
        tma_load_4d(smem + Dq2Smem::DO_OFF, &tm_do, q_full, 0, h, q0, b);
        tma_load_4d(smem + Dq2Smem::O_OFF, &tm_o, q_full, 0, h, q0, b);
        if (two) {
          tma_load_4d(smem + Dq2Smem::Q_OFF + TQ * HD * 2, &tm_q, q_full, 0, h, q0 + TQ, b);
          tma_load_4d(smem + Dq2Smem::DO_OFF + TQ * HD * 2, &tm_do, q_full, 0, h, q0 + TQ, b);
          tma_load_4d(smem + Dq2Smem::O_OFF + TQ * HD * 2, &tm_o, q_full, 0, h, q0 + TQ, b);
        }
        for (int j = 0; j < n_tiles; ++j, ++g) {
          const int st = g % B2_ST;
          mbar_wait(&kv_empty[st], ((g / B2_ST) & 1) ^ 1);
          mbar_expect_tx(&kv_full[st], 2 * BT * HD * 2);
          tma_load_4d(smem + Dq2Smem::K_OFF + st * BT * HD * 2, &tm_k, &kv_full[st], 0, h, j * BT, b);
          tma_load_4d(smem + Dq2Smem::V_OFF + st * BT * HD * 2, &tm_v, &kv_full[st], 0, h, j * BT, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    {
      const bool leader = elect_one();   // warp-uniform loops; only the MMA / commit instructions are predicated
      constexpr uint32_t idesc_s = umma_idesc_bf16(TQ, BT, 0, 0);     // S = Q K^T, dP = dO V^T   (N = 64 keys)
      constexpr uint32_t idesc_dq = umma_idesc_bf16(TQ, HD, 0, 1);    // dQ += dS K             (K MN-major, N = d)
      uint32_t g = 0, it = 0, ds_cnt[2] = {0, 0}, dq_cnt[2] = {0, 0}, sdp_cnt[2] = {0, 0};
      // S_t / dP_t of key tile j+1 are issued as soon as tile j sits in the softmax warps' registers (sdp_free) and run under
      // their exp work; dQ_t(j) follows when dS_t(j) is in smem.
      // low descriptor words (umma_lo): K-major operands step 2 units (32 B) per K=16; K as the MN-major B operand of the dQ
      // MMA steps 16 key rows = 2048 B = 128 units
      const uint32_t q_lo = umma_lo(smem_u32(smem + Dq2Smem::Q_OFF), 16), do_lo = umma_lo(smem_u32(smem + Dq2Smem::DO_OFF), 16);
      const uint32_t k_lo = umma_lo(smem_u32(smem + Dq2Smem::K_OFF), 16), v_lo = umma_lo(smem_u32(smem + Dq2Smem::V_OFF), 16);
      const uint32_t kmn_lo = umma_lo(smem_u32(smem + Dq2Smem::K_OFF), BT * 128), ds_lo = umma_lo(smem_u32(smem + Dq2Smem::DS_OFF), 16);
      constexpr uint32_t QT16 = TQ * HD * 2 / 16, KT16 = BT * HD * 2 / 16, DST16 = TQ * BT * 2 / 16;
      auto issue_sdp = [&](int t, int st) {
        mbar_wait(&sdp_free[t], (sdp_cnt[t] & 1) ^ 1);   // first use passes immediately
        ++sdp_cnt[t];
        tc_fence_after();
        const uint32_t t_s = tmem_base + t * 192;
        if (leader) umma_chain<HD / 16>(t_s, q_lo + t * QT16, 2, k_lo + st * KT16, 2, idesc_s, 0);
        if (leader) umma_chain<HD / 16>(t_s + 64, do_lo + t * QT16, 2, v_lo + st * KT16, 2, idesc_s, 0);
        if (leader) umma_commit(&sdp_full[t]);
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int qb = item % n_qblk;
        const int nt = (qb * 2 * TQ + TQ < p.Lq) ? 2 : 1;
        mbar_wait(q_full, it & 1);
        {
          const int st = g % B2_ST;
          mbar_wait(&kv_full[st], (g / B2_ST) & 1);
          tc_fence_after();
          for (int t = 0; t < nt; ++t) issue_sdp(t, st);
          if (n_tiles == 1) if (leader) umma_commit(q_empty);
        }
        for (int j = 0; j < n_tiles; ++j) {
          const int st = (g + j) % B2_ST;
          if (j + 1 < n_tiles) {
            const int stn = (g + j + 1) % B2_ST;
            mbar_wait(&kv_full[stn], ((g + j + 1) / B2_ST) & 1);
            tc_fence_after();
            for (int t = 0; t < nt; ++t) issue_sdp(t, stn);
            if (j + 2 == n_tiles) if (leader) umma_commit(q_empty);
          }
          for (int t = 0; t < nt; ++t) {
            mbar_wait(&ds_ready[t], ds_cnt[t] & 1);
            ++ds_cnt[t];
            if (j == 0) { mbar_wait(&dq_empty[t], (dq_cnt[t] & 1) ^ 1); ++dq_cnt[t]; }   // the previous item's epilogue has drained dQ_t
            tc_fence_after();
            if (leader) umma_chain<BT / 16>(tmem_base + t * 192 + 128, ds_lo + t * DST16, 2, kmn_lo + st * KT16, 128, idesc_dq, j > 0);
            if (leader) umma_commit(&ds_free[t]);
          }
          if (leader) umma_commit(&kv_empty[st]);
        }
        g += n_tiles;
      }
    }
    __syncwarp();
  }
  } else {
    setmaxnreg_inc<B2_SOFTMAX_REGS>();
    const int t = warp >> 2;
    const int grp = warp & 3;
    const int row = grp * 32 + lane;
    const uint32_t lane_addr = uint32_t(grp * 32) << 16;
    const uint32_t t_s = tmem_base + t * 192, t_dp = t_s + 64, t_dq = t_s + 128;
    uint8_t* ds_tile = smem + Dq2Smem::DS_OFF + t * TQ * BT * 2;
    const float c = p.scale * LOG2E;
    const uint64_t c2 = f2_pack(c, c), sc2 = f2_pack(p.scale, p.scale);
    uint32_t sf_cnt = 0, dsf_cnt = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int qb = item % n_qblk, h = (item / n_qblk) % p.H, b = item / (n_qblk * p.H);
      const int q0 = qb * 2 * TQ + t * TQ;
      mbar_wait(q_full, it & 1);                     // dO / O of this item are in smem (also paces a warpgroup whose tile does not exist)
      if (q0 >= p.Lq) {                              // this warpgroup's tile does not exist (the MMA warp skips it too)
        __syncwarp();
        if (lane == 0) mbar_arrive(q_empty);
        continue;
      }
      const int qrow = q0 + row;
      const long long ri = ((long long)b * p.H + h) * p.Lq + qrow;
      const float nlse2 = -(qrow < p.Lq ? p.LSE[ri] : 0.f) * LOG2E;
      // delta = rowsum(dO o O) from the TMA-loaded tiles (rows past Lq are zero-filled): no separate pass over dO and O
      float dl = 0.f;
      {
        const uint8_t* so = smem + Dq2Smem::O_OFF + t * TQ * HD * 2;
        const uint8_t* sg = smem + Dq2Smem::DO_OFF + t * TQ * HD * 2;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          Vec16<bf16> vo, vg;
          vo.raw = *reinterpret_cast<const uint4*>(so + sw128_offset(row, ch));
          vg.raw = *reinterpret_cast<const uint4*>(sg + sw128_offset(row, ch));
          float fo[8], fg[8]; vo.unpack(fo); vg.unpack(fg);
#pragma unroll
          for (int e = 0; e < 8; ++e) dl = fmaf(fo[e], fg[e], dl);
        }
      }
      fence_proxy_async_smem();                      // generic reads of the TMA-loaded dO / O tiles before the TMA refill (WAR across proxies)
      __syncwarp();
      if (lane == 0) mbar_arrive(q_empty);           // this warp is done with the item's dO / O tiles
      if (qrow < p.Lq) p.delta[ri] = dl;             // the dKdV kernel (next launch on the stream) reads it
      const float ndl = -dl * p.scale;
      const uint64_t nlse22 = f2_pack(nlse2, nlse2), ndl2 = f2_pack(ndl, ndl);
      // dropout: dP = keep/(1-p) * (dO V^T), so the mask and the scale ride on the multiplier of dP
      const uint32_t rk = DROP ? drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, uint32_t((b * p.H + h) * p.Lq + qrow)) : 0u;
      const float sck = p.scale * p.inv_keep;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&sdp_full[t], sf_cnt & 1);
        ++sf_cnt;
        tc_fence_after();
        uint32_t rs[2][32], rp[2][32];               // the whole 64-key row of S and dP: four loads in flight, one wait
        tmem_ld32(t_s + lane_addr, rs[0]);
        tmem_ld32(t_dp + lane_addr, rp[0]);
        tmem_ld32(t_s + lane_addr + 32, rs[1]);
        tmem_ld32(t_dp + lane_addr + 32, rp[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sdp_free[t]);      // the next S_t / dP_t may be computed under this tile's exp work
        const uint32_t rkh = rk ^ (uint32_t(j) * kDropColMulHi);
        uint32_t pk[2][16];                            // all of dS_t(j) in registers first: the wait for the smem tile comes after the math
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {     // dS = P o (dP - delta) * scale, two elements per FFMA2 / FMUL2
            float x0, x1, d0, d1;
            f2_unpack(f2_fma(f2_pack_u(rs[hf][i], rs[hf][i + 1]), c2, nlse22), x0, x1);
            uint64_t m2 = sc2;
            if (DROP) {                            // key = j*64 + hf*32 + i: 64-key group j (hoisted xor), offset hf*32 + i (immediate)
              m2 = f2_pack(drop_keep_c(rkh, uint32_t(hf * 32 + i) * kDropColMul, p.drop_thr) ? sck : 0.f,
                           drop_keep_c(rkh, uint32_t(hf * 32 + i + 1) * kDropColMul, p.drop_thr) ? sck : 0.f);
            }
            const uint64_t t2 = f2_fma(f2_pack_u(rp[hf][i], rp[hf][i + 1]), m2, ndl2);
            f2_unpack(f2_mul(f2_pack(ex2_approx(x0), ex2_approx(x1)), t2), d0, d1);
            pk[hf][i >> 1] = pack_bf16(d0, d1);
          }
        }
        if (j > 0) {                                   // dQ_t(j-1) must have retired before the dS_t tile is rewritten
          mbar_wait(&ds_free[t], dsf_cnt & 1);
          ++dsf_cnt;
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            *reinterpret_cast<uint4*>(ds_tile + sw128_offset(row, hf * 4 + q4)) =
                make_uint4(pk[hf][q4 * 4], pk[hf][q4 * 4 + 1], pk[hf][q4 * 4 + 2], pk[hf][q4 * 4 + 3]);
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ds_ready[t]);
      }
      mbar_wait(&ds_free[t], dsf_cnt & 1);           // the item's last dQ_t MMA
      ++dsf_cnt;
      tc_fence_after();
      // the dS_t tile is free once the last dQ_t MMA retired: stage dQ through this warp's 4 KB slice of it
      store_rows64(t_dq + lane_addr, ds_tile + grp * (32 * 128), p.dQ + ((long long)b * p.Lq + q0 + grp * 32) * p.lddq + h * HD, p.lddq,
                   p.Lq - (q0 + grp * 32), lane, 1.f, p.dbq ? p.dbq + h * HD : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dq_empty[t]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

struct Dkv2Smem {
  static constexpr int K_OFF = 0;                              // 2 x [128 keys x 64]
  static constexpr int V_OFF = K_OFF + 2 * TK * HD * 2;        // 2 x [128 keys x 64]
  static constexpr int Q_OFF = V_OFF + 2 * TK * HD * 2;        // B2_ST x [64 queries x 64]
  static constexpr int DO_OFF = Q_OFF + B2_ST * BT * HD * 2;   // B2_ST x [64 queries x 64]
  static constexpr int P_OFF = DO_OFF + B2_ST * BT * HD * 2;   // 2 x P^T  [128 keys x 64 queries] bf16
  static constexpr int DS_OFF = P_OFF + 2 * TK * BT * 2;       // 2 x dS^T [128 keys x 64 queries] bf16
  static constexpr int STAT_OFF = DS_OFF + 2 * TK * BT * 2;    // 2 warpgroups x 2 buffers x {lse2[64], delta[64], dropout row key[64]}
  static constexpr int BAR_OFF = STAT_OFF + 2 * 2 * 3 * BT * 4;
  static constexpr int TOTAL = BAR_OFF + 256;
};

// work item = (batch, head, 256-key block); TMEM per tile t: S^T_t [256t, +64) dP^T_t [+64, +128) dV_t [+128, +192) dK_t [+192, +256)
// DS_OUT: every warp also sends its 32-key slab of each dS^T tile to the scratch tensor [B, H, Lk, Lq] (tm_ds) with one bulk tensor
// store from the shared-memory operand tile -- the dQ kernel below then needs no softmax work at all.
template <bool DROP, bool DS_OUT>
__global__ void __launch_bounds__(B2_THREADS, 1)
attn_bwd_dkv_tc2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                        const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                        const __grid_constant__ CUtensorMap tm_ds, const AttnTcParams p,
                        const int n_kblk, const int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Dkv2Smem::BAR_OFF);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  uint64_t* q_full = bars + 2;                // [B2_ST]
  uint64_t* q_empty = q_full + B2_ST;         // [B2_ST]
  uint64_t* sdp_full = q_empty + B2_ST;       // [2]  S^T_t(i), dP^T_t(i) complete
  uint64_t* sdp_free = sdp_full + 2;          // [2]  ... and in the softmax warps' registers: may be overwritten
  uint64_t* ds_ready = sdp_free + 2;          // [2]  P^T_t(i), dS^T_t(i) in smem
  uint64_t* ds_free = ds_ready + 2;           // [2]  dV_t / dK_t MMAs of tile i retired (the last one of an item = accumulators complete)
  uint64_t* acc_empty = ds_free + 2;          // [2]  epilogue has read dV_t / dK_t
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.Lq + BT - 1) / BT;

  if (warp == 8 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    for (int i = 0; i < B2_ST; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&sdp_full[t], 1); mbar_init(&sdp_free[t], 4); mbar_init(&ds_ready[t], 4); mbar_init(&ds_free[t], 1); mbar_init(&acc_empty[t], 4);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
  setmaxnreg_dec<B2_IO_REGS>();
  if (warp == 8) {
    if (lane == 0) {
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int kb = item % n_kblk, h = (item / n_kblk) % p.H, b = item / (n_kblk * p.H);
        const int k0 = kb * 2 * TK;
        const bool two = k0 + TK < p.Lk;
        mbar_wait(kv_empty, (it & 1) ^ 1);
        mbar_expect_tx(kv_full, (two ? 4 : 2) * TK * HD * 2);
        tma_load_4d(smem + Dkv2Smem::K_OFF, &tm_k, kv_full, 0, h, k0, b);
        tma_load_4d(smem + Dkv2Smem::V_OFF, &tm_v, kv_full, 0, h, k0, b);
        if (two) {
          tma_load_4d(smem + Dkv2Smem::K_OFF + TK * HD * 2, &tm_k, kv_full, 0, h, k0 + TK, b);
          tma_load_4d(smem + Dkv2Smem::V_OFF + TK * HD * 2, &tm_v, kv_full, 0, h, k0 + TK, b);
        }
        for (int i = 0; i < n_tiles; ++i, ++g) {
          const int st = g % B2_ST;
          mbar_wait(&q_empty[st], ((g / B2_ST) & 1) ^ 1);
          mbar_expect_tx(&q_full[st], 2 * BT * HD * 2);
          tma_load_4d(smem + Dkv2Smem::Q_OFF + st * BT * HD * 2, &tm_q, &q_full[st], 0, h, i * BT, b);
          tma_load_4d(smem + Dkv2Smem::DO_OFF + st * BT * HD * 2, &tm_do, &q_full[st], 0, h, i * BT, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    {
      const bool leader = elect_one();   // warp-uniform loops; only the MMA / commit instructions are predicated
      constexpr uint32_t idesc_s = umma_idesc_bf16(TK, BT, 0, 0);     // S^T = K Q^T, dP^T = V dO^T   (N = 64 queries)
      constexpr uint32_t idesc_acc = umma_idesc_bf16(TK, HD, 0, 1);   // dV += P^T dO, dK += dS^T Q   (B MN-major, N = d)
      uint32_t g = 0, it = 0, ds_cnt[2] = {0, 0}, acc_cnt[2] = {0, 0}, sdp_cnt[2] = {0, 0};
      // low descriptor words (umma_lo): K-major operands step 2 units (32 B) per K=16; Q / dO as MN-major B operands of the
      // accumulation MMAs step 16 query rows = 2048 B = 128 units
      const uint32_t k_lo = umma_lo(smem_u32(smem + Dkv2Smem::K_OFF), 16), v_lo = umma_lo(smem_u32(smem + Dkv2Smem::V_OFF), 16);
      const uint32_t q_lo = umma_lo(smem_u32(smem + Dkv2Smem::Q_OFF), 16), do_lo = umma_lo(smem_u32(smem + Dkv2Smem::DO_OFF), 16);
      const uint32_t qmn_lo = umma_lo(smem_u32(smem + Dkv2Smem::Q_OFF), BT * 128), domn_lo = umma_lo(smem_u32(smem + Dkv2Smem::DO_OFF), BT * 128);
      const uint32_t p_lo = umma_lo(smem_u32(smem + Dkv2Smem::P_OFF), 16), ds_lo = umma_lo(smem_u32(smem + Dkv2Smem::DS_OFF), 16);
      constexpr uint32_t KT16 = TK * HD * 2 / 16, QT16 = BT * HD * 2 / 16, PT16 = TK * BT * 2 / 16;
      auto issue_sdp = [&](int t, int st) {
        mbar_wait(&sdp_free[t], (sdp_cnt[t] & 1) ^ 1);   // first use passes immediately
        ++sdp_cnt[t];
        tc_fence_after();
        const uint32_t t_s = tmem_base + t * 256;
        if (leader) umma_chain<HD / 16>(t_s, k_lo + t * KT16, 2, q_lo + st * QT16, 2, idesc_s, 0);
        if (leader) umma_chain<HD / 16>(t_s + 64, v_lo + t * KT16, 2, do_lo + st * QT16, 2, idesc_s, 0);
        if (leader) umma_commit(&sdp_full[t]);
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int kb = item % n_kblk;
        const int nt = (kb * 2 * TK + TK < p.Lk) ? 2 : 1;
        mbar_wait(kv_full, it & 1);
        {
          const int st = g % B2_ST;
          mbar_wait(&q_full[st], (g / B2_ST) & 1);
          tc_fence_after();
          for (int t = 0; t < nt; ++t) issue_sdp(t, st);
          if (n_tiles == 1) if (leader) umma_commit(kv_empty);
        }
        for (int i = 0; i < n_tiles; ++i) {
          const int st = (g + i) % B2_ST;
          if (i + 1 < n_tiles) {                       // run ahead: next query tile's S^T / dP^T under this tile's exp work
            const int stn = (g + i + 1) % B2_ST;
            mbar_wait(&q_full[stn], ((g + i + 1) / B2_ST) & 1);
            tc_fence_after();
            for (int t = 0; t < nt; ++t) issue_sdp(t, stn);
            if (i + 2 == n_tiles) if (leader) umma_commit(kv_empty);
          }
          for (int t = 0; t < nt; ++t) {
            mbar_wait(&ds_ready[t], ds_cnt[t] & 1);
            ++ds_cnt[t];
            if (i == 0) { mbar_wait(&acc_empty[t], (acc_cnt[t] & 1) ^ 1); ++acc_cnt[t]; }
            tc_fence_after();
            const uint32_t t_dv = tmem_base + t * 256 + 128;
            if (leader) umma_chain<BT / 16>(t_dv, p_lo + t * PT16, 2, domn_lo + st * QT16, 128, idesc_acc, i > 0);
            if (leader) umma_chain<BT / 16>(t_dv + 64, ds_lo + t * PT16, 2, qmn_lo + st * QT16, 128, idesc_acc, i > 0);
            if (leader) umma_commit(&ds_free[t]);
          }
          if (leader) umma_commit(&q_empty[st]);
        }
        g += n_tiles;
      }
    }
    __syncwarp();
  }
  } else {
    setmaxnreg_inc<B2_SOFTMAX_REGS>();
    const int t = warp >> 2;
    const int grp = warp & 3;
    const int row = grp * 32 + lane;                 // key row of the tile == TMEM lane
    const int t128 = threadIdx.x - t * 128;          // 0..127 inside this warpgroup
    const uint32_t lane_addr = uint32_t(grp * 32) << 16;
    const uint32_t t_s = tmem_base + t * 256, t_dp = t_s + 64, t_dv = t_s + 128, t_dk = t_s + 192;
    uint8_t* p_tile = smem + Dkv2Smem::P_OFF + t * TK * BT * 2;
    uint8_t* ds_tile = smem + Dkv2Smem::DS_OFF + t * TK * BT * 2;
    float* stats = reinterpret_cast<float*>(smem + Dkv2Smem::STAT_OFF) + t * (2 * 3 * BT);
    const float c = p.scale * LOG2E;
    const uint64_t c2 = f2_pack(c, c), sc2 = f2_pack(p.scale, p.scale);
    uint32_t sf_cnt = 0, dsf_cnt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int kb = item % n_kblk, h = (item / n_kblk) % p.H, b = item / (n_kblk * p.H);
      const int k0 = kb * 2 * TK + t * TK;
      if (k0 >= p.Lk) continue;
      const long long stat_base = ((long long)b * p.H + h) * p.Lq;
      const uint32_t jc = drop_col_attn(uint32_t(k0 + row));      // dropout: this thread's key column
      const float sck = p.scale * p.inv_keep;
      // per-query statistics of a tile: thread t128 < 64 fetches LSE, the others delta; the NEXT tile's value is requested one
      // iteration ahead so the global-load latency hides under this tile's exp work
      auto load_stat = [&](int i) -> float {
        const int qi = i * BT + (t128 & 63);
        return qi < p.Lq ? (t128 < 64 ? p.LSE[stat_base + qi] : p.delta[stat_base + qi]) : 0.f;
      };
      float stat_next = load_stat(0);
      for (int i = 0; i < n_tiles; ++i) {
        float* sl = stats + (sf_cnt & 1) * 3 * BT;   // [lse2[64] | delta[64] | row key[64]] of this tile's queries; parity runs across items
        {
          const int qi = i * BT + (t128 & 63);
          // stored negated (and delta pre-scaled) so the inner loop is pure FFMA2
          sl[t128] = t128 < 64 ? -stat_next * LOG2E : -stat_next * p.scale;
          if (DROP && t128 < 64) reinterpret_cast<uint32_t*>(sl)[2 * BT + t128] = drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, uint32_t(stat_base + qi));
          if (i + 1 < n_tiles) stat_next = load_stat(i + 1);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");
        mbar_wait(&sdp_full[t], sf_cnt & 1);
        ++sf_cnt;
        tc_fence_after();
        uint32_t rs[2][32], rp[2][32];               // the whole 64-query row of S^T and dP^T: four loads in flight, one wait
        tmem_ld32(t_s + lane_addr, rs[0]);
        tmem_ld32(t_dp + lane_addr, rp[0]);
        tmem_ld32(t_s + lane_addr + 32, rs[1]);
        tmem_ld32(t_dp + lane_addr + 32, rp[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sdp_free[t]);      // the next S^T_t / dP^T_t may be computed under this tile's exp work
        uint32_t pp[2][16], pd[2][16];                 // P^T / dS^T of the tile in registers first: the wait for the smem tiles comes after the math
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const uint64_t* nl2 = reinterpret_cast<const uint64_t*>(sl + hf * 32);         // -lse2 of the queries, pairs (warp-broadcast reads)
          const uint64_t* nd2 = reinterpret_cast<const uint64_t*>(sl + BT + hf * 32);    // -delta*scale
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float x0, x1, d0, d1;
            f2_unpack(f2_fma(f2_pack_u(rs[hf][e], rs[hf][e + 1]), c2, nl2[e >> 1]), x0, x1);
            const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            uint64_t m2 = sc2;
            float pd0 = p0, pd1 = p1;                // what multiplies dO in dV: the dropped, rescaled probabilities
            if (DROP) {
              const uint32_t* rk = reinterpret_cast<const uint32_t*>(sl) + 2 * BT + hf * 32 + e;
              const bool k0_ = drop_keep_c(rk[0], jc, p.drop_thr), k1_ = drop_keep_c(rk[1], jc, p.drop_thr);
              m2 = f2_pack(k0_ ? sck : 0.f, k1_ ? sck : 0.f);
              // P o m2 = keep/(1-p) * P * scale in ONE packed multiply; the epilogue takes the `scale` out of dV again
              // (4 fewer instructions per pair than two selects and two multiplies on the probabilities)
              f2_unpack(f2_mul(f2_pack(p0, p1), m2), pd0, pd1);
            }
            const uint64_t t2 = f2_fma(f2_pack_u(rp[hf][e], rp[hf][e + 1]), m2, nd2[e >> 1]);
            f2_unpack(f2_mul(f2_pack(p0, p1), t2), d0, d1);
            pp[hf][e >> 1] = pack_bf16(pd0, pd1);
            pd[hf][e >> 1] = pack_bf16(d0, d1);
          }
        }
        if (i > 0) {                                   // dV_t / dK_t MMAs of the previous tile must have retired before P^T / dS^T are rewritten
          mbar_wait(&ds_free[t], dsf_cnt & 1);
          ++dsf_cnt;
        }
        if (DS_OUT) {                                  // ... and the bulk store of this warp's previous dS^T slab must be done reading it
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint32_t off = sw128_offset(row, hf * 4 + q4);
            *reinterpret_cast<uint4*>(p_tile + off) = make_uint4(pp[hf][q4 * 4], pp[hf][q4 * 4 + 1], pp[hf][q4 * 4 + 2], pp[hf][q4 * 4 + 3]);
            *reinterpret_cast<uint4*>(ds_tile + off) = make_uint4(pd[hf][q4 * 4], pd[hf][q4 * 4 + 1], pd[hf][q4 * 4 + 2], pd[hf][q4 * 4 + 3]);
          }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&ds_ready[t]);
          if (DS_OUT) {                                // [32 keys x 64 queries] of dS^T (scaled, dropout applied) -> scratch[b*H + h, query chunk i, k0 + 32 grp .., :]:
            tma_store_4d(&tm_ds, ds_tile + grp * (32 * 128), 0, k0 + grp * 32, i, b * p.H + h);   // 4 KB contiguous; keys past Lk are clipped
            tma_store_commit();
          }
        }
      }
      mbar_wait(&ds_free[t], dsf_cnt & 1);           // the item's last accumulation MMAs
      ++dsf_cnt;
      tc_fence_after();
      if (DS_OUT) {                                  // the staging below reuses the dS^T slab
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
      // P^T / dS^T tiles are free once the last accumulation MMA retired: stage dV / dK through this warp's slices of them
      store_rows64(t_dv + lane_addr, p_tile + grp * (32 * 128), p.dV + ((long long)b * p.Lk + k0 + grp * 32) * p.lddv + h * HD, p.lddv,
                   p.Lk - (k0 + grp * 32), lane, DROP ? 1.f / p.scale : 1.f, p.dbv ? p.dbv + h * HD : nullptr);
      store_rows64(t_dk + lane_addr, ds_tile + grp * (32 * 128), p.dK + ((long long)b * p.Lk + k0 + grp * 32) * p.lddk + h * HD, p.lddk,
                   p.Lk - (k0 + grp * 32), lane, 1.f, p.dbk ? p.dbk + h * HD : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[t]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// dQ from the stored score gradient:  dQ[b, q, h, :] = sum_k dS[b, h, q, k] K[b, k, h, :]   (dS already carries scale and dropout).
// A memory-bound batched GEMM: work item = (batch, head, 128-query tile); per 128-key step one TMA stage = the dS^T box
// [128 keys x 128 queries] (two 64-query SW128 chunks: the A operand, MN-major -- queries contiguous) + the K tile [128 keys x 64]
// (B operand, MN-major); 8 UMMAs (M = 128 queries, N = 64, K = 16 keys) accumulate into one of two TMEM accumulators; four epilogue
// warps drain the other one (bf16 rows + the in-proj bias gradient column sums).  4-stage ring = 192 KB in flight per SM.
// ------------------------------------------------------------------------------------------------
static constexpr int DQS_THREADS = 192;    // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (TMEM lane group = warp & 3)
static constexpr int DQS_ST = 4;
struct DqDsSmem {
  static constexpr int DS_OFF = 0;                                   // DQS_ST x 2 x [128 keys x 64 queries]
  static constexpr int K_OFF = DS_OFF + DQS_ST * 2 * TK * 64 * 2;    // DQS_ST x [128 keys x 64]
  static constexpr int STG_OFF = K_OFF + DQS_ST * TK * HD * 2;       // 4 warps x [32 x 128 B]
  static constexpr int BAR_OFF = STG_OFF + 4 * 32 * 128;
  static constexpr int TOTAL = BAR_OFF + 256;
};

__global__ void __launch_bounds__(DQS_THREADS, 1)
attn_bwd_dq_ds_kernel(const __grid_constant__ CUtensorMap tm_ds, const __grid_constant__ CUtensorMap tm_k, const AttnTcParams p,
                      const int n_qt, const int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DqDsSmem::BAR_OFF);
  uint64_t* full = bars;                      // [DQS_ST]
  uint64_t* empty = full + DQS_ST;            // [DQS_ST]
  uint64_t* acc_full = empty + DQS_ST;        // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ks = (p.Lk + TK - 1) / TK;
  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tm_ds); tma_prefetch_desc(&tm_k);
    for (int i = 0; i < DQS_ST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int qt = item % n_qt, h = (item / n_qt) % p.H, b = item / (n_qt * p.H);
        for (int ks = 0; ks < n_ks; ++ks, ++g) {
          const int st = g % DQS_ST;
          mbar_wait(&empty[st], ((g / DQS_ST) & 1) ^ 1);
          mbar_expect_tx(&full[st], 2 * TK * 64 * 2 + TK * HD * 2);
          uint8_t* ds = smem + DqDsSmem::DS_OFF + st * (2 * TK * 64 * 2);
          tma_load_4d(ds, &tm_ds, &full[st], 0, ks * TK, 2 * qt, b * p.H + h);               // 16 KB contiguous each; out-of-range keys / chunks arrive as zeros
          tma_load_4d(ds + TK * 64 * 2, &tm_ds, &full[st], 0, ks * TK, 2 * qt + 1, b * p.H + h);
          tma_load_4d(smem + DqDsSmem::K_OFF + st * (TK * HD * 2), &tm_k, &full[st], 0, h, ks * TK, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(TQ, HD, 1, 1);             // A = dS^T tile read MN-major, B = K tile read MN-major
    const uint32_t a_lo0 = umma_lo(smem_u32(smem + DqDsSmem::DS_OFF), TK * 128), b_lo0 = umma_lo(smem_u32(smem + DqDsSmem::K_OFF), TK * 128);
    constexpr uint32_t A16 = 2 * TK * 64 * 2 / 16, B16 = TK * HD * 2 / 16;
    uint32_t g = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int ks = 0; ks < n_ks; ++ks, ++g) {
        const int st = g % DQS_ST;
        mbar_wait(&full[st], (g / DQS_ST) & 1);
        tc_fence_after();
        if (leader) umma_chain<TK / 16>(tmem_base + acc * HD, a_lo0 + st * A16, 128, b_lo0 + st * B16, 128, idesc, ks > 0);   // 16 keys = 2048 B per step
        if (leader) umma_commit(&empty[st]);
      }
      if (leader) umma_commit(&acc_full[acc]);
    }
    __syncwarp();
  } else {
    const int grp = warp & 3;
    const uint32_t lane_addr = uint32_t(grp * 32) << 16;
    uint8_t* stage = smem + DqDsSmem::STG_OFF + (warp - 2) * (32 * 128);
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int qt = item % n_qt, h = (item / n_qt) % p.H, b = item / (n_qt * p.H);
      const int acc = it & 1;
      const int q0 = qt * TQ + grp * 32;
      mbar_wait(&acc_full[acc], (it >> 1) & 1);
      tc_fence_after();
      store_rows64(tmem_base + acc * HD + lane_addr, stage, p.dQ + ((long long)b * p.Lq + q0) * p.lddq + h * HD, p.lddq, p.Lq - q0, lane, 1.f,
                   p.dbq ? p.dbq + h * HD : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// scratch of the dS route: bf16 [B * H, ceil(Lq / 64), Lk, 64] -- per 64-query chunk a [Lk x 64] matrix with 128-byte rows, so the 32-key
// slabs the dK/dV kernel stores and the 128-key boxes the dQ kernel loads are contiguous (the first version, [B, H, Lk, Lq], made
// every box 128 separate 128-byte pieces 1 KB apart and the dQ kernel ran at half the memory bandwidth)
extern int g_attn_bwd_ds_route;
int64_t attn_bwd_ws_bytes(const b200f_attn_args& a) {
  if (!g_attn_bwd_ds_route || a.dtype != B200F_BF16 || a.D != HD || a.Lq < 128 || a.Lk < 128) return 0;
  return (int64_t)a.B * a.H * a.Lk * ((a.Lq + 63) / 64 * 64) * 2;
}
// b200f_debug_set(15, 1) turns the dS route on.  OFF by default: measured on the B = 4096 step (same box, alternating runs, profiles/
// r02_v_*): delta 0.17 + dQ-from-dS 0.95-1.1 ms replace the 1.83 ms recomputing dQ kernel per launch, but the 8.6 GB of extra HBM traffic
// per launch costs what the saved exp / hash / MMA work gains on this power-capped part (SM clock 1.51 vs 1.54 GHz) -- 277.1 vs 277.7 ms
// per step, i.e. nothing, for 4.3 GB more memory.  Kept, tested in both forms, for parts that are not power-bound.
int g_attn_bwd_ds_route = 0;

int g_attn_bwd_variant = 0;      // 0 = persistent two-tile kernels, 1 = one tile per CTA (debug / A-B; b200f_debug_set(5, v))

int attn_bwd_tc(const b200f_attn_args& a, cudaStream_t st) {
  int rc = attn_tc_check(a);
  if (rc) return rc;
  B200F_REQUIRE(a.lddo % 8 == 0 && a.lddq % 8 == 0 && a.lddk % 8 == 0 && a.lddv % 8 == 0, B200F_ERR_ALIGN, "attention(tcgen05): gradient leading dims must be multiples of 8");
  B200F_REQUIRE(aligned16(a.dO) && aligned16(a.dQ) && aligned16(a.dK) && aligned16(a.dV) && a.delta, B200F_ERR_ALIGN, "attention(tcgen05): gradient alignment");
  AttnTcParams p = {};
  p.B = a.B; p.H = a.H; p.Lq = a.Lq; p.Lk = a.Lk; p.scale = a.scale;
  p.LSE = a.LSE; p.delta = a.delta;
  p.dQ = static_cast<bf16*>(a.dQ); p.lddq = a.lddq;
  p.dK = static_cast<bf16*>(a.dK); p.lddk = a.lddk;
  p.dV = static_cast<bf16*>(a.dV); p.lddv = a.lddv;
  p.dbq = a.dbq; p.dbk = a.dbk; p.dbv = a.dbv;
  p.drop_thr = drop_threshold(a.dropout_p); p.drop_seed_lo = a.drop_seed_lo; p.drop_seed_hi = a.drop_seed_hi;
  p.inv_keep = 1.f / (1.f - a.dropout_p);
  const bool drop = p.drop_thr != 0;
  B200F_REQUIRE(!drop || g_attn_bwd_variant == 0, B200F_ERR_UNSUPPORTED, "attention(tcgen05): dropout needs the persistent kernels");
  const long long rows = (long long)a.B * a.H * a.Lq;
  if (g_attn_bwd_variant != 0) {                     // the persistent dQ kernel computes delta from its own dO / O tiles
    if ((rc = launch_delta(a, st))) return rc;
  }
  static PerDeviceOnce configured;
  if (configured.first()) {
    B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::TOTAL));
    B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvSmem::TOTAL));
    B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Dq2Smem::TOTAL));
    B200F_CHECK_CUDA(cudaFuncSetAttribute((attn_bwd_dkv_tc2_kernel<false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, Dkv2Smem::TOTAL));
    B200F_CHECK_CUDA(cudaFuncSetAttribute((attn_bwd_dkv_tc2_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, Dkv2Smem::TOTAL));
    B200F_CHECK_CUDA(cudaFuncSetAttribute((attn_bwd_dkv_tc2_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, Dkv2Smem::TOTAL));
    B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_ds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DqDsSmem::TOTAL));
    B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Dq2Smem::TOTAL));
    B200F_CHECK_CUDA(cudaFuncSetAttribute((attn_bwd_dkv_tc2_kernel<true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, Dkv2Smem::TOTAL));
  }
  CUtensorMap tq, tdo, tk, tv, to, tds;
  memset(&tds, 0, sizeof(tds));
  const int64_t ws_need = attn_bwd_ws_bytes(a);
  const bool ds_route = g_attn_bwd_variant == 0 && g_attn_bwd_ds_route && a.bwd_ws && ws_need > 0 && a.bwd_ws_bytes >= ws_need && aligned16(a.bwd_ws);
  if (ds_route) {
    // delta = rowsum(dO o O) first (the recomputing dQ kernel used to produce it), dK/dV + dS^T, then dQ = dS K
    if ((rc = launch_delta(a, st))) return rc;
    const uint64_t n_qc = uint64_t((a.Lq + 63) / 64);
    const uint64_t dims[4] = {64, uint64_t(a.Lk), n_qc, uint64_t(a.B) * a.H};
    const uint64_t strides[3] = {128, uint64_t(a.Lk) * 128, n_qc * a.Lk * 128};
    {  // dKdV + dS^T slabs [32 keys x 64 queries]
      const uint32_t box[4] = {64, 32, 1, 1};
      if ((rc = make_tmap_bf16(&tds, a.bwd_ws, 4, dims, strides, box))) return rc;
      if ((rc = make_head_tmap(&tq, a.Q, a.ldq, a.B, a.H, a.Lq, BT))) return rc;
      if ((rc = make_head_tmap(&tdo, a.dO, a.lddo, a.B, a.H, a.Lq, BT))) return rc;
      if ((rc = make_head_tmap(&tk, a.K, a.ldk, a.B, a.H, a.Lk, TK))) return rc;
      if ((rc = make_head_tmap(&tv, a.V, a.ldv, a.B, a.H, a.Lk, TK))) return rc;
      const int n_kblk = (a.Lk + 2 * TK - 1) / (2 * TK);
      const long long n_items = (long long)n_kblk * a.H * a.B;
      B200F_REQUIRE(n_items < (1ll << 31), B200F_ERR_SHAPE, "attention(tcgen05): too many work items");
      const int grid = int(n_items < num_sms() ? n_items : num_sms());
      if (drop) attn_bwd_dkv_tc2_kernel<true, true><<<grid, B2_THREADS, Dkv2Smem::TOTAL, st>>>(tq, tdo, tk, tv, tds, p, n_kblk, int(n_items));
      else attn_bwd_dkv_tc2_kernel<false, true><<<grid, B2_THREADS, Dkv2Smem::TOTAL, st>>>(tq, tdo, tk, tv, tds, p, n_kblk, int(n_items));
      if ((rc = check_launch("attn_bwd_dkv_tc2_kernel"))) return rc;
    }
    {  // dQ = dS K: dS^T boxes [128 keys x 64 queries], K boxes [128 keys x 64]
      const uint32_t box[4] = {64, uint32_t(TK), 1, 1};
      if ((rc = make_tmap_bf16(&tds, a.bwd_ws, 4, dims, strides, box))) return rc;
      if ((rc = make_head_tmap(&tk, a.K, a.ldk, a.B, a.H, a.Lk, TK))) return rc;
      const int n_qt = (a.Lq + TQ - 1) / TQ;
      const long long n_items = (long long)n_qt * a.H * a.B;
      B200F_REQUIRE(n_items < (1ll << 31), B200F_ERR_SHAPE, "attention(tcgen05): too many work items");
      const int grid = int(n_items < num_sms() ? n_items : num_sms());
      attn_bwd_dq_ds_kernel<<<grid, DQS_THREADS, DqDsSmem::TOTAL, st>>>(tds, tk, p, n_qt, int(n_items));
      if ((rc = check_launch("attn_bwd_dq_ds_kernel"))) return rc;
    }
    return B200F_OK;
  }
  if (g_attn_bwd_variant == 0) {
    B200F_REQUIRE(a.ldo % 8 == 0 && aligned16(a.O), B200F_ERR_ALIGN, "attention(tcgen05): O alignment");
    if ((rc = make_head_tmap(&to, a.O, a.ldo, a.B, a.H, a.Lq, TQ))) return rc;
    {  // dQ: 128-row Q/dO boxes, 64-row K/V boxes; item = (batch, head, 256-query block)
      if ((rc = make_head_tmap(&tq, a.Q, a.ldq, a.B, a.H, a.Lq, TQ))) return rc;
      if ((rc = make_head_tmap(&tdo, a.dO, a.lddo, a.B, a.H, a.Lq, TQ))) return rc;
      if ((rc = make_head_tmap(&tk, a.K, a.ldk, a.B, a.H, a.Lk, BT))) return rc;
      if ((rc = make_head_tmap(&tv, a.V, a.ldv, a.B, a.H, a.Lk, BT))) return rc;
      const int n_qblk = (a.Lq + 2 * TQ - 1) / (2 * TQ);
      const long long n_items = (long long)n_qblk * a.H * a.B;
      B200F_REQUIRE(n_items < (1ll << 31), B200F_ERR_SHAPE, "attention(tcgen05): too many work items");
      const int grid = int(n_items < num_sms() ? n_items : num_sms());
      if (drop) attn_bwd_dq_tc2_kernel<true><<<grid, B2_THREADS, Dq2Smem::TOTAL, st>>>(tq, tdo, tk, tv, to, p, n_qblk, int(n_items));
      else attn_bwd_dq_tc2_kernel<false><<<grid, B2_THREADS, Dq2Smem::TOTAL, st>>>(tq, tdo, tk, tv, to, p, n_qblk, int(n_items));
      if ((rc = check_launch("attn_bwd_dq_tc2_kernel"))) return rc;
    }
    {  // dKdV: 128-row K/V boxes, 64-row Q/dO boxes; item = (batch, head, 256-key block)
      if ((rc = make_head_tmap(&tq, a.Q, a.ldq, a.B, a.H, a.Lq, BT))) return rc;
      if ((rc = make_head_tmap(&tdo, a.dO, a.lddo, a.B, a.H, a.Lq, BT))) return rc;
      if ((rc = make_head_tmap(&tk, a.K, a.ldk, a.B, a.H, a.Lk, TK))) return rc;
      if ((rc = make_head_tmap(&tv, a.V, a.ldv, a.B, a.H, a.Lk, TK))) return rc;
      const int n_kblk = (a.Lk + 2 * TK - 1) / (2 * TK);
      const long long n_items = (long long)n_kblk * a.H * a.B;
      B200F_REQUIRE(n_items < (1ll << 31), B200F_ERR_SHAPE, "attention(tcgen05): too many work items");
      const int grid = int(n_items < num_sms() ? n_items : num_sms());
      if (drop) attn_bwd_dkv_tc2_kernel<true, false><<<grid, B2_THREADS, Dkv2Smem::TOTAL, st>>>(tq, tdo, tk, tv, tds, p, n_kblk, int(n_items));
      else attn_bwd_dkv_tc2_kernel<false, false><<<grid, B2_THREADS, Dkv2Smem::TOTAL, st>>>(tq, tdo, tk, tv, tds, p, n_kblk, int(n_items));
      if ((rc = check_launch("attn_bwd_dkv_tc2_kernel"))) return rc;
    }
    return B200F_OK;
  }
  {  // dQ kernel: 128-row Q/dO boxes, 64-row K/V boxes
    if ((rc = make_head_tmap(&tq, a.Q, a.ldq, a.B, a.H, a.Lq, TQ))) return rc;
    if ((rc = make_head_tmap(&tdo, a.dO, a.lddo, a.B, a.H, a.Lq, TQ))) return rc;
    if ((rc = make_head_tmap(&tk, a.K, a.ldk, a.B, a.H, a.Lk, BT))) return rc;
    if ((rc = make_head_tmap(&tv, a.V, a.ldv, a.B, a.H, a.Lk, BT))) return rc;
    dim3 grid((a.Lq + TQ - 1) / TQ, a.H, a.B);
    attn_bwd_dq_tc_kernel<<<grid, ATT_THREADS, DqSmem::TOTAL, st>>>(tq, tdo, tk, tv, p);
    if ((rc = check_launch("attn_bwd_dq_tc_kernel"))) return rc;
  }
  {  // dKdV kernel: 128-row K/V boxes, 64-row Q/dO boxes
    if ((rc = make_head_tmap(&tq, a.Q, a.ldq, a.B, a.H, a.Lq, BT))) return rc;
    if ((rc = make_head_tmap(&tdo, a.dO, a.lddo, a.B, a.H, a.Lq, BT))) return rc;
    if ((rc = make_head_tmap(&tk, a.K, a.ldk, a.B, a.H, a.Lk, TK))) return rc;
    if ((rc = make_head_tmap(&tv, a.V, a.ldv, a.B, a.H, a.Lk, TK))) return rc;
    dim3 grid((a.Lk + TK - 1) / TK, a.H, a.B);
    attn_bwd_dkv_tc_kernel<<<grid, ATT_THREADS, DkvSmem::TOTAL, st>>>(tq, tdo, tk, tv, p);
    if ((rc = check_launch("attn_bwd_dkv_tc_kernel"))) return rc;
  }
  return B200F_OK;
}

B200F_DEFINE_EPOCH_HOOK(attn_tc)

}  // namespace b200f
