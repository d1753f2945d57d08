// bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands by TMA).
//
//   C[m,n] = epi( alpha * sum_k A(m,k) * B(n,k) ),  fp32 accumulation
//
// One persistent CTA per SM, 10 warps, warp-specialised:
//   warp 0    TMA producer   : cp.async.bulk.tensor -> STAGES-deep ring of {A,B} smem tiles (SWIZZLE_128B)
//   warp 1    MMA issuer     : one lane issues tcgen05.mma (128 x BN x 16) into one of two TMEM accumulators
//   warps 2-9 epilogue       : tcgen05.ld -> bias / residual / ReLU / mask -> global, while the MMA warp
//                              already fills the other accumulator (mainloop/epilogue overlap); each TMEM lane
//                              group is served by two warps that split the tile's columns
// Both operands may be K-major (reduction dim contiguous) or MN-major, so forward (X W^T), input-gradient
// (dY W) and weight-gradient (dY^T X, fp32 atomics with split-K over tokens) all run without a transpose.
#include "common.cuh"
#include "ptx.cuh"

namespace b200f {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int EPI_STAGE_BYTES = 32 * 128;   // smem staging tile per epilogue warp
static constexpr int GEMM_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane group)

struct GemmTcParams {
  int M, N, K;
  int num_m, num_n, split_k, kb_total, kb_per_split;
  void* C;
  long long ldc;
  const float* bias;
  const bf16* residual;
  long long ldr;
  const bf16* mask;
  long long ldm;
  float alpha;
  int flags;
  // UMMA descriptor geometry (bytes), filled by the host so it can be probed without recompiling
  uint32_t a_lbo, a_sbo, a_kadv, b_lbo, b_sbo, b_kadv;
  int vec_ok;  // C / residual / mask rows are 16-byte aligned -> 128-bit epilogue accesses
  float* colsum;  // optional [N]: += column sums of the stored bf16 C (N % 64 == 0; staged bf16 epilogue only)
  uint32_t drop_thr, drop_seed_lo, drop_seed_hi;   // inverted dropout on the output (staged bf16 epilogue only); 0 = off
  float inv_keep;
  int tma_store;   // bf16 output without residual / mask: staged tiles leave through cp.async.bulk.tensor stores (tensor map of C)
  int dbg_colsum;  // debug (b200f_debug_set(13, v)): 1 = column sums computed but the REDs skipped, 2 = column sums skipped (timing experiments only)
  int late_aux;    // debug (b200f_debug_set(12, 1)): prefetch the aux block one column block ahead instead of all at tile start
  int epi_stride;  // byte distance of a warp's two epilogue staging tiles (EPI_STAGE_BYTES), or 0 when the launch has a single tile per warp
  // one-bit ReLU' mask (b200f_gemm_args::sign_bits*): [M, ldsb] words, word c of a row = columns 32c..32c+31, element e of the word at
  // bit 8*(e % 4) + e / 4 (the order the PRMT-based packing below produces)
  uint32_t* sbits_out;
  const uint32_t* sbits;
  long long ldsb;
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 256;                         // 2 x BN fp32: this / next tile's bias slice
  static constexpr int EPI_OFFSET = (BIAS_OFFSET + 2 * BN * 4 + 1023) / 1024 * 1024;   // 8 warps x 2 x 4 KB staging tiles, 1024-byte aligned (TMA SWIZZLE_128B)
  static constexpr int TOTAL = EPI_OFFSET + 8 * 2 * EPI_STAGE_BYTES;
};

// Drain columns [c_begin, c_end) of one accumulator row (this thread's TMEM lane) through the fused epilogue.
__device__ __forceinline__ void epilogue_columns(const GemmTcParams& p, uint32_t t_row, long long row, bool row_ok, int n_base, int c_begin,
                                                 int c_end, bool out_f32, bool accum, bool relu) {
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(t_row + c0, r);
    tmem_ld_wait();
    const int col0 = n_base + c0;
    if (row_ok && col0 < p.N) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = col0 + g * 8;
        if (col >= p.N) break;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]) * p.alpha;
        const bool full8 = (col + 8 <= p.N) && p.vec_ok;
        if (p.bias) {
          if (full8) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (col + j < p.N) v[j] += __ldg(p.bias + col + j);
          }
        }
        if (p.residual) {
          const bf16* rp = p.residual + row * p.ldr + col;
          if (full8) {
            Vec16<bf16> rv; rv.load(rp);
            float f[8]; rv.unpack(f);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += f[j];
          } else {
            for (int j = 0; j < 8 && col + j < p.N; ++j) v[j] += to_f32(rp[j]);
          }
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.mask) {
          const bf16* mp = p.mask + row * p.ldm + col;
          if (full8) {
            Vec16<bf16> mv; mv.load(mp);
            float f[8]; mv.unpack(f);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = f[j] > 0.f ? v[j] : 0.f;
          } else {
            for (int j = 0; j < 8 && col + j < p.N; ++j) v[j] = to_f32(mp[j]) > 0.f ? v[j] : 0.f;
          }
        }
        if (out_f32) {
          float* cp = reinterpret_cast<float*>(p.C) + row * p.ldc + col;
          if (accum) {
            for (int j = 0; j < 8 && col + j < p.N; ++j) atomicAdd(cp + j, v[j]);
          } else if (full8) {
            *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(cp + 4) = make_float4(v[4], v[5], v[6], v[7]);
          } else {
            for (int j = 0; j < 8 && col + j < p.N; ++j) cp[j] = v[j];
          }
        } else {
          bf16* cp = reinterpret_cast<bf16*>(p.C) + row * p.ldc + col;
          if (full8) {
            Vec16<bf16> ov; ov.pack(v); ov.store(cp);
          } else {
            for (int j = 0; j < 8 && col + j < p.N; ++j) cp[j] = __float2bfloat16_rn(v[j]);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Staged epilogues.  A thread owns one accumulator ROW (its TMEM lane), so direct 16-byte accesses from a warp touch 32
// different 128-byte lines per instruction.  Instead every warp moves column blocks through private 4 KB shared-memory
// tiles [32 rows x 128 B] (16-byte chunks XOR-swizzled by row, conflict-free row-wise and line-wise):
//   * the residual / ReLU-mask block is PREFETCHED with cp.async (4 full lines per instruction, no registers) one block
//     ahead -- the first one before the accumulator is even complete -- so its latency never sits on the critical path;
//   * the bias slice of the tile sits in shared memory (loaded once per tile by the epilogue warps, read as broadcasts);
//   * results are written row-wise into the tile and leave as full 128-byte lines (bf16 / fp32 stores, or vector REDs
//     for the fp32 split-K accumulation of weight gradients).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void red_add_v4(float* dst, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void red_add_v2(float* dst, float x, float y) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(x), "f"(y) : "memory");
}

// Prefetch the 32 x 64 aux (residual, else mask) block at (row0, col0) into `stage`; always commits one (maybe empty) group.
__device__ __forceinline__ void epi_issue_aux(const GemmTcParams& p, uint8_t* stage, long long row0, int col0, int lane) {
  const bf16* aux = p.residual ? p.residual : p.mask;
  if (aux && col0 + 64 <= p.N) {
    const long long ldaux = p.residual ? p.ldr : p.ldm;
    const int crow = lane >> 3, cchunk = lane & 7;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + crow;
      if (row0 + r < p.M) cp_async16(stage + sw128_offset(r, cchunk), aux + (row0 + r) * ldaux + col0 + cchunk * 8);
    }
  }
  cp_async_commit();
}

// bf16 output.  `stage`: this warp's two 4 KB tiles; `bias_s`: the tile's bias slice starting at this warp's first column;
// columns [col_base, col_base + NCOLS) of rows [row0, row0 + 32).  Block 0's aux prefetch was issued by the caller.
// EXTRAS = dropout and/or column sums requested: a separate instantiation, so the plain epilogue (which bounds the K = 512
// GEMMs) carries none of their instructions or registers (measured: the runtime-flag version cost those GEMMs 10-15 %).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__host__ __device__ constexpr int sbit_pos(int e) { return 8 * (e & 3) + (e >> 2); }   // element e (0..31) of a 32-column word

template <int NCOLS, bool EXTRAS>
__device__ __forceinline__ void epilogue_tile_bf16(const GemmTcParams& p, const CUtensorMap* tmc, uint8_t* stage, const float* bias_s, uint32_t t_row,
                                                   long long row0, int lane, int n_base, int c_begin, bool relu, uint2 sb0, uint2 sb1) {
  constexpr int NBLK = NCOLS / 64;
  const long long row = row0 + lane;
  const int crow = lane >> 3, cchunk = lane & 7;             // coalesced phase: 4 rows x 8 chunks per instruction
  const bool has_aux = p.residual || p.mask;
  const uint32_t rk = (EXTRAS && p.drop_thr) ? drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, uint32_t(row)) : 0u;
  // dropout without a residual: 1/(1-p) rides on alpha and on the bias slice (pre-scaled by epilogue_tile), relu commutes with the
  // positive scale -- the per-element work is then one select
  const bool fold_keep = EXTRAS && p.drop_thr && !p.residual;
  const float al = fold_keep ? p.alpha * p.inv_keep : p.alpha;
  const uint64_t al2 = f2_pack(al, al);
  const float keep_mul = fold_keep ? 1.f : p.inv_keep;
#pragma unroll 1
  for (int blk = 0; blk < NBLK; ++blk) {
    const int c0 = c_begin + blk * 64;
    const int col0 = n_base + c0;
    uint8_t* st = stage + (blk & 1) * p.epi_stride;
    if (has_aux && !p.late_aux) {
      // the aux (residual / ReLU-mask) blocks of ALL of this warp's column blocks were prefetched by epilogue_tile before it waited
      // for the accumulator (one cp.async group per block, NBLK <= 2 = the number of staging tiles): their global-memory latency
      // hides under the mainloop instead of under one block of epilogue work
      if (blk + 1 < NBLK) cp_async_wait<1>(); else cp_async_wait<0>();
    } else {
      if (p.tma_store) {                                      // the bulk store that last read this staging tile must be done reading it
        // ... and with an aux block, also the store that last read the OTHER tile, which the prefetch below is about to overwrite
        const bool all = p.epi_stride == 0 || (has_aux && blk + 1 < NBLK);
        if (lane == 0) { if (all) tma_store_wait_read(); else tma_store_wait_read1(); }
        __syncwarp();
      }
      if (has_aux) {                                          // b200f_debug_set(12, 1): the round-1 schedule, one block ahead (A/B)
        if (blk + 1 < NBLK) {
          epi_issue_aux(p, stage + ((blk + 1) & 1) * p.epi_stride, row0, col0 + 64, lane);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
      }
    }
    __syncwarp();
    if (col0 >= p.N) continue;                                // warp-uniform
    if (col0 + 64 > p.N) {                                    // ragged right edge: direct path
      epilogue_columns(p, t_row, row, row < p.M, n_base, c0, c0 + 64, false, false, relu);
      continue;
    }
    // Three phases so that shared-memory loads and stores never interleave (an LDS must stay behind an earlier STS that may
    // alias it): (1) aux cells and the accumulator into registers, (2) all the math, (3) eight STS; then eight LDS of full rows
    // followed by eight global stores.  ncu on the K = 512 shapes: the epilogue warps are busy all the time (short-scoreboard
    // stalls on shared-memory operations) while the tensor pipe is 56 % active -- the staging traffic competes with the TMA
    // writes and UMMA operand reads of the mainloop for the 128 B/clk of shared-memory bandwidth; open item for round 2.
    uint32_t r[2][32];
    tmem_ld32(t_row + c0, r[0]);
    tmem_ld32(t_row + c0 + 32, r[1]);
    const uint2 sbw = blk ? sb1 : sb0;                        // this block's two words of the one-bit mask (if any)
    uint4 aux[8];
    if (has_aux) {
#pragma unroll
      for (int q = 0; q < 8; ++q) aux[q] = *reinterpret_cast<const uint4*>(st + sw128_offset(lane, q));
    }
    tmem_ld_wait();
    uint4 outv[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v[8];
        {   // v = acc * alpha + bias as four FFMA2 (two accumulator columns per instruction)
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
          if (p.bias) {
            b0 = *reinterpret_cast<const float4*>(bias_s + blk * 64 + h * 32 + g * 8);
            b1 = *reinterpret_cast<const float4*>(bias_s + blk * 64 + h * 32 + g * 8 + 4);
          }
          const uint32_t* a = &r[h][g * 8];
          f2_unpack(f2_fma(f2_pack_u(a[0], a[1]), al2, f2_pack(b0.x, b0.y)), v[0], v[1]);
          f2_unpack(f2_fma(f2_pack_u(a[2], a[3]), al2, f2_pack(b0.z, b0.w)), v[2], v[3]);
          f2_unpack(f2_fma(f2_pack_u(a[4], a[5]), al2, f2_pack(b1.x, b1.y)), v[4], v[5]);
          f2_unpack(f2_fma(f2_pack_u(a[6], a[7]), al2, f2_pack(b1.z, b1.w)), v[6], v[7]);
        }
        float f[8];
        if (has_aux) { Vec16<bf16> av; av.raw = aux[h * 4 + g]; av.unpack(f); }
        if (p.residual) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += f[j];
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (EXTRAS && p.drop_thr) {                          // nn.Dropout behind Linear(+ReLU): mask from (seed, row, column)
          const uint32_t cc = uint32_t(col0 + h * 32 + g * 8) * kDropColMul;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = drop_keep_c(rk, cc + uint32_t(j) * kDropColMul, p.drop_thr) ? v[j] * keep_mul : 0.f;
        }
        if (p.mask) {
          if (p.residual) {                                  // both present: the mask comes straight from global
            if (row < p.M) { Vec16<bf16> mv; mv.load(p.mask + row * p.ldm + col0 + h * 32 + g * 8); mv.unpack(f); }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = f[j] > 0.f ? v[j] : 0.f;
        }
        if (EXTRAS && p.sbits) {                             // the same mask, one bit per element (loaded before the accumulator wait)
          const uint32_t w = h ? sbw.y : sbw.x;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (w & (1u << sbit_pos(g * 8 + j))) ? v[j] : 0.f;
        }
        Vec16<bf16> ov; ov.pack(v);
        outv[h * 4 + g] = ov.raw;
      }
    }
    if (EXTRAS && p.sbits_out) {
      // One bit per stored element: set <=> the bf16 value is > 0.  The outputs are ReLU'd (halves in [+0, +inf] -- max(-0, +0) = +0 --
      // or NaN), so half + 0x7FFF carries into its top bit exactly when the half is non-zero, without crossing into the other half;
      // PRMT in sign-replicate mode spreads the two top bits of two words over four bytes, one LOP3 drops them at this group's bit
      // position: 1 integer instruction per element.
      uint32_t wb[2] = {0u, 0u};
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t t0 = outv[q].x + 0x7FFF7FFFu, t1 = outv[q].y + 0x7FFF7FFFu, t2 = outv[q].z + 0x7FFF7FFFu, t3 = outv[q].w + 0x7FFF7FFFu;
        wb[q >> 2] |= prmt(t0, t1, 0xFDB9u) & (0x01010101u << (2 * (q & 3)));
        wb[q >> 2] |= prmt(t2, t3, 0xFDB9u) & (0x01010101u << (2 * (q & 3) + 1));
      }
      if (row < p.M) *reinterpret_cast<uint2*>(p.sbits_out + row * p.ldsb + (col0 >> 5)) = make_uint2(wb[0], wb[1]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(st + sw128_offset(lane, q)) = outv[q];   // the cells this lane consumed
    if (p.tma_store) {
      // the staged [32 x 128 B] tile (SWIZZLE_128B layout, 1024-byte aligned) leaves as ONE bulk tensor store issued by one lane:
      // no LDS / STG instructions and no address arithmetic in the epilogue warps; rows past M are clipped by the tensor map
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { tma_store_2d(tmc, st, col0, int(row0)); tma_store_commit(); }
    } else {
      __syncwarp();
      bf16* cbase = reinterpret_cast<bf16*>(p.C) + row0 * p.ldc + col0 + cchunk * 8;
#pragma unroll
      for (int it = 0; it < 8; ++it) outv[it] = *reinterpret_cast<const uint4*>(st + sw128_offset(it * 4 + crow, cchunk));
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + crow;
        if (row0 + rr < p.M) *reinterpret_cast<uint4*>(cbase + (long long)rr * p.ldc) = outv[it];
      }
    }
    if (EXTRAS && p.colsum && p.dbg_colsum != 2) {
      // bias gradient: column sums of the 32 x 64 block just staged (the rounded values that were stored).  A lane owns
      // columns 2*lane, 2*lane+1 (one conflict-free word per row); even lanes collect 4 columns and issue one vector RED.
      // (round 2: on the warp-level tensor path -- ONES x tile -- instead of a 32-row LDS loop per lane; ptx.cuh)
      const int nr = p.M - row0 < 32 ? int(p.M - row0) : 32;
      float cs[8][2];
      colsum32x64_hmma(st, nr, lane, cs);
      if (lane < 4 && (p.dbg_colsum != 1 || cs[0][0] == 12345.678f)) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) red_add_v2(p.colsum + col0 + nt * 8 + 2 * lane, cs[nt][0], cs[nt][1]);
      }
    }
    __syncwarp();
  }
}

// fp32 output (plain store or split-K accumulation with vector REDs), 32-column blocks through one 4 KB tile.
template <int NCOLS>
__device__ __forceinline__ void epilogue_tile_f32(const GemmTcParams& p, uint8_t* stage, const float* bias_s, uint32_t t_row, long long row0,
                                                  int lane, int n_base, int c_begin, bool accum, bool relu) {
  const long long row = row0 + lane;
  const int crow = lane >> 3, cchunk = lane & 7;
#pragma unroll 1
  for (int cb = 0; cb < NCOLS; cb += 32) {
    const int c0 = c_begin + cb;
    const int col0 = n_base + c0;
    if (col0 >= p.N) break;                                   // warp-uniform
    if (col0 + 32 > p.N || p.residual || p.mask) {            // ragged edge / rare fp32 + aux combination: direct path
      epilogue_columns(p, t_row, row, row < p.M, n_base, c0, c0 + 32, true, accum, relu);
      continue;
    }
    uint32_t r[32];
    tmem_ld32(t_row + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 v = make_float4(__uint_as_float(r[q * 4]) * p.alpha, __uint_as_float(r[q * 4 + 1]) * p.alpha, __uint_as_float(r[q * 4 + 2]) * p.alpha,
                             __uint_as_float(r[q * 4 + 3]) * p.alpha);
      if (p.bias) {
        const float4 b = *reinterpret_cast<const float4*>(bias_s + cb + q * 4);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      *reinterpret_cast<float4*>(stage + sw128_offset(lane, q)) = v;
    }
    __syncwarp();
    float* cbase = reinterpret_cast<float*>(p.C) + row0 * p.ldc + col0 + cchunk * 4;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + crow;
      if (row0 + rr < p.M) {
        const float4 v = *reinterpret_cast<const float4*>(stage + sw128_offset(rr, cchunk));
        if (accum) red_add_v4(cbase + (long long)rr * p.ldc, v);
        else *reinterpret_cast<float4*>(cbase + (long long)rr * p.ldc) = v;
      }
    }
    __syncwarp();
  }
}

// One tile's epilogue for one warp: bias slice -> smem (all 8 epilogue warps, named barrier 1), first aux prefetch, wait for
// the accumulator, drain this warp's 32 rows x BN/2 columns.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmTcParams& p, const CUtensorMap* tmc, uint8_t* stage, float* bias_tile, uint64_t* full_bar, uint32_t full_phase,
                                              uint32_t t_row, long long row0, int lane, int et, int n_base, int col_half) {
  const bool out_f32 = (p.flags & (B200F_EPI_OUT_F32 | B200F_EPI_ACCUM)) != 0;
  const bool accum = (p.flags & B200F_EPI_ACCUM) != 0;
  const bool relu = (p.flags & B200F_EPI_RELU) != 0;
  const int c_begin = col_half * (BN / 2);
  const bool staged16_ = !((p.flags & (B200F_EPI_OUT_F32 | B200F_EPI_ACCUM)) != 0) && p.vec_ok;
  const float bias_mul = (p.drop_thr && !p.residual && staged16_) ? p.inv_keep : 1.f;      // see fold_keep in epilogue_tile_bf16
  if (p.bias && et < BN) bias_tile[et] = (n_base + et < p.N) ? __ldg(p.bias + n_base + et) * bias_mul : 0.f;
  const bool staged16 = !out_f32 && p.vec_ok;
  if (staged16 && (p.residual || p.mask)) {
    static_assert(BN / 2 / 64 <= 2, "one staging tile per column block of a warp");
    if (p.tma_store) {                                         // both staging tiles are about to be overwritten: every bulk store of the
      if (lane == 0) tma_store_wait_read();                    // previous tile must be done reading them (it has had the whole mainloop)
      __syncwarp();
    }
    if (p.late_aux) {
      epi_issue_aux(p, stage, row0, n_base + c_begin, lane);
    } else {
#pragma unroll
      for (int blk = 0; blk < BN / 2 / 64; ++blk)
        epi_issue_aux(p, stage + (blk & 1) * p.epi_stride, row0, n_base + c_begin + blk * 64, lane);
    }
  }
  uint2 sb0 = make_uint2(0u, 0u), sb1 = sb0;                   // one-bit mask words of this lane's row: fetched before the accumulator wait
  if (staged16 && p.sbits && row0 + lane < p.M) {
    const uint32_t* sp = p.sbits + (row0 + lane) * p.ldsb + ((n_base + c_begin) >> 5);
    if (n_base + c_begin + 64 <= p.N) sb0 = __ldg(reinterpret_cast<const uint2*>(sp));
    if (BN / 2 > 64 && n_base + c_begin + 128 <= p.N) sb1 = __ldg(reinterpret_cast<const uint2*>(sp + 2));
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");             // bias slice visible to all epilogue warps
  mbar_wait(full_bar, full_phase);
  tc_fence_after();
  if (staged16)
    if (p.drop_thr || p.colsum || p.sbits || p.sbits_out) epilogue_tile_bf16<BN / 2, true>(p, tmc, stage, bias_tile + c_begin, t_row, row0, lane, n_base, c_begin, relu, sb0, sb1);
    else epilogue_tile_bf16<BN / 2, false>(p, tmc, stage, bias_tile + c_begin, t_row, row0, lane, n_base, c_begin, relu, sb0, sb1);
  else if (out_f32 && p.vec_ok)
    epilogue_tile_f32<BN / 2>(p, stage, bias_tile + c_begin, t_row, row0, lane, n_base, c_begin, accum, relu);
  else
    epilogue_columns(p, t_row, row0 + lane, row0 + lane < p.M, n_base, c_begin, c_begin + BN / 2, out_f32, accum, relu);
}


// ------------------------------------------------------------------------------------------------
// Specialised epilogues (round 2).  The generic epilogue above decides every feature at run time inside fully unrolled code; ncu on the
// K = 512 -> N = 2048 GEMMs of a MulT block showed what that costs once any "extra" is on: the two epilogue warps of a scheduler run
// a serial chain (6-7 clk per instruction), so every instruction is on the critical path, the kernel's 200 KB of SASS misses the
// instruction cache (15 % of the samples stalled on `no_inst`), and the 168-register cap (3 warps per scheduler) spills.  The hot
// feature sets therefore get their own kernel instantiation with the features as template constants (EPI bit set): no runtime branch,
// no dead registers, a fraction of the code.  Preconditions checked by the host: bf16 output, 16-byte aligned rows, N % 64 == 0 (a
// 64-column block is inside the matrix or skipped), bulk tensor stores, two staging tiles per warp.
// ------------------------------------------------------------------------------------------------
enum : int { E_SPEC = 1, E_BIAS = 2, E_RES = 4, E_RELU = 8, E_DROP = 16, E_SBOUT = 32, E_SBIN = 64, E_COLSUM = 128 };

template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile_spec(const GemmTcParams& p, const CUtensorMap* tmc, uint8_t* stage, float* bias_tile, uint64_t* full_bar,
                                                   uint32_t full_phase, uint32_t t_row, long long row0, int lane, int et, int n_base, int col_half) {
  constexpr bool BIAS = (EPI & E_BIAS) != 0, RES = (EPI & E_RES) != 0, RELU = (EPI & E_RELU) != 0, DROP = (EPI & E_DROP) != 0;
  constexpr bool SBOUT = (EPI & E_SBOUT) != 0, SBIN = (EPI & E_SBIN) != 0, COLSUM = (EPI & E_COLSUM) != 0;
  constexpr int NBLK = BN / 2 / 64;
  static_assert(NBLK == 2, "two 64-column blocks per warp, one staging tile each");
  static_assert(!(DROP && RES), "dropout rides on alpha and the bias slice: not combined with a residual here");
  const int c_begin = col_half * (BN / 2);
  const long long row = row0 + lane;
  if (BIAS && et < BN) bias_tile[et] = (n_base + et < p.N) ? __ldg(p.bias + n_base + et) * (DROP ? p.inv_keep : 1.f) : 0.f;
  if (RES) {                                                  // both staging tiles are about to receive the residual blocks: the bulk stores of
    if (lane == 0) tma_store_wait_read();                     // the previous tile must be done reading them (they have had the whole mainloop)
    __syncwarp();
#pragma unroll
    for (int blk = 0; blk < NBLK; ++blk) epi_issue_aux(p, stage + blk * EPI_STAGE_BYTES, row0, n_base + c_begin + blk * 64, lane);
  }
  uint2 sb[NBLK];
  if (SBIN) {                                                 // one-bit mask words of this lane's row, in flight across the accumulator wait
#pragma unroll
    for (int blk = 0; blk < NBLK; ++blk) {
      sb[blk] = make_uint2(0u, 0u);
      if (row < p.M && n_base + c_begin + blk * 64 + 64 <= p.N)
        sb[blk] = __ldg(reinterpret_cast<const uint2*>(p.sbits + row * p.ldsb + ((n_base + c_begin + blk * 64) >> 5)));
    }
  }
  if (BIAS) asm volatile("bar.sync 1, 256;" ::: "memory");    // bias slice visible to all epilogue warps
  mbar_wait(full_bar, full_phase);
  tc_fence_after();
  const float* bias_s = bias_tile + c_begin;
  const uint32_t rk = DROP ? drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, uint32_t(row)) : 0u;
  const float al = DROP ? p.alpha * p.inv_keep : p.alpha;     // 1/(1-p) folded into alpha and the bias slice; relu commutes with it
  const uint64_t al2 = f2_pack(al, al);
#pragma unroll
  for (int blk = 0; blk < NBLK; ++blk) {
    const int c0 = c_begin + blk * 64;
    const int col0 = n_base + c0;
    uint8_t* st = stage + blk * EPI_STAGE_BYTES;
    if (RES) {
      if (blk + 1 < NBLK) cp_async_wait<1>(); else cp_async_wait<0>();
    } else {
      if (lane == 0) tma_store_wait_read1();                  // the bulk store that last read THIS staging tile (two blocks ago) is done with it
    }
    __syncwarp();
    if (col0 >= p.N) continue;                                // warp-uniform; N % 64 == 0: a block is never ragged
    uint32_t r[2][32];
    tmem_ld32(t_row + c0, r[0]);
    tmem_ld32(t_row + c0 + 32, r[1]);
    uint4 aux[RES ? 8 : 1];
    if (RES) {
#pragma unroll
      for (int q = 0; q < 8; ++q) aux[q] = *reinterpret_cast<const uint4*>(st + sw128_offset(lane, q));
    }
    tmem_ld_wait();
    uint4 outv[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v[8];
        const uint32_t* a = &r[h][g * 8];
        if (BIAS) {
          const float4 b0 = *reinterpret_cast<const float4*>(bias_s + blk * 64 + h * 32 + g * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(bias_s + blk * 64 + h * 32 + g * 8 + 4);
          f2_unpack(f2_fma(f2_pack_u(a[0], a[1]), al2, f2_pack(b0.x, b0.y)), v[0], v[1]);
          f2_unpack(f2_fma(f2_pack_u(a[2], a[3]), al2, f2_pack(b0.z, b0.w)), v[2], v[3]);
          f2_unpack(f2_fma(f2_pack_u(a[4], a[5]), al2, f2_pack(b1.x, b1.y)), v[4], v[5]);
          f2_unpack(f2_fma(f2_pack_u(a[6], a[7]), al2, f2_pack(b1.z, b1.w)), v[6], v[7]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) f2_unpack(f2_mul(f2_pack_u(a[2 * j], a[2 * j + 1]), al2), v[2 * j], v[2 * j + 1]);
        }
        if (RES) {
          float f[8];
          Vec16<bf16> av; av.raw = aux[h * 4 + g]; av.unpack(f);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += f[j];
        }
        if (RELU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (DROP) {
          const uint32_t cc = uint32_t(col0 + h * 32 + g * 8) * kDropColMul;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = drop_keep_c(rk, cc + uint32_t(j) * kDropColMul, p.drop_thr) ? v[j] : 0.f;
        }
        if (SBIN) {
          const uint32_t w = h ? sb[blk].y : sb[blk].x;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (w & (1u << sbit_pos(g * 8 + j))) ? v[j] : 0.f;
        }
        Vec16<bf16> ov; ov.pack(v);
        outv[h * 4 + g] = ov.raw;
      }
    }
    if (SBOUT) {                                              // see the generic epilogue for the bit arithmetic
      uint32_t wb[2] = {0u, 0u};
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t t0 = outv[q].x + 0x7FFF7FFFu, t1 = outv[q].y + 0x7FFF7FFFu, t2 = outv[q].z + 0x7FFF7FFFu, t3 = outv[q].w + 0x7FFF7FFFu;
        wb[q >> 2] |= prmt(t0, t1, 0xFDB9u) & (0x01010101u << (2 * (q & 3)));
        wb[q >> 2] |= prmt(t2, t3, 0xFDB9u) & (0x01010101u << (2 * (q & 3) + 1));
      }
      if (row < p.M) *reinterpret_cast<uint2*>(p.sbits_out + row * p.ldsb + (col0 >> 5)) = make_uint2(wb[0], wb[1]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(st + sw128_offset(lane, q)) = outv[q];
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) { tma_store_2d(tmc, st, col0, int(row0)); tma_store_commit(); }
    if (COLSUM) {
      const int nr = p.M - row0 < 32 ? int(p.M - row0) : 32;
      float cs[8][2];
      colsum32x64_hmma(st, nr, lane, cs);
      if (lane < 4) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) red_add_v2(p.colsum + col0 + nt * 8 + 2 * lane, cs[nt][0], cs[nt][1]);
      }
      __syncwarp();
    }
  }
}

template <int BN, int STAGES, int A_MN, int B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_c,
               const GemmTcParams p) {
  using S = GemmSmem<BN, STAGES>;
  constexpr int TMEM_COLS = 2 * BN;
  extern __shared__ __align__(1024) uint8_t smem[];          // SWIZZLE_128B tiles need a 1024-byte aligned base (checked below)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int units = p.num_m * p.num_n * p.split_k;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int n_blk = u % p.num_n;
        const int m_blk = (u / p.num_n) % p.num_m;
        const int ks = u / (p.num_n * p.num_m);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * (BK * 128), &tma_a, &full_bar[stage], m_blk * BM + c * 64, kb * BK);
          } else {
            tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * (BK * 128), &tma_b, &full_bar[stage], n_blk * BN + c * 64, kb * BK);
          } else {
            tma_load_2d(sb, &tma_b, &full_bar[stage], kb * BK, n_blk * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      // Warp-uniform issue loop; only tcgen05.mma / commit are predicated on the elected lane.  The descriptors of stage 0 are
      // built once: a later stage / K step is one 64-bit add of (byte offset >> 4) to the address field (offsets < 256 KB).
      const bool one = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      const uint64_t a_desc0 = umma_desc(smem_u32(smem), p.a_lbo, p.a_sbo);
      const uint64_t b_desc0 = umma_desc(smem_u32(smem) + S::A_BYTES, p.b_lbo, p.b_sbo);
      const uint32_t a_step = p.a_kadv >> 4, b_step = p.b_kadv >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        const int ks = u / (p.num_n * p.num_m);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = a_desc0 + uint64_t(stage * (S::STAGE_BYTES >> 4)), db = b_desc0 + uint64_t(stage * (S::STAGE_BYTES >> 4));
          if (one) {
            umma_ss(d_tmem, da, db, idesc, kb > kb0 ? 1u : 0u);
#pragma unroll
            for (int k = 1; k < BK / 16; ++k) umma_ss_acc(d_tmem, da + uint64_t(k * a_step), db + uint64_t(k * b_step), idesc);
            umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (one) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    const int lane_grp = warp & 3;  // TMEM lanes [32*lane_grp, +32) are the ones this warp may touch
    const int col_half = (warp - 2) >> 2;  // which half of the tile's columns this warp drains
    uint8_t* stage = smem + S::EPI_OFFSET + (warp - 2) * 2 * EPI_STAGE_BYTES;
    float* bias_s = reinterpret_cast<float*>(smem + S::BIAS_OFFSET);
    int it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      const int n_blk = u % p.num_n;
      const int m_blk = (u / p.num_n) % p.num_m;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long long row0 = (long long)m_blk * BM + lane_grp * 32;
      const uint32_t t_row = tmem_base + (uint32_t(lane_grp * 32) << 16) + acc * BN;
      epilogue_tile<BN>(p, &tma_c, stage, bias_s + acc * BN, &tmem_full[acc], acc_phase, t_row, row0, lane, threadIdx.x - 64, n_blk * BN, col_half);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  if (p.tma_store && warp >= 2 && lane == 0) tma_store_wait_read();   // pending bulk stores still read this CTA's staging tiles
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster (= two SMs of a TPC) own one 256 x BN tile.  Each CTA
// stages its own 128 rows of A and HALF of the B tile, the leader issues 256 x BN x 16 MMAs that read both CTAs'
// shared memory and write 128 accumulator rows into each CTA's TMEM.  Per FLOP this moves 2/3 of the L2->SM bytes
// of the single-CTA kernel, which is what bounds the K=512 GEMMs of the path (L2, not HBM or the tensor pipe).
// Barrier ownership: smem-full and tmem-empty live in the LEADER (TMA of both CTAs completes on the leader's full
// barrier; epilogue warps of both CTAs arrive on the leader's tmem-empty); smem-empty and tmem-full are local to each
// CTA and signalled by multicast tcgen05.commit.
// ------------------------------------------------------------------------------------------------
// EB = epilogue staging tiles per warp: 2 when a residual / ReLU-mask block is prefetched one block ahead, 1 otherwise -- the
// 32 KB saved buy a sixth TMA stage (192 KB in flight per SM: the 5-stage ring holds exactly bandwidth x latency and runs dry)
template <int BN, int STAGES, int EB = 2>
struct PairSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 256;
  static constexpr int EPI_OFFSET = (BIAS_OFFSET + 2 * BN * 4 + 1023) / 1024 * 1024;
  static constexpr int TOTAL = EPI_OFFSET + 8 * EB * EPI_STAGE_BYTES;
  static_assert(TOTAL <= 232448, "exceeds the 227 KB dynamic shared memory limit");
};

template <int BN, int STAGES, int A_MN, int B_MN, int EB, int EPI = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_c,
                    const GemmTcParams p) {
  using S = PairSmem<BN, STAGES, EB>;
  constexpr int TMEM_COLS = 2 * BN;
  constexpr int HALF_N = BN / 2;
  extern __shared__ __align__(1024) uint8_t smem[];          // SWIZZLE_128B tiles need a 1024-byte aligned base (checked below)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);      // leader: armed with the bytes of both CTAs
      mbar_init(&empty_bar[i], 1);     // local: multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);     // local: multicast commit
      mbar_init(&tmem_empty[i], 16);   // leader: 8 epilogue warps x 2 CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // peer barriers initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int units = p.num_m * p.num_n * p.split_k;   // num_m counts 256-row pair tiles here

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = cluster_id; u < units; u += num_clusters) {
        const int n_blk = u % p.num_n;
        const int m_blk = (u / p.num_n) % p.num_m;
        const int ks = u / (p.num_n * p.num_m);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int m0 = m_blk * 2 * BM + int(rank) * BM;
        const int n0 = n_blk * BN + int(rank) * HALF_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d_pair(sa + c * (BK * 128), &tma_a, &full_bar[stage], m0 + c * 64, kb * BK);
          } else {
            tma_load_2d_pair(sa, &tma_a, &full_bar[stage], kb * BK, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < HALF_N / 64; ++c) tma_load_2d_pair(sb + c * (BK * 128), &tma_b, &full_bar[stage], n0 + c * 64, kb * BK);
          } else {
            tma_load_2d_pair(sb, &tma_b, &full_bar[stage], kb * BK, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader) {                                      // the whole warp of the leader CTA (warp-uniform); see the single-CTA kernel
      const bool one = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, A_MN, B_MN);
      const uint64_t a_desc0 = umma_desc(smem_u32(smem), p.a_lbo, p.a_sbo);
      const uint64_t b_desc0 = umma_desc(smem_u32(smem) + S::A_BYTES, p.b_lbo, p.b_sbo);
      const uint32_t a_step = p.a_kadv >> 4, b_step = p.b_kadv >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = cluster_id; u < units; u += num_clusters, ++it) {
        const int ks = u / (p.num_n * p.num_m);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = a_desc0 + uint64_t(stage * (S::STAGE_BYTES >> 4)), db = b_desc0 + uint64_t(stage * (S::STAGE_BYTES >> 4));
          if (one) {
            umma_ss_pair(d_tmem, da, db, idesc, kb > kb0 ? 1u : 0u);
#pragma unroll
            for (int k = 1; k < BK / 16; ++k) umma_ss_pair_acc(d_tmem, da + uint64_t(k * a_step), db + uint64_t(k * b_step), idesc);
            umma_commit_pair(&empty_bar[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (one) umma_commit_pair(&tmem_full[acc]);
      }
    }
    __syncwarp();
  } else {
    const int lane_grp = warp & 3;
    const int col_half = (warp - 2) >> 2;
    uint8_t* stage = smem + S::EPI_OFFSET + (warp - 2) * EB * EPI_STAGE_BYTES;
    float* bias_s = reinterpret_cast<float*>(smem + S::BIAS_OFFSET);
    int it = 0;
    for (int u = cluster_id; u < units; u += num_clusters, ++it) {
      const int n_blk = u % p.num_n;
      const int m_blk = (u / p.num_n) % p.num_m;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long long row0 = (long long)m_blk * 2 * BM + (long long)rank * BM + lane_grp * 32;
      const uint32_t t_row = tmem_base + (uint32_t(lane_grp * 32) << 16) + acc * BN;
      if constexpr (EPI != 0)
        epilogue_tile_spec<BN, EPI>(p, &tma_c, stage, bias_s + acc * BN, &tmem_full[acc], acc_phase, t_row, row0, lane, threadIdx.x - 64, n_blk * BN, col_half);
      else
        epilogue_tile<BN>(p, &tma_c, stage, bias_s + acc * BN, &tmem_full[acc], acc_phase, t_row, row0, lane, threadIdx.x - 64, n_blk * BN, col_half);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
    }
  }

  if (p.tma_store && warp >= 2 && lane == 0) tma_store_wait_read();   // pending bulk stores still read this CTA's staging tiles
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // the peer may still read this CTA's smem / signal its barriers
  if (warp == 1) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// bf16 tensor map of rank 2..4 with SWIZZLE_128B; dims[0] is the contiguous dimension.
int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(B200F_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t d[4]; cuuint64_t s[3]; cuuint32_t b[4]; cuuint32_t e[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(B200F_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu] stride %llu box [%u,%u] base %p", int(r), rank,
                (unsigned long long)d[0], (unsigned long long)d[1], (unsigned long long)s[0], b[0], b[1], base);
  return B200F_OK;
}

// debug overrides for the MN-major descriptor geometry (b200f_debug_set); 0 = use the default
uint32_t g_dbg_mn_lbo = 0, g_dbg_mn_sbo = 0, g_dbg_mn_kadv = 0;

bool g_dbg_disable_pair = false;     // b200f_debug_set(3, 1): force the single-CTA kernel (A/B testing)
bool g_dbg_no_tma_store = false;     // b200f_debug_set(8, 1): LDS + STG copy-out instead of bulk tensor stores (A/B testing)
int g_dbg_colsum = 0;                // b200f_debug_set(13, v): see GemmTcParams::dbg_colsum
bool g_dbg_no_spec_epi = false;      // b200f_debug_set(14, 1): every launch through the generic (run-time flag) epilogue (A/B testing)
bool g_dbg_late_aux = false;         // b200f_debug_set(12, 1): aux blocks prefetched one column block ahead (round-1 schedule) instead of at tile start
bool g_dbg_no_tma_store_aux = false; // b200f_debug_set(11, 1): launches with a residual / mask block keep the LDS + STG copy-out (A/B testing)
bool g_dbg_six_stages = false;       // b200f_debug_set(7, 1): 6-stage / one-staging-tile pair kernel for launches without an aux block.
                                     // Measured no faster than 5 stages on any MulT shape (profiles/r01_e): the ring depth is not the limiter.

template <int BN, int STAGES, int A_MN, int B_MN, int EB, int EPI = 0>
static int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, GemmTcParams p, int grid, cudaStream_t st) {
  using S = PairSmem<BN, STAGES, EB>;
  auto kern = gemm_tc_pair_kernel<BN, STAGES, A_MN, B_MN, EB, EPI>;
  p.epi_stride = (EB - 1) * EPI_STAGE_BYTES;
  static PerDeviceOnce configured;
  if (configured.first()) {
    B200F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  }
  kern<<<grid, GEMM_THREADS, S::TOTAL, st>>>(ta, tb, tc, p);
  return check_launch("gemm_tc_pair_kernel");
}

template <int BN, int STAGES, int A_MN, int B_MN>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmTcParams& p, int grid, cudaStream_t st) {
  using S = GemmSmem<BN, STAGES>;
  auto kern = gemm_tc_kernel<BN, STAGES, A_MN, B_MN>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    B200F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  }
  kern<<<grid, GEMM_THREADS, S::TOTAL, st>>>(ta, tb, tc, p);
  return check_launch("gemm_tc_kernel");
}

// The TMA path needs 16-byte aligned bases and row strides for A and B; everything else is handled in-kernel.
bool gemm_tc_eligible(const b200f_gemm_args& a) {
  return a.lda % 8 == 0 && a.ldb % 8 == 0 && aligned16(a.A) && aligned16(a.B) && a.N >= 8 && a.K >= 8 && a.M >= 1;
}

// The column-sum epilogue lives in the staged bf16 path only: full 64-column blocks, vector accesses, 16-byte aligned accumulator.
bool gemm_tc_colsum_fused(const b200f_gemm_args& a) {
  const bool out_f32 = (a.flags & (B200F_EPI_OUT_F32 | B200F_EPI_ACCUM)) != 0;
  const bool vec_ok = aligned16(a.C) && a.ldc % 8 == 0 && (!a.residual || (a.ldr % 8 == 0 && aligned16(a.residual))) &&
                      (!a.relu_mask || (a.ldm % 8 == 0 && aligned16(a.relu_mask))) && (!a.bias || aligned16(a.bias));
  return gemm_tc_eligible(a) && !out_f32 && vec_ok && a.N % 64 == 0 && (!a.colsum || aligned16(a.colsum));
}

int gemm_bf16_tc(const b200f_gemm_args& a, cudaStream_t st) {
  B200F_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, B200F_ERR_SHAPE, "gemm: empty shape M=%lld N=%lld K=%lld", (long long)a.M, (long long)a.N, (long long)a.K);
  B200F_REQUIRE(gemm_tc_eligible(a), B200F_ERR_ALIGN, "gemm(bf16/tcgen05): lda/ldb must be multiples of 8 elements and A/B 16-byte aligned");
  const bool out_f32 = (a.flags & (B200F_EPI_OUT_F32 | B200F_EPI_ACCUM)) != 0;
  const bool vec_ok = aligned16(a.C) && a.ldc % (out_f32 ? 4 : 8) == 0 && (!a.residual || (a.ldr % 8 == 0 && aligned16(a.residual))) &&
                      (!a.relu_mask || (a.ldm % 8 == 0 && aligned16(a.relu_mask))) && (!a.bias || aligned16(a.bias));
  const int split = a.split_k > 1 ? a.split_k : 1;
  B200F_REQUIRE(split == 1 || (a.flags & B200F_EPI_ACCUM), B200F_ERR_UNSUPPORTED, "gemm: split_k needs B200F_EPI_ACCUM");

  const int BN = (a.N > 128) ? 256 : 128;
  // CTA pairs when there is enough work for the 74 pairs of the chip; tiny-M problems stay on single CTAs
  const bool pair = !g_dbg_disable_pair && BN == 256 && a.M > 2 * BM;
  CUtensorMap ta, tb;
  {
    uint64_t dims[2], strides[1]; uint32_t box[2];
    if (a.a_layout == 0) { dims[0] = a.K; dims[1] = a.M; box[0] = 64; box[1] = BM; }
    else                 { dims[0] = a.M; dims[1] = a.K; box[0] = 64; box[1] = BK; }
    strides[0] = uint64_t(a.lda) * 2;
    int rc = make_tmap_bf16(&ta, a.A, 2, dims, strides, box);
    if (rc) return rc;
    if (a.b_layout == 0) { dims[0] = a.K; dims[1] = a.N; box[0] = 64; box[1] = pair ? BN / 2 : BN; }
    else                 { dims[0] = a.N; dims[1] = a.K; box[0] = 64; box[1] = BK; }
    strides[0] = uint64_t(a.ldb) * 2;
    rc = make_tmap_bf16(&tb, a.B, 2, dims, strides, box);
    if (rc) return rc;
  }
  GemmTcParams p;
  p.M = int(a.M); p.N = int(a.N); p.K = int(a.K);
  p.num_m = pair ? int((a.M + 2 * BM - 1) / (2 * BM)) : int((a.M + BM - 1) / BM);
  p.num_n = int((a.N + BN - 1) / BN);
  p.kb_total = int((a.K + BK - 1) / BK);
  p.split_k = split > p.kb_total ? p.kb_total : split;
  p.kb_per_split = (p.kb_total + p.split_k - 1) / p.split_k;
  p.split_k = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.C = a.C; p.ldc = a.ldc;
  p.bias = a.bias;
  p.residual = static_cast<const bf16*>(a.residual); p.ldr = a.ldr;
  p.mask = static_cast<const bf16*>(a.relu_mask); p.ldm = a.ldm;
  p.alpha = a.alpha; p.flags = a.flags;
  p.vec_ok = vec_ok ? 1 : 0;
  // bf16 output, nothing to prefetch into the staging tiles, full 64-column blocks: the epilogue stores through TMA
  // (round 2: also with a residual / ReLU-mask block -- it is prefetched into the staging tile by cp.async, consumed into registers,
  //  and the tile then leaves through the bulk store like a plain one; b200f_debug_set(11, 1) restores LDS + STG for those launches)
  const bool has_aux = a.residual || a.relu_mask;
  p.tma_store = (!out_f32 && vec_ok && (!has_aux || !g_dbg_no_tma_store_aux) && a.N % 64 == 0 && !g_dbg_no_tma_store) ? 1 : 0;
  CUtensorMap tc;
  memset(&tc, 0, sizeof(tc));
  if (p.tma_store) {
    uint64_t dims[2] = {uint64_t(a.N), uint64_t(a.M)}, strides[1] = {uint64_t(a.ldc) * 2};
    uint32_t box[2] = {64, 32};
    int rc = make_tmap_bf16(&tc, a.C, 2, dims, strides, box);
    if (rc) return rc;
  }
  p.epi_stride = EPI_STAGE_BYTES;
  p.late_aux = g_dbg_late_aux ? 1 : 0;
  p.dbg_colsum = g_dbg_colsum;
  p.colsum = a.colsum;
  p.drop_thr = drop_threshold(a.dropout_p); p.drop_seed_lo = a.drop_seed_lo; p.drop_seed_hi = a.drop_seed_hi;
  p.inv_keep = 1.f / (1.f - a.dropout_p);
  if (a.colsum || p.drop_thr)
    B200F_REQUIRE(gemm_tc_colsum_fused(a), B200F_ERR_UNSUPPORTED, "gemm(tcgen05): fused colsum / dropout need bf16 output, N %% 64 == 0 and aligned rows");
  p.sbits_out = a.sign_bits_out; p.sbits = a.sign_bits; p.ldsb = a.ldsb;
  if (a.sign_bits_out || a.sign_bits) {
    B200F_REQUIRE(gemm_tc_colsum_fused(a), B200F_ERR_UNSUPPORTED, "gemm(tcgen05): sign_bits need bf16 output, N %% 64 == 0 and aligned rows");
    B200F_REQUIRE(a.ldsb % 2 == 0 && a.ldsb * 32 >= a.N && (reinterpret_cast<uintptr_t>(a.sign_bits_out) & 7) == 0 && (reinterpret_cast<uintptr_t>(a.sign_bits) & 7) == 0,
                  B200F_ERR_ALIGN, "gemm: sign_bits rows must be 8-byte aligned (ldsb even, ldsb * 32 >= N)");
    B200F_REQUIRE(!a.sign_bits_out || (a.flags & B200F_EPI_RELU), B200F_ERR_UNSUPPORTED, "gemm: sign_bits_out is the mask of a ReLU output (B200F_EPI_RELU)");
    B200F_REQUIRE(!(a.sign_bits && a.relu_mask), B200F_ERR_UNSUPPORTED, "gemm: relu_mask and sign_bits are two forms of the same mask; pass one");
  }
  // K-major SW128: rows of 128 B, 8-row groups 1024 B apart, +32 B per K=16 step inside the swizzle row.
  // MN-major SW128: 64-element (128 B) MN chunks, k rows 128 B apart, 8-k-row groups 1024 B apart (SBO),
  //                 MN chunks BK*128 B apart (LBO), +16 k rows = 2048 B per K=16 step.
  const uint32_t mn_lbo = g_dbg_mn_lbo ? g_dbg_mn_lbo : BK * 128, mn_sbo = g_dbg_mn_sbo ? g_dbg_mn_sbo : 1024,
                 mn_kadv = g_dbg_mn_kadv ? g_dbg_mn_kadv : 2048;
  if (a.a_layout == 0) { p.a_lbo = 16; p.a_sbo = 1024; p.a_kadv = 32; } else { p.a_lbo = mn_lbo; p.a_sbo = mn_sbo; p.a_kadv = mn_kadv; }
  if (a.b_layout == 0) { p.b_lbo = 16; p.b_sbo = 1024; p.b_kadv = 32; } else { p.b_lbo = mn_lbo; p.b_sbo = mn_sbo; p.b_kadv = mn_kadv; }

  const int units = p.num_m * p.num_n * p.split_k;
  if (pair) {
    const int clusters = units < num_sms() / 2 ? units : num_sms() / 2;
    const int pkey = (a.a_layout ? 2 : 0) | (a.b_layout ? 1 : 0);
    if (!a.residual && !a.relu_mask && g_dbg_six_stages) {     // experiment (b200f_debug_set(7, 1)): one staging tile per warp, six TMA stages
      switch (pkey) {
        case 0: return launch_pair<256, 6, 0, 0, 1>(ta, tb, tc, p, 2 * clusters, st);
        case 1: return launch_pair<256, 6, 0, 1, 1>(ta, tb, tc, p, 2 * clusters, st);
        case 2: return launch_pair<256, 6, 1, 0, 1>(ta, tb, tc, p, 2 * clusters, st);
        default: return launch_pair<256, 6, 1, 1, 1>(ta, tb, tc, p, 2 * clusters, st);
      }
    }
    // the hot feature sets of the MulT schedule run kernels whose epilogue has them as template constants (epilogue_tile_spec)
    if (p.tma_store && vec_ok && a.N % 64 == 0 && split == 1 && !g_dbg_no_spec_epi && !g_dbg_late_aux && !g_dbg_colsum && !a.relu_mask) {
      const int feat = E_SPEC | (a.bias ? E_BIAS : 0) | (a.residual ? E_RES : 0) | ((a.flags & B200F_EPI_RELU) ? E_RELU : 0) | (p.drop_thr ? E_DROP : 0) |
                       (a.sign_bits_out ? E_SBOUT : 0) | (a.sign_bits ? E_SBIN : 0) | (a.colsum ? E_COLSUM : 0);
      if (pkey == 0 && feat == (E_SPEC | E_BIAS | E_RES)) return launch_pair<256, 5, 0, 0, 2, E_SPEC | E_BIAS | E_RES>(ta, tb, tc, p, 2 * clusters, st);
      if (pkey == 0 && feat == (E_SPEC | E_BIAS | E_RELU | E_DROP | E_SBOUT))
        return launch_pair<256, 5, 0, 0, 2, E_SPEC | E_BIAS | E_RELU | E_DROP | E_SBOUT>(ta, tb, tc, p, 2 * clusters, st);
      if (pkey == 0 && feat == (E_SPEC | E_BIAS | E_RELU | E_SBOUT)) return launch_pair<256, 5, 0, 0, 2, E_SPEC | E_BIAS | E_RELU | E_SBOUT>(ta, tb, tc, p, 2 * clusters, st);
      if (pkey == 1 && feat == (E_SPEC | E_SBIN | E_COLSUM)) return launch_pair<256, 5, 0, 1, 2, E_SPEC | E_SBIN | E_COLSUM>(ta, tb, tc, p, 2 * clusters, st);
      if (pkey == 1 && feat == (E_SPEC | E_RES)) return launch_pair<256, 5, 0, 1, 2, E_SPEC | E_RES>(ta, tb, tc, p, 2 * clusters, st);
      if (pkey == 0 && feat == (E_SPEC | E_BIAS)) return launch_pair<256, 5, 0, 0, 2, E_SPEC | E_BIAS>(ta, tb, tc, p, 2 * clusters, st);   // packed projections
      if (pkey == 1 && feat == E_SPEC) return launch_pair<256, 5, 0, 1, 2, E_SPEC>(ta, tb, tc, p, 2 * clusters, st);                        // plain input gradients
    }
    switch (pkey) {
      case 0: return launch_pair<256, 5, 0, 0, 2>(ta, tb, tc, p, 2 * clusters, st);
      case 1: return launch_pair<256, 5, 0, 1, 2>(ta, tb, tc, p, 2 * clusters, st);
      case 2: return launch_pair<256, 5, 1, 0, 2>(ta, tb, tc, p, 2 * clusters, st);
      default: return launch_pair<256, 5, 1, 1, 2>(ta, tb, tc, p, 2 * clusters, st);
    }
  }
  const int grid = units < num_sms() ? units : num_sms();
  const int key = (BN == 256 ? 4 : 0) | (a.a_layout ? 2 : 0) | (a.b_layout ? 1 : 0);
  switch (key) {
    case 0: return launch_cfg<128, 4, 0, 0>(ta, tb, tc, p, grid, st);
    case 1: return launch_cfg<128, 4, 0, 1>(ta, tb, tc, p, grid, st);
    case 2: return launch_cfg<128, 4, 1, 0>(ta, tb, tc, p, grid, st);
    case 3: return launch_cfg<128, 4, 1, 1>(ta, tb, tc, p, grid, st);
    case 4: return launch_cfg<256, 3, 0, 0>(ta, tb, tc, p, grid, st);
    case 5: return launch_cfg<256, 3, 0, 1>(ta, tb, tc, p, grid, st);
    case 6: return launch_cfg<256, 3, 1, 0>(ta, tb, tc, p, grid, st);
    default: return launch_cfg<256, 3, 1, 1>(ta, tb, tc, p, grid, st);
  }
}

B200F_DEFINE_EPOCH_HOOK(gemm_tc)

}  // namespace b200f
