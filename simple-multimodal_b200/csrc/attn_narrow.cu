// Attention with ONE NARROW SIDE (bf16, head dim 64): the cross-modal blocks that touch the video stream (30 frames) --
// text->video / audio->video (Lq = 512, Lk = 30) and video->text / video->audio (Lq = 30, Lk = 512); reference
// models/fusion_layers.py:146-153 -- and the 30 x 30 video self-attention (:165-167).
//
// These blocks carry 1/17 of the FLOPs of a 512 x 512 block but the same bytes on their long side, so they are HBM-bound:
// per (batch, head) a forward reads or writes ~2 x L_long x 128 B and needs ~0.5 MFLOP per KB moved.  The 128-wide tcgen05
// tiles of attn_tc.cu pad the narrow side 30 -> 128 and run one latency chain (TMA -> MMA -> TMEM -> softmax -> MMA -> store)
// per work item with two items in flight per SM: 0.08-0.21 ms per launch against a 0.045-0.09 ms HBM time.  Here the narrow
// operand (<= 32 rows) stays resident in shared memory / registers for the whole (batch, head), the long side streams through a
// cp.async ring in 64-row tiles, and the 16 x 32 x 64 products run on warp-level mma.sync (m16n8k16, fp32 accumulate): at this
// arithmetic intensity the legacy tensor path is far from being the limit, needs no TMEM round trip or mbarrier protocol, and
// lets 8-20 warps per SM hide the memory latency.  Scores / probabilities / dS never leave registers (the accumulator layout of
// one product is the A-operand layout of the next; transposed operands via movmatrix).
//
//   NK = narrow keys    (Lk <= 32):  forward: CTA = 128 query rows of one (b, h); backward: CTA = one (b, h), loops over the
//                                    queries, dK / dV (32 x 64 each) accumulate in registers, one cross-warp reduction at the end.
//   NQ = narrow queries (Lq <= 32):  CTA = one (b, h), the four warps split every 64-key tile; forward merges the four partial
//                                    (max, sum, O) at the end (split-key softmax); backward writes dK / dV of a tile as soon as it
//                                    is computed (all <= 32 queries are in the product) and reduces dQ across warps at the end.
// Dropout, LSE, delta and the bias-gradient column sums follow attn_tc.cu exactly (same counter-based mask, csrc/common.cuh).
#include "common.cuh"
#include "ptx.cuh"

namespace b200f {

namespace {

constexpr int HD = 64;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int NARROW = 32;     // rows of the narrow operand (padded with zero rows)

struct NarrowParams {
  int B, H, Lq, Lk;
  float scale;
  const bf16* Q; long long ldq;
  const bf16* K; long long ldk;
  const bf16* V; long long ldv;
  bf16* O; long long ldo;               // forward: output; backward: the forward output (read for delta)
  float* LSE;
  const bf16* dO; long long lddo;
  bf16* dQ; long long lddq;
  bf16* dK; long long lddk;
  bf16* dV; long long lddv;
  float* dbq; float* dbk; float* dbv;
  float* pool;                          // forward, optional [B, parts, H*64] fp32: partial column sums of the stored O (NK: one part per 16 query
                                        // rows; NQ: one part), each written by exactly one warp / thread set
  uint32_t drop_thr, drop_seed_lo, drop_seed_hi;
  float inv_keep;
};

// ---- PTX helpers -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp16(uint32_t saddr, const void* g, bool pred) {   // 16-byte async copy; !pred: zero fill, no read
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D(16x8, fp32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// transpose an 8x8 bf16 block held in fragment layout (thread T: row T/4, columns 2(T%4), 2(T%4)+1)
__device__ __forceinline__ uint32_t movm_t(uint32_t x) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}

// Shared-memory tiles are [rows x 128 B] (64 bf16 per row), 16-byte chunks XOR-swizzled by row % 8 (conflict-free ldmatrix).
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return uint32_t(row) * 128u + (uint32_t((chunk ^ row) & 7) << 4); }

// A fragment (16 rows x 16 k) of a row-major tile: rows row0.., k = kk*16..
__device__ __forceinline__ void ld_a(uint32_t (&a)[4], uint32_t tile, int row0, int kk, int lane) {
  const int mi = lane >> 3, r = lane & 7;
  ldsm4(a, tile + swz(row0 + (mi & 1) * 8 + r, kk * 2 + (mi >> 1)));
}
// B fragments of two adjacent n-tiles from a tile whose ROWS are n (k contiguous): n = n0..n0+15, k = kk*16..; b[0],b[1] -> n-tile
// n0, b[2],b[3] -> n-tile n0+8
__device__ __forceinline__ void ld_b_nk(uint32_t (&b)[4], uint32_t tile, int n0, int kk, int lane) {
  const int mi = lane >> 3, r = lane & 7;
  ldsm4(b, tile + swz(n0 + (mi >> 1) * 8 + r, kk * 2 + (mi & 1)));
}
// B fragments of two adjacent n-tiles from a tile whose ROWS are k (n contiguous): k = k0..k0+15, n = np*16..; b[0],b[1] -> n-tile
// 2np, b[2],b[3] -> n-tile 2np+1
__device__ __forceinline__ void ld_b_kn(uint32_t (&b)[4], uint32_t tile, int k0, int np, int lane) {
  const int mi = lane >> 3, r = lane & 7;
  ldsm4t(b, tile + swz(k0 + (mi & 1) * 8 + r, np * 2 + (mi >> 1)));
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// async load of `rows` token rows (row index r -> token tok0 + r, valid while < L) of one head into a swizzled tile
__device__ __forceinline__ void load_rows(uint32_t tile, const bf16* base, long long ld, long long tok_base, int tok0, int L, int rows, int h, int tid,
                                          int nthreads) {
  for (int i = tid; i < rows * 8; i += nthreads) {
    const int r = i >> 3, c = i & 7;
    const bool ok = tok0 + r < L;
    cp16(tile + swz(r, c), base + (tok_base + (ok ? tok0 + r : 0)) * ld + h * HD + c * 8, ok);
  }
}

// a warp's 16 x 64 fp32 accumulator tile (x mul) -> bf16 -> rows row0..row0+15 of a swizzled tile (warp-private rows)
__device__ __forceinline__ void stage_tile(uint8_t* tile, int row0, const float (&acc)[8][4], float mul0, float mul1, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(tile + swz(row0 + g, nt) + t * 4) = pack_bf16(acc[nt][0] * mul0, acc[nt][1] * mul0);
    *reinterpret_cast<uint32_t*>(tile + swz(row0 + g + 8, nt) + t * 4) = pack_bf16(acc[nt][2] * mul1, acc[nt][3] * mul1);
  }
}
// rows row0..row0+15 of a swizzled tile -> global (full 128-byte lines), rows valid while tok0 + r < L; optionally returns this
// lane's two column sums (columns 2*lane, 2*lane+1) over the valid rows
__device__ __forceinline__ void store_tile(const uint8_t* tile, int row0, bf16* base, long long ld, long long tok_base, int tok0, int L, int h, int lane,
                                           float* cs0 = nullptr, float* cs1 = nullptr) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = i * 32 + lane, r = idx >> 3, c = idx & 7;
    if (tok0 + r < L)
      *reinterpret_cast<uint4*>(base + (tok_base + tok0 + r) * ld + h * HD + c * 8) = *reinterpret_cast<const uint4*>(tile + swz(row0 + r, c));
  }
  if (cs0) {
    float a0 = 0.f, a1 = 0.f;
    const int nr = min(16, L - tok0);
    for (int r = 0; r < nr; ++r) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(tile + swz(row0 + r, lane >> 2) + (lane & 3) * 4);
      a0 += __uint_as_float(w << 16);
      a1 += __uint_as_float(w & 0xFFFF0000u);
    }
    *cs0 += a0; *cs1 += a1;
  }
}

// ==============================================================================================================================
// NK forward: Lk <= 32.  grid (ceil(Lq/128), H, B), 128 threads, each warp two 16-row query groups.
// ==============================================================================================================================
template <bool DROP>
__global__ void __launch_bounds__(128) attn_nk_fwd_kernel(const NarrowParams p) {
  __shared__ __align__(128) uint8_t Qs[128 * 128];
  __shared__ __align__(128) uint8_t Ks[NARROW * 128];
  __shared__ __align__(128) uint8_t Vs[NARROW * 128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const uint32_t sq = smem_u32(Qs), sk = smem_u32(Ks), sv = smem_u32(Vs);
  load_rows(sk, p.K, p.ldk, (long long)b * p.Lk, 0, p.Lk, NARROW, h, tid, 128);
  load_rows(sv, p.V, p.ldv, (long long)b * p.Lk, 0, p.Lk, NARROW, h, tid, 128);
  load_rows(sq, p.Q, p.ldq, (long long)b * p.Lq, q0, p.Lq, 128, h, tid, 128);
  cp_commit();
  cp_wait<0>();
  __syncthreads();
  const float c = p.scale * LOG2E;
#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    const int r0 = warp * 32 + mt * 16;
    if (q0 + r0 >= p.Lq) break;
    float s[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4], bb[4];
      ld_a(a, sq, r0, kk, lane);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        ld_b_nk(bb, sk, np * 16, kk, lane);
        mma16816(s[2 * np], a, bb[0], bb[1]);
        mma16816(s[2 * np + 1], a, bb[2], bb[3]);
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (nt * 8 + 2 * t + e >= p.Lk) { s[nt][e] = -INFINITY; s[nt][2 + e] = -INFINITY; }
        mx0 = fmaxf(mx0, s[nt][e]);
        mx1 = fmaxf(mx1, s[nt][2 + e]);
      }
    mx0 = quad_max(mx0); mx1 = quad_max(mx1);
    const int row_a = q0 + r0 + g, row_b = row_a + 8;
    const uint32_t rbase = uint32_t((b * p.H + h) * p.Lq);
    const uint32_t rk0 = DROP ? drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, rbase + row_a) : 0u;
    const uint32_t rk1 = DROP ? drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, rbase + row_b) : 0u;
    float l0 = 0.f, l1 = 0.f;
    uint32_t pa[2][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float pv[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        pv[e] = ex2_approx((s[nt][e] - mx0) * c);
        pv[2 + e] = ex2_approx((s[nt][2 + e] - mx1) * c);
      }
      l0 += pv[0] + pv[1];
      l1 += pv[2] + pv[3];
      if (DROP) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const uint32_t cm = drop_col_attn(uint32_t(nt * 8 + 2 * t + e));
          if (!drop_keep_c(rk0, cm, p.drop_thr)) pv[e] = 0.f;
          if (!drop_keep_c(rk1, cm, p.drop_thr)) pv[2 + e] = 0.f;
        }
      }
      pa[nt >> 1][(nt & 1) * 2] = pack_bf16(pv[0], pv[1]);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(pv[2], pv[3]);
    }
    l0 = quad_sum(l0); l1 = quad_sum(l1);
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t bb[4];
        ld_b_kn(bb, sv, ks * 16, dp, lane);
        mma16816(o[2 * dp], pa[ks], bb[0], bb[1]);
        mma16816(o[2 * dp + 1], pa[ks], bb[2], bb[3]);
      }
    if (t == 0) {
      const long long lb = ((long long)b * p.H + h) * p.Lq;
      if (row_a < p.Lq) p.LSE[lb + row_a] = mx0 * p.scale + logf(l0);
      if (row_b < p.Lq) p.LSE[lb + row_b] = mx1 * p.scale + logf(l1);
    }
    const float ik = DROP ? p.inv_keep : 1.f;
    __syncwarp();
    stage_tile(Qs, r0, o, ik / l0, ik / l1, lane);      // this warp's own (already consumed) query rows
    __syncwarp();
    float ps0 = 0.f, ps1 = 0.f;
    store_tile(Qs, r0, p.O, p.ldo, (long long)b * p.Lq, q0 + r0, p.Lq, h, lane, p.pool ? &ps0 : nullptr, p.pool ? &ps1 : nullptr);
    if (p.pool)                                           // this 16-row group's own slot of the partial sums: a plain store
      *reinterpret_cast<float2*>(p.pool + (((long long)b * ((p.Lq + 15) / 16) + (q0 + r0) / 16) * p.H + h) * HD + 2 * lane) = make_float2(ps0, ps1);
  }
}

// ==============================================================================================================================
// NK backward: Lk <= 32.  grid (H, B), 128 threads; 64-query tiles of {Q, dO, O} through a 3-deep cp.async ring, a warp owns 16
// query rows per tile.  Per 16 rows: S = Q K^T, dP = dO V^T, dS, dQ = dS K (stored), dV += Pd^T dO, dK += dS^T Q.
// ==============================================================================================================================
constexpr int NKB_ST = 3;
constexpr int NKB_TILE = 64 * 128;                         // one operand tile: 64 rows x 128 B
constexpr int NKB_STAGE = 3 * NKB_TILE;                    // Q, dO, O
constexpr int NKB_SMEM = 2 * NARROW * 128 + NKB_ST * NKB_STAGE + 256;
constexpr int NKB_RSTRIDE = HD + 8;                        // fp32 row stride of the per-warp dK / dV partials (bank spread)
static_assert(4 * 2 * NARROW * NKB_RSTRIDE * 4 <= NKB_ST * NKB_STAGE, "the per-warp partials must fit in the ring");

template <bool DROP>
__global__ void __launch_bounds__(128, 2) attn_nk_bwd_kernel(const NarrowParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* Ks = smem;
  uint8_t* Vs = smem + NARROW * 128;
  uint8_t* ring = smem + 2 * NARROW * 128;
  float* dbq_s = reinterpret_cast<float*>(ring + NKB_ST * NKB_STAGE);      // 64 floats
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int h = blockIdx.x, b = blockIdx.y;
  const uint32_t sk = smem_u32(Ks), sv = smem_u32(Vs), sring = smem_u32(ring);
  const int n_tiles = (p.Lq + 63) / 64;
  const long long qtok = (long long)b * p.Lq;
  auto issue = [&](int i) {
    if (i < n_tiles) {
      const uint32_t st = sring + (i % NKB_ST) * NKB_STAGE;
      load_rows(st, p.Q, p.ldq, qtok, i * 64, p.Lq, 64, h, tid, 128);
      load_rows(st + NKB_TILE, p.dO, p.lddo, qtok, i * 64, p.Lq, 64, h, tid, 128);
      load_rows(st + 2 * NKB_TILE, p.O, p.ldo, qtok, i * 64, p.Lq, 64, h, tid, 128);
    }
    cp_commit();
  };
  load_rows(sk, p.K, p.ldk, (long long)b * p.Lk, 0, p.Lk, NARROW, h, tid, 128);
  load_rows(sv, p.V, p.ldv, (long long)b * p.Lk, 0, p.Lk, NARROW, h, tid, 128);
  issue(0);
  issue(1);
  if (tid < 64) dbq_s[tid] = 0.f;
  float dv[2][8][4], dk[2][8][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) { dv[j][nt][e] = 0.f; dk[j][nt][e] = 0.f; }
  float dbq0 = 0.f, dbq1 = 0.f;
  const float c = p.scale * LOG2E;
  const long long lb = ((long long)b * p.H + h) * p.Lq;
#pragma unroll 1
  for (int i = 0; i < n_tiles; ++i) {
    cp_wait<1>();
    __syncthreads();                 // tile i visible to all; every warp is done with tile i-1 (its buffer is refilled next)
    issue(i + 2);
    uint8_t* stg = ring + (i % NKB_ST) * NKB_STAGE;
    const uint32_t sq = sring + (i % NKB_ST) * NKB_STAGE, sdo = sq + NKB_TILE;
    const uint8_t* Qt = stg; (void)Qt;
    const uint8_t* dOt = stg + NKB_TILE;
    const uint8_t* Ot = stg + 2 * NKB_TILE;
    const int r0 = warp * 16;
    const int tok0 = i * 64 + r0;
    if (tok0 >= p.Lq) continue;      // warp-uniform; the barriers above were already passed
    const int row_a = tok0 + g, row_b = row_a + 8;
    // delta = rowsum(dO o O): a quad shares a row, each thread 16 of its 64 columns
    float dl0 = 0.f, dl1 = 0.f;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      Vec16<bf16> x, y;
      float fx[8], fy[8];
      x.raw = *reinterpret_cast<const uint4*>(dOt + swz(r0 + g, 2 * t + cc));
      y.raw = *reinterpret_cast<const uint4*>(Ot + swz(r0 + g, 2 * t + cc));
      x.unpack(fx); y.unpack(fy);
#pragma unroll
      for (int e = 0; e < 8; ++e) dl0 = fmaf(fx[e], fy[e], dl0);
      x.raw = *reinterpret_cast<const uint4*>(dOt + swz(r0 + g + 8, 2 * t + cc));
      y.raw = *reinterpret_cast<const uint4*>(Ot + swz(r0 + g + 8, 2 * t + cc));
      x.unpack(fx); y.unpack(fy);
#pragma unroll
      for (int e = 0; e < 8; ++e) dl1 = fmaf(fx[e], fy[e], dl1);
    }
    dl0 = quad_sum(dl0); dl1 = quad_sum(dl1);
    const float nl0 = -(row_a < p.Lq ? p.LSE[lb + row_a] : 0.f) * LOG2E;
    const float nl1 = -(row_b < p.Lq ? p.LSE[lb + row_b] : 0.f) * LOG2E;
    float s[4][4], dp[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[nt][e] = 0.f; dp[nt][e] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4], a2[4], bb[4];
      ld_a(a, sq, r0, kk, lane);
      ld_a(a2, sdo, r0, kk, lane);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        ld_b_nk(bb, sk, np * 16, kk, lane);
        mma16816(s[2 * np], a, bb[0], bb[1]);
        mma16816(s[2 * np + 1], a, bb[2], bb[3]);
        ld_b_nk(bb, sv, np * 16, kk, lane);
        mma16816(dp[2 * np], a2, bb[0], bb[1]);
        mma16816(dp[2 * np + 1], a2, bb[2], bb[3]);
      }
    }
    const uint32_t rbase = uint32_t((b * p.H + h) * p.Lq);
    const uint32_t rk0 = DROP ? drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, rbase + row_a) : 0u;
    const uint32_t rk1 = DROP ? drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, rbase + row_b) : 0u;
    uint32_t pa[2][4], dsa[2][4];     // Pd (what multiplies dO in dV) and dS as packed bf16 blocks
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float pd[4], ds[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = nt * 8 + 2 * t + e;
        const bool valid = key < p.Lk;
        const float p0 = valid ? ex2_approx(fmaf(s[nt][e], c, nl0)) : 0.f;
        const float p1 = valid ? ex2_approx(fmaf(s[nt][2 + e], c, nl1)) : 0.f;
        float k0 = 1.f, k1 = 1.f;
        if (DROP) {
          const uint32_t cm = drop_col_attn(uint32_t(key));
          k0 = drop_keep_c(rk0, cm, p.drop_thr) ? p.inv_keep : 0.f;
          k1 = drop_keep_c(rk1, cm, p.drop_thr) ? p.inv_keep : 0.f;
        }
        pd[e] = p0 * k0;
        pd[2 + e] = p1 * k1;
        ds[e] = p0 * (dp[nt][e] * k0 - dl0) * p.scale;
        ds[2 + e] = p1 * (dp[nt][2 + e] * k1 - dl1) * p.scale;
      }
      pa[nt >> 1][(nt & 1) * 2] = pack_bf16(pd[0], pd[1]);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(pd[2], pd[3]);
      dsa[nt >> 1][(nt & 1) * 2] = pack_bf16(ds[0], ds[1]);
      dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
    }
    // dV += Pd^T dO, dK += dS^T Q  (A = transposed blocks, B = this warp's dO / Q rows read k-major)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t at[4] = {movm_t(pa[j][0]), movm_t(pa[j][2]), movm_t(pa[j][1]), movm_t(pa[j][3])};
      uint32_t dt[4] = {movm_t(dsa[j][0]), movm_t(dsa[j][2]), movm_t(dsa[j][1]), movm_t(dsa[j][3])};
#pragma unroll
      for (int d4 = 0; d4 < 4; ++d4) {
        uint32_t bb[4];
        ld_b_kn(bb, sdo, r0, d4, lane);
        mma16816(dv[j][2 * d4], at, bb[0], bb[1]);
        mma16816(dv[j][2 * d4 + 1], at, bb[2], bb[3]);
        ld_b_kn(bb, sq, r0, d4, lane);
        mma16816(dk[j][2 * d4], dt, bb[0], bb[1]);
        mma16816(dk[j][2 * d4 + 1], dt, bb[2], bb[3]);
      }
    }
    // dQ = dS K
    float dq[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int d4 = 0; d4 < 4; ++d4) {
        uint32_t bb[4];
        ld_b_kn(bb, sk, ks * 16, d4, lane);
        mma16816(dq[2 * d4], dsa[ks], bb[0], bb[1]);
        mma16816(dq[2 * d4 + 1], dsa[ks], bb[2], bb[3]);
      }
    __syncwarp();
    stage_tile(stg, r0, dq, 1.f, 1.f, lane);             // over this warp's own (consumed) Q rows
    __syncwarp();
    store_tile(stg, r0, p.dQ, p.lddq, qtok, tok0, p.Lq, h, lane, p.dbq ? &dbq0 : nullptr, p.dbq ? &dbq1 : nullptr);
  }
  cp_wait<0>();
  __syncthreads();                   // the ring is free: per-warp partials [4 warps][dV | dK][32][NKB_RSTRIDE] fp32, summed in a FIXED order
  float* red = reinterpret_cast<float*>(ring);          // (no atomics: dK / dV are bit-reproducible from run to run)
  {
    float* mine = red + warp * (2 * NARROW * NKB_RSTRIDE);
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int d = nt * 8 + 2 * t;
        *reinterpret_cast<float2*>(mine + (j * 16 + g) * NKB_RSTRIDE + d) = make_float2(dv[j][nt][0], dv[j][nt][1]);
        *reinterpret_cast<float2*>(mine + (j * 16 + g + 8) * NKB_RSTRIDE + d) = make_float2(dv[j][nt][2], dv[j][nt][3]);
        *reinterpret_cast<float2*>(mine + (NARROW + j * 16 + g) * NKB_RSTRIDE + d) = make_float2(dk[j][nt][0], dk[j][nt][1]);
        *reinterpret_cast<float2*>(mine + (NARROW + j * 16 + g + 8) * NKB_RSTRIDE + d) = make_float2(dk[j][nt][2], dk[j][nt][3]);
      }
  }
  if (p.dbq) { atomicAdd(&dbq_s[2 * lane], dbq0); atomicAdd(&dbq_s[2 * lane + 1], dbq1); }
  __syncthreads();
  const long long ktok = (long long)b * p.Lk;
  for (int i = tid; i < 2 * NARROW * 8; i += 128) {     // (tensor, key, 16-byte chunk): sum the four warps, round, store, keep the rounded
    const int which = i >> 8, key = (i >> 3) & 31, ch = i & 7;                                      // values for the column sums
    float* src = red + (which * NARROW + key) * NKB_RSTRIDE + ch * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = src[e];
#pragma unroll
    for (int w = 1; w < 4; ++w)
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += src[w * (2 * NARROW * NKB_RSTRIDE) + e];
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
#pragma unroll
    for (int e = 0; e < 8; ++e) src[e] = key < p.Lk ? bf16_round(v[e]) : 0.f;
    if (key >= p.Lk) continue;
    bf16* dst = which == 0 ? p.dV + (ktok + key) * p.lddv : p.dK + (ktok + key) * p.lddk;
    *reinterpret_cast<uint4*>(dst + h * HD + ch * 8) = o;
  }
  __syncthreads();
  if (tid < 64) {                                        // bias gradients: column sums of the stored (bf16) values
    if (p.dbq) atomicAdd(p.dbq + h * HD + tid, dbq_s[tid]);
    float sv_ = 0.f, sk_ = 0.f;
    for (int key = 0; key < p.Lk; ++key) {
      sv_ += red[key * NKB_RSTRIDE + tid];
      sk_ += red[(NARROW + key) * NKB_RSTRIDE + tid];
    }
    if (p.dbv) atomicAdd(p.dbv + h * HD + tid, sv_);
    if (p.dbk) atomicAdd(p.dbk + h * HD + tid, sk_);
  }
}

// ==============================================================================================================================
// NK backward, version 2 ("d-split", the default): the first version above keeps dK and dV (32 x 64 fp32 each) in every warp's
// registers -- 255 registers per thread, two CTAs (8 warps) per SM, 45 % issue utilisation under ncu because nothing hides the
// ldmatrix -> mma and exp latencies.  Here a tile is processed in two phases:
//   A  warp w owns query rows 16w..16w+15 of the 64-row tile (as before): S, dP, Pd, dS in registers, dQ = dS K stored at once,
//      and Pd / dS written as bf16 [64 rows x 32 keys] tiles to shared memory;
//   B  warp w owns the 16-wide slice d = 16w..16w+15 of dK / dV: dV[:, slice] += Pd^T dO[:, slice] and dK[:, slice] += dS^T Q[:, slice]
//      over ALL 64 rows of the tile (A operands = the shared Pd / dS tiles read transposed by ldmatrix, no movmatrix), 32 fp32
//      accumulator registers per thread instead of 128.
// Same number of HMMAs; ~150 registers -> three CTAs per SM with a 2-deep ring; no cross-warp reduction at the end (every warp owns
// its own columns of dK / dV, summed in a fixed order: bit-reproducible).
// ==============================================================================================================================
constexpr int NKB2_ST = 2;
constexpr int NKB2_PS = 64 * 64;                           // one [64 rows x 32 keys] bf16 tile (64-byte rows)
constexpr int NKB2_SMEM = 2 * NARROW * 128 + NKB2_ST * NKB_STAGE + 2 * NKB2_PS + 256;

__device__ __forceinline__ uint32_t swz64(int row, int chunk) {   // [rows x 64 B] tile, 16-byte chunks XOR-swizzled by (row / 2) % 4
  return uint32_t(row) * 64u + (uint32_t((chunk ^ (row >> 1)) & 3) << 4);
}

template <bool DROP>
__global__ void __launch_bounds__(128, 3) attn_nk_bwd2_kernel(const NarrowParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* Ks = smem;
  uint8_t* Vs = smem + NARROW * 128;
  uint8_t* ring = smem + 2 * NARROW * 128;
  uint8_t* Pt = ring + NKB2_ST * NKB_STAGE;                // Pd  [64 x 32] bf16
  uint8_t* St = Pt + NKB2_PS;                              // dS  [64 x 32] bf16
  float* dbq_s = reinterpret_cast<float*>(St + NKB2_PS);   // 64 floats
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int h = blockIdx.x, b = blockIdx.y;
  const uint32_t sk = smem_u32(Ks), sv = smem_u32(Vs), sring = smem_u32(ring), spt = smem_u32(Pt), sst = smem_u32(St);
  const int n_tiles = (p.Lq + 63) / 64;
  const long long qtok = (long long)b * p.Lq;
  auto issue = [&](int i) {
    if (i < n_tiles) {
      const uint32_t st = sring + (i % NKB2_ST) * NKB_STAGE;
      load_rows(st, p.Q, p.ldq, qtok, i * 64, p.Lq, 64, h, tid, 128);
      load_rows(st + NKB_TILE, p.dO, p.lddo, qtok, i * 64, p.Lq, 64, h, tid, 128);
      load_rows(st + 2 * NKB_TILE, p.O, p.ldo, qtok, i * 64, p.Lq, 64, h, tid, 128);
    }
    cp_commit();
  };
  load_rows(sk, p.K, p.ldk, (long long)b * p.Lk, 0, p.Lk, NARROW, h, tid, 128);
  load_rows(sv, p.V, p.ldv, (long long)b * p.Lk, 0, p.Lk, NARROW, h, tid, 128);
  issue(0);
  if (tid < 64) dbq_s[tid] = 0.f;
  float dv[2][2][4], dk[2][2][4];                          // [key m-tile][n-tile of this warp's d slice]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) { dv[mt][nt][e] = 0.f; dk[mt][nt][e] = 0.f; }
  float dbq0 = 0.f, dbq1 = 0.f;
  const float c = p.scale * LOG2E;
  const long long lb = ((long long)b * p.H + h) * p.Lq;
  const int r0 = warp * 16;
#pragma unroll 1
  for (int i = 0; i < n_tiles; ++i) {
    cp_wait<0>();
    __syncthreads();                 // tile i landed; every warp is done with phase B of tile i-1 (Pd / dS tiles and the other ring slot are free)
    issue(i + 1);
    uint8_t* stg = ring + (i % NKB2_ST) * NKB_STAGE;
    const uint32_t sq = sring + (i % NKB2_ST) * NKB_STAGE, sdo = sq + NKB_TILE;
    const uint8_t* dOt = stg + NKB_TILE;
    uint8_t* Ot = stg + 2 * NKB_TILE;
    // ---------------- phase A: this warp's 16 query rows (rows past Lq are zero-filled: they contribute exact zeros everywhere)
    const int tok0 = i * 64 + r0;
    const int row_a = tok0 + g, row_b = row_a + 8;
    float dl0 = 0.f, dl1 = 0.f;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      Vec16<bf16> x, y;
      float fx[8], fy[8];
      x.raw = *reinterpret_cast<const uint4*>(dOt + swz(r0 + g, 2 * t + cc));
      y.raw = *reinterpret_cast<const uint4*>(Ot + swz(r0 + g, 2 * t + cc));
      x.unpack(fx); y.unpack(fy);
#pragma unroll
      for (int e = 0; e < 8; ++e) dl0 = fmaf(fx[e], fy[e], dl0);
      x.raw = *reinterpret_cast<const uint4*>(dOt + swz(r0 + g + 8, 2 * t + cc));
      y.raw = *reinterpret_cast<const uint4*>(Ot + swz(r0 + g + 8, 2 * t + cc));
      x.unpack(fx); y.unpack(fy);
#pragma unroll
      for (int e = 0; e < 8; ++e) dl1 = fmaf(fx[e], fy[e], dl1);
    }
    dl0 = quad_sum(dl0); dl1 = quad_sum(dl1);
    const float nl0 = -(row_a < p.Lq ? p.LSE[lb + row_a] : 0.f) * LOG2E;
    const float nl1 = -(row_b < p.Lq ? p.LSE[lb + row_b] : 0.f) * LOG2E;
    uint32_t dsa[2][4];
    {
      float s[4][4], dp[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { s[nt][e] = 0.f; dp[nt][e] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4], a2[4], bb[4];
        ld_a(a, sq, r0, kk, lane);
        ld_a(a2, sdo, r0, kk, lane);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          ld_b_nk(bb, sk, np * 16, kk, lane);
          mma16816(s[2 * np], a, bb[0], bb[1]);
          mma16816(s[2 * np + 1], a, bb[2], bb[3]);
          ld_b_nk(bb, sv, np * 16, kk, lane);
          mma16816(dp[2 * np], a2, bb[0], bb[1]);
          mma16816(dp[2 * np + 1], a2, bb[2], bb[3]);
        }
      }
      const uint32_t rbase = uint32_t((b * p.H + h) * p.Lq);
      const uint32_t rk0 = DROP ? drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, rbase + row_a) : 0u;
      const uint32_t rk1 = DROP ? drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, rbase + row_b) : 0u;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float pd[4], ds[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = nt * 8 + 2 * t + e;
          const bool valid = key < p.Lk;
          const float p0 = valid ? ex2_approx(fmaf(s[nt][e], c, nl0)) : 0.f;
          const float p1 = valid ? ex2_approx(fmaf(s[nt][2 + e], c, nl1)) : 0.f;
          float k0 = 1.f, k1 = 1.f;
          if (DROP) {
            const uint32_t cm = drop_col_attn(uint32_t(key));
            k0 = drop_keep_c(rk0, cm, p.drop_thr) ? p.inv_keep : 0.f;
            k1 = drop_keep_c(rk1, cm, p.drop_thr) ? p.inv_keep : 0.f;
          }
          pd[e] = p0 * k0;
          pd[2 + e] = p1 * k1;
          ds[e] = p0 * (dp[nt][e] * k0 - dl0) * p.scale;
          ds[2 + e] = p1 * (dp[nt][2 + e] * k1 - dl1) * p.scale;
        }
        const uint32_t plo = pack_bf16(pd[0], pd[1]), phi = pack_bf16(pd[2], pd[3]);
        const uint32_t slo = pack_bf16(ds[0], ds[1]), shi = pack_bf16(ds[2], ds[3]);
        dsa[nt >> 1][(nt & 1) * 2] = slo;
        dsa[nt >> 1][(nt & 1) * 2 + 1] = shi;
        // [row][key] bf16 tiles for phase B: this thread's two keys (4 bytes) of rows r0+g and r0+g+8
        *reinterpret_cast<uint32_t*>(Pt + swz64(r0 + g, nt) + t * 4) = plo;
        *reinterpret_cast<uint32_t*>(Pt + swz64(r0 + g + 8, nt) + t * 4) = phi;
        *reinterpret_cast<uint32_t*>(St + swz64(r0 + g, nt) + t * 4) = slo;
        *reinterpret_cast<uint32_t*>(St + swz64(r0 + g + 8, nt) + t * 4) = shi;
      }
    }
    {                                // dQ = dS K for this warp's rows
      float dq[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) dq[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int d4 = 0; d4 < 4; ++d4) {
          uint32_t bb[4];
          ld_b_kn(bb, sk, ks * 16, d4, lane);
          mma16816(dq[2 * d4], dsa[ks], bb[0], bb[1]);
          mma16816(dq[2 * d4 + 1], dsa[ks], bb[2], bb[3]);
        }
      __syncwarp();
      stage_tile(Ot, r0, dq, 1.f, 1.f, lane);             // over this warp's own (consumed) O rows: phase B reads Q and dO of ALL rows
      __syncwarp();
      if (tok0 < p.Lq)
        store_tile(Ot, r0, p.dQ, p.lddq, qtok, tok0, p.Lq, h, lane, p.dbq ? &dbq0 : nullptr, p.dbq ? &dbq1 : nullptr);
    }
    __syncthreads();                 // Pd / dS of all 64 rows are in shared memory
    // ---------------- phase B: this warp's 16-wide slice of d, all 64 rows of the tile
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t bo[4], bq[4];
      ld_b_kn(bo, sdo, ks * 16, warp, lane);              // dO[k = rows 16ks.., n = d slice]: two n-tiles
      ld_b_kn(bq, sq, ks * 16, warp, lane);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t ap[4], as[4];
        const int mi = lane >> 3, r = lane & 7;
        const int row = ks * 16 + (mi >> 1) * 8 + r, chunk = mt * 2 + (mi & 1);
        ldsm4t(ap, spt + swz64(row, chunk));              // A = Pd^T (m = keys 16mt.., k = rows 16ks..)
        ldsm4t(as, sst + swz64(row, chunk));
        mma16816(dv[mt][0], ap, bo[0], bo[1]);
        mma16816(dv[mt][1], ap, bo[2], bo[3]);
        mma16816(dk[mt][0], as, bq[0], bq[1]);
        mma16816(dk[mt][1], as, bq[2], bq[3]);
      }
    }
  }
  cp_wait<0>();
  if (p.dbq) { atomicAdd(&dbq_s[2 * lane], dbq0); atomicAdd(&dbq_s[2 * lane + 1], dbq1); }
  // dV / dK: warp w owns columns 16w..16w+15 of both; rows = keys
  const long long ktok = (long long)b * p.Lk;
  float csv[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, csk[2][2] = {{0.f, 0.f}, {0.f, 0.f}};     // column sums of the stored (rounded) values
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int col = h * HD + warp * 16 + nt * 8 + 2 * t;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int key = mt * 16 + g + hf * 8;
        if (key < p.Lk) {
          const uint32_t wv = pack_bf16(dv[mt][nt][2 * hf], dv[mt][nt][2 * hf + 1]);
          const uint32_t wk = pack_bf16(dk[mt][nt][2 * hf], dk[mt][nt][2 * hf + 1]);
          *reinterpret_cast<uint32_t*>(p.dV + (ktok + key) * p.lddv + col) = wv;
          *reinterpret_cast<uint32_t*>(p.dK + (ktok + key) * p.lddk + col) = wk;
          csv[nt][0] += __uint_as_float(wv << 16); csv[nt][1] += __uint_as_float(wv & 0xFFFF0000u);
          csk[nt][0] += __uint_as_float(wk << 16); csk[nt][1] += __uint_as_float(wk & 0xFFFF0000u);
        }
      }
    }
  if (p.dbv || p.dbk) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {                 // over the 8 row groups g (lanes with equal t)
          csv[nt][e] += __shfl_xor_sync(0xffffffffu, csv[nt][e], o);
          csk[nt][e] += __shfl_xor_sync(0xffffffffu, csk[nt][e], o);
        }
        if (g == 0) {
          const int col = h * HD + warp * 16 + nt * 8 + 2 * t + e;
          if (p.dbv) atomicAdd(p.dbv + col, csv[nt][e]);
          if (p.dbk) atomicAdd(p.dbk + col, csk[nt][e]);
        }
      }
  }
  __syncthreads();
  if (p.dbq && tid < 64) atomicAdd(p.dbq + h * HD + tid, dbq_s[tid]);
}

// ==============================================================================================================================
// NQ forward: Lq <= 32.  grid (H, B), 128 threads; 64-key tiles of {K, V} through a 4-deep ring, warp w owns keys 16w..16w+15 of
// every tile and keeps its own running (max, sum, O[32 x 64]); the four partials are merged at the end.
// ==============================================================================================================================
constexpr int NQF_ST = 4;
constexpr int NQ_STAGE = 2 * 64 * 128;                    // K, V
constexpr int NQ_OSTRIDE = HD + 8;                        // fp32 row stride of the merge buffers (bank spread)
constexpr int NQF_SMEM = NARROW * 128 + NQF_ST * NQ_STAGE + 256;
static_assert(4 * NARROW * NQ_OSTRIDE * 4 + 2 * 4 * NARROW * 4 + NARROW * HD * 4 <= NQF_ST * NQ_STAGE, "merge buffers (+ the pooled-output tile) must fit in the ring");
static_assert(4 * NARROW * NQ_OSTRIDE * 4 <= 3 * NQ_STAGE, "the backward's per-warp dQ partials must fit in its ring");

template <bool DROP>
__global__ void __launch_bounds__(128, 2) attn_nq_fwd_kernel(const NarrowParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* Qs = smem;
  uint8_t* ring = smem + NARROW * 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int h = blockIdx.x, b = blockIdx.y;
  const uint32_t sq = smem_u32(Qs), sring = smem_u32(ring);
  const int n_tiles = (p.Lk + 63) / 64;
  const long long ktok = (long long)b * p.Lk;
  auto issue = [&](int i) {
    if (i < n_tiles) {
      const uint32_t st = sring + (i % NQF_ST) * NQ_STAGE;
      load_rows(st, p.K, p.ldk, ktok, i * 64, p.Lk, 64, h, tid, 128);
      load_rows(st + 64 * 128, p.V, p.ldv, ktok, i * 64, p.Lk, 64, h, tid, 128);
    }
    cp_commit();
  };
  load_rows(sq, p.Q, p.ldq, (long long)b * p.Lq, 0, p.Lq, NARROW, h, tid, 128);
  issue(0); issue(1); issue(2);
  const float c = p.scale * LOG2E;
  float o[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[mt][nt][e] = 0.f;
  float m[2][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}}, l[2][2] = {{0.f, 0.f}, {0.f, 0.f}};   // [m-tile][row g / g+8]
  uint32_t rk[2][2] = {{0u, 0u}, {0u, 0u}};
  if (DROP) {
    const uint32_t rbase = uint32_t((b * p.H + h) * p.Lq);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) rk[mt][hf] = drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, rbase + mt * 16 + hf * 8 + g);
  }
#pragma unroll 1
  for (int i = 0; i < n_tiles; ++i) {
    cp_wait<2>();
    __syncthreads();
    issue(i + 3);
    const uint32_t skt = sring + (i % NQF_ST) * NQ_STAGE, svt = skt + 64 * 128;
    const int kw = warp * 16;                             // this warp's keys inside the tile
    const int key0 = i * 64 + kw;
    if (key0 >= p.Lk) continue;
    float s[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[mt][nt][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t bb[4], a[4];
      ld_b_nk(bb, skt, kw, kk, lane);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        ld_a(a, sq, mt * 16, kk, lane);
        mma16816(s[mt][0], a, bb[0], bb[1]);
        mma16816(s[mt][1], a, bb[2], bb[3]);
      }
    }
    uint32_t pa[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (key0 + nt * 8 + 2 * t + e >= p.Lk) { s[mt][nt][e] = -INFINITY; s[mt][nt][2 + e] = -INFINITY; }
          tm0 = fmaxf(tm0, s[mt][nt][e]);
          tm1 = fmaxf(tm1, s[mt][nt][2 + e]);
        }
      tm0 = quad_max(tm0); tm1 = quad_max(tm1);            // key0 < Lk: at least one valid key, the maxima are finite
      const float mn0 = fmaxf(m[mt][0], tm0), mn1 = fmaxf(m[mt][1], tm1);
      const float al0 = ex2_approx((m[mt][0] - mn0) * c), al1 = ex2_approx((m[mt][1] - mn1) * c);     // first tile: ex2(-inf) = 0
      m[mt][0] = mn0; m[mt][1] = mn1;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        float pv[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          pv[e] = ex2_approx((s[mt][nt][e] - mn0) * c);
          pv[2 + e] = ex2_approx((s[mt][nt][2 + e] - mn1) * c);
        }
        ps0 += pv[0] + pv[1];
        ps1 += pv[2] + pv[3];
        if (DROP) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            // drop_col_attn(key0 + ...): key0 = 64 i + kw with kw + nt * 8 + 2 t + e < 64, so the 64-key group term is i * MulHi (hoisted per
            // tile) and the in-group term needs no shift / mask per element
            const uint32_t cm = (uint32_t(i) * kDropColMulHi) ^ (uint32_t(kw + nt * 8 + 2 * t + e) * kDropColMul);
            if (!drop_keep_c(rk[mt][0], cm, p.drop_thr)) pv[e] = 0.f;
            if (!drop_keep_c(rk[mt][1], cm, p.drop_thr)) pv[2 + e] = 0.f;
          }
        }
        pa[mt][nt * 2] = pack_bf16(pv[0], pv[1]);
        pa[mt][nt * 2 + 1] = pack_bf16(pv[2], pv[3]);
      }
      l[mt][0] = l[mt][0] * al0 + ps0;                     // per-thread partial sums; reduced over the quad at the end
      l[mt][1] = l[mt][1] * al1 + ps1;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        o[mt][nt][0] *= al0; o[mt][nt][1] *= al0;
        o[mt][nt][2] *= al1; o[mt][nt][3] *= al1;
      }
    }
#pragma unroll
    for (int d4 = 0; d4 < 4; ++d4) {
      uint32_t bb[4];
      ld_b_kn(bb, svt, kw, d4, lane);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        mma16816(o[mt][2 * d4], pa[mt], bb[0], bb[1]);
        mma16816(o[mt][2 * d4 + 1], pa[mt], bb[2], bb[3]);
      }
    }
  }
  cp_wait<0>();
  __syncthreads();                   // ring free: merge buffers  O_w [4][32][NQ_OSTRIDE] fp32, m_w [4][32], l_w [4][32]
  float* Ow = reinterpret_cast<float*>(ring);
  float* mw = Ow + 4 * NARROW * NQ_OSTRIDE;
  float* lw = mw + 4 * NARROW;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const float l0 = quad_sum(l[mt][0]), l1 = quad_sum(l[mt][1]);
    const int ra = mt * 16 + g, rb = ra + 8;
    if (t == 0) {
      mw[warp * NARROW + ra] = m[mt][0]; lw[warp * NARROW + ra] = l0;
      mw[warp * NARROW + rb] = m[mt][1]; lw[warp * NARROW + rb] = l1;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<float2*>(Ow + (warp * NARROW + ra) * NQ_OSTRIDE + nt * 8 + 2 * t) = make_float2(o[mt][nt][0], o[mt][nt][1]);
      *reinterpret_cast<float2*>(Ow + (warp * NARROW + rb) * NQ_OSTRIDE + nt * 8 + 2 * t) = make_float2(o[mt][nt][2], o[mt][nt][3]);
    }
  }
  __syncthreads();
  {
    const int row = tid >> 2, c0 = (tid & 3) * 16;         // 32 rows x 4 column quarters
    if (row < p.Lq) {
      float mm = -INFINITY;
#pragma unroll
      for (int w = 0; w < 4; ++w) mm = fmaxf(mm, mw[w * NARROW + row]);
      float f[4], lt = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float mv = mw[w * NARROW + row];
        f[w] = mv == -INFINITY ? 0.f : ex2_approx((mv - mm) * c);   // a warp that saw no valid key (short Lk) contributes nothing
        lt += f[w] * lw[w * NARROW + row];
      }
      const float inv = (DROP ? p.inv_keep : 1.f) / lt;
      float acc[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float* src = Ow + (w * NARROW + row) * NQ_OSTRIDE + c0;
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = fmaf(f[w], src[e], acc[e]);
      }
      uint4 o0, o1;
      o0.x = pack_bf16(acc[0] * inv, acc[1] * inv); o0.y = pack_bf16(acc[2] * inv, acc[3] * inv);
      o0.z = pack_bf16(acc[4] * inv, acc[5] * inv); o0.w = pack_bf16(acc[6] * inv, acc[7] * inv);
      o1.x = pack_bf16(acc[8] * inv, acc[9] * inv); o1.y = pack_bf16(acc[10] * inv, acc[11] * inv);
      o1.z = pack_bf16(acc[12] * inv, acc[13] * inv); o1.w = pack_bf16(acc[14] * inv, acc[15] * inv);
      bf16* dst = p.O + ((long long)b * p.Lq + row) * p.ldo + h * HD + c0;
      *reinterpret_cast<uint4*>(dst) = o0;
      *reinterpret_cast<uint4*>(dst + 8) = o1;
      if ((tid & 3) == 0) p.LSE[((long long)b * p.H + h) * p.Lq + row] = mm * p.scale + logf(lt);
      if (p.pool) {                                        // the rounded values that were stored -> a [32][64] fp32 tile behind the merge buffers
        float* pf = lw + 4 * NARROW + row * HD + c0;
#pragma unroll
        for (int e = 0; e < 16; ++e) pf[e] = bf16_round(acc[e] * inv);
      }
    }
  }
  if (p.pool) {                                            // column sums over the (<= 32) query rows in a fixed order: one part per (b, h)
    __syncthreads();
    if (tid < HD) {
      const float* pf = lw + 4 * NARROW + tid;
      float s_ = 0.f;
      for (int row = 0; row < p.Lq; ++row) s_ += pf[row * HD];
      p.pool[((long long)b * p.H + h) * HD + tid] = s_;
    }
  }
}

// ==============================================================================================================================
// NQ backward: Lq <= 32.  grid (H, B), 128 threads; Q, dO, O resident; 64-key tiles of {K, V} through a 3-deep ring, warp w owns
// keys 16w..16w+15 of every tile.  Per 16 keys: S^T = K Q^T, dP^T = V dO^T, dS^T, dV = Pd^T dO and dK = dS^T Q (complete: stored
// at once, staged over the warp's own K / V rows), dQ += dS K (reduced over warps and tiles at the end).
// ==============================================================================================================================
constexpr int NQB_ST = 3;
constexpr int NQB_FIXED = 3 * NARROW * 128;               // Q, dO, O
constexpr int NQB_RED = NARROW * NQ_OSTRIDE * 4;          // dQ reduction buffer (fp32)
constexpr int NQB_SMEM = NQB_FIXED + NQB_ST * NQ_STAGE + NQB_RED + 2 * NARROW * 4 + 3 * HD * 4 + 128;

template <bool DROP>
__global__ void __launch_bounds__(128, 2) attn_nq_bwd_kernel(const NarrowParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* Qs = smem;
  uint8_t* dOs = smem + NARROW * 128;
  uint8_t* Os = smem + 2 * NARROW * 128;
  uint8_t* ring = smem + NQB_FIXED;
  float* red = reinterpret_cast<float*>(ring + NQB_ST * NQ_STAGE);       // [32][NQ_OSTRIDE]
  float* nl2s = red + NARROW * NQ_OSTRIDE;                               // -lse * log2e   per query
  float* ndls = nl2s + NARROW;                                           // -delta * scale per query
  float* dbs = ndls + NARROW;                                            // [3][64]: dbq, dbk, dbv partial sums
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int h = blockIdx.x, b = blockIdx.y;
  const uint32_t sq = smem_u32(Qs), sdo = smem_u32(dOs), sring = smem_u32(ring);
  const int n_tiles = (p.Lk + 63) / 64;
  const long long ktok = (long long)b * p.Lk, qtok = (long long)b * p.Lq;
  auto issue = [&](int i) {
    if (i < n_tiles) {
      const uint32_t st = sring + (i % NQB_ST) * NQ_STAGE;
      load_rows(st, p.K, p.ldk, ktok, i * 64, p.Lk, 64, h, tid, 128);
      load_rows(st + 64 * 128, p.V, p.ldv, ktok, i * 64, p.Lk, 64, h, tid, 128);
    }
    cp_commit();
  };
  load_rows(sq, p.Q, p.ldq, qtok, 0, p.Lq, NARROW, h, tid, 128);
  load_rows(sdo, p.dO, p.lddo, qtok, 0, p.Lq, NARROW, h, tid, 128);
  load_rows(smem_u32(Os), p.O, p.ldo, qtok, 0, p.Lq, NARROW, h, tid, 128);
  issue(0);
  issue(1);
  for (int i = tid; i < 3 * HD; i += 128) dbs[i] = 0.f;
  cp_wait<1>();                      // the resident tiles + K/V tile 0 (one group may still be in flight)
  __syncthreads();
  {                                  // per-query statistics: thread -> (row = tid / 4, 16 of its 64 columns)
    const int row = tid >> 2, cq = (tid & 3) * 2;
    float dl = 0.f;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      Vec16<bf16> x, y;
      float fx[8], fy[8];
      x.raw = *reinterpret_cast<const uint4*>(dOs + swz(row, cq + cc));
      y.raw = *reinterpret_cast<const uint4*>(Os + swz(row, cq + cc));
      x.unpack(fx); y.unpack(fy);
#pragma unroll
      for (int e = 0; e < 8; ++e) dl = fmaf(fx[e], fy[e], dl);
    }
    dl = quad_sum(dl);
    if ((tid & 3) == 0) {
      ndls[row] = -dl * p.scale;
      nl2s[row] = -(row < p.Lq ? p.LSE[((long long)b * p.H + h) * p.Lq + row] : 0.f) * LOG2E;
    }
  }
  __syncthreads();
  // this thread's 8 query columns: q = nt*8 + 2t + e
  float nl2[4][2], ndl[4][2];
  uint32_t rk[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int q = nt * 8 + 2 * t + e;
      nl2[nt][e] = nl2s[q];
      ndl[nt][e] = ndls[q];
      rk[nt][e] = DROP ? drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, uint32_t((b * p.H + h) * p.Lq + q)) : 0u;
    }
  float dq[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[mt][nt][e] = 0.f;
  float dbv0 = 0.f, dbv1 = 0.f, dbk0 = 0.f, dbk1 = 0.f;
  const float c = p.scale * LOG2E;
  const float sck = p.scale * (DROP ? p.inv_keep : 1.f);
#pragma unroll 1
  for (int i = 0; i < n_tiles; ++i) {
    cp_wait<1>();
    __syncthreads();
    issue(i + 2);
    uint8_t* Kt = ring + (i % NQB_ST) * NQ_STAGE;
    uint8_t* Vt = Kt + 64 * 128;
    const uint32_t skt = sring + (i % NQB_ST) * NQ_STAGE, svt = skt + 64 * 128;
    const int kw = warp * 16;
    const int key0 = i * 64 + kw;
    if (key0 >= p.Lk) continue;
    float st[4][4], dpt[4][4];        // S^T, dP^T: rows = this warp's 16 keys, columns = the 32 queries
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) { st[nt][e] = 0.f; dpt[nt][e] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4], a2[4], bb[4];
      ld_a(a, skt, kw, kk, lane);
      ld_a(a2, svt, kw, kk, lane);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        ld_b_nk(bb, sq, np * 16, kk, lane);
        mma16816(st[2 * np], a, bb[0], bb[1]);
        mma16816(st[2 * np + 1], a, bb[2], bb[3]);
        ld_b_nk(bb, sdo, np * 16, kk, lane);
        mma16816(dpt[2 * np], a2, bb[0], bb[1]);
        mma16816(dpt[2 * np + 1], a2, bb[2], bb[3]);
      }
    }
    const bool va = key0 + g < p.Lk, vb = key0 + g + 8 < p.Lk;
    const uint32_t cma = drop_col_attn(uint32_t(key0 + g)), cmb = drop_col_attn(uint32_t(key0 + g + 8));
    uint32_t pa[2][4], dsa[2][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float pd[4], ds[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float p0 = va ? ex2_approx(fmaf(st[nt][e], c, nl2[nt][e])) : 0.f;
        const float p1 = vb ? ex2_approx(fmaf(st[nt][2 + e], c, nl2[nt][e])) : 0.f;
        bool k0 = true, k1 = true;
        if (DROP) {
          k0 = drop_keep_c(rk[nt][e], cma, p.drop_thr);
          k1 = drop_keep_c(rk[nt][e], cmb, p.drop_thr);
        }
        pd[e] = k0 ? p0 * (DROP ? p.inv_keep : 1.f) : 0.f;
        pd[2 + e] = k1 ? p1 * (DROP ? p.inv_keep : 1.f) : 0.f;
        ds[e] = p0 * fmaf(dpt[nt][e], k0 ? sck : 0.f, ndl[nt][e]);
        ds[2 + e] = p1 * fmaf(dpt[nt][2 + e], k1 ? sck : 0.f, ndl[nt][e]);
      }
      pa[nt >> 1][(nt & 1) * 2] = pack_bf16(pd[0], pd[1]);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(pd[2], pd[3]);
      dsa[nt >> 1][(nt & 1) * 2] = pack_bf16(ds[0], ds[1]);
      dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
    }
    // dQ += dS K   (A = transposed dS^T blocks, B = this warp's K rows read k-major) -- before the K rows are overwritten below
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      uint32_t at[4] = {movm_t(dsa[mt][0]), movm_t(dsa[mt][2]), movm_t(dsa[mt][1]), movm_t(dsa[mt][3])};
#pragma unroll
      for (int d4 = 0; d4 < 4; ++d4) {
        uint32_t bb[4];
        ld_b_kn(bb, skt, kw, d4, lane);
        mma16816(dq[mt][2 * d4], at, bb[0], bb[1]);
        mma16816(dq[mt][2 * d4 + 1], at, bb[2], bb[3]);
      }
    }
    float acc[8][4];
    // dV = Pd^T dO over all (<= 32) queries: complete for these 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int d4 = 0; d4 < 4; ++d4) {
        uint32_t bb[4];
        ld_b_kn(bb, sdo, ks * 16, d4, lane);
        mma16816(acc[2 * d4], pa[ks], bb[0], bb[1]);
        mma16816(acc[2 * d4 + 1], pa[ks], bb[2], bb[3]);
      }
    __syncwarp();
    stage_tile(Vt, kw, acc, 1.f, 1.f, lane);             // over this warp's own (consumed) V rows
    __syncwarp();
    store_tile(Vt, kw, p.dV, p.lddv, ktok, key0, p.Lk, h, lane, p.dbv ? &dbv0 : nullptr, p.dbv ? &dbv1 : nullptr);
    // dK = dS^T Q
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int d4 = 0; d4 < 4; ++d4) {
        uint32_t bb[4];
        ld_b_kn(bb, sq, ks * 16, d4, lane);
        mma16816(acc[2 * d4], dsa[ks], bb[0], bb[1]);
        mma16816(acc[2 * d4 + 1], dsa[ks], bb[2], bb[3]);
      }
    __syncwarp();
    stage_tile(Kt, kw, acc, 1.f, 1.f, lane);             // over this warp's own K rows (the dQ product above has read them)
    __syncwarp();
    store_tile(Kt, kw, p.dK, p.lddk, ktok, key0, p.Lk, h, lane, p.dbk ? &dbk0 : nullptr, p.dbk ? &dbk1 : nullptr);
  }
  cp_wait<0>();
  __syncthreads();                   // the ring is free: per-warp dQ partials [4][32][NQ_OSTRIDE] fp32, summed in a FIXED order (no atomics:
  float* part = reinterpret_cast<float*>(ring);          // dQ is bit-reproducible from run to run)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<float2*>(part + (warp * NARROW + mt * 16 + g) * NQ_OSTRIDE + nt * 8 + 2 * t) = make_float2(dq[mt][nt][0], dq[mt][nt][1]);
      *reinterpret_cast<float2*>(part + (warp * NARROW + mt * 16 + g + 8) * NQ_OSTRIDE + nt * 8 + 2 * t) = make_float2(dq[mt][nt][2], dq[mt][nt][3]);
    }
  if (p.dbv) { atomicAdd(&dbs[2 * HD + 2 * lane], dbv0); atomicAdd(&dbs[2 * HD + 2 * lane + 1], dbv1); }
  if (p.dbk) { atomicAdd(&dbs[HD + 2 * lane], dbk0); atomicAdd(&dbs[HD + 2 * lane + 1], dbk1); }
  __syncthreads();
  {
    const int row = tid >> 2, c0 = (tid & 3) * 16;
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = part[row * NQ_OSTRIDE + c0 + e];
#pragma unroll
    for (int w = 1; w < 4; ++w)
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] += part[(w * NARROW + row) * NQ_OSTRIDE + c0 + e];
#pragma unroll
    for (int e = 0; e < 16; ++e) red[row * NQ_OSTRIDE + c0 + e] = row < p.Lq ? bf16_round(v[e]) : 0.f;
    if (row < p.Lq) {
      uint4 o0, o1;
      o0.x = pack_bf16(v[0], v[1]); o0.y = pack_bf16(v[2], v[3]); o0.z = pack_bf16(v[4], v[5]); o0.w = pack_bf16(v[6], v[7]);
      o1.x = pack_bf16(v[8], v[9]); o1.y = pack_bf16(v[10], v[11]); o1.z = pack_bf16(v[12], v[13]); o1.w = pack_bf16(v[14], v[15]);
      bf16* dst = p.dQ + (qtok + row) * p.lddq + h * HD + c0;
      *reinterpret_cast<uint4*>(dst) = o0;
      *reinterpret_cast<uint4*>(dst + 8) = o1;
    }
  }
  __syncthreads();
  if (tid < 64) {
    if (p.dbq) {
      float s_ = 0.f;
      for (int row = 0; row < p.Lq; ++row) s_ += red[row * NQ_OSTRIDE + tid];
      atomicAdd(p.dbq + h * HD + tid, s_);
    }
    if (p.dbk) atomicAdd(p.dbk + h * HD + tid, dbs[HD + tid]);
    if (p.dbv) atomicAdd(p.dbv + h * HD + tid, dbs[2 * HD + tid]);
  }
}

// ==============================================================================================================================
// NQ backward, version 2 ("d-split", the default; same idea as attn_nk_bwd2_kernel): phase A -- warp w owns keys 16w..16w+15 of the
// 64-key tile: S^T, dP^T, Pd^T, dS^T in registers, dV and dK of its keys stored at once (both staged over the warp's own V rows),
// dS^T written as a bf16 [64 keys x 32 queries] tile to shared memory; phase B -- warp w owns the slice d = 16w..16w+15 of dQ:
// dQ[:, slice] += dS K[:, slice] over ALL 64 keys of the tile (A = the shared dS^T tile read transposed).  16 instead of 64
// accumulator registers for dQ, three CTAs per SM, no cross-warp reduction at the end.
// ==============================================================================================================================
constexpr int NQB2_SMEM = NQB_FIXED + NQB_ST * NQ_STAGE + NKB2_PS + 2 * NARROW * 4 + 3 * HD * 4 + 128;

template <bool DROP>
__global__ void __launch_bounds__(128, 3) attn_nq_bwd2_kernel(const NarrowParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* Qs = smem;
  uint8_t* dOs = smem + NARROW * 128;
  uint8_t* Os = smem + 2 * NARROW * 128;
  uint8_t* ring = smem + NQB_FIXED;
  uint8_t* DSt = ring + NQB_ST * NQ_STAGE;                               // dS^T [64 keys x 32 queries] bf16 (64-byte rows)
  float* nl2s = reinterpret_cast<float*>(DSt + NKB2_PS);                 // -lse * log2e   per query
  float* ndls = nl2s + NARROW;                                           // -delta * scale per query
  float* dbs = ndls + NARROW;                                            // [3][64]: (unused), dbk, dbv partial sums
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int h = blockIdx.x, b = blockIdx.y;
  const uint32_t sq = smem_u32(Qs), sdo = smem_u32(dOs), sring = smem_u32(ring), sds = smem_u32(DSt);
  const int n_tiles = (p.Lk + 63) / 64;
  const long long ktok = (long long)b * p.Lk, qtok = (long long)b * p.Lq;
  auto issue = [&](int i) {
    if (i < n_tiles) {
      const uint32_t st = sring + (i % NQB_ST) * NQ_STAGE;
      load_rows(st, p.K, p.ldk, ktok, i * 64, p.Lk, 64, h, tid, 128);
      load_rows(st + 64 * 128, p.V, p.ldv, ktok, i * 64, p.Lk, 64, h, tid, 128);
    }
    cp_commit();
  };
  load_rows(sq, p.Q, p.ldq, qtok, 0, p.Lq, NARROW, h, tid, 128);
  load_rows(sdo, p.dO, p.lddo, qtok, 0, p.Lq, NARROW, h, tid, 128);
  load_rows(smem_u32(Os), p.O, p.ldo, qtok, 0, p.Lq, NARROW, h, tid, 128);
  issue(0);
  issue(1);
  for (int i = tid; i < 3 * HD; i += 128) dbs[i] = 0.f;
  cp_wait<1>();
  __syncthreads();
  {                                  // per-query statistics: thread -> (row = tid / 4, 16 of its 64 columns)
    const int row = tid >> 2, cq = (tid & 3) * 2;
    float dl = 0.f;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      Vec16<bf16> x, y;
      float fx[8], fy[8];
      x.raw = *reinterpret_cast<const uint4*>(dOs + swz(row, cq + cc));
      y.raw = *reinterpret_cast<const uint4*>(Os + swz(row, cq + cc));
      x.unpack(fx); y.unpack(fy);
#pragma unroll
      for (int e = 0; e < 8; ++e) dl = fmaf(fx[e], fy[e], dl);
    }
    dl = quad_sum(dl);
    if ((tid & 3) == 0) {
      ndls[row] = -dl * p.scale;
      nl2s[row] = -(row < p.Lq ? p.LSE[((long long)b * p.H + h) * p.Lq + row] : 0.f) * LOG2E;
    }
  }
  __syncthreads();
  float nl2[4][2], ndl[4][2];
  uint32_t rk[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int q = nt * 8 + 2 * t + e;
      nl2[nt][e] = nl2s[q];
      ndl[nt][e] = ndls[q];
      rk[nt][e] = DROP ? drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, uint32_t((b * p.H + h) * p.Lq + q)) : 0u;
    }
  float dq[2][2][4];                                       // [query m-tile][n-tile of this warp's d slice]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[mt][nt][e] = 0.f;
  float dbv0 = 0.f, dbv1 = 0.f, dbk0 = 0.f, dbk1 = 0.f;
  const float c = p.scale * LOG2E;
  const float sck = p.scale * (DROP ? p.inv_keep : 1.f);
  const int kw = warp * 16;
#pragma unroll 1
  for (int i = 0; i < n_tiles; ++i) {
    cp_wait<1>();
    __syncthreads();                 // tile i landed; every warp is done with phase B of tile i-1 (the dS^T tile and a ring slot are free)
    issue(i + 2);
    uint8_t* Vt = ring + (i % NQB_ST) * NQ_STAGE + 64 * 128;
    const uint32_t skt = sring + (i % NQB_ST) * NQ_STAGE, svt = skt + 64 * 128;
    const int key0 = i * 64 + kw;
    // ---------------- phase A: this warp's 16 keys (keys past Lk are zero-filled and forced to P = 0: exact zeros everywhere)
    uint32_t pa[2][4], dsa[2][4];
    {
      float st[4][4], dpt[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { st[nt][e] = 0.f; dpt[nt][e] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4], a2[4], bb[4];
        ld_a(a, skt, kw, kk, lane);
        ld_a(a2, svt, kw, kk, lane);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          ld_b_nk(bb, sq, np * 16, kk, lane);
          mma16816(st[2 * np], a, bb[0], bb[1]);
          mma16816(st[2 * np + 1], a, bb[2], bb[3]);
          ld_b_nk(bb, sdo, np * 16, kk, lane);
          mma16816(dpt[2 * np], a2, bb[0], bb[1]);
          mma16816(dpt[2 * np + 1], a2, bb[2], bb[3]);
        }
      }
      const bool va = key0 + g < p.Lk, vb = key0 + g + 8 < p.Lk;
      const uint32_t cma = drop_col_attn(uint32_t(key0 + g)), cmb = drop_col_attn(uint32_t(key0 + g + 8));
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float pd[4], ds[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float p0 = va ? ex2_approx(fmaf(st[nt][e], c, nl2[nt][e])) : 0.f;
          const float p1 = vb ? ex2_approx(fmaf(st[nt][2 + e], c, nl2[nt][e])) : 0.f;
          bool k0 = true, k1 = true;
          if (DROP) {
            k0 = drop_keep_c(rk[nt][e], cma, p.drop_thr);
            k1 = drop_keep_c(rk[nt][e], cmb, p.drop_thr);
          }
          pd[e] = k0 ? p0 * (DROP ? p.inv_keep : 1.f) : 0.f;
          pd[2 + e] = k1 ? p1 * (DROP ? p.inv_keep : 1.f) : 0.f;
          ds[e] = p0 * fmaf(dpt[nt][e], k0 ? sck : 0.f, ndl[nt][e]);
          ds[2 + e] = p1 * fmaf(dpt[nt][2 + e], k1 ? sck : 0.f, ndl[nt][e]);
        }
        pa[nt >> 1][(nt & 1) * 2] = pack_bf16(pd[0], pd[1]);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(pd[2], pd[3]);
        const uint32_t slo = pack_bf16(ds[0], ds[1]), shi = pack_bf16(ds[2], ds[3]);
        dsa[nt >> 1][(nt & 1) * 2] = slo;
        dsa[nt >> 1][(nt & 1) * 2 + 1] = shi;
        // dS^T tile for phase B: this thread's two queries (4 bytes) of key rows kw+g and kw+g+8
        *reinterpret_cast<uint32_t*>(DSt + swz64(kw + g, nt) + t * 4) = slo;
        *reinterpret_cast<uint32_t*>(DSt + swz64(kw + g + 8, nt) + t * 4) = shi;
      }
    }
    {
      float acc[8][4];
      // dV = Pd^T dO over all (<= 32) queries: complete for these 16 keys
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int d4 = 0; d4 < 4; ++d4) {
          uint32_t bb[4];
          ld_b_kn(bb, sdo, ks * 16, d4, lane);
          mma16816(acc[2 * d4], pa[ks], bb[0], bb[1]);
          mma16816(acc[2 * d4 + 1], pa[ks], bb[2], bb[3]);
        }
      __syncwarp();
      stage_tile(Vt, kw, acc, 1.f, 1.f, lane);             // over this warp's own (consumed) V rows; the K rows stay intact for phase B
      __syncwarp();
      store_tile(Vt, kw, p.dV, p.lddv, ktok, key0, p.Lk, h, lane, p.dbv ? &dbv0 : nullptr, p.dbv ? &dbv1 : nullptr);
      // dK = dS^T Q
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int d4 = 0; d4 < 4; ++d4) {
          uint32_t bb[4];
          ld_b_kn(bb, sq, ks * 16, d4, lane);
          mma16816(acc[2 * d4], dsa[ks], bb[0], bb[1]);
          mma16816(acc[2 * d4 + 1], dsa[ks], bb[2], bb[3]);
        }
      __syncwarp();                                        // the dV rows above have been read back by this warp's store_tile
      stage_tile(Vt, kw, acc, 1.f, 1.f, lane);
      __syncwarp();
      store_tile(Vt, kw, p.dK, p.lddk, ktok, key0, p.Lk, h, lane, p.dbk ? &dbk0 : nullptr, p.dbk ? &dbk1 : nullptr);
    }
    __syncthreads();                 // dS^T of all 64 keys is in shared memory
    // ---------------- phase B: this warp's 16-wide slice of d, all 64 keys of the tile
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t bk[4];
      ld_b_kn(bk, skt, ks * 16, warp, lane);               // K[k = keys 16ks.., n = d slice]: two n-tiles
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t as[4];
        const int mi = lane >> 3, r = lane & 7;
        ldsm4t(as, sds + swz64(ks * 16 + (mi >> 1) * 8 + r, mt * 2 + (mi & 1)));     // A = dS (m = queries 16mt.., k = keys 16ks..)
        mma16816(dq[mt][0], as, bk[0], bk[1]);
        mma16816(dq[mt][1], as, bk[2], bk[3]);
      }
    }
  }
  cp_wait<0>();
  if (p.dbv) { atomicAdd(&dbs[2 * HD + 2 * lane], dbv0); atomicAdd(&dbs[2 * HD + 2 * lane + 1], dbv1); }
  if (p.dbk) { atomicAdd(&dbs[HD + 2 * lane], dbk0); atomicAdd(&dbs[HD + 2 * lane + 1], dbk1); }
  // dQ: warp w owns columns 16w..16w+15; rows = queries
  float csq[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int col = h * HD + warp * 16 + nt * 8 + 2 * t;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int q = mt * 16 + g + hf * 8;
        if (q < p.Lq) {
          const uint32_t w = pack_bf16(dq[mt][nt][2 * hf], dq[mt][nt][2 * hf + 1]);
          *reinterpret_cast<uint32_t*>(p.dQ + (qtok + q) * p.lddq + col) = w;
          csq[nt][0] += __uint_as_float(w << 16); csq[nt][1] += __uint_as_float(w & 0xFFFF0000u);
        }
      }
    }
  if (p.dbq) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) csq[nt][e] += __shfl_xor_sync(0xffffffffu, csq[nt][e], o);
        if (g == 0) atomicAdd(p.dbq + h * HD + warp * 16 + nt * 8 + 2 * t + e, csq[nt][e]);
      }
  }
  __syncthreads();
  if (tid < 64) {
    if (p.dbk) atomicAdd(p.dbk + h * HD + tid, dbs[HD + tid]);
    if (p.dbv) atomicAdd(p.dbv + h * HD + tid, dbs[2 * HD + tid]);
  }
}

NarrowParams make_params(const b200f_attn_args& a) {
  NarrowParams p = {};
  p.B = a.B; p.H = a.H; p.Lq = a.Lq; p.Lk = a.Lk; p.scale = a.scale;
  p.Q = static_cast<const bf16*>(a.Q); p.ldq = a.ldq;
  p.K = static_cast<const bf16*>(a.K); p.ldk = a.ldk;
  p.V = static_cast<const bf16*>(a.V); p.ldv = a.ldv;
  p.O = static_cast<bf16*>(a.O); p.ldo = a.ldo;
  p.LSE = a.LSE;
  p.dO = static_cast<const bf16*>(a.dO); p.lddo = a.lddo;
  p.dQ = static_cast<bf16*>(a.dQ); p.lddq = a.lddq;
  p.dK = static_cast<bf16*>(a.dK); p.lddk = a.lddk;
  p.dV = static_cast<bf16*>(a.dV); p.lddv = a.lddv;
  p.dbq = a.dbq; p.dbk = a.dbk; p.dbv = a.dbv;
  p.pool = a.pool_sum;
  p.drop_thr = drop_threshold(a.dropout_p); p.drop_seed_lo = a.drop_seed_lo; p.drop_seed_hi = a.drop_seed_hi;
  p.inv_keep = 1.f / (1.f - a.dropout_p);
  return p;
}

}  // namespace

// number of partial-sum slots per (batch, head) the forward kernel serving this shape writes into pool_sum (0: none)
int attn_narrow_pool_parts(const b200f_attn_args& a);
int g_attn_narrow = 1;           // b200f_debug_set(10, 0) routes the narrow shapes back to the 128-wide tcgen05 tiles (A/B)

// which narrow kernel family serves this shape: 1 = NK (Lk <= 32), 2 = NQ (Lq <= 32 < Lk), 0 = none
int attn_narrow_kind(const b200f_attn_args& a) {
  if (!g_attn_narrow || a.D != HD || a.dtype != B200F_BF16) return 0;
  if (a.H > 65535 || a.B > 65535) return 0;
  if (a.Lk <= NARROW) return 1;
  if (a.Lq <= NARROW) return 2;
  return 0;
}

int attn_narrow_pool_parts(const b200f_attn_args& a) {
  const int kind = attn_narrow_kind(a);
  return kind == 1 ? (a.Lq + 15) / 16 : (kind == 2 ? 1 : 0);
}

static int narrow_check(const b200f_attn_args& a, bool bwd) {
  B200F_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0, B200F_ERR_ALIGN, "attention(narrow): leading dims must be multiples of 8");
  B200F_REQUIRE(aligned16(a.Q) && aligned16(a.K) && aligned16(a.V) && aligned16(a.O), B200F_ERR_ALIGN, "attention(narrow): 16-byte alignment");
  if (bwd) {
    B200F_REQUIRE(a.lddo % 8 == 0 && a.lddq % 8 == 0 && a.lddk % 8 == 0 && a.lddv % 8 == 0, B200F_ERR_ALIGN, "attention(narrow): gradient leading dims must be multiples of 8");
    B200F_REQUIRE(aligned16(a.dO) && aligned16(a.dQ) && aligned16(a.dK) && aligned16(a.dV), B200F_ERR_ALIGN, "attention(narrow): gradient alignment");
  }
  return B200F_OK;
}

int attn_fwd_narrow(const b200f_attn_args& a, cudaStream_t st) {
  int rc = narrow_check(a, false);
  if (rc) return rc;
  const NarrowParams p = make_params(a);
  const bool drop = p.drop_thr != 0;
  if (attn_narrow_kind(a) == 1) {
    dim3 grid((a.Lq + 127) / 128, a.H, a.B);
    if (drop) attn_nk_fwd_kernel<true><<<grid, 128, 0, st>>>(p);
    else attn_nk_fwd_kernel<false><<<grid, 128, 0, st>>>(p);
    return check_launch("attn_nk_fwd_kernel");
  }
  dim3 grid(a.H, a.B);
  if (drop) {
    B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_nq_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NQF_SMEM));
    attn_nq_fwd_kernel<true><<<grid, 128, NQF_SMEM, st>>>(p);
  } else {
    B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_nq_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NQF_SMEM));
    attn_nq_fwd_kernel<false><<<grid, 128, NQF_SMEM, st>>>(p);
  }
  return check_launch("attn_nq_fwd_kernel");
}

int attn_bwd_narrow(const b200f_attn_args& a, cudaStream_t st) {
  int rc = narrow_check(a, true);
  if (rc) return rc;
  const NarrowParams p = make_params(a);
  const bool drop = p.drop_thr != 0;
  dim3 grid(a.H, a.B);
  if (attn_narrow_kind(a) == 1) {
    if (g_attn_narrow != 2) {                              // default: the d-split kernel; b200f_debug_set(10, 2) = the first version (A/B)
      if (drop) {
        B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_nk_bwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NKB2_SMEM));
        attn_nk_bwd2_kernel<true><<<grid, 128, NKB2_SMEM, st>>>(p);
      } else {
        B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_nk_bwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NKB2_SMEM));
        attn_nk_bwd2_kernel<false><<<grid, 128, NKB2_SMEM, st>>>(p);
      }
      return check_launch("attn_nk_bwd2_kernel");
    }
    if (drop) {
      B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_nk_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NKB_SMEM));
      attn_nk_bwd_kernel<true><<<grid, 128, NKB_SMEM, st>>>(p);
    } else {
      B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_nk_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NKB_SMEM));
      attn_nk_bwd_kernel<false><<<grid, 128, NKB_SMEM, st>>>(p);
    }
    return check_launch("attn_nk_bwd_kernel");
  }
  if (g_attn_narrow != 2) {                                // default: the d-split kernel; b200f_debug_set(10, 2) = the first version (A/B)
    if (drop) {
      B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_nq_bwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NQB2_SMEM));
      attn_nq_bwd2_kernel<true><<<grid, 128, NQB2_SMEM, st>>>(p);
    } else {
      B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_nq_bwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NQB2_SMEM));
      attn_nq_bwd2_kernel<false><<<grid, 128, NQB2_SMEM, st>>>(p);
    }
    return check_launch("attn_nq_bwd2_kernel");
  }
  if (drop) {
    B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_nq_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NQB_SMEM));
    attn_nq_bwd_kernel<true><<<grid, 128, NQB_SMEM, st>>>(p);
  } else {
    B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_nq_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NQB_SMEM));
    attn_nq_bwd_kernel<false><<<grid, 128, NQB_SMEM, st>>>(p);
  }
  return check_launch("attn_nq_bwd_kernel");
}

B200F_DEFINE_EPOCH_HOOK(attn_narrow)

}  // namespace b200f
