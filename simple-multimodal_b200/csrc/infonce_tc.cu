// Fused InfoNCE on the 5th-gen tensor cores (bf16 embeddings, D = 64 / 128 / 192 / 256): the [Bl x Bg] similarity block
// never leaves the SM (ContrastiveFusion.contrastive_loss, reference models/fusion_layers.py:361-375; SURVEY 8e option A).
//
//   lse  kernel : S = x y^T on tcgen05 (M = 128 rows of x, N = 128 rows of y, K = D), scores read once from TMEM into
//                 registers, online (max, sum-exp) in the log2 domain with warp-uniform control flow, the diagonal logit
//                 picked up on the way.  Grid = (row tiles, key splits); per-split partial (max, sum) pairs are merged by a
//                 small combine kernel, so 32 row tiles still fill 148 SMs.
//   grad kernel : the same S tiles are recomputed, W = c (exp(S - lse_x[i]) + exp(S - lse_y[j]) - 2 [j == off + i]) goes
//                 to shared memory as the bf16 A operand of dX[128 x D] += W Y (Y is the tile already resident for S, read
//                 MN-major), dX accumulates in TMEM over the split's key tiles and is added to dx with fp32 vector REDs.
//
// Per CTA: warp 0 TMA producer (x tile once, 2-deep ring of y tiles), warp 1 MMA issuer, warps 2-5 softmax.  The S
// accumulator is double-buffered in TMEM and released as soon as the row sits in registers, so the next tile's MMAs run
// under the exp work (the attention kernels' run-ahead schedule).
#include "common.cuh"
#include "ptx.cuh"

namespace b200f {

int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);

static constexpr int NCE_THREADS = 192;
static constexpr int NT = 128;                 // rows of x per CTA, rows of y per tile
static constexpr int SUB = NT * 64 * 2;        // one [128 x 64] bf16 K-major SW128 sub-tile: 16 KB
static constexpr int SUB16 = SUB / 16;
static constexpr float NCE_LOG2E = 1.4426950408889634f;

struct NceParams {
  int Bl, Bg, nk;                              // nk = D / 64
  int n_tiles, tiles_per_split;
  long long diag_off;
  float inv_tau;
  // lse
  float* part_m; float* part_l;                // [splits][Bl] log2-domain partial max / sum
  float* diag;                                 // [Bl] or null
  // grad
  const float* lse_x; const float* lse_y;      // natural-log LSE of the rows of x against all y / of the rows of y against all x
  const float* gscale; float coef;
  float* dx; int D;
};

struct NceSmem {
  static constexpr int X_OFF = 0;                         // 4 sub-tiles [128 x 64]
  static constexpr int Y_OFF = X_OFF + 4 * SUB;           // 2 stages x 4 sub-tiles
  static constexpr int W_OFF = Y_OFF + 2 * 4 * SUB;       // grad: W [128 x 128] bf16 as two K-major halves of 64 keys
  static constexpr int STAT_OFF = W_OFF + 2 * SUB;        // grad: 2 x 128 floats (-lse_y * log2e of the tile's columns)
  static constexpr int BAR_OFF = STAT_OFF + 2 * NT * 4;
  static constexpr int TOTAL = BAR_OFF + 256;
  static constexpr int TOTAL_LSE = W_OFF + 256;           // the lse kernel has no W tile: its barriers sit at W_OFF
};

__device__ __forceinline__ void nce_load_tile(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int row0, int nk) {
  mbar_expect_tx(bar, uint32_t(nk) * SUB);
  for (int kk = 0; kk < nk; ++kk) tma_load_2d(dst + kk * SUB, tm, bar, kk * 64, row0);
}

// S[buf] = X Y(stage)^T, K = 64 * nk
__device__ __forceinline__ void nce_issue_s(uint32_t t_s, uint32_t x_lo, uint32_t y_lo, int nk, uint32_t idesc) {
  for (int kk = 0; kk < nk; ++kk) umma_chain<4>(t_s, x_lo + kk * SUB16, 2, y_lo + kk * SUB16, 2, idesc, kk > 0);
}

__global__ void __launch_bounds__(NCE_THREADS, 1)
infonce_lse_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y, const NceParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NceSmem::W_OFF);
  uint64_t* x_full = bars + 0;
  uint64_t* y_full = bars + 1;     // [2]
  uint64_t* y_empty = bars + 3;    // [2]
  uint64_t* s_full = bars + 5;     // [2]
  uint64_t* s_free = bars + 7;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * NT, split = blockIdx.y;
  const int jt0 = split * p.tiles_per_split;
  const int jt1 = min(p.n_tiles, jt0 + p.tiles_per_split);
  const int nt = jt1 - jt0;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_y);
    mbar_init(x_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      nce_load_tile(smem + NceSmem::X_OFF, &tm_x, x_full, r0, p.nk);
      for (int n = 0; n < nt; ++n) {
        const int st = n & 1;
        mbar_wait(&y_empty[st], ((n >> 1) & 1) ^ 1);
        nce_load_tile(smem + NceSmem::Y_OFF + st * 4 * SUB, &tm_y, &y_full[st], (jt0 + n) * NT, p.nk);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      const bool leader = elect_one();   // warp-uniform loops; only the MMA / commit instructions are predicated
      constexpr uint32_t idesc = umma_idesc_bf16(NT, NT, 0, 0);
      const uint32_t x_lo = umma_lo(smem_u32(smem + NceSmem::X_OFF), 16), y_lo = umma_lo(smem_u32(smem + NceSmem::Y_OFF), 16);
      mbar_wait(x_full, 0);
      for (int n = 0; n < nt; ++n) {
        const int st = n & 1;
        mbar_wait(&y_full[st], (n >> 1) & 1);
        mbar_wait(&s_free[st], ((n >> 1) & 1) ^ 1);
        tc_fence_after();
        if (leader) nce_issue_s(tmem_base + st * NT, x_lo, y_lo + st * 4 * SUB16, p.nk, idesc);
        if (leader) umma_commit(&s_full[st]);
        if (leader) umma_commit(&y_empty[st]);
      }
    }
    __syncwarp();
  } else {
    const int grp = warp & 3;
    const int row = grp * 32 + lane;
    const uint32_t lane_addr = uint32_t(grp * 32) << 16;
    const float c = p.inv_tau * NCE_LOG2E;
    const uint64_t c2 = f2_pack(c, c);
    const long long grow = r0 + row;
    const long long dcol = p.diag_off + grow;            // column of this row's positive
    float m = -INFINITY, l = 0.f, dval = 0.f;
    bool dfound = false;
    for (int n = 0; n < nt; ++n) {
      const int st = n & 1;
      mbar_wait(&s_full[st], (n >> 1) & 1);
      tc_fence_after();
      uint32_t r[4][32];
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_ld32(tmem_base + st * NT + lane_addr + q * 32, r[q]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[st]);
      const long long col0 = (long long)(jt0 + n) * NT;
      const int valid = int(min((long long)NT, p.Bg - col0));       // rows of y past Bg were zero-filled by TMA: mask them
      if (p.diag && dcol >= col0 && dcol < col0 + NT) {              // this tile holds the row's positive logit
        const int dj = int(dcol - col0);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (q * 32 + i == dj) dval = __uint_as_float(r[q][i]) * p.inv_tau;
        dfound = true;
      }
      if (valid < NT) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (q * 32 + i >= valid) r[q][i] = 0xff800000u;
      }
      float tmax = -INFINITY, tmax_b = -INFINITY;
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          tmax = fmax3(tmax, __uint_as_float(r[q][i]), __uint_as_float(r[q][i + 1]));
          tmax_b = fmax3(tmax_b, __uint_as_float(r[q][i + 2]), __uint_as_float(r[q][i + 3]));
        }
      const float m_new = fmaxf(m, fmaxf(tmax, tmax_b) * c);          // log2 domain (c > 0)
      const float nm = -m_new;
      const uint64_t nm2 = f2_pack(nm, nm);
      uint64_t sum2 = 0ull;
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float x0, x1;
          f2_unpack(f2_fma(f2_pack_u(r[q][i], r[q][i + 1]), c2, nm2), x0, x1);
          sum2 = f2_add(sum2, f2_pack(ex2_approx(x0), ex2_approx(x1)));
        }
      float s0, s1;
      f2_unpack(sum2, s0, s1);
      l = l * ex2_approx(m - m_new) + (s0 + s1);                      // first tile: m = -inf -> factor 0
      m = m_new;
    }
    if (grow < p.Bl) {
      p.part_m[(long long)split * p.Bl + grow] = m;
      p.part_l[(long long)split * p.Bl + grow] = l;
      if (dfound) p.diag[grow] = dval;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// lse[i] = ln sum_j exp(S_ij) from the per-split (max, sum) pairs (log2 domain)
__global__ void infonce_lse_combine_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l, float* __restrict__ lse, int Bl, int splits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bl) return;
  float M = -INFINITY;
  for (int s = 0; s < splits; ++s) M = fmaxf(M, part_m[(long long)s * Bl + i]);
  float L = 0.f;
  for (int s = 0; s < splits; ++s) L += part_l[(long long)s * Bl + i] * exp2f(part_m[(long long)s * Bl + i] - M);
  lse[i] = (M + log2f(L)) * 0.6931471805599453f;
}

__global__ void __launch_bounds__(NCE_THREADS, 1)
infonce_grad_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y, const NceParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NceSmem::BAR_OFF);
  uint64_t* x_full = bars + 0;
  uint64_t* y_full = bars + 1;     // [2]
  uint64_t* y_empty = bars + 3;    // [2]
  uint64_t* s_full = bars + 5;     // [2]
  uint64_t* s_free = bars + 7;     // [2]
  uint64_t* w_ready = bars + 9;    // W(n) in smem
  uint64_t* w_free = bars + 10;    // dX MMAs of tile n retired: W may be rewritten; the last one = dX complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  float* stats = reinterpret_cast<float*>(smem + NceSmem::STAT_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * NT, split = blockIdx.y;
  const int jt0 = split * p.tiles_per_split;
  const int jt1 = min(p.n_tiles, jt0 + p.tiles_per_split);
  const int nt = jt1 - jt0;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_y);
    mbar_init(x_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4); }
    mbar_init(w_ready, 4);
    mbar_init(w_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_dx = tmem_base + 2 * NT;

  if (warp == 0) {
    if (lane == 0) {
      nce_load_tile(smem + NceSmem::X_OFF, &tm_x, x_full, r0, p.nk);
      for (int n = 0; n < nt; ++n) {
        const int st = n & 1;
        mbar_wait(&y_empty[st], ((n >> 1) & 1) ^ 1);
        nce_load_tile(smem + NceSmem::Y_OFF + st * 4 * SUB, &tm_y, &y_full[st], (jt0 + n) * NT, p.nk);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      const bool leader = elect_one();   // warp-uniform loops; only the MMA / commit instructions are predicated
      constexpr uint32_t idesc_s = umma_idesc_bf16(NT, NT, 0, 0);
      constexpr uint32_t idesc_x = umma_idesc_bf16(NT, 64, 0, 1);      // dX[:, 64 kk .. +64) += W Y_kk   (Y MN-major, N = 64 columns of D)
      const uint32_t x_lo = umma_lo(smem_u32(smem + NceSmem::X_OFF), 16), y_lo = umma_lo(smem_u32(smem + NceSmem::Y_OFF), 16);
      const uint32_t ymn_lo = umma_lo(smem_u32(smem + NceSmem::Y_OFF), NT * 128), w_lo = umma_lo(smem_u32(smem + NceSmem::W_OFF), 16);
      auto issue_dx = [&](int n) {                                    // tile n's W against its Y stage
        const int st = n & 1;
        mbar_wait(w_ready, n & 1);
        tc_fence_after();
        for (int kk = 0; kk < p.nk; ++kk) {
          const uint32_t yb = ymn_lo + st * 4 * SUB16 + kk * SUB16;
          if (leader) umma_chain<4>(t_dx + kk * 64, w_lo, 2, yb, 128, idesc_x, n > 0);
          if (leader) umma_chain<4>(t_dx + kk * 64, w_lo + SUB16, 2, yb + 4 * 128, 128, idesc_x, 1);
        }
        if (leader) umma_commit(w_free);
        if (leader) umma_commit(&y_empty[st]);
      };
      mbar_wait(x_full, 0);
      for (int n = 0; n < nt; ++n) {
        const int st = n & 1;
        mbar_wait(&y_full[st], (n >> 1) & 1);
        mbar_wait(&s_free[st], ((n >> 1) & 1) ^ 1);
        tc_fence_after();
        if (leader) nce_issue_s(tmem_base + st * NT, x_lo, y_lo + st * 4 * SUB16, p.nk, idesc_s);
        if (leader) umma_commit(&s_full[st]);
        if (n > 0) issue_dx(n - 1);                                   // S(n) runs ahead of the accumulation of tile n-1
      }
      if (nt > 0) issue_dx(nt - 1);
    }
    __syncwarp();
  } else {
    const int grp = warp & 3;
    const int row = grp * 32 + lane;
    const int t128 = threadIdx.x - 64;
    const uint32_t lane_addr = uint32_t(grp * 32) << 16;
    const float c1 = p.inv_tau * NCE_LOG2E;
    const uint64_t c12 = f2_pack(c1, c1);
    const long long grow = r0 + row;
    const long long dcol = p.diag_off + grow;
    const float cw = p.coef * (p.gscale ? *p.gscale : 1.f);
    const float nlx = grow < p.Bl ? -p.lse_x[grow] * NCE_LOG2E : 0.f;
    const uint64_t nlx2 = f2_pack(nlx, nlx);
    auto load_stat = [&](int n) -> float {                            // -lse_y * log2e of column t128 of tile n (prefetched one tile ahead)
      const long long col = (long long)(jt0 + n) * NT + t128;
      return col < p.Bg ? -p.lse_y[col] * NCE_LOG2E : 0.f;
    };
    float stat_next = nt > 0 ? load_stat(0) : 0.f;
    for (int n = 0; n < nt; ++n) {
      const int st = n & 1;
      float* sl = stats + st * NT;
      sl[t128] = stat_next;
      if (n + 1 < nt) stat_next = load_stat(n + 1);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&s_full[st], (n >> 1) & 1);
      tc_fence_after();
      uint32_t r[4][32];
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_ld32(tmem_base + st * NT + lane_addr + q * 32, r[q]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[st]);
      const long long col0 = (long long)(jt0 + n) * NT;
      const int valid = int(min((long long)NT, p.Bg - col0));
      const int dj = (dcol >= col0 && dcol < col0 + NT) ? int(dcol - col0) : -1;
      uint32_t pk[4][16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint64_t* nly2 = reinterpret_cast<const uint64_t*>(sl + q * 32);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t sc = f2_pack_u(r[q][i], r[q][i + 1]);
          float a0, a1, b0, b1;
          f2_unpack(f2_fma(sc, c12, nlx2), a0, a1);                    // log2 of the row-softmax probability
          f2_unpack(f2_fma(sc, c12, nly2[i >> 1]), b0, b1);            // log2 of the column-softmax probability
          float w0 = ex2_approx(a0) + ex2_approx(b0), w1 = ex2_approx(a1) + ex2_approx(b1);
          if (q * 32 + i == dj) w0 -= 2.f;
          if (q * 32 + i + 1 == dj) w1 -= 2.f;
          if (q * 32 + i >= valid) w0 = 0.f;
          if (q * 32 + i + 1 >= valid) w1 = 0.f;
          pk[q][i >> 1] = pack_bf16(w0 * cw, w1 * cw);
        }
      }
      if (n > 0) mbar_wait(w_free, (n - 1) & 1);                       // dX MMAs of the previous tile have read W
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint8_t* half = smem + NceSmem::W_OFF + (q >> 1) * SUB;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          *reinterpret_cast<uint4*>(half + sw128_offset(row, (q & 1) * 4 + q4)) = make_uint4(pk[q][q4 * 4], pk[q][q4 * 4 + 1], pk[q][q4 * 4 + 2], pk[q][q4 * 4 + 3]);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(w_ready);
    }
    if (nt > 0) {
      mbar_wait(w_free, (nt - 1) & 1);
      tc_fence_after();
      float* out = p.dx + grow * p.D;
      for (int cc = 0; cc < p.D; cc += 32) {
        uint32_t v[32];
        tmem_ld32(t_dx + lane_addr + cc, v);             // .sync.aligned: every lane of the warp takes part, rows past Bl only skip the REDs
        tmem_ld_wait();
        if (grow < p.Bl) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + cc + i), "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])),
                         "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3])) : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

bool infonce_tc_eligible(const void* x, const void* y, int64_t Bl, int64_t Bg, int32_t D, int32_t dtype) {
  return dtype == B200F_BF16 && D >= 64 && D <= 256 && D % 64 == 0 && aligned16(x) && aligned16(y) && Bl > 0 && Bg > 0 && Bg < (1ll << 31) && Bl < (1ll << 31);
}

static int nce_setup(const void* x, const void* y, int64_t Bl, int64_t Bg, int32_t D, CUtensorMap* tx, CUtensorMap* ty, NceParams* p, dim3* grid) {
  uint64_t dims[2] = {uint64_t(D), uint64_t(Bl)}, strides[1] = {uint64_t(D) * 2};
  uint32_t box[2] = {64, NT};
  int rc = make_tmap_bf16(tx, x, 2, dims, strides, box);
  if (rc) return rc;
  dims[1] = uint64_t(Bg);
  if ((rc = make_tmap_bf16(ty, y, 2, dims, strides, box))) return rc;
  p->Bl = int(Bl); p->Bg = int(Bg); p->nk = D / 64; p->D = D;
  p->n_tiles = int((Bg + NT - 1) / NT);
  const int row_tiles = int((Bl + NT - 1) / NT);
  // as many key splits as fit in ONE wave of CTAs (one CTA per SM: the kernels use > 113 KB of shared memory).  Rounding up instead
  // (round 1: 32 row tiles x 5 splits = 160 CTAs on 148 SMs) put 12 CTAs into a second wave and doubled the kernel's time.
  int splits = num_sms() / row_tiles;
  if (splits > p->n_tiles) splits = p->n_tiles;
  if (splits > 16) splits = 16;
  if (splits < 1) splits = 1;
  p->tiles_per_split = (p->n_tiles + splits - 1) / splits;
  splits = (p->n_tiles + p->tiles_per_split - 1) / p->tiles_per_split;
  *grid = dim3(row_tiles, splits);
  return B200F_OK;
}

size_t infonce_tc_workspace_bytes(int64_t Bl) { return size_t(2) * 16 * size_t(Bl) * sizeof(float); }

int infonce_lse_tc(const void* x, const void* y, float* lse, float* diag, int64_t Bl, int64_t Bg, int32_t D, int64_t diag_off, float inv_tau,
                   void* workspace, cudaStream_t st) {
  CUtensorMap tx, ty;
  NceParams p = {};
  dim3 grid;
  int rc = nce_setup(x, y, Bl, Bg, D, &tx, &ty, &p, &grid);
  if (rc) return rc;
  p.diag_off = diag_off; p.inv_tau = inv_tau; p.diag = diag;
  p.part_m = static_cast<float*>(workspace);
  p.part_l = p.part_m + size_t(grid.y) * Bl;
  static PerDeviceOnce configured;
  if (configured.first()) {
    B200F_CHECK_CUDA(cudaFuncSetAttribute(infonce_lse_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NceSmem::TOTAL_LSE));
    B200F_CHECK_CUDA(cudaFuncSetAttribute(infonce_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NceSmem::TOTAL));
  }
  infonce_lse_tc_kernel<<<grid, NCE_THREADS, NceSmem::TOTAL_LSE, st>>>(tx, ty, p);
  if ((rc = check_launch("infonce_lse_tc_kernel"))) return rc;
  infonce_lse_combine_kernel<<<(unsigned)((Bl + 255) / 256), 256, 0, st>>>(p.part_m, p.part_l, lse, int(Bl), int(grid.y));
  return check_launch("infonce_lse_combine_kernel");
}

int infonce_grad_tc(const void* x, const void* y, const float* lse_x, const float* lse_y, float coef, const float* gscale_dev, float* dx, int64_t Bl,
                    int64_t Bg, int32_t D, int64_t diag_off, float inv_tau, cudaStream_t st) {
  CUtensorMap tx, ty;
  NceParams p = {};
  dim3 grid;
  int rc = nce_setup(x, y, Bl, Bg, D, &tx, &ty, &p, &grid);
  if (rc) return rc;
  p.diag_off = diag_off; p.inv_tau = inv_tau;
  p.lse_x = lse_x; p.lse_y = lse_y; p.gscale = gscale_dev; p.coef = coef; p.dx = dx;
  static PerDeviceOnce configured;
  if (configured.first()) {
    B200F_CHECK_CUDA(cudaFuncSetAttribute(infonce_lse_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NceSmem::TOTAL_LSE));
    B200F_CHECK_CUDA(cudaFuncSetAttribute(infonce_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NceSmem::TOTAL));
  }
  infonce_grad_tc_kernel<<<grid, NCE_THREADS, NceSmem::TOTAL, st>>>(tx, ty, p);
  return check_launch("infonce_grad_tc_kernel");
}

}  // namespace b200f
