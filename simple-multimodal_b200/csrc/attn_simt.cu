// CUDA-core multi-head attention (forward, backward) for the fp32 parity mode and for shapes the
// tcgen05 kernel does not take (head_dim != 64).  One warp per (batch, head, row); fp32 arithmetic;
// token-major strided Q/K/V so they can alias slices of the packed projection outputs.
//   fwd : s_j = scale * <q_i, k_j>; p = softmax_j(s); o_i = sum_j p_j v_j; lse_i = log sum_j exp(s_j)
//   bwd : delta_i = <do_i, o_i>;  ds_ij = p_ij (<do_i, v_j> - delta_i)
//         dq_i = scale * sum_j ds_ij k_j;   dk_j = scale * sum_i ds_ij q_i;   dv_j = sum_i p_ij do_i
#include "common.cuh"

namespace b200f {

static constexpr int MAXD = 64;

struct AttnP {
  int B, H, Lq, Lk, D;
  const void *Q, *K, *V, *dO;
  void *O, *dQ, *dK, *dV;
  long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  float* LSE;
  float* delta;
  float scale;
  uint32_t drop_thr, drop_seed_lo, drop_seed_hi;   // attention-probability dropout (0 = off), same mask as the tcgen05 kernels
  float inv_keep;
};

template <typename T>
__device__ __forceinline__ const T* tok(const void* base, long long ld, int L, int b, int l, int h, int D) {
  return static_cast<const T*>(base) + ((long long)b * L + l) * ld + h * D;
}
template <typename T>
__device__ __forceinline__ T* tokw(void* base, long long ld, int L, int b, int l, int h, int D) {
  return static_cast<T*>(base) + ((long long)b * L + l) * ld + h * D;
}

template <typename T>
__global__ void __launch_bounds__(128) attn_fwd_simt(const AttnP p) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * 4 + warp;
  const long long total = (long long)p.B * p.H * p.Lq;
  float* sc = sm + warp * (p.Lk + MAXD);
  float* qs = sc + p.Lk;
  if (gw >= total) return;
  const int i = int(gw % p.Lq), h = int((gw / p.Lq) % p.H), b = int(gw / ((long long)p.Lq * p.H));
  const T* q = tok<T>(p.Q, p.ldq, p.Lq, b, i, h, p.D);
  for (int d = lane; d < p.D; d += 32) qs[d] = to_f32(q[d]) * p.scale;
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < p.Lk; j += 32) {
    const T* k = tok<T>(p.K, p.ldk, p.Lk, b, j, h, p.D);
    float s = 0.f;
    for (int d = 0; d < p.D; ++d) s = fmaf(qs[d], to_f32(k[d]), s);
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < p.Lk; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (p.drop_thr) {                                  // mask what multiplies V; the row sum stays that of the undropped scores
    const uint32_t rk = drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, uint32_t(gw));
    for (int j = lane; j < p.Lk; j += 32) sc[j] = drop_keep_attn(rk, uint32_t(j), p.drop_thr) ? sc[j] * p.inv_keep : 0.f;
  }
  __syncwarp();
  const float inv = 1.f / sum;
  T* o = tokw<T>(p.O, p.ldo, p.Lq, b, i, h, p.D);
  for (int d = lane; d < p.D; d += 32) {
    float acc = 0.f;
    for (int j = 0; j < p.Lk; ++j) acc = fmaf(sc[j], to_f32(tok<T>(p.V, p.ldv, p.Lk, b, j, h, p.D)[d]), acc);
    o[d] = from_f32<T>(acc * inv);
  }
  if (lane == 0) p.LSE[((long long)b * p.H + h) * p.Lq + i] = mx + logf(sum);
}

template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_dq_simt(const AttnP p) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * 4 + warp;
  const long long total = (long long)p.B * p.H * p.Lq;
  float* ds = sm + warp * (p.Lk + 2 * MAXD);
  float* qs = ds + p.Lk;
  float* gs = qs + MAXD;
  if (gw >= total) return;
  const int i = int(gw % p.Lq), h = int((gw / p.Lq) % p.H), b = int(gw / ((long long)p.Lq * p.H));
  const T* q = tok<T>(p.Q, p.ldq, p.Lq, b, i, h, p.D);
  const T* g = tok<T>(p.dO, p.lddo, p.Lq, b, i, h, p.D);
  const T* o = tok<T>(p.O, p.ldo, p.Lq, b, i, h, p.D);
  float dl = 0.f;
  for (int d = lane; d < p.D; d += 32) {
    qs[d] = to_f32(q[d]) * p.scale;
    gs[d] = to_f32(g[d]);
    dl += gs[d] * to_f32(o[d]);
  }
  dl = warp_sum(dl);
  const long long ri = ((long long)b * p.H + h) * p.Lq + i;
  if (lane == 0) p.delta[ri] = dl;
  const float lse = p.LSE[ri];
  __syncwarp();
  for (int j = lane; j < p.Lk; j += 32) {
    const T* k = tok<T>(p.K, p.ldk, p.Lk, b, j, h, p.D);
    const T* v = tok<T>(p.V, p.ldv, p.Lk, b, j, h, p.D);
    float s = 0.f, dp = 0.f;
    for (int d = 0; d < p.D; ++d) {
      s = fmaf(qs[d], to_f32(k[d]), s);
      dp = fmaf(gs[d], to_f32(v[d]), dp);
    }
    if (p.drop_thr) dp = drop_keep_attn(drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, uint32_t(gw)), uint32_t(j), p.drop_thr) ? dp * p.inv_keep : 0.f;
    ds[j] = expf(s - lse) * (dp - dl);
  }
  __syncwarp();
  T* dq = tokw<T>(p.dQ, p.lddq, p.Lq, b, i, h, p.D);
  for (int d = lane; d < p.D; d += 32) {
    float acc = 0.f;
    for (int j = 0; j < p.Lk; ++j) acc = fmaf(ds[j], to_f32(tok<T>(p.K, p.ldk, p.Lk, b, j, h, p.D)[d]), acc);
    dq[d] = from_f32<T>(acc * p.scale);
  }
}

// one warp per (b, h, key j); lanes stride over the queries; needs delta from the dq kernel.
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_dkv_simt(const AttnP p) {
  __shared__ float ks[4][MAXD], vs[4][MAXD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * 4 + warp;
  const long long total = (long long)p.B * p.H * p.Lk;
  if (gw >= total) return;
  const int j = int(gw % p.Lk), h = int((gw / p.Lk) % p.H), b = int(gw / ((long long)p.Lk * p.H));
  const T* k = tok<T>(p.K, p.ldk, p.Lk, b, j, h, p.D);
  const T* v = tok<T>(p.V, p.ldv, p.Lk, b, j, h, p.D);
  for (int d = lane; d < p.D; d += 32) { ks[warp][d] = to_f32(k[d]); vs[warp][d] = to_f32(v[d]); }
  __syncwarp();
  float dk[MAXD], dv[MAXD];
#pragma unroll
  for (int d = 0; d < MAXD; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
  for (int i = lane; i < p.Lq; i += 32) {
    const T* q = tok<T>(p.Q, p.ldq, p.Lq, b, i, h, p.D);
    const T* g = tok<T>(p.dO, p.lddo, p.Lq, b, i, h, p.D);
    float s = 0.f, dp = 0.f;
    for (int d = 0; d < p.D; ++d) {
      s = fmaf(to_f32(q[d]), ks[warp][d], s);
      dp = fmaf(to_f32(g[d]), vs[warp][d], dp);
    }
    const long long ri = ((long long)b * p.H + h) * p.Lq + i;
    const float pij = expf(s * p.scale - p.LSE[ri]);
    float pd = pij;                                  // dropped + rescaled probability (multiplies dO in dV)
    if (p.drop_thr) {
      const bool keep = drop_keep_attn(drop_row_key_e(p.drop_seed_lo, p.drop_seed_hi, uint32_t(ri)), uint32_t(j), p.drop_thr);
      dp = keep ? dp * p.inv_keep : 0.f;
      pd = keep ? pij * p.inv_keep : 0.f;
    }
    const float dsij = pij * (dp - p.delta[ri]) * p.scale;
#pragma unroll
    for (int d = 0; d < MAXD; ++d)
      if (d < p.D) {
        dv[d] = fmaf(pd, to_f32(g[d]), dv[d]);
        dk[d] = fmaf(dsij, to_f32(q[d]), dk[d]);
      }
  }
  T* dkp = tokw<T>(p.dK, p.lddk, p.Lk, b, j, h, p.D);
  T* dvp = tokw<T>(p.dV, p.lddv, p.Lk, b, j, h, p.D);
#pragma unroll
  for (int d = 0; d < MAXD; ++d)
    if (d < p.D) {
      const float a = warp_sum(dk[d]), c = warp_sum(dv[d]);
      if (lane == 0) { dkp[d] = from_f32<T>(a); dvp[d] = from_f32<T>(c); }
    }
}

static AttnP make_params(const b200f_attn_args& a) {
  AttnP p;
  p.B = a.B; p.H = a.H; p.Lq = a.Lq; p.Lk = a.Lk; p.D = a.D;
  p.Q = a.Q; p.K = a.K; p.V = a.V; p.dO = a.dO; p.O = a.O; p.dQ = a.dQ; p.dK = a.dK; p.dV = a.dV;
  p.ldq = a.ldq; p.ldk = a.ldk; p.ldv = a.ldv; p.ldo = a.ldo; p.lddo = a.lddo; p.lddq = a.lddq; p.lddk = a.lddk; p.lddv = a.lddv;
  p.LSE = a.LSE; p.delta = a.delta; p.scale = a.scale;
  p.drop_thr = drop_threshold(a.dropout_p); p.drop_seed_lo = a.drop_seed_lo; p.drop_seed_hi = a.drop_seed_hi;
  p.inv_keep = 1.f / (1.f - a.dropout_p);
  return p;
}

template <typename T>
static int attn_fwd_simt_launch(const b200f_attn_args& a, cudaStream_t st) {
  const AttnP p = make_params(a);
  const long long warps = (long long)a.B * a.H * a.Lq;
  const size_t smem = 4 * (a.Lk + MAXD) * sizeof(float);
  B200F_REQUIRE(smem <= 200 * 1024, B200F_ERR_SHAPE, "attention(simt): Lk=%d too long", a.Lk);
  if (smem > 48 * 1024) B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_simt<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_fwd_simt<T><<<(unsigned)((warps + 3) / 4), 128, smem, st>>>(p);
  return check_launch("attn_fwd_simt");
}

template <typename T>
static int attn_bwd_simt_launch(const b200f_attn_args& a, cudaStream_t st) {
  const AttnP p = make_params(a);
  const long long wq = (long long)a.B * a.H * a.Lq, wk = (long long)a.B * a.H * a.Lk;
  const size_t smem = 4 * (a.Lk + 2 * MAXD) * sizeof(float);
  B200F_REQUIRE(smem <= 200 * 1024, B200F_ERR_SHAPE, "attention(simt): Lk=%d too long", a.Lk);
  if (smem > 48 * 1024) B200F_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_simt<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_bwd_dq_simt<T><<<(unsigned)((wq + 3) / 4), 128, smem, st>>>(p);
  int rc = check_launch("attn_bwd_dq_simt");
  if (rc) return rc;
  attn_bwd_dkv_simt<T><<<(unsigned)((wk + 3) / 4), 128, 0, st>>>(p);
  return check_launch("attn_bwd_dkv_simt");
}

int attn_fwd_simt_dispatch(const b200f_attn_args& a, cudaStream_t st) {
  B200F_REQUIRE(a.D > 0 && a.D <= MAXD, B200F_ERR_SHAPE, "attention(simt): head dim %d not in [1,%d]", a.D, MAXD);
  return a.dtype == B200F_F32 ? attn_fwd_simt_launch<float>(a, st) : attn_fwd_simt_launch<bf16>(a, st);
}
int attn_bwd_simt_dispatch(const b200f_attn_args& a, cudaStream_t st) {
  B200F_REQUIRE(a.D > 0 && a.D <= MAXD, B200F_ERR_SHAPE, "attention(simt): head dim %d not in [1,%d]", a.D, MAXD);
  return a.dtype == B200F_F32 ? attn_bwd_simt_launch<float>(a, st) : attn_bwd_simt_launch<bf16>(a, st);
}

B200F_DEFINE_EPOCH_HOOK(attn_simt)

}  // namespace b200f
