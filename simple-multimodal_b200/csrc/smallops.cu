// Per-sample heads of the fusion path that are too small for tensor cores: the 3-node graph attention
// (GraphFusion), attention over the 3 modality tokens and the gated mix (AdaptiveFusion), the late-fusion
// combine, the modality-dropout mask and inverted dropout.  One warp per sample, 128-bit vector loads,
// warp-shuffle reductions (segmented per head where a head spans a sub-group of lanes).
#include "common.cuh"

namespace b200f {

static constexpr int GH = 4;      // GAT heads handled by the register layout below
static constexpr int MAXV = 16;   // max 16-byte vectors per lane per row (C <= 16*32*VN)

// ------------------------------------------------------------------------------------------------
// GATConv core on dense 3-node graphs
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void gat_scores(const T* __restrict__ xp, const float* __restrict__ att_src, const float* __restrict__ att_dst,
                                           int heads, int C, int lane, float (&a_src)[3][GH], float (&a_dst)[3][GH]) {
  constexpr int VN = Vec16<T>::N;
  const int nvec = C / VN;
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int h = 0; h < GH; ++h) { a_src[j][h] = 0.f; a_dst[j][h] = 0.f; }
#pragma unroll
  for (int h = 0; h < GH; ++h)
    if (h < heads)
    for (int vi = lane; vi < nvec; vi += 32) {
      float ws[VN], wd[VN];
#pragma unroll
      for (int e = 0; e < VN; ++e) { ws[e] = att_src[h * C + vi * VN + e]; wd[e] = att_dst[h * C + vi * VN + e]; }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        Vec16<T> t; t.load(xp + ((long long)j * heads + h) * C + vi * VN);
        float f[VN]; t.unpack(f);
#pragma unroll
        for (int e = 0; e < VN; ++e) { a_src[j][h] += f[e] * ws[e]; a_dst[j][h] += f[e] * wd[e]; }
      }
    }
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int h = 0; h < GH; ++h) { a_src[j][h] = warp_sum(a_src[j][h]); a_dst[j][h] = warp_sum(a_dst[j][h]); }
}

template <typename T>
__global__ void __launch_bounds__(128) gat_fwd_kernel(const T* __restrict__ xp, const float* __restrict__ att_src, const float* __restrict__ att_dst,
                                                      const float* __restrict__ bias, T* __restrict__ out, float* __restrict__ alpha_out,
                                                      long long B, int heads, int C, float slope, uint32_t drop_thr, uint32_t seed_lo, uint32_t seed_hi, float inv_keep) {
  constexpr int VN = Vec16<T>::N;
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  const T* x = xp + b * 3LL * heads * C;
  float a_src[3][GH], a_dst[3][GH];
  gat_scores<T>(x, att_src, att_dst, heads, C, lane, a_src, a_dst);
  float al[3][GH][3];  // [i][h][j]
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int h = 0; h < GH; ++h) {
      float e[3], mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 3; ++j) { const float v = a_src[j][h] + a_dst[i][h]; e[j] = v > 0.f ? v : slope * v; mx = fmaxf(mx, e[j]); }
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j) { e[j] = expf(e[j] - mx); s += e[j]; }
#pragma unroll
      for (int j = 0; j < 3; ++j) al[i][h][j] = e[j] / s;
    }
  if (lane == 0) {
    float* ao = alpha_out + b * 3LL * heads * 3;
    for (int i = 0; i < 3; ++i)
      for (int h = 0; h < heads; ++h)
        for (int j = 0; j < 3; ++j) ao[(i * heads + h) * 3 + j] = al[i][h][j];   // the softmax output; dropout is regenerated in backward
  }
  if (drop_thr) {            // GATConv(dropout=p): F.dropout on the attention coefficients before aggregation (training mode)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int h = 0; h < GH; ++h) {
        const uint32_t rk = drop_row_key_e(seed_lo, seed_hi, uint32_t((b * heads + h) * 3 + i));
#pragma unroll
        for (int j = 0; j < 3; ++j) al[i][h][j] = drop_keep(rk, uint32_t(j), drop_thr) ? al[i][h][j] * inv_keep : 0.f;
      }
  }
  const int nvec = C / VN;
  const float invh = 1.f / heads;
  for (int vi = lane; vi < nvec; vi += 32) {
    float acc[3][VN];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int e = 0; e < VN; ++e) acc[i][e] = 0.f;
#pragma unroll
    for (int h = 0; h < GH; ++h)
      if (h < heads)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        Vec16<T> t; t.load(x + ((long long)j * heads + h) * C + vi * VN);
        float f[VN]; t.unpack(f);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float w = al[i][h][j];
#pragma unroll
          for (int e = 0; e < VN; ++e) acc[i][e] += w * f[e];
        }
      }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int e = 0; e < VN; ++e) acc[i][e] = fmaxf(acc[i][e] * invh + bias[vi * VN + e], 0.f);
      Vec16<T> o; o.pack(acc[i]); o.store(out + (b * 3 + i) * C + vi * VN);
    }
  }
}

// backward; dynamic smem: [2*heads*C] datt partials + [C] dbias partials
template <typename T>
__global__ void __launch_bounds__(128) gat_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ out, const T* __restrict__ xp,
                                                      const float* __restrict__ alpha_in, const float* __restrict__ att_src,
                                                      const float* __restrict__ att_dst, T* __restrict__ dxp, float* __restrict__ datt_src,
                                                      float* __restrict__ datt_dst, float* __restrict__ dbias, long long B, int heads, int C,
                                                      float slope, uint32_t drop_thr, uint32_t seed_lo, uint32_t seed_hi, float inv_keep) {
  constexpr int VN = Vec16<T>::N;
  extern __shared__ float sm[];
  const int HC = heads * C;
  float* s_src = sm;
  float* s_dst = sm + HC;
  float* s_bias = sm + 2 * HC;
  for (int i = threadIdx.x; i < 2 * HC + C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nvec = C / VN;
  const float invh = 1.f / heads;
  for (long long b = (long long)blockIdx.x * 4 + (threadIdx.x >> 5); b < B; b += (long long)gridDim.x * 4) {
    const T* x = xp + b * 3LL * HC;
    float a_src[3][GH], a_dst[3][GH];
    gat_scores<T>(x, att_src, att_dst, heads, C, lane, a_src, a_dst);
    float al[3][GH][3], dal[3][GH][3];
    {
      const float* ai = alpha_in + b * 3LL * heads * 3;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int h = 0; h < GH; ++h)
#pragma unroll
          for (int j = 0; j < 3; ++j) { al[i][h][j] = h < heads ? ai[(i * heads + h) * 3 + j] : 0.f; dal[i][h][j] = 0.f; }
    }
    // pass 1: g = dout * (out > 0); dbias; dalpha[i][h][j] = invh * sum_c g[i][c] xp[j][h][c]
    for (int vi = lane; vi < nvec; vi += 32) {
      float g[3][VN];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        Vec16<T> a, o; a.load(dout + (b * 3 + i) * C + vi * VN); o.load(out + (b * 3 + i) * C + vi * VN);
        float fo[VN]; a.unpack(g[i]); o.unpack(fo);
#pragma unroll
        for (int e = 0; e < VN; ++e) g[i][e] = fo[e] > 0.f ? g[i][e] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < VN; ++e) atomicAdd(&s_bias[vi * VN + e], g[0][e] + g[1][e] + g[2][e]);
#pragma unroll
      for (int h = 0; h < GH; ++h)
        if (h < heads)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          Vec16<T> t; t.load(x + ((long long)j * heads + h) * C + vi * VN);
          float f[VN]; t.unpack(f);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            float d = 0.f;
#pragma unroll
            for (int e = 0; e < VN; ++e) d += g[i][e] * f[e];
            dal[i][h][j] += d;
          }
        }
    }
    float da_src[3][GH], da_dst[3][GH];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int h = 0; h < GH; ++h) { da_src[j][h] = 0.f; da_dst[j][h] = 0.f; }
    float mk[3][GH][3];      // dropout multiplier of each coefficient (1 without dropout): keep / (1 - p)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int h = 0; h < GH; ++h) {
        const uint32_t rk = drop_thr ? drop_row_key_e(seed_lo, seed_hi, uint32_t((b * heads + h) * 3 + i)) : 0u;
#pragma unroll
        for (int j = 0; j < 3; ++j) mk[i][h][j] = !drop_thr ? 1.f : (drop_keep(rk, uint32_t(j), drop_thr) ? inv_keep : 0.f);
      }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int h = 0; h < GH; ++h) {
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) { dal[i][h][j] = warp_sum(dal[i][h][j]) * invh * mk[i][h][j]; dot += al[i][h][j] * dal[i][h][j]; }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float pre = a_src[j][h] + a_dst[i][h];
          const float de = al[i][h][j] * (dal[i][h][j] - dot) * (pre > 0.f ? 1.f : slope);
          da_src[j][h] += de;
          da_dst[i][h] += de;
        }
      }
    // pass 2: dxp[j][h][c] = invh * sum_i alpha[i][h][j] g[i][c] + da_src[j][h] att_src[h][c] + da_dst[j][h] att_dst[h][c]
    for (int vi = lane; vi < nvec; vi += 32) {
      float g[3][VN];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        Vec16<T> a, o; a.load(dout + (b * 3 + i) * C + vi * VN); o.load(out + (b * 3 + i) * C + vi * VN);
        float fo[VN]; a.unpack(g[i]); o.unpack(fo);
#pragma unroll
        for (int e = 0; e < VN; ++e) g[i][e] = fo[e] > 0.f ? g[i][e] : 0.f;
      }
#pragma unroll
      for (int h = 0; h < GH; ++h) {
        if (h >= heads) break;
        const int hh = h;
        float ws[VN], wd[VN], ps[VN], pd[VN];
#pragma unroll
        for (int e = 0; e < VN; ++e) { ws[e] = att_src[h * C + vi * VN + e]; wd[e] = att_dst[h * C + vi * VN + e]; ps[e] = 0.f; pd[e] = 0.f; }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          Vec16<T> t; t.load(x + ((long long)j * heads + h) * C + vi * VN);
          float f[VN], o[VN]; t.unpack(f);
#pragma unroll
          for (int e = 0; e < VN; ++e) {
            o[e] = invh * (al[0][hh][j] * mk[0][hh][j] * g[0][e] + al[1][hh][j] * mk[1][hh][j] * g[1][e] + al[2][hh][j] * mk[2][hh][j] * g[2][e]) +
                   da_src[j][hh] * ws[e] + da_dst[j][hh] * wd[e];
            ps[e] += da_src[j][hh] * f[e];
            pd[e] += da_dst[j][hh] * f[e];
          }
          Vec16<T> ov; ov.pack(o); ov.store(dxp + b * 3LL * HC + ((long long)j * heads + h) * C + vi * VN);
        }
#pragma unroll
        for (int e = 0; e < VN; ++e) { atomicAdd(&s_src[h * C + vi * VN + e], ps[e]); atomicAdd(&s_dst[h * C + vi * VN + e], pd[e]); }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < HC; i += blockDim.x) { atomicAdd(datt_src + i, s_src[i]); atomicAdd(datt_dst + i, s_dst[i]); }
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dbias + i, s_bias[i]);
}

// ------------------------------------------------------------------------------------------------
// attention over the 3 modality tokens; a head (D dims) spans D/VN consecutive lanes
// ------------------------------------------------------------------------------------------------
template <int LPH>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, int LPH>  // LPH = lanes per head = D / VN (power of two <= 32)
__global__ void __launch_bounds__(128) tok3_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ ctx, float* __restrict__ probs,
                                                       float* __restrict__ avgw, long long B, int heads, int H, float scale, uint32_t drop_thr, uint32_t seed_lo, uint32_t seed_hi, float inv_keep) {
  constexpr int VN = Vec16<T>::N;
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  const T* base = qkv + b * 9LL * H;
  const int nvec = H / VN;
  float wsum[3][3] = {};
  for (int v0 = 0; v0 < nvec; v0 += 32) {
    // lanes past the row (H/VN not a multiple of 32) idle through the loads/stores but still take part in the
    // shuffles; head groups (LPH lanes) never straddle the boundary because nvec = heads * LPH
    const bool live = v0 + lane < nvec;
    const int vi = live ? v0 + lane : 0;
    const int head = vi / LPH;
    float q[3][VN], k[3][VN], vv[3][VN];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      Vec16<T> a, c, d;
      a.load(base + t * 3LL * H + vi * VN); c.load(base + t * 3LL * H + H + vi * VN); d.load(base + t * 3LL * H + 2 * H + vi * VN);
      a.unpack(q[t]); c.unpack(k[t]); d.unpack(vv[t]);
    }
    float p[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < VN; ++e) s += q[i][e] * k[j][e];
        p[i][j] = group_sum<LPH>(s) * scale;
        mx = fmaxf(mx, p[i][j]);
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j) { p[i][j] = expf(p[i][j] - mx); sum += p[i][j]; }
#pragma unroll
      for (int j = 0; j < 3; ++j) p[i][j] /= sum;
      float pd[3] = {p[i][0], p[i][1], p[i][2]};     // nn.MultiheadAttention(dropout=p): the weights are dropped before P V, and the
      if (drop_thr) {                                // returned (head-averaged) weights are the dropped ones (torch functional.py:6647-6665)
        const uint32_t rk = drop_row_key_e(seed_lo, seed_hi, uint32_t((b * heads + head) * 3 + i));
#pragma unroll
        for (int j = 0; j < 3; ++j) pd[j] = drop_keep(rk, uint32_t(j), drop_thr) ? pd[j] * inv_keep : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) if (live && lane % LPH == 0) wsum[i][j] += pd[j];
      float o[VN];
#pragma unroll
      for (int e = 0; e < VN; ++e) o[e] = pd[0] * vv[0][e] + pd[1] * vv[1][e] + pd[2] * vv[2][e];
      Vec16<T> ov; ov.pack(o);
      if (live) ov.store(ctx + (b * 3 + i) * H + vi * VN);
    }
    if (live && lane % LPH == 0) {
      float* po = probs + (b * heads + head) * 9;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) po[i * 3 + j] = p[i][j];
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float s = warp_sum(wsum[i][j]);
      if (lane == 0) avgw[b * 9 + i * 3 + j] = s / heads;
    }
}

template <typename T, int LPH>
__global__ void __launch_bounds__(128) tok3_bwd_kernel(const T* __restrict__ dctx, const float* __restrict__ davgw, const T* __restrict__ qkv,
                                                       const float* __restrict__ probs, T* __restrict__ dqkv, long long B, int heads, int H,
                                                       float scale, uint32_t drop_thr, uint32_t seed_lo, uint32_t seed_hi, float inv_keep) {
  constexpr int VN = Vec16<T>::N;
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  const T* base = qkv + b * 9LL * H;
  T* dbase = dqkv + b * 9LL * H;
  const int nvec = H / VN;
  for (int v0 = 0; v0 < nvec; v0 += 32) {
    const bool live = v0 + lane < nvec;
    const int vi = live ? v0 + lane : 0;
    const int head = vi / LPH;
    float q[3][VN], k[3][VN], vv[3][VN], g[3][VN];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      Vec16<T> a, c, d, e;
      a.load(base + t * 3LL * H + vi * VN); c.load(base + t * 3LL * H + H + vi * VN); d.load(base + t * 3LL * H + 2 * H + vi * VN);
      e.load(dctx + (b * 3 + t) * H + vi * VN);
      a.unpack(q[t]); c.unpack(k[t]); d.unpack(vv[t]); e.unpack(g[t]);
    }
    const float* pi = probs + (b * heads + head) * 9;
    float p[3][3], ds[3][3], mk[3][3];               // mk: dropout multiplier keep / (1 - p) of each weight (1 without dropout)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float dot = 0.f;
      const uint32_t rk = drop_thr ? drop_row_key_e(seed_lo, seed_hi, uint32_t((b * heads + head) * 3 + i)) : 0u;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        p[i][j] = pi[i * 3 + j];
        mk[i][j] = !drop_thr ? 1.f : (drop_keep(rk, uint32_t(j), drop_thr) ? inv_keep : 0.f);
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < VN; ++e) s += g[i][e] * vv[j][e];
        ds[i][j] = (group_sum<LPH>(s) + (davgw ? davgw[b * 9 + i * 3 + j] / heads : 0.f)) * mk[i][j];
        dot += p[i][j] * ds[i][j];
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) ds[i][j] = p[i][j] * (ds[i][j] - dot) * scale;
    }
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      float dq[VN], dk[VN], dv[VN];
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        dq[e] = ds[t][0] * k[0][e] + ds[t][1] * k[1][e] + ds[t][2] * k[2][e];
        dk[e] = ds[0][t] * q[0][e] + ds[1][t] * q[1][e] + ds[2][t] * q[2][e];
        dv[e] = p[0][t] * mk[0][t] * g[0][e] + p[1][t] * mk[1][t] * g[1][e] + p[2][t] * mk[2][t] * g[2][e];
      }
      Vec16<T> a, c, d; a.pack(dq); c.pack(dk); d.pack(dv);
      if (live) { a.store(dbase + t * 3LL * H + vi * VN); c.store(dbase + t * 3LL * H + H + vi * VN); d.store(dbase + t * 3LL * H + 2 * H + vi * VN); }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// gated mix
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) gate_mix_fwd_kernel(const T* __restrict__ att, const T* __restrict__ logits, float* __restrict__ gate,
                                                           T* __restrict__ mixed, long long B, int H) {
  constexpr int VN = Vec16<T>::N;
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  float l[3], mx = -INFINITY, s = 0.f;
#pragma unroll
  for (int m = 0; m < 3; ++m) { l[m] = to_f32(logits[b * 3 + m]); mx = fmaxf(mx, l[m]); }
#pragma unroll
  for (int m = 0; m < 3; ++m) { l[m] = expf(l[m] - mx); s += l[m]; }
#pragma unroll
  for (int m = 0; m < 3; ++m) { l[m] /= s; if (lane == 0) gate[b * 3 + m] = l[m]; }
  for (int vi = lane; vi < H / VN; vi += 32) {
    float o[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) o[e] = 0.f;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      Vec16<T> t; t.load(att + (b * 3 + m) * H + vi * VN); float f[VN]; t.unpack(f);
#pragma unroll
      for (int e = 0; e < VN; ++e) o[e] += l[m] * f[e];
    }
    Vec16<T> ov; ov.pack(o); ov.store(mixed + b * H + vi * VN);
  }
}

template <typename T>
__global__ void __launch_bounds__(128) gate_mix_bwd_kernel(const T* __restrict__ dmixed, const float* __restrict__ dgate_ext, const T* __restrict__ att,
                                                           const float* __restrict__ gate, T* __restrict__ datt, T* __restrict__ dlogits,
                                                           long long B, int H) {
  constexpr int VN = Vec16<T>::N;
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  float gt[3], dg[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int m = 0; m < 3; ++m) gt[m] = gate[b * 3 + m];
  for (int vi = lane; vi < H / VN; vi += 32) {
    Vec16<T> d; d.load(dmixed + b * H + vi * VN); float g[VN]; d.unpack(g);
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      Vec16<T> t; t.load(att + (b * 3 + m) * H + vi * VN); float f[VN], o[VN]; t.unpack(f);
#pragma unroll
      for (int e = 0; e < VN; ++e) { dg[m] += g[e] * f[e]; o[e] = g[e] * gt[m]; }
      Vec16<T> ov; ov.pack(o); ov.store(datt + (b * 3 + m) * H + vi * VN);
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int m = 0; m < 3; ++m) { dg[m] = warp_sum(dg[m]) + (dgate_ext ? dgate_ext[b * 3 + m] : 0.f); dot += gt[m] * dg[m]; }
  if (lane < 3) dlogits[b * 3 + lane] = from_f32<T>(gt[lane] * (dg[lane] - dot));
}

// ------------------------------------------------------------------------------------------------
// late fusion combine (tiny: a single block)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void late_fwd_kernel(const T* lt, const T* la, const T* lv, const float* w3, float* wsoft, T* fused, long long n) {
  const float mx = fmaxf(w3[0], fmaxf(w3[1], w3[2]));
  const float e0 = expf(w3[0] - mx), e1 = expf(w3[1] - mx), e2 = expf(w3[2] - mx);
  const float s = e0 + e1 + e2;
  const float w0 = e0 / s, w1 = e1 / s, w2 = e2 / s;
  if (blockIdx.x == 0 && threadIdx.x == 0) { wsoft[0] = w0; wsoft[1] = w1; wsoft[2] = w2; }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    fused[i] = from_f32<T>(w0 * to_f32(lt[i]) + w1 * to_f32(la[i]) + w2 * to_f32(lv[i]));
}

template <typename T>
__global__ void __launch_bounds__(256) late_bwd_kernel(const T* dfused, const T* lt, const T* la, const T* lv, const float* wsoft,
                                                       const float* dwsoft_ext, T* dlt, T* dla, T* dlv, float* dw3, long long n) {
  __shared__ float red[3][8];
  const float w0 = wsoft[0], w1 = wsoft[1], w2 = wsoft[2];
  float d0 = 0.f, d1 = 0.f, d2 = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float g = to_f32(dfused[i]);
    d0 += g * to_f32(lt[i]); d1 += g * to_f32(la[i]); d2 += g * to_f32(lv[i]);
    dlt[i] = from_f32<T>(g * w0); dla[i] = from_f32<T>(g * w1); dlv[i] = from_f32<T>(g * w2);
  }
  d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = d0; red[1][threadIdx.x >> 5] = d1; red[2][threadIdx.x >> 5] = d2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t[3] = {0.f, 0.f, 0.f};
    for (int w = 0; w < 8; ++w) { t[0] += red[0][w]; t[1] += red[1][w]; t[2] += red[2][w]; }
    if (dwsoft_ext) { t[0] += dwsoft_ext[0]; t[1] += dwsoft_ext[1]; t[2] += dwsoft_ext[2]; }
    const float dot = w0 * t[0] + w1 * t[1] + w2 * t[2];
    dw3[0] += w0 * (t[0] - dot); dw3[1] += w1 * (t[1] - dot); dw3[2] += w2 * (t[2] - dot);
  }
}

// ------------------------------------------------------------------------------------------------
// The dropout epoch (common.cuh) is folded into the seed like in dropout_kernel: a mask generated inside a captured training
// step is therefore re-drawn by every replay (the host-side offset of ModalityDropout.sample_mask is frozen at capture).
__global__ void modality_mask_kernel(float* mask, long long B, float rate, uint64_t seed, uint64_t offset) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  seed ^= uint64_t(g_drop_epoch) << 32;
  float k[3];
  bool any = false;
#pragma unroll
  for (int m = 0; m < 3; ++m) { k[m] = rng_uniform(seed, offset + 4 * b + m) > rate ? 1.f : 0.f; any |= k[m] != 0.f; }
  if (!any) {
    int pick = int(rng_uniform(seed, offset + 4 * b + 3) * 3.f);
    pick = pick > 2 ? 2 : pick;
    k[pick] = 1.f;
  }
#pragma unroll
  for (int m = 0; m < 3; ++m) mask[b * 3 + m] = k[m];
}

// in-place inverted dropout of an [M, N] activation with the (seed, row, column) mask the GEMM epilogue generates
// (the route for outputs the tcgen05 epilogue does not cover: fp32 parity mode, ragged N)
template <typename T>
__global__ void dropout_rowcol_kernel(T* __restrict__ x, long long ldx, long long M, long long N, uint32_t thr, uint32_t seed_lo, uint32_t seed_hi,
                                      float inv_keep) {
  const long long total = M * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / N, c = i - r * N;
    T* e = x + r * ldx + c;
    *e = drop_keep(drop_row_key_e(seed_lo, seed_hi, uint32_t(r)), uint32_t(c), thr) ? from_f32<T>(to_f32(*e) * inv_keep) : from_f32<T>(0.f);
  }
}

template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, float p, float inv_keep, uint64_t seed, uint64_t offset) {
  constexpr int VN = Vec16<T>::N;
  const long long nvec = n / VN;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    Vec16<T> t; t.load(x + i * VN); float f[VN]; t.unpack(f);
#pragma unroll
    for (int e = 0; e < VN; ++e) f[e] = rng_uniform(seed ^ (uint64_t(g_drop_epoch) << 32), offset + i * VN + e) >= p ? f[e] * inv_keep : 0.f;
    t.pack(f); t.store(y + i * VN);
  }
}

static inline int ew_grid2(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}

#define DISPATCH_DTYPE(dtype, T, ...)                                              \
  if ((dtype) == B200F_F32) { using T = float; __VA_ARGS__ }                       \
  else if ((dtype) == B200F_BF16) { using T = bf16; __VA_ARGS__ }                  \
  else return fail(B200F_ERR_DTYPE, "unknown dtype %d", int(dtype));

B200F_DEFINE_EPOCH_HOOK(smallops)

}  // namespace b200f

using namespace b200f;

extern "C" {

int b200f_gat_fwd(const void* xp, const float* att_src, const float* att_dst, const float* bias, void* out, float* alpha, int64_t B,
                  int32_t heads, int32_t C, float slope, float dropout_p, uint32_t seed_lo, uint32_t seed_hi, int32_t dtype, void* stream) {
  if (B == 0) return B200F_OK;
  B200F_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, B200F_ERR_SHAPE, "gat: dropout_p=%f", dropout_p);
  B200F_REQUIRE(heads >= 1 && heads <= GH, B200F_ERR_UNSUPPORTED, "gat: heads=%d (max %d)", heads, GH);
  DISPATCH_DTYPE(dtype, T, {
    B200F_REQUIRE(C % Vec16<T>::N == 0, B200F_ERR_SHAPE, "gat: C=%d must be a multiple of %d", C, Vec16<T>::N);
    B200F_REQUIRE(aligned16(xp) && aligned16(out), B200F_ERR_ALIGN, "gat: alignment");
    gat_fwd_kernel<T><<<(unsigned)((B + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const T*>(xp), att_src, att_dst, bias,
                                                                                               static_cast<T*>(out), alpha, B, heads, C, slope,
                                                                                               drop_threshold(dropout_p), seed_lo, seed_hi, 1.f / (1.f - dropout_p));
  })
  return check_launch("gat_fwd");
}

int b200f_gat_bwd(const void* dout, const void* out, const void* xp, const float* alpha, const float* att_src, const float* att_dst, void* dxp,
                  float* datt_src, float* datt_dst, float* dbias, int64_t B, int32_t heads, int32_t C, float slope, float dropout_p,
                  uint32_t seed_lo, uint32_t seed_hi, int32_t dtype, void* stream) {
  if (B == 0) return B200F_OK;
  B200F_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, B200F_ERR_SHAPE, "gat: dropout_p=%f", dropout_p);
  B200F_REQUIRE(heads >= 1 && heads <= GH, B200F_ERR_UNSUPPORTED, "gat: heads=%d (max %d)", heads, GH);
  const size_t smem = (2 * (size_t)heads * C + C) * sizeof(float);
  B200F_REQUIRE(smem <= 160 * 1024, B200F_ERR_SHAPE, "gat: heads*C too large");
  DISPATCH_DTYPE(dtype, T, {
    B200F_REQUIRE(C % Vec16<T>::N == 0, B200F_ERR_SHAPE, "gat: C=%d must be a multiple of %d", C, Vec16<T>::N);
    B200F_REQUIRE(aligned16(xp) && aligned16(out) && aligned16(dout) && aligned16(dxp), B200F_ERR_ALIGN, "gat: alignment");
    if (smem > 48 * 1024) B200F_CHECK_CUDA(cudaFuncSetAttribute(gat_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (B + 3) / 4;
    if (blocks > 2LL * num_sms()) blocks = 2LL * num_sms();
    gat_bwd_kernel<T><<<(unsigned)blocks, 128, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const T*>(dout), static_cast<const T*>(out), static_cast<const T*>(xp), alpha, att_src, att_dst, static_cast<T*>(dxp), datt_src,
        datt_dst, dbias, B, heads, C, slope, drop_threshold(dropout_p), seed_lo, seed_hi, 1.f / (1.f - dropout_p));
  })
  return check_launch("gat_bwd");
}

#define TOK3_DISPATCH(KERNEL, ...)                                                                                    \
  DISPATCH_DTYPE(dtype, T, {                                                                                          \
    constexpr int VN = Vec16<T>::N;                                                                                   \
    const int D = H / heads;                                                                                          \
    B200F_REQUIRE(heads > 0 && H % heads == 0 && D % VN == 0, B200F_ERR_SHAPE, "tok3: H=%d heads=%d", H, heads); \
    const int lph = D / VN;                                                                                           \
    const unsigned grid = (unsigned)((B + 3) / 4);                                                                    \
    cudaStream_t st = static_cast<cudaStream_t>(stream);                                                              \
    switch (lph) {                                                                                                    \
      case 1: KERNEL<T, 1><<<grid, 128, 0, st>>>(__VA_ARGS__); break;                                                 \
      case 2: KERNEL<T, 2><<<grid, 128, 0, st>>>(__VA_ARGS__); break;                                                 \
      case 4: KERNEL<T, 4><<<grid, 128, 0, st>>>(__VA_ARGS__); break;                                                 \
      case 8: KERNEL<T, 8><<<grid, 128, 0, st>>>(__VA_ARGS__); break;                                                 \
      case 16: KERNEL<T, 16><<<grid, 128, 0, st>>>(__VA_ARGS__); break;                                               \
      case 32: KERNEL<T, 32><<<grid, 128, 0, st>>>(__VA_ARGS__); break;                                               \
      default: return fail(B200F_ERR_SHAPE, "tok3: head dim %d unsupported", D);                                      \
    }                                                                                                                 \
  })

int b200f_tok3_attn_fwd(const void* qkv, void* ctx, float* probs, float* avgw, int64_t B, int32_t heads, int32_t H, float scale, float dropout_p,
                        uint32_t seed_lo, uint32_t seed_hi, int32_t dtype, void* stream) {
  if (B == 0) return B200F_OK;
  B200F_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, B200F_ERR_SHAPE, "tok3: dropout_p=%f", dropout_p);
  B200F_REQUIRE(aligned16(qkv) && aligned16(ctx), B200F_ERR_ALIGN, "tok3: alignment");
  TOK3_DISPATCH(tok3_fwd_kernel, static_cast<const T*>(qkv), static_cast<T*>(ctx), probs, avgw, B, heads, H, scale, drop_threshold(dropout_p), seed_lo,
                seed_hi, 1.f / (1.f - dropout_p))
  return check_launch("tok3_fwd");
}

int b200f_tok3_attn_bwd(const void* dctx, const float* davgw, const void* qkv, const float* probs, void* dqkv, int64_t B, int32_t heads, int32_t H,
                        float scale, float dropout_p, uint32_t seed_lo, uint32_t seed_hi, int32_t dtype, void* stream) {
  if (B == 0) return B200F_OK;
  B200F_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, B200F_ERR_SHAPE, "tok3: dropout_p=%f", dropout_p);
  B200F_REQUIRE(aligned16(qkv) && aligned16(dctx) && aligned16(dqkv), B200F_ERR_ALIGN, "tok3: alignment");
  TOK3_DISPATCH(tok3_bwd_kernel, static_cast<const T*>(dctx), davgw, static_cast<const T*>(qkv), probs, static_cast<T*>(dqkv), B, heads, H, scale,
                drop_threshold(dropout_p), seed_lo, seed_hi, 1.f / (1.f - dropout_p))
  return check_launch("tok3_bwd");
}

int b200f_gate_mix_fwd(const void* att, const void* logits, float* gate, void* mixed, int64_t B, int32_t H, int32_t dtype, void* stream) {
  if (B == 0) return B200F_OK;
  DISPATCH_DTYPE(dtype, T, {
    B200F_REQUIRE(H % Vec16<T>::N == 0 && aligned16(att) && aligned16(mixed), B200F_ERR_ALIGN, "gate_mix: alignment");
    gate_mix_fwd_kernel<T><<<(unsigned)((B + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const T*>(att), static_cast<const T*>(logits), gate,
                                                                                                    static_cast<T*>(mixed), B, H);
  })
  return check_launch("gate_mix_fwd");
}

int b200f_gate_mix_bwd(const void* dmixed, const float* dgate_ext, const void* att, const float* gate, void* datt, void* dlogits, int64_t B, int32_t H,
                       int32_t dtype, void* stream) {
  if (B == 0) return B200F_OK;
  DISPATCH_DTYPE(dtype, T, {
    B200F_REQUIRE(H % Vec16<T>::N == 0 && aligned16(att) && aligned16(dmixed) && aligned16(datt), B200F_ERR_ALIGN, "gate_mix: alignment");
    gate_mix_bwd_kernel<T><<<(unsigned)((B + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const T*>(dmixed), dgate_ext, static_cast<const T*>(att),
                                                                                                    gate, static_cast<T*>(datt), static_cast<T*>(dlogits), B, H);
  })
  return check_launch("gate_mix_bwd");
}

int b200f_late_combine_fwd(const void* lt, const void* la, const void* lv, const float* w3, float* wsoft, void* fused, int64_t B, int32_t E, int32_t dtype,
                           void* stream) {
  if (B == 0) return B200F_OK;
  DISPATCH_DTYPE(dtype, T, {
    late_fwd_kernel<T><<<ew_grid2(B * E, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const T*>(lt), static_cast<const T*>(la), static_cast<const T*>(lv),
                                                                                             w3, wsoft, static_cast<T*>(fused), B * E);
  })
  return check_launch("late_fwd");
}

int b200f_late_combine_bwd(const void* dfused, const void* lt, const void* la, const void* lv, const float* wsoft, const float* dwsoft_ext, void* dlt,
                           void* dla, void* dlv, float* dw3, int64_t B, int32_t E, int32_t dtype, void* stream) {
  if (B == 0) return B200F_OK;
  DISPATCH_DTYPE(dtype, T, {
    late_bwd_kernel<T><<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const T*>(dfused), static_cast<const T*>(lt), static_cast<const T*>(la),
                                                                          static_cast<const T*>(lv), wsoft, dwsoft_ext, static_cast<T*>(dlt), static_cast<T*>(dla),
                                                                          static_cast<T*>(dlv), dw3, B * E);
  })
  return check_launch("late_bwd");
}

int b200f_modality_mask(float* mask, int64_t B, float rate, uint64_t seed, uint64_t offset, void* stream) {
  if (B == 0) return B200F_OK;
  modality_mask_kernel<<<(unsigned)((B + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, B, rate, seed, offset);
  return check_launch("modality_mask");
}

int b200f_dropout_rowcol(void* x, int64_t ldx, int64_t M, int64_t N, float p, uint32_t seed_lo, uint32_t seed_hi, int32_t dtype, void* stream) {
  if (M == 0 || N == 0 || p <= 0.f) return B200F_OK;
  B200F_REQUIRE(p < 1.f, B200F_ERR_SHAPE, "dropout: p=%f", p);
  const long long total = (long long)M * N;
  long long blocks = (total + 255) / 256;
  if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
  DISPATCH_DTYPE(dtype, T, {
    dropout_rowcol_kernel<T><<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<T*>(x), ldx, M, N, drop_threshold(p), seed_lo, seed_hi,
                                                                                             1.f / (1.f - p));
  })
  return check_launch("dropout_rowcol");
}

int b200f_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, uint64_t offset, int32_t dtype, void* stream) {
  if (n == 0) return B200F_OK;
  B200F_REQUIRE(p >= 0.f && p < 1.f, B200F_ERR_SHAPE, "dropout: p=%f", p);
  DISPATCH_DTYPE(dtype, T, {
    B200F_REQUIRE(n % Vec16<T>::N == 0 && aligned16(x) && aligned16(y), B200F_ERR_ALIGN, "dropout: alignment");
    dropout_kernel<T><<<ew_grid2(n / Vec16<T>::N, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const T*>(x), static_cast<T*>(y), n, p,
                                                                                                      1.f / (1.f - p), seed, offset);
  })
  return check_launch("dropout");
}

}  // extern "C"
