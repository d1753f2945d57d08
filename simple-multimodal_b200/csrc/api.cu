// C-ABI entry points that need no kernel of their own: version, error buffer, device check,
// GEMM dtype dispatch, debug knobs.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace b200f {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
char* err_buf() { return g_err; }

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {
  static int n[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!n[dev]) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

int gemm_bf16_tc(const b200f_gemm_args& a, cudaStream_t st);
int gemm_f32_simt(const b200f_gemm_args& a, cudaStream_t st);
int gemm_bf16_simt(const b200f_gemm_args& a, cudaStream_t st);
bool gemm_tc_eligible(const b200f_gemm_args& a);
bool gemm_tc_colsum_fused(const b200f_gemm_args& a);
extern uint32_t g_dbg_mn_lbo, g_dbg_mn_sbo, g_dbg_mn_kadv;
extern bool g_dbg_disable_pair;
extern bool g_dbg_six_stages;
extern bool g_dbg_no_tma_store;
extern bool g_dbg_no_tma_store_aux;
extern bool g_dbg_late_aux;
extern int g_dbg_colsum;
extern bool g_dbg_no_spec_epi;
extern bool g_dbg_no_ln_tma;
extern int g_attn_fwd_variant;
extern int g_attn_bwd_ds_route;
extern int g_attn_bwd_variant;
extern int g_infonce_variant;
void attn_tc_epoch(uint32_t v, int add, cudaStream_t st);
void attn_narrow_epoch(uint32_t v, int add, cudaStream_t st);
extern int g_attn_narrow;
void attn_simt_epoch(uint32_t v, int add, cudaStream_t st);
void gemm_tc_epoch(uint32_t v, int add, cudaStream_t st);
void smallops_epoch(uint32_t v, int add, cudaStream_t st);

}  // namespace b200f

extern "C" {

int b200f_version(void) { return 100; }

unsigned long long b200f_launch_count(void) { return b200f::g_launches; }

const char* b200f_last_error(void) { return b200f::err_buf(); }

int b200f_device_supported(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return b200f::fail(B200F_ERR_CUDA, "cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
  if (prop.major != 10) return b200f::fail(B200F_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is sm_100a only", device, prop.major, prop.minor);
  return B200F_OK;
}

int b200f_gemm(const b200f_gemm_args* a, void* stream) {
  if (!a) return b200f::fail(B200F_ERR_SHAPE, "gemm: null args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->colsum && (a->flags & (B200F_EPI_OUT_F32 | B200F_EPI_ACCUM)) && a->dtype == B200F_BF16)
    return b200f::fail(B200F_ERR_UNSUPPORTED, "gemm: colsum needs the output in the operand dtype");
  if (!(a->dropout_p >= 0.f && a->dropout_p < 1.f)) return b200f::fail(B200F_ERR_SHAPE, "gemm: dropout_p=%f not in [0,1)", a->dropout_p);
  const bool bits = a->sign_bits_out || a->sign_bits;
  if (bits && !(a->dtype == B200F_BF16 && b200f::gemm_tc_eligible(*a) && b200f::gemm_tc_colsum_fused(*a)))
    return b200f::fail(B200F_ERR_UNSUPPORTED, "gemm: sign_bits / sign_bits_out exist on the bf16 tcgen05 path only (bf16 output, N %% 64 == 0, aligned rows)");
  const bool extras = a->colsum || a->dropout_p > 0.f || bits;
  if (a->dropout_p > 0.f && (a->flags & (B200F_EPI_OUT_F32 | B200F_EPI_ACCUM)) && a->dtype == B200F_BF16)
    return b200f::fail(B200F_ERR_UNSUPPORTED, "gemm: dropout needs the output in the operand dtype");
  if (a->dtype == B200F_BF16 && b200f::gemm_tc_eligible(*a) && (!extras || b200f::gemm_tc_colsum_fused(*a)))
    return b200f::gemm_bf16_tc(*a, st);                       // dropout / column sums (if any) come out of the epilogue
  b200f_gemm_args plain = *a;                                // other paths: GEMM, then one pass each over the stored C
  plain.colsum = nullptr;
  plain.dropout_p = 0.f;
  int rc;
  if (a->dtype == B200F_BF16) rc = b200f::gemm_tc_eligible(plain) ? b200f::gemm_bf16_tc(plain, st) : b200f::gemm_bf16_simt(plain, st);
  else if (a->dtype == B200F_F32) rc = b200f::gemm_f32_simt(plain, st);
  else return b200f::fail(B200F_ERR_DTYPE, "gemm: unknown dtype %d", a->dtype);
  if (rc) return rc;
  if (a->dropout_p > 0.f && (rc = b200f_dropout_rowcol(a->C, a->ldc, a->M, a->N, a->dropout_p, a->drop_seed_lo, a->drop_seed_hi, a->dtype, stream))) return rc;
  if (!a->colsum) return B200F_OK;
  return b200f_colsum_accum(a->C, a->ldc, a->colsum, a->M, a->N, a->dtype, stream);
}

// Dropout epoch (see csrc/common.cuh): add != 0 -> epoch += value, else epoch = value; stream-ordered, capturable in a CUDA graph.
int b200f_dropout_epoch(uint32_t value, int32_t add, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  b200f::attn_tc_epoch(value, add, st);
  b200f::attn_narrow_epoch(value, add, st);
  b200f::attn_simt_epoch(value, add, st);
  b200f::gemm_tc_epoch(value, add, st);
  b200f::smallops_epoch(value, add, st);
  return b200f::check_launch("dropout_epoch");
}

// Debug / A-B switches (not part of the product API, not declared in include/b200_fusion.h; B200F_DEBUG_SET="key=value,..." sets
// them from the environment when the Python binding loads the library).  Every value 0 restores the default unless noted.
//   0-2  MN-major UMMA descriptor geometry (LBO / SBO / K advance in bytes)      3  no CTA-pair GEMM kernel
//   4    attention forward: 1 = one tile per CTA, 2 = persistent two-tile        5  attention backward: 1 = one tile per CTA
//   6    InfoNCE: GEMM + row-kernel route instead of the fused kernels           7  6-stage / one-staging-tile pair GEMM
//   8    no bulk tensor stores in the GEMM epilogue                              9  LayerNorm without the TMA-staged kernels
//   10   narrow-side attention: 0 = off (tcgen05 tiles), 1 = default, 2 = v1 backward kernels
//   11   LDS + STG copy-out for GEMMs with a residual / mask block               12 aux block prefetched one column block ahead
//   13   column sums: 1 = computed, REDs skipped; 2 = skipped (timing only)      14 generic (run-time flag) GEMM epilogue everywhere
//   15   attention backward dS route: 1 = on (default 0, see csrc/attn_tc.cu)
int b200f_debug_set(int key, unsigned value) {
  switch (key) {
    case 0: b200f::g_dbg_mn_lbo = value; break;
    case 1: b200f::g_dbg_mn_sbo = value; break;
    case 2: b200f::g_dbg_mn_kadv = value; break;
    case 3: b200f::g_dbg_disable_pair = value != 0; break;
    case 4: b200f::g_attn_fwd_variant = int(value); break;
    case 5: b200f::g_attn_bwd_variant = int(value); break;
    case 6: b200f::g_infonce_variant = int(value); break;
    case 7: b200f::g_dbg_six_stages = value != 0; break;
    case 8: b200f::g_dbg_no_tma_store = value != 0; break;
    case 9: b200f::g_dbg_no_ln_tma = value != 0; break;
    case 10: b200f::g_attn_narrow = int(value); break;
    case 11: b200f::g_dbg_no_tma_store_aux = value != 0; break;
    case 12: b200f::g_dbg_late_aux = value != 0; break;
    case 13: b200f::g_dbg_colsum = int(value); break;
    case 14: b200f::g_dbg_no_spec_epi = value != 0; break;
    case 15: b200f::g_attn_bwd_ds_route = int(value); break;
    default: return b200f::fail(B200F_ERR_UNSUPPORTED, "unknown debug key %d", key);
  }
  return B200F_OK;
}

}  // extern "C"

namespace b200f {
int attn_fwd_simt_dispatch(const b200f_attn_args& a, cudaStream_t st);
int attn_bwd_simt_dispatch(const b200f_attn_args& a, cudaStream_t st);
int attn_fwd_tc(const b200f_attn_args& a, cudaStream_t st);
int attn_bwd_tc(const b200f_attn_args& a, cudaStream_t st);
int64_t attn_bwd_ws_bytes(const b200f_attn_args& a);
int attn_narrow_kind(const b200f_attn_args& a);
int attn_narrow_pool_parts(const b200f_attn_args& a);
int attn_fwd_narrow(const b200f_attn_args& a, cudaStream_t st);
int attn_bwd_narrow(const b200f_attn_args& a, cudaStream_t st);
bool g_force_simt_attention = false;

static int attn_check(const b200f_attn_args* a) {
  if (!a) return fail(B200F_ERR_SHAPE, "attention: null args");
  B200F_REQUIRE(a->B >= 0 && a->H > 0 && a->Lq > 0 && a->Lk > 0 && a->D > 0, B200F_ERR_SHAPE, "attention: bad shape B=%d H=%d Lq=%d Lk=%d D=%d", a->B,
                a->H, a->Lq, a->Lk, a->D);
  B200F_REQUIRE(a->dtype == B200F_F32 || a->dtype == B200F_BF16, B200F_ERR_DTYPE, "attention: unknown dtype %d", a->dtype);
  B200F_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, B200F_ERR_SHAPE, "attention: dropout_p=%f not in [0,1)", a->dropout_p);
  return B200F_OK;
}
}  // namespace b200f

extern "C" {

int32_t b200f_attn_pool_parts(const b200f_attn_args* a) {
  if (!a || a->dtype != B200F_BF16 || a->D != 64 || b200f::g_force_simt_attention) return 0;
  const int n = b200f::attn_narrow_pool_parts(*a);
  return n ? n : (a->Lq + 31) / 32;                       // tcgen05 kernels: one partial per 32-row tile
}

int b200f_attn_fwd(const b200f_attn_args* a, void* stream) {
  int rc = b200f::attn_check(a);
  if (rc || a->B == 0) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->dtype == B200F_BF16 && a->D == 64 && !b200f::g_force_simt_attention)      // one side <= 32 rows: the HBM-bound narrow kernels
    return b200f::attn_narrow_kind(*a) ? b200f::attn_fwd_narrow(*a, st) : b200f::attn_fwd_tc(*a, st);
  if (a->pool_sum) return b200f::fail(B200F_ERR_UNSUPPORTED, "attention: pool_sum needs the bf16 / head-dim-64 kernels");
  return b200f::attn_fwd_simt_dispatch(*a, st);
}

int64_t b200f_attn_bwd_ws_bytes(const b200f_attn_args* a) {
  if (!a || b200f::attn_narrow_kind(*a)) return 0;
  return b200f::attn_bwd_ws_bytes(*a);
}

int b200f_attn_bwd(const b200f_attn_args* a, void* stream) {
  int rc = b200f::attn_check(a);
  if (rc || a->B == 0) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool tc = a->dtype == B200F_BF16 && a->D == 64 && !b200f::g_force_simt_attention;
  if (tc && b200f::attn_narrow_kind(*a)) return b200f::attn_bwd_narrow(*a, st);    // sums the bias gradients itself
  rc = tc ? b200f::attn_bwd_tc(*a, st) : b200f::attn_bwd_simt_dispatch(*a, st);
  if (rc || (tc && b200f::g_attn_bwd_variant == 0)) return rc;      // the persistent tcgen05 kernels sum the bias gradients in their epilogue
  const int64_t W = (int64_t)a->H * a->D;
  if (a->dbq && (rc = b200f_colsum_accum(a->dQ, a->lddq, a->dbq, (int64_t)a->B * a->Lq, W, a->dtype, stream))) return rc;
  if (a->dbk && (rc = b200f_colsum_accum(a->dK, a->lddk, a->dbk, (int64_t)a->B * a->Lk, W, a->dtype, stream))) return rc;
  if (a->dbv && (rc = b200f_colsum_accum(a->dV, a->lddv, a->dbv, (int64_t)a->B * a->Lk, W, a->dtype, stream))) return rc;
  return B200F_OK;
}

// debug: route bf16 attention through the CUDA-core kernels (for A/B comparison against the tcgen05 path)
int b200f_debug_force_simt_attention(int on) { b200f::g_force_simt_attention = on != 0; return B200F_OK; }

}  // extern "C"
