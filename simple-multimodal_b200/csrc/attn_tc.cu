// placeholder until the tcgen05 attention kernels land: bf16 attention runs on the CUDA-core kernels
#include "common.cuh"
namespace b200f {
int attn_fwd_simt_dispatch(const b200f_attn_args& a, cudaStream_t st);
int attn_bwd_simt_dispatch(const b200f_attn_args& a, cudaStream_t st);
int attn_fwd_tc(const b200f_attn_args& a, cudaStream_t st) { return attn_fwd_simt_dispatch(a, st); }
int attn_bwd_tc(const b200f_attn_args& a, cudaStream_t st) { return attn_bwd_simt_dispatch(a, st); }
}  // namespace b200f
