// tcgen05 / TMEM / TMA flash attention for bf16 (placeholder until the kernels land: reports "unsupported"
// so that AUTO falls to the CUDA-core path and an explicit AGA_ATTN_TCGEN05 request fails loudly).
#include "aga_common.cuh"
#include "attn_common.cuh"

namespace aga {
bool attn_tc_supported(const aga_attn_params&) { return false; }
size_t attn_tc_fwd_workspace(const aga_attn_params&) { return 0; }
int attn_tc_fwd(const aga_attn_params&, void*, cudaStream_t) { return AGA_ERR_UNSUPPORTED; }
size_t attn_tc_bwd_workspace(const aga_attn_params&) { return 0; }
int attn_tc_bwd(const aga_attn_bwd_params&, void*, cudaStream_t) { return AGA_ERR_UNSUPPORTED; }
}  // namespace aga
