// Internal interface between the attention C ABI (attn_api.cu) and its two implementations.
#pragma once
#include "aga_common.cuh"

namespace aga {

// CUDA-core fp32-accumulate path (attn_simt.cu): fp32 or bf16 in/out, any Tq/Tk, causal, column export.
int attn_simt_fwd(const aga_attn_params& p, cudaStream_t s);
// single-query (decoding step) forward: Tq == 1, no mask / export; a cluster of CTAs per (hypothesis, head)
bool attn_decode_shape(const aga_attn_params& p);
int attn_decode_fwd(const aga_attn_params& p, cudaStream_t s);
size_t attn_simt_bwd_workspace(const aga_attn_params& p);
int attn_simt_bwd(const aga_attn_bwd_params& bp, void* ws, cudaStream_t s);

// tcgen05 / TMEM / TMA path (attn_tc.cu): bf16 in/out.
bool attn_tc_supported(const aga_attn_params& p);
size_t attn_tc_fwd_workspace(const aga_attn_params& p);
int attn_tc_fwd(const aga_attn_params& p, void* ws, cudaStream_t s);
bool attn_tc_bwd_supported(const aga_attn_params& p);
size_t attn_tc_bwd_workspace(const aga_attn_params& p);
int attn_tc_bwd(const aga_attn_bwd_params& bp, void* ws, cudaStream_t s);

}  // namespace aga
