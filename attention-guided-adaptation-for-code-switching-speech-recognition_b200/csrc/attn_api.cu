// C ABI of the attention core: validation, implementation choice, export pre-fill.
#include "aga_common.cuh"
#include "attn_common.cuh"

#include <algorithm>

namespace aga {
namespace {

// Causal kernels skip key tiles above the diagonal, so exported entries there are set up front:
// -inf is what the reference's `qk + mask` holds (whisper/model.py:103) and exp(-inf - lse) = 0 for probs.
__global__ void __launch_bounds__(256)
export_prefill_kernel(float* __restrict__ buf, const uint8_t* __restrict__ head_sel, int H, int64_t per_head,
                      int64_t total, float value) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int h = int((i / per_head) % H);
    if (head_sel && !head_sel[h]) continue;
    buf[i] = value;
  }
}

int validate(const aga_attn_params* p) {
  if (!p || !p->q || !p->k || !p->v || !p->out) return AGA_ERR_INVALID_ARGUMENT;
  if (p->B <= 0 || p->H <= 0 || p->Tq <= 0 || p->Tk <= 0) return AGA_ERR_INVALID_ARGUMENT;
  if (p->dtype != AGA_F32 && p->dtype != AGA_BF16) return AGA_ERR_INVALID_ARGUMENT;
  if (p->causal && p->Tq != p->Tk) return AGA_ERR_INVALID_ARGUMENT;
  if (p->kv_len && (p->causal || p->export_kind != AGA_EXPORT_NONE)) return AGA_ERR_UNSUPPORTED;
  if ((p->guided_pattern != nullptr) != (p->guided_part != nullptr)) return AGA_ERR_INVALID_ARGUMENT;
  if (p->guided_part && (!p->causal || p->Tq > 128 || p->Tq < 3)) return AGA_ERR_UNSUPPORTED;
  if (p->B > 65535 || p->H > 65535) return AGA_ERR_UNSUPPORTED;
  if (p->export_kind != AGA_EXPORT_NONE) {
    if (p->export_kind != AGA_EXPORT_LOGITS && p->export_kind != AGA_EXPORT_PROBS) return AGA_ERR_INVALID_ARGUMENT;
    if (!p->export_buf || p->export_lo < 0 || p->export_hi > p->Tk || p->export_lo >= p->export_hi)
      return AGA_ERR_INVALID_ARGUMENT;
    if (p->export_kind == AGA_EXPORT_PROBS && !p->lse) return AGA_ERR_INVALID_ARGUMENT;
  }
  // 16-byte vector access on every row of every head
  const int64_t vec = p->dtype == AGA_BF16 ? 8 : 4;
  const int64_t strides[] = {p->q_stride_b, p->q_stride_t, p->k_stride_b, p->k_stride_t,
                             p->v_stride_b, p->v_stride_t, p->o_stride_b, p->o_stride_t};
  for (int64_t s : strides)
    if (s % vec != 0) return AGA_ERR_UNSUPPORTED;
  const void* ptrs[] = {p->q, p->k, p->v, p->out};
  for (const void* x : ptrs)
    if (reinterpret_cast<uintptr_t>(x) & 15) return AGA_ERR_UNSUPPORTED;
  return AGA_OK;
}

bool use_tc(const aga_attn_params& p) {
  if (p.impl == AGA_ATTN_SIMT) return false;
  if (p.dtype != AGA_BF16) return false;
  return attn_tc_supported(p);
}
bool use_tc_bwd(const aga_attn_params& p) {
  if (p.impl == AGA_ATTN_SIMT) return false;
  if (p.dtype != AGA_BF16) return false;
  return attn_tc_bwd_supported(p);
}

}  // namespace
}  // namespace aga

using namespace aga;

extern "C" int aga_attn_fwd_workspace_bytes(const aga_attn_params* p, size_t* bytes) {
  int st = validate(p);
  if (st != AGA_OK) return st;
  if (!bytes) return AGA_ERR_INVALID_ARGUMENT;
  // (the CUDA-core forward serves kv_len too — the fp32 decoding step; the backward and the guided epilogue are tcgen05 only)
  if ((p->impl == AGA_ATTN_TCGEN05 || p->guided_part) && !use_tc(*p)) return AGA_ERR_UNSUPPORTED;
  *bytes = (use_tc(*p) && !attn_decode_shape(*p)) ? attn_tc_fwd_workspace(*p) : 0;
  return AGA_OK;
}

extern "C" int aga_attn_fwd(const aga_attn_params* p, void* workspace, size_t workspace_bytes, void* stream) {
  size_t need = 0;
  int st = aga_attn_fwd_workspace_bytes(p, &need);
  if (st != AGA_OK) return st;
  if (need > 0 && (!workspace || workspace_bytes < need)) return AGA_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (attn_decode_shape(*p)) return attn_decode_fwd(*p, s);
  // (the tcgen05 kernel never skips key tile 0, so columns below 128 are always written by the kernel itself)
  const bool kernel_writes_all = use_tc(*p) && p->export_hi <= 128;
  if (p->causal && p->export_kind != AGA_EXPORT_NONE && !kernel_writes_all) {
    const int64_t per_head = int64_t(p->Tq) * (p->export_hi - p->export_lo);
    const int64_t total = per_head * p->H * p->B;
    const unsigned gx = unsigned(std::min<int64_t>((total + 255) / 256, 148 * 8));
    export_prefill_kernel<<<gx ? gx : 1, 256, 0, s>>>(p->export_buf, p->head_sel, p->H, per_head, total, -INFINITY);
    AGA_AFTER_LAUNCH();
  }
  return use_tc(*p) ? attn_tc_fwd(*p, workspace, s) : attn_simt_fwd(*p, s);
}

extern "C" int aga_attn_bwd_workspace_bytes(const aga_attn_bwd_params* p, size_t* bytes) {
  if (!p) return AGA_ERR_INVALID_ARGUMENT;
  int st = validate(&p->fwd);
  if (st != AGA_OK) return st;
  if (!bytes || !p->dout || !p->dq || !p->dk || !p->dv || !p->fwd.lse) return AGA_ERR_INVALID_ARGUMENT;
  const void* ptrs[] = {p->dout, p->dq, p->dk, p->dv};
  for (const void* x : ptrs)
    if (reinterpret_cast<uintptr_t>(x) & 15) return AGA_ERR_UNSUPPORTED;
  if ((p->fwd.impl == AGA_ATTN_TCGEN05 || p->fwd.kv_len || p->d_guided_part) && !use_tc_bwd(p->fwd)) return AGA_ERR_UNSUPPORTED;
  if (p->d_guided_part && !p->fwd.guided_pattern) return AGA_ERR_INVALID_ARGUMENT;
  *bytes = use_tc_bwd(p->fwd) ? attn_tc_bwd_workspace(p->fwd) : attn_simt_bwd_workspace(p->fwd);
  return AGA_OK;
}

extern "C" int aga_attn_bwd(const aga_attn_bwd_params* p, void* workspace, size_t workspace_bytes, void* stream) {
  size_t need = 0;
  int st = aga_attn_bwd_workspace_bytes(p, &need);
  if (st != AGA_OK) return st;
  if (!workspace || workspace_bytes < need) return AGA_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return use_tc_bwd(p->fwd) ? attn_tc_bwd(*p, workspace, s) : attn_simt_bwd(*p, workspace, s);
}
