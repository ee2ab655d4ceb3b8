// Shared helpers for the aga_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>

#include "aga_b200.h"

namespace aga {

extern thread_local int g_last_cuda_error;
extern std::atomic<uint64_t> g_launch_count;

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = static_cast<int>(e);
  return AGA_ERR_CUDA;
}

#define AGA_CUDA_TRY(expr)                                 \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) return ::aga::cuda_fail(_e);    \
  } while (0)

// Call after every kernel launch: counts the launch and surfaces launch-configuration errors.
#define AGA_AFTER_LAUNCH()                                          \
  do {                                                              \
    ::aga::g_launch_count.fetch_add(1, std::memory_order_relaxed);  \
    cudaError_t _e = cudaGetLastError();                            \
    if (_e != cudaSuccess) return ::aga::cuda_fail(_e);             \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// cuTensorMapEncodeTiled is a driver-API call: it fails in a thread that has no current context yet (PyTorch's autograd
// threads have only had cudaSetDevice called, which binds lazily).  One runtime call per thread makes the primary
// context current before the first tensor map is encoded there.
inline void ensure_context_in_this_thread() {
  thread_local bool done = false;
  if (!done) {
    cudaFree(nullptr);
    done = true;
  }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Monotonic float <-> uint32 key (total order matches float order; key 0 is below every float).
__device__ __forceinline__ uint32_t float_to_key(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}

template <typename T> __device__ __forceinline__ float to_f32(T x);
template <> __device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

}  // namespace aga
