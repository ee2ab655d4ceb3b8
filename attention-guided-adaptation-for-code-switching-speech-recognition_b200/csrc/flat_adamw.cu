// AdamW on the FLAT parameter / gradient buffers of the trainable (adapter) parameters, with the gradient clipping, the
// non-finite-norm skip and the bf16 shadow refresh folded in.
//
// Reference: espnet2/train/trainer.py:649-716 — clip_grad_norm_ over all trainable parameters, `if not
// torch.isfinite(grad_norm): skip the update`, optimizer.step() (torch.optim.AdamW from the recipe's `optim: adamw`).
// Stock PyTorch runs this as a chain over ~200 small tensors: multi-tensor L2 norm, clip coefficient, a multiply pass over
// the gradients, 8 fused-AdamW launches (chunked by tensor count: 39 us each, latency-bound on 147 K-element tensors)
// and 5 multi-tensor copies for the low-precision parameter copies autocast uses — 0.5 ms of a 24 ms step.  Here:
//   kernel 1  sum of squares of the flat gradient -> norm; the LAST block (ticket) also decides "skip", advances the step
//             counter or the skipped-updates counter: deterministic (fixed-order reduction of the block partials)
//   kernel 2  one pass: g' = g * min(1, max_norm / (norm + 1e-6)); torch.optim.AdamW's update (same operation order as
//             its fused kernel: decoupled decay, moments, double-precision hyper-parameters); bf16 copy of the new
//             parameters.  Nothing is written when the norm is not finite.
// HBM-bound: 7 fp32 + 1 bf16 streams over n elements (30 B per element).
#include "aga_common.cuh"

#include <algorithm>

namespace aga {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 1024;

struct NormWs {
  double partial[kMaxBlocks];
  unsigned int ticket;
};

__global__ void __launch_bounds__(kThreads) flat_norm_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ norm_out,
                                                            float* __restrict__ step, float* __restrict__ skipped, NormWs* ws) {
  float acc = 0.f;
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  // four independent 16-byte loads per trip (one load per trip ran at 1.6 TB/s on load latency: ncu, profiles/r2_new_kernels.md)
  const int64_t stride = int64_t(gridDim.x) * kThreads;
  int64_t i = int64_t(blockIdx.x) * kThreads + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    const float4 v0 = __ldg(g4 + i), v1 = __ldg(g4 + i + stride), v2 = __ldg(g4 + i + 2 * stride), v3 = __ldg(g4 + i + 3 * stride);
    acc = fmaf(v0.x, v0.x, acc); acc = fmaf(v0.y, v0.y, acc); acc = fmaf(v0.z, v0.z, acc); acc = fmaf(v0.w, v0.w, acc);
    acc = fmaf(v1.x, v1.x, acc); acc = fmaf(v1.y, v1.y, acc); acc = fmaf(v1.z, v1.z, acc); acc = fmaf(v1.w, v1.w, acc);
    acc = fmaf(v2.x, v2.x, acc); acc = fmaf(v2.y, v2.y, acc); acc = fmaf(v2.z, v2.z, acc); acc = fmaf(v2.w, v2.w, acc);
    acc = fmaf(v3.x, v3.x, acc); acc = fmaf(v3.y, v3.y, acc); acc = fmaf(v3.z, v3.z, acc); acc = fmaf(v3.w, v3.w, acc);
  }
  for (; i < n4; i += stride) {
    const float4 v = __ldg(g4 + i);
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  if (blockIdx.x == 0)
    for (int64_t j = (n4 << 2) + threadIdx.x; j < n; j += kThreads) acc = fmaf(g[j], g[j], acc);
  __shared__ double red[kThreads / 32];
  __shared__ bool last;
  const double d = double(warp_sum(acc));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    ws->partial[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {  // the whole last block: strided partial sums, then a fixed-shape tree — the same bits on every run
    __shared__ double tree[kThreads];
    __threadfence();
    double t = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += kThreads) t += *(volatile double*)&ws->partial[b];
    tree[threadIdx.x] = t;
    __syncthreads();
    for (int w = kThreads / 2; w > 0; w >>= 1) {
      if (int(threadIdx.x) < w) tree[threadIdx.x] += tree[threadIdx.x + w];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const float norm = float(sqrt(tree[0]));
      *norm_out = norm;
      if (isfinite(norm)) {
        if (step) *step += 1.0f;
      } else if (skipped) {
        *skipped += 1.0f;
      }
      ws->ticket = 0;  // ready for the next launch
    }
  }
}

// torch's fused AdamW (aten/src/ATen/native/cuda/fused_adam_utils.cuh, adam_math) keeps param / moments as float and the
// hyper-parameters as double, so each statement is evaluated in double and rounded to float once; mirrored here
// statement by statement so that a run with this optimizer tracks a stock run to the last bits.
__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, double lr, double wd, double beta1, double beta2,
                                          float step_size, double bc2_sqrt, double eps) {
  p = float(double(p) - lr * wd * double(p));
  m = float(beta1 * double(m) + (1.0 - beta1) * double(g));
  v = float(beta2 * double(v) + (1.0 - beta2) * double(g) * double(g));
  const float denom = float(double(sqrtf(v)) / bc2_sqrt + eps);
  p -= step_size * m / denom;
}

__global__ void __launch_bounds__(kThreads) flat_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                             float* __restrict__ v, int64_t n, const float* __restrict__ lr_ptr,
                                                             double beta1, double beta2, double eps, double wd,
                                                             const float* __restrict__ step_ptr, const float* __restrict__ norm_ptr,
                                                             float max_norm, __nv_bfloat16* __restrict__ shadow) {
  const float norm = norm_ptr ? *norm_ptr : 0.f;
  if (!isfinite(norm)) return;  // the reference skips the update (trainer.py:677); the step counter was not advanced either
  __shared__ double s_c[2];
  if (threadIdx.x == 0) {
    const double t = double(*step_ptr);
    s_c[0] = double(*lr_ptr) / (1.0 - pow(beta1, t));   // step_size = lr / bias_correction1
    s_c[1] = sqrt(1.0 - pow(beta2, t));                  // bias_correction2_sqrt
  }
  __syncthreads();
  const double lr = double(*lr_ptr), bc2_sqrt = s_c[1];
  const float step_size = float(s_c[0]);
  const float coef = (norm_ptr && max_norm > 0.f) ? fminf(1.0f, max_norm / (norm + 1e-6f)) : 1.0f;
  const int64_t n4 = n >> 2;
  for (int64_t i = int64_t(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += int64_t(gridDim.x) * kThreads) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    adamw_one(pp.x, gg.x * coef, mm.x, vv.x, lr, wd, beta1, beta2, step_size, bc2_sqrt, eps);
    adamw_one(pp.y, gg.y * coef, mm.y, vv.y, lr, wd, beta1, beta2, step_size, bc2_sqrt, eps);
    adamw_one(pp.z, gg.z * coef, mm.z, vv.z, lr, wd, beta1, beta2, step_size, bc2_sqrt, eps);
    adamw_one(pp.w, gg.w * coef, mm.w, vv.w, lr, wd, beta1, beta2, step_size, bc2_sqrt, eps);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow) {
      __nv_bfloat162 a = __floats2bfloat162_rn(pp.x, pp.y), b = __floats2bfloat162_rn(pp.z, pp.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&a);
      u.y = *reinterpret_cast<uint32_t*>(&b);
      reinterpret_cast<uint2*>(shadow)[i] = u;
    }
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += kThreads) {
      float pp = p[i], mm = m[i], vv = v[i];
      adamw_one(pp, g[i] * coef, mm, vv, lr, wd, beta1, beta2, step_size, bc2_sqrt, eps);
      p[i] = pp;
      m[i] = mm;
      v[i] = vv;
      if (shadow) shadow[i] = __float2bfloat16_rn(pp);
    }
}

int n_blocks(int64_t n) {
  static const int n_sm = []() {
    int dev = 0, k = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&k, cudaDevAttrMultiProcessorCount, dev);
    return k > 0 ? k : 148;
  }();
  const int64_t want = (n / 4 + kThreads - 1) / kThreads;
  return int(std::max<int64_t>(1, std::min<int64_t>(want, std::min(kMaxBlocks, n_sm * 6))));
}

}  // namespace
}  // namespace aga

using namespace aga;

extern "C" int aga_flat_grad_norm_workspace_bytes(size_t* bytes) {
  if (!bytes) return AGA_ERR_INVALID_ARGUMENT;
  *bytes = sizeof(NormWs);
  return AGA_OK;
}

extern "C" int aga_flat_grad_norm(const float* g, int64_t n, float* norm_out, float* step, float* skipped, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (!g || n <= 0 || !norm_out || !workspace) return AGA_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < sizeof(NormWs)) return AGA_ERR_WORKSPACE_TOO_SMALL;
  if ((reinterpret_cast<uintptr_t>(g) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 7)) return AGA_ERR_UNSUPPORTED;
  flat_norm_kernel<<<n_blocks(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(g, n, norm_out, step, skipped,
                                                                                      static_cast<NormWs*>(workspace));
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

extern "C" int aga_flat_adamw(float* p, const float* g, float* m, float* v, int64_t n, const float* lr, double beta1, double beta2,
                              double eps, double weight_decay, const float* step, const float* grad_norm, float max_norm,
                              void* shadow_bf16, void* stream) {
  if (!p || !g || !m || !v || n <= 0 || !lr || !step) return AGA_ERR_INVALID_ARGUMENT;
  if (!(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0) || eps < 0.0) return AGA_ERR_INVALID_ARGUMENT;
  const uintptr_t all = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                        reinterpret_cast<uintptr_t>(v);
  if ((all & 15) || (reinterpret_cast<uintptr_t>(shadow_bf16) & 7)) return AGA_ERR_UNSUPPORTED;
  flat_adamw_kernel<<<n_blocks(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_norm, max_norm, static_cast<__nv_bfloat16*>(shadow_bf16));
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}
