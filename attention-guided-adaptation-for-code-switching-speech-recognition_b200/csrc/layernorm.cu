// LayerNorm with fp32 statistics on bf16 / fp32 rows (HBM-bound; SURVEY.md §8f #2, the first "next" row).
//
// Reference: whisper/whisper/model.py:30-32  `super().forward(x.float()).type(x.dtype)` — every block runs it up to
// four times (attn_ln, adapter_attn_ln, mlp_ln, adapter_mlp_ln; the adapter LNs are the trainable ones), and eager
// PyTorch turns each call into an fp32 up-cast, the normalisation and a down-cast (5x the algorithmic traffic),
// with a separate slow column reduction for the gamma/beta gradients.
//   forward : one warp per row, the row lives in registers (D/32 values per lane, 8- or 16-byte loads),
//             two-pass mean / variance in fp32, writes y in the input dtype and (mean, rstd) for backward.
//   backward: persistent warps stride over rows; dx = rstd (g - mean(g) - xhat mean(g xhat)), g = dy*gamma; each lane
//             keeps its columns' dgamma / dbeta partial sums in registers across all its rows, one smem reduction
//             and one atomicAdd per column per CTA at the end.
// Algorithmic bytes per row: forward 2*D*sizeof(T); backward 3*D*sizeof(T).
#include "aga_common.cuh"

#include <algorithm>

namespace aga {
namespace {

constexpr int kLnWarps = 8;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

// lane l owns columns {128*c + 4*l .. +3 : c < NC}
template <typename T> __device__ __forceinline__ float round_to(float v);
template <> __device__ __forceinline__ float round_to<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_to<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// residual != nullptr: normalises s = T(x + residual) (the sum is rounded to the row dtype first, as the reference's
// `x + self.model(x)` is) and, when sum_out != nullptr, also writes s — the tensor the backward needs.
template <typename T, int NC>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ residual, int64_t rows, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, T* __restrict__ y, T* __restrict__ sum_out,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  constexpr int D = NC * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * D;
  float v[NC][4];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    Vec4<T>::load(xr + c * 128 + lane * 4, v[c]);
    if (residual) {
      float rv[4];
      Vec4<T>::load(residual + row * D + c * 128 + lane * 4, rv);
#pragma unroll
      for (int e = 0; e < 4; ++e) v[c][e] = round_to<T>(v[c][e] + rv[e]);
      if (sum_out) Vec4<T>::store(sum_out + row * D + c * 128 + lane * 4, v[c]);
    }
    sum += (v[c][0] + v[c][1]) + (v[c][2] + v[c][3]);
  }
  const float mean = warp_sum(sum) * (1.0f / D);
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float d = v[c][e] - mean;
      sq = fmaf(d, d, sq);
    }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
  T* yr = y + row * D;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c * 128 + lane * 4));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c * 128 + lane * 4));
    float o[4];
    o[0] = fmaf((v[c][0] - mean) * rstd, g.x, b.x);
    o[1] = fmaf((v[c][1] - mean) * rstd, g.y, b.y);
    o[2] = fmaf((v[c][2] - mean) * rstd, g.z, b.z);
    o[3] = fmaf((v[c][3] - mean) * rstd, g.w, b.w);
    Vec4<T>::store(yr + c * 128 + lane * 4, o);
  }
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

// kParamGrads == 2 additionally accumulates dxsum[c] = sum_rows dx[row, c]: when the normalised tensor is
// x + Linear(...)(x), that column sum IS the gradient of the Linear's bias (saves a separate reduction pass).
template <typename T, int NC, int kParamGrads>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, int64_t rows, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum) {
  constexpr int D = NC * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[NC][4];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(gamma + c * 128 + lane * 4));
    g[c][0] = t.x; g[c][1] = t.y; g[c][2] = t.z; g[c][3] = t.w;
  }
  float ag[NC][4], ab[NC][4], ax[kParamGrads == 2 ? NC : 1][4];
  if (kParamGrads) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) ag[c][e] = ab[c][e] = 0.f;
  }
  if (kParamGrads == 2) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) ax[c][e] = 0.f;
  }
  const int64_t stride = int64_t(gridDim.x) * kLnWarps;
  for (int64_t row = int64_t(blockIdx.x) * kLnWarps + warp; row < rows; row += stride) {
    const float mu = mean[row], rs = rstd[row];
    float xh[NC][4], gy[NC][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float xv[4], dv[4];
      Vec4<T>::load(x + row * D + c * 128 + lane * 4, xv);
      Vec4<T>::load(dy + row * D + c * 128 + lane * 4, dv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        xh[c][e] = (xv[e] - mu) * rs;
        gy[c][e] = dv[e] * g[c][e];
        s1 += gy[c][e];
        s2 = fmaf(gy[c][e], xh[c][e], s2);
        if (kParamGrads) {
          ag[c][e] = fmaf(dv[e], xh[c][e], ag[c][e]);
          ab[c][e] += dv[e];
        }
      }
    }
    const float c1 = warp_sum(s1) * (1.0f / D), c2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[e] = rs * (gy[c][e] - c1 - xh[c][e] * c2);
        if (kParamGrads == 2) ax[c][e] += o[e];
      }
      Vec4<T>::store(dx + row * D + c * 128 + lane * 4, o);
    }
  }
  if (kParamGrads) {
    __shared__ float red[kLnWarps][128];
#pragma unroll 1
    for (int pass = 0; pass < (kParamGrads == 2 ? 3 : 2); ++pass) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; ++e)
          red[warp][lane * 4 + e] = pass == 0 ? ag[c][e] : (pass == 1 ? ab[c][e] : ax[kParamGrads == 2 ? c : 0][e]);
        __syncthreads();
        if (threadIdx.x < 128) {
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < kLnWarps; ++w) t += red[w][threadIdx.x];
          atomicAdd((pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum)) + c * 128 + threadIdx.x, t);
        }
      }
    }
  }
}

template <typename T, int NC>
int launch_fwd(const void* x, const void* residual, int64_t rows, const float* gamma, const float* beta, float eps, void* y,
               void* sum_out, float* mean, float* rstd, cudaStream_t s) {
  const unsigned grid = unsigned((rows + kLnWarps - 1) / kLnWarps);
  layernorm_fwd_kernel<T, NC><<<grid, kLnWarps * 32, 0, s>>>(static_cast<const T*>(x), static_cast<const T*>(residual), rows,
                                                              gamma, beta, eps, static_cast<T*>(y), static_cast<T*>(sum_out),
                                                              mean, rstd);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

template <typename T, int NC>
int launch_bwd(const void* dy, const void* x, int64_t rows, const float* gamma, const float* mean, const float* rstd,
               void* dx, float* dgamma, float* dbeta, float* dxsum, cudaStream_t s) {
  const int64_t want = (rows + kLnWarps - 1) / kLnWarps;
  const unsigned grid = unsigned(std::max<int64_t>(1, std::min<int64_t>(want, 148 * 4)));
  if (dgamma && dbeta) {
    AGA_CUDA_TRY(cudaMemsetAsync(dgamma, 0, NC * 128 * sizeof(float), s));
    AGA_CUDA_TRY(cudaMemsetAsync(dbeta, 0, NC * 128 * sizeof(float), s));
    if (dxsum) {
      AGA_CUDA_TRY(cudaMemsetAsync(dxsum, 0, NC * 128 * sizeof(float), s));
      layernorm_bwd_kernel<T, NC, 2><<<grid, kLnWarps * 32, 0, s>>>(static_cast<const T*>(dy), static_cast<const T*>(x), rows,
                                                                     gamma, mean, rstd, static_cast<T*>(dx), dgamma, dbeta, dxsum);
    } else {
      layernorm_bwd_kernel<T, NC, 1><<<grid, kLnWarps * 32, 0, s>>>(static_cast<const T*>(dy), static_cast<const T*>(x), rows,
                                                                     gamma, mean, rstd, static_cast<T*>(dx), dgamma, dbeta, nullptr);
    }
  } else {
    layernorm_bwd_kernel<T, NC, 0><<<grid, kLnWarps * 32, 0, s>>>(static_cast<const T*>(dy), static_cast<const T*>(x), rows, gamma,
                                                                   mean, rstd, static_cast<T*>(dx), nullptr, nullptr, nullptr);
  }
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

#define AGA_LN_DISPATCH(FN, ...)                                                            \
  switch (D / 128) {                                                                        \
    case 3:  return bf16 ? FN<__nv_bfloat16, 3>(__VA_ARGS__)  : FN<float, 3>(__VA_ARGS__);  \
    case 4:  return bf16 ? FN<__nv_bfloat16, 4>(__VA_ARGS__)  : FN<float, 4>(__VA_ARGS__);  \
    case 6:  return bf16 ? FN<__nv_bfloat16, 6>(__VA_ARGS__)  : FN<float, 6>(__VA_ARGS__);  \
    case 8:  return bf16 ? FN<__nv_bfloat16, 8>(__VA_ARGS__)  : FN<float, 8>(__VA_ARGS__);  \
    case 10: return bf16 ? FN<__nv_bfloat16, 10>(__VA_ARGS__) : FN<float, 10>(__VA_ARGS__); \
    default: return AGA_ERR_UNSUPPORTED;                                                    \
  }

int check(const void* a, const void* b, int dtype, int64_t rows, int D) {
  if (!a || !b || rows <= 0 || D <= 0) return AGA_ERR_INVALID_ARGUMENT;
  if (dtype != AGA_F32 && dtype != AGA_BF16) return AGA_ERR_INVALID_ARGUMENT;
  if (D % 128 != 0) return AGA_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) return AGA_ERR_UNSUPPORTED;
  return AGA_OK;
}

}  // namespace
}  // namespace aga

using namespace aga;

extern "C" int aga_layernorm_fwd(const void* x, const void* residual, int dtype, int64_t rows, int D, const float* gamma,
                                 const float* beta, float eps, void* y, void* sum_out, float* mean, float* rstd,
                                 void* stream) {
  int st = check(x, y, dtype, rows, D);
  if (st != AGA_OK) return st;
  if (!gamma || !beta || !mean || !rstd || (sum_out && !residual)) return AGA_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(sum_out)) & 15) return AGA_ERR_UNSUPPORTED;
  const bool bf16 = dtype == AGA_BF16;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  AGA_LN_DISPATCH(launch_fwd, x, residual, rows, gamma, beta, eps, y, sum_out, mean, rstd, s)
}

extern "C" int aga_layernorm_bwd(const void* dy, const void* x, int dtype, int64_t rows, int D, const float* gamma,
                                 const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                                 float* dxsum, void* stream) {
  int st = check(x, dx, dtype, rows, D);
  if (st != AGA_OK) return st;
  if (!dy || !gamma || !mean || !rstd || ((dgamma == nullptr) != (dbeta == nullptr)) || (dxsum && !dgamma))
    return AGA_ERR_INVALID_ARGUMENT;
  if (reinterpret_cast<uintptr_t>(dy) & 15) return AGA_ERR_UNSUPPORTED;
  const bool bf16 = dtype == AGA_BF16;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  AGA_LN_DISPATCH(launch_bwd, dy, x, rows, gamma, mean, rstd, dx, dgamma, dbeta, dxsum, s)
}
