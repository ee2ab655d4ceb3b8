// Backward of the Adapter's GELU fused with the bias gradient of its first Linear.
//
// Reference: whisper/whisper/model.py:181-194  Adapter = x + Linear(GELU(Linear(x))).  In the backward pass
//   dh = dg * gelu'(h)   and   db1 = sum_rows dh
// are an elementwise pass and a column reduction over the same (rows, bottleneck) matrix; eager PyTorch runs them as
// gelu_backward + a strided reduce_kernel (30 us for 9 MB).  One pass here: each thread owns 16 bytes of a row, walks
// rows with a grid stride keeping its columns' partial sums in registers, one smem reduction + one atomicAdd per
// column per CTA at the end.  gelu' is the exact (erf) form in fp32, as at::gelu_backward computes it.
#include "aga_common.cuh"
#include "gelu_math.cuh"

#include <algorithm>

namespace aga {
namespace {

constexpr int kThreads = 256;

template <typename T> struct Pack;
template <> struct Pack<float> {
  static constexpr int kVec = 4;
  using Raw = float4;
  static __device__ __forceinline__ Raw load_raw(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& t, float (&v)[4]) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Pack<__nv_bfloat16> {
  static constexpr int kVec = 8;
  using Raw = uint4;
  static __device__ __forceinline__ Raw load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] = __uint_as_float(w[e] << 16);
      v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
      v[2 * e] = f.x;
      v[2 * e + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      w[e] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
gelu_bwd_colsum_kernel(const T* __restrict__ dg, const T* __restrict__ h, int64_t rows, int cols, int tpr, int rpb,
                       T* __restrict__ dh, float* __restrict__ colsum) {
  constexpr int V = Pack<T>::kVec;
  extern __shared__ float red[];  // rpb x cols
  const int tid = threadIdx.x;
  const bool active = tid < tpr * rpb;
  const int cg = tid % tpr, ro = tid / tpr;
  float acc[V];
#pragma unroll
  for (int e = 0; e < V; ++e) acc[e] = 0.f;
  if (active) {
    // four rows per trip, their eight 16-byte loads issued before any arithmetic: with one row per trip the kernel ran
    // at 2.6x its HBM time on load latency (240 threads x 32 bytes in flight per CTA)
    constexpr int kU = 4;
    const int64_t step = int64_t(gridDim.x) * rpb;
    for (int64_t r0 = int64_t(blockIdx.x) * rpb + ro; r0 < rows; r0 += kU * step) {
      typename Pack<T>::Raw xr[kU], gr[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t r = r0 + u * step;
        if (r < rows) {
          xr[u] = Pack<T>::load_raw(h + r * cols + cg * V);
          gr[u] = Pack<T>::load_raw(dg + r * cols + cg * V);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t r = r0 + u * step;
        if (r >= rows) break;
        float x[V], g[V], o[V];
        Pack<T>::unpack(xr[u], x);
        Pack<T>::unpack(gr[u], g);
        // gelu'(x) = Phi(x) + x phi(x) on PAIRS (packed fp32x2 FMAs, erf by Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7:
        // gelu_math.cuh)
#pragma unroll
        for (int e = 0; e < V; e += 2) {
          const float2 d = dgelu_erf2(make_float2(x[e], x[e + 1]));
          o[e] = g[e] * d.x;
          o[e + 1] = g[e + 1] * d.y;
          acc[e] += to_f32<T>(from_f32<T>(o[e]));  // the bias gradient sums the ROUNDED dh, as dh.sum(0) would
          acc[e + 1] += to_f32<T>(from_f32<T>(o[e + 1]));
        }
        Pack<T>::store(dh + r * cols + cg * V, o);
      }
    }
#pragma unroll
    for (int e = 0; e < V; ++e) red[ro * cols + cg * V + e] = acc[e];
  }
  __syncthreads();
  for (int c = tid; c < cols; c += kThreads) {
    float t = 0.f;
    for (int r = 0; r < rpb; ++r) t += red[r * cols + c];
    atomicAdd(colsum + c, t);
  }
}

template <typename T>
int launch(const void* dg, const void* h, int64_t rows, int cols, void* dh, float* colsum, bool accumulate, cudaStream_t s) {
  constexpr int V = Pack<T>::kVec;
  const int tpr = cols / V;
  if (cols % V != 0 || tpr > kThreads) return AGA_ERR_UNSUPPORTED;
  const int rpb = kThreads / tpr;
  const int64_t want = (rows + rpb - 1) / rpb;
  const unsigned grid = unsigned(std::max<int64_t>(1, std::min<int64_t>(want, 148 * 4)));
  if (!accumulate) AGA_CUDA_TRY(cudaMemsetAsync(colsum, 0, size_t(cols) * sizeof(float), s));
  gelu_bwd_colsum_kernel<T><<<grid, kThreads, size_t(rpb) * cols * sizeof(float), s>>>(
      static_cast<const T*>(dg), static_cast<const T*>(h), rows, cols, tpr, rpb, static_cast<T*>(dh), colsum);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

}  // namespace
}  // namespace aga

namespace {
int gelu_bwd_colsum_impl(const void* dg, const void* h, int dtype, int64_t rows, int cols, void* dh, float* colsum, bool accumulate,
                         void* stream) {
  using namespace aga;
  if (!dg || !h || !dh || !colsum || rows <= 0 || cols <= 0) return AGA_ERR_INVALID_ARGUMENT;
  if (dtype != AGA_F32 && dtype != AGA_BF16) return AGA_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(dg) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(dh)) & 15)
    return AGA_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return dtype == AGA_BF16 ? launch<__nv_bfloat16>(dg, h, rows, cols, dh, colsum, accumulate, s)
                           : launch<float>(dg, h, rows, cols, dh, colsum, accumulate, s);
}
}  // namespace

extern "C" int aga_gelu_bwd_colsum(const void* dg, const void* h, int dtype, int64_t rows, int cols, void* dh,
                                   float* colsum, void* stream) {
  return gelu_bwd_colsum_impl(dg, h, dtype, rows, cols, dh, colsum, false, stream);
}
extern "C" int aga_gelu_bwd_colsum_acc(const void* dg, const void* h, int dtype, int64_t rows, int cols, void* dh,
                                       float* colsum, void* stream) {
  return gelu_bwd_colsum_impl(dg, h, dtype, rows, cols, dh, colsum, true, stream);
}
