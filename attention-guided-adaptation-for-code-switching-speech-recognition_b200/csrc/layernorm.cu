// LayerNorm with fp32 statistics on bf16 / fp32 rows (HBM-bound; SURVEY.md §8f #2, the first "next" row).
//
// Reference: whisper/whisper/model.py:30-32  `super().forward(x.float()).type(x.dtype)` — every block runs it up to
// four times (attn_ln, adapter_attn_ln, mlp_ln, adapter_mlp_ln; the adapter LNs are the trainable ones), and eager
// PyTorch turns each call into an fp32 up-cast, the normalisation and a down-cast (5x the algorithmic traffic),
// with a separate slow column reduction for the gamma/beta gradients.
//   forward : one warp per row, the row lives in registers (D/32 values per lane, 16-byte loads, all issued up front),
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

// V consecutive elements of a row <-> fp32 registers; 16-byte accesses wherever the row length allows (bf16: V = 8)
template <typename T, int V> struct Vec;
template <> struct Vec<float, 4> {
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
template <> struct Vec<__nv_bfloat16, 4> {
  using Raw = uint2;
  static __device__ __forceinline__ Raw load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint2*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& t, float (&v)[4]) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
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
template <> struct Vec<__nv_bfloat16, 8> {
  using Raw = uint4;
  static __device__ __forceinline__ Raw load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
      v[2 * e] = f.x;
      v[2 * e + 1] = f.y;
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
template <int V> __device__ __forceinline__ void load_f32(const float* p, float (&v)[V]) {
#pragma unroll
  for (int q = 0; q < V / 4; ++q) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}

template <int V> __device__ __forceinline__ void load_smem_f32(const float* p, float (&v)[V]) {
#pragma unroll
  for (int q = 0; q < V / 4; ++q) {
    const float4 t = *(reinterpret_cast<const float4*>(p) + q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}

// lane l owns columns {32*V*c + V*l .. + V-1 : c < NC}
template <typename T> __device__ __forceinline__ float round_to(float v);
template <> __device__ __forceinline__ float round_to<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_to<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// residual != nullptr: normalises s = T(x + residual) (the sum is rounded to the row dtype first, as the reference's
// `x + self.model(x)` is) and, when sum_out != nullptr, also writes s — the tensor the backward needs.
template <typename T, int V, int NC, bool kPair>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ residual, int64_t rows, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, T* __restrict__ y, T* __restrict__ sum_out,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, const float* __restrict__ gamma2,
                     const float* __restrict__ beta2, float eps2, T* __restrict__ y2, float* __restrict__ mean2_out,
                     float* __restrict__ rstd2_out) {
  constexpr int D = NC * 32 * V;
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * D;
  float v[NC][V];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) Vec<T, V>::load(xr + c * 32 * V + lane * V, v[c]);  // all loads in flight first
  if (residual) {
    float rv[NC][V];
#pragma unroll
    for (int c = 0; c < NC; ++c) Vec<T, V>::load(residual + row * D + c * 32 * V + lane * V, rv[c]);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
      for (int e = 0; e < V; ++e) v[c][e] = round_to<T>(v[c][e] + rv[c][e]);
      if (sum_out) Vec<T, V>::store(sum_out + row * D + c * 32 * V + lane * V, v[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int e = 0; e < V; ++e) sum += v[c][e];
  const float mean = warp_sum(sum) * (1.0f / D);
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float d = v[c][e] - mean;
      sq = fmaf(d, d, sq);
    }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
  T* yr = y + row * D;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float g[V], b[V], o[V];
    load_f32<V>(gamma + c * 32 * V + lane * V, g);
    load_f32<V>(beta + c * 32 * V + lane * V, b);
#pragma unroll
    for (int e = 0; e < V; ++e) o[e] = fmaf((v[c][e] - mean) * rstd, g[e], b[e]);
    Vec<T, V>::store(yr + c * 32 * V + lane * V, o);
    if (kPair) {  // keep the row as it was STORED (rounded to T): that is what a LayerNorm launched next would read
#pragma unroll
      for (int e = 0; e < V; ++e) v[c][e] = round_to<T>(o[e]);
    }
  }
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  if (!kPair) return;
  // ---- the pre-LayerNorm of the NEXT residual branch on the row just produced (whisper/model.py:231-246: the adapter's
  //      post-LN output is at once the residual stream and the input of the next `*_ln`): no second pass over HBM
  float sum2 = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int e = 0; e < V; ++e) sum2 += v[c][e];
  const float m2 = warp_sum(sum2) * (1.0f / D);
  float sq2 = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float d = v[c][e] - m2;
      sq2 = fmaf(d, d, sq2);
    }
  const float r2 = rsqrtf(warp_sum(sq2) * (1.0f / D) + eps2);
  T* y2r = y2 + row * D;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float g[V], b[V], o[V];
    load_f32<V>(gamma2 + c * 32 * V + lane * V, g);
    load_f32<V>(beta2 + c * 32 * V + lane * V, b);
#pragma unroll
    for (int e = 0; e < V; ++e) o[e] = fmaf((v[c][e] - m2) * r2, g[e], b[e]);
    Vec<T, V>::store(y2r + c * 32 * V + lane * V, o);
  }
  if (lane == 0) {
    mean2_out[row] = m2;
    rstd2_out[row] = r2;
  }
}

// kParamGrads == 2 additionally accumulates dxsum[c] = sum_rows dx[row, c]: when the normalised tensor is
// x + Linear(...)(x), that column sum IS the gradient of the Linear's bias (saves a separate reduction pass).
template <typename T, int V, int NC, int kParamGrads>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, int64_t rows, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const T* __restrict__ dres,
                     T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum) {
  constexpr int D = NC * 32 * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[NC][V];
#pragma unroll
  for (int c = 0; c < NC; ++c) load_f32<V>(gamma + c * 32 * V + lane * V, g[c]);
  float ag[NC][V], ab[NC][V], ax[kParamGrads == 2 ? NC : 1][V];
  if (kParamGrads) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int e = 0; e < V; ++e) ag[c][e] = ab[c][e] = 0.f;
  }
  if (kParamGrads == 2) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int e = 0; e < V; ++e) ax[c][e] = 0.f;
  }
  const int64_t stride = int64_t(gridDim.x) * kLnWarps;
  for (int64_t row = int64_t(blockIdx.x) * kLnWarps + warp; row < rows; row += stride) {
    const float mu = mean[row], rs = rstd[row];
    float xh[NC][V], gy[NC][V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      Vec<T, V>::load(x + row * D + c * 32 * V + lane * V, xh[c]);
      Vec<T, V>::load(dy + row * D + c * 32 * V + lane * V, gy[c]);
    }
    // the residual-path gradient is fetched together with x and dy (kept packed): behind the two warp reductions its
    // latency would be exposed a second time per row
    typename Vec<T, V>::Raw rr[NC];
    if (dres) {
#pragma unroll
      for (int c = 0; c < NC; ++c) rr[c] = Vec<T, V>::load_raw(dres + row * D + c * 32 * V + lane * V);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float dv = gy[c][e];
        xh[c][e] = (xh[c][e] - mu) * rs;
        gy[c][e] = dv * g[c][e];
        s1 += gy[c][e];
        s2 = fmaf(gy[c][e], xh[c][e], s2);
        if (kParamGrads) {
          ag[c][e] = fmaf(dv, xh[c][e], ag[c][e]);
          ab[c][e] += dv;
        }
      }
    }
    const float c1 = warp_sum(s1) * (1.0f / D), c2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float o[V];
#pragma unroll
      for (int e = 0; e < V; ++e) {
        o[e] = rs * (gy[c][e] - c1 - xh[c][e] * c2);
        if (kParamGrads == 2) ax[c][e] += o[e];
      }
      if (dres) {  // the gradient that reached the same tensor through the residual connection: one pass instead of an add kernel
        float r[V];
        Vec<T, V>::unpack(rr[c], r);
#pragma unroll
        for (int e = 0; e < V; ++e) o[e] = round_to<T>(o[e]) + r[e];
      }
      Vec<T, V>::store(dx + row * D + c * 32 * V + lane * V, o);
    }
  }
  if (kParamGrads) {
    __shared__ float red[kLnWarps][32 * V];
#pragma unroll 1
    for (int pass = 0; pass < (kParamGrads == 2 ? 3 : 2); ++pass) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        __syncthreads();
#pragma unroll
        for (int e = 0; e < V; ++e)
          red[warp][lane * V + e] = pass == 0 ? ag[c][e] : (pass == 1 ? ab[c][e] : ax[kParamGrads == 2 ? c : 0][e]);
        __syncthreads();
        for (int col = threadIdx.x; col < 32 * V; col += kLnWarps * 32) {
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < kLnWarps; ++w) t += red[w][col];
          atomicAdd((pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum)) + c * 32 * V + col, t);
        }
      }
    }
  }
}


// The parameter-gradient variants (trainable adapter LayerNorms) carry 2-3 fp32 accumulators per owned column
// (72 registers at D = 768), which leaves ONE 8-warp CTA per SM: with the loads of a row issued from registers that is
// 24 KiB in flight per SM and 2.9 TB/s (ncu, round 2).  Here every lane streams its 16-byte pieces of the next kStages-1
// rows into a private shared-memory ring with cp.async (no barriers: a lane only reads back what it copied itself), so
// ~100 KiB per SM are in flight whatever the register budget.  The row is read from the ring twice (sums, then dx) instead of
// being held unpacked in registers, which fits 12 warps beside the accumulators without spills.  One persistent CTA per SM.
template <int kBytes>
__device__ __forceinline__ void cp_async_piece(uint32_t smem_dst, const void* gsrc) {
  if constexpr (kBytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_dst), "l"(gsrc), "n"(kBytes) : "memory");
}
template <typename T, int V, int NC, int kParamGrads, int kStages, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1)
layernorm_bwd_ring_kernel(const T* __restrict__ dy, const T* __restrict__ x, int64_t rows, const float* __restrict__ gamma,
                          const float* __restrict__ mean, const float* __restrict__ rstd, const T* __restrict__ dres,
                          T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum) {
  static_assert(kParamGrads >= 1, "the ring kernel is the parameter-gradient path");
  constexpr int D = NC * 32 * V;
  constexpr int kRowBytes = D * int(sizeof(T));
  constexpr int kPiece = V * int(sizeof(T));
  extern __shared__ __align__(16) uint8_t ln_ring[];
  __shared__ __align__(16) float sg[D];             // gamma
  __shared__ float red[kWarps][32 * V];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_t = dres ? 3 : 2;  // row buffers per stage: x, dy (, dres)
  uint8_t* my = ln_ring + size_t(warp) * kStages * n_t * kRowBytes;
  const uint32_t my_addr = static_cast<uint32_t>(__cvta_generic_to_shared(my));
  for (int i = threadIdx.x; i < D; i += kWarps * 32) sg[i] = gamma[i];
  float ag[NC][V], ab[NC][V], ax[kParamGrads == 2 ? NC : 1][V];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int e = 0; e < V; ++e) {
      ag[c][e] = ab[c][e] = 0.f;
      if (kParamGrads == 2) ax[c][e] = 0.f;
    }
  const int64_t stride = int64_t(gridDim.x) * kWarps;
  const int64_t row0 = int64_t(blockIdx.x) * kWarps + warp;
  auto issue = [&](int64_t row, int stage) {
    if (row < rows) {
      const uint32_t base = my_addr + uint32_t(stage * n_t * kRowBytes);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int off = (c * 32 * V + lane * V);
        cp_async_piece<kPiece>(base + off * int(sizeof(T)), x + row * D + off);
        cp_async_piece<kPiece>(base + kRowBytes + off * int(sizeof(T)), dy + row * D + off);
        if (dres) cp_async_piece<kPiece>(base + 2 * kRowBytes + off * int(sizeof(T)), dres + row * D + off);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per row slot, empty past the end: uniform counting
  };
#pragma unroll
  for (int p = 0; p < kStages - 1; ++p) issue(row0 + p * stride, p);
  __syncthreads();  // gamma is in shared memory
  int it = 0;
  for (int64_t row = row0; row < rows; row += stride, ++it) {
    issue(row + (kStages - 1) * stride, (it + kStages - 1) % kStages);
    asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");
    const float mu = mean[row], rs = rstd[row];
    const T* sx = reinterpret_cast<const T*>(my + size_t(it % kStages) * n_t * kRowBytes);
    const T* sdy = sx + D;
    // pass 1 over the staged row: the two row sums and the parameter-gradient accumulators; nothing of the row is kept
    // in registers (the accumulators need them), pass 2 reads the row from shared memory again
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float xv[V], dv[V], g[V];
      Vec<T, V>::load(sx + c * 32 * V + lane * V, xv);
      Vec<T, V>::load(sdy + c * 32 * V + lane * V, dv);
      load_smem_f32<V>(sg + c * 32 * V + lane * V, g);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float xh = (xv[e] - mu) * rs;
        const float gy = dv[e] * g[e];
        s1 += gy;
        s2 = fmaf(gy, xh, s2);
        ag[c][e] = fmaf(dv[e], xh, ag[c][e]);
        ab[c][e] += dv[e];
      }
    }
    const float c1 = warp_sum(s1) * (1.0f / D), c2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float xv[V], dv[V], g[V], o[V];
      Vec<T, V>::load(sx + c * 32 * V + lane * V, xv);
      Vec<T, V>::load(sdy + c * 32 * V + lane * V, dv);
      load_smem_f32<V>(sg + c * 32 * V + lane * V, g);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float xh = (xv[e] - mu) * rs;
        o[e] = rs * (dv[e] * g[e] - c1 - xh * c2);
        if (kParamGrads == 2) ax[c][e] += o[e];
      }
      if (dres) {
        float r[V];
        Vec<T, V>::load(sdy + D + c * 32 * V + lane * V, r);
#pragma unroll
        for (int e = 0; e < V; ++e) o[e] = round_to<T>(o[e]) + r[e];
      }
      Vec<T, V>::store(dx + row * D + c * 32 * V + lane * V, o);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll 1
  for (int pass = 0; pass < (kParamGrads == 2 ? 3 : 2); ++pass) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      __syncthreads();
#pragma unroll
      for (int e = 0; e < V; ++e)
        red[warp][lane * V + e] = pass == 0 ? ag[c][e] : (pass == 1 ? ab[c][e] : ax[kParamGrads == 2 ? c : 0][e]);
      __syncthreads();
      for (int col = threadIdx.x; col < 32 * V; col += kWarps * 32) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += red[w][col];
        atomicAdd((pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum)) + c * 32 * V + col, t);
      }
    }
  }
}

template <typename T, int V, int NC>
int launch_fwd(const void* x, const void* residual, int64_t rows, const float* gamma, const float* beta, float eps, void* y,
               void* sum_out, float* mean, float* rstd, const float* gamma2, const float* beta2, float eps2, void* y2,
               float* mean2, float* rstd2, cudaStream_t s) {
  const unsigned grid = unsigned((rows + kLnWarps - 1) / kLnWarps);
  if (y2)
    layernorm_fwd_kernel<T, V, NC, true><<<grid, kLnWarps * 32, 0, s>>>(static_cast<const T*>(x), static_cast<const T*>(residual), rows,
                                                                         gamma, beta, eps, static_cast<T*>(y), static_cast<T*>(sum_out),
                                                                         mean, rstd, gamma2, beta2, eps2, static_cast<T*>(y2), mean2, rstd2);
  else
    layernorm_fwd_kernel<T, V, NC, false><<<grid, kLnWarps * 32, 0, s>>>(static_cast<const T*>(x), static_cast<const T*>(residual), rows,
                                                                          gamma, beta, eps, static_cast<T*>(y), static_cast<T*>(sum_out),
                                                                          mean, rstd, nullptr, nullptr, 0.f, nullptr, nullptr, nullptr);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

template <typename T, int V, int NC>
int launch_bwd(const void* dy, const void* x, int64_t rows, const float* gamma, const float* mean, const float* rstd,
               const void* dres, void* dx, float* dgamma, float* dbeta, float* dxsum, bool accumulate, cudaStream_t s) {
  constexpr int D = NC * 32 * V;
  const int64_t want = (rows + kLnWarps - 1) / kLnWarps;
  const unsigned grid = unsigned(std::max<int64_t>(1, std::min<int64_t>(want, 148 * 4)));
  if (dgamma && dbeta) {
    // one memset node when the caller packed the parameter-gradient rows back to back (ops.py does)
    const bool packed = dbeta == dgamma + D && (!dxsum || dxsum == dbeta + D);
    if (accumulate) {
      // the caller cleared the rows (one clear for a whole step's worth of them): nothing to do here
    } else if (packed) {
      AGA_CUDA_TRY(cudaMemsetAsync(dgamma, 0, (dxsum ? 3 : 2) * D * sizeof(float), s));
    } else {
      AGA_CUDA_TRY(cudaMemsetAsync(dgamma, 0, D * sizeof(float), s));
      AGA_CUDA_TRY(cudaMemsetAsync(dbeta, 0, D * sizeof(float), s));
      if (dxsum) AGA_CUDA_TRY(cudaMemsetAsync(dxsum, 0, D * sizeof(float), s));
    }
    // ring kernel: 12 warps x 4 stages of (x, dy[, dres]) rows; rows too long for that keep the register-fed kernel
    constexpr int kRingWarps = 12, kRing = 4;
    const size_t ring_bytes = size_t(kRingWarps) * kRing * (dres ? 3 : 2) * D * sizeof(T);
    constexpr size_t kRingStatic = (size_t(D) + kRingWarps * 32 * V) * sizeof(float);
    if (ring_bytes + kRingStatic <= 224 * 1024 && rows >= 4096) {
      static const int n_sm = []() {
        int dev = 0, n = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n > 0 ? n : 148;
      }();
      const unsigned rgrid = unsigned(std::max<int64_t>(1, std::min<int64_t>((rows + kRingWarps - 1) / kRingWarps, n_sm)));
      if (dxsum) {
        auto kern = layernorm_bwd_ring_kernel<T, V, NC, 2, kRing, kRingWarps>;
        AGA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(224 * 1024 - kRingStatic)));
        kern<<<rgrid, kRingWarps * 32, ring_bytes, s>>>(static_cast<const T*>(dy), static_cast<const T*>(x), rows, gamma, mean, rstd,
                                                       static_cast<const T*>(dres), static_cast<T*>(dx), dgamma, dbeta, dxsum);
      } else {
        auto kern = layernorm_bwd_ring_kernel<T, V, NC, 1, kRing, kRingWarps>;
        AGA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(224 * 1024 - kRingStatic)));
        kern<<<rgrid, kRingWarps * 32, ring_bytes, s>>>(static_cast<const T*>(dy), static_cast<const T*>(x), rows, gamma, mean, rstd,
                                                       static_cast<const T*>(dres), static_cast<T*>(dx), dgamma, dbeta, nullptr);
      }
      AGA_AFTER_LAUNCH();
      return AGA_OK;
    }
    if (dxsum) {
      layernorm_bwd_kernel<T, V, NC, 2><<<grid, kLnWarps * 32, 0, s>>>(static_cast<const T*>(dy), static_cast<const T*>(x), rows,
                                                                        gamma, mean, rstd, static_cast<const T*>(dres), static_cast<T*>(dx), dgamma, dbeta, dxsum);
    } else {
      layernorm_bwd_kernel<T, V, NC, 1><<<grid, kLnWarps * 32, 0, s>>>(static_cast<const T*>(dy), static_cast<const T*>(x), rows,
                                                                        gamma, mean, rstd, static_cast<const T*>(dres), static_cast<T*>(dx), dgamma, dbeta, nullptr);
    }
  } else {
    layernorm_bwd_kernel<T, V, NC, 0><<<grid, kLnWarps * 32, 0, s>>>(static_cast<const T*>(dy), static_cast<const T*>(x), rows, gamma,
                                                                      mean, rstd, static_cast<const T*>(dres), static_cast<T*>(dx), nullptr, nullptr, nullptr);
  }
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

// bf16 rows of 512 / 768 / 1024 / 1280 columns use 16-byte (8-element) accesses; 384 (whisper-tiny) and fp32 rows 4 elements
#define AGA_LN_DISPATCH(FN, ...)                                                                 \
  switch (D) {                                                                                   \
    case 384:  return bf16 ? FN<__nv_bfloat16, 4, 3>(__VA_ARGS__) : FN<float, 4, 3>(__VA_ARGS__);  \
    case 512:  return bf16 ? FN<__nv_bfloat16, 8, 2>(__VA_ARGS__) : FN<float, 4, 4>(__VA_ARGS__);  \
    case 768:  return bf16 ? FN<__nv_bfloat16, 8, 3>(__VA_ARGS__) : FN<float, 4, 6>(__VA_ARGS__);  \
    case 1024: return bf16 ? FN<__nv_bfloat16, 8, 4>(__VA_ARGS__) : FN<float, 4, 8>(__VA_ARGS__);  \
    case 1280: return bf16 ? FN<__nv_bfloat16, 8, 5>(__VA_ARGS__) : FN<float, 4, 10>(__VA_ARGS__); \
    default: return AGA_ERR_UNSUPPORTED;                                                         \
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
  AGA_LN_DISPATCH(launch_fwd, x, residual, rows, gamma, beta, eps, y, sum_out, mean, rstd, nullptr, nullptr, 0.f, nullptr, nullptr,
                  nullptr, s)
}

extern "C" int aga_layernorm_pair_fwd(const void* x, const void* residual, int dtype, int64_t rows, int D, const float* gamma,
                                      const float* beta, float eps, void* y, void* sum_out, float* mean, float* rstd,
                                      const float* gamma2, const float* beta2, float eps2, void* y2, float* mean2, float* rstd2,
                                      void* stream) {
  int st = check(x, y, dtype, rows, D);
  if (st != AGA_OK) return st;
  if (!gamma || !beta || !mean || !rstd || (sum_out && !residual) || !gamma2 || !beta2 || !y2 || !mean2 || !rstd2)
    return AGA_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(sum_out) | reinterpret_cast<uintptr_t>(y2)) & 15)
    return AGA_ERR_UNSUPPORTED;
  const bool bf16 = dtype == AGA_BF16;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  AGA_LN_DISPATCH(launch_fwd, x, residual, rows, gamma, beta, eps, y, sum_out, mean, rstd, gamma2, beta2, eps2, y2, mean2, rstd2, s)
}

namespace {
int layernorm_bwd_impl(const void* dy, const void* x, int dtype, int64_t rows, int D, const float* gamma, const float* mean,
                       const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta, float* dxsum, bool accumulate,
                       void* stream) {
  int st = check(x, dx, dtype, rows, D);
  if (st != AGA_OK) return st;
  if (!dy || !gamma || !mean || !rstd || ((dgamma == nullptr) != (dbeta == nullptr)) || (dxsum && !dgamma))
    return AGA_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dres)) & 15) return AGA_ERR_UNSUPPORTED;
  const bool bf16 = dtype == AGA_BF16;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  AGA_LN_DISPATCH(launch_bwd, dy, x, rows, gamma, mean, rstd, dres, dx, dgamma, dbeta, dxsum, accumulate, s)
}
}  // namespace

extern "C" int aga_layernorm_bwd(const void* dy, const void* x, int dtype, int64_t rows, int D, const float* gamma,
                                 const float* mean, const float* rstd, const void* dres, void* dx, float* dgamma,
                                 float* dbeta, float* dxsum, void* stream) {
  return layernorm_bwd_impl(dy, x, dtype, rows, D, gamma, mean, rstd, dres, dx, dgamma, dbeta, dxsum, false, stream);
}
extern "C" int aga_layernorm_bwd_acc(const void* dy, const void* x, int dtype, int64_t rows, int D, const float* gamma,
                                     const float* mean, const float* rstd, const void* dres, void* dx, float* dgamma,
                                     float* dbeta, float* dxsum, void* stream) {
  return layernorm_bwd_impl(dy, x, dtype, rows, D, gamma, mean, rstd, dres, dx, dgamma, dbeta, dxsum, true, stream);
}
