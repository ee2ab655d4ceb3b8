// Label-smoothing cross entropy + accuracy on the decoder's vocabulary logits, without fp32 logits.
//
// Reference: OpenAIWhisperDecoder.forward returns `(x @ tok_emb^T).float()` (espnet2/asr/decoder/whisper_decoder.py:164-166,
// (B,T,51865) fp32 = 212 MB), LabelSmoothingLoss clones it twice (espnet/nets/pytorch_backend/transformer/
// label_smoothing_loss.py:41-63: log_softmax, true_dist, KLDiv) and th_accuracy takes an argmax over it
// (espnet/nets/pytorch_backend/nets_utils.py:304-324).  Here the (rows, ld) logits stay in the GEMM's dtype:
//   forward : one CTA per row, ONE pass — online max / sum-exp, sum of the logits, arg-max, the target's logit —
//             -> per-row KL loss, log-sum-exp and "prediction == target" flag;
//   backward: dlogits = g/denom * (softmax - smoothed one-hot), recomputed from the logits and the stored log-sum-exp
//             (one read, one write), zero in the padding columns [V, ld) and in ignored rows.
// KL(t || softmax) with t = 1-eps on the target and eps/(V-1) elsewhere:
//   kl = conf*log(conf) + eps*log(eps/(V-1)) - [ eps/(V-1) * (sum_v logp_v - logp_t) + conf * logp_t ].
#include "aga_common.cuh"

namespace aga {
namespace {

constexpr int kCeThreads = 256;

template <typename T> struct Ld;
template <> struct Ld<float> {
  static constexpr int kVec = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Ld<__nv_bfloat16> {
  static constexpr int kVec = 8;
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

struct RowStat {
  float m, s, sx;  // running max, sum exp(x - m), sum x
  float bv;        // best value
  int bi;          // its (lowest) index
};
__device__ __forceinline__ RowStat merge(const RowStat& a, const RowStat& b) {
  RowStat r;
  r.m = fmaxf(a.m, b.m);
  r.s = (a.m == -INFINITY ? 0.f : a.s * __expf(a.m - r.m)) + (b.m == -INFINITY ? 0.f : b.s * __expf(b.m - r.m));
  r.sx = a.sx + b.sx;
  const bool take_b = b.bv > a.bv || (b.bv == a.bv && b.bi < a.bi);
  r.bv = take_b ? b.bv : a.bv;
  r.bi = take_b ? b.bi : a.bi;
  return r;
}

template <typename T>
__global__ void __launch_bounds__(kCeThreads)
ls_ce_fwd_kernel(const T* __restrict__ logits, int V, int64_t ld, const int64_t* __restrict__ target, int64_t padding_idx,
                 float smoothing, float* __restrict__ row_loss, float* __restrict__ row_lse, int32_t* __restrict__ row_correct) {
  constexpr int VEC = Ld<T>::kVec;
  const int64_t row = blockIdx.x;
  const T* x = logits + row * ld;
  RowStat st{-INFINITY, 0.f, 0.f, -INFINITY, 0x7fffffff};
  const int n_vec = (V + VEC - 1) / VEC;  // the row is padded to a multiple of VEC elements (ld % VEC == 0)
  for (int i = threadIdx.x; i < n_vec; i += kCeThreads) {
    float v[VEC];
    Ld<T>::load(x + i * VEC, v);
    float mloc = -INFINITY;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      if (i * VEC + e >= V) v[e] = -INFINITY;  // padding columns of the aligned GEMM
      mloc = fmaxf(mloc, v[e]);
    }
    if (mloc > st.m) {
      st.s *= __expf(st.m - mloc);  // exp(-inf) = 0 on the first vector
      st.m = mloc;
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      if (i * VEC + e < V) {
        st.s += __expf(v[e] - st.m);
        st.sx += v[e];
        if (v[e] > st.bv) {  // strictly greater: the lowest index of the maximum wins, as torch.argmax
          st.bv = v[e];
          st.bi = i * VEC + e;
        }
      }
    }
  }
  // block reduction
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    RowStat other;
    other.m = __shfl_xor_sync(0xffffffffu, st.m, o);
    other.s = __shfl_xor_sync(0xffffffffu, st.s, o);
    other.sx = __shfl_xor_sync(0xffffffffu, st.sx, o);
    other.bv = __shfl_xor_sync(0xffffffffu, st.bv, o);
    other.bi = __shfl_xor_sync(0xffffffffu, st.bi, o);
    st = merge(st, other);
  }
  __shared__ RowStat sh[kCeThreads / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = st;
  __syncthreads();
  if (threadIdx.x == 0) {
    RowStat t = sh[0];
    for (int w = 1; w < kCeThreads / 32; ++w) t = merge(t, sh[w]);
    const float lse = t.m + logf(t.s);
    const int64_t tg = target[row];
    const bool ignore = tg == padding_idx;
    float kl = 0.f;
    int32_t correct = 0;
    if (!ignore && (tg < 0 || tg >= V)) {
      // a target outside [0, V) that is not the padding index (token list / vocabulary mismatch): the reference's
      // scatter_ raises there (label_smoothing_loss.py:56).  No out-of-bounds read here: the row's loss is NaN, which
      // the caller sees in the loss and the training step's non-finite check turns into a skipped update.
      kl = __int_as_float(0x7fc00000);
    } else if (!ignore) {
      const float conf = 1.0f - smoothing, eps = smoothing / float(V - 1);
      float c = conf > 0.f ? conf * logf(conf) : 0.f;
      if (eps > 0.f) c += float(V - 1) * eps * logf(eps);
      const float lp_t = to_f32<T>(x[tg]) - lse;
      const float sum_lp = t.sx - float(V) * lse;
      kl = c - (eps * (sum_lp - lp_t) + conf * lp_t);
      correct = (int64_t(t.bi) == tg) ? 1 : 0;
    }
    row_loss[row] = kl;
    row_lse[row] = lse;
    row_correct[row] = correct;
  }
}

template <typename T>
__global__ void __launch_bounds__(kCeThreads)
ls_ce_bwd_kernel(const T* __restrict__ logits, int V, int64_t ld, int width, const int64_t* __restrict__ target,
                 int64_t padding_idx, float smoothing, const float* __restrict__ row_lse, const float* __restrict__ gscale,
                 float inv_denom, T* __restrict__ dlogits) {
  constexpr int VEC = Ld<T>::kVec;
  const int64_t row = blockIdx.y;
  const T* x = logits + row * ld;
  T* dx = dlogits + row * ld;
  const int64_t tg = target[row];
  const bool ignore = tg == padding_idx;
  const bool bad = !ignore && (tg < 0 || tg >= V);  // out-of-range target: NaN gradient row (see the forward kernel)
  const float g = bad ? __int_as_float(0x7fc00000) : (gscale ? *gscale : 1.0f) * inv_denom;
  const float lse = row_lse[row];
  const float conf = 1.0f - smoothing, eps = smoothing / float(V - 1);
  const int n_vec = width / VEC;
  for (int i = blockIdx.x * kCeThreads + threadIdx.x; i < n_vec; i += gridDim.x * kCeThreads) {
    float v[VEC], o[VEC];
    if (ignore) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] = 0.f;
    } else {
      Ld<T>::load(x + i * VEC, v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const int j = i * VEC + e;
        o[e] = j < V ? g * (__expf(v[e] - lse) - (j == tg ? conf : eps)) : 0.f;
      }
    }
    Ld<T>::store(dx + i * VEC, o);
  }
}

int check(const void* logits, int dtype, int64_t rows, int V, int64_t ld, const void* target) {
  if (!logits || !target || rows <= 0 || V <= 1 || ld < V) return AGA_ERR_INVALID_ARGUMENT;
  if (dtype != AGA_F32 && dtype != AGA_BF16) return AGA_ERR_INVALID_ARGUMENT;
  const int vec = dtype == AGA_BF16 ? 8 : 4;
  if (ld % vec != 0 || (reinterpret_cast<uintptr_t>(logits) & 15)) return AGA_ERR_UNSUPPORTED;
  if (rows > 2147483647LL) return AGA_ERR_UNSUPPORTED;
  return AGA_OK;
}

}  // namespace
}  // namespace aga

using namespace aga;

extern "C" int aga_ls_ce_fwd(const void* logits, int dtype, int64_t rows, int V, int64_t ld, const int64_t* target,
                             int64_t padding_idx, float smoothing, float* row_loss, float* row_lse, int32_t* row_correct,
                             void* stream) {
  int st = check(logits, dtype, rows, V, ld, target);
  if (st != AGA_OK) return st;
  if (!row_loss || !row_lse || !row_correct) return AGA_ERR_INVALID_ARGUMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == AGA_BF16)
    ls_ce_fwd_kernel<__nv_bfloat16><<<unsigned(rows), kCeThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(logits), V, ld, target,
                                                                           padding_idx, smoothing, row_loss, row_lse, row_correct);
  else
    ls_ce_fwd_kernel<float><<<unsigned(rows), kCeThreads, 0, s>>>(static_cast<const float*>(logits), V, ld, target, padding_idx,
                                                                   smoothing, row_loss, row_lse, row_correct);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

extern "C" int aga_ls_ce_bwd(const void* logits, int dtype, int64_t rows, int V, int64_t ld, const int64_t* target,
                             int64_t padding_idx, float smoothing, const float* row_lse, const float* gscale, float inv_denom,
                             void* dlogits, void* stream) {
  int st = check(logits, dtype, rows, V, ld, target);
  if (st != AGA_OK) return st;
  if (!row_lse || !dlogits || (reinterpret_cast<uintptr_t>(dlogits) & 15)) return AGA_ERR_INVALID_ARGUMENT;
  if (rows > 65535) return AGA_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int vec = dtype == AGA_BF16 ? 8 : 4;
  const int n_vec = int(ld / vec);
  const unsigned gx = unsigned(std::max(1, std::min((n_vec + kCeThreads - 1) / kCeThreads, 8)));
  dim3 grid(gx, unsigned(rows));
  if (dtype == AGA_BF16)
    ls_ce_bwd_kernel<__nv_bfloat16><<<grid, kCeThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(logits), V, ld, int(ld), target,
                                                                 padding_idx, smoothing, row_lse, gscale, inv_denom,
                                                                 static_cast<__nv_bfloat16*>(dlogits));
  else
    ls_ce_bwd_kernel<float><<<grid, kCeThreads, 0, s>>>(static_cast<const float*>(logits), V, ld, int(ld), target, padding_idx,
                                                         smoothing, row_lse, gscale, inv_denom, static_cast<float*>(dlogits));
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}
