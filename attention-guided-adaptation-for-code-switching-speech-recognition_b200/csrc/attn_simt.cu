// fp32-accumulate CUDA-core flash attention, forward and backward, head dim 64.
//
// This is the exact-arithmetic path of the library: it serves fp32 inputs (1e-4 parity with the
// reference's fp32 run, greedy decode of config 1) and the small decoder self-attention maps whose
// columns are exported.  bf16 encoder / cross attention goes to the tcgen05 kernels in attn_tc.cu.
//
// Reference: MultiHeadAttention.qkv_attention, whisper/whisper/model.py:93-109
//   S = (q d^-1/4)(k d^-1/4)^T + mask ; qk = S.float() ; w = softmax(qk) ; out = w v ; returns (out, qk)
// With d = 64 the two d^-1/4 factors are one exact 1/8 applied to the fp32 dot product.
//
// Tiling: one CTA (256 threads) per 64-query tile of one (batch, head); 64-key tiles stream through
// shared memory; every thread owns a 4x4 micro-tile.  Tiles are stored with a 68-float row pitch so
// that all inner-loop reads are conflict-free LDS.128.
#include "aga_common.cuh"
#include "attn_common.cuh"

#include <cooperative_groups.h>

namespace aga {
namespace {

constexpr int kBM = 64;
constexpr int kBN = 64;
constexpr int kD = 64;
constexpr int kLd = 68;
constexpr int kTile = kBM * kLd;  // floats per smem tile
constexpr int kThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kScale = 0.125f;  // (64^-1/4)^2, exact

// ---- global -> smem tile load (rows [row0, row0+64) of one head), zero-filled past `rows`
template <typename T>
__device__ __forceinline__ void load_tile(float* __restrict__ dst, const T* __restrict__ src, int64_t stride_t,
                                          int row0, int rows, int tid);

template <>
__device__ __forceinline__ void load_tile<float>(float* __restrict__ dst, const float* __restrict__ src,
                                                 int64_t stride_t, int row0, int rows, int tid) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = tid + it * kThreads;  // 1024 float4 per tile
    const int r = idx >> 4, c4 = idx & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < rows) v = __ldg(reinterpret_cast<const float4*>(src + int64_t(row0 + r) * stride_t) + c4);
    *reinterpret_cast<float4*>(dst + r * kLd + c4 * 4) = v;
  }
}

template <>
__device__ __forceinline__ void load_tile<__nv_bfloat16>(float* __restrict__ dst, const __nv_bfloat16* __restrict__ src,
                                                         int64_t stride_t, int row0, int rows, int tid) {
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = tid + it * kThreads;  // 512 x 16-byte chunks per tile
    const int r = idx >> 3, c8 = idx & 7;
    uint4 raw = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + r < rows) raw = __ldg(reinterpret_cast<const uint4*>(src + int64_t(row0 + r) * stride_t) + c8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
    float* d = dst + r * kLd + c8 * 8;
    const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
    const float2 f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
    *reinterpret_cast<float4*>(d) = make_float4(f0.x, f0.y, f1.x, f1.y);
    *reinterpret_cast<float4*>(d + 4) = make_float4(f2.x, f2.y, f3.x, f3.y);
  }
}

template <typename T>
__device__ __forceinline__ void store_row4(T* dst, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store_row4<float>(float* dst, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(dst) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store_row4<__nv_bfloat16>(__nv_bfloat16* dst, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(dst) = u;
}

// acc[i][j] = sum_c A[ty*4+i][c] * Bm[tx+16j][c]        (A, Bm: [64][kLd], contraction over the 64 columns)
__device__ __forceinline__ void micro_nt(const float* __restrict__ A, const float* __restrict__ Bm, int ty, int tx,
                                         float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int c4 = 0; c4 < kD / 4; ++c4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ty * 4 + i) * kLd + c4 * 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bm + (tx + 16 * j) * kLd + c4 * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
        acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
        acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
        acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
      }
  }
}

// acc[i][jj] += sum_k P[ty*4+i][k] * V[k][tx*4+jj]
__device__ __forceinline__ void micro_nn(const float* __restrict__ P, const float* __restrict__ V, int ty, int tx,
                                         float (&acc)[4][4]) {
#pragma unroll 4
  for (int k4 = 0; k4 < kBN / 4; ++k4) {
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(P + (ty * 4 + i) * kLd + k4 * 4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 b = *reinterpret_cast<const float4*>(V + (k4 * 4 + kk) * kLd + tx * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
        acc[i][0] = fmaf(av, b.x, acc[i][0]);
        acc[i][1] = fmaf(av, b.y, acc[i][1]);
        acc[i][2] = fmaf(av, b.z, acc[i][2]);
        acc[i][3] = fmaf(av, b.w, acc[i][3]);
      }
    }
  }
}

// acc[i][jj] += sum_q P[q][ty*4+i] * V[q][tx*4+jj]
__device__ __forceinline__ void micro_tn(const float* __restrict__ P, const float* __restrict__ V, int ty, int tx,
                                         float (&acc)[4][4]) {
#pragma unroll 8
  for (int q = 0; q < kBM; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(P + q * kLd + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(V + q * kLd + tx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
      acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
      acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
      acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
    }
  }
}

__device__ __forceinline__ float group16_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct SimtArgs {
  int B, H, Tq, Tk, causal;
  int export_kind, export_lo, export_hi;
  int64_t q_sb, q_st, k_sb, k_st, v_sb, v_st, o_sb, o_st;
  const void* q;
  const void* k;
  const void* v;
  void* out;
  float* lse;
  const uint8_t* head_sel;
  float* export_buf;
  // backward
  const void* dout;
  const float* d_export;
  const float* delta;
  void* dq;
  void* dk;
  void* dv;
  const int32_t* kv_len;  // forward only: device scalar, keys at or past it do not exist (static-shape decoding step)
};

__device__ __forceinline__ bool head_selected(const SimtArgs& a, int h) {
  return a.export_kind != AGA_EXPORT_NONE && a.export_buf != nullptr && (a.head_sel == nullptr || a.head_sel[h] != 0);
}

// ------------------------------------------------------------------------------------------ forward
template <typename T>
__global__ void __launch_bounds__(kThreads) attn_fwd_simt_kernel(const SimtArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;
  float* Ks = Qs + kTile;
  float* Vs = Ks + kTile;
  float* Ps = Vs + kTile;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row0 = qt * kBM;
  const T* qg = static_cast<const T*>(a.q) + b * a.q_sb + h * kD;
  const T* kg = static_cast<const T*>(a.k) + b * a.k_sb + h * kD;
  const T* vg = static_cast<const T*>(a.v) + b * a.v_sb + h * kD;
  const bool exp_on = head_selected(a, h);
  const int W = a.export_hi - a.export_lo;
  float* ebuf = exp_on ? a.export_buf + (int64_t(b) * a.H + h) * int64_t(a.Tq) * W : nullptr;

  load_tile<T>(Qs, qg, a.q_st, row0, a.Tq, tid);

  float m[4], l[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }
  const int Tk = a.kv_len ? max(1, min(__ldg(a.kv_len), a.Tk)) : a.Tk;
  int n_kt = (Tk + kBN - 1) / kBN;
  if (a.causal) n_kt = min(n_kt, qt + 1);

  for (int kt = 0; kt < n_kt; ++kt) {
    __syncthreads();  // previous tile's P V done (and Q visible on the first trip)
    load_tile<T>(Ks, kg, a.k_st, kt * kBN, Tk, tid);
    load_tile<T>(Vs, vg, a.v_st, kt * kBN, Tk, tid);
    __syncthreads();
    float s[4][4];
    micro_nt(Qs, Ks, ty, tx, s);
    float alpha[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = row0 + ty * 4 + i;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = kt * kBN + tx + 16 * j;
        float sv = s[i][j] * kScale;
        if (col >= Tk || (a.causal && col > row)) sv = -INFINITY;
        s[i][j] = sv;
        mx = fmaxf(mx, sv);
        if (exp_on && row < a.Tq && col >= a.export_lo && col < a.export_hi)
          ebuf[int64_t(row) * W + (col - a.export_lo)] = sv;
      }
      mx = group16_max(mx);
      const float m_new = fmaxf(m[i], mx);
      alpha[i] = exp2f((m[i] - m_new) * kLog2e);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = exp2f((s[i][j] - m_new) * kLog2e);
        rs += p;
        Ps[(ty * 4 + i) * kLd + tx + 16 * j] = p;
      }
      l[i] = l[i] * alpha[i] + rs;  // per-thread partial; the 16 lanes of a row are summed once at the end
      m[i] = m_new;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= alpha[i];
    }
    __syncthreads();
    micro_nn(Ps, Vs, ty, tx, o);
  }

  T* og = static_cast<T*>(a.out) + b * a.o_sb + h * kD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row0 + ty * 4 + i;
    const float lt = group16_sum(l[i]);
    if (row < a.Tq) {
      const float inv = 1.0f / lt;
      store_row4<T>(og + int64_t(row) * a.o_st + tx * 4, o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
      if (tx == 0 && a.lse) a.lse[(int64_t(b) * a.H + h) * a.Tq + row] = m[i] + logf(lt);
    }
  }
}

// ------------------------------------------------------------------------------------------ single-query forward
// Decoding step (Tq = 1 per hypothesis; whisper_decoder.py:172-244 recomputes the whole prefix instead): one query
// row against Tk keys is far too little work for a query-tile kernel — 12 CTAs walk 1500 keys serially (12.4 us on
// the tcgen05 kernel).  Here a CLUSTER of kDecSplit CTAs shares one (hypothesis, head): each CTA scans its slice of
// the keys (a lane owns a key for the dot product, then a pair of output columns for P V), the per-CTA
// (max, sum, out[64]) partials are combined by rank 0 through distributed shared memory.  No workspace, no atomics.
constexpr int kDecSplit = 8;
constexpr int kDecWarps = 4;

// two adjacent elements as they sit in memory; converted only where they are used, so that a batch of loads is issued
// back to back (a conversion right behind each load makes every load wait for the one before it)
__device__ __forceinline__ float2 load2_raw(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ uint32_t load2_raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ float2 cvt2(float2 r) { return r; }
__device__ __forceinline__ float2 cvt2(uint32_t r) { return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u)); }
template <typename T> struct Raw2;
template <> struct Raw2<float> { using type = float2; static __device__ __forceinline__ float2 zero() { return make_float2(0.f, 0.f); } };
template <> struct Raw2<__nv_bfloat16> { using type = uint32_t; static __device__ __forceinline__ uint32_t zero() { return 0u; } };
template <typename T> __device__ __forceinline__ float dot64(const float (&q)[kD], const T* __restrict__ row);
template <> __device__ __forceinline__ float dot64<float>(const float (&q)[kD], const float* __restrict__ row) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < kD / 4; ++c) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row) + c);
    acc = fmaf(q[4 * c], v.x, acc); acc = fmaf(q[4 * c + 1], v.y, acc);
    acc = fmaf(q[4 * c + 2], v.z, acc); acc = fmaf(q[4 * c + 3], v.w, acc);
  }
  return acc;
}
template <> __device__ __forceinline__ float dot64<__nv_bfloat16>(const float (&q)[kD], const __nv_bfloat16* __restrict__ row) {
  uint4 raw[kD / 8];
#pragma unroll
  for (int c = 0; c < kD / 8; ++c) raw[c] = __ldg(reinterpret_cast<const uint4*>(row) + c);
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < kD / 8; ++c) {
    const uint32_t w[4] = {raw[c].x, raw[c].y, raw[c].z, raw[c].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = cvt2(w[e]);
      acc = fmaf(q[8 * c + 2 * e], f.x, acc);
      acc = fmaf(q[8 * c + 2 * e + 1], f.y, acc);
    }
  }
  return acc;
}
__device__ __forceinline__ void store2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store2(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T>
__global__ void __cluster_dims__(kDecSplit, 1, 1) __launch_bounds__(kDecWarps * 32) attn_decode_kernel(const SimtArgs a) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ __align__(16) float qs[kD];
  __shared__ float wpart[kDecWarps][kD + 2];  // per warp: out[64], max, sum
  __shared__ __align__(8) float cpart[kD + 2];  // this CTA's partial, read by rank 0 of the cluster
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rank = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const T* qg = static_cast<const T*>(a.q) + b * a.q_sb + h * kD;
  const T* kg = static_cast<const T*>(a.k) + b * a.k_sb + h * kD;
  const T* vg = static_cast<const T*>(a.v) + b * a.v_sb + h * kD;
  if (threadIdx.x < kD) qs[threadIdx.x] = to_f32(qg[threadIdx.x]) * kScale;  // the exact 1/8 goes onto q
  __syncthreads();
  float q[kD];
#pragma unroll
  for (int c = 0; c < kD / 4; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(qs + 4 * c);
    q[4 * c] = v.x; q[4 * c + 1] = v.y; q[4 * c + 2] = v.z; q[4 * c + 3] = v.w;
  }
  const int Tk = a.kv_len ? max(1, min(__ldg(a.kv_len), a.Tk)) : a.Tk;
  const int chunk = ((Tk + kDecSplit - 1) / kDecSplit + 31) & ~31;  // whole 32-key blocks per CTA
  const int k0 = rank * chunk, k1 = min(Tk, k0 + chunk);
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int kb = k0 + 32 * warp; kb < k1; kb += 32 * kDecWarps) {
    const int key = kb + lane;
    // the block's V rows (lane = column pair) are requested BEFORE the scores exist: all of a block's global loads are
    // in flight together (a load per key behind the softmax cost 32 exposed latencies per block: 16.5 us per call)
    typename Raw2<T>::type vv[32];
#pragma unroll
    for (int j = 0; j < 32; ++j)
      vv[j] = kb + j < k1 ? load2_raw(vg + int64_t(kb + j) * a.v_st + 2 * lane) : Raw2<T>::zero();
    const float sv = key < k1 ? dot64<T>(q, kg + int64_t(key) * a.k_st) : -INFINITY;
    const float m_new = fmaxf(m, warp_max_f(sv));  // finite: the block holds at least one key
    const float alpha = exp2f((m - m_new) * kLog2e);
    const float pv = exp2f((sv - m_new) * kLog2e);  // 0 for the lanes past the end
    l = l * alpha + warp_sum(pv);
    o0 *= alpha;
    o1 *= alpha;
    m = m_new;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float pj = __shfl_sync(0xffffffffu, pv, j);  // 0 past the end
      const float2 v2 = cvt2(vv[j]);
      o0 = fmaf(pj, v2.x, o0);
      o1 = fmaf(pj, v2.y, o1);
    }
  }
  wpart[warp][2 * lane] = o0;
  wpart[warp][2 * lane + 1] = o1;
  if (lane == 0) {
    wpart[warp][kD] = m;
    wpart[warp][kD + 1] = l;
  }
  __syncthreads();
  if (warp == 0) {  // the CTA's partial
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) M = fmaxf(M, wpart[w][kD]);
    float L = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) {
      const float mw = wpart[w][kD];
      const float sc = mw == -INFINITY ? 0.f : exp2f((mw - M) * kLog2e);
      L = fmaf(sc, wpart[w][kD + 1], L);
      a0 = fmaf(sc, wpart[w][2 * lane], a0);
      a1 = fmaf(sc, wpart[w][2 * lane + 1], a1);
    }
    cpart[2 * lane] = a0;
    cpart[2 * lane + 1] = a1;
    if (lane == 0) {
      cpart[kD] = M;
      cpart[kD + 1] = L;
    }
  }
  cluster.sync();
  if (rank == 0 && warp == 0) {
    // all remote reads first (a distributed-shared-memory load is ~200 clk: 40 of them one after the other were the
    // kernel's fixed cost), then the combination in registers
    float mr[kDecSplit], lr[kDecSplit], p0[kDecSplit], p1[kDecSplit];
#pragma unroll
    for (int r = 0; r < kDecSplit; ++r) {
      const float* pr = cluster.map_shared_rank(cpart, r);
      mr[r] = pr[kD];
      lr[r] = pr[kD + 1];
      const float2 t = *reinterpret_cast<const float2*>(pr + 2 * lane);
      p0[r] = t.x;
      p1[r] = t.y;
    }
    float M = -INFINITY;
#pragma unroll
    for (int r = 0; r < kDecSplit; ++r) M = fmaxf(M, mr[r]);
    float L = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int r = 0; r < kDecSplit; ++r) {
      const float sc = mr[r] == -INFINITY ? 0.f : exp2f((mr[r] - M) * kLog2e);
      L = fmaf(sc, lr[r], L);
      a0 = fmaf(sc, p0[r], a0);
      a1 = fmaf(sc, p1[r], a1);
    }
    const float inv = 1.0f / L;
    store2(static_cast<T*>(a.out) + b * a.o_sb + h * kD + 2 * lane, a0 * inv, a1 * inv);
    if (lane == 0 && a.lse) a.lse[int64_t(b) * a.H + h] = M + logf(L);
  }
  cluster.sync();  // nobody leaves while rank 0 may still read its shared memory
}

// probs export = exp(exported logits - lse), in place (-inf -> 0)
__global__ void __launch_bounds__(256) export_logits_to_probs_kernel(const SimtArgs a) {
  const int W = a.export_hi - a.export_lo;
  const int64_t rows = int64_t(a.B) * a.H * a.Tq;
  const int64_t total = rows * W;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / W;
    const int h = int((r / a.Tq) % a.H);
    if (a.head_sel && !a.head_sel[h]) continue;
    a.export_buf[i] = expf(a.export_buf[i] - a.lse[r]);
  }
}

// ------------------------------------------------------------------------------------------ backward
// delta[b,h,t] = sum_c dO*O  (+ sum_w P*G over the exported columns when probabilities were exported)
template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(const SimtArgs a) {
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t rows = int64_t(a.B) * a.H * a.Tq;
  if (warp >= rows) return;
  const int t = int(warp % a.Tq);
  const int h = int((warp / a.Tq) % a.H);
  const int b = int(warp / (int64_t(a.Tq) * a.H));
  const T* o = static_cast<const T*>(a.out) + b * a.o_sb + int64_t(t) * a.o_st + h * kD;
  const T* d = static_cast<const T*>(a.dout) + b * a.o_sb + int64_t(t) * a.o_st + h * kD;
  float acc = to_f32(o[lane]) * to_f32(d[lane]) + to_f32(o[lane + 32]) * to_f32(d[lane + 32]);
  if (a.export_kind == AGA_EXPORT_PROBS && a.d_export && a.export_buf && (!a.head_sel || a.head_sel[h])) {
    const int W = a.export_hi - a.export_lo;
    for (int w = lane; w < W; w += 32) acc += a.export_buf[warp * W + w] * a.d_export[warp * W + w];
  }
  acc = warp_sum(acc);
  if (lane == 0) const_cast<float*>(a.delta)[warp] = acc;
}

// Shared by both backward kernels: S, P and dS micro-tiles for (q tile, k tile) from smem tiles.
__device__ __forceinline__ void bwd_tile_p_ds(const SimtArgs& a, const float* Qs, const float* Ks, const float* Vs,
                                              const float* dOs, int ty, int tx, int row0, int col0, int b, int h,
                                              bool exp_on, const float (&lse)[4], const float (&dl)[4],
                                              float (&p)[4][4], float (&ds)[4][4]) {
  float s[4][4], dp[4][4];
  micro_nt(Qs, Ks, ty, tx, s);
  micro_nt(dOs, Vs, ty, tx, dp);
  const int W = a.export_hi - a.export_lo;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx + 16 * j;
      const bool masked = col >= a.Tk || (a.causal && col > row) || row >= a.Tq;
      const float sv = s[i][j] * kScale;
      const float pv = masked ? 0.f : exp2f((sv - lse[i]) * kLog2e);
      float dpv = dp[i][j];
      float g = 0.f;
      if (exp_on && !masked && col >= a.export_lo && col < a.export_hi)
        g = a.d_export[((int64_t(b) * a.H + h) * a.Tq + row) * W + (col - a.export_lo)];
      if (a.export_kind == AGA_EXPORT_PROBS) dpv += g;
      float dsv = pv * (dpv - dl[i]);
      if (a.export_kind == AGA_EXPORT_LOGITS) dsv += g;
      p[i][j] = pv;
      ds[i][j] = dsv;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) attn_bwd_dq_simt_kernel(const SimtArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;
  float* dOs = Qs + kTile;
  float* Ks = dOs + kTile;
  float* Vs = Ks + kTile;
  float* dSs = Vs + kTile;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row0 = qt * kBM;
  const T* qg = static_cast<const T*>(a.q) + b * a.q_sb + h * kD;
  const T* kg = static_cast<const T*>(a.k) + b * a.k_sb + h * kD;
  const T* vg = static_cast<const T*>(a.v) + b * a.v_sb + h * kD;
  const T* dog = static_cast<const T*>(a.dout) + b * a.o_sb + h * kD;
  const bool exp_on = head_selected(a, h) && a.d_export != nullptr;
  load_tile<T>(Qs, qg, a.q_st, row0, a.Tq, tid);
  load_tile<T>(dOs, dog, a.o_st, row0, a.Tq, tid);
  float lse[4], dl[4], dq[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row0 + ty * 4 + i;
    const int64_t r = (int64_t(b) * a.H + h) * a.Tq + row;
    lse[i] = row < a.Tq ? a.lse[r] : 0.f;
    dl[i] = row < a.Tq ? a.delta[r] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
  }
  int n_kt = (a.Tk + kBN - 1) / kBN;
  if (a.causal) n_kt = min(n_kt, qt + 1);
  for (int kt = 0; kt < n_kt; ++kt) {
    __syncthreads();
    load_tile<T>(Ks, kg, a.k_st, kt * kBN, a.Tk, tid);
    load_tile<T>(Vs, vg, a.v_st, kt * kBN, a.Tk, tid);
    __syncthreads();
    float p[4][4], ds[4][4];
    bwd_tile_p_ds(a, Qs, Ks, Vs, dOs, ty, tx, row0, kt * kBN, b, h, exp_on, lse, dl, p, ds);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dSs[(ty * 4 + i) * kLd + tx + 16 * j] = ds[i][j];
    __syncthreads();
    micro_nn(dSs, Ks, ty, tx, dq);
  }
  T* dqg = static_cast<T*>(a.dq) + b * a.q_sb + h * kD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row0 + ty * 4 + i;
    if (row < a.Tq)
      store_row4<T>(dqg + int64_t(row) * a.q_st + tx * 4, dq[i][0] * kScale, dq[i][1] * kScale, dq[i][2] * kScale,
                    dq[i][3] * kScale);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) attn_bwd_dkv_simt_kernel(const SimtArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;
  float* Vs = Ks + kTile;
  float* Qs = Vs + kTile;
  float* dOs = Qs + kTile;
  float* Ps = dOs + kTile;
  float* dSs = Ps + kTile;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int col0 = kt * kBN;
  const T* qg = static_cast<const T*>(a.q) + b * a.q_sb + h * kD;
  const T* kg = static_cast<const T*>(a.k) + b * a.k_sb + h * kD;
  const T* vg = static_cast<const T*>(a.v) + b * a.v_sb + h * kD;
  const T* dog = static_cast<const T*>(a.dout) + b * a.o_sb + h * kD;
  const bool exp_on = head_selected(a, h) && a.d_export != nullptr;
  load_tile<T>(Ks, kg, a.k_st, col0, a.Tk, tid);
  load_tile<T>(Vs, vg, a.v_st, col0, a.Tk, tid);
  float dk[4][4], dv[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dk[i][j] = dv[i][j] = 0.f;
  const int n_qt = (a.Tq + kBM - 1) / kBM;
  const int qt0 = a.causal ? kt : 0;
  for (int qt = qt0; qt < n_qt; ++qt) {
    const int row0 = qt * kBM;
    __syncthreads();
    load_tile<T>(Qs, qg, a.q_st, row0, a.Tq, tid);
    load_tile<T>(dOs, dog, a.o_st, row0, a.Tq, tid);
    float lse[4], dl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = row0 + ty * 4 + i;
      const int64_t r = (int64_t(b) * a.H + h) * a.Tq + row;
      lse[i] = row < a.Tq ? a.lse[r] : 0.f;
      dl[i] = row < a.Tq ? a.delta[r] : 0.f;
    }
    __syncthreads();
    float p[4][4], ds[4][4];
    bwd_tile_p_ds(a, Qs, Ks, Vs, dOs, ty, tx, row0, col0, b, h, exp_on, lse, dl, p, ds);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Ps[(ty * 4 + i) * kLd + tx + 16 * j] = p[i][j];
        dSs[(ty * 4 + i) * kLd + tx + 16 * j] = ds[i][j];
      }
    __syncthreads();
    micro_tn(Ps, dOs, ty, tx, dv);
    micro_tn(dSs, Qs, ty, tx, dk);
  }
  T* dkg = static_cast<T*>(a.dk) + b * a.k_sb + h * kD;
  T* dvg = static_cast<T*>(a.dv) + b * a.v_sb + h * kD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int key = col0 + ty * 4 + i;
    if (key < a.Tk) {
      store_row4<T>(dkg + int64_t(key) * a.k_st + tx * 4, dk[i][0] * kScale, dk[i][1] * kScale, dk[i][2] * kScale,
                    dk[i][3] * kScale);
      store_row4<T>(dvg + int64_t(key) * a.v_st + tx * 4, dv[i][0], dv[i][1], dv[i][2], dv[i][3]);
    }
  }
}

SimtArgs make_args(const aga_attn_params& p) {
  SimtArgs a{};
  a.B = p.B; a.H = p.H; a.Tq = p.Tq; a.Tk = p.Tk; a.causal = p.causal;
  a.export_kind = p.export_buf ? p.export_kind : AGA_EXPORT_NONE;
  a.export_lo = p.export_lo; a.export_hi = p.export_hi;
  a.q_sb = p.q_stride_b; a.q_st = p.q_stride_t; a.k_sb = p.k_stride_b; a.k_st = p.k_stride_t;
  a.v_sb = p.v_stride_b; a.v_st = p.v_stride_t; a.o_sb = p.o_stride_b; a.o_st = p.o_stride_t;
  a.q = p.q; a.k = p.k; a.v = p.v; a.out = p.out; a.lse = p.lse; a.head_sel = p.head_sel; a.export_buf = p.export_buf;
  a.kv_len = p.kv_len;
  return a;
}

template <typename T>
int launch_fwd(const aga_attn_params& p, cudaStream_t s) {
  SimtArgs a = make_args(p);
  const size_t smem = 4 * kTile * sizeof(float);
  AGA_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  dim3 grid((p.Tq + kBM - 1) / kBM, p.H, p.B);
  attn_fwd_simt_kernel<T><<<grid, kThreads, smem, s>>>(a);
  AGA_AFTER_LAUNCH();
  if (a.export_kind == AGA_EXPORT_PROBS) {
    const int64_t total = int64_t(p.B) * p.H * p.Tq * (p.export_hi - p.export_lo);
    const unsigned gx = unsigned(std::min<int64_t>((total + 255) / 256, 148 * 8));
    export_logits_to_probs_kernel<<<gx ? gx : 1, 256, 0, s>>>(a);
    AGA_AFTER_LAUNCH();
  }
  return AGA_OK;
}

template <typename T>
int launch_bwd(const aga_attn_bwd_params& bp, float* delta, cudaStream_t s) {
  const aga_attn_params& p = bp.fwd;
  SimtArgs a = make_args(p);
  a.dout = bp.dout; a.d_export = bp.d_export; a.delta = delta; a.dq = bp.dq; a.dk = bp.dk; a.dv = bp.dv;
  if (!bp.d_export) a.export_kind = AGA_EXPORT_NONE;
  const int64_t rows = int64_t(p.B) * p.H * p.Tq;
  attn_bwd_delta_kernel<T><<<unsigned((rows * 32 + 255) / 256), 256, 0, s>>>(a);
  AGA_AFTER_LAUNCH();
  const size_t smem_dq = 5 * kTile * sizeof(float), smem_dkv = 6 * kTile * sizeof(float);
  AGA_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dq_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_dq)));
  AGA_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dkv_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_dkv)));
  attn_bwd_dq_simt_kernel<T><<<dim3((p.Tq + kBM - 1) / kBM, p.H, p.B), kThreads, smem_dq, s>>>(a);
  AGA_AFTER_LAUNCH();
  attn_bwd_dkv_simt_kernel<T><<<dim3((p.Tk + kBN - 1) / kBN, p.H, p.B), kThreads, smem_dkv, s>>>(a);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

}  // namespace

bool attn_decode_shape(const aga_attn_params& p) {
  return p.Tq == 1 && !p.causal && (p.export_kind == AGA_EXPORT_NONE || !p.export_buf) && !p.guided_part &&
         p.impl != AGA_ATTN_TCGEN05;
}
int attn_decode_fwd(const aga_attn_params& p, cudaStream_t s) {
  SimtArgs a = make_args(p);
  const dim3 grid(kDecSplit, p.H, p.B);
  if (p.dtype == AGA_BF16) attn_decode_kernel<__nv_bfloat16><<<grid, kDecWarps * 32, 0, s>>>(a);
  else attn_decode_kernel<float><<<grid, kDecWarps * 32, 0, s>>>(a);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}
int attn_simt_fwd(const aga_attn_params& p, cudaStream_t s) {
  return p.dtype == AGA_BF16 ? launch_fwd<__nv_bfloat16>(p, s) : launch_fwd<float>(p, s);
}
size_t attn_simt_bwd_workspace(const aga_attn_params& p) {
  return align_up(size_t(p.B) * p.H * p.Tq * sizeof(float), 256);
}
int attn_simt_bwd(const aga_attn_bwd_params& bp, void* ws, cudaStream_t s) {
  float* delta = static_cast<float*>(ws);
  return bp.fwd.dtype == AGA_BF16 ? launch_bwd<__nv_bfloat16>(bp, delta, s) : launch_bwd<float>(bp, delta, s);
}

}  // namespace aga
