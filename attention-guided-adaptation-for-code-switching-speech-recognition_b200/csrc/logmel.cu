// Whisper log-mel frontend for sm_100a (fp32, CUDA cores; HBM-bound by design).
//
// Replaces OpenAIWhisperEncoder.log_mel_spectrogram (espnet2/asr/encoder/whisper_encoder.py:105-135):
//   torch.stft(center, reflect, hann 400, hop 160) -> drop last frame -> |.|^2 -> mel_80 @ P ->
//   log10(clamp 1e-10) -> max(x, utt_max - 8) -> (x + 4) / 4
// as two kernels:
//   1. logmel_frames_kernel: PERSISTENT CTAs (two per SM) walk blocks of 32 consecutive frames of one utterance;
//      the constant tables (window, twiddles, filter bands) are loaded once per CTA and the NEXT block's samples
//      are fetched into registers with 128-bit loads while the current block is transformed (HBM latency off
//      the critical path; it was 30 % of the stall samples).  A block's 5360 samples sit in shared memory ONCE
//      (each HBM sample is read once, the 2.5x frame overlap is served from smem); per frame a 400-point real DFT
//      factored 400 = 16 x 25 (16 real 25-point DFTs with constant roots in registers, W_400 twiddles,
//      13 complex 16-point FFTs; the other 12 residues follow from Hermitian symmetry), the power
//      spectrum (kept [bin][frame], frame fastest, so that the projection reads it conflict-free), the banded mel
//      projection and log10 (MUFU lg2); it stores the un-normalised log-mel tile with 128-byte coalesced rows and
//      folds the per-utterance maximum into one atomicMax per warp.
//   2. logmel_normalise_kernel: max(x, m_b - 8), (x + 4)/4 in place (the tile is L2-resident).
// Algorithmic HBM bytes per utterance: N*4 read + n_mels*(N/160)*4 written.
#include "aga_common.cuh"

#include <algorithm>
#include <type_traits>

namespace aga {
namespace {

#include "logmel_tables.inc"

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kNfreq = 201;
constexpr int kFramesPerCta = 32;
constexpr int kFramesPerPass = 16;
constexpr int kThreads = 256;
constexpr int kChunk = (kFramesPerCta - 1) * kHop + kNfft;  // 5360 samples staged per CTA
constexpr int kMaxMels = 256;
constexpr int kWSmem = 2048;  // filter weights cached in smem (80-mel: 391, 128-mel: 394)

// smem sample index: 16 floats of padding per hop so that the two frames a warp works on in
// stage 1 (160 samples apart = same banks) land 16 banks apart.
__host__ __device__ constexpr int pad_idx(int i) { return i + 16 * (i / kHop); }
constexpr int kSampSmem = pad_idx(kChunk - 1) + 1 + 3;  // 5891 -> keep 4-aligned below
constexpr int kSampSmemAl = (kSampSmem + 3) / 4 * 4;
constexpr int kXchK1Stride = 36;                    // floats: 16 complex + 4 pad (conflict-free LDS.128)
constexpr int kXchFrameStride = 13 * kXchK1Stride;  // 468
constexpr int kXchSmem = kFramesPerPass * kXchFrameStride;
constexpr int kPowStride = kFramesPerCta + 1;  // [bin][frame]: odd stride -> stage 2's strided bins and the projection's rows are conflict-free
constexpr int kPowSmem = (kNfreq * kPowStride + 3) / 4 * 4;  // keeps the tables behind it 16-byte aligned
constexpr int kStageVecs = (kChunk / 4 + kThreads - 1) / kThreads;  // float4 per thread of a block's samples (6)
constexpr int kTwSmem = 13 * 16 * 2;
constexpr int kSmemFloats = kSampSmemAl + kXchSmem + kPowSmem + kNfft + kTwSmem + 3 * kMaxMels + kWSmem;
constexpr size_t kSmemBytes = size_t(kSmemFloats) * 4;

struct PackedHeader {
  int32_t n_mels;
  int32_t total;
  int32_t pad0, pad1;
};
// packed layout: PackedHeader | int start[n_mels] | int count[n_mels] | int offset[n_mels] | float w[n_mels*201]

// Compile-time loop: the body sees the index as a constant expression, so table lookups become immediates.
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

__device__ __forceinline__ float2 cmul(float2 a, float cr, float ci) {
  return make_float2(a.x * cr - a.y * ci, a.x * ci + a.y * cr);
}

// In-register 16-point complex DFT (4 x 4 Cooley-Tukey), natural order in and out.
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  float2 t[4][4];
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) {
    const float2 a0 = v[n2], a1 = v[4 + n2], a2 = v[8 + n2], a3 = v[12 + n2];
    const float2 s0 = make_float2(a0.x + a2.x, a0.y + a2.y);
    const float2 s1 = make_float2(a0.x - a2.x, a0.y - a2.y);
    const float2 s2 = make_float2(a1.x + a3.x, a1.y + a3.y);
    const float2 s3 = make_float2(a1.y - a3.y, a3.x - a1.x);  // (a1 - a3) * (-i)
    t[n2][0] = make_float2(s0.x + s2.x, s0.y + s2.y);
    t[n2][1] = make_float2(s1.x + s3.x, s1.y + s3.y);
    t[n2][2] = make_float2(s0.x - s2.x, s0.y - s2.y);
    t[n2][3] = make_float2(s1.x - s3.x, s1.y - s3.y);
  }
  static_for<0, 4>([&](auto k1c) {
    constexpr int k1 = decltype(k1c)::value;
    float2 b[4];
    static_for<0, 4>([&](auto n2c) {
      constexpr int n2 = decltype(n2c)::value;
      constexpr int m = (n2 * k1) & 15;
      if constexpr (m == 0) {
        b[n2] = t[n2][k1];
      } else {
        constexpr float cr = kC16[m], ci = -kS16[m];
        b[n2] = cmul(t[n2][k1], cr, ci);
      }
    });
    const float2 s0 = make_float2(b[0].x + b[2].x, b[0].y + b[2].y);
    const float2 s1 = make_float2(b[0].x - b[2].x, b[0].y - b[2].y);
    const float2 s2 = make_float2(b[1].x + b[3].x, b[1].y + b[3].y);
    const float2 s3 = make_float2(b[1].y - b[3].y, b[3].x - b[1].x);
    v[k1] = make_float2(s0.x + s2.x, s0.y + s2.y);
    v[k1 + 4] = make_float2(s1.x + s3.x, s1.y + s3.y);
    v[k1 + 8] = make_float2(s0.x - s2.x, s0.y - s2.y);
    v[k1 + 12] = make_float2(s1.x - s3.x, s1.y - s3.y);
  });
}

__device__ __forceinline__ float load_sample_reflect(const float* __restrict__ row, int64_t g, int64_t N) {
  if (g < 0) g = -g;                   // reflect, no edge repeat: x_pad[199 - i] = x[i + 1]
  if (g >= N) g = 2 * (N - 1) - g;
  return (g >= 0 && g < N) ? __ldg(row + g) : 0.0f;
}

__global__ void __launch_bounds__(kThreads, 2)
logmel_frames_kernel(const float* __restrict__ audio, int64_t N, int64_t ld, int F, int n_fblocks, int n_blocks,
                     const unsigned char* __restrict__ packed, int n_mels, float* __restrict__ out,
                     uint32_t* __restrict__ maxkey) {
  extern __shared__ __align__(16) float smem[];
  float* s_samp = smem;
  float* s_xch = s_samp + kSampSmemAl;
  float* s_pow = s_xch + kXchSmem;
  float* s_win = s_pow + kPowSmem;
  float* s_tw = s_win + kNfft;
  int* s_fst = reinterpret_cast<int*>(s_tw + kTwSmem);
  float* s_fw = reinterpret_cast<float*>(s_fst + 3 * kMaxMels);

  const int tid = threadIdx.x;

  // ---- constant tables -> smem, once per (persistent) CTA
  const PackedHeader* hdr = reinterpret_cast<const PackedHeader*>(packed);
  const int* g_start = reinterpret_cast<const int*>(packed + sizeof(PackedHeader));
  const int* g_count = g_start + n_mels;
  const int* g_off = g_count + n_mels;
  const float* g_w = reinterpret_cast<const float*>(g_off + n_mels);
  const int total_w = hdr->total;
  const bool w_in_smem = total_w <= kWSmem;  // every stock filterbank (391 / 394 weights); dense custom ones read global
  for (int i = tid; i < n_mels; i += kThreads) {
    s_fst[i] = g_start[i];
    s_fst[kMaxMels + i] = g_count[i];
    s_fst[2 * kMaxMels + i] = g_off[i];
  }
  for (int i = tid; i < min(total_w, kWSmem); i += kThreads) s_fw[i] = g_w[i];
  for (int i = tid; i < kNfft; i += kThreads) s_win[i] = k_hann400[i];
  for (int i = tid; i < kTwSmem; i += kThreads) s_tw[i] = k_tw400[i];

  // samples of one block as kStageVecs float4 per thread (128-bit loads where the chunk is interior and aligned)
  auto fetch = [&](int blk, float4 (&val)[kStageVecs]) {
    const int b = blk / n_fblocks, f0 = (blk - b * n_fblocks) * kFramesPerCta;
    const float* arow = audio + int64_t(b) * ld;
    const int64_t g0 = int64_t(f0) * kHop - kNfft / 2;  // multiple of 8 samples
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(arow) & 15) == 0);
#pragma unroll
    for (int u = 0; u < kStageVecs; ++u) {
      const int c = tid + u * kThreads;
      const int64_t g = g0 + 4 * c;
      if (c >= kChunk / 4) {
        val[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (vec_ok && g >= 0 && g + 3 < N) {
        val[u] = __ldg(reinterpret_cast<const float4*>(arow + g));
      } else {
        val[u].x = load_sample_reflect(arow, g, N);
        val[u].y = load_sample_reflect(arow, g + 1, N);
        val[u].z = load_sample_reflect(arow, g + 2, N);
        val[u].w = load_sample_reflect(arow, g + 3, N);
      }
    }
  };

  const int slot = tid & 15;
  const int fpass = tid >> 4;  // frame within the pass
  const int lane = tid & 31, warp = tid >> 5;
  float4 nxt[kStageVecs];
  int blk = blockIdx.x;
  if (blk < n_blocks) fetch(blk, nxt);

  for (; blk < n_blocks; blk += gridDim.x) {
    const int b = blk / n_fblocks, f0 = (blk - b * n_fblocks) * kFramesPerCta;
    // ---- this block's samples: registers -> smem; then the next block's loads go out and fly during the transforms
#pragma unroll
    for (int u = 0; u < kStageVecs; ++u) {
      const int c = tid + u * kThreads;
      if (c < kChunk / 4) *reinterpret_cast<float4*>(s_samp + pad_idx(4 * c)) = nxt[u];
    }
    __syncthreads();
    if (blk + int(gridDim.x) < n_blocks) fetch(blk + gridDim.x, nxt);

#pragma unroll 1
    for (int pass = 0; pass < kFramesPerCta / kFramesPerPass; ++pass) {
      const int fl = pass * kFramesPerPass + fpass;
      // ---- stage 1: thread (frame, n2) — real 25-point DFT over n1 of x[16*n1 + n2], outputs k1 = 0..12
      {
        const int n2 = slot;
        const float* sp = s_samp + fl * (kHop + 16) + n2;
        float x[25];
#pragma unroll
        for (int n1 = 0; n1 < 25; ++n1) {
          const int j = 16 * n1;
          x[n1] = sp[j + 16 * (n1 / 10)] * s_win[j + n2];
        }
        float a[13], bb[13];
#pragma unroll
        for (int j = 1; j <= 12; ++j) {
          a[j] = x[j] + x[25 - j];
          bb[j] = x[j] - x[25 - j];
        }
        float2* xrow = reinterpret_cast<float2*>(s_xch + fpass * kXchFrameStride) + n2;
        static_for<0, 13>([&](auto k1c) {
          constexpr int k1 = decltype(k1c)::value;
          float re = x[0], im = 0.0f;
          static_for<1, 13>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            constexpr int m = (j * k1) % 25;
            constexpr float cr = kC25[m], nsi = -kS25[m];
            re = fmaf(a[j], cr, re);
            im = fmaf(bb[j], nsi, im);
          });
          const float2 tw = reinterpret_cast<const float2*>(s_tw)[k1 * 16 + n2];
          xrow[k1 * (kXchK1Stride / 2)] = make_float2(re * tw.x - im * tw.y, re * tw.y + im * tw.x);
        });
      }
      __syncthreads();
      // ---- stage 2: thread (frame, k1 < 13) — 16-point FFT over n2; bins k1 + 25*k2 and their mirrors
      if (slot < 13) {
        const int k1 = slot;
        const float4* src = reinterpret_cast<const float4*>(s_xch + fpass * kXchFrameStride + k1 * kXchK1Stride);
        float2 v[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = src[i];
          v[2 * i] = make_float2(t.x, t.y);
          v[2 * i + 1] = make_float2(t.z, t.w);
        }
        fft16(v);
        float* pcol = s_pow + fl;
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
          const float p = v[k2].x * v[k2].x + v[k2].y * v[k2].y;
          if (k2 < 8) {
            pcol[(k1 + 25 * k2) * kPowStride] = p;
          } else if (k1 > 0 || k2 == 8) {
            pcol[(kNfft - k1 - 25 * k2) * kPowStride] = p;  // |X[400-k]| = |X[k]|
          }
        }
      }
      __syncthreads();
    }

    // ---- mel projection + log10; lane = frame (coalesced 128-byte rows), warp strides over mel bins
    const int f = f0 + lane;
    const float* pcol = s_pow + lane;
    float vmax = -INFINITY;
    for (int m = warp; m < n_mels; m += kThreads / 32) {
      const int st = s_fst[m], cnt = s_fst[kMaxMels + m], off = s_fst[2 * kMaxMels + m];
      const float* pp = pcol + st * kPowStride;
      float acc0 = 0.0f, acc1 = 0.0f;
      if (w_in_smem) {
        const float* wp = s_fw + off;
        int i = 0;
        for (; i + 1 < cnt; i += 2) {
          acc0 = fmaf(wp[i], pp[i * kPowStride], acc0);
          acc1 = fmaf(wp[i + 1], pp[(i + 1) * kPowStride], acc1);
        }
        if (i < cnt) acc0 = fmaf(wp[i], pp[i * kPowStride], acc0);
      } else {
        for (int i = 0; i < cnt; ++i) acc0 = fmaf(__ldg(g_w + off + i), pp[i * kPowStride], acc0);
      }
      const float v = __log2f(fmaxf(acc0 + acc1, 1e-10f)) * 0.30102999566398120f;  // log10 via MUFU lg2 (abs error < 1e-7)
      if (f < F) {
        out[(int64_t(b) * n_mels + m) * F + f] = v;
        vmax = fmaxf(vmax, v);
      }
    }
    vmax = warp_max(vmax);
    if (lane == 0 && vmax > -INFINITY) atomicMax(maxkey + b, float_to_key(vmax));
    // s_samp / s_pow of this block are dead once every warp is here; the next iteration's first barrier orders the
    // sample stores against stage 1, this one orders them against the projection's reads of s_pow (next stage 2)
    __syncthreads();
  }
}

// max(x, m_b - 8), (x + 4) / 4 in place.  ``Fv`` < F (a batch zero-padded past its true common length, see
// aga_logmel_tc_fwd's n_valid): frames >= Fv were never computed and are written as exact zeros — what the conv stem's
// own zero padding would have supplied to the reference, which never sees those frames.
__global__ void __launch_bounds__(256)
logmel_normalise_kernel(float* __restrict__ out, const uint32_t* __restrict__ maxkey, int64_t per_utt, int vec, int F,
                        const int32_t* __restrict__ n_valid, int raw_power) {
  const int b = blockIdx.y;
  // raw_power (logmel_tc.cu): `out` and the maximum hold the mel POWER; log10(clamp 1e-10) happens here (monotone, so
  // the maximum of the logs is the log of the maximum)
  const float kLog10Of2 = 0.30102999566398120f;
  const float mx = key_to_float(maxkey[b]);
  const float floor_v = (raw_power ? __log2f(fmaxf(mx, 1e-10f)) * kLog10Of2 : mx) - 8.0f;
  auto fin = [&](float x) {
    if (raw_power) x = __log2f(fmaxf(x, 1e-10f)) * kLog10Of2;
    return (fmaxf(x, floor_v) + 4.0f) * 0.25f;
  };
  const int Fv = n_valid ? min(F, *n_valid / kHop) : F;
  float* o = out + int64_t(b) * per_utt;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (vec) {
    float4* o4 = reinterpret_cast<float4*>(o);
    const int64_t n4 = per_utt / 4;
    for (; i < n4; i += stride) {
      float4 v = o4[i];
      const int f = int((4 * i) % F);  // F % 4 == 0 on this path: the four elements share a row
      v.x = f + 0 < Fv ? fin(v.x) : 0.0f;
      v.y = f + 1 < Fv ? fin(v.y) : 0.0f;
      v.z = f + 2 < Fv ? fin(v.z) : 0.0f;
      v.w = f + 3 < Fv ? fin(v.w) : 0.0f;
      o4[i] = v;
    }
  } else {
    for (; i < per_utt; i += stride) o[i] = int(i % F) < Fv ? fin(o[i]) : 0.0f;
  }
}

// One CTA; thread m owns filter row m: first/last non-zero bin, then a serial prefix over rows.
__global__ void logmel_pack_filters_kernel(const float* __restrict__ fb, int n_mels, unsigned char* __restrict__ packed) {
  __shared__ int s_start[kMaxMels], s_count[kMaxMels], s_off[kMaxMels];
  PackedHeader* hdr = reinterpret_cast<PackedHeader*>(packed);
  int* g_start = reinterpret_cast<int*>(packed + sizeof(PackedHeader));
  int* g_count = g_start + n_mels;
  int* g_off = g_count + n_mels;
  float* g_w = reinterpret_cast<float*>(g_off + n_mels);
  const int m = threadIdx.x;
  if (m < n_mels) {
    int lo = kNfreq, hi = -1;
    for (int k = 0; k < kNfreq; ++k) {
      if (fb[m * kNfreq + k] != 0.0f) {
        lo = min(lo, k);
        hi = max(hi, k);
      }
    }
    s_start[m] = (hi < 0) ? 0 : lo;
    s_count[m] = (hi < 0) ? 0 : hi - lo + 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < n_mels; ++i) {
      s_off[i] = acc;
      acc += s_count[i];
    }
    hdr->n_mels = n_mels;
    hdr->total = acc;
    hdr->pad0 = hdr->pad1 = 0;
  }
  __syncthreads();
  if (m < n_mels) {
    g_start[m] = s_start[m];
    g_count[m] = s_count[m];
    g_off[m] = s_off[m];
    for (int i = 0; i < s_count[m]; ++i) g_w[s_off[m] + i] = fb[m * kNfreq + s_start[m] + i];
  }
}

}  // namespace

// shared with logmel_tc.cu
int logmel_launch_normalise(float* out, const uint32_t* maxkey, int64_t B, int64_t per_utt, int64_t F, const int32_t* n_valid,
                            int raw_power, cudaStream_t s) {
  const int vec = (F % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const int64_t work = vec ? per_utt / 4 : per_utt;
  unsigned gx = unsigned(std::min<int64_t>((work + 255) / 256, 1024));
  if (gx == 0) gx = 1;
  logmel_normalise_kernel<<<dim3(gx, unsigned(B)), 256, 0, s>>>(out, maxkey, per_utt, vec, int(F), n_valid, raw_power);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}
}  // namespace aga

using namespace aga;

extern "C" int aga_logmel_packed_filter_bytes(int n_mels, size_t* bytes) {
  if (!bytes || n_mels <= 0 || n_mels > kMaxMels) return AGA_ERR_INVALID_ARGUMENT;
  *bytes = sizeof(PackedHeader) + size_t(n_mels) * (3 * sizeof(int) + kNfreq * sizeof(float));
  return AGA_OK;
}

extern "C" int aga_logmel_pack_filters(const float* melfb, int n_mels, void* packed, size_t packed_bytes, void* stream) {
  size_t need = 0;
  int st = aga_logmel_packed_filter_bytes(n_mels, &need);
  if (st != AGA_OK) return st;
  if (!melfb || !packed) return AGA_ERR_INVALID_ARGUMENT;
  if (packed_bytes < need) return AGA_ERR_WORKSPACE_TOO_SMALL;
  logmel_pack_filters_kernel<<<1, kMaxMels, 0, static_cast<cudaStream_t>(stream)>>>(
      melfb, n_mels, static_cast<unsigned char*>(packed));
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

extern "C" int aga_logmel_workspace_bytes(int64_t B, int64_t N, int n_mels, size_t* bytes) {
  if (!bytes || B <= 0 || N <= kNfft / 2 || n_mels <= 0 || n_mels > kMaxMels) return AGA_ERR_INVALID_ARGUMENT;
  *bytes = align_up(size_t(B) * sizeof(uint32_t), 256);
  return AGA_OK;
}

extern "C" int aga_logmel_fwd(const float* audio, int64_t B, int64_t N, int64_t ld, const void* packed_filters,
                              int n_mels, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  size_t need = 0;
  int st = aga_logmel_workspace_bytes(B, N, n_mels, &need);
  if (st != AGA_OK) return st;
  if (!audio || !packed_filters || !out || !workspace || ld < N) return AGA_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < need) return AGA_ERR_WORKSPACE_TOO_SMALL;
  if (B > 65535) return AGA_ERR_UNSUPPORTED;
  const int64_t F = N / kHop;  // 1 + N/160 frames, the last one dropped (whisper_encoder.py:117)
  if (F <= 0) return AGA_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint32_t* maxkey = static_cast<uint32_t*>(workspace);
  AGA_CUDA_TRY(cudaMemsetAsync(maxkey, 0, size_t(B) * sizeof(uint32_t), s));
  AGA_CUDA_TRY(cudaFuncSetAttribute(logmel_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes)));
  const int64_t n_fblocks = (F + kFramesPerCta - 1) / kFramesPerCta;
  const int64_t n_blocks = n_fblocks * B;
  if (n_blocks > INT32_MAX) return AGA_ERR_UNSUPPORTED;
  static const int n_sm = []() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
  }();
  const unsigned grid = unsigned(std::min<int64_t>(n_blocks, 2 * int64_t(n_sm)));  // persistent: two resident CTAs per SM
  logmel_frames_kernel<<<grid, kThreads, kSmemBytes, s>>>(
      audio, N, ld, int(F), int(n_fblocks), int(n_blocks), static_cast<const unsigned char*>(packed_filters), n_mels, out,
      maxkey);
  AGA_AFTER_LAUNCH();
  return logmel_launch_normalise(out, maxkey, B, int64_t(n_mels) * F, F, nullptr, /*raw_power=*/0, s);
}
