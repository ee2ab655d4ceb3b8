// Whisper log-mel frontend on the 5th-generation tensor cores (sm_100a): the 400-point real DFT of every frame as
// split-precision fp16 GEMMs on tcgen05 with fp32 TMEM accumulators.
//
// Replaces OpenAIWhisperEncoder.log_mel_spectrogram (espnet2/asr/encoder/whisper_encoder.py:105-135) like logmel.cu
// does, for BANDED filterbanks (every bin feeds at most two adjacent filters, filter index monotone in the bin: the
// Slaney triangles of whisper/audio.py:92-107 and their 128-bin sibling).  logmel.cu remains the general path.
//
// Why tensor cores: the CUDA-core transform of logmel.cu issues ~17 k thread-instructions per frame and is fp32-issue
// bound at 0.13 of the HBM roofline.  Folding the windowed frame turns the DFT into small dense products:
//     x_n = frame sample n (n = 0..399), w_n = periodic Hann (w_0 = 0, w_n = w_{400-n}), th = 2 pi n k / 400
//     Re X_k =  sum_{n=1..199} w_n (x_n + x_{400-n}) cos th + w_200 x_200 cos(pi k)
//     Im X_k = -sum_{n=1..199} w_n (x_n - x_{400-n}) sin th
//   and, splitting n by parity, bins k and 200-k share their partial sums (cos th' = (-1)^n cos th, sin th' = -(-1)^n sin th):
//     Ce_k = sum_{n even} w e_n cos th   Co_k = sum_{n odd} w e_n cos th   Se_k, So_k alike with o_n = x_n - x_{400-n}, sin
//     |X_k|^2 = (Ce+Co)^2 + (Se+So)^2        |X_{200-k}|^2 = (Ce-Co)^2 + (Se-So)^2          k = 0..100
//   i.e. FOUR GEMMs (frames x 100) @ (100 x 101) per frame tile instead of two (frames x 400) @ (400 x 201): 8x fewer MACs
//   than the plain DFT-as-GEMM.  The Hann window lives in the tables.
// Precision: operands are fp16 PAIRS, v = hi + lo with hi = fp16(v), lo = fp16(v - hi) (22 significant bits; the
//   samples of a tile are pre-scaled by a power of two so that hi / lo stay in fp16's normal range), three MMAs per
//   product (hi*hi + hi*lo + lo*hi; the dropped lo*lo term is 2^-22 relative), fp32 accumulation in TMEM: the result
//   is as accurate as an fp32 FFT (tests/test_gpu_logmel.py holds it to the same 1e-4 gate as logmel.cu).
//
// Kernel (persistent, one CTA per SM, 18 warps):
//   tile = 128 consecutive frames of one utterance; its 20 720 samples are staged ONCE in shared memory (128-bit HBM
//   loads, reflect padding at the utterance ends, tile maximum -> power-of-two scale);
//   8 producer warps (lane = frame) build the A operands k-step by k-step straight in the UMMA canonical
//   (no-swizzle, K-major) shared-memory layout: each (frame, k-step) task reads 2 x 32 consecutive samples, forms
//   16 values of each of the four folded sequences, splits them and writes 16 x 16 bytes;
//   1 loader warp streams the matching table slices (cp.async.bulk, L2-resident, 28 KB per k-step);
//   1 warp issues tcgen05.mma (M128 N112 K16, 12 per k-step) into four TMEM accumulators (Ce | Co | Se | So);
//   8 epilogue warps (two per TMEM lane quarter: bins 0..100 upwards and 200..101 downwards) read the accumulators, form
//   the power spectrum and stream it through the banded mel projection with two running sums per thread; a completed
//   filter leaves as its raw power (128-byte coalesced stores), one atomicMax per warp keeps the per-utterance maximum.
//   logmel_normalise_kernel (logmel.cu) then applies log10(clamp), max(x, m - 8), (x + 4) / 4 in place (the tile is still
//   in L2).
// Roofline: HBM — N*4 bytes read + n_mels*(N/160)*4 written per utterance (2.88 MB per 30 s at 80 bins).
#include "aga_common.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include <cuda_fp16.h>

// Debug timeline (compile with -DAGA_TIMELINE): CTA 0 records (tag, clock64()) per warp role into a global buffer set with
// aga_debug_set_logmel_timeline(); each role owns a row of 4096 int64 slots (tools/timeline_logmel.py).
#ifdef AGA_TIMELINE
__device__ long long* g_lm_timeline = nullptr;
#define LTL_DECL(role) long long* tl_ptr = (g_lm_timeline && (role) >= 0 && blockIdx.x == 0) ? g_lm_timeline + (role) * 4096 : nullptr; int tl_n = 0
#define LTL(tag) do { if (tl_ptr && tl_n < 2046) { tl_ptr[2 * tl_n] = (tag); tl_ptr[2 * tl_n + 1] = clock64(); ++tl_n; tl_ptr[2 * tl_n] = -1; } } while (0)
extern "C" __attribute__((visibility("default"))) int aga_debug_set_logmel_timeline(long long* p) {
  return cudaMemcpyToSymbol(g_lm_timeline, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}
#else
#define LTL_DECL(role) do { } while (0)
#define LTL(tag) do { } while (0)
#endif

namespace aga {
// defined in logmel.cu
int logmel_launch_normalise(float* out, const uint32_t* maxkey, int64_t B, int64_t per_utt, int64_t F, const int32_t* n_valid,
                            int raw_power, cudaStream_t s);

namespace {
using namespace ptx;

constexpr int kHop = 160;
constexpr int kNfft = 400;
constexpr int kNfreq = 201;
constexpr int kTileFrames = 128;
constexpr int kChunk = (kTileFrames - 1) * kHop + kNfft;  // 20 720 samples per tile
constexpr int kSampPad = 4;                                // floats of padding per hop: lane = frame reads are conflict-free
__host__ __device__ constexpr int pidx(int i) { return i + kSampPad * (i / kHop); }
constexpr int kFrameStride = kHop + kSampPad;              // 164
constexpr int kSampFloats = (pidx(kChunk - 1) + 1 + 3) / 4 * 4;
constexpr int kKSteps = 7;     // K = 100 padded to 112 = 7 x 16
constexpr int kN = 112;        // bins 0..100 padded to a multiple of 16
constexpr int kBins = 101;
constexpr int kASub = kTileFrames * 32;  // one A sub-chunk: 128 rows x 16 fp16 (4096 B)
constexpr int kTSub = kN * 32;           // one table sub-chunk: 112 rows x 16 fp16 (3584 B)
constexpr int kStageA = 4 * kASub;       // the LO halves of the 4 sequences (the HI halves live in TMEM)
constexpr int kStageT = 8 * kTSub;       // 4 tables x (hi, lo)
constexpr int kStageBytes = kStageA + kStageT;
constexpr int kStages = 2;
constexpr int kTableBytes = kKSteps * kStageT;  // 200 704 B of DFT tables in global memory
constexpr int kAccStride = kN;                  // TMEM columns between the four accumulators: [0, 448)
constexpr int kColAhi = 4 * kN;                 // [448, 512): HI halves of the A operands, 2 ring slots x 4 sequences x 8 columns
constexpr int kTmemCols = 512;

// Warp roles.  The SM sub-partition schedulers prefer the HIGHEST warp id among the ready warps: the epilogue (which holds
// the accumulators and therefore the tensor core) gets the top ids, the producers the middle, the two single-thread
// issuers the bottom; every long wait backs off with nanosleep instead of polling.  (With the epilogue on warps 0-7 and
// polling waits above it, it ran at one instruction per ~10 cycles.)  Warps 2, 3 only keep the role blocks aligned to
// the TMEM lane quarters (warp % 4) that tcgen05.ld / .st impose.
constexpr int kEpiWarps = 8, kPrepWarps = 8;
constexpr int kWarpMma = 0, kWarpLoad = 1, kWarpPrep0 = 4, kWarpEpi0 = kWarpPrep0 + kPrepWarps;
constexpr int kThreads = (kWarpEpi0 + kEpiWarps) * 32;  // 640
constexpr int kPrepThreads = kPrepWarps * 32;
constexpr int kStageVecs = (kChunk / 4 + kPrepThreads - 1) / kPrepThreads;  // float4 per producer thread per tile (21)
constexpr int kMaxMels = 256;

// named barriers
constexpr int kBarPrep = 1;       // the 256 producer threads
constexpr int kBarQuarter0 = 2;   // +q: the two epilogue warps of TMEM lane quarter q (64 threads)

struct TcHeader {
  int32_t magic, n_mels;
  int32_t L100, L101;   // lower filter index of bins 100 / 101: where the ascending and the descending stream meet
  int32_t first_id[2];  // filter the "older" running sum of stream 0 (ascending) / 1 (descending) starts on
  int32_t pad[2];
};
constexpr int32_t kMagic = 0x4c4d5443;  // "LMTC"
// Mel projection as two STREAMS per frame (one epilogue warp each): stream 0 walks bins 0..100 upwards, stream 1 walks bins
// 200..101 downwards; position p of either stream is TMEM column p of the accumulators.  A stream keeps two running sums
// (filters "older" and "newer"); before a position it hands over `rot` (0..2) times: the older filter is complete and is
// written out, newer becomes older.  Per position: float2 {weight into older, weight into newer}; per 16 positions two
// 16-bit masks (rot >= 1, rot >= 2).
// packed layout: TcHeader | float4 bin[201] = {wA, wB, L (int bits), 0} | float2 wt[2][112] | uint32 mask[2][7][2] | pad to 256 |
//                DFT tables (kTableBytes)
constexpr size_t kBinTableOff = sizeof(TcHeader);
constexpr size_t kStreamOff = kBinTableOff + kNfreq * 16;
constexpr size_t kStreamBytes = 2 * 112 * 8 + 2 * 7 * 2 * 4;
constexpr size_t kDftOff = (kStreamOff + kStreamBytes + 255) / 256 * 256;
constexpr size_t kPackedBytes = kDftOff + kTableBytes;

struct Smem {
  static constexpr size_t kSamp = 0;
  static constexpr size_t kStage = (size_t(kSampFloats) * 4 + 1023) / 1024 * 1024;
  static constexpr size_t kBin = kStage + size_t(kStages) * kStageBytes;  // stream tables (kStreamBytes)
  static constexpr size_t kPartial = kBin + 208 * 16;                     // float[2][128]
  static constexpr size_t kRed = kPartial + 2 * 128 * 4;                  // float[8] + float2 scale[2]
  static constexpr size_t kBars = kRed + 64;
  static constexpr size_t kTotal = kBars + 16 * 8 + 16;
};

// UMMA shared-memory descriptor, K-major, no swizzle ("interleave"): 8 rows x 16 bytes form a contiguous 128-byte core
// matrix; LBO = distance between the two core matrices of one K16 step, SBO = distance between 8-row groups.
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t(lbo >> 4) << 16;
  d |= uint64_t(sbo >> 4) << 32;
  d |= uint64_t(1) << 46;  // descriptor version (Blackwell)
  return d;                // layout type 0 = SWIZZLE_NONE
}
// kind::f16 instruction descriptor, fp16 x fp16 -> fp32, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

__device__ __forceinline__ float load_reflect(const float* __restrict__ row, int64_t g, int64_t N) {
  if (g < 0) g = -g;                    // reflect, no edge repeat: x_pad[199 - i] = x[i + 1]
  if (g >= N) g = 2 * (N - 1) - g;
  return (g >= 0 && g < N) ? __ldg(row + g) : 0.0f;
}

__device__ __forceinline__ void st_global_pred(float* ptr, float v, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.global.f32 [%0], %1;\n\t}" ::"l"(ptr), "f"(v), "r"(int(pred)) : "memory");
}

// v -> (hi, lo) fp16 pairs of two values, packed
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

struct Params {
  const float* audio;
  int64_t N, ld;
  int F, tiles_per_utt, n_tiles, n_mels;
  const unsigned char* packed;
  float* out;
  uint32_t* maxkey;
  const int32_t* n_valid;  // optional device scalar: true common length of the (zero-padded) batch
};

__global__ void __launch_bounds__(kThreads, 1) logmel_tc_kernel(const Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  float* s_samp = reinterpret_cast<float*>(smem + Smem::kSamp);
  const uint32_t stage_u32 = smem_u32(smem + Smem::kStage);
  float* s_partial = reinterpret_cast<float*>(smem + Smem::kPartial);
  float* s_red = reinterpret_cast<float*>(smem + Smem::kRed);
  float2* s_scale = reinterpret_cast<float2*>(smem + Smem::kRed + 32);  // [tile parity] = {S, 1 / S^2}
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::kBars);
  uint64_t* bar_full = bars;        // [2]  A written (8 producer warps) + table slice landed
  uint64_t* bar_empty = bars + 2;   // [2]  the stage's MMAs have completed
  uint64_t* bar_acc_full = bars + 4;   // accumulators of the tile complete
  uint64_t* bar_acc_empty = bars + 5;  // the 8 epilogue warps have read them
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TcHeader* hdr = reinterpret_cast<const TcHeader*>(p.packed);
  const unsigned char* g_tables = p.packed + kDftOff;

  // effective sample / frame counts (the batch may be zero-padded beyond a common true length)
  int64_t N = p.N;
  if (p.n_valid) N = min(int64_t(*p.n_valid), p.N);
  const int Fv = int(N / kHop);

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&bar_full[0], kPrepWarps + 1);
      mbar_init(&bar_full[1], kPrepWarps + 1);
      mbar_init(&bar_empty[0], 1);
      mbar_init(&bar_empty[1], 1);
      mbar_init(bar_acc_full, 1);
      mbar_init(bar_acc_empty, kEpiWarps);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(s_tmem, kTmemCols);
    tmem_relinquish();
  }
  {  // mel stream tables -> smem
    const uint32_t* g_st = reinterpret_cast<const uint32_t*>(p.packed + kStreamOff);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem + Smem::kBin);
    for (int i = tid; i < int(kStreamBytes / 4); i += kThreads) dst[i] = g_st[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  const int n_my_tiles = (p.n_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

  if (warp >= kWarpPrep0 && warp < kWarpEpi0) {
    // =========================================================== producers: staging + A operands
    const int ptid = tid - kWarpPrep0 * 32;
    const int quarter = warp & 3, khalf = (warp - kWarpPrep0) >> 2;  // this warp builds K columns [8 khalf, 8 khalf + 8) of a step
    const int r = quarter * 32 + lane;  // frame (row of the MMA) this thread owns
    const float* sp = s_samp + r * kFrameStride;
    // the thread's 16-byte slot inside a 4 KB sub-chunk: row r, K half khalf
    const uint32_t row_off = uint32_t(r >> 3) * 256u + uint32_t(r & 7) * 16u + uint32_t(khalf) * 128u;
    const uint32_t t_ahi = tmem + (uint32_t(quarter * 32) << 16) + kColAhi + uint32_t(khalf) * 4;  // this warp's lanes / K half
    LTL_DECL((lane == 0 && quarter == 0) ? khalf : -1);
    for (int ti = 0; ti < n_my_tiles; ++ti) {
      LTL(100);
      const int t = blockIdx.x + ti * gridDim.x;
      const int b = t / p.tiles_per_utt, f0 = (t - b * p.tiles_per_utt) * kTileFrames;
      const float* arow = p.audio + int64_t(b) * p.ld;
      // ---- stage the tile's samples: HBM -> registers (+ running max) -> scaled -> smem.  Interior float4s take the
      //      unrolled 128-bit path; the few that touch an utterance end (reflect padding, zero fill) or an unaligned row
      //      are recomputed by ROLLED loops (kept out of the unrolled code: instruction-cache footprint).
      float4 v[kStageVecs];
      const int64_t g0 = int64_t(f0) * kHop - kNfft / 2;  // multiple of 8 samples
      const bool vec_ok = ((reinterpret_cast<uintptr_t>(arow) & 15) == 0);
      const bool edge_tile = !vec_ok || g0 < 0 || g0 + kChunk > N;
      float amax = 0.0f;
#pragma unroll
      for (int u = 0; u < kStageVecs; ++u) {
        const int c = ptid + u * kPrepThreads;
        const int64_t g = g0 + 4 * int64_t(c);
        const bool interior = c < kChunk / 4 && vec_ok && g >= 0 && g + 3 < N;
        v[u] = interior ? __ldg(reinterpret_cast<const float4*>(arow + (interior ? g : 0))) : make_float4(0.f, 0.f, 0.f, 0.f);
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w))));
      }
      if (edge_tile) {
#pragma unroll 1
        for (int c = ptid; c < kChunk / 4; c += kPrepThreads) {
          const int64_t g = g0 + 4 * int64_t(c);
          if (vec_ok && g >= 0 && g + 3 < N) continue;
#pragma unroll 1
          for (int e2 = 0; e2 < 4; ++e2) amax = fmaxf(amax, fabsf(load_reflect(arow, g + e2, N)));
        }
      }
      amax = warp_max(amax);
      LTL(101);
      // previous tile: every producer has finished READING the sample buffer (its last k-step is written)
      named_bar_sync(kBarPrep, kPrepThreads);
      LTL(102);
      if (lane == 0) s_red[warp - kWarpPrep0] = amax;
      named_bar_sync(kBarPrep, kPrepThreads);
      float m = s_red[0];
#pragma unroll
      for (int i = 1; i < kPrepWarps; ++i) m = fmaxf(m, s_red[i]);
      // power-of-two scale: the tile's largest |sample| lands in [2^12, 2^13), folded sums stay below fp16's 65504
      int e = int((__float_as_uint(m) >> 23) & 0xff) - 127;
      e = (m > 0.0f) ? max(-50, min(e, 60)) : 12;
      const float S = __uint_as_float(uint32_t(127 + 12 - e) << 23);
      if (ptid == 0) s_scale[ti & 1] = make_float2(S, __uint_as_float(uint32_t(127 - 2 * (12 - e)) << 23));
#pragma unroll
      for (int u = 0; u < kStageVecs; ++u) {
        const int c = ptid + u * kPrepThreads;
        if (c < kChunk / 4) {
          float4 x = v[u];
          x.x *= S; x.y *= S; x.z *= S; x.w *= S;
          *reinterpret_cast<float4*>(s_samp + pidx(4 * c)) = x;
        }
      }
      if (edge_tile) {
#pragma unroll 1
        for (int c = ptid; c < kChunk / 4; c += kPrepThreads) {
          const int64_t g = g0 + 4 * int64_t(c);
          if (vec_ok && g >= 0 && g + 3 < N) continue;
          float* dst = s_samp + pidx(4 * c);
#pragma unroll 1
          for (int e2 = 0; e2 < 4; ++e2) dst[e2] = load_reflect(arow, g + e2, N) * S;
        }
      }
      __threadfence_block();
      named_bar_sync(kBarPrep, kPrepThreads);
      LTL(103);

      // ---- A operands, k-step by k-step (ring slot = step parity); all 8 producer warps work on every step, 4 on each K half
      for (int j = 0; j < kKSteps; ++j) {
        const int it = ti * kKSteps + j, slot_i = it & 1, use = it >> 1;
        LTL(110);
        if (use > 0) mbar_wait_sleep(&bar_empty[slot_i], (use - 1) & 1, 40);
        tc_fence_after();
        LTL(111);
        // K columns kk = 8 khalf + c (c = 0..7) need n = 32 j + t, t = 16 khalf + t', t' = 1..16, and x_{400-n}:
        //   fw[i] = x[32 j + 16 khalf + i], i = 0..19;   bw[i] = x[368 - 32 j + 16 (1 - khalf) + i], i = 0..15;   x_{400-n} = bw[16 - t']
        float fw[20], bw[16];
        const int nf = 32 * j + 16 * khalf, nb = 368 - 32 * j + 16 * (1 - khalf);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const int n = nf + 4 * q;
          const float4 x = *reinterpret_cast<const float4*>(sp + n + kSampPad * (n / kHop));
          fw[4 * q] = x.x; fw[4 * q + 1] = x.y; fw[4 * q + 2] = x.z; fw[4 * q + 3] = x.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = nb + 4 * q;
          const float4 x = *reinterpret_cast<const float4*>(sp + n + kSampPad * (n / kHop));
          bw[4 * q] = x.x; bw[4 * q + 1] = x.y; bw[4 * q + 2] = x.z; bw[4 * q + 3] = x.w;
        }
        // sequences: 0 = e even n (t' = 2+2c), 1 = e odd n (t' = 1+2c), 2 = o even, 3 = (-1)^kk o odd  ((-1)^kk = (-1)^c)
        const int c_valid = ((j == kKSteps - 1) ? 4 : 16) - 8 * khalf;  // K index 16 j + kk < 100
        const uint32_t slot = stage_u32 + uint32_t(slot_i) * kStageBytes + row_off;
#pragma unroll
        for (int seq = 0; seq < 4; ++seq) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int c2 = 0; c2 < 4; ++c2) {  // c = 2 c2, 2 c2 + 1
            float val[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c = 2 * c2 + h;
              const int tt = ((seq & 1) ? 1 : 2) + 2 * c;
              const float a = fw[tt], bb = bw[16 - tt];
              float x;
              if (seq < 2) x = a + bb;
              else if (seq == 2) x = a - bb;
              else x = (c & 1) ? (bb - a) : (a - bb);
              val[h] = (c < c_valid) ? x : 0.0f;
            }
            split2(val[0], val[1], hi[c2], lo[c2]);
          }
          // HI half: 8 fp16 = 4 TMEM columns of this row (the A operand of a TS MMA: lane = row, 2 K elements per column);
          // LO half: the row's 16-byte slot of the sequence's shared-memory sub-chunk
          tmem_st4(t_ahi + uint32_t(slot_i * 32 + seq * 8), hi);
          sts128(slot + uint32_t(seq) * kASub, lo[0], lo[1], lo[2], lo[3]);
        }
        tmem_wait_st();
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_full[slot_i]);
        LTL(112);
      }
    }
  } else if (warp == kWarpLoad) {
    // =========================================================== table loader
    if (elect_one()) {
      LTL_DECL(3);
      const int total = n_my_tiles * kKSteps;
      for (int it = 0; it < total; ++it) {
        const int s = it & 1, use = it >> 1, j = it % kKSteps;
        LTL(400);
        if (use > 0) mbar_wait_sleep(&bar_empty[s], (use - 1) & 1, 100);
        LTL(401);
        mbar_arrive_expect_tx(&bar_full[s], kStageT);
        bulk_load(smem + Smem::kStage + size_t(s) * kStageBytes + kStageA, g_tables + size_t(j) * kStageT, kStageT, &bar_full[s]);
      }
    }
  } else if (warp == kWarpMma) {
    // =========================================================== MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(kTileFrames, kN);
    LTL_DECL(lane == 0 ? 2 : -1);
    for (int ti = 0; ti < n_my_tiles; ++ti) {
      LTL(200);
      if (ti > 0) {
        mbar_wait_sleep(bar_acc_empty, (ti - 1) & 1, 100);
        tc_fence_after();
      }
      LTL(201);
      for (int j = 0; j < kKSteps; ++j) {
        const int it = ti * kKSteps + j, s = it & 1, use = it >> 1;
        LTL(210);
        mbar_wait_sleep(&bar_full[s], use & 1, 20);
        LTL(211);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a0 = stage_u32 + uint32_t(s) * kStageBytes, t0 = a0 + kStageA;
#pragma unroll
          for (int seq = 0; seq < 4; ++seq) {
            const uint32_t a_hi = tmem + kColAhi + uint32_t(s * 32 + seq * 8);  // TMEM: 128 lanes x 8 columns
            const uint64_t a_lo = make_desc_nosw(a0 + uint32_t(seq) * kASub, 128, 256);
            const uint64_t t_hi = make_desc_nosw(t0 + uint32_t(2 * seq) * kTSub, 128, 256);
            const uint64_t t_lo = make_desc_nosw(t0 + uint32_t(2 * seq + 1) * kTSub, 128, 256);
            const uint32_t d = tmem + uint32_t(seq) * kAccStride;
            mma_ts(d, a_hi, t_hi, idesc, j > 0 ? 1u : 0u);
            mma_ts(d, a_hi, t_lo, idesc, 1u);
            mma_ss(d, a_lo, t_hi, idesc, 1u);
          }
          tc_commit(&bar_empty[s]);
          if (j == kKSteps - 1) tc_commit(bar_acc_full);
        }
        __syncwarp();
        LTL(212);
      }
    }
  } else if (warp >= kWarpEpi0) {
    // =========================================================== epilogue: power spectrum -> banded mel -> log10
    // Two streams per TMEM lane quarter (see TcHeader): warp q walks bins 0..100 upwards, warp 4 + q bins 200..101
    // downwards, with the SAME code (sign of the Co / So terms, tables and output direction are data): a rolled loop over
    // 7 blocks of 16 accumulator columns whose body is unrolled over the 16 register-resident columns.  The hand-over of
    // the running sums is a warp-uniform branch on two 16-bit masks per block.  (Fully unrolled over the filterbank the
    // epilogue was 50-200 KB of straight-line code and instruction-fetch bound — the MMAs wait for the accumulators.)
    const int quarter = warp & 3, role = (warp - kWarpEpi0) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t tq = tmem + (uint32_t(quarter * 32) << 16);
    const int n_mels = p.n_mels;
    const int64_t F = p.F;
    const float sgn = role ? -1.0f : 1.0f;
    const float2* wt = reinterpret_cast<const float2*>(smem + Smem::kBin) + role * 112;
    const uint32_t* mk = reinterpret_cast<const uint32_t*>(smem + Smem::kBin + 2 * 112 * 8) + role * 14;
    const int first_id = hdr->first_id[role], step = role ? -1 : 1;
    const int L100 = hdr->L100, L101 = hdr->L101;
    LTL_DECL((lane == 0 && quarter == 0) ? 4 + role : -1);
    for (int ti = 0; ti < n_my_tiles; ++ti) {
      LTL(300);
      const int t = blockIdx.x + ti * gridDim.x;
      const int b = t / p.tiles_per_utt, f = (t - b * p.tiles_per_utt) * kTileFrames + r;
      const bool live = f < Fv;
      float* obase = p.out + int64_t(b) * n_mels * F + f;
      mbar_wait_sleep(bar_acc_full, ti & 1, 200);
      LTL(301);
      tc_fence_after();
      const float inv_s2 = s_scale[ti & 1].y;
      float vmax = 0.f;  // of the RAW mel power (>= 0): log10 is monotone, logmel_normalise_kernel takes it once per element
      // a completed filter leaves as its raw (un-logged) power: one multiply, one predicated store
      auto emit = [&](int m, float acc) {  // m is warp-uniform
        const float v = acc * inv_s2;
        if (live && m >= 0 && m < n_mels) {
          obase[int64_t(m) * F] = v;
          vmax = fmaxf(vmax, v == v ? v : INFINITY);
        }
      };
      float accA = 0.f, accB = 0.f;  // running sums of the older / newer filter
      int m_id = first_id;           // the older filter
      const int64_t step_f = int64_t(step) * F;
      float* optr = obase + int64_t(m_id) * F;  // where the older filter's value of this frame goes (dereferenced only when valid)
#pragma unroll 1
      for (int blk = 0; blk < 7; ++blk) {
        float pw[16];
        {
          uint32_t ce[16], co[16], se[16], so[16];
          tmem_ld16(tq + 0 * kAccStride + blk * 16, ce);
          tmem_ld16(tq + 1 * kAccStride + blk * 16, co);
          tmem_ld16(tq + 2 * kAccStride + blk * 16, se);
          LTL(310 + blk);
          tmem_ld16(tq + 3 * kAccStride + blk * 16, so);
          tmem_wait_ld();
          LTL(320 + blk);
          if (blk == 6) {  // last TMEM read of the tile: the accumulators are free for the next tile's MMAs
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty);
            LTL(302);
          }
          // phase 1 (no branches, 16 independent chains): the power spectrum of the block's 16 bins
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float re = fmaf(sgn, __uint_as_float(co[i]), __uint_as_float(ce[i]));
            const float im = fmaf(sgn, __uint_as_float(so[i]), __uint_as_float(se[i]));
            pw[i] = fmaf(re, re, im * im);
          }
        }
        LTL(330 + blk);
        const uint32_t m1 = mk[2 * blk], m2 = mk[2 * blk + 1];
        float2 wv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) wv[i] = wt[blk * 16 + i];
        // phase 2: the running sums.  The hand-over (where the block's masks say so) is PREDICATED, not branched: a
        // completed filter leaves through a predicated store and the sums rotate through selects.  (With a branch per
        // bin the warp spent its time on instruction-fetch bubbles behind the reconvergence points: 15 k cycles per tile.)
        auto hand_over = [&](bool rot) {
          const bool pe = rot && live && unsigned(m_id) < unsigned(n_mels);
          const float v = accA * inv_s2;
          st_global_pred(optr, v, pe);
          // (a NaN power — non-finite samples — must not vanish in fmaxf: it raises the utterance maximum to +inf, which
          //  makes the whole utterance non-finite downstream, as the reference's NaN-propagating max does)
          vmax = fmaxf(vmax, pe ? (v == v ? v : INFINITY) : 0.f);
          accA = rot ? accB : accA;
          accB = rot ? 0.f : accB;
          m_id += rot ? step : 0;
          optr += rot ? step_f : 0;
        };
        if (m2 == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            hand_over((m1 >> i) & 1);
            accA = fmaf(wv[i].x, pw[i], accA);
            accB = fmaf(wv[i].y, pw[i], accB);
          }
        } else {  // a bin that completes two filters at once (once per stock filterbank)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            hand_over((m1 >> i) & 1);
            hand_over((m2 >> i) & 1);
            accA = fmaf(wv[i].x, pw[i], accA);
            accB = fmaf(wv[i].y, pw[i], accB);
          }
        }
      }
      // the streams meet between bins 100 and 101.  ascending: accA <-> filter L100, accB <-> L100 + 1;
      // descending: accA <-> filter L101 + 1, accB <-> L101.  The ascending warp writes the (at most four) filters out.
      if (role == 1) {
        s_partial[r] = accB;
        s_partial[128 + r] = accA;
        __threadfence_block();
        named_bar_arrive(kBarQuarter0 + quarter, 64);
      } else {
        named_bar_sync(kBarQuarter0 + quarter, 64);
        const float u0 = s_partial[r], u1 = s_partial[128 + r];  // filters L101, L101 + 1 as far as bins >= 101 go
        const int d = L101 - L100;                                // 0, 1 or 2 (checked when the tables are built)
        if (d == 0) {
          emit(L100, accA + u0);
          emit(L100 + 1, accB + u1);
        } else if (d == 1) {
          emit(L100, accA);
          emit(L100 + 1, accB + u0);
          emit(L100 + 2, u1);
        } else {
          emit(L100, accA);
          emit(L100 + 1, accB);
          emit(L100 + 2, u0);
          emit(L100 + 3, u1);
        }
        // the descending warp may overwrite its partial sums for the next tile only after they have been read
        named_bar_arrive(kBarQuarter0 + 4 + quarter, 64);
      }
      vmax = warp_max(vmax);
      if (lane == 0) atomicMax(p.maxkey + b, float_to_key(vmax));
      if (role == 1) named_bar_sync(kBarQuarter0 + 4 + quarter, 64);
      LTL(303);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ host side
// Band analysis of an (n_mels x 201) filterbank: L[k] = index of the LOWER of the (at most two, adjacent) filters bin k
// feeds, non-decreasing in k; wA / wB = its weights in filters L[k] / L[k] + 1.  Returns false for any other structure.
bool analyse_bands(const float* fb, int n_mels, std::vector<float>& wA, std::vector<float>& wB, std::vector<int>& L) {
  wA.assign(kNfreq, 0.f);
  wB.assign(kNfreq, 0.f);
  L.assign(kNfreq, -1);
  int prev = -1;
  for (int k = 0; k < kNfreq; ++k) {
    int first = -1, last = -1, cnt = 0;
    for (int m = 0; m < n_mels; ++m) {
      if (fb[size_t(m) * kNfreq + k] != 0.0f) {
        if (first < 0) first = m;
        last = m;
        ++cnt;
      }
    }
    if (cnt == 0) {
      L[k] = prev;
      continue;
    }
    if (cnt > 2 || last - first != cnt - 1) return false;
    int l;
    if (cnt == 2) {
      l = first;
    } else {
      l = std::max(prev, first - 1);
      if (l != first && l != first - 1) return false;
    }
    if (l < prev) return false;
    L[k] = l;
    if (l >= 0 && l < n_mels) wA[k] = fb[size_t(l) * kNfreq + k];
    if (l + 1 < n_mels) wB[k] = fb[size_t(l + 1) * kNfreq + k];
    if (cnt == 2 && l + 1 != last) return false;
    prev = l;
  }
  return true;
}

// Stream tables of the epilogue (see TcHeader).  Returns false when a hand-over count exceeds 2 or the two streams do not
// meet within two filters of each other (never the case for triangular banks).
bool build_streams(const std::vector<float>& wA, const std::vector<float>& wB, const std::vector<int>& L, int n_mels,
                   float* wt /*[2][112][2]*/, uint32_t* mask /*[2][7][2]*/, int32_t* first_id /*[2]*/) {
  std::memset(wt, 0, 2 * 112 * 2 * sizeof(float));
  std::memset(mask, 0, 2 * 7 * 2 * sizeof(uint32_t));
  const int d = L[101] - L[100];
  if (d < 0 || d > 2) return false;
  first_id[0] = L[0];        // ascending: older = filter L[bin] (weight wA), newer = L[bin] + 1 (weight wB)
  first_id[1] = L[200] + 1;  // descending: older = filter L[bin] + 1 (weight wB), newer = L[bin] (weight wA)
  for (int pos = 0; pos <= 100; ++pos) {
    const int bin = pos, rot = pos == 0 ? 0 : L[bin] - L[bin - 1];
    if (rot < 0 || rot > 2) return false;
    wt[(0 * 112 + pos) * 2 + 0] = wA[bin];
    wt[(0 * 112 + pos) * 2 + 1] = wB[bin];
    if (rot >= 1) mask[(0 * 7 + pos / 16) * 2 + 0] |= 1u << (pos % 16);
    if (rot >= 2) mask[(0 * 7 + pos / 16) * 2 + 1] |= 1u << (pos % 16);
  }
  for (int pos = 0; pos <= 99; ++pos) {
    const int bin = 200 - pos, rot = pos == 0 ? 0 : L[bin + 1] - L[bin];
    if (rot < 0 || rot > 2) return false;
    wt[(1 * 112 + pos) * 2 + 0] = wB[bin];
    wt[(1 * 112 + pos) * 2 + 1] = wA[bin];
    if (rot >= 1) mask[(1 * 7 + pos / 16) * 2 + 0] |= 1u << (pos % 16);
    if (rot >= 2) mask[(1 * 7 + pos / 16) * 2 + 1] |= 1u << (pos % 16);
  }
  (void)n_mels;
  return true;
}

// One fp16 (hi, lo) pair of a table entry.
inline void split_half(double v, __half& hi, __half& lo) {
  hi = __float2half_rn(float(v));
  lo = __float2half_rn(float(v - double(__half2float(hi))));
}

// DFT tables in the exact shared-memory image of a pipeline stage: [k-step][table 0..3][hi, lo][112 rows x 16 fp16] with
// row n at (n / 8) * 256 + (n % 8) * 16 and the two K halves 128 bytes apart.  Tables (row = bin k, column = K index):
//   0: w_n cos(2 pi n k / 400), n = 2 (K + 1)  (n = 200 carries 1/2: the folded operand holds 2 x_200)
//   1: w_n cos(2 pi n k / 400), n = 2 K + 1
//   2: w_n sin(2 pi n k / 400), n = 2 (K + 1)
//   3: (-1)^K w_n sin(2 pi n k / 400), n = 2 K + 1     (the sign is folded into the operand as well: (-1)^K o_n)
void build_dft_tables(unsigned char* dst) {
  std::memset(dst, 0, kTableBytes);
  const double two_pi = 6.283185307179586476925286766559;
  for (int j = 0; j < kKSteps; ++j) {
    for (int tb = 0; tb < 4; ++tb) {
      __half* hi = reinterpret_cast<__half*>(dst + size_t(j) * kStageT + size_t(2 * tb) * kTSub);
      __half* lo = reinterpret_cast<__half*>(dst + size_t(j) * kStageT + size_t(2 * tb + 1) * kTSub);
      for (int k = 0; k < kBins; ++k) {
        for (int kk = 0; kk < 16; ++kk) {
          const int K = 16 * j + kk;
          if (K >= 100) continue;
          const int n = (tb == 0 || tb == 2) ? 2 * (K + 1) : 2 * K + 1;
          const double w = 0.5 - 0.5 * std::cos(two_pi * n / 400.0);
          // exact argument reduction: n k mod 400
          const double ang = two_pi * double((int64_t(n) * k) % 400) / 400.0;
          double v;
          if (tb == 0) v = w * std::cos(ang) * (n == 200 ? 0.5 : 1.0);
          else if (tb == 1) v = w * std::cos(ang);
          else if (tb == 2) v = w * std::sin(ang);
          else v = ((K & 1) ? -1.0 : 1.0) * w * std::sin(ang);  // operand 3 carries the same (-1)^K: the signs cancel
          const size_t e = size_t(k / 8) * 128 + size_t(kk / 8) * 64 + size_t(k % 8) * 8 + size_t(kk % 8);  // in fp16 units
          split_half(v, hi[e], lo[e]);
        }
      }
    }
  }
}

}  // namespace
}  // namespace aga

using namespace aga;

extern "C" int aga_logmel_tc_packed_bytes(int n_mels, size_t* bytes) {
  if (!bytes || n_mels <= 0 || n_mels > kMaxMels) return AGA_ERR_INVALID_ARGUMENT;
  *bytes = kPackedBytes;
  return AGA_OK;
}

extern "C" int aga_logmel_filters_banded(const float* melfb_host, int n_mels) {
  if (!melfb_host || n_mels <= 0 || n_mels > kMaxMels) return 0;
  std::vector<float> wA, wB;
  std::vector<int> L;
  if (!analyse_bands(melfb_host, n_mels, wA, wB, L)) return 0;
  std::vector<float> wt(2 * 112 * 2);
  uint32_t mask[2 * 7 * 2];
  int32_t first_id[2];
  return build_streams(wA, wB, L, n_mels, wt.data(), mask, first_id) ? 1 : 0;
}

extern "C" int aga_logmel_tc_build_host(const float* melfb_host, int n_mels, void* packed_host, size_t packed_bytes) {
  if (!melfb_host || !packed_host || n_mels <= 0 || n_mels > kMaxMels) return AGA_ERR_INVALID_ARGUMENT;
  if (packed_bytes < kPackedBytes) return AGA_ERR_WORKSPACE_TOO_SMALL;
  std::vector<float> wA, wB;
  std::vector<int> L;
  if (!analyse_bands(melfb_host, n_mels, wA, wB, L)) return AGA_ERR_UNSUPPORTED;
  unsigned char* host = static_cast<unsigned char*>(packed_host);
  std::memset(host, 0, kPackedBytes);
  TcHeader* h = reinterpret_cast<TcHeader*>(host);
  h->magic = kMagic;
  h->n_mels = n_mels;
  h->L100 = L[100];
  h->L101 = L[101];
  float* bins = reinterpret_cast<float*>(host + kBinTableOff);
  for (int k = 0; k < kNfreq; ++k) {
    bins[4 * k + 0] = wA[k];
    bins[4 * k + 1] = wB[k];
    std::memcpy(&bins[4 * k + 2], &L[k], 4);
    bins[4 * k + 3] = 0.f;
  }
  if (!build_streams(wA, wB, L, n_mels, reinterpret_cast<float*>(host + kStreamOff),
                     reinterpret_cast<uint32_t*>(host + kStreamOff + 2 * 112 * 8), h->first_id))
    return AGA_ERR_UNSUPPORTED;
  build_dft_tables(host + kDftOff);
  return AGA_OK;
}

extern "C" int aga_logmel_tc_pack(const float* melfb_host, int n_mels, void* packed, size_t packed_bytes, void* stream) {
  if (!packed) return AGA_ERR_INVALID_ARGUMENT;
  std::vector<unsigned char> host(kPackedBytes, 0);
  const int st = aga_logmel_tc_build_host(melfb_host, n_mels, host.data(), packed_bytes < kPackedBytes ? packed_bytes : host.size());
  if (st != AGA_OK) return st;
  // pageable source: the runtime stages it before returning, so `host` may die at the end of this scope
  AGA_CUDA_TRY(cudaMemcpyAsync(packed, host.data(), kPackedBytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  return AGA_OK;
}

extern "C" int aga_logmel_tc_fwd(const float* audio, int64_t B, int64_t N, int64_t ld, const void* packed_tc, int n_mels,
                                 float* out, const int32_t* n_valid, void* workspace, size_t workspace_bytes, void* stream) {
  size_t need = 0;
  int st = aga_logmel_workspace_bytes(B, N, n_mels, &need);
  if (st != AGA_OK) return st;
  if (!audio || !packed_tc || !out || !workspace || ld < N) return AGA_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < need) return AGA_ERR_WORKSPACE_TOO_SMALL;
  if (B > 65535 || (reinterpret_cast<uintptr_t>(packed_tc) & 15)) return AGA_ERR_UNSUPPORTED;
  const int64_t F = N / kHop;  // 1 + N/160 frames, the last one dropped (whisper_encoder.py:117)
  if (F <= 0) return AGA_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint32_t* maxkey = static_cast<uint32_t*>(workspace);
  AGA_CUDA_TRY(cudaMemsetAsync(maxkey, 0, size_t(B) * sizeof(uint32_t), s));
  static_assert(Smem::kTotal <= 232448, "shared memory budget");
  auto kernel = logmel_tc_kernel;
  AGA_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Smem::kTotal)));
  const int64_t tiles_per_utt = (F + kTileFrames - 1) / kTileFrames;
  const int64_t n_tiles = tiles_per_utt * B;
  if (n_tiles > INT32_MAX) return AGA_ERR_UNSUPPORTED;
  static const int n_sm = []() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
  }();
  Params p;
  p.audio = audio;
  p.N = N;
  p.ld = ld;
  p.F = int(F);
  p.tiles_per_utt = int(tiles_per_utt);
  p.n_tiles = int(n_tiles);
  p.n_mels = n_mels;
  p.packed = static_cast<const unsigned char*>(packed_tc);
  p.out = out;
  p.maxkey = maxkey;
  p.n_valid = n_valid;
  const unsigned grid = unsigned(std::min<int64_t>(n_tiles, n_sm));
  kernel<<<grid, kThreads, Smem::kTotal, s>>>(p);
  AGA_AFTER_LAUNCH();
  return logmel_launch_normalise(out, maxkey, B, int64_t(n_mels) * F, F, n_valid, /*raw_power=*/1, s);
}
