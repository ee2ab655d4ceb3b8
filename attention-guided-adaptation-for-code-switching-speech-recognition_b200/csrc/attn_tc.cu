// bf16 flash attention forward on the 5th-gen tensor cores: tcgen05.mma with TMEM accumulators, TMA-staged
// Q/K/V tiles, warp-specialised (2 softmax warpgroups + 1 TMA warp + 2 MMA warps), head dim 64.
//
// Reference: MultiHeadAttention.qkv_attention, whisper/whisper/model.py:93-109 (non-causal: encoder self
// attention 1500x1500 and decoder cross attention Tx1500 — 99.9 % of the attention FLOPs of a step).
//
// One CTA = 256 query rows of one (batch, head): two 128-row tiles A and B, each owned by one softmax
// warpgroup, sharing every K/V tile that TMA brings in (halves L2->smem traffic per FLOP).
//   TMEM (512 columns): S_A [0,128) | S_B [128,256) | O_A [256,320) | O_B [320,384) | P_A [384,448) | P_B [448,512)
//     S_t = Q_t K^T   : tcgen05.mma  M=128 N=128 K=16 x4, A/B from smem (K-major, SWIZZLE_128B)
//     P_t (bf16, two keys per 32-bit column) has its own 64 columns, so S_t(j+1) can be issued as soon as the
//     softmax has pulled S_t(j) into registers — the tensor core refills S while the exponentials run
//     O_t += P_t V    : tcgen05.mma  M=128 N=64  K=16 x8, A from TMEM, B = V tile from smem (MN-major)
//   The MMA warp issues  S_A(j+1), S_B(j+1) (when the S registers were read), then PV_A(j), PV_B(j) (when P is
//   written): both GEMMs of tile j+1 / j hide under the softmax of tile j.
//   Online softmax in the exp2 domain with lazy rescaling: O_t is only rescaled (TMEM round trip) when the
//   running row maximum grew by more than 2^8, otherwise the stale maximum keeps being used.
#include "aga_common.cuh"
#include "attn_common.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>

// Debug timeline (compile with -DAGA_TIMELINE): CTA (0,0,0) records clock64() at named points into a global
// buffer set with aga_debug_set_timeline(); each role owns a row of 4096 slots.
#ifdef AGA_TIMELINE
__device__ long long* g_timeline = nullptr;
#define TL_DECL(role) long long* tl_ptr = (g_timeline && (role) >= 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_timeline + (role) * 4096 : nullptr; int tl_n = 0
#define TL(tag) do { if (tl_ptr && tl_n < 2046) { tl_ptr[2 * tl_n] = (tag); tl_ptr[2 * tl_n + 1] = clock64(); ++tl_n; } } while (0)
#define TL_END() do { if (tl_ptr) { tl_ptr[2 * tl_n] = -1; } } while (0)
extern "C" __attribute__((visibility("default"))) int aga_debug_set_timeline(long long* p) {
  return cudaMemcpyToSymbol(g_timeline, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}
// Per-CTA log: (clock64, globaltimer) at entry and exit plus the SM id, 8 slots per CTA.
__device__ long long* g_cta_log = nullptr;
extern "C" __attribute__((visibility("default"))) int aga_debug_set_cta_log(long long* p) {
  return cudaMemcpyToSymbol(g_cta_log, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}
__device__ __forceinline__ long long global_timer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void cta_log(int slot) {
  if (g_cta_log && threadIdx.x == 0) {
    long long* rec = g_cta_log + ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    rec[slot * 2] = clock64();
    rec[slot * 2 + 1] = global_timer_ns();
    rec[4] = smid;
  }
}
__device__ __forceinline__ void cta_mark(int idx) {  // extra clock marks: rec[5] main loop entered, rec[6] main loop left
  if (g_cta_log && threadIdx.x == 0)
    g_cta_log[((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + idx] = clock64();
}
#define CTA_LOG(slot) cta_log(slot)
#define CTA_MARK(idx) cta_mark(idx)
#else
#define CTA_LOG(slot) do { } while (0)
#define CTA_MARK(idx) do { } while (0)
#define TL_DECL(role) do { } while (0)
#define TL(tag) do { } while (0)
#define TL_END() do { } while (0)
#endif

namespace aga {
namespace {

using namespace ptx;

constexpr int kBlockM = 128;  // query rows per tile (= TMEM lanes)
constexpr int kBlockN = 128;  // keys per K/V tile
constexpr int kHeadDim = 64;
constexpr int kStages = 3;
constexpr int kTileBytes = kBlockN * kHeadDim * 2;  // 16 KiB
constexpr uint32_t kTmemCols = 512;
// Forward CTA shapes.  NT = query tiles (of 128 rows) per CTA:
//   NT = 2: two softmax warpgroups share every K/V tile (half the L2 -> smem traffic), one CTA per SM, all 512 TMEM columns;
//   NT = 1: 192 threads, 256 TMEM columns, 97 KiB of smem -> TWO CTAs per SM, whose prologues / epilogues and
//           non-MUFU phases overlap with the other CTA's exponentials without any cross-warpgroup protocol.
template <int NT> struct FwdCfg {
  static constexpr int kSoftmaxWarps = 4 * NT;
  static constexpr int kTmaWarp = 4 * NT;
  static constexpr int kMmaWarp = 4 * NT + 1;  // NT warps: one MMA-issuing warp per query tile
  static constexpr int kThreads = 32 * (4 * NT + 1 + NT);
  static constexpr uint32_t kCols = 256 * NT;
  static constexpr uint32_t kColS = 0, kColO = 128 * NT, kColP = 192 * NT;
  static constexpr int kStages = NT == 1 ? 2 : 3;  // K/V ring depth (an iteration is several TMA latencies long)
  static constexpr size_t kSmemBytes = 1024 /*align slack*/ + size_t(NT + 2 * kStages) * kTileBytes + 256 /*FwdSmem*/;
};
constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;  // d^-1/2 * log2(e)
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;
#ifndef AGA_FWD_POLY
#define AGA_FWD_POLY 0  // measured on B200: 0 -> 0.188 ms, 2 -> 0.192, 3 -> 0.198, 4 -> 0.194 (the loop is not MUFU-bound)
#endif
constexpr int kFwdPolyPairs = AGA_FWD_POLY;  // of every 8 element pairs of the forward softmax, this many use ex2_poly2

struct FwdSmem {
  // barriers
  uint64_t q_full;
  uint64_t k_full[kStages], k_empty[kStages], v_full[kStages], v_empty[kStages];
  uint64_t s_full[2], s_free[2], p_ready[2], pv_done[2];
  uint32_t tmem_base;
};
static_assert(sizeof(FwdSmem) <= 256, "FwdCfg::kSmemBytes reserves 256 bytes for the barrier block");

struct FwdArgs {
  int B, H, Tq, Tk;
  int64_t o_sb, o_st;
  __nv_bfloat16* out;
  float* lse;
  int causal;                  // key j visible to query i iff j <= i (Tq == Tk): key tiles above the diagonal are skipped
  int export_lo, export_hi;    // exported key columns [lo, hi) of the scaled, masked logits (decoder self attention)
  float* export_buf;           // (B, H, Tq, hi - lo) fp32 or nullptr
  const uint8_t* head_sel;     // (H) or nullptr: heads whose columns are exported
  const int32_t* kv_len;       // nullptr, or device scalar: only keys [0, min(*kv_len, Tk)) exist (zero-padded static shapes)
  // guided-loss reduction in the epilogue (aga_attn_params::guided_*): causal, one query tile
  const float* g_pattern;      // (B, Tq, 2) or nullptr
  float* g_part;               // (B, H, 4, 2)
  int g_early;
};

// One row's term of the guided loss (espnet_model.py:496-509) from the scaled logits of key columns 1, 2 (-inf where the
// causal mask hides them) and the row's target pattern: returns r and the two residuals (S~ - c) the gradient needs
// (zero where the reference zeroes the entry: hidden by the mask, or a pad row of a late layer).
__device__ __forceinline__ float guided_row(float s1, float s2, float2 pt, int early, float& res1, float& res2) {
  const bool h1 = isinf(s1), h2 = isinf(s2);
  float a1 = h1 ? 0.f : s1, a2 = h2 ? 0.f : s2;   // A[isinf(A)] = 0 (:497)
  float t1 = pt.x, t2 = pt.y;
  bool z = false;
  if (early) { t1 = 0.f; t2 = 0.f; }             // early layers: all-zero target in columns 1:3, pad rows are kept (:479-481)
  else if (isinf(pt.x) || isinf(pt.y)) { a1 = a2 = t1 = t2 = 0.f; z = true; }  // pad row: A and P zeroed (:496, :498)
  const float e1 = a1 - t1, e2 = a2 - t2;
  res1 = (h1 || z) ? 0.f : e1;                   // a zeroed entry is a constant: no gradient
  res2 = (h2 || z) ? 0.f : e2;
  return e1 * e1 + e2 * e2;
}

// effective key count of a launch whose key length is padded to a static Tk (aga_attn_params::kv_len)
__device__ __forceinline__ int effective_tk(const int32_t* kv_len, int Tk) {
  return kv_len ? max(1, min(__ldg(kv_len), Tk)) : Tk;
}

// 32 fp32 TMEM columns (scaled) -> the bf16 half `half` (64 bytes) of a 128-byte row of a SWIZZLE_128B staging tile
__device__ __forceinline__ void stage_half_row_bf16(uint32_t row_addr, int row, int half, const uint32_t (&v)[32], float scale) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(v[8 * c + 2 * e]) * scale, __uint_as_float(v[8 * c + 2 * e + 1]) * scale);
      w[e] = *reinterpret_cast<uint32_t*>(&hb);
    }
    sts128(row_addr + uint32_t(((half * 4 + c) ^ (row & 7)) * 16), w[0], w[1], w[2], w[3]);
  }
}
__device__ __forceinline__ void stage_row_bf16(uint32_t row_addr, int row, const uint32_t (&lo)[32], const uint32_t (&hi)[32], float scale) {
  stage_half_row_bf16(row_addr, row, 0, lo, scale);
  stage_half_row_bf16(row_addr, row, 1, hi, scale);
}

template <int NT>
__global__ void __launch_bounds__(FwdCfg<NT>::kThreads, NT == 1 ? 2 : 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o, const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using Cfg = FwdCfg<NT>;
  constexpr int kTmaWarp = Cfg::kTmaWarp, kMmaWarp = Cfg::kMmaWarp, kStages = Cfg::kStages;
  constexpr uint32_t kColS = Cfg::kColS, kColO = Cfg::kColO, kColP = Cfg::kColP;
  uint8_t* sQ = smem;                               // NT tiles
  uint8_t* sK = sQ + NT * kTileBytes;               // kStages tiles
  uint8_t* sV = sK + kStages * kTileBytes;          // kStages tiles
  FwdSmem* sb = reinterpret_cast<FwdSmem*>(sV + kStages * kTileBytes);

  CTA_LOG(0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row0 = qt * NT * kBlockM;
  const int Tk = effective_tk(a.kv_len, a.Tk);
  const int n_kt_all = (Tk + kBlockN - 1) / kBlockN;
  const int n_kt = a.causal ? min(n_kt_all, (row0 + NT * kBlockM - 1) / kBlockN + 1) : n_kt_all;
  const bool active_b = NT == 2 && row0 + kBlockM < a.Tq;  // tile B exists and holds at least one valid row

  if (threadIdx.x == 0) {
    mbar_init(&sb->q_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sb->k_full[s], 1);
      mbar_init(&sb->k_empty[s], active_b ? 2 : 1);  // one tcgen05.commit per MMA warp
      mbar_init(&sb->v_full[s], 1);
      mbar_init(&sb->v_empty[s], active_b ? 2 : 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&sb->s_full[t], 1);
      mbar_init(&sb->s_free[t], 4);   // one arrival per softmax warp of the warpgroup
      mbar_init(&sb->p_ready[t], 4);
      mbar_init(&sb->pv_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(&sb->tmem_base, Cfg::kCols);
    tmem_relinquish();
  }
  if (warp == kTmaWarp && lane == 0) {
    prefetch_tensormap(&map_q);
    prefetch_tensormap(&map_k);
    prefetch_tensormap(&map_v);
    prefetch_tensormap(&map_o);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sb->tmem_base;

  if (warp == kTmaWarp) {
    // ============================== TMA producer ==============================
    // the whole warp walks the loop (warp-uniform control flow); one elected lane issues
    if (elect_one()) {
      mbar_arrive_expect_tx(&sb->q_full, (active_b ? 2 : 1) * kTileBytes);
      tma_load_4d(sQ, &map_q, &sb->q_full, 0, h, row0, b);
      if (active_b) tma_load_4d(sQ + kTileBytes, &map_q, &sb->q_full, 0, h, row0 + kBlockM, b);
    }
    for (int j = 0; j < n_kt; ++j) {
      const int s = j % kStages;
      const uint32_t ph = (j / kStages) & 1;
      mbar_wait(&sb->k_empty[s], ph ^ 1);  // first pass through the ring returns immediately
      if (elect_one()) {
        mbar_arrive_expect_tx(&sb->k_full[s], kTileBytes);
        tma_load_4d(sK + s * kTileBytes, &map_k, &sb->k_full[s], 0, h, j * kBlockN, b);
      }
      mbar_wait(&sb->v_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&sb->v_full[s], kTileBytes);
        tma_load_4d(sV + s * kTileBytes, &map_v, &sb->v_full[s], 0, h, j * kBlockN, b);
      }
    }
  } else if (warp >= kMmaWarp) {
    // ============================== MMA issuers: one converged warp per query tile, elected lane issues ==========
    // (a single issuing thread for both tiles serialises ~6 barrier waits + 24 MMAs per key tile and becomes the
    //  critical path; with one warp per tile the two in-order streams interleave on the tensor pipe)
    const int t = warp - kMmaWarp;
    if (t == 0 || active_b) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(kBlockM, kBlockN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kBlockM, kHeadDim, 0, 1);
      TL_DECL(lane == 0 ? 3 * t : -1);
      const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ + t * kTileBytes));
      auto issue_s = [&](int stage) {
        const uint64_t dk = make_smem_desc_sw128(smem_u32(sK + stage * kTileBytes));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kHeadDim / 16; ++kk)  // +32 bytes along K inside the 128-byte swizzle row
            mma_ss(tmem + kColS + t * kBlockN, dq + uint64_t(kk * 2), dk + uint64_t(kk * 2), idesc_qk, kk > 0);
          tc_commit(&sb->s_full[t]);
          tc_commit(&sb->k_empty[stage]);
        }
        __syncwarp();
      };
      mbar_wait(&sb->q_full, 0);
      mbar_wait(&sb->k_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < n_kt; ++j) {
        const int s = j % kStages;
        const uint32_t ph = (j / kStages) & 1;
        if (j + 1 < n_kt) {  // S_t(j+1) as soon as the softmax holds S_t(j) in registers
          const int s1 = (j + 1) % kStages;
          TL(10);
          mbar_wait(&sb->k_full[s1], ((j + 1) / kStages) & 1);
          mbar_wait(&sb->s_free[t], j & 1);
          TL(12);
          tc_fence_after();
          issue_s(s1);
          TL(14);
        }
        mbar_wait(&sb->v_full[s], ph);
        mbar_wait(&sb->p_ready[t], j & 1);  // PV_t(j) as soon as P_t(j) is written
        TL(16);
        tc_fence_after();
        const uint64_t dv = make_smem_desc_sw128(smem_u32(sV + s * kTileBytes));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kBlockN / 16; ++kk)  // A: +8 TMEM columns (16 bf16); B: +16 key rows = 2048 bytes
            mma_ts(tmem + kColO + t * kHeadDim, tmem + kColP + t * 64 + kk * 8, dv + uint64_t(kk * 128), idesc_pv,
                   (j > 0 || kk > 0) ? 1u : 0u);
          tc_commit(&sb->pv_done[t]);
          tc_commit(&sb->v_empty[s]);
        }
        __syncwarp();
        TL(18);
      }
      TL_END();
    }
  } else {
    // ============================== softmax / correction / epilogue ==============================
    const int t = warp >> 2;                        // warpgroup 0 -> tile A, 1 -> tile B
    const uint32_t lane_base = uint32_t((warp & 3) * 32);
    const int row = row0 + t * kBlockM + int(lane_base) + lane;
    if (t == 0 || active_b) {
      const uint32_t t_s = tmem + (lane_base << 16) + kColS + t * kBlockN;
      const uint32_t t_o = tmem + (lane_base << 16) + kColO + t * kHeadDim;
      const uint32_t t_p = tmem + (lane_base << 16) + kColP + t * 64;
      float m_used = -INFINITY, l = 0.f;
      TL_DECL((lane == 0 && (warp & 3) == 0) ? 1 + t : -1);
      const int W = a.export_hi - a.export_lo;
      float* erow = (a.export_buf && (a.head_sel == nullptr || a.head_sel[h] != 0) && row < a.Tq)
                        ? a.export_buf + ((int64_t(b) * a.H + h) * a.Tq + row) * W - a.export_lo : nullptr;

      // The two warpgroups share each SM sub-partition's MUFU.  Their exp2 phases are forced to ALTERNATE with a pair
      // of named barriers (id 2: "A may start its exp2 phase", id 3: "B may start"): while one warpgroup runs its
      // exponentials at the full MUFU rate, the other does its barrier waits, TMEM loads, row maxima and stores.
#ifdef AGA_FWD_TOKEN  // measured slower than free-running warpgroups (a lone warp per sub-partition issues MUFU at half rate)
      const bool pingpong = active_b;
#else
      const bool pingpong = false;
#endif
      if (pingpong && t == 1) named_bar_arrive(2, 256);
      for (int j = 0; j < n_kt; ++j) {
        TL(20);
        mbar_wait(&sb->s_full[t], j & 1);
        TL(21);
        tc_fence_after();
        // ---- the whole S row (128 fp32) into registers, then hand the TMEM columns back to the MMA warp
        uint32_t sr[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(t_s + c * 32, sr[c]);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb->s_free[t]);
        TL(22);
        int valid = Tk - j * kBlockN;  // keys of this tile this row may see (>= 128 except on the last / diagonal tile)
        if (a.causal) valid = min(valid, row - j * kBlockN + 1);
        if (valid < kBlockN) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= valid) sr[c][i] = 0xff800000u;  // -inf: exp2 -> 0
        }
        if (erow != nullptr && j * kBlockN < a.export_hi && (j + 1) * kBlockN > a.export_lo) {
          // side buffer of the guided loss: scaled, masked logits of the selected key columns (whisper/model.py:103,
          // what the reference returns as `qk`), written straight from the S registers
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int key = j * kBlockN + c * 32 + i;
              if (key >= a.export_lo && key < a.export_hi) erow[key] = __uint_as_float(sr[c][i]) * 0.125f;
            }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[0][i]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[1][i]));
          mx2 = fmaxf(mx2, __uint_as_float(sr[2][i]));
          mx3 = fmaxf(mx3, __uint_as_float(sr[3][i]));
        }
        const float m_new = fmaxf(m_used, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * kScaleLog2);
        if (j > 0) {
          TL(23);
          mbar_wait(&sb->pv_done[t], (j - 1) & 1);  // P_t buffer consumed and O_t stable
          tc_fence_after();
          TL(24);
        }
        if (j == 0) {
          m_used = m_new;
        } else if (__any_sync(0xffffffffu, m_new - m_used > kRescaleThreshold)) {
          const float alpha = ex2(m_used - m_new);
          l *= alpha;
          m_used = m_new;
#pragma unroll 1
          for (int c = 0; c < kHeadDim / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(t_o + c * 32, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st32(t_o + c * 32, r);
          }
        }
        // ---- P = exp2(S*c - m) -> bf16 pairs -> its own TMEM columns (packed f32x2 FMA / ADD halve the issue slots)
        if (j == 0) CTA_MARK(5);
        float neg_m = -m_used;
        if (pingpong) neg_m = named_bar_sync_dep(2 + t, 256, neg_m);
        float2 rs = make_float2(0.f, 0.f);
        const float2 sc2 = make_float2(kScaleLog2, kScaleLog2), nm2 = make_float2(neg_m, neg_m);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(sr[c][2 * i]), __uint_as_float(sr[c][2 * i + 1])), sc2, nm2);
            // kFwdPolyPairs of every 8 pairs take the FMA-pipe exp2, the rest the MUFU
            const float2 pp = (i & 7) < kFwdPolyPairs ? ex2_poly2(x) : make_float2(ex2(x.x), ex2(x.y));
            rs = __fadd2_rn(rs, pp);
            __nv_bfloat162 hb = __floats2bfloat162_rn(pp.x, pp.y);
            pk[i] = *reinterpret_cast<uint32_t*>(&hb);
          }
          tmem_st16(t_p + c * 16, pk);
        }
        float tile_sum = rs.x + rs.y;
        if (pingpong && !(t == 1 && j == n_kt - 1)) tile_sum = named_bar_arrive_dep(3 - t, 256, tile_sum);
        l += tile_sum;
        TL(25);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb->p_ready[t]);
        TL(26);
      }
      CTA_MARK(6);
      TL_END();
      // ---- epilogue: O / l -> bf16 rows staged in this tile's Q buffer (its last S GEMM is long done) -> one TMA tile
      //      store per warp (rows past Tq are clipped).  A lane-per-row store of 16-byte pieces costs one L1 transaction per
      //      lane per instruction and was 7 % of the CTA's lifetime.
      mbar_wait(&sb->pv_done[t], (n_kt - 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / l;
      {
        uint8_t* stage = sQ + t * kTileBytes + (warp & 3) * (32 * 128);
#pragma unroll 1
        for (int c = 0; c < kHeadDim / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_o + c * 32, r);
          tmem_wait_ld();
          stage_half_row_bf16(smem_u32(stage + lane * 128), lane, c, r, inv);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&map_o, stage, 0, h, row0 + t * kBlockM + int(lane_base), b);
          bulk_commit_group();
          bulk_wait_group_read0();  // the staging rows must outlive the store's reads (the CTA exits next)
        }
      }
      if (row < a.Tq && a.lse) a.lse[(int64_t(b) * a.H + h) * a.Tq + row] = (m_used + log2f(l)) * kLn2;
    }
  }
  tc_fence_before();
  __syncthreads();
  CTA_LOG(1);
  if (warp == kMmaWarp) tmem_dealloc(tmem, Cfg::kCols);
}

// ------------------------------------------------------------------------------------------ forward, column-split softmax
// Same pipeline as attn_fwd_tc_kernel<1> (one 128-row query tile per CTA, two CTAs per SM, same TMEM map), but every
// 32-row group of S is shared by TWO softmax warps, 64 key columns each: 8 softmax warps per CTA, FOUR per SM
// sub-partition.  With one warp per sub-partition and CTA, a tile's serial chain (wait S, TMEM load, row max, wait PV,
// 128 exp2, TMEM store: 2900 cycles, of which 1024 on the MUFU) left the MUFU 30 % idle; half-rows shorten every link
// of the chain and give the scheduler four chains per MUFU to interleave.  The two warps of a pair agree on the
// running row maximum through a parity-double-buffered smem slot and a 64-thread named barrier per tile; the row sums
// are combined once at the end.
#ifndef AGA_FS_POLY
#define AGA_FS_POLY 2  // measured on B200 (B16 H12 1500x1500): 0 -> 631, 2 -> 679, 3 -> 663 TFLOP/s
#endif
constexpr int kFsPolyPairs = AGA_FS_POLY;
constexpr int kFsSoftmaxWarps = 8;
constexpr int kFsTmaWarp = 8, kFsMmaWarp = 9;
constexpr int kFsThreads = 10 * 32;
constexpr int kFsStages = 2;
struct FsSmem {
  uint64_t q_full;
  uint64_t k_full[kFsStages], k_empty[kFsStages], v_full[kFsStages], v_empty[kFsStages];
  uint64_t s_full, s_free, p_ready, pv_done;
  uint32_t tmem_base;
  alignas(16) float xch[2][2][kBlockM];  // [tile parity][column half][row]: local row maxima (and the row sums at the end)
};
constexpr size_t kFsSmemBytes = 1024 + size_t(1 + 2 * kFsStages) * kTileBytes + sizeof(FsSmem);
static_assert(2 * kFsSmemBytes <= 227 * 1024, "two CTAs per SM");

__global__ void __launch_bounds__(kFsThreads, 2)
attn_fwd_tc_split_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                         const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o, const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t kColS = 0, kColO = 128, kColP = 192;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTileBytes;
  uint8_t* sV = sK + kFsStages * kTileBytes;
  FsSmem* sb = reinterpret_cast<FsSmem*>(sV + kFsStages * kTileBytes);

  CTA_LOG(0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row0 = qt * kBlockM;
  const int Tk = effective_tk(a.kv_len, a.Tk);
  const int n_kt_all = (Tk + kBlockN - 1) / kBlockN;
  const int n_kt = a.causal ? min(n_kt_all, (row0 + kBlockM - 1) / kBlockN + 1) : n_kt_all;

  if (threadIdx.x == 0) {
    mbar_init(&sb->q_full, 1);
    for (int s = 0; s < kFsStages; ++s) {
      mbar_init(&sb->k_full[s], 1);
      mbar_init(&sb->k_empty[s], 1);
      mbar_init(&sb->v_full[s], 1);
      mbar_init(&sb->v_empty[s], 1);
    }
    mbar_init(&sb->s_full, 1);
    mbar_init(&sb->s_free, kFsSoftmaxWarps);
    mbar_init(&sb->p_ready, kFsSoftmaxWarps);
    mbar_init(&sb->pv_done, 1);
    fence_barrier_init();
  }
  if (warp == kFsMmaWarp) {
    tmem_alloc(&sb->tmem_base, 256);
    tmem_relinquish();
  }
  if (warp == kFsTmaWarp && lane == 0) {
    prefetch_tensormap(&map_q);
    prefetch_tensormap(&map_k);
    prefetch_tensormap(&map_v);
    prefetch_tensormap(&map_o);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sb->tmem_base;

  if (warp == kFsTmaWarp) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      mbar_arrive_expect_tx(&sb->q_full, kTileBytes);
      tma_load_4d(sQ, &map_q, &sb->q_full, 0, h, row0, b);
    }
    for (int j = 0; j < n_kt; ++j) {
      const int s = j % kFsStages;
      const uint32_t ph = (j / kFsStages) & 1;
      mbar_wait(&sb->k_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&sb->k_full[s], kTileBytes);
        tma_load_4d(sK + s * kTileBytes, &map_k, &sb->k_full[s], 0, h, j * kBlockN, b);
      }
      mbar_wait(&sb->v_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&sb->v_full[s], kTileBytes);
        tma_load_4d(sV + s * kTileBytes, &map_v, &sb->v_full[s], 0, h, j * kBlockN, b);
      }
    }
  } else if (warp == kFsMmaWarp) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc_qk = make_idesc_bf16(kBlockM, kBlockN, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(kBlockM, kHeadDim, 0, 1);
    const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ));
    auto issue_s = [&](int stage) {
      const uint64_t dk = make_smem_desc_sw128(smem_u32(sK + stage * kTileBytes));
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk) mma_ss(tmem + kColS, dq + uint64_t(kk * 2), dk + uint64_t(kk * 2), idesc_qk, kk > 0);
        tc_commit(&sb->s_full);
        tc_commit(&sb->k_empty[stage]);
      }
      __syncwarp();
    };
    mbar_wait(&sb->q_full, 0);
    mbar_wait(&sb->k_full[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < n_kt; ++j) {
      const int s = j % kFsStages;
      const uint32_t ph = (j / kFsStages) & 1;
      if (j + 1 < n_kt) {  // S(j+1) as soon as the softmax warps hold S(j) in registers
        const int s1 = (j + 1) % kFsStages;
        mbar_wait(&sb->k_full[s1], ((j + 1) / kFsStages) & 1);
        mbar_wait(&sb->s_free, j & 1);
        tc_fence_after();
        issue_s(s1);
      }
      mbar_wait(&sb->v_full[s], ph);
      mbar_wait(&sb->p_ready, j & 1);
      tc_fence_after();
      const uint64_t dv = make_smem_desc_sw128(smem_u32(sV + s * kTileBytes));
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kBlockN / 16; ++kk)
          mma_ts(tmem + kColO, tmem + kColP + kk * 8, dv + uint64_t(kk * 128), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        tc_commit(&sb->pv_done);
        tc_commit(&sb->v_empty[s]);
      }
      __syncwarp();
    }
  } else {
    // ============================== softmax / correction / epilogue: warp = (row quadrant, column half) ==============
    const int quad = warp & 3, ch = warp >> 2;
    const uint32_t lane_base = uint32_t(quad * 32);
    const int rloc = int(lane_base) + lane;
    const int row = row0 + rloc;
    const uint32_t t_s = tmem + (lane_base << 16) + kColS + ch * 64;
    const uint32_t t_o = tmem + (lane_base << 16) + kColO + ch * 32;  // this warp's half of the O columns
    const uint32_t t_p = tmem + (lane_base << 16) + kColP + ch * 32;  // bf16 pairs of its 64 keys
    const int pair_bar = 2 + quad;                                     // named barrier of the two warps sharing these rows
    float m_used = -INFINITY, l = 0.f;
    const int W = a.export_hi - a.export_lo;
    float* erow = (a.export_buf && (a.head_sel == nullptr || a.head_sel[h] != 0) && row < a.Tq)
                      ? a.export_buf + ((int64_t(b) * a.H + h) * a.Tq + row) * W - a.export_lo : nullptr;
    if (row0 + int(lane_base) >= a.Tq) {
      // all 32 rows of this warp (and of its pair) lie past Tq (decoder tiles: Tq = 64): only keep the MMA warp's
      // barriers moving — no TMEM traffic, no exponentials; the PV GEMM's rows for them are never read
      for (int j = 0; j < n_kt; ++j) {
        mbar_wait(&sb->s_full, j & 1);
        if (lane == 0) mbar_arrive(&sb->s_free);
        if (j > 0) mbar_wait(&sb->pv_done, (j - 1) & 1);
        if (lane == 0) mbar_arrive(&sb->p_ready);
      }
    } else {
    for (int j = 0; j < n_kt; ++j) {
      mbar_wait(&sb->s_full, j & 1);
      tc_fence_after();
      uint32_t sr[2][32];
      tmem_ld32(t_s, sr[0]);
      tmem_ld32(t_s + 32, sr[1]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb->s_free);
      const int kbase = j * kBlockN + ch * 64;
      int valid = Tk - kbase;  // keys of this 64-column slice the row may see
      if (a.causal) valid = min(valid, row - kbase + 1);
      if (valid < 64) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) sr[c][i] = 0xff800000u;  // -inf: exp2 -> 0
      }
      if (erow != nullptr && kbase < a.export_hi && kbase + 64 > a.export_lo) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int key = kbase + c * 32 + i;
            if (key >= a.export_lo && key < a.export_hi) erow[key] = __uint_as_float(sr[c][i]) * 0.125f;
          }
      }
      if (a.g_part != nullptr && kbase == 0) {
        // guided loss, reduced here: this warp holds key columns 1, 2 of its 32 query rows
        float r = 0.f, u1, u2;
        if (row < a.Tq) {
          const float2 pt = *reinterpret_cast<const float2*>(a.g_pattern + (int64_t(b) * a.Tq + row) * 2);
          r = guided_row(__uint_as_float(sr[0][1]) * 0.125f, __uint_as_float(sr[0][2]) * 0.125f, pt, a.g_early, u1, u2);
        }
        const float rs = warp_sum(r), rc = warp_sum(r != 0.f ? 1.f : 0.f);
        if (lane == 0) {
          float* dst = a.g_part + ((int64_t(b) * a.H + h) * 4 + quad) * 2;
          dst[0] = rs;
          dst[1] = rc;
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx0 = fmaxf(mx0, __uint_as_float(sr[0][i]));
        mx1 = fmaxf(mx1, __uint_as_float(sr[1][i]));
      }
      // ---- agree on the row maximum with the warp that holds the other 64 columns of these rows
      float* slot = sb->xch[j & 1][0];
      slot[ch * kBlockM + rloc] = fmaxf(mx0, mx1);
      named_bar_sync(pair_bar, 64);
      const float m_new = fmaxf(m_used, fmaxf(slot[rloc], slot[kBlockM + rloc]) * kScaleLog2);
      if (j > 0) {
        mbar_wait(&sb->pv_done, (j - 1) & 1);  // P buffer consumed and O stable
        tc_fence_after();
      }
      if (j == 0) {
        m_used = m_new;
      } else if (__any_sync(0xffffffffu, m_new - m_used > kRescaleThreshold)) {  // same rows, same decision in both warps
        const float alpha = ex2(m_used - m_new);
        l *= alpha;
        m_used = m_new;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {  // 16 columns at a time: the S row stays in registers across this rare path
          uint32_t r[16];
          tmem_ld16(t_o + c * 16, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st16(t_o + c * 16, r);
        }
      }
      const float neg_m = -m_used;
      float2 rs = make_float2(0.f, 0.f);
      const float2 sc2 = make_float2(kScaleLog2, kScaleLog2), nm2 = make_float2(neg_m, neg_m);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(sr[c][2 * i]), __uint_as_float(sr[c][2 * i + 1])), sc2, nm2);
          // kFsPolyPairs of every 8 pairs take the FMA-pipe exp2 (tc_ptx.cuh), the rest the MUFU
          const float2 pp = (i & 7) < kFsPolyPairs ? ex2_poly2(x) : make_float2(ex2(x.x), ex2(x.y));
          rs = __fadd2_rn(rs, pp);
          __nv_bfloat162 hb = __floats2bfloat162_rn(pp.x, pp.y);
          pk[i] = *reinterpret_cast<uint32_t*>(&hb);
        }
        tmem_st16(t_p + c * 16, pk);
      }
      l += rs.x + rs.y;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb->p_ready);
    }
    // ---- epilogue: combine the two partial row sums, O / l -> bf16 half rows staged in the dead Q buffer, one TMA tile
    //      store per 32-row group (issued by the pair's first warp), lse
    float* slot = sb->xch[n_kt & 1][0];
    slot[ch * kBlockM + rloc] = l;
    named_bar_sync(pair_bar, 64);
    l = slot[rloc] + slot[kBlockM + rloc];
    mbar_wait(&sb->pv_done, (n_kt - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    uint8_t* stage = sQ + quad * (32 * 128);
    {
      uint32_t r[32];
      tmem_ld32(t_o, r);
      tmem_wait_ld();
      stage_half_row_bf16(smem_u32(stage + lane * 128), lane, ch, r, inv);
    }
    fence_proxy_async_smem();
    named_bar_sync(pair_bar, 64);
    if (ch == 0) {
      if (lane == 0) {
        tma_store_4d(&map_o, stage, 0, h, row0 + int(lane_base), b);
        bulk_commit_group();
        bulk_wait_group_read0();  // the staging rows must outlive the store's reads (the CTA exits next)
      }
      if (row < a.Tq && a.lse) a.lse[(int64_t(b) * a.H + h) * a.Tq + row] = (m_used + log2f(l)) * kLn2;
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  CTA_LOG(1);
  if (warp == kFsMmaWarp) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------ host
PFN_cuTensorMapEncodeTiled get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = []() -> PFN_cuTensorMapEncodeTiled {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  }();
  return fn;
}

// (B, T, H*64) bf16 activations as a rank-4 tensor (c=64, h, t, b); box = 64 x 1 x rows x 1, SWIZZLE_128B.
// Rows past T are zero-filled by the TMA unit.
int make_map(CUtensorMap* map, const void* base, int B, int H, int T, int64_t stride_b, int64_t stride_t, int rows) {
  ensure_context_in_this_thread();
  PFN_cuTensorMapEncodeTiled enc = get_encode_fn();
  if (!enc) return AGA_ERR_UNSUPPORTED;
  cuuint64_t dims[4] = {cuuint64_t(kHeadDim), cuuint64_t(H), cuuint64_t(T), cuuint64_t(B)};
  cuuint64_t strides[3] = {cuuint64_t(kHeadDim) * 2, cuuint64_t(stride_t) * 2, cuuint64_t(stride_b) * 2};
  cuuint32_t box[4] = {cuuint32_t(kHeadDim), 1, cuuint32_t(rows), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? AGA_OK : AGA_ERR_INVALID_ARGUMENT;
}

}  // namespace

// exported columns are written from registers with fully unrolled, predicated stores: a narrow window only
// (the guided loss reads key columns 1:3); full maps and probability export stay on the CUDA-core path
constexpr int kMaxExportCols = 16;

bool attn_tc_supported(const aga_attn_params& p) {
  if (p.dtype != AGA_BF16) return false;
  if (p.export_kind != AGA_EXPORT_NONE &&
      (p.export_kind != AGA_EXPORT_LOGITS || p.export_hi - p.export_lo > kMaxExportCols))
    return false;
  // TMA: global strides are multiples of 16 bytes (validated by the caller) and below 2^40 bytes
  return get_encode_fn() != nullptr;
}
size_t attn_tc_fwd_workspace(const aga_attn_params&) { return 0; }

int attn_tc_fwd(const aga_attn_params& p, void*, cudaStream_t s) {
  CUtensorMap mq, mk, mv;
  int st;
  if ((st = make_map(&mq, p.q, p.B, p.H, p.Tq, p.q_stride_b, p.q_stride_t, kBlockM)) != AGA_OK) return st;
  if ((st = make_map(&mk, p.k, p.B, p.H, p.Tk, p.k_stride_b, p.k_stride_t, kBlockN)) != AGA_OK) return st;
  if ((st = make_map(&mv, p.v, p.B, p.H, p.Tk, p.v_stride_b, p.v_stride_t, kBlockN)) != AGA_OK) return st;
  CUtensorMap mo;  // 32-row boxes: one store per softmax warp
  if ((st = make_map(&mo, p.out, p.B, p.H, p.Tq, p.o_stride_b, p.o_stride_t, 32)) != AGA_OK) return st;
  const bool exporting = p.export_kind == AGA_EXPORT_LOGITS && p.export_buf != nullptr;
  FwdArgs a{p.B, p.H, p.Tq, p.Tk, p.o_stride_b, p.o_stride_t, static_cast<__nv_bfloat16*>(p.out), p.lse, p.causal,
            exporting ? p.export_lo : 0, exporting ? p.export_hi : 0, exporting ? p.export_buf : nullptr,
            exporting ? p.head_sel : nullptr, p.kv_len, p.guided_pattern, p.guided_part, p.guided_early};
#ifdef AGA_FWD_TWO_TILES  // one CTA per SM, two query tiles sharing each K/V tile
  constexpr int NT = 2;
#else                     // two independent single-tile CTAs per SM (measured faster: see DESIGN.md)
  constexpr int NT = 1;
#endif
#ifndef AGA_FWD_NO_SPLIT  // column-split softmax: 8 softmax warps per CTA (measured faster, see the kernel's header)
  (void)NT;
  AGA_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kFsSmemBytes)));
  dim3 grid_s((p.Tq + kBlockM - 1) / kBlockM, p.H, p.B);
  attn_fwd_tc_split_kernel<<<grid_s, kFsThreads, kFsSmemBytes, s>>>(mq, mk, mv, mo, a);
#else
  AGA_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    int(FwdCfg<NT>::kSmemBytes)));
  dim3 grid((p.Tq + NT * kBlockM - 1) / (NT * kBlockM), p.H, p.B);
  attn_fwd_tc_kernel<NT><<<grid, FwdCfg<NT>::kThreads, FwdCfg<NT>::kSmemBytes, s>>>(mq, mk, mv, mo, a);
#endif
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

namespace {
// =============================================================================================== backward
// PERSISTENT kernel: one CTA per SM walks work items (batch, head, 128-key tile); for each item K_j, V_j sit in smem
// and the CTA streams the 128-row query tiles of that (batch, head) past them.  All rings and barriers run on a
// CTA-global tile counter: the next item's K_j (double-buffered: the current one is read until the item's last dQ GEMM),
// V_j, first (Q, dO) tiles, S^T and dP^T are in flight while the current item's last tiles drain — no per-item
// prologue / epilogue / launch gap (they were 28 % of a one-item CTA).
//
// The score tiles are computed TRANSPOSED (keys in the TMEM lanes, queries along the columns), so that P^T and
// dS^T — the M x K operands of the dV and dK GEMMs — never leave tensor memory.  Each query tile is handled as two
// independent 64-query halves g = 0, 1:
//   S^T_g  = K_j Q_ig^T       cols [64g, 64g+64)          SS, both K-major, N = 64
//   dP^T_g = V_j dO_ig^T      cols [128+64g, 128+64g+64)  SS, both K-major, N = 64
//   dV_j  += P^T_g dO_ig      [256,320)   A = P^T_g  (TMEM, bf16 pairs written over the S^T_g columns), B = dO rows (MN-major)
//   dK_j  += dS^T_g Q_ig      [320,384)   A = dS^T_g (TMEM, bf16 pairs written over the dP^T_g columns), B = Q rows (MN-major)
//   dQ_i   = dS K_j           [384,448)   A = dS^T rows in smem read as an MN-major operand (both halves), B = K_j (MN-major)
// Warp roles (768 threads):
//   0-15  softmax: half g = warp >> 3, 32-column slice sub = (warp >> 2) & 1, lane quadrant warp & 3 (lane = key).
//         phase 1  P = exp2(S c - lse[q]) -> TMEM (MUFU);  phase 2  dS = P (dP - delta[q]) -> TMEM + one swizzled smem row
//         chunk per thread (FMA / LSU).  The two halves alternate their phase 1 (a pair of named barriers), so half 1
//         runs half a tile behind half 0.  At the end of an item half 0's warps store dV_j, half 1's warps dK_j.
//   16-19 dQ drain: TMEM -> swizzled fp32 staging -> cp.reduce.async.bulk add into the tile-major global accumulator
//         (b, h, q-tile, column half, 128 rows, 32).
//   20    TMA producer: K_j, V_j per item; (Q_i, dO_i, lse/delta rows) per tile.
//   21,22 MMA streams of halves 0, 1 (S^T, dV, dP^T, dK);  23  MMA stream of dQ.  One in-order issuer per stream:
//         every "softmax event -> group of MMAs" costs a few hundred cycles of wait + descriptor set-up.
// Rows/keys past the tensor ends are zero-filled by TMA, which makes their contributions exactly zero.
#ifndef AGA_BWD_POLY
#define AGA_BWD_POLY 0
#endif
constexpr int kBwdPolyQuads = AGA_BWD_POLY;  // of every 4 element quads of phase 1, this many use ex2_poly2
constexpr int kBwdThreads = 768;
constexpr int kBwdSoftmaxWarps = 16;
[[maybe_unused]] constexpr int kBwdDrainWarp0 = 16;  // (timeline build)
constexpr int kBwdTmaWarp = 20;
constexpr int kBwdMmaWarp = 21;
constexpr uint32_t kColBS = 0, kColBdP = 128, kColBdV = 256, kColBdK = 320, kColBdQ = 384;
constexpr int kPanelBytes = kBlockM * 128;             // 128 rows x 64 bf16
constexpr int kDqStageBytes = kBlockM * 32 * 4;        // one 32-column half of a dQ tile, fp32: 16 KiB
constexpr int kBwdStages = 3;                          // (Q_i, dO_i, stats) ring: a tile's TMA is issued two tiles ahead

struct BwdSmem {
  uint64_t k_full[2], k_empty[2], v_full, v_empty;
  uint64_t qdo_full[kBwdStages], qdo_empty[kBwdStages];
  uint64_t s_full[2], dp_full[2], p_ready[2], ds_ready[2], ds_free[2], dq_full, dq_empty;
  uint64_t dv_init, dk_init, dv_final, dk_final, dk_read;
  uint32_t tmem_base;
  alignas(16) float stats[kBwdStages][2][kBlockM];  // per stage: lse * log2(e) | delta of the 128 queries
};
// 2 x K, V | kBwdStages x (Q, dO) | dS^T panels: half 0 double-buffered, half 1 single | dQ staging
constexpr size_t kBwdSmemBytes = 1024 + size_t(3 + 2 * kBwdStages) * kTileBytes + 3 * size_t(kPanelBytes) + kDqStageBytes + sizeof(BwdSmem);
static_assert(kBwdSmemBytes <= 227 * 1024, "backward kernel exceeds the 227 KiB shared-memory limit");

struct BwdArgs {
  int B, H, Tq, Tk;
  int n_items;      // B * H * ceil(Tk / 128)
  int64_t k_sb, k_st, v_sb, v_st;
  const float* stats;  // (B, H, ceil(Tq/128), 2, 128) fp32: lse * log2(e) | delta per query tile, zero past Tq
  float* dq_accum;  // (B, H, ceil(Tq/128), 2, 128, 32) fp32, zero-initialised, 16-byte chunks XOR-swizzled with (row & 7)
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  const int32_t* kv_len;  // nullptr, or device scalar: keys at or past it do not exist (their P is forced to 0; key tiles past it are skipped)
  // decoder self attention with more than one query tile (Tq = Tk in (128, 448], whisper/model.py:322): causal mask, and
  // the gradient of the exported logit columns [lo, hi) (the guided loss, espnet_model.py:463-530) added to dS
  int causal;
  int export_lo, export_hi;
  const float* d_export;      // (B, H, Tq, hi - lo) fp32 or nullptr
  const uint8_t* head_sel;    // (H) or nullptr
};

// MN-major operand spanning two 64-element panels along M (dS^T rows as the A of dQ): LBO = panel stride
__device__ __forceinline__ uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t(lbo_bytes >> 4) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                   const __grid_constant__ CUtensorMap map_dk, const __grid_constant__ CUtensorMap map_dv,
                   const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                            // 2 item buffers
  uint8_t* sV = sK + 2 * kTileBytes;
  uint8_t* sQ = sV + kTileBytes;                 // kBwdStages stages
  uint8_t* sdO = sQ + kBwdStages * kTileBytes;   // kBwdStages stages
  // dS^T rows [key][64 queries]: half 0's panel is double-buffered (its tile c+1 is written while dQ(c) still waits for
  // half 1), half 1's is not (dQ(c) is issued as soon as half 1 has written tile c): [half 0, c even][half 0, c odd][half 1]
  uint8_t* sdS = sdO + kBwdStages * kTileBytes;
  uint8_t* sdQ = sdS + 3 * kPanelBytes;          // fp32 staging, 4 warps x (32 rows x 128 B)
  BwdSmem* sb = reinterpret_cast<BwdSmem*>(sdQ + kDqStageBytes);

  CTA_LOG(0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = (a.Tq + kBlockM - 1) / kBlockM;
  const int Tk = effective_tk(a.kv_len, a.Tk);
  const int n_kt = (Tk + kBlockN - 1) / kBlockN;
  // work items of this CTA: w = blockIdx.x + it * gridDim.x;  item w = ((b * H + h) * n_kt + kt)
  // (with kv_len the item list is re-enumerated over the key tiles that exist: a.n_items is the static upper bound)
  const int n_items = a.kv_len ? a.B * a.H * n_kt : a.n_items;
  const int first_item = blockIdx.x, item_step = gridDim.x;
  const int my_items = first_item < n_items ? (n_items - first_item + item_step - 1) / item_step : 0;
  const int total_tiles = my_items * n_qt;  // tiles this CTA processes (the global tile counter c runs over them)

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sb->k_full[s], 1);
      mbar_init(&sb->k_empty[s], 3);  // commits of the three MMA streams after their last read of K_j
    }
    mbar_init(&sb->v_full, 1);
    mbar_init(&sb->v_empty, 2);  // commits of the two half streams after the item's last dP^T
    for (int s = 0; s < kBwdStages; ++s) {
      mbar_init(&sb->qdo_full[s], 1);
      mbar_init(&sb->qdo_empty[s], 2);  // one tcgen05.commit per half stream
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&sb->s_full[g], 1);
      mbar_init(&sb->dp_full[g], 1);
      mbar_init(&sb->p_ready[g], 8);   // one arrival per warp of the half
      mbar_init(&sb->ds_ready[g], 8);
      mbar_init(&sb->ds_free[g], 1);   // indexed by dS^T buffer
    }
    mbar_init(&sb->dq_full, 1);
    mbar_init(&sb->dq_empty, 4);
    mbar_init(&sb->dv_init, 1);
    mbar_init(&sb->dk_init, 1);
    mbar_init(&sb->dv_final, 2);
    mbar_init(&sb->dk_final, 2);
    mbar_init(&sb->dk_read, 8);
    fence_barrier_init();
  }
  if (warp == kBwdMmaWarp) {
    tmem_alloc(&sb->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  if (warp == kBwdTmaWarp && lane == 0) {
    prefetch_tensormap(&map_q);
    prefetch_tensormap(&map_k);
    prefetch_tensormap(&map_v);
    prefetch_tensormap(&map_do);
    prefetch_tensormap(&map_dk);
    prefetch_tensormap(&map_dv);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sb->tmem_base;
  auto decode = [&](int it, int& kt, int& h, int& b) {
    const int w = first_item + it * item_step;
    kt = w % n_kt;
    const int bh = w / n_kt;
    h = bh % a.H;
    b = bh / a.H;
  };

  if (warp == kBwdTmaWarp) {
    // ============================== TMA producer ==============================
    // per item: K_j (the buffer item it-2 used), V_j (free since the previous item's last dP^T), then the (Q, dO) tiles
    int c = 0;
    auto load_qdo = [&](int i, int h, int b) {
      const int s = c % kBwdStages;
      mbar_wait(&sb->qdo_empty[s], ((c / kBwdStages) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&sb->qdo_full[s], 2 * kTileBytes + 2 * kBlockM * 4);
        bulk_load(sb->stats[s], a.stats + ((int64_t(b) * a.H + h) * n_qt + i) * (2 * kBlockM), 2 * kBlockM * 4, &sb->qdo_full[s]);
        tma_load_4d(sQ + s * kTileBytes, &map_q, &sb->qdo_full[s], 0, h, i * kBlockM, b);
        tma_load_4d(sdO + s * kTileBytes, &map_do, &sb->qdo_full[s], 0, h, i * kBlockM, b);
      }
      ++c;
    };
    for (int it = 0; it < my_items; ++it) {
      int kt, h, b;
      decode(it, kt, h, b);
      const int kb = it & 1;
      mbar_wait(&sb->k_empty[kb], ((it >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&sb->k_full[kb], kTileBytes);
        tma_load_4d(sK + kb * kTileBytes, &map_k, &sb->k_full[kb], 0, h, kt * kBlockN, b);
      }
      // the waits are taken in the order in which they clear: ring slots of the item's first two tiles (freed by tiles
      // n-3, n-2 of the previous item), then V (freed when the previous item's last dP^T was issued, during its tile
      // n-1), then the remaining ring slots
      const int pre = n_qt < 2 ? n_qt : 2;
      for (int i = 0; i < pre; ++i) load_qdo(i, h, b);
      mbar_wait(&sb->v_empty, (it & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&sb->v_full, kTileBytes);
        tma_load_4d(sV, &map_v, &sb->v_full, 0, h, kt * kBlockN, b);
      }
      for (int i = pre; i < n_qt; ++i) load_qdo(i, h, b);
    }
  } else if (warp == kBwdMmaWarp || warp == kBwdMmaWarp + 1) {
    // ============================== MMA stream of query half g ==============================
    const int g = warp - kBwdMmaWarp;
    constexpr uint32_t idesc_nt = make_idesc_bf16(kBlockN, 64, 0, 0);        // S^T_g, dP^T_g: A and B K-major, N = 64
    constexpr uint32_t idesc_ts = make_idesc_bf16(kBlockN, kHeadDim, 0, 1);  // dV, dK: A in TMEM, B MN-major
    const uint32_t t_sg = tmem + kColBS + g * 64, t_dpg = tmem + kColBdP + g * 64;
    TL_DECL((lane == 0 && g == 0) ? 0 : -1);
    // query half g of a 128-row tile = rows 64g .. 64g+63 = byte offset 64 * 128 (a multiple of the 1024-byte swizzle atom)
    auto q_desc = [&](int c) { return make_smem_desc_sw128(smem_u32(sQ + (c % kBwdStages) * kTileBytes + g * 8192)); };
    auto do_desc = [&](int c) { return make_smem_desc_sw128(smem_u32(sdO + (c % kBwdStages) * kTileBytes + g * 8192)); };
    const uint64_t dv = make_smem_desc_sw128(smem_u32(sV));
    auto issue_s = [&](int c, int it) {  // S^T_g(c) = K(it) Q_g(c)^T
      const uint64_t dk = make_smem_desc_sw128(smem_u32(sK + (it & 1) * kTileBytes)), dq = q_desc(c);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk) mma_ss(t_sg, dk + uint64_t(kk * 2), dq + uint64_t(kk * 2), idesc_nt, kk > 0);
        tc_commit(&sb->s_full[g]);
      }
      __syncwarp();
    };
    auto issue_dp = [&](int c, bool last_of_item) {  // dP^T_g(c) = V dO_g(c)^T
      const uint64_t ddo = do_desc(c);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk) mma_ss(t_dpg, dv + uint64_t(kk * 2), ddo + uint64_t(kk * 2), idesc_nt, kk > 0);
        tc_commit(&sb->dp_full[g]);
        if (last_of_item) tc_commit(&sb->v_empty);  // this stream's last read of V_j
      }
      __syncwarp();
    };
    if (my_items > 0) {
      mbar_wait(&sb->k_full[0], 0);
      mbar_wait(&sb->v_full, 0);
      mbar_wait(&sb->qdo_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_dp(0, n_qt == 1);
    }
    int c = 0;
    for (int it = 0; it < my_items; ++it) {
      for (int i = 0; i < n_qt; ++i, ++c) {
        const uint32_t par = c & 1;
        const bool last = i == n_qt - 1;
        const bool more = c + 1 < total_tiles;
        const int it_next = last ? it + 1 : it;            // item of tile c + 1
        const bool next_last = last ? n_qt == 1 : i + 2 == n_qt;  // tile c + 1 is the last tile of its item
        const uint64_t dq_c = q_desc(c), ddo_c = do_desc(c);
        // ---- phase 1 of half g done: dV += P^T_g dO_g, then S^T_g(c+1) over the same columns (in-order tensor pipe)
        TL(10);
        mbar_wait(&sb->p_ready[g], par);
        if (more) {
          mbar_wait(&sb->qdo_full[(c + 1) % kBwdStages], ((c + 1) / kBwdStages) & 1);
          if (last) mbar_wait(&sb->k_full[it_next & 1], (it_next >> 1) & 1);
        }
        if (i == 0 && g == 1) mbar_wait(&sb->dv_init, it & 1);  // half 0's overwriting dV MMA of this item has executed
        TL(11);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_ts(tmem + kColBdV, t_sg + (kk >> 1) * 32 + (kk & 1) * 8, ddo_c + uint64_t(kk * 128), idesc_ts,
                   (i > 0 || g > 0 || kk > 0) ? 1u : 0u);
          if (i == 0 && g == 0) tc_commit(&sb->dv_init);
          if (last) tc_commit(&sb->dv_final);
        }
        __syncwarp();
        if (more) issue_s(c + 1, it_next);
        // ---- phase 2 of half g done: dK += dS^T_g Q_g, then dP^T_g(c+1) over the same columns
        TL(12);
        mbar_wait(&sb->ds_ready[g], par);
        if (i == 0 && g == 1) mbar_wait(&sb->dk_init, it & 1);
        if (i == 0 && g == 0 && it > 0) mbar_wait(&sb->dk_read, (it - 1) & 1);  // half 1's warps have read dK of item it-1
        if (more && last) mbar_wait(&sb->v_full, it_next & 1);
        TL(13);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_ts(tmem + kColBdK, t_dpg + (kk >> 1) * 32 + (kk & 1) * 8, dq_c + uint64_t(kk * 128), idesc_ts,
                   (i > 0 || g > 0 || kk > 0) ? 1u : 0u);
          if (i == 0 && g == 0) tc_commit(&sb->dk_init);
          tc_commit(&sb->qdo_empty[c % kBwdStages]);  // this half's last read of (Q, dO) of tile c
          if (last) {
            tc_commit(&sb->dk_final);
            tc_commit(&sb->k_empty[it & 1]);  // this stream's last read of K_j
          }
        }
        __syncwarp();
        if (more) issue_dp(c + 1, next_last);
        TL(14);
      }
    }
    TL_END();
  } else if (warp == kBwdMmaWarp + 2) {
    // ============================== dQ stream: dQ(c) = dS K_j once both halves of dS^T(c) are in smem ====================
    constexpr uint32_t idesc_tn = make_idesc_bf16(kBlockM, kHeadDim, 1, 1);  // A and B MN-major
    int c = 0;
    for (int it = 0; it < my_items; ++it) {
      const uint64_t dK_d = make_smem_desc_sw128(smem_u32(sK + (it & 1) * kTileBytes));
      mbar_wait(&sb->k_full[it & 1], (it >> 1) & 1);
      for (int i = 0; i < n_qt; ++i, ++c) {
        const uint32_t par = c & 1;
        // A = [half 0 panel of buffer par | half 1 panel]: LBO = distance between the two 64-query panels
        const uint64_t dS_mn = make_smem_desc_sw128_mn(smem_u32(sdS + par * kPanelBytes), (2 - par) * kPanelBytes);
        mbar_wait(&sb->ds_ready[0], par);
        mbar_wait(&sb->ds_ready[1], par);
        if (c > 0) mbar_wait(&sb->dq_empty, (c - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kBlockN / 16; ++kk)  // contraction over 128 keys: 16 key rows = 2048 bytes in both operands
            mma_ss(tmem + kColBdQ, dS_mn + uint64_t(kk * 128), dK_d + uint64_t(kk * 128), idesc_tn, kk > 0);
          tc_commit(&sb->dq_full);
          tc_commit(&sb->ds_free[par]);
          if (i == n_qt - 1) tc_commit(&sb->k_empty[it & 1]);
        }
        __syncwarp();
      }
    }
  } else if (warp < kBwdSoftmaxWarps) {
    // ============================== P^T / dS^T producers ==============================
    // Query half g (64 columns) is owned by two warpgroups: `sub` selects a 32-column slice, warp & 3 the lane quadrant.
    // (A lone warp per SM sub-partition issues MUFU at about half the unit's rate; two co-resident warps saturate it.)
    const int g = warp >> 3, sub = (warp >> 2) & 1;
    const uint32_t lane_base = uint32_t((warp & 3) * 32);
    const int r = int(lane_base) + lane;  // key row inside the tile
    // S^T / dP^T columns of this slice; the bf16 pairs (16 columns) are written over the slice's own first 16 columns
    const uint32_t t_s = tmem + (lane_base << 16) + kColBS + g * 64 + sub * 32;
    const uint32_t t_dp = tmem + (lane_base << 16) + kColBdP + g * 64 + sub * 32;
    const uint32_t stats0 = smem_u32(sb->stats[0][0]) + (g * 64 + sub * 32) * 4;  // this slice's 32 queries of stage 0's lse row
    const uint32_t ds_base = smem_u32(sdS + g * 2 * kPanelBytes + r * 128);  // half 0: buffers 0, 1; half 1: the third panel
#ifndef AGA_BWD_NO_TOKEN
    if (g == 1 && total_tiles > 0) named_bar_arrive(4, 512);  // half 0 may run the first phase 1
#endif
    const bool issuer = (warp & 7) == 0 && lane == 0;  // issues this half's dV_j / dK_j tile stores
    TL_DECL(((warp & 7) == 0 && lane == 0) ? 1 + g : -1);
    int c = 0;
    for (int it = 0; it < my_items; ++it) {
      // keys past the effective length (kv_len inside a static Tk) keep P^T = 0: this lane's key is masked for the item
      uint32_t keep = 0xffffffffu;
      int kt_i = 0, h_i = 0, b_i = 0;
      if (a.kv_len || a.causal) decode(it, kt_i, h_i, b_i);
      if (a.kv_len) keep = (kt_i * kBlockN + r < Tk) ? 0xffffffffu : 0u;
      const int key_g = kt_i * kBlockN + r;  // this lane's key (causal / export paths)
      // gradient of the exported logits of this key (one of the columns [lo, hi)): row `query` of (B, H, Tq, W)
      const float* gcol = (a.d_export && key_g >= a.export_lo && key_g < a.export_hi && (a.head_sel == nullptr || a.head_sel[h_i] != 0))
                              ? a.d_export + (int64_t(b_i) * a.H + h_i) * a.Tq * (a.export_hi - a.export_lo) + (key_g - a.export_lo)
                              : nullptr;
      for (int i = 0; i < n_qt; ++i, ++c) {
        const uint32_t par = c & 1;
        TL(20);
        const int stage = c % kBwdStages;
        const uint32_t lse4 = stats0 + stage * (2 * kBlockM * 4), del4 = lse4 + kBlockM * 4;
        mbar_wait(&sb->qdo_full[stage], (c / kBwdStages) & 1);  // the tile's lse / delta rows have landed (long ago)
        TL(26);
        mbar_wait(&sb->s_full[g], par);
        TL(21);
        tc_fence_after();
        uint32_t pk[16];
        {
          uint32_t sv[32];
          tmem_ld32(t_s, sv);
          tmem_wait_ld();
          // ---- phase 1 (MUFU): wait for the other half to leave its phase 1
#ifndef AGA_BWD_NO_TOKEN
          const float sc = named_bar_sync_dep(4 + g, 512, kScaleLog2);
#else
          const float sc = kScaleLog2;
#endif
          TL(22);
          if (c == 0) CTA_MARK(5);
          const float2 sc2 = make_float2(sc, sc);
          float chk = 0.f;
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4) {
            const float4 L = lds128f(lse4 + e4 * 16);
            const float2 x0 = __ffma2_rn(make_float2(__uint_as_float(sv[4 * e4 + 0]), __uint_as_float(sv[4 * e4 + 1])), sc2,
                                         make_float2(-L.x, -L.y));
            const float2 x1 = __ffma2_rn(make_float2(__uint_as_float(sv[4 * e4 + 2]), __uint_as_float(sv[4 * e4 + 3])), sc2,
                                         make_float2(-L.z, -L.w));
            float p0, p1, p2, p3;
            if ((e4 & 3) < kBwdPolyQuads) {  // this share of the exponentials runs on the FMA pipe instead of the MUFU
              const float2 q0 = ex2_poly2(x0), q1 = ex2_poly2(x1);
              p0 = q0.x, p1 = q0.y, p2 = q1.x, p3 = q1.y;
            } else {
              p0 = ex2(x0.x), p1 = ex2(x0.y), p2 = ex2(x1.x), p3 = ex2(x1.y);
            }
            __nv_bfloat162 h0 = __floats2bfloat162_rn(p0, p1), h1 = __floats2bfloat162_rn(p2, p3);
            pk[2 * e4] = *reinterpret_cast<uint32_t*>(&h0);
            pk[2 * e4 + 1] = *reinterpret_cast<uint32_t*>(&h1);
            chk = __uint_as_float(pk[2 * e4] ^ pk[2 * e4 + 1] ^ __float_as_uint(chk));
          }
          // hand the MUFU to the other half (the value threaded through depends on every exponential above)
#ifndef AGA_BWD_NO_TOKEN
          if (!(g == 1 && c == total_tiles - 1)) pk[15] ^= __float_as_uint(named_bar_arrive_dep(5 - g, 512, chk)) ^ __float_as_uint(chk);
#endif
        }
        if (a.kv_len) {
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] &= keep;
        }
        // causal (Tq = Tk): query q sees key k iff k <= q.  Query tiles above the diagonal (i < kt) are walked with P = 0
        // (the decoder's maps are at most 448 x 448: 6 wasted tile pairs of 16), the diagonal tile is masked per element.
        const int q0 = i * kBlockM + g * 64 + sub * 32;  // first query of this warp's 32-column slice
        if (a.causal && q0 < key_g) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const uint32_t lo_ok = (q0 + 2 * e >= key_g) ? 0x0000ffffu : 0u, hi_ok = (q0 + 2 * e + 1 >= key_g) ? 0xffff0000u : 0u;
            pk[e] &= lo_ok | hi_ok;
          }
        }
        tmem_st16(t_s, pk);  // 32 queries as bf16 pairs over the first 16 of this slice's S^T columns
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb->p_ready[g]);
        TL(23);
        // ---- phase 2 (FMA / LSU)
        const uint32_t dsrow = ds_base + (g == 0 ? par * kPanelBytes : 0);
        mbar_wait(&sb->dp_full[g], par);
        // the panel about to be written has been consumed: half 0 (double-buffered) by dQ(c-2), half 1 by dQ(c-1)
        if (g == 0) {
          if (c >= 2) mbar_wait(&sb->ds_free[par], ((c >> 1) - 1) & 1);
        } else if (c >= 1) {
          mbar_wait(&sb->ds_free[(c - 1) & 1], ((c - 1) >> 1) & 1);
        }
        TL(24);
        tc_fence_after();
        uint32_t dd[16];
        {
          uint32_t dv[32];
          tmem_ld32(t_dp, dv);
          tmem_wait_ld();
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4) {
            const float4 D = lds128f(del4 + e4 * 16);
            const uint32_t w0 = pk[2 * e4], w1 = pk[2 * e4 + 1];
            const float2 a0 = __fadd2_rn(make_float2(__uint_as_float(dv[4 * e4 + 0]), __uint_as_float(dv[4 * e4 + 1])),
                                         make_float2(-D.x, -D.y));
            const float2 a1 = __fadd2_rn(make_float2(__uint_as_float(dv[4 * e4 + 2]), __uint_as_float(dv[4 * e4 + 3])),
                                         make_float2(-D.z, -D.w));
            float2 d0 = __fmul2_rn(make_float2(__uint_as_float(w0 << 16), __uint_as_float(w0 & 0xffff0000u)), a0);
            float2 d1 = __fmul2_rn(make_float2(__uint_as_float(w1 << 16), __uint_as_float(w1 & 0xffff0000u)), a1);
            if (gcol != nullptr) {  // (two lanes of key tile 0 only) the exported columns' gradient adds to dS on visible entries
              const int W = a.export_hi - a.export_lo;
              const int q = q0 + 4 * e4;
              const int qmin = a.causal ? key_g : 0;
              if (q >= qmin && q < a.Tq) d0.x += gcol[int64_t(q) * W];
              if (q + 1 >= qmin && q + 1 < a.Tq) d0.y += gcol[int64_t(q + 1) * W];
              if (q + 2 >= qmin && q + 2 < a.Tq) d1.x += gcol[int64_t(q + 2) * W];
              if (q + 3 >= qmin && q + 3 < a.Tq) d1.y += gcol[int64_t(q + 3) * W];
            }
            __nv_bfloat162 h0 = __floats2bfloat162_rn(d0.x, d0.y), h1 = __floats2bfloat162_rn(d1.x, d1.y);
            dd[2 * e4] = *reinterpret_cast<uint32_t*>(&h0);
            dd[2 * e4 + 1] = *reinterpret_cast<uint32_t*>(&h1);
          }
        }
        tmem_st16(t_dp, dd);
        if (i == 0 && it > 0) {  // this panel staged the previous item's dV_j / dK_j: its tile store must have read it
          if (issuer) bulk_wait_group_read0();
          named_bar_sync(6 + g, 256);
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)  // 16-byte chunk sub*4 + q4 = queries 32 sub + 8 q4 .. + 7, XOR-swizzled with (row & 7)
          sts128(dsrow + (((sub * 4 + q4) ^ (r & 7)) * 16), dd[4 * q4], dd[4 * q4 + 1], dd[4 * q4 + 2], dd[4 * q4 + 3]);
        fence_proxy_async_smem();
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb->ds_ready[g]);
        TL(25);
      }
      // ---- item done for this half: half 0's warps hand out dV_j (final after both streams' last dV MMA), half 1's warps
      //      dK_j.  Half 0 reads dV before its next p_ready arrival (which gates the next item's overwriting dV MMA);
      //      half 1 signals dk_read, which gates the next item's overwriting dK MMA.  The rows are staged (bf16, swizzled)
      //      in the dS^T panel this half writes NEXT — free once the dQ GEMM that read it is done, the very wait the next
      //      tile's phase 2 takes — and leave as one TMA tile store per half (keys past Tk are clipped).  Lane-per-row
      //      global stores cost one L1 transaction per lane per instruction and starved the tensor core's operand reads:
      //      the item boundary was 4800 cycles longer than a tile.
      {
        int kt, h, b;
        decode(it, kt, h, b);
        mbar_wait(g == 0 ? &sb->dv_final : &sb->dk_final, it & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem + (lane_base << 16) + (g == 0 ? kColBdV : kColBdK) + sub * 32, v);
        tmem_wait_ld();
        if (g == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sb->dk_read);
        }
        if (g == 0) {
          if (c >= 2) mbar_wait(&sb->ds_free[c & 1], ((c >> 1) - 1) & 1);
        } else if (c >= 1) {
          mbar_wait(&sb->ds_free[(c - 1) & 1], ((c - 1) >> 1) & 1);
        }
        uint8_t* panel = sdS + (g == 0 ? (c & 1) : 2) * kPanelBytes;
        stage_half_row_bf16(smem_u32(panel + r * 128), r, sub, v, g == 0 ? 1.0f : 0.125f);
        fence_proxy_async_smem();
        named_bar_sync(6 + g, 256);
        if (issuer) {
          tma_store_4d(g == 0 ? &map_dv : &map_dk, panel, 0, h, kt * kBlockN, b);
          bulk_commit_group();
        }
      }
    }
    if (issuer) bulk_wait_group_read0();
    CTA_MARK(6);
    TL_END();
  } else if (warp < kBwdTmaWarp) {
    // ============================== dQ drain ==============================
    const int w4 = warp & 3;
    const uint32_t lane_base = uint32_t(w4 * 32);
    const uint32_t t_dq = tmem + (lane_base << 16) + kColBdQ;
    uint8_t* stage = sdQ + w4 * (32 * 128);  // this warp's 32 rows x 32 fp32
    const uint32_t my_row = smem_u32(stage + lane * 128);
    TL_DECL((warp == kBwdDrainWarp0 && lane == 0) ? 3 : -1);
    int c = 0;
    for (int it = 0; it < my_items; ++it) {
      int kt, h, b;
      decode(it, kt, h, b);
      // tile (b, h, i) = 2 halves x (128 rows x 32 floats); this warp owns rows [32 w4, 32 w4 + 32) of each half
      float* gdst = a.dq_accum + (int64_t(b) * a.H + h) * n_qt * (kBlockM * kHeadDim) + lane_base * 32;
      for (int i = 0; i < n_qt; ++i, ++c) {
        TL(30);
        mbar_wait(&sb->dq_full, c & 1);
        TL(31);
        tc_fence_after();
        uint32_t lo[32], hi[32];
        tmem_ld32(t_dq, lo);
        tmem_ld32(t_dq + 32, hi);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb->dq_empty);
        float* gtile = gdst + int64_t(i) * (kBlockM * kHeadDim);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (lane == 0) bulk_wait_read0();  // the previous bulk reduction has finished reading the staging rows
          __syncwarp();
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t* v = half ? hi : lo;
            sts128(my_row + ((e ^ (lane & 7)) * 16), v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
#ifndef AGA_BWD_NO_RED
          if (lane == 0) bulk_reduce_add_f32(gtile + half * (kBlockM * 32), stage, 32 * 128);
#endif
        }
        TL(32);
      }
    }
    if (lane == 0) bulk_wait_read0();  // the staging rows must outlive the last bulk reduction's reads; the adds themselves
                                       // complete asynchronously (kernel completion orders them before the convert kernel)
    TL(33);
    TL_END();
  }
  tc_fence_before();
  __syncthreads();
  CTA_LOG(1);
  if (warp == kBwdMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

// Per-query statistics of the backward, laid out per 128-row query tile so that one 1 KiB bulk copy brings a tile's
// rows into shared memory:  stats[b,h,tile] = { lse[t] * log2(e) : 128 } { delta[t] = sum_c dO[b,t,h,c] * O[b,t,h,c] : 128 },
// zero for rows past Tq.  8 lanes per row, 16-byte loads.
__global__ void __launch_bounds__(256)
attn_bwd_stats_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, int64_t o_sb, int64_t o_st,
                      const float* __restrict__ lse, int B, int H, int Tq, int n_qt, float* __restrict__ stats,
                      float4* __restrict__ dq_acc4, int64_t n_acc4) {
  const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  // the fp32 dQ accumulator is cleared here as well (one launch and one pass less than a separate memset)
  for (int64_t i = gid; i < n_acc4; i += int64_t(gridDim.x) * blockDim.x) dq_acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t row_id = gid >> 3;  // (b, h, padded t)
  const int sub = int(gid & 7);
  const int Tp = n_qt * kBlockM;
  if (row_id >= int64_t(B) * H * Tp) return;
  const int t = int(row_id % Tp), h = int((row_id / Tp) % H), b = int(row_id / (int64_t(Tp) * H));
  float acc = 0.f, l2 = 0.f;
  if (t < Tq) {
    const int64_t off = b * o_sb + int64_t(t) * o_st + h * kHeadDim + sub * 8;
    const uint4 x = *reinterpret_cast<const uint4*>(o + off);
    const uint4 y = *reinterpret_cast<const uint4*>(d_o + off);
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 xf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xs[e]));
      const float2 yf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ys[e]));
      acc = fmaf(xf.x, yf.x, fmaf(xf.y, yf.y, acc));
    }
    l2 = lse[(int64_t(b) * H + h) * Tq + t] * 1.4426950408889634f;
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (sub == 0) {
    float* dst = stats + ((int64_t(b) * H + h) * n_qt + (t >> 7)) * (2 * kBlockM) + (t & 127);
    dst[0] = l2;
    dst[kBlockM] = acc;
  }
}

// dq = bf16(0.125 * dq_accum): un-tiles / un-swizzles the (b, h, q-tile, 128, 64) fp32 accumulator; 8 columns per thread
__global__ void __launch_bounds__(256)
attn_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int64_t q_sb, int64_t q_st,
                           int H, int Tq, int n_qt, int64_t total8) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total8; i += int64_t(gridDim.x) * blockDim.x) {
    const int c8 = int(i & 7);
    const int64_t bth = i >> 3;
    const int h = int(bth % H);
    const int64_t bt = bth / H;
    const int t = int(bt % Tq);
    const int64_t b = bt / Tq;
    // tile (b, h, t >> 7) = 2 column halves x (128 rows x 32 floats); 8 columns = chunks 2 c8 and 2 c8 + 1 of half c8 >> 2
    const float* row = acc + ((b * H + h) * n_qt + (t >> 7)) * int64_t(kBlockM * kHeadDim) + (c8 >> 2) * (kBlockM * 32) +
                       (t & 127) * 32;
    const int ch = (2 * c8) & 7;
    const float4 v0 = *reinterpret_cast<const float4*>(row + ((ch ^ (t & 7)) * 4));
    const float4 v1 = *reinterpret_cast<const float4*>(row + (((ch + 1) ^ (t & 7)) * 4));
    __nv_bfloat162 w0 = __floats2bfloat162_rn(v0.x * 0.125f, v0.y * 0.125f), w1 = __floats2bfloat162_rn(v0.z * 0.125f, v0.w * 0.125f);
    __nv_bfloat162 w2 = __floats2bfloat162_rn(v1.x * 0.125f, v1.y * 0.125f), w3 = __floats2bfloat162_rn(v1.z * 0.125f, v1.w * 0.125f);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&w0);
    u.y = *reinterpret_cast<uint32_t*>(&w1);
    u.z = *reinterpret_cast<uint32_t*>(&w2);
    u.w = *reinterpret_cast<uint32_t*>(&w3);
    *reinterpret_cast<uint4*>(dq + b * q_sb + int64_t(t) * q_st + h * kHeadDim + c8 * 8) = u;
  }
}


// =============================================================================================== backward, Tq <= 128
// Decoder cross attention (Tq = text length <= 128, Tk = 1500): one query tile, so the roles of the persistent kernel
// are swapped — Q, dO and the row statistics stay RESIDENT (per-thread scalars: lane = query row) and the CTA streams a
// chunk of key tiles past them.  dV_j and dK_j are complete after one tile (no accumulation across tiles: they are
// drained and stored as bf16 directly), dQ accumulates in TMEM over the chunk and is added once into the fp32
// accumulator (a few chunk-CTAs per (batch, head) keep 148 SMs busy).  No per-item pipeline restart: with the
// persistent kernel every tile was an item boundary (97 us per call, 120 TFLOP/s).
//   S  = Q K_j^T      cols [0,128)     SS, both K-major, N = 128
//   dP = dO V_j^T     cols [128,256)   SS, both K-major, N = 128
//   softmax warps (lane = query, 32 keys each): P = exp2(S c - lse), dS = P (dP - delta) -> bf16 rows [query][key] in
//   two 64-key swizzled smem panels each
//   dV_j = P^T dO     [256,320)   A = P panels read MN-major (M = keys), B = dO rows (MN-major)
//   dK_j = dS^T Q     [320,384)   A = dS panels read MN-major,           B = Q rows (MN-major)
//   dQ  += dS K_j     [384,448)   A = dS panels read K-major,           B = K_j rows (MN-major)
// Warps: 0-15 softmax (quadrant = warp & 3, key slice = warp >> 2), 16 TMA, 17 MMA, 18-21 drain (dV_j, dK_j per tile,
// dQ at the end).  The drain warps never store to global memory themselves: a lane-per-row store of 16-byte pieces
// costs one L1 transaction per lane per instruction (4000 cycles per tile, and it starves the tensor core's own
// shared-memory operand reads).  Rows go to swizzled staging tiles and leave through the TMA unit (tile stores for
// dV_j / dK_j, which also clips keys past Tk; a bulk reduce-add for dQ).
constexpr int kQrSoftmaxWarps = 16;
constexpr int kQrTmaWarp = 16, kQrMmaWarp = 17, kQrDrainWarp0 = 18;
constexpr int kQrThreads = 22 * 32;
constexpr uint32_t kColQS = 0, kColQdP = 128, kColQdV = 256, kColQdK = 320, kColQdQ = 384;
constexpr int kQrKStages = 3;  // K_j stays until dQ(j) has run, two tiles after its load: a third buffer keeps TMA ahead

struct QrSmem {
  uint64_t qdo_full, k_full[kQrKStages], k_empty[kQrKStages], v_full[2], v_empty[2];
  uint64_t s_full, dp_full, s_free, dp_free, p_ready, pds_free, dvk_full, dvk_free, dq_full;
  uint32_t tmem_base;
};
// Q, dO | 3 x K, 2 x V | P panels (2) | dS panels (2) | dV, dK staging (bf16 rows; reused as fp32 dQ staging at the end)
constexpr size_t kQrSmemBytes = 1024 + size_t(2 + kQrKStages + 2) * kTileBytes + 4 * size_t(kPanelBytes) + 2 * size_t(kTileBytes) + sizeof(QrSmem);
static_assert(kQrSmemBytes <= 227 * 1024, "q-resident backward kernel exceeds the shared-memory limit");

struct QrArgs {
  int B, H, Tq, Tk;
  int tiles_per_cta;  // key tiles per chunk
  int causal;                 // Tq == Tk <= 128: one key tile, keys above the diagonal masked
  int export_lo, export_hi;   // columns [lo, hi) whose scaled logits were exported; d_export is their gradient
  const float* d_export;      // (B, H, Tq, hi - lo) fp32 or nullptr
  const uint8_t* head_sel;    // (H) or nullptr
  const float* stats;  // (B, H, 1, 2, 128): lse * log2(e) | delta, zero past Tq
  float* dq_accum;     // (B, H, 1, 2, 128, 32) fp32, zero-initialised, chunk-swizzled like the persistent kernel's
  const int32_t* kv_len;  // as BwdArgs::kv_len
  const float* g_pattern; // guided loss fused into the forward epilogue: (B, Tq, 2) pattern,
  const float* g_dpart;   //   (B, H, 4, 2) gradient of the per-row-group partial sums, or nullptr
  int g_early;
};

__global__ void __launch_bounds__(kQrThreads, 1)
attn_bwd_tc_qres_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                        const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                        const __grid_constant__ CUtensorMap map_dk, const __grid_constant__ CUtensorMap map_dv,
                        const QrArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + kTileBytes;
  uint8_t* sK = sdO + kTileBytes;       // kQrKStages stages
  uint8_t* sV = sK + kQrKStages * kTileBytes;  // 2 stages
  uint8_t* sP = sV + 2 * kTileBytes;    // 2 panels: keys [0,64), [64,128)
  uint8_t* sdS = sP + 2 * kPanelBytes;  // 2 panels
  uint8_t* sStage = sdS + 2 * kPanelBytes;  // dV rows | dK rows (32 rows = 4 KiB per drain warp each)
  QrSmem* sb = reinterpret_cast<QrSmem*>(sStage + 2 * kTileBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int Tk = effective_tk(a.kv_len, a.Tk);
  const int n_kt = (Tk + kBlockN - 1) / kBlockN;
  const int kt0 = blockIdx.x * a.tiles_per_cta;
  const int n_my = min(a.tiles_per_cta, n_kt - kt0);  // >= 1 by construction of the grid, unless kv_len cut the key range
  if (n_my <= 0) return;  // (whole CTA, before any barrier / TMEM allocation: its key tiles do not exist)

  if (threadIdx.x == 0) {
    mbar_init(&sb->qdo_full, 1);
    for (int s = 0; s < kQrKStages; ++s) {
      mbar_init(&sb->k_full[s], 1);
      mbar_init(&sb->k_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sb->v_full[s], 1);
      mbar_init(&sb->v_empty[s], 1);
    }
    mbar_init(&sb->s_full, 1);
    mbar_init(&sb->dp_full, 1);
    mbar_init(&sb->s_free, kQrSoftmaxWarps);
    mbar_init(&sb->dp_free, kQrSoftmaxWarps);
    mbar_init(&sb->p_ready, kQrSoftmaxWarps);
    mbar_init(&sb->pds_free, 1);
    mbar_init(&sb->dvk_full, 1);
    mbar_init(&sb->dvk_free, 4);
    mbar_init(&sb->dq_full, 1);
    fence_barrier_init();
  }
  if (warp == kQrMmaWarp) {
    tmem_alloc(&sb->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  if (warp == kQrTmaWarp && lane == 0) {
    prefetch_tensormap(&map_q);
    prefetch_tensormap(&map_k);
    prefetch_tensormap(&map_v);
    prefetch_tensormap(&map_do);
    prefetch_tensormap(&map_dk);
    prefetch_tensormap(&map_dv);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sb->tmem_base;

  if (warp == kQrTmaWarp) {
    // ============================== TMA producer ==============================
    TL_DECL(lane == 0 ? 3 : -1);
    TL(40);
    if (elect_one()) {
      mbar_arrive_expect_tx(&sb->qdo_full, 2 * kTileBytes);
      tma_load_4d(sQ, &map_q, &sb->qdo_full, 0, h, 0, b);
      tma_load_4d(sdO, &map_do, &sb->qdo_full, 0, h, 0, b);
    }
    for (int jj = 0; jj < n_my; ++jj) {
      const int s = jj & 1, sk = jj % kQrKStages;
      const uint32_t ph = (jj >> 1) & 1;
      mbar_wait(&sb->k_empty[sk], ((jj / kQrKStages) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&sb->k_full[sk], kTileBytes);
        tma_load_4d(sK + sk * kTileBytes, &map_k, &sb->k_full[sk], 0, h, (kt0 + jj) * kBlockN, b);
      }
      mbar_wait(&sb->v_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&sb->v_full[s], kTileBytes);
        tma_load_4d(sV + s * kTileBytes, &map_v, &sb->v_full[s], 0, h, (kt0 + jj) * kBlockN, b);
      }
      TL(42);
    }
    TL_END();
  } else if (warp == kQrMmaWarp) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc_s = make_idesc_bf16(kBlockM, kBlockN, 0, 0);    // S, dP: A and B K-major, N = 128
    constexpr uint32_t idesc_mn = make_idesc_bf16(kBlockN, kHeadDim, 1, 1);  // dV, dK: A and B MN-major
    constexpr uint32_t idesc_dq = make_idesc_bf16(kBlockM, kHeadDim, 0, 1);  // dQ: A K-major, B MN-major
    const uint64_t dQd = make_smem_desc_sw128(smem_u32(sQ)), ddO = make_smem_desc_sw128(smem_u32(sdO));
    const uint64_t aP = make_smem_desc_sw128_mn(smem_u32(sP), kPanelBytes);
    const uint64_t adS = make_smem_desc_sw128_mn(smem_u32(sdS), kPanelBytes);
    auto issue_s = [&](int jj) {
      const uint64_t dk = make_smem_desc_sw128(smem_u32(sK + (jj % kQrKStages) * kTileBytes));
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk) mma_ss(tmem + kColQS, dQd + uint64_t(kk * 2), dk + uint64_t(kk * 2), idesc_s, kk > 0);
        tc_commit(&sb->s_full);
      }
      __syncwarp();
    };
    auto issue_dp = [&](int jj) {
      const uint64_t dv = make_smem_desc_sw128(smem_u32(sV + (jj & 1) * kTileBytes));
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk) mma_ss(tmem + kColQdP, ddO + uint64_t(kk * 2), dv + uint64_t(kk * 2), idesc_s, kk > 0);
        tc_commit(&sb->dp_full);
        tc_commit(&sb->v_empty[jj & 1]);
      }
      __syncwarp();
    };
    TL_DECL(lane == 0 ? 0 : -1);
    TL(9);
    mbar_wait(&sb->qdo_full, 0);
    mbar_wait(&sb->k_full[0], 0);
    tc_fence_after();
    issue_s(0);
    mbar_wait(&sb->v_full[0], 0);
    tc_fence_after();
    issue_dp(0);
    for (int jj = 0; jj < n_my; ++jj) {
      const uint32_t par = jj & 1;
      TL(10);
      if (jj + 1 < n_my) {
        const uint32_t ph1 = ((jj + 1) >> 1) & 1;
        mbar_wait(&sb->k_full[(jj + 1) % kQrKStages], ((jj + 1) / kQrKStages) & 1);
        mbar_wait(&sb->s_free, par);
        tc_fence_after();
        issue_s(jj + 1);
        TL(11);
        mbar_wait(&sb->v_full[(jj + 1) & 1], ph1);
        mbar_wait(&sb->dp_free, par);
        tc_fence_after();
        issue_dp(jj + 1);
        TL(12);
      }
      mbar_wait(&sb->p_ready, par);
      if (jj > 0) mbar_wait(&sb->dvk_free, (jj - 1) & 1);
      TL(13);
      tc_fence_after();
      const uint64_t dkm = make_smem_desc_sw128(smem_u32(sK + (jj % kQrKStages) * kTileBytes));
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kBlockM / 16; ++kk)  // contraction over the 128 queries: 16 rows = 2048 bytes in both operands
          mma_ss(tmem + kColQdV, aP + uint64_t(kk * 128), ddO + uint64_t(kk * 128), idesc_mn, kk > 0);
#pragma unroll
        for (int kk = 0; kk < kBlockM / 16; ++kk)
          mma_ss(tmem + kColQdK, adS + uint64_t(kk * 128), dQd + uint64_t(kk * 128), idesc_mn, kk > 0);
        tc_commit(&sb->dvk_full);
#pragma unroll
        for (int kk = 0; kk < kBlockN / 16; ++kk) {  // contraction over the 128 keys: 4 k-steps per 64-key panel
          const uint64_t ads_k = make_smem_desc_sw128(smem_u32(sdS + (kk >> 2) * kPanelBytes)) + uint64_t((kk & 3) * 2);
          mma_ss(tmem + kColQdQ, ads_k, dkm + uint64_t(kk * 128), idesc_dq, (jj > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(&sb->pds_free);
        tc_commit(&sb->k_empty[jj % kQrKStages]);
        if (jj == n_my - 1) tc_commit(&sb->dq_full);
      }
      __syncwarp();
      TL(14);
    }
    TL_END();
  } else if (warp < kQrSoftmaxWarps) {
    // ============================== P / dS producers ==============================
    const int quad = warp & 3, cs = warp >> 2;
    const uint32_t lane_base = uint32_t(quad * 32);
    const int row = int(lane_base) + lane;  // query row
    const float* st = a.stats + (int64_t(b) * a.H + h) * (2 * kBlockM);
    const float neg_lse = -st[row], delta = st[kBlockM + row];
    const uint32_t t_s = tmem + (lane_base << 16) + kColQS + cs * 32;
    const uint32_t t_dp = tmem + (lane_base << 16) + kColQdP + cs * 32;
    const uint32_t p_row = smem_u32(sP + (cs >> 1) * kPanelBytes + row * 128);
    const uint32_t ds_row = smem_u32(sdS + (cs >> 1) * kPanelBytes + row * 128);
    const float2 sc2 = make_float2(kScaleLog2, kScaleLog2), nl2 = make_float2(neg_lse, neg_lse);
    const float2 nd2 = make_float2(-delta, -delta);
    const int W = a.export_hi - a.export_lo;
    const float* grow = (a.d_export && (a.head_sel == nullptr || a.head_sel[h] != 0) && row < a.Tq)
                            ? a.d_export + ((int64_t(b) * a.H + h) * a.Tq + row) * W - a.export_lo : nullptr;
    TL_DECL((warp == 0 && lane == 0) ? 1 : -1);
    TL(19);
    if (int(lane_base) >= a.Tq) {
      // all 32 query rows of this warp lie past Tq (Tq = 64: half of the softmax warps): their P / dS rows are zero —
      // written once — and only the barrier protocol is kept going (no TMEM loads, no exponentials, no smem traffic)
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint32_t off = uint32_t((((cs & 1) * 4 + q4) ^ (row & 7)) * 16);
        sts128(p_row + off, 0u, 0u, 0u, 0u);
        sts128(ds_row + off, 0u, 0u, 0u, 0u);
      }
      fence_proxy_async_smem();
      __syncwarp();
      for (int jj = 0; jj < n_my; ++jj) {
        const uint32_t par = jj & 1;
        mbar_wait(&sb->s_full, par);
        if (lane == 0) mbar_arrive(&sb->s_free);
        mbar_wait(&sb->dp_full, par);
        if (lane == 0) mbar_arrive(&sb->dp_free);
        // one arrival per phase: the live warps may still be writing tile jj-1 when dP(jj) lands
        if (jj > 0) mbar_wait(&sb->p_ready, (jj - 1) & 1);
        if (lane == 0) mbar_arrive(&sb->p_ready);
      }
    } else
    for (int jj = 0; jj < n_my; ++jj) {
      const uint32_t par = jj & 1;
      uint32_t pk[16], dd[16];
      float gg1 = 0.f, gg2 = 0.f;  // fused guided loss: gradient on the scaled logits of keys 1, 2 of this row
      TL(20);
      mbar_wait(&sb->s_full, par);
      TL(21);
      tc_fence_after();
      {
        uint32_t sv[32];
        tmem_ld32(t_s, sv);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb->s_free);
        if (a.causal || a.kv_len) {  // key > query (causal) or key >= kv_len: -inf -> P = 0
          const int kbase = (kt0 + jj) * kBlockN + cs * 32;
          int vis = a.causal ? row - kbase + 1 : 32;  // keys of this 32-column slice the row may see
          if (a.kv_len) vis = min(vis, Tk - kbase);
          if (vis < 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= vis) sv[i] = 0xff800000u;
          }
        }
        if (a.g_dpart != nullptr && kt0 + jj == 0 && cs == 0 && row < a.Tq) {
          // d loss / d S~[row, 1..2] of the fused guided loss: 2 (S~ - c) * d(sum of the row's group)
          const float2 pt = *reinterpret_cast<const float2*>(a.g_pattern + (int64_t(b) * a.Tq + row) * 2);
          float u1, u2;
          guided_row(__uint_as_float(sv[1]) * 0.125f, __uint_as_float(sv[2]) * 0.125f, pt, a.g_early, u1, u2);
          const float gsum = 2.0f * a.g_dpart[((int64_t(b) * a.H + h) * 4 + (row >> 5)) * 2];
          gg1 = gsum * u1;
          gg2 = gsum * u2;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), sc2, nl2);
          __nv_bfloat162 hb = __floats2bfloat162_rn(ex2(x.x), ex2(x.y));
          pk[i] = *reinterpret_cast<uint32_t*>(&hb);
        }
      }
      TL(23);
      mbar_wait(&sb->dp_full, par);
      tc_fence_after();
      {
        uint32_t dv[32];
        tmem_ld32(t_dp, dv);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb->dp_free);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 t = __fadd2_rn(make_float2(__uint_as_float(dv[2 * i]), __uint_as_float(dv[2 * i + 1])), nd2);
          float2 d = __fmul2_rn(make_float2(__uint_as_float(pk[i] << 16), __uint_as_float(pk[i] & 0xffff0000u)), t);
          if (grow != nullptr) {  // gradient of the exported (scaled, masked) logits adds to dS on the visible entries
            const int key = (kt0 + jj) * kBlockN + cs * 32 + 2 * i;
            const int lim = a.causal ? min(a.export_hi, row + 1) : a.export_hi;
            if (key >= a.export_lo && key < lim) d.x += grow[key];
            if (key + 1 >= a.export_lo && key + 1 < lim) d.y += grow[key + 1];
          }
          if (i == 0) d.y += gg1;  // key 1  (gg1 / gg2 are zero unless this warp holds keys 0..31 of key tile 0)
          if (i == 1) d.x += gg2;  // key 2
          __nv_bfloat162 hb = __floats2bfloat162_rn(d.x, d.y);
          dd[i] = *reinterpret_cast<uint32_t*>(&hb);
        }
      }
      TL(25);
      if (jj > 0) mbar_wait(&sb->pds_free, (jj - 1) & 1);  // the previous tile's GEMMs have read the panels
      TL(26);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {  // 16-byte chunk (cs & 1) * 4 + q4 of the 128-byte row, XOR-swizzled with (row & 7)
        const uint32_t off = uint32_t((((cs & 1) * 4 + q4) ^ (row & 7)) * 16);
        sts128(p_row + off, pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
        sts128(ds_row + off, dd[4 * q4], dd[4 * q4 + 1], dd[4 * q4 + 2], dd[4 * q4 + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb->p_ready);
      TL(27);
    }
    TL_END();
  } else if (warp >= kQrDrainWarp0) {
    // ============================== drain: dV_j, dK_j per tile; dQ at the end ==============================
    const int w4 = warp & 3;
    const uint32_t lane_base = uint32_t(w4 * 32);
    const int r = int(lane_base) + lane;
    const uint32_t t_base = tmem + (lane_base << 16);
    uint8_t* st_dv = sStage + w4 * (32 * 128);               // this warp's 32 rows of the dV staging tile
    uint8_t* st_dk = sStage + kTileBytes + w4 * (32 * 128);
    const uint32_t row_dv = smem_u32(st_dv + lane * 128), row_dk = smem_u32(st_dk + lane * 128);
    TL_DECL((warp == kQrDrainWarp0 && lane == 0) ? 2 : -1);
    for (int jj = 0; jj < n_my; ++jj) {
      TL(30);
      const int key0 = (kt0 + jj) * kBlockN + int(lane_base);
      mbar_wait(&sb->dvk_full, jj & 1);
      TL(31);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(t_base + kColQdV, v0);
      tmem_ld32(t_base + kColQdV + 32, v1);
      tmem_wait_ld();
      if (lane == 0) bulk_wait_group_read0();  // the previous tile's stores have read the staging rows
      __syncwarp();
      stage_row_bf16(row_dv, r, v0, v1, 1.0f);
      tmem_ld32(t_base + kColQdK, v0);
      tmem_ld32(t_base + kColQdK + 32, v1);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb->dvk_free);
      stage_row_bf16(row_dk, r, v0, v1, 0.125f);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&map_dv, st_dv, 0, h, key0, b);
        tma_store_4d(&map_dk, st_dk, 0, h, key0, b);
        bulk_commit_group();
      }
      TL(32);
    }
    // dQ: this warp's 32 rows of the (b, h) tile, two 32-column fp32 halves, staged (16-byte chunks XOR-swizzled with
    // the row, the accumulator's own layout) and added with one bulk reduction each
    mbar_wait(&sb->dq_full, 0);
    TL(33);
    tc_fence_after();
    float* gtile = a.dq_accum + (int64_t(b) * a.H + h) * (kBlockM * kHeadDim) + lane_base * 32;
    if (lane == 0) bulk_wait_group_read0();
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      tmem_ld32(t_base + kColQdQ + half * 32, v);
      tmem_wait_ld();
      const uint32_t my_row = half ? row_dk : row_dv;
#pragma unroll
      for (int e = 0; e < 8; ++e) sts128(my_row + ((e ^ (lane & 7)) * 16), v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) bulk_reduce_add_f32(gtile + half * (kBlockM * 32), half ? st_dk : st_dv, 32 * 128);
    }
    if (lane == 0) bulk_wait_group_read0();
    TL(34);
    TL_END();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kQrMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace

// causal masking and the gradient of the exported columns exist in the one-query-tile kernel only (Tq <= 128: the
// decoder self attention of a training step); longer causal sequences take the CUDA-core path
bool attn_tc_bwd_supported(const aga_attn_params& p) {
  if (!attn_tc_supported(p)) return false;
#ifdef AGA_BWD_NO_QRES
  if (p.causal || p.export_kind != AGA_EXPORT_NONE) return false;
#endif
  return true;
}

size_t attn_tc_bwd_workspace(const aga_attn_params& p) {
  const size_t n_qt = size_t(p.Tq + kBlockM - 1) / kBlockM;
  const size_t stats = align_up(size_t(p.B) * p.H * n_qt * 2 * kBlockM * sizeof(float), 256);
  const size_t dq = align_up(size_t(p.B) * p.H * n_qt * kBlockM * kHeadDim * sizeof(float), 256);
  return stats + dq;
}

int attn_tc_bwd(const aga_attn_bwd_params& bp, void* ws, cudaStream_t s) {
  const aga_attn_params& p = bp.fwd;
  const int n_qt = (p.Tq + kBlockM - 1) / kBlockM;
  float* stats = static_cast<float*>(ws);
  float* dq_acc = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) +
                                           align_up(size_t(p.B) * p.H * n_qt * 2 * kBlockM * sizeof(float), 256));
  const size_t dq_bytes = size_t(p.B) * p.H * n_qt * kBlockM * kHeadDim * sizeof(float);
  const int64_t stat_threads = int64_t(p.B) * p.H * n_qt * kBlockM * 8;
  attn_bwd_stats_kernel<<<unsigned((stat_threads + 255) / 256), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(p.out), static_cast<const __nv_bfloat16*>(bp.dout), p.o_stride_b, p.o_stride_t,
      p.lse, p.B, p.H, p.Tq, n_qt, stats, reinterpret_cast<float4*>(dq_acc), int64_t(dq_bytes / 16));
  AGA_AFTER_LAUNCH();
  CUtensorMap mq, mk, mv, mdo;
  int st;
  if ((st = make_map(&mq, p.q, p.B, p.H, p.Tq, p.q_stride_b, p.q_stride_t, kBlockM)) != AGA_OK) return st;
  if ((st = make_map(&mk, p.k, p.B, p.H, p.Tk, p.k_stride_b, p.k_stride_t, kBlockN)) != AGA_OK) return st;
  if ((st = make_map(&mv, p.v, p.B, p.H, p.Tk, p.v_stride_b, p.v_stride_t, kBlockN)) != AGA_OK) return st;
  if ((st = make_map(&mdo, bp.dout, p.B, p.H, p.Tq, p.o_stride_b, p.o_stride_t, kBlockM)) != AGA_OK) return st;
  const int n_kt = (p.Tk + kBlockN - 1) / kBlockN;
  static const int n_sm = []() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
  }();
#ifndef AGA_BWD_NO_QRES
  if (n_qt == 1) {
    // one query tile (decoder cross attention): Q-resident kernel, ~4 waves of chunk-CTAs
    const int64_t tiles = int64_t(p.B) * p.H * n_kt;
    int per = int((tiles + int64_t(n_sm) * 4 - 1) / (int64_t(n_sm) * 4));
    per = std::max(1, std::min(per, n_kt));
    const int n_chunks = (n_kt + per - 1) / per;
    const bool dexp = p.export_kind == AGA_EXPORT_LOGITS && bp.d_export != nullptr;
    QrArgs qa{p.B, p.H, p.Tq, p.Tk, per, p.causal, dexp ? p.export_lo : 0, dexp ? p.export_hi : 0,
              dexp ? bp.d_export : nullptr, dexp ? p.head_sel : nullptr, stats, dq_acc, p.kv_len,
              bp.d_guided_part ? p.guided_pattern : nullptr, bp.d_guided_part, p.guided_early};
    CUtensorMap mdk, mdv;  // 32-row boxes: one store per drain warp
    if ((st = make_map(&mdk, bp.dk, p.B, p.H, p.Tk, p.k_stride_b, p.k_stride_t, 32)) != AGA_OK) return st;
    if ((st = make_map(&mdv, bp.dv, p.B, p.H, p.Tk, p.v_stride_b, p.v_stride_t, 32)) != AGA_OK) return st;
    AGA_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_tc_qres_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kQrSmemBytes)));
    attn_bwd_tc_qres_kernel<<<dim3(n_chunks, p.H, p.B), kQrThreads, kQrSmemBytes, s>>>(mq, mk, mv, mdo, mdk, mdv, qa);
    AGA_AFTER_LAUNCH();
  } else
#endif
  {
  const int n_items = p.B * p.H * n_kt;
  const bool dexp_p = p.export_kind == AGA_EXPORT_LOGITS && bp.d_export != nullptr;
  BwdArgs a{p.B, p.H, p.Tq, p.Tk, n_items, p.k_stride_b, p.k_stride_t, p.v_stride_b, p.v_stride_t, stats, dq_acc,
            static_cast<__nv_bfloat16*>(bp.dk), static_cast<__nv_bfloat16*>(bp.dv), p.kv_len, p.causal,
            dexp_p ? p.export_lo : 0, dexp_p ? p.export_hi : 0, dexp_p ? bp.d_export : nullptr, dexp_p ? p.head_sel : nullptr};
  AGA_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kBwdSmemBytes)));
  const unsigned grid = unsigned(std::min(n_items, n_sm));  // persistent: one CTA per SM walks the items
  CUtensorMap mdk, mdv;
  if ((st = make_map(&mdk, bp.dk, p.B, p.H, p.Tk, p.k_stride_b, p.k_stride_t, kBlockN)) != AGA_OK) return st;
  if ((st = make_map(&mdv, bp.dv, p.B, p.H, p.Tk, p.v_stride_b, p.v_stride_t, kBlockN)) != AGA_OK) return st;
  attn_bwd_tc_kernel<<<grid, kBwdThreads, kBwdSmemBytes, s>>>(mq, mk, mv, mdo, mdk, mdv, a);
  AGA_AFTER_LAUNCH();
  }
  const int64_t total8 = int64_t(p.B) * p.Tq * p.H * (kHeadDim / 8);
  const unsigned gx = unsigned(std::min<int64_t>((total8 + 255) / 256, 148 * 16));
  attn_bwd_dq_convert_kernel<<<gx ? gx : 1, 256, 0, s>>>(dq_acc, static_cast<__nv_bfloat16*>(bp.dq), p.q_stride_b,
                                                          p.q_stride_t, p.H, p.Tq, n_qt, total8);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

}  // namespace aga
