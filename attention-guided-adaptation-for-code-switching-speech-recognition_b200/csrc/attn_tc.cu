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
#else
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
constexpr int kNumSoftmaxWarps = 8;
constexpr int kTmaWarp = 8;
constexpr int kMmaWarp = 9;   // warps 9 and 10: one MMA-issuing warp per query tile
constexpr int kThreads = 352;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS = 0, kColO = 256, kColP = 384;
constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;  // d^-1/2 * log2(e)
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;

struct FwdSmem {
  // barriers
  uint64_t q_full;
  uint64_t k_full[kStages], k_empty[kStages], v_full[kStages], v_empty[kStages];
  uint64_t s_full[2], s_free[2], p_ready[2], pv_done[2];
  uint32_t tmem_base;
};
constexpr size_t kFwdSmemBytes = 1024 /*align slack*/ + size_t(2 + 2 * kStages) * kTileBytes + sizeof(FwdSmem);

struct FwdArgs {
  int B, H, Tq, Tk;
  int64_t o_sb, o_st;
  __nv_bfloat16* out;
  float* lse;
};

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                               // 2 tiles
  uint8_t* sK = sQ + 2 * kTileBytes;                // kStages tiles
  uint8_t* sV = sK + kStages * kTileBytes;          // kStages tiles
  FwdSmem* sb = reinterpret_cast<FwdSmem*>(sV + kStages * kTileBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row0 = qt * 2 * kBlockM;
  const int n_kt = (a.Tk + kBlockN - 1) / kBlockN;
  const bool active_b = row0 + kBlockM < a.Tq;  // tile B holds at least one valid row

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
    tmem_alloc(&sb->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  if (warp == kTmaWarp && lane == 0) {
    prefetch_tensormap(&map_q);
    prefetch_tensormap(&map_k);
    prefetch_tensormap(&map_v);
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
  } else if (warp == kMmaWarp || warp == kMmaWarp + 1) {
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

#ifndef AGA_FWD_STAGGER
#define AGA_FWD_STAGGER 1200
#endif
      if (t == 1 && n_kt > 2) {
        // The two warpgroups share each SM sub-partition's MUFU.  Starting tile B half a period late puts its
        // exp2 phase under tile A's load / max / wait phases (and vice versa) instead of on top of A's exp2 phase.
        const long long t0 = clock64();
        while (clock64() - t0 < AGA_FWD_STAGGER) {
        }
      }
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
        const int valid = a.Tk - j * kBlockN;  // keys of this tile that exist (>= 128 except on the last tile)
        if (valid < kBlockN) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= valid) sr[c][i] = 0xff800000u;  // -inf: exp2 -> 0
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
        // ---- P = exp2(S*c - m) -> bf16 pairs -> its own TMEM columns
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = ex2(fmaf(__uint_as_float(sr[c][2 * i]), kScaleLog2, -m_used));
            const float p1 = ex2(fmaf(__uint_as_float(sr[c][2 * i + 1]), kScaleLog2, -m_used));
            rs0 += p0;
            rs1 += p1;
            __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
            pk[i] = *reinterpret_cast<uint32_t*>(&hb);
          }
          tmem_st16(t_p + c * 16, pk);
        }
        l += rs0 + rs1;
        TL(25);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb->p_ready[t]);
        TL(26);
      }
      TL_END();
      // ---- epilogue: O / l -> bf16 -> global ; lse
      mbar_wait(&sb->pv_done[t], (n_kt - 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / l;
      __nv_bfloat16* orow = a.out + int64_t(b) * a.o_sb + int64_t(row) * a.o_st + h * kHeadDim;
#pragma unroll 1
      for (int c = 0; c < kHeadDim / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(t_o + c * 32, r);
        tmem_wait_ld();
        if (row < a.Tq) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(r[8 * i + 2 * e]) * inv,
                                                        __uint_as_float(r[8 * i + 2 * e + 1]) * inv);
              w[e] = *reinterpret_cast<uint32_t*>(&hb);
            }
            *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      if (row < a.Tq && a.lse) a.lse[(int64_t(b) * a.H + h) * a.Tq + row] = (m_used + log2f(l)) * kLn2;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
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

bool attn_tc_supported(const aga_attn_params& p) {
  if (p.dtype != AGA_BF16 || p.causal || p.export_kind != AGA_EXPORT_NONE) return false;
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
  FwdArgs a{p.B, p.H, p.Tq, p.Tk, p.o_stride_b, p.o_stride_t, static_cast<__nv_bfloat16*>(p.out), p.lse};
  AGA_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kFwdSmemBytes)));
  dim3 grid((p.Tq + 2 * kBlockM - 1) / (2 * kBlockM), p.H, p.B);
  attn_fwd_tc_kernel<<<grid, kThreads, kFwdSmemBytes, s>>>(mq, mk, mv, a);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

namespace {
// =============================================================================================== backward
// One CTA = one 128-key tile (K_j, V_j resident in smem) of one (batch, head); it walks the 128-row query tiles.
// Five tcgen05 GEMMs per (i, j) pair, accumulators in TMEM (448 of 512 columns):
//   S  = Q_i K_j^T          [  0,128)   SS, both K-major
//   dP = dO_i V_j^T         [128,256)   SS, both K-major
//   dV_j += P^T dO_i        [256,320)   A = P  (smem, MN-major: M = keys), B = dO_i (MN-major)
//   dK_j += dS^T Q_i        [320,384)   A = dS (smem, MN-major),           B = Q_i  (MN-major)
//   dQ_i  = dS K_j          [384,448)   A = dS (smem, K-major),            B = K_j  (MN-major)
// The 8 softmax warps (two warpgroups, 64 key columns each) turn S, dP into P = exp2(S c - lse), dS = P (dP - delta)
// (bf16, written to smem in the 128-byte-swizzled UMMA layout); 4 epilogue warps drain dQ_i with vector
// red.global.add into an fp32 accumulator (converted to bf16 and scaled by a tiny kernel afterwards) and, at the
// end, store dK_j, dV_j.  Rows/keys past the tensor ends are zero-filled by TMA, which makes their contributions
// exactly zero — no masking is needed in the non-causal backward.
constexpr int kBwdThreads = 448;
constexpr int kBwdSoftmaxWarps = 8;
constexpr int kBwdDqWarp0 = 8;
constexpr int kBwdTmaWarp = 12;
constexpr int kBwdMmaWarp = 13;
constexpr uint32_t kColBS = 0, kColBdP = 128, kColBdV = 256, kColBdK = 320, kColBdQ = 384;
constexpr int kPanelBytes = kBlockM * 128;  // 128 rows x 64 bf16

struct BwdSmem {
  uint64_t kv_full;
  uint64_t qdo_full[2], qdo_empty[2];
  uint64_t sdp_full, pds_ready, pds_free[2], dq_full, dq_empty;
  uint32_t tmem_base;
};
// K, V | 2 x (Q, dO) | 2 x (P, dS) of 2 panels each  (= 224 KiB: P/dS are double-buffered so that the softmax of
// tile i+1 overlaps the dV/dK/dQ GEMMs of tile i)
constexpr size_t kBwdSmemBytes = 1024 + size_t(2 + 4) * kTileBytes + 8 * size_t(kPanelBytes) + sizeof(BwdSmem);

struct BwdArgs {
  int B, H, Tq, Tk;
  int64_t k_sb, k_st, v_sb, v_st;
  const float* lse;
  const float* delta;
  float* dq_accum;  // (B, Tq, H*64) fp32, zero-initialised
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
};

// MN-major operand spanning two 64-element panels along M (P^T / dS^T as A): LBO = panel stride
__device__ __forceinline__ uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t(lbo_bytes >> 4) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                   const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kTileBytes;
  uint8_t* sQ = sV + kTileBytes;        // 2 stages
  uint8_t* sdO = sQ + 2 * kTileBytes;   // 2 stages
  uint8_t* sP = sdO + 2 * kTileBytes;   // 2 buffers x 2 panels
  uint8_t* sdS = sP + 4 * kPanelBytes;  // 2 buffers x 2 panels
  BwdSmem* sb = reinterpret_cast<BwdSmem*>(sdS + 4 * kPanelBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int key0 = kt * kBlockN;
  const int n_qt = (a.Tq + kBlockM - 1) / kBlockM;

  if (threadIdx.x == 0) {
    mbar_init(&sb->kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sb->qdo_full[s], 1);
      mbar_init(&sb->qdo_empty[s], 1);
    }
    mbar_init(&sb->sdp_full, 1);
    mbar_init(&sb->pds_ready, kBwdSoftmaxWarps);
    mbar_init(&sb->pds_free[0], 1);
    mbar_init(&sb->pds_free[1], 1);
    mbar_init(&sb->dq_full, 1);
    mbar_init(&sb->dq_empty, 4);
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
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sb->tmem_base;

  if (warp == kBwdTmaWarp) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&sb->kv_full, 2 * kTileBytes);
      tma_load_4d(sK, &map_k, &sb->kv_full, 0, h, key0, b);
      tma_load_4d(sV, &map_v, &sb->kv_full, 0, h, key0, b);
    }
    for (int i = 0; i < n_qt; ++i) {
      const int s = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      mbar_wait(&sb->qdo_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&sb->qdo_full[s], 2 * kTileBytes);
        tma_load_4d(sQ + s * kTileBytes, &map_q, &sb->qdo_full[s], 0, h, i * kBlockM, b);
        tma_load_4d(sdO + s * kTileBytes, &map_do, &sb->qdo_full[s], 0, h, i * kBlockM, b);
      }
    }
  } else if (warp == kBwdMmaWarp) {
    {
      constexpr uint32_t idesc_nt = make_idesc_bf16(kBlockM, kBlockN, 0, 0);   // S, dP
      constexpr uint32_t idesc_tn = make_idesc_bf16(kBlockN, kHeadDim, 1, 1);  // dV, dK: A and B MN-major
      constexpr uint32_t idesc_nn = make_idesc_bf16(kBlockM, kHeadDim, 0, 1);  // dQ: A K-major, B MN-major
      const uint64_t dK_k = make_smem_desc_sw128(smem_u32(sK));
      const uint64_t dV_k = make_smem_desc_sw128(smem_u32(sV));
      mbar_wait(&sb->kv_full, 0);
      TL_DECL(lane == 0 ? 0 : -1);
      auto issue_s_dp = [&](int i) {  // S = Q_i K^T, dP = dO_i V^T into TMEM
        const int s = i & 1;
        const uint64_t dQ_s = make_smem_desc_sw128(smem_u32(sQ + s * kTileBytes));
        const uint64_t ddO_s = make_smem_desc_sw128(smem_u32(sdO + s * kTileBytes));
        mbar_wait(&sb->qdo_full[s], (i >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kHeadDim / 16; ++kk)
            mma_ss(tmem + kColBS, dQ_s + uint64_t(kk * 2), dK_k + uint64_t(kk * 2), idesc_nt, kk > 0);
#pragma unroll
          for (int kk = 0; kk < kHeadDim / 16; ++kk)
            mma_ss(tmem + kColBdP, ddO_s + uint64_t(kk * 2), dV_k + uint64_t(kk * 2), idesc_nt, kk > 0);
          tc_commit(&sb->sdp_full);
        }
        __syncwarp();
      };
      issue_s_dp(0);
      for (int i = 0; i < n_qt; ++i) {
        const int s = i & 1;
        const uint64_t dQ_s = make_smem_desc_sw128(smem_u32(sQ + s * kTileBytes));
        const uint64_t ddO_s = make_smem_desc_sw128(smem_u32(sdO + s * kTileBytes));
        const uint8_t* bP = sP + s * 2 * kPanelBytes;
        const uint8_t* bdS = sdS + s * 2 * kPanelBytes;
        const uint64_t dP_mn = make_smem_desc_sw128_mn(smem_u32(bP), kPanelBytes);
        const uint64_t dS_mn = make_smem_desc_sw128_mn(smem_u32(bdS), kPanelBytes);
        const uint64_t dS_k0 = make_smem_desc_sw128(smem_u32(bdS));
        const uint64_t dS_k1 = make_smem_desc_sw128(smem_u32(bdS + kPanelBytes));
        TL(10);
        mbar_wait(&sb->pds_ready, i & 1);  // softmax(i) is done with the S / dP columns and has written P / dS
        TL(13);
        // next tile's S, dP first: its softmax then runs under this tile's dV / dK / dQ GEMMs
        if (i + 1 < n_qt) issue_s_dp(i + 1);
        TL(12);
        if (i > 0) mbar_wait(&sb->dq_empty, (i - 1) & 1);
        TL(14);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kBlockM / 16; ++kk)  // contraction over the 128 query rows: 16 rows = 2048 bytes
            mma_ss(tmem + kColBdV, dP_mn + uint64_t(kk * 128), ddO_s + uint64_t(kk * 128), idesc_tn, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < kBlockM / 16; ++kk)
            mma_ss(tmem + kColBdK, dS_mn + uint64_t(kk * 128), dQ_s + uint64_t(kk * 128), idesc_tn, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < kBlockN / 16; ++kk)  // contraction over the 128 keys: panel kk/4, +32 bytes per step
            mma_ss(tmem + kColBdQ, ((kk >> 2) ? dS_k1 : dS_k0) + uint64_t((kk & 3) * 2), dK_k + uint64_t(kk * 128), idesc_nn, kk > 0);
          tc_commit(&sb->dq_full);
          tc_commit(&sb->pds_free[s]);
          tc_commit(&sb->qdo_empty[s]);
        }
        __syncwarp();
        TL(15);
      }
      TL_END();
    }
  } else if (warp < kBwdSoftmaxWarps) {
    // ============================== P / dS producers ==============================
    const int g = warp >> 2;  // column half
    const uint32_t lane_base = uint32_t((warp & 3) * 32);
    const int r = int(lane_base) + lane;  // row inside the query tile
    const uint32_t t_s = tmem + (lane_base << 16) + kColBS + g * 64;
    const uint32_t t_dp = tmem + (lane_base << 16) + kColBdP + g * 64;
    uint8_t* prow0 = sP + g * kPanelBytes + r * 128;
    uint8_t* dsrow0 = sdS + g * kPanelBytes + r * 128;
    TL_DECL((warp == 0 && lane == 0) ? 1 : -1);
    for (int i = 0; i < n_qt; ++i) {
      TL(20);
      const int row = i * kBlockM + r;
      float lse2 = 0.f, dl = 0.f;
      if (row < a.Tq) {
        const int64_t idx = (int64_t(b) * a.H + h) * a.Tq + row;
        lse2 = a.lse[idx] * 1.4426950408889634f;
        dl = a.delta[idx];
      }
      uint8_t* prow = prow0 + (i & 1) * 2 * kPanelBytes;
      uint8_t* dsrow = dsrow0 + (i & 1) * 2 * kPanelBytes;
      mbar_wait(&sb->sdp_full, i & 1);
      TL(21);
      if (i >= 2) mbar_wait(&sb->pds_free[i & 1], ((i >> 1) - 1) & 1);  // buffer consumed by the GEMMs of tile i-2
      TL(22);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld32(t_s + c * 32, sv);
        tmem_ld32(t_dp + c * 32, dv);
        tmem_wait_ld();
        uint32_t pp[16], dd[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = ex2(fmaf(__uint_as_float(sv[2 * e]), kScaleLog2, -lse2));
          const float p1 = ex2(fmaf(__uint_as_float(sv[2 * e + 1]), kScaleLog2, -lse2));
          const float d0 = p0 * (__uint_as_float(dv[2 * e]) - dl);
          const float d1 = p1 * (__uint_as_float(dv[2 * e + 1]) - dl);
          __nv_bfloat162 hp = __floats2bfloat162_rn(p0, p1), hd = __floats2bfloat162_rn(d0, d1);
          pp[e] = *reinterpret_cast<uint32_t*>(&hp);
          dd[e] = *reinterpret_cast<uint32_t*>(&hd);
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {  // 16-byte chunk index inside the 128-byte row, XOR-swizzled with (row & 7)
          const int chunk = (c * 4 + q4) ^ (r & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pp[4 * q4], pp[4 * q4 + 1], pp[4 * q4 + 2], pp[4 * q4 + 3]);
          *reinterpret_cast<uint4*>(dsrow + chunk * 16) = make_uint4(dd[4 * q4], dd[4 * q4 + 1], dd[4 * q4 + 2], dd[4 * q4 + 3]);
        }
      }
      TL(23);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb->pds_ready);
      TL(24);
    }
    TL_END();
  } else if (warp < kBwdTmaWarp) {
    // ============================== dQ drain, then dK / dV store ==============================
    const uint32_t lane_base = uint32_t((warp & 3) * 32);
    const int r = int(lane_base) + lane;
    const uint32_t t_dq = tmem + (lane_base << 16) + kColBdQ;
    TL_DECL((warp == kBwdDqWarp0 && lane == 0) ? 2 : -1);
    for (int i = 0; i < n_qt; ++i) {
      TL(30);
      mbar_wait(&sb->dq_full, i & 1);
      TL(31);
      tc_fence_after();
      uint32_t lo[32], hi[32];
      tmem_ld32(t_dq, lo);
      tmem_ld32(t_dq + 32, hi);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb->dq_empty);
      TL(32);
      const int row = i * kBlockM + r;
      if (row < a.Tq) {
        float* dst = a.dq_accum + (int64_t(b) * a.Tq + row) * (int64_t(a.H) * kHeadDim) + h * kHeadDim;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          red_add_v4(dst + 4 * e, __uint_as_float(lo[4 * e]), __uint_as_float(lo[4 * e + 1]),
                     __uint_as_float(lo[4 * e + 2]), __uint_as_float(lo[4 * e + 3]));
#pragma unroll
        for (int e = 0; e < 8; ++e)
          red_add_v4(dst + 32 + 4 * e, __uint_as_float(hi[4 * e]), __uint_as_float(hi[4 * e + 1]),
                     __uint_as_float(hi[4 * e + 2]), __uint_as_float(hi[4 * e + 3]));
      }
    }
    TL(33);
    TL_END();
    // all MMAs of the last tile are complete once dq_full(n_qt-1) fired (commit covers every earlier op)
    const int key = key0 + r;
    auto store_rows = [&](uint32_t col, __nv_bfloat16* base, int64_t sb_, int64_t st_, float scale) {
      __nv_bfloat16* dst = base + int64_t(b) * sb_ + int64_t(key) * st_ + h * kHeadDim;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem + (lane_base << 16) + col + c * 32, v);
        tmem_wait_ld();
        if (key < a.Tk) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(v[8 * q4 + 2 * e]) * scale,
                                                        __uint_as_float(v[8 * q4 + 2 * e + 1]) * scale);
              w[e] = *reinterpret_cast<uint32_t*>(&hb);
            }
            *reinterpret_cast<uint4*>(dst + c * 32 + q4 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    };
    store_rows(kColBdV, a.dv, a.v_sb, a.v_st, 1.0f);
    store_rows(kColBdK, a.dk, a.k_sb, a.k_st, 0.125f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBwdMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

// delta[b,h,t] = sum_c dO[b,t,h,c] * O[b,t,h,c]   (one warp per row)
__global__ void __launch_bounds__(256)
attn_bwd_delta_bf16_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, int64_t o_sb,
                           int64_t o_st, int B, int H, int Tq, float* __restrict__ delta) {
  const int64_t w = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= int64_t(B) * H * Tq) return;
  const int t = int(w % Tq), h = int((w / Tq) % H), b = int(w / (int64_t(Tq) * H));
  const int64_t off = b * o_sb + int64_t(t) * o_st + h * kHeadDim + lane * 2;
  const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(o + off));
  const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(d_o + off));
  const float acc = warp_sum(x.x * y.x + x.y * y.y);
  if (lane == 0) delta[w] = acc;
}

// dq = bf16(0.125 * dq_accum)
__global__ void __launch_bounds__(256)
attn_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int64_t q_sb, int64_t q_st,
                           int Tq, int D, int64_t total4) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total4; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t e = i * 4;
    const int c = int(e % D);
    const int64_t bt = e / D;
    const int t = int(bt % Tq);
    const int64_t b = bt / Tq;
    const float4 v = *reinterpret_cast<const float4*>(acc + e);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x * 0.125f, v.y * 0.125f), hi = __floats2bfloat162_rn(v.z * 0.125f, v.w * 0.125f);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dq + b * q_sb + int64_t(t) * q_st + c) = u;
  }
}

}  // namespace

bool attn_tc_bwd_supported(const aga_attn_params& p) { return attn_tc_supported(p); }

size_t attn_tc_bwd_workspace(const aga_attn_params& p) {
  const size_t delta = align_up(size_t(p.B) * p.H * p.Tq * sizeof(float), 256);
  const size_t dq = align_up(size_t(p.B) * p.Tq * p.H * kHeadDim * sizeof(float), 256);
  return delta + dq;
}

int attn_tc_bwd(const aga_attn_bwd_params& bp, void* ws, cudaStream_t s) {
  const aga_attn_params& p = bp.fwd;
  float* delta = static_cast<float*>(ws);
  float* dq_acc = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + align_up(size_t(p.B) * p.H * p.Tq * sizeof(float), 256));
  const size_t dq_bytes = size_t(p.B) * p.Tq * p.H * kHeadDim * sizeof(float);
  AGA_CUDA_TRY(cudaMemsetAsync(dq_acc, 0, dq_bytes, s));
  const int64_t rows = int64_t(p.B) * p.H * p.Tq;
  attn_bwd_delta_bf16_kernel<<<unsigned((rows * 32 + 255) / 256), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(p.out), static_cast<const __nv_bfloat16*>(bp.dout), p.o_stride_b, p.o_stride_t,
      p.B, p.H, p.Tq, delta);
  AGA_AFTER_LAUNCH();
  CUtensorMap mq, mk, mv, mdo;
  int st;
  if ((st = make_map(&mq, p.q, p.B, p.H, p.Tq, p.q_stride_b, p.q_stride_t, kBlockM)) != AGA_OK) return st;
  if ((st = make_map(&mk, p.k, p.B, p.H, p.Tk, p.k_stride_b, p.k_stride_t, kBlockN)) != AGA_OK) return st;
  if ((st = make_map(&mv, p.v, p.B, p.H, p.Tk, p.v_stride_b, p.v_stride_t, kBlockN)) != AGA_OK) return st;
  if ((st = make_map(&mdo, bp.dout, p.B, p.H, p.Tq, p.o_stride_b, p.o_stride_t, kBlockM)) != AGA_OK) return st;
  BwdArgs a{p.B, p.H, p.Tq, p.Tk, p.k_stride_b, p.k_stride_t, p.v_stride_b, p.v_stride_t, p.lse, delta, dq_acc,
            static_cast<__nv_bfloat16*>(bp.dk), static_cast<__nv_bfloat16*>(bp.dv)};
  AGA_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kBwdSmemBytes)));
  dim3 grid((p.Tk + kBlockN - 1) / kBlockN, p.H, p.B);
  attn_bwd_tc_kernel<<<grid, kBwdThreads, kBwdSmemBytes, s>>>(mq, mk, mv, mdo, a);
  AGA_AFTER_LAUNCH();
  const int D = p.H * kHeadDim;
  const int64_t total4 = int64_t(p.B) * p.Tq * D / 4;
  const unsigned gx = unsigned(std::min<int64_t>((total4 + 255) / 256, 148 * 16));
  attn_bwd_dq_convert_kernel<<<gx ? gx : 1, 256, 0, s>>>(dq_acc, static_cast<__nv_bfloat16*>(bp.dq), p.q_stride_b,
                                                          p.q_stride_t, p.Tq, D, total4);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

}  // namespace aga
