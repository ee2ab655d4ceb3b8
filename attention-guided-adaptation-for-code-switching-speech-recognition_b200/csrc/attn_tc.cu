// bf16 flash attention forward on the 5th-gen tensor cores: tcgen05.mma with TMEM accumulators, TMA-staged
// Q/K/V tiles, warp-specialised (2 softmax warpgroups + 1 TMA warp + 1 MMA warp), head dim 64.
//
// Reference: MultiHeadAttention.qkv_attention, whisper/whisper/model.py:93-109 (non-causal: encoder self
// attention 1500x1500 and decoder cross attention Tx1500 — 99.9 % of the attention FLOPs of a step).
//
// One CTA = 256 query rows of one (batch, head): two 128-row tiles A and B, each owned by one softmax
// warpgroup, sharing every K/V tile that TMA brings in (halves L2->smem traffic per FLOP).
//   TMEM (512 columns): S_A [0,128) | S_B [128,256) | O_A [256,320) | O_B [320,384)
//     S_t = Q_t K^T   : tcgen05.mma  M=128 N=128 K=16 x4, A/B from smem (K-major, SWIZZLE_128B)
//     P_t (bf16) overwrites the first 64 columns of S_t (two keys per 32-bit column)
//     O_t += P_t V    : tcgen05.mma  M=128 N=64  K=16 x8, A from TMEM, B = V tile from smem (MN-major)
//   The MMA warp issues  PV_A(j), S_A(j+1), PV_B(j), S_B(j+1): while warpgroup A runs the softmax of tile
//   j+1 the tensor core works for warpgroup B, and vice versa.
//   Online softmax in the exp2 domain with lazy rescaling: O_t is only rescaled (TMEM round trip) when the
//   running row maximum grew by more than 2^8, otherwise the stale maximum keeps being used.
#include "aga_common.cuh"
#include "attn_common.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>

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
constexpr int kMmaWarp = 9;
constexpr int kThreads = 320;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS = 0, kColO = 256;
constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;  // d^-1/2 * log2(e)
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;

struct FwdSmem {
  // barriers
  uint64_t q_full;
  uint64_t k_full[kStages], k_empty[kStages], v_full[kStages], v_empty[kStages];
  uint64_t s_full[2], p_ready[2], pv_done[2];
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
      mbar_init(&sb->k_empty[s], 1);
      mbar_init(&sb->v_full[s], 1);
      mbar_init(&sb->v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&sb->s_full[t], 1);
      mbar_init(&sb->p_ready[t], 4);  // one arrival per softmax warp of the warpgroup
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
    if (lane == 0) {
      mbar_arrive_expect_tx(&sb->q_full, (active_b ? 2 : 1) * kTileBytes);
      tma_load_4d(sQ, &map_q, &sb->q_full, 0, h, row0, b);
      if (active_b) tma_load_4d(sQ + kTileBytes, &map_q, &sb->q_full, 0, h, row0 + kBlockM, b);
      for (int j = 0; j < n_kt; ++j) {
        const int s = j % kStages;
        const uint32_t ph = (j / kStages) & 1;
        mbar_wait(&sb->k_empty[s], ph ^ 1);  // first pass through the ring returns immediately
        mbar_arrive_expect_tx(&sb->k_full[s], kTileBytes);
        tma_load_4d(sK + s * kTileBytes, &map_k, &sb->k_full[s], 0, h, j * kBlockN, b);
        mbar_wait(&sb->v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&sb->v_full[s], kTileBytes);
        tma_load_4d(sV + s * kTileBytes, &map_v, &sb->v_full[s], 0, h, j * kBlockN, b);
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA issuer (one thread) ==============================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(kBlockM, kBlockN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kBlockM, kHeadDim, 0, 1);
      const int n_tiles = active_b ? 2 : 1;
      auto issue_s = [&](int t, int stage) {
        const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ + t * kTileBytes));
        const uint64_t dk = make_smem_desc_sw128(smem_u32(sK + stage * kTileBytes));
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk)  // +32 bytes along K inside the 128-byte swizzle row
          mma_ss(tmem + kColS + t * kBlockN, dq + uint64_t(kk * 2), dk + uint64_t(kk * 2), idesc_qk, kk > 0);
        tc_commit(&sb->s_full[t]);
      };
      mbar_wait(&sb->q_full, 0);
      mbar_wait(&sb->k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < n_tiles; ++t) issue_s(t, 0);
      tc_commit(&sb->k_empty[0]);
      for (int j = 0; j < n_kt; ++j) {
        const int s = j % kStages;
        const uint32_t ph = (j / kStages) & 1;
        const int s1 = (j + 1) % kStages;
        const uint32_t ph1 = ((j + 1) / kStages) & 1;
        const bool more = j + 1 < n_kt;
        mbar_wait(&sb->v_full[s], ph);
        if (more) mbar_wait(&sb->k_full[s1], ph1);
        for (int t = 0; t < n_tiles; ++t) {
          mbar_wait(&sb->p_ready[t], j & 1);
          tc_fence_after();
          const uint64_t dv = make_smem_desc_sw128(smem_u32(sV + s * kTileBytes));
#pragma unroll
          for (int kk = 0; kk < kBlockN / 16; ++kk)  // A: +8 TMEM columns (16 bf16); B: +16 key rows = 2048 bytes
            mma_ts(tmem + kColO + t * kHeadDim, tmem + kColS + t * kBlockN + kk * 8, dv + uint64_t(kk * 128), idesc_pv,
                   (j > 0 || kk > 0) ? 1u : 0u);
          tc_commit(&sb->pv_done[t]);
          if (more) issue_s(t, s1);
        }
        tc_commit(&sb->v_empty[s]);
        if (more) tc_commit(&sb->k_empty[s1]);
      }
    }
  } else {
    // ============================== softmax / correction / epilogue ==============================
    const int t = warp >> 2;                        // warpgroup 0 -> tile A, 1 -> tile B
    const uint32_t lane_base = uint32_t((warp & 3) * 32);
    const int row = row0 + t * kBlockM + int(lane_base) + lane;
    if (t == 0 || active_b) {
      const uint32_t t_s = tmem + (lane_base << 16) + kColS + t * kBlockN;
      const uint32_t t_o = tmem + (lane_base << 16) + kColO + t * kHeadDim;
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < n_kt; ++j) {
        mbar_wait(&sb->s_full[t], j & 1);
        tc_fence_after();
        const int valid = a.Tk - j * kBlockN;  // keys of this tile that exist (>= 128 except on the last tile)
        // ---- pass 1: row maximum
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < kBlockN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_s + c * 32, r);
          tmem_wait_ld();
          if (valid >= (c + 1) * 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(r[i]));
          }
        }
        const float m_new = fmaxf(m_used, mx * kScaleLog2);
        if (j == 0) {
          m_used = m_new;
        } else if (__any_sync(0xffffffffu, m_new - m_used > kRescaleThreshold)) {
          // s_full(j) completing implies PV(j-1) completed (in-order tensor pipe): O_t is stable here
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
        // ---- pass 2: P = exp2(S*c - m), packed to bf16 over the S columns already consumed
        float rs = 0.f;
#pragma unroll 1
        for (int c = 0; c < kBlockN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_s + c * 32, r);
          tmem_wait_ld();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = ex2(fmaf(__uint_as_float(r[2 * i]), kScaleLog2, -m_used));
            float p1 = ex2(fmaf(__uint_as_float(r[2 * i + 1]), kScaleLog2, -m_used));
            if (c * 32 + 2 * i >= valid) p0 = 0.f;
            if (c * 32 + 2 * i + 1 >= valid) p1 = 0.f;
            rs += p0 + p1;
            __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
            pk[i] = *reinterpret_cast<uint32_t*>(&hb);
          }
          tmem_st16(t_s + c * 16, pk);
        }
        l += rs;
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb->p_ready[t]);
      }
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

bool attn_tc_bwd_supported(const aga_attn_params&) { return false; }
size_t attn_tc_bwd_workspace(const aga_attn_params&) { return 0; }
int attn_tc_bwd(const aga_attn_bwd_params&, void*, cudaStream_t) { return AGA_ERR_UNSUPPORTED; }

}  // namespace aga
