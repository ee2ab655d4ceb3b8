// bf16 GEMM with the exact (erf) GELU fused into its epilogue, on tcgen05 / TMEM / TMA.
//
// Reference: ResidualAttentionBlock.mlp = Linear(D, 4D) -> GELU -> Linear(4D, D), whisper/whisper/model.py:213,242.
// With cuBLAS the GELU is a separate HBM pass each way (forward: read h, write g; backward: read h and dg, write dh —
// 2.9 ms of a 26 ms step), and cuBLASLt's GELU epilogue is the tanh approximation.  Two modes of one kernel:
//   mode 0 (forward) : h = bf16(A W^T + bias),  g = bf16(gelu(h))             -> writes h (kept for backward) and g
//   mode 1 (backward): dh = bf16( bf16(A W^T) * gelu'(h) )                     -> reads h tiles in the epilogue
// A (M, K) and W (N, K) are row-major bf16 (both K-major operands; the backward passes the cached transpose of the
// second Linear's frozen weight).  Persistent CTAs, one per SM; CTA tile 128 x 256, K in steps of 64 through a 3-stage
// TMA ring; tcgen05.mma M128 N256 K16 (128 clk each, 96 B/clk of shared-memory operand traffic); the fp32 accumulator is
// double-buffered in TMEM (2 x 256 columns) so that the 8 epilogue warps (erf costs ~25 instructions per element)
// work on tile i while the tensor core runs tile i+1; results leave through swizzled staging rows and TMA tile stores.
// What bounds it is the L2 -> SM operand stream: the MMA warp spends 46 % of its time waiting for `full` and 3 % waiting
// for the epilogue (tools/gemm_waits.py), and this kernel, its variants and cuBLAS's 256x256 nvjet kernel all settle at
// ~36 B/clk per SM of TMA requests out of L2 (~8.7 TB/s chip-wide); FLOP per requested byte decides the rest.  CTAs are
// therefore clustered and share loads by TMA multicast.  Variants (bit-identical results, profiles/r2_gemm_gelu.md):
//   1 (default) multicast pair, 256 x 256 per cluster, 128 FLOP/B                      934 TFLOP/s
//   2 one tcgen05.mma.cta_group::2 M256 N256 per pair, W halves not duplicated, 5 stages   856 - 904 TFLOP/s
//   3 multicast quad 2 x 2, 256 x 512 per cluster, 171 FLOP/B (cuBLAS's ratio); only 33 four-CTA clusters
//     (132 of 148 SMs) are co-resident                                                     888 - 898 TFLOP/s
#include "aga_common.cuh"
#include "tc_ptx.cuh"
#include "gelu_math.cuh"

#include <cudaTypedefs.h>

// Debug build (-DAGA_TIMELINE, tools/gemm_waits.py): every CTA adds up the clock64() cycles its producer, MMA and first
// epilogue warp spend inside their mbarrier waits; 16 int64 slots per CTA in a buffer set with aga_debug_set_gemm_waits().
#ifdef AGA_TIMELINE
__device__ long long* g_gemm_waits = nullptr;
extern "C" __attribute__((visibility("default"))) int aga_debug_set_gemm_waits(long long* p) {
  return cudaMemcpyToSymbol(g_gemm_waits, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}
#define GW_DECL() long long gw_acc[3] = {0, 0, 0}; const long long gw_t0 = clock64()
#define GW_WAIT(slot, stmt) do { const long long gw_a = clock64(); stmt; gw_acc[slot] += clock64() - gw_a; } while (0)
#define GW_STORE(base, n) do { if (g_gemm_waits && lane == 0) { long long* r = g_gemm_waits + (long long)blockIdx.x * 16 + (base); \
    for (int i_ = 0; i_ < (n); ++i_) r[i_] = gw_acc[i_]; r[n] = clock64() - gw_t0; } } while (0)
#else
#define GW_DECL() do { } while (0)
#define GW_WAIT(slot, stmt) stmt
#define GW_STORE(base, n) do { } while (0)
#endif

namespace aga {
namespace {

using namespace ptx;

constexpr int kTM = 128, kTN = 256, kTK = 64;
constexpr int kGStages = 3;      // variants 0 / 1: A tile + whole W tile per stage
constexpr int kGStages2 = 5;     // variant 2 (two-SM MMA): A tile + half W tile per stage
constexpr int kGMaxStages = 5;
constexpr int kATile = kTM * kTK * 2;   // 16 KiB
constexpr int kBTile = kTN * kTK * 2;   // 32 KiB
constexpr int kEpiWarps = 8;
constexpr int kGTmaWarp = 0, kGMmaWarp = 1, kGEpiWarp0 = 4;
constexpr int kGThreads = (kGEpiWarp0 + kEpiWarps) * 32;  // 384
constexpr int kChunkBytes = 32 * 128;   // one warp's 32 rows x 64 bf16 columns

struct GSmem {
  uint64_t full[kGMaxStages], empty[kGMaxStages];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t h_ready[kEpiWarps];
  uint32_t tmem_base;
  alignas(16) float bias[kTN];
};
// stages | per epilogue warp: two 4 KiB staging chunks (h / g, or h_in / dh)
constexpr size_t kGSmemBytes = 1024 + size_t(kGStages) * (kATile + kBTile) + size_t(kEpiWarps) * 2 * kChunkBytes + sizeof(GSmem);
constexpr size_t kGSmemBytes2 = 1024 + size_t(kGStages2) * (kATile + kBTile / 2) + size_t(kEpiWarps) * 2 * kChunkBytes + sizeof(GSmem);
static_assert(kGSmemBytes <= 227 * 1024 && kGSmemBytes2 <= 227 * 1024, "gemm_gelu exceeds the shared-memory limit");

struct GArgs {
  int M, N, K;
  int mode;             // 0: forward (bias, h and g out), 1: backward (h in, dh out)
  const __nv_bfloat16* bias;  // (N) or nullptr
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// the same tile delivered to the same shared-memory offset of every CTA in `mask`; each destination's mbarrier (same
// offset) receives the bytes
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                      uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// mbarrier arrive (same offset) in every CTA of `mask` once this thread's tcgen05 ops have completed
__device__ __forceinline__ void tc_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// ---- two-SM (cta_group::2) forms: the leader CTA (cluster rank 0) issues the MMAs and owns the `full` / `acc_empty` barriers
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {  // shared::cta address -> the same offset in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// tile into THIS CTA's shared memory, bytes counted on a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// D[tmem of both CTAs: 128 lanes each] (+)= A[each CTA's 128 rows] * B[N/2 rows from each CTA]
__device__ __forceinline__ void mma_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// kVar 0: one CTA per 128 x 256 tile (M < 256).
// kVar 1: clusters of two CTAs work on two vertically adjacent 128-row tiles of the same 256-column panel; each CTA
// fetches HALF of the W tile and multicasts it into both CTAs' shared memory, so the L2 -> SM operand traffic per FLOP is
// that of a 256 x 256 tile (the kernel is L2-bandwidth-bound with 128 x 256 tiles: 1.33 GB per Whisper-small MLP GEMM).
// kVar 2: the same pairing, but the W halves stay where they land and the leader CTA issues cta_group::2 MMAs over both
// CTAs' shared memory and TMEM: half the B-operand reads and half the TMA fill bytes per SM.
// kVar 3: clusters of four CTAs = 2 m-tiles x 2 n-tiles.  CTA (r, c) loads rows [64c, +64) of ITS m-tile's A tile and
// multicasts them to its row mate, loads rows [128r, +128) of ITS n-tile's W tile and multicasts them to its column mate:
// 24 KiB from L2 per CTA and K step instead of 32.
template <int kVar>
__global__ void __launch_bounds__(kGThreads, 1)
gemm_gelu_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a64,
                 const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_h,
                 const __grid_constant__ CUtensorMap map_o, const GArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr bool kPair = kVar != 0;  // any clustered variant
  constexpr bool kTwoSm = kVar == 2;
  constexpr bool kQuad = kVar == 3;
  constexpr int kCluster = kQuad ? 4 : (kPair ? 2 : 1);
  constexpr int kNStages = kTwoSm ? kGStages2 : kGStages;
  constexpr int kBStage = kTwoSm ? kBTile / 2 : kBTile;  // bytes of W per stage in THIS CTA
  uint8_t* sA = smem;
  uint8_t* sB = sA + kNStages * kATile;
  uint8_t* sE = sB + kNStages * kBStage;  // epilogue staging: warp w -> [2][32 rows x 128 B]
  GSmem* sb = reinterpret_cast<GSmem*>(sE + kEpiWarps * 2 * kChunkBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_mt = (a.M + kTM - 1) / kTM, n_nt = (a.N + kTN - 1) / kTN, n_k = (a.K + kTK - 1) / kTK;
  // work units: tiles, (pair of m-tiles, n-tile) per pair, or (pair of m-tiles, pair of n-tiles) per quad;
  // `rank` (pairs) / `qr`, `qc` (quads) select this CTA's tile of the unit
  const int crank = kPair ? int(cluster_ctarank()) : 0;
  const int rank = kQuad ? 0 : crank;
  const int qr = crank >> 1, qc = crank & 1;
  const int n_nu = kQuad ? (n_nt + 1) / 2 : n_nt;  // units along n
  const int n_units = kPair ? ((n_mt + 1) / 2) * n_nu : n_mt * n_nt;
  const int unit0 = int(blockIdx.x) / kCluster;
  const int unit_step = int(gridDim.x) / kCluster;
  auto tile_of = [&](int u, int& mt, int& nt) {
    nt = u % n_nu;  // n fastest: neighbouring CTAs share the A rows in L2
    mt = kPair ? 2 * (u / n_nu) + (kQuad ? qr : rank) : u / n_nu;
    if (kQuad) nt = 2 * nt + qc;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kNStages; ++s) {
      mbar_init(&sb->full[s], 1);  // two-SM: the leader's producer announces both CTAs' bytes on its barrier
      // multicast variants: every CTA that writes into a stage waits for the MMA warps of all CTAs it writes to
      mbar_init(&sb->empty[s], kVar == 1 ? 2 : (kQuad ? 3 : 1));
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sb->acc_full[i], 1);
      mbar_init(&sb->acc_empty[i], kTwoSm ? 2 * kEpiWarps : kEpiWarps);  // two-SM: the leader waits for both CTAs' epilogues
    }
    for (int w = 0; w < kEpiWarps; ++w) mbar_init(&sb->h_ready[w], 1);
    fence_barrier_init();
  }
  if (warp == kGMmaWarp) {
    if (kTwoSm) {  // the same warp of both CTAs takes part; both get the same columns
      tmem_alloc_2sm(&sb->tmem_base, 512);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(&sb->tmem_base, 512);
      tmem_relinquish();
    }
  }
  if (warp == kGTmaWarp && lane == 0) {
    prefetch_tensormap(kQuad ? &map_a64 : &map_a);
    prefetch_tensormap(&map_w);
    prefetch_tensormap(&map_h);
    prefetch_tensormap(&map_o);
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast into this CTA
  tc_fence_after();
  const uint32_t tmem = sb->tmem_base;

  if (warp == kGTmaWarp) {
    // ============================== TMA producer ==============================
    int it = 0;
    GW_DECL();
    for (int u = unit0; u < n_units; u += unit_step) {
      int mt, nt;
      tile_of(u, mt, nt);
      for (int k = 0; k < n_k; ++k, ++it) {
        const int s = it % kNStages;
        GW_WAIT(0, mbar_wait(&sb->empty[s], ((it / kNStages) & 1) ^ 1));
        if (elect_one()) {
          if (kTwoSm) {  // own A rows + own half of the W tile, counted on the leader's barrier (which expects both CTAs' bytes)
            const uint32_t full_leader = mapa_rank(smem_u32(&sb->full[s]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&sb->full[s], 2 * (kATile + kBTile / 2));
            tma_load_2d_2sm(sA + s * kATile, &map_a, full_leader, k * kTK, mt * kTM);
            tma_load_2d_2sm(sB + s * kBStage, &map_w, full_leader, k * kTK, nt * kTN + rank * (kTN / 2));
            continue;
          }
          mbar_arrive_expect_tx(&sb->full[s], kATile + kBTile);
          if (kQuad) {  // half of the A tile to the row mates, half of the W tile to the column mates
            tma_load_2d_multicast(sA + s * kATile + qc * (kATile / 2), &map_a64, &sb->full[s], k * kTK, mt * kTM + qc * (kTM / 2),
                                  uint16_t(3u << (2 * qr)));
            tma_load_2d_multicast(sB + s * kBTile + qr * (kBTile / 2), &map_w, &sb->full[s], k * kTK, nt * kTN + qr * (kTN / 2),
                                  uint16_t(5u << qc));
            continue;
          }
          tma_load_2d(sA + s * kATile, &map_a, &sb->full[s], k * kTK, mt * kTM);
          if (kPair) {  // this CTA's half of the W tile (128 of its 256 rows), delivered to both CTAs
            tma_load_2d_multicast(sB + s * kBTile + rank * (kBTile / 2), &map_w, &sb->full[s], k * kTK,
                                  nt * kTN + rank * (kTN / 2), uint16_t(3));
          } else {
            tma_load_2d(sB + s * kBTile, &map_w, &sb->full[s], k * kTK, nt * kTN);
            tma_load_2d(sB + s * kBTile + kBTile / 2, &map_w, &sb->full[s], k * kTK, nt * kTN + kTN / 2);
          }
        }
      }
    }
    GW_STORE(0, 1);  // [0] wait empty, [1] producer total
  } else if (warp == kGMmaWarp) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc = make_idesc_bf16(kTwoSm ? 2 * kTM : kTM, kTN, 0, 0);
    int it = 0, tile_i = 0;
    GW_DECL();
    for (int u = unit0; u < n_units && !(kTwoSm && rank != 0); u += unit_step, ++tile_i) {  // two-SM: the leader issues for the pair
      const int buf = tile_i & 1;
      GW_WAIT(0, mbar_wait(&sb->acc_empty[buf], ((tile_i >> 1) & 1) ^ 1));  // the epilogue has drained this accumulator
      tc_fence_after();
      for (int k = 0; k < n_k; ++k, ++it) {
        const int s = it % kNStages;
        GW_WAIT(1, mbar_wait(&sb->full[s], (it / kNStages) & 1));
        tc_fence_after();
        const uint64_t da = make_smem_desc_sw128(smem_u32(sA + s * kATile));
        const uint64_t db = make_smem_desc_sw128(smem_u32(sB + s * kBStage));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kTK / 16; ++kk) {
            if (kTwoSm) mma_ss_2sm(tmem + buf * kTN, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc, (k > 0 || kk > 0) ? 1u : 0u);
            else mma_ss(tmem + buf * kTN, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc, (k > 0 || kk > 0) ? 1u : 0u);
          }
          if (kTwoSm) {  // both CTAs' producers / epilogues wait on their own copies of these barriers
            tc_commit_2sm(&sb->empty[s], uint16_t(3));
            if (k == n_k - 1) tc_commit_2sm(&sb->acc_full[buf], uint16_t(3));
          } else {
            if (kQuad) tc_commit_multicast(&sb->empty[s], uint16_t((3u << (2 * qr)) | (5u << qc)));  // self, row mate, column mate
            else if (kPair) tc_commit_multicast(&sb->empty[s], uint16_t(3));
            else tc_commit(&sb->empty[s]);
            if (k == n_k - 1) tc_commit(&sb->acc_full[buf]);
          }
        }
        __syncwarp();
      }
    }
    GW_STORE(2, 2);  // [2] wait acc_empty, [3] wait full, [4] MMA warp total
  } else if (warp >= kGEpiWarp0) {
    // ============================== epilogue: warp = (row quadrant, column half) ==============================
    const int ew = warp - kGEpiWarp0;
    const int quad = warp & 3, half = ew >> 2;  // TMEM lanes [32 quad, +32); columns [128 half, +128)
    const uint32_t lane_base = uint32_t(quad * 32);
    uint8_t* st0 = sE + ew * 2 * kChunkBytes;   // h (mode 0) / h_in (mode 1)
    uint8_t* st1 = st0 + kChunkBytes;           // g (mode 0) / dh  (mode 1)
    const uint32_t row0_addr = smem_u32(st0 + lane * 128), row1_addr = smem_u32(st1 + lane * 128);
    int tile_i = 0, hph = 0;
    GW_DECL();
    for (int u = unit0; u < n_units; u += unit_step, ++tile_i) {
      int mt, nt;
      tile_of(u, mt, nt);
      const int buf = tile_i & 1;
      const int grow = mt * kTM + int(lane_base);  // first global row of this warp
      if (a.mode == 0) {
        // this tile's bias slice -> smem (all epilogue warps of the tile; 256 threads, one value each)
        named_bar_sync(1, kEpiWarps * 32);  // the previous tile's readers are done
        const int n = nt * kTN + ew * 32 + lane;
        sb->bias[ew * 32 + lane] = (a.bias && n < a.N) ? __bfloat162float(a.bias[n]) : 0.f;
        named_bar_sync(1, kEpiWarps * 32);
      }
      GW_WAIT(0, mbar_wait(&sb->acc_full[buf], (tile_i >> 1) & 1));
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {  // two 64-column chunks per warp
        const int col0 = half * 128 + c * 64;         // column inside the tile
        const int gcol = nt * kTN + col0;
        GW_WAIT(1, if (lane == 0) bulk_wait_group_read0(); __syncwarp());  // the previous chunk's stores have read the staging rows
        if (a.mode == 1) {
          if (lane == 0) {
            mbar_arrive_expect_tx(&sb->h_ready[ew], kChunkBytes);
            tma_load_2d(st0, &map_h, &sb->h_ready[ew], gcol, grow);
          }
        }
        uint32_t v[2][32];
        GW_WAIT(2, tmem_ld32(tmem + (lane_base << 16) + buf * kTN + col0, v[0]);
                   tmem_ld32(tmem + (lane_base << 16) + buf * kTN + col0 + 32, v[1]); tmem_wait_ld());
        if (c == 1) {  // both chunks of this warp are in registers: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kTwoSm) mbar_arrive_cluster(mapa_rank(smem_u32(&sb->acc_empty[buf]), 0)); else mbar_arrive(&sb->acc_empty[buf]);
          }
        }
        if (a.mode == 1) {
          mbar_wait(&sb->h_ready[ew], hph & 1);
          ++hph;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {  // 16-byte chunk q of the 128-byte row = 8 columns
          const uint32_t* src = v[q >> 2];
          const int e0 = (q & 3) * 8;
          uint32_t w0[4], w1[4];
          if (a.mode == 0) {
            const float4 b0 = *reinterpret_cast<const float4*>(&sb->bias[col0 + q * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sb->bias[col0 + q * 8 + 4]);
            const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 acc = __fadd2_rn(make_float2(__uint_as_float(src[e0 + 2 * e]), __uint_as_float(src[e0 + 2 * e + 1])), bb[e]);
              __nv_bfloat162 hh = __floats2bfloat162_rn(acc.x, acc.y);  // h as the reference's separate GELU kernel sees it
              w0[e] = *reinterpret_cast<uint32_t*>(&hh);
              const float2 g2 = gelu_erf2(make_float2(__uint_as_float(w0[e] << 16), __uint_as_float(w0[e] & 0xffff0000u)));
              __nv_bfloat162 gg = __floats2bfloat162_rn(g2.x, g2.y);
              w1[e] = *reinterpret_cast<uint32_t*>(&gg);
            }
            sts128(row0_addr + uint32_t((q ^ (lane & 7)) * 16), w0[0], w0[1], w0[2], w0[3]);
          } else {
            uint4 hv;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(hv.x), "=r"(hv.y), "=r"(hv.z), "=r"(hv.w)
                         : "r"(row0_addr + uint32_t((q ^ (lane & 7)) * 16)) : "memory");
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 dgp = dgelu_erf2(make_float2(__uint_as_float(hw[e] << 16), __uint_as_float(hw[e] & 0xffff0000u)));
              __nv_bfloat162 db = __floats2bfloat162_rn(__uint_as_float(src[e0 + 2 * e]), __uint_as_float(src[e0 + 2 * e + 1]));
              const uint32_t dw = *reinterpret_cast<uint32_t*>(&db);  // the dgrad GEMM's result as bf16, like cuBLAS writes it
              const float2 d2 = __fmul2_rn(make_float2(__uint_as_float(dw << 16), __uint_as_float(dw & 0xffff0000u)), dgp);
              __nv_bfloat162 dd = __floats2bfloat162_rn(d2.x, d2.y);
              w1[e] = *reinterpret_cast<uint32_t*>(&dd);
            }
          }
          sts128(row1_addr + uint32_t((q ^ (lane & 7)) * 16), w1[0], w1[1], w1[2], w1[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (a.mode == 0) tma_store_2d(&map_h, st0, gcol, grow);
          tma_store_2d(&map_o, st1, gcol, grow);
          bulk_commit_group();
        }
      }
    }
    if (lane == 0) bulk_wait_group_read0();
    if (ew == 0) GW_STORE(5, 3);  // [5] wait acc_full, [6] wait store-read, [7] TMEM loads, [8] epilogue warp total
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it or signal its barriers
  if (warp == kGMmaWarp) {
    if (kTwoSm) tmem_dealloc_2sm(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}

PFN_cuTensorMapEncodeTiled encode_fn() {
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

// row-major (rows, cols) bf16 matrix; box = 64 columns x box_rows rows, SWIZZLE_128B; out-of-range parts zero-filled / clipped
int make_map_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  ensure_context_in_this_thread();
  PFN_cuTensorMapEncodeTiled enc = encode_fn();
  if (!enc) return AGA_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
  cuuint32_t box[2] = {64, cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? AGA_OK : AGA_ERR_INVALID_ARGUMENT;
}

#ifndef AGA_GEMM_VARIANT
#define AGA_GEMM_VARIANT 1
#endif
int g_gemm_variant = AGA_GEMM_VARIANT;  // 0 single CTA, 1 multicast pair, 2 two-SM MMA, 3 multicast quad (see the kernel's comment)

// how many 4-CTA clusters of the quad kernel the device can hold at once (GPC boundaries decide; 0 = cannot launch)
int quad_clusters() {
  static const int n = []() {
    if (cudaFuncSetAttribute(gemm_gelu_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kGSmemBytes)) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * 64);
    cfg.blockDim = dim3(kGThreads);
    cfg.dynamicSmemBytes = kGSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int c = 0;
    if (cudaOccupancyMaxActiveClusters(&c, gemm_gelu_kernel<3>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    return c;
  }();
  return n;
}

}  // namespace
}  // namespace aga

using namespace aga;

// measurement hook (tools/bench_gemm_gelu.py): choose the kernel variant for M >= 256; returns the previous one
extern "C" __attribute__((visibility("default"))) int aga_debug_set_gemm_variant(int variant) {
  const int prev = g_gemm_variant;
  if (variant >= 0 && variant <= 3) g_gemm_variant = variant;
  return prev;
}
extern "C" __attribute__((visibility("default"))) int aga_debug_gemm_quad_clusters() { return quad_clusters(); }

// mode 0: h (M,N) = bf16(a (M,K) @ w (N,K)^T + bias (N)),  out (M,N) = bf16(gelu(h));  h is an OUTPUT
// mode 1: out (M,N) = bf16(bf16(a @ w^T) * gelu'(h));                                   h is an INPUT, bias ignored
extern "C" int aga_gemm_gelu(const void* a, const void* w, const void* bias, void* h, void* out, int mode, int64_t M, int N,
                             int K, void* stream) {
  if (!a || !w || !h || !out || M <= 0 || N <= 0 || K <= 0 || (mode != 0 && mode != 1)) return AGA_ERR_INVALID_ARGUMENT;
  if (K % 8 != 0 || N % 8 != 0) return AGA_ERR_UNSUPPORTED;  // 16-byte global strides for TMA
  const uintptr_t all = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(h) |
                        reinterpret_cast<uintptr_t>(out);
  if (all & 15) return AGA_ERR_UNSUPPORTED;
  if (M > 2147483647LL) return AGA_ERR_UNSUPPORTED;
  CUtensorMap ma, ma64, mw, mh, mo;
  int st;
  if ((st = make_map_2d(&ma, a, M, K, K, kTM)) != AGA_OK) return st;
  if ((st = make_map_2d(&ma64, a, M, K, K, kTM / 2)) != AGA_OK) return st;  // quads fetch A tiles as two 64-row halves
  if ((st = make_map_2d(&mw, w, N, K, K, kTN / 2)) != AGA_OK) return st;  // W tiles arrive as two 128-row halves
  if ((st = make_map_2d(&mh, h, M, N, N, 32)) != AGA_OK) return st;
  if ((st = make_map_2d(&mo, out, M, N, N, 32)) != AGA_OK) return st;
  GArgs ga{int(M), N, K, mode, static_cast<const __nv_bfloat16*>(bias)};
  static const int n_sm = []() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
  }();
  const int n_mt = int((M + kTM - 1) / kTM), n_nt = (N + kTN - 1) / kTN;
  if (n_mt >= 2 && g_gemm_variant != 0) {
    const bool two_sm = g_gemm_variant == 2;
    const bool quad = g_gemm_variant == 3 && n_nt >= 2 && quad_clusters() > 0;
    const int csize = quad ? 4 : 2;
    const int n_units = ((n_mt + 1) / 2) * (quad ? (n_nt + 1) / 2 : n_nt);
    const int n_clusters = std::max(1, std::min(n_units, quad ? quad_clusters() : n_sm / 2));
    auto kern = quad ? gemm_gelu_kernel<3> : (two_sm ? gemm_gelu_kernel<2> : gemm_gelu_kernel<1>);
    const size_t smem_bytes = two_sm ? kGSmemBytes2 : kGSmemBytes;
    AGA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(csize * n_clusters));
    cfg.blockDim = dim3(kGThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = unsigned(csize);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    AGA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ma, ma64, mw, mh, mo, ga));
    AGA_AFTER_LAUNCH();
    return AGA_OK;
  }
  AGA_CUDA_TRY(cudaFuncSetAttribute(gemm_gelu_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kGSmemBytes)));
  gemm_gelu_kernel<0><<<std::min(n_mt * n_nt, n_sm), kGThreads, kGSmemBytes, static_cast<cudaStream_t>(stream)>>>(ma, ma64, mw, mh, mo, ga);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}
