// Weight gradients of the adapter Linears on tcgen05: C (M x N, fp32) += A^T B with A (rows x M) and B (rows x N) bf16
// row-major and a HUGE contraction (rows = batch x frames = 24000) over a small output (768 x 192).
//
// Reference: autograd of Adapter.model = Linear -> GELU -> Linear (whisper/whisper/model.py:181-194): for each adapter
//   dW2 = ds^T g   (D x D/4)      and      dW1 = dh1^T x  (D/4 x D)  — written here as (x^T dh1)^T, the same shape class.
// These are the only weight gradients of the training step (--freeze_param leaves the adapters trainable), 48 per step.
// cuBLAS runs them as split-K kernels + a reduce pass: 18.5 + 2.7 us each against an HBM floor of 7 us (46 MB of
// operands) — 1.0 ms of a 24.4 ms step.  Here: one CTA per (128-row slice of M, K split); both operands are MN-major
// for the tensor core (the contraction index is the row index, the contiguous direction is M / N), staged by TMA as
// SWIZZLE_128B panels of 64 rows x 64 elements straight from the row-major activations — no transposed copy; one
// tcgen05.mma M128 N(<=256) K16 per 16 rows; a 4-stage ring; the K splits are combined with coalesced fp32 atomics into
// the caller's zero-initialised output (the step's cleared arena, ops.ZeroArena).  The accumulator's lane index is the
// contiguous index of the output, so the kernel produces (a^T b)^T; the other layout is the same kernel with a and b
// exchanged.
#include "aga_common.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>

#include <algorithm>

namespace aga {
namespace {

using namespace ptx;

constexpr int kWK = 64;                  // contraction rows per stage
constexpr int kWPanel = kWK * 128;       // one panel: 64 rows x 64 bf16 = 8 KiB
constexpr int kWTmaWarp = 0, kWMmaWarp = 1, kWEpiWarp0 = 2;
constexpr int kWEpiWarps = 8;           // two per TMEM lane quadrant, half of the tile's columns each
constexpr int kWThreads = (2 + kWEpiWarps) * 32;

struct WSmem {
  uint64_t full[4], empty[4], acc_full;
  uint32_t tmem_base;
};

struct WArgs {
  int rows, M, N;     // the kernel's view: out (N, M) += (a^T b)^T, a (rows, M), b (rows, N)
  int k_chunk;        // contraction rows per split (multiple of kWK)
  int n_stages;
  int mt;             // 128-row slices of M per CTA (1 or 2): two slices share every B panel that is fetched
  float* out;
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// MN-major SWIZZLE_128B operand made of 64-element panels along M / N: LBO = distance between panels, SBO = 8 contraction rows
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t(lbo_bytes >> 4) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// grid: x = slices of 128 * mt columns of a (output columns), y = K splits, z = tiles of 256 columns of b (output rows).
// Columns of a / b past M / N are zero-filled by TMA (M = 192: the second 128-slice is half empty) and masked at the end.
__global__ void __launch_bounds__(kWThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const WArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int m0 = blockIdx.x * 128 * a.mt, n0 = blockIdx.z * 256;
  const int n_tile = min(256, a.N - n0);           // multiple of 64
  const int a_panels = 2 * a.mt, b_panels = n_tile / 64;
  const int stage_bytes = (a_panels + 4) * kWPanel;  // sized for a full 256-column tile of b
  WSmem* sb = reinterpret_cast<WSmem*>(smem + a.n_stages * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_begin = blockIdx.y * a.k_chunk;
  if (k_begin >= a.rows) return;  // (whole CTA: nothing to add)
  const int n_ksteps = (min(a.k_chunk, a.rows - k_begin) + kWK - 1) / kWK;
  const uint32_t tmem_cols = a.mt == 1 ? 256u : 512u;

  // The producer warp initialises the barriers itself and starts fetching at once (a short kernel: the first loads'
  // DRAM latency is a fifth of its life); everybody else meets it at a named barrier, after the TMEM allocation.
  if (warp == kWTmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < a.n_stages; ++s) {
        mbar_init(&sb->full[s], 1);
        mbar_init(&sb->empty[s], 1);
      }
      mbar_init(&sb->acc_full, 1);
      fence_barrier_init();
      prefetch_tensormap(&map_a);
      prefetch_tensormap(&map_b);
    }
    __syncwarp();
    named_bar_arrive(1, kWThreads);
  } else {
    if (warp == kWMmaWarp) {
      tmem_alloc(&sb->tmem_base, tmem_cols);
      tmem_relinquish();
    }
    tc_fence_before();
    named_bar_sync(1, kWThreads);
    tc_fence_after();
  }
  const uint32_t tmem = warp == kWTmaWarp ? 0u : sb->tmem_base;

  if (warp == kWTmaWarp) {
    for (int it = 0; it < n_ksteps; ++it) {
      const int s = it % a.n_stages;
      mbar_wait(&sb->empty[s], ((it / a.n_stages) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* st = smem + s * stage_bytes;
        const int krow = k_begin + it * kWK;  // rows past the end are zero-filled: they add nothing
        mbar_arrive_expect_tx(&sb->full[s], uint32_t((a_panels + b_panels) * kWPanel));
        for (int q = 0; q < a_panels; ++q) tma_load_2d(st + q * kWPanel, &map_a, &sb->full[s], m0 + 64 * q, krow);
        for (int q = 0; q < b_panels; ++q) tma_load_2d(st + (a_panels + q) * kWPanel, &map_b, &sb->full[s], n0 + 64 * q, krow);
      }
      __syncwarp();
    }
  } else if (warp == kWMmaWarp) {
    const uint32_t idesc = make_idesc_bf16(128, n_tile, 1, 1);
    for (int it = 0; it < n_ksteps; ++it) {
      const int s = it % a.n_stages;
      mbar_wait(&sb->full[s], (it / a.n_stages) & 1);
      tc_fence_after();
      const uint32_t st = smem_u32(smem + s * stage_bytes);
      const uint64_t da = make_desc_mn(st, kWPanel);
      const uint64_t da2 = make_desc_mn(st + 2 * kWPanel, kWPanel);  // second 128-column slice of a (mt == 2)
      const uint64_t db = make_desc_mn(st + a_panels * kWPanel, kWPanel);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kWK / 16; ++kk) {  // 16 contraction rows = 2048 bytes further into every panel
          const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
          mma_ss(tmem, da + uint64_t(kk * 128), db + uint64_t(kk * 128), idesc, acc);
          if (a.mt == 2) mma_ss(tmem + 256, da2 + uint64_t(kk * 128), db + uint64_t(kk * 128), idesc, acc);
        }
        tc_commit(&sb->empty[s]);
        if (it == n_ksteps - 1) tc_commit(&sb->acc_full);
      }
      __syncwarp();
    }
  } else {
    // ============================== epilogue: TMEM -> fp32 atomics into out (N, M) ==============================
    // TMEM lane = column of a = the CONTIGUOUS index of the output: for a fixed output row the 32 lanes of a warp add
    // 32 consecutive floats
    const int quad = warp & 3;  // TMEM lanes [32 quad, +32)
    const int chalf = (warp - kWEpiWarp0) >> 2;  // which half of the tile's columns
    const uint32_t lane_base = uint32_t(quad * 32);
    const int c_lo = chalf * (n_tile / 2), c_hi = c_lo + n_tile / 2;  // n_tile is a multiple of 64: halves of whole 32-column chunks
    mbar_wait(&sb->acc_full, 0);
    tc_fence_after();
    for (int t = 0; t < a.mt; ++t) {
      const int m = m0 + 128 * t + int(lane_base) + lane;
      if (m0 + 128 * t + int(lane_base) >= a.M) break;  // (warp-uniform: this quadrant is past the matrix)
      float* o = a.out + int64_t(n0) * a.M + m;
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + (lane_base << 16) + 256 * t + c0, v);
        tmem_wait_ld();
        if (m < a.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(o + int64_t(c0 + j) * a.M, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWMmaWarp) tmem_dealloc(tmem, tmem_cols);
}

PFN_cuTensorMapEncodeTiled w_encode_fn() {
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

// (rows, cols) bf16 row-major; box = 64 columns x 64 rows, SWIZZLE_128B, zero fill outside
int make_panel_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols) {
  ensure_context_in_this_thread();
  PFN_cuTensorMapEncodeTiled enc = w_encode_fn();
  if (!enc) return AGA_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(cols) * 2};
  cuuint32_t box[2] = {64, cuuint32_t(kWK)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? AGA_OK : AGA_ERR_INVALID_ARGUMENT;
}

}  // namespace
}  // namespace aga

using namespace aga;

extern "C" int aga_wgrad_bf16(const void* a, const void* b, int64_t rows, int M, int N, float* out, int transpose_out,
                              void* stream) {
  if (!a || !b || !out || rows <= 0 || M <= 0 || N <= 0) return AGA_ERR_INVALID_ARGUMENT;
  if (M % 64 != 0 || N % 64 != 0 || rows > 2147483647LL - 4096) return AGA_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) return AGA_ERR_UNSUPPORTED;
  // The kernel writes (b-columns, a-columns) row-major, i.e. (a^T b)^T: the (M, N) layout is the same kernel with the
  // operands exchanged — (b^T a)^T = a^T b.
  if (!transpose_out) {
    std::swap(a, b);
    std::swap(M, N);
  }
  CUtensorMap ma, mb;
  int st;
  if ((st = make_panel_map(&ma, a, rows, M)) != AGA_OK) return st;
  if ((st = make_panel_map(&mb, b, rows, N)) != AGA_OK) return st;
  static const int n_sm = []() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
  }();
  // two 128-column slices of a per CTA: every panel of b fetched from L2 feeds twice the MMAs
  const int mt = M > 128 ? 2 : 1;
  const int gx = (M + 128 * mt - 1) / (128 * mt), gz = (N + 255) / 256;
  const int stage_bytes = (2 * mt + 4) * kWPanel;
  const int n_stages = mt == 2 ? 3 : 4;  // 64 KiB (mt = 2) or 48 KiB per stage
  const size_t smem_bytes = 1024 + size_t(n_stages) * stage_bytes + sizeof(WSmem);
  const int want_splits = std::max(1, n_sm / (gx * gz));
  const int64_t k_chunk = ((rows + want_splits - 1) / want_splits + kWK - 1) / kWK * kWK;
  const int n_splits = int((rows + k_chunk - 1) / k_chunk);
  WArgs wa{int(rows), M, N, int(k_chunk), n_stages, mt, out};
  AGA_CUDA_TRY(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes)));
  wgrad_kernel<<<dim3(unsigned(gx), unsigned(n_splits), unsigned(gz)), kWThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(ma, mb, wa);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}
