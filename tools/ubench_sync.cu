// Micro-benchmark: cost (cycles, one warp timing itself, 4 or 8 warps running the same sequence) of the synchronisation
// and TMEM primitives that sit on the critical chain of the attention kernels' softmax warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I<pkg>/csrc -o ubench_sync.bin ubench_sync.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

using namespace aga::ptx;

constexpr int kReps = 64;
enum { LD32 = 0, LD32x2, ST16, ST32, FENCE_BEFORE, FENCE_AFTER, FENCE_PROXY, STS4_FENCE_PROXY, MBAR_ARRIVE, MBAR_WAIT_DONE, NAMED_BAR,
       PUBLISH_TMEM, PUBLISH_TMEM_SMEM, CLOCK_ONLY, N_TESTS };
const char* kNames[N_TESTS] = {"tcgen05.ld 32x32b.x32 + wait::ld", "2 x tcgen05.ld x32 + wait::ld", "tcgen05.st x16 + wait::st",
                               "tcgen05.st x32 + wait::st", "tcgen05.fence::before_thread_sync", "tcgen05.fence::after_thread_sync",
                               "fence.proxy.async.shared::cta", "4 x st.shared.v4 + fence.proxy.async", "mbarrier.arrive (lane 0)",
                               "mbarrier.try_wait on a completed phase", "bar.sync (all warps of the CTA)",
                               "publish: st x16, wait::st, fence::before, syncwarp, arrive",
                               "publish: st x16, 4 sts, fence.proxy, wait::st, fence::before, syncwarp, arrive", "clock64 pair only"};

__global__ void __launch_bounds__(256) k_sync(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) mbar_arrive(&done_bar);  // phase 0 of done_bar is complete from now on
  __syncthreads();
  const uint32_t t = tmem_slot + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  const uint32_t srow = smem_u32(smem) + threadIdx.x * 128;
  uint32_t r[32], r2[32];
  for (int i = 0; i < 32; ++i) r[i] = i + lane, r2[i] = 0;
  for (int test = 0; test < N_TESTS; ++test) {
    __syncthreads();
    long long total = 0;
    for (int rep = 0; rep < kReps; ++rep) {
      __syncwarp();
      const long long t0 = clock64();
      switch (test) {
        case LD32: tmem_ld32(t, r2); tmem_wait_ld(); break;
        case LD32x2: tmem_ld32(t, r2); tmem_ld32(t + 32, r); tmem_wait_ld(); break;
        case ST16: tmem_st16(t, reinterpret_cast<uint32_t(&)[16]>(r)); tmem_wait_st(); break;
        case ST32: tmem_st32(t, r); tmem_wait_st(); break;
        case FENCE_BEFORE: tc_fence_before(); break;
        case FENCE_AFTER: tc_fence_after(); break;
        case FENCE_PROXY: fence_proxy_async_smem(); break;
        case STS4_FENCE_PROXY:
          for (int q = 0; q < 4; ++q) sts128(srow + ((q ^ (lane & 7)) * 16), r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
          fence_proxy_async_smem();
          break;
        case MBAR_ARRIVE: if (lane == 0) mbar_arrive(&bars[warp]); break;
        case MBAR_WAIT_DONE: mbar_wait(&done_bar, 0); break;
        case NAMED_BAR: named_bar_sync(1, blockDim.x); break;
        case PUBLISH_TMEM:
          tmem_st16(t, reinterpret_cast<uint32_t(&)[16]>(r)); tmem_wait_st(); tc_fence_before(); __syncwarp();
          if (lane == 0) mbar_arrive(&bars[warp]);
          break;
        case PUBLISH_TMEM_SMEM:
          tmem_st16(t, reinterpret_cast<uint32_t(&)[16]>(r));
          for (int q = 0; q < 4; ++q) sts128(srow + ((q ^ (lane & 7)) * 16), r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
          fence_proxy_async_smem(); tmem_wait_st(); tc_fence_before(); __syncwarp();
          if (lane == 0) mbar_arrive(&bars[warp]);
          break;
        default: break;
      }
      __syncwarp();
      total += clock64() - t0;
      for (int i = 0; i < 32; ++i) r[i] += r2[i] & 1;
    }
    if (lane == 0) out[test * 8 + warp] = total / kReps;
  }
  uint32_t acc = 0;
  for (int i = 0; i < 32; ++i) acc += r[i];
  if (acc == 0x12345678u) out[1000] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, 512);
  (void)nw;
}

int main() {
  long long* out;
  cudaMalloc(&out, 2048 * sizeof(long long));
  cudaFuncSetAttribute(k_sync, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int warps : {1, 4, 8}) {
    cudaMemset(out, 0, 2048 * sizeof(long long));
    k_sync<<<1, warps * 32, 64 * 1024>>>(out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[N_TESTS * 8];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("---- %d warp(s) (%s)\n", warps, cudaGetErrorString(e));
    for (int t = 0; t < N_TESTS; ++t) {
      long long mx = 0;
      for (int w = 0; w < warps; ++w) mx = h[t * 8 + w] > mx ? h[t * 8 + w] : mx;
      printf("  %-82s %5lld cycles (warp 0: %lld)\n", kNames[t], mx, h[t * 8]);
    }
  }
  return 0;
}
