// Micro-benchmark: tcgen05.ld throughput per SM as a function of the number of warps issuing loads, and how much a
// concurrent stream of tcgen05.mma (M128 N64 K16, SS) slows them down (and vice versa).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I<pkg>/csrc -o ubench_tmem.bin ubench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

using namespace aga::ptx;

constexpr int kIters = 256;

// mode 0: loads only; mode 1: loads + MMA stream from warp nw (one extra warp); mode 2: MMA stream only
__global__ void __launch_bounds__(1024) k_tmem(long long* out, int n_ld_warps, int mode, int mma_n) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0;
  uint32_t acc = 0;
  if (warp < n_ld_warps && mode != 2) {
    const uint32_t t = tmem + (uint32_t((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 32;
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
      uint32_t r[32];
      tmem_ld32(t, r);
      tmem_wait_ld();
      acc ^= r[it & 31];
    }
    t1 = clock64();
  } else if (warp == n_ld_warps && mode != 0) {
    // MMA stream: M128 x N x K16, A/B from smem (zeros), accumulate into columns [256, 256 + N)
    const uint32_t idesc = make_idesc_bf16(128, mma_n, 0, 0);
    const uint64_t da = make_smem_desc_sw128(smem_u32(smem)), db = make_smem_desc_sw128(smem_u32(smem + 16384));
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < kIters / 8; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) mma_ss(tmem + 256, da + uint64_t((kk & 3) * 2), db + uint64_t((kk & 3) * 2), idesc, 1);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
  }
  if (lane == 0) {
    out[warp * 2] = t1 - t0;
    out[warp * 2 + 1] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * 2 * sizeof(long long));
  cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mma_n : {64, 128}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int nw : {1, 4, 8, 16}) {
        if (mode == 2 && nw != 4) continue;
        if (mode == 0 && mma_n == 128) continue;
        cudaMemset(d, 0, 64 * 2 * sizeof(long long));
        k_tmem<<<1, (nw + 1) * 32, 64 * 1024>>>(d, nw, mode, mma_n);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[128];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < nw; ++w) mx = h[2 * w] > mx ? h[2 * w] : mx;
        const double bytes = double(nw) * kIters * 4096.0;
        printf("N=%3d mode %d  ld warps %2d: ", mma_n, mode, nw);
        if (mode != 2) printf("loads %lld clk (%.1f clk per ld per warp, %.1f B/clk/SM)  ", mx, double(mx) / kIters, bytes / mx);
        if (mode != 0) printf("mma stream %lld clk (%.1f clk per MMA)", h[2 * nw], double(h[2 * nw]) / kIters);
        printf("  [%s]\n", cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
