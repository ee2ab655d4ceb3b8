// Micro-benchmark: cycles per tcgen05.mma (M = 128, K = 16, bf16 -> fp32) for the operand forms the attention kernels
// use.  One elected thread issues `n` MMAs back to back, commits to an mbarrier and waits; clock64() brackets it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I<pkg>/csrc -o ubench_mma.bin ubench_mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

using namespace aga::ptx;

__device__ __forceinline__ uint64_t desc_mn2(uint32_t addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= uint64_t((addr & 0x3FFFF) >> 4);
  d |= uint64_t(lbo >> 4) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// mode: 0 SS K-major N=128 | 1 SS K-major N=64 | 2 TS (A in TMEM) B MN-major N=64 | 3 SS A MN-major (2 panels) B MN-major N=64
//       4 SS K-major N=256 | 5 TS B K-major N=128 | 6 SS K-major N=64 alternating two accumulators | 7 SS K-major N=32
template <int mode>
__global__ void __launch_bounds__(128) k_mma(int n, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {
    const uint64_t dA = make_smem_desc_sw128(smem_u32(smem));             // 128 rows x 128 B
    const uint64_t dB = make_smem_desc_sw128(smem_u32(smem + 32768));     // up to 256 rows x 128 B
    const uint64_t dAmn = desc_mn2(smem_u32(smem + 65536), 16384);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < n; i += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int kk = u & 3;
            if (mode == 0) mma_ss(tmem, dA + kk * 2, dB + kk * 2, make_idesc_bf16(128, 128, 0, 0), 1);
            if (mode == 1) mma_ss(tmem, dA + kk * 2, dB + kk * 2, make_idesc_bf16(128, 64, 0, 0), 1);
            if (mode == 2) mma_ts(tmem + 256, tmem + kk * 8, dB + kk * 128, make_idesc_bf16(128, 64, 0, 1), 1);
            if (mode == 3) mma_ss(tmem, dAmn + kk * 128, dB + kk * 128, make_idesc_bf16(128, 64, 1, 1), 1);
            if (mode == 4) mma_ss(tmem, dA + kk * 2, dB + kk * 2, make_idesc_bf16(128, 256, 0, 0), 1);
            if (mode == 5) mma_ts(tmem + 256, tmem + kk * 8, dB + kk * 2, make_idesc_bf16(128, 128, 0, 0), 1);
            if (mode == 6) mma_ss(tmem + (u & 4 ? 64 : 0), dA + kk * 2, dB + kk * 2, make_idesc_bf16(128, 64, 0, 0), 1);
            if (mode == 7) mma_ss(tmem, dA + kk * 2, dB + kk * 2, make_idesc_bf16(128, 32, 0, 0), 1);
          }
        }
        tc_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(k_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_mma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_mma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_mma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_mma<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_mma<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_mma<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[] = {"SS K-major N=128", "SS K-major N=64", "TS B MN-major N=64", "SS A MN-major(2 panels) B MN-major N=64",
                         "SS K-major N=256", "TS B K-major N=128", "SS K-major N=64, two accumulators", "SS K-major N=32"};
  for (int grid : {1}) {
    for (int mode = 0; mode < 8; ++mode) {
      for (int n : {32, 256}) {
        switch (mode) {
          case 0: k_mma<0><<<grid, 128, 100 * 1024>>>(n, out); break;
          case 1: k_mma<1><<<grid, 128, 100 * 1024>>>(n, out); break;
          case 2: k_mma<2><<<grid, 128, 100 * 1024>>>(n, out); break;
          case 3: k_mma<3><<<grid, 128, 100 * 1024>>>(n, out); break;
          case 4: k_mma<4><<<grid, 128, 100 * 1024>>>(n, out); break;
          case 5: k_mma<5><<<grid, 128, 100 * 1024>>>(n, out); break;
          case 6: k_mma<6><<<grid, 128, 100 * 1024>>>(n, out); break;
          default: k_mma<7><<<grid, 128, 100 * 1024>>>(n, out); break;
        }
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("grid %3d  %-42s n=%3d: %6lld cycles total, %.1f cycles/MMA (%s)\n", grid, names[mode], n, mx, double(mx) / n,
               cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
