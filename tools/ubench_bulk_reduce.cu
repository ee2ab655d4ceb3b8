// Micro-benchmark: how fast can the dQ partial tiles of the attention backward be accumulated in L2?
//   mode 0: red.global.add.v4.f32, one 16-byte piece per lane, lanes 3 KiB apart (what attn_bwd_tc_kernel did)
//   mode 1: red.global.add.v4.f32, lanes contiguous (512 B per warp instruction)
//   mode 2: cp.reduce.async.bulk.global.shared::cta.add.f32 of a contiguous 32 KiB tile staged in shared memory
// Every CTA walks `tiles_per_cta` destination tiles of 128 x 64 fp32 (12 CTAs share each destination, like the
// 12 key tiles of one (batch, head)).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_bulk_reduce ...
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(128) k_red(float* acc, int n_tiles, int tiles_per_cta, int mode, int row_stride) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* tile = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) tile[i] = 1.0f;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int first = (blockIdx.x / 12) * tiles_per_cta;  // 12 consecutive CTAs hit the same tiles
  for (int t = 0; t < tiles_per_cta; ++t) {
    const int tile_id = (first + t) % n_tiles;
    if (mode == 0) {
      float* dst = acc + (size_t(tile_id) * 128 + threadIdx.x) * row_stride;
#pragma unroll
      for (int e = 0; e < 16; ++e)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(dst + 4 * e), "f"(1.0f) : "memory");
    } else if (mode == 1) {
      float* dst = acc + size_t(tile_id) * 128 * 64;
#pragma unroll
      for (int e = 0; e < 16; ++e)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(dst + (e * 128 + threadIdx.x) * 4), "f"(1.0f)
                     : "memory");
    } else {
      if (threadIdx.x == 0) {
        float* dst = acc + size_t(tile_id) * 128 * 64;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
                     "r"(smem_u32(tile)), "r"(32768)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem reusable (what the kernel must wait for)
      }
      __syncthreads();
    }
  }
  if (mode == 2 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const int n_tiles = 16 * 12 * 12;  // (b, h, q-tile) of one Whisper-small layer
  const int tiles_per_cta = 12, grid = 16 * 12 * 12;
  float* acc;
  const size_t bytes = size_t(n_tiles) * 128 * 64 * 4;
  cudaMalloc(&acc, bytes * 12);  // mode 0 uses a 768-float row stride
  cudaMemset(acc, 0, bytes * 12);
  cudaFuncSetAttribute(k_red, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_red<<<grid, 128, 32768>>>(acc, n_tiles, tiles_per_cta, mode, mode == 0 ? 768 : 64);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double gb = double(grid) * tiles_per_cta * 32768 / 1e9;
      printf("mode %d rep %d: %.3f ms  %.1f GB/s of fp32 reduce payload (%s)\n", mode, rep, ms, gb / (ms / 1e3),
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  // correctness of mode 2: every element of tile 0 received 12 (CTAs) x 3 (reps) adds in mode 2 plus mode 1's
  float h[4];
  cudaMemcpy(h, acc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("acc[0..3] = %g %g %g %g\n", h[0], h[1], h[2], h[3]);
  return 0;
}
