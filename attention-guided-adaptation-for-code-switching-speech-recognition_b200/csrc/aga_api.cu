// Library-wide C ABI helpers (status strings, counters, device check).
#include "aga_common.cuh"

namespace aga {
thread_local int g_last_cuda_error = 0;
std::atomic<uint64_t> g_launch_count{0};
}  // namespace aga

extern "C" int aga_version(void) { return AGA_B200_VERSION; }

extern "C" const char* aga_status_str(int status) {
  switch (status) {
    case AGA_OK: return "AGA_OK";
    case AGA_ERR_INVALID_ARGUMENT: return "AGA_ERR_INVALID_ARGUMENT";
    case AGA_ERR_UNSUPPORTED: return "AGA_ERR_UNSUPPORTED";
    case AGA_ERR_CUDA: return "AGA_ERR_CUDA";
    case AGA_ERR_WORKSPACE_TOO_SMALL: return "AGA_ERR_WORKSPACE_TOO_SMALL";
    case AGA_ERR_NO_SM100: return "AGA_ERR_NO_SM100";
    default: return "AGA_ERR_UNKNOWN";
  }
}

extern "C" int aga_last_cuda_error(void) { return aga::g_last_cuda_error; }

extern "C" uint64_t aga_launch_count(void) { return aga::g_launch_count.load(std::memory_order_relaxed); }

extern "C" int aga_device_is_sm100(int dev) {
  int major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return aga::cuda_fail(e);
  return major == 10 ? AGA_OK : AGA_ERR_NO_SM100;
}
