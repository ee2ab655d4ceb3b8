// out = x W^T + bias + residual in ONE cuBLASLt GEMM ("out-proj + residual").
//
// Reference: ResidualAttentionBlock.forward, whisper/whisper/model.py:231-242 — `x = x + self.attn(...)`,
// `x = x + self.mlp(...)`: a Linear (GEMM with a bias epilogue) followed by a separate full-size add kernel that
// re-reads the GEMM's result and the residual stream.  cuBLASLt can read C (the residual) and write a different D in
// the same GEMM with the bias epilogue (beta = 1); PyTorch's addmm only reaches that path by first copying the
// residual into the output.  This is a plain library GEMM (cuBLAS is the right tool for it); the library is resolved
// with dlopen so that libaga_b200.so has no link-time dependency on it (the process — PyTorch — has it loaded).
#include "aga_common.cuh"

#include <cublasLt.h>
#include <dlfcn.h>

#include <map>
#include <mutex>
#include <tuple>

namespace aga {
namespace {

struct LtApi {
  void* lib = nullptr;
  decltype(&cublasLtCreate) create = nullptr;
  decltype(&cublasLtMatmul) matmul = nullptr;
  decltype(&cublasLtMatmulDescCreate) desc_create = nullptr;
  decltype(&cublasLtMatmulDescDestroy) desc_destroy = nullptr;
  decltype(&cublasLtMatmulDescSetAttribute) desc_set = nullptr;
  decltype(&cublasLtMatrixLayoutCreate) layout_create = nullptr;
  decltype(&cublasLtMatrixLayoutDestroy) layout_destroy = nullptr;
  decltype(&cublasLtMatmulPreferenceCreate) pref_create = nullptr;
  decltype(&cublasLtMatmulPreferenceDestroy) pref_destroy = nullptr;
  decltype(&cublasLtMatmulPreferenceSetAttribute) pref_set = nullptr;
  decltype(&cublasLtMatmulAlgoGetHeuristic) heuristic = nullptr;
  bool ok = false;
};

const LtApi& lt_api() {
  static LtApi api = []() {
    LtApi a;
    const char* names[] = {"libcublasLt.so.12", "libcublasLt.so"};
    for (const char* n : names) {
      a.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD);  // the copy PyTorch already loaded
      if (a.lib) break;
    }
    for (const char* n : names) {
      if (a.lib) break;
      a.lib = dlopen(n, RTLD_NOW);
    }
    if (!a.lib) return a;
#define AGA_LT_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, name))
    AGA_LT_SYM(create, "cublasLtCreate");
    AGA_LT_SYM(matmul, "cublasLtMatmul");
    AGA_LT_SYM(desc_create, "cublasLtMatmulDescCreate");
    AGA_LT_SYM(desc_destroy, "cublasLtMatmulDescDestroy");
    AGA_LT_SYM(desc_set, "cublasLtMatmulDescSetAttribute");
    AGA_LT_SYM(layout_create, "cublasLtMatrixLayoutCreate");
    AGA_LT_SYM(layout_destroy, "cublasLtMatrixLayoutDestroy");
    AGA_LT_SYM(pref_create, "cublasLtMatmulPreferenceCreate");
    AGA_LT_SYM(pref_destroy, "cublasLtMatmulPreferenceDestroy");
    AGA_LT_SYM(pref_set, "cublasLtMatmulPreferenceSetAttribute");
    AGA_LT_SYM(heuristic, "cublasLtMatmulAlgoGetHeuristic");
#undef AGA_LT_SYM
    a.ok = a.create && a.matmul && a.desc_create && a.desc_destroy && a.desc_set && a.layout_create && a.layout_destroy &&
           a.pref_create && a.pref_destroy && a.pref_set && a.heuristic;
    return a;
  }();
  return api;
}

struct Plan {
  cublasLtMatmulDesc_t op = nullptr;
  cublasLtMatrixLayout_t a = nullptr, b = nullptr, c = nullptr, d = nullptr;
  cublasLtMatmulAlgo_t algo;
  bool valid = false;
};

struct Ctx {
  std::mutex mu;
  std::map<int, cublasLtHandle_t> handles;  // per device
  std::map<std::tuple<int, int, int64_t, int, int, int, size_t>, Plan> plans;  // (device, dtype | layout, rows, N, K, bias, ws)
};
Ctx& ctx() {
  static Ctx c;
  return c;
}

}  // namespace
}  // namespace aga

using namespace aga;

extern "C" int aga_linear_residual_workspace_bytes(size_t* bytes) {
  if (!bytes) return AGA_ERR_INVALID_ARGUMENT;
  *bytes = size_t(32) << 20;  // what the heuristic may use (Blackwell kernels want up to 32 MiB)
  return AGA_OK;
}

// out (rows, N) = x (rows, K) @ W + bias (N) + residual (rows, N); all row-major, contiguous, one dtype.
// w_layout 0: w is (N, K) — an nn.Linear weight, W = w^T; 1: w is (K, N), W = w (the dgrad form dY @ weight).
extern "C" int aga_linear_residual(const void* x, const void* w, int w_layout, const void* bias, const void* residual,
                                   void* out, int dtype, int64_t rows, int N, int K, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (!x || !w || !residual || !out || rows <= 0 || N <= 0 || K <= 0) return AGA_ERR_INVALID_ARGUMENT;
  if ((dtype != AGA_F32 && dtype != AGA_BF16) || (w_layout != 0 && w_layout != 1)) return AGA_ERR_INVALID_ARGUMENT;
  const LtApi& lt = lt_api();
  if (!lt.ok) return AGA_ERR_UNSUPPORTED;
  const uintptr_t all = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(residual) |
                        reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(workspace);
  if (all & 15) return AGA_ERR_UNSUPPORTED;
  int dev = 0;
  AGA_CUDA_TRY(cudaGetDevice(&dev));
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lock(c.mu);
  cublasLtHandle_t& handle = c.handles[dev];
  if (!handle && lt.create(&handle) != CUBLAS_STATUS_SUCCESS) return AGA_ERR_CUDA;
  const cudaDataType_t dt = dtype == AGA_BF16 ? CUDA_R_16BF : CUDA_R_32F;
  const auto key = std::make_tuple(dev, dtype * 2 + w_layout, rows, N, K, bias ? 1 : 0, workspace_bytes);
  Plan& p = c.plans[key];
  if (!p.valid) {
    // column-major view: D^T (N x rows) = A (N x K) * x^T (K x rows) + C^T.  w row-major (N,K) = col-major (K,N), ld K -> op T;
    // w row-major (K,N) = col-major (N,K), ld N -> op N
    if (lt.desc_create(&p.op, CUBLAS_COMPUTE_32F, CUDA_R_32F) != CUBLAS_STATUS_SUCCESS) return AGA_ERR_CUDA;
    const cublasOperation_t ta = w_layout == 0 ? CUBLAS_OP_T : CUBLAS_OP_N, tb = CUBLAS_OP_N;
    lt.desc_set(p.op, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta));
    lt.desc_set(p.op, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb));
    const cublasLtEpilogue_t epi = bias ? CUBLASLT_EPILOGUE_BIAS : CUBLASLT_EPILOGUE_DEFAULT;
    lt.desc_set(p.op, CUBLASLT_MATMUL_DESC_EPILOGUE, &epi, sizeof(epi));
    if (bias) lt.desc_set(p.op, CUBLASLT_MATMUL_DESC_BIAS_DATA_TYPE, &dt, sizeof(dt));
    bool ok = (w_layout == 0 ? lt.layout_create(&p.a, dt, K, N, K) : lt.layout_create(&p.a, dt, N, K, N)) == CUBLAS_STATUS_SUCCESS &&
              lt.layout_create(&p.b, dt, K, rows, K) == CUBLAS_STATUS_SUCCESS &&
              lt.layout_create(&p.c, dt, N, rows, N) == CUBLAS_STATUS_SUCCESS &&
              lt.layout_create(&p.d, dt, N, rows, N) == CUBLAS_STATUS_SUCCESS;
    if (!ok) return AGA_ERR_CUDA;
    if (bias) lt.desc_set(p.op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias));  // needed by the heuristic's checks
    cublasLtMatmulPreference_t pref = nullptr;
    if (lt.pref_create(&pref) != CUBLAS_STATUS_SUCCESS) return AGA_ERR_CUDA;
    lt.pref_set(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &workspace_bytes, sizeof(workspace_bytes));
    cublasLtMatmulHeuristicResult_t res;
    int found = 0;
    const cublasStatus_t hs = lt.heuristic(handle, p.op, p.a, p.b, p.c, p.d, pref, 1, &res, &found);
    lt.pref_destroy(pref);
    if (hs != CUBLAS_STATUS_SUCCESS || found == 0) return AGA_ERR_UNSUPPORTED;
    p.algo = res.algo;
    p.valid = true;
  }
  if (bias) lt.desc_set(p.op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias));
  const float alpha = 1.0f, beta = 1.0f;
  const cublasStatus_t st = lt.matmul(handle, p.op, &alpha, w, p.a, x, p.b, &beta, residual, p.c, out, p.d, &p.algo, workspace,
                                      workspace_bytes, static_cast<cudaStream_t>(stream));
  if (st != CUBLAS_STATUS_SUCCESS) return AGA_ERR_CUDA;
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return AGA_OK;
}
