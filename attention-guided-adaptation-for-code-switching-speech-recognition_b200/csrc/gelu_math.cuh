// Exact-GELU arithmetic shared by the GEMM+GELU epilogue (gemm_gelu.cu) and the adapter's GELU backward (adapter.cu).
#pragma once
#include "tc_ptx.cuh"

namespace aga {

using ptx::ex2;

// erf through Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 rounding of the result), evaluated on PAIRS
// with the packed fp32x2 FMA / MUL / ADD of sm_100: two MUFU ops (rcp, ex2) and ~8 issue slots per element instead of
// erff's ~25.  (At 22 scalar instructions per element the epilogue of a 128 x 256 tile needs 5600 issue cycles per
// sub-partition against the 6144 cycles of its MMAs — the kernel was epilogue-issue-bound.)
//   cdf(x) = Phi(x) = 0.5 (1 + erf(x / sqrt 2)) = 0.5 + sign(x) (0.5 - 0.5 erfc(|x| / sqrt 2)),  e = exp(-x^2 / 2)
__device__ __forceinline__ void cdf_exp2(float2 x, float2& cdf, float2& e) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 z = __fmul2_rn(ax, make_float2(0.70710678118654752440f, 0.70710678118654752440f));
  const float2 d = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(d.y));
  // exp(-z^2) = 2^(-(z sqrt(log2 e))^2)
  const float2 w = __fmul2_rn(ax, make_float2(0.84932180028801904272f, 0.84932180028801904272f));
  const float2 m = __fmul2_rn(w, w);
  e = make_float2(ex2(-m.x), ex2(-m.y));
  // 0.5 * (a1 + a2 t + a3 t^2 + a4 t^3 + a5 t^4) (coefficients pre-scaled by 0.5)
  float2 p = __ffma2_rn(make_float2(0.5307027145f, 0.5307027145f), t, make_float2(-0.7265760135f, -0.7265760135f));
  p = __ffma2_rn(p, t, make_float2(0.7107068705f, 0.7107068705f));
  p = __ffma2_rn(p, t, make_float2(-0.142248368f, -0.142248368f));
  p = __ffma2_rn(p, t, make_float2(0.127414796f, 0.127414796f));
  const float2 he = __fmul2_rn(__fmul2_rn(p, t), e);                      // 0.5 erfc(|x| / sqrt 2)
  const float2 u = __fadd2_rn(make_float2(0.5f, 0.5f), make_float2(-he.x, -he.y));  // >= 0
  cdf = __fadd2_rn(make_float2(copysignf(u.x, x.x), copysignf(u.y, x.y)), make_float2(0.5f, 0.5f));
}
#ifdef AGA_GEMM_NOMATH  // experiment builds: the epilogue's transcendental work removed
__device__ __forceinline__ float2 gelu_erf2(float2 x) { return x; }
__device__ __forceinline__ float2 dgelu_erf2(float2 x) { return x; }
#else
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  float2 cdf, e;
  cdf_exp2(x, cdf, e);
  return __fmul2_rn(x, cdf);
}
__device__ __forceinline__ float2 dgelu_erf2(float2 x) {
  float2 cdf, e;
  cdf_exp2(x, cdf, e);
  return __ffma2_rn(x, __fmul2_rn(e, make_float2(0.39894228040143267794f, 0.39894228040143267794f)), cdf);
}
#endif
}  // namespace aga
