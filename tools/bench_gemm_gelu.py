"""MLP GEMM + GELU: the tcgen05 GEMM with the erf-GELU epilogue vs cuBLAS + separate GELU kernels (graph-timed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import torch
import torch.nn.functional as F
import aga_b200  # noqa: F401
if len(sys.argv) > 1:
    from aga_b200 import _lib
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
from aga_b200 import ops
from bench_cross import graph_time

def main():
    M, K, N = 24000, 768, 3072
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, K, generator=g).bfloat16().cuda()
    w1 = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().cuda()
    b1 = torch.randn(N, generator=g).bfloat16().cuda()
    w2 = (torch.randn(K, N, generator=g) / N ** 0.5).bfloat16().cuda()
    w2t = w2.t().contiguous()
    dy = torch.randn(M, K, generator=g).bfloat16().cuda()
    h = F.linear(x, w1, b1)
    fl = 2.0 * M * K * N
    from aga_b200 import _lib
    import ctypes
    lib = ctypes.CDLL(os.environ.get("LD_PRELOAD") or _lib.LIB_PATH)  # the instance the extension's calls resolve to
    ref_g = F.gelu(h.float()).bfloat16()
    with torch.no_grad():
        print("quad clusters resident:", lib.aga_debug_gemm_quad_clusters())
        for variant, name in ((1, "multicast pair"), (2, "two-SM MMA"), (3, "multicast quad")):
            lib.aga_debug_set_gemm_variant(variant)
            hh, gg = ops.gemm_gelu_fwd(x, w1, b1)
            err = (gg.float() - ref_g.float()).abs().max().item()
            t = graph_time(lambda: ops.gemm_gelu_fwd(x, w1, b1))
            print(f"fwd fused [{name:14s}]: {t * 1e3:7.1f} us  {fl / t / 1e9:7.1f} TFLOP/s   max|g - ref| {err:.3e}")
            t = graph_time(lambda: ops.gemm_gelu_bwd(dy, w2t, h))
            print(f"bwd fused [{name:14s}]: {t * 1e3:7.1f} us  {fl / t / 1e9:7.1f} TFLOP/s")
        t = graph_time(lambda: ops.gemm_gelu_fwd(x, w1, b1))
        print(f"fwd fused   : {t * 1e3:7.1f} us  {fl / t / 1e9:7.1f} TFLOP/s")
        t = graph_time(lambda: F.gelu(F.linear(x, w1, b1)))
        print(f"fwd cuBLAS+gelu: {t * 1e3:7.1f} us")
        t = graph_time(lambda: F.linear(x, w1, b1))
        print(f"fwd cuBLAS alone: {t * 1e3:7.1f} us  {fl / t / 1e9:7.1f} TFLOP/s")
        t = graph_time(lambda: ops.gemm_gelu_bwd(dy, w2t, h))
        print(f"bwd fused   : {t * 1e3:7.1f} us  {fl / t / 1e9:7.1f} TFLOP/s")
        t = graph_time(lambda: torch.ops.aten.gelu_backward(dy @ w2, h))
        print(f"bwd cuBLAS+gelu_backward: {t * 1e3:7.1f} us")

if __name__ == "__main__":
    main()
