"""Where the GEMM+GELU kernel's warps wait (debug build: make -C csrc timeline): cycles inside mbarrier waits of the TMA
producer, the MMA issuer and the first epilogue warp, averaged over CTAs, per cluster variant.
  LD_PRELOAD=$PWD/<pkg>/libaga_b200_timeline.so python tools/gemm_waits.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import aga_b200  # noqa: F401
from aga_b200 import ops, _lib

lib = C.CDLL(os.path.join(os.path.dirname(_lib.LIB_PATH), "libaga_b200_timeline.so"))
lib.aga_debug_set_gemm_waits.argtypes = [C.c_void_p]
M, K, N = 24000, 768, 3072
g = torch.Generator().manual_seed(0)
x = torch.randn(M, K, generator=g).bfloat16().cuda()
w1 = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().cuda()
b1 = torch.randn(N, generator=g).bfloat16().cuda()
for variant in (1, 2, 3):
    lib.aga_debug_set_gemm_variant(variant)
    for _ in range(3):
        ops.gemm_gelu_fwd(x, w1, b1)
    buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    lib.aga_debug_set_gemm_waits(C.c_void_p(buf.data_ptr()))
    ops.gemm_gelu_fwd(x, w1, b1)
    torch.cuda.synchronize()
    lib.aga_debug_set_gemm_waits(C.c_void_p(0))
    t = buf.cpu().view(148, 16).double()
    t = t[t[:, 8] > 0]
    m = t.mean(0)
    print(f"variant {variant}: CTAs {len(t)}  producer: wait-empty {m[0]:9.0f} of {m[1]:9.0f} | mma: wait-acc_empty {m[2]:9.0f} "
          f"wait-full {m[3]:9.0f} of {m[4]:9.0f} | epilogue warp 0: wait-acc_full {m[5]:9.0f} wait-store-read {m[6]:9.0f} tmem-ld {m[7]:9.0f} of {m[8]:9.0f}")
    lead = t[t[:, 4] > 0]
    print(f"           (MMA-issuing CTAs: {len(lead)}; their mma wait-full share {float((lead[:, 3] / lead[:, 4]).mean()):.3f}, "
          f"wait-acc_empty share {float((lead[:, 2] / lead[:, 4]).mean()):.3f})")
