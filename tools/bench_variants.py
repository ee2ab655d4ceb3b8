"""Time the attention forward/backward of alternative builds of the library (libaga_<tag>.so) on the BASELINE shape."""
import os, sys, glob, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attention-guided-adaptation-for-code-switching-speech-recognition_b200")
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path[:0] = [ROOT]
    import torch
    import aga_b200 as A
    from aga_b200 import _lib
    _lib.LIB_PATH = sys.argv[2]
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from bench_attn import timeit
    B, H, T = 16, 12, 1500
    q, k, v = (torch.randn(B, T, H * 64, device="cuda").bfloat16().requires_grad_() for _ in range(3))
    fl = 4.0 * B * H * T * T * 64
    ms = timeit(lambda: A.qkv_attention(q, k, v, H), iters=20)
    out, _, _ = A.qkv_attention(q, k, v, H)
    do = torch.randn_like(out)
    msb = timeit(lambda: torch.autograd.grad(out, (q, k, v), do, retain_graph=True), iters=10)
    print(f"{os.path.basename(sys.argv[2]):28s} fwd {ms:.4f} ms {fl/ms/1e9:7.1f} TFLOP/s | bwd {msb:.4f} ms {2.5*fl/msb/1e9:7.1f} TFLOP/s", flush=True)
else:
    for lib in sorted(glob.glob(os.path.join(PKG, "libaga_*.so"))):
        if "timeline" in lib: continue
        subprocess.run([sys.executable, __file__, "--one", lib])
