"""Single-query (decoding step) attention kernel: time per call vs number of keys, both dtypes (CUDA-graph timed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import torch
import aga_b200 as A
from bench_cross import graph_time

H = 12
for dtype in (torch.bfloat16, torch.float32):
    for Tk in (8, 128, 512, 1500, 3000):
        q = torch.randn(1, 1, H * 64, device="cuda", dtype=dtype)
        kv = [torch.randn(1, Tk, 2 * H * 64, device="cuda", dtype=dtype) for _ in range(12)]  # 12 layers' caches in turn
        st = {"i": 0}
        def fn():
            st["i"] = (st["i"] + 1) % 12
            c = kv[st["i"]]
            return A.qkv_attention(q, c[..., :H * 64], c[..., H * 64:], H)
        with torch.no_grad():
            t = graph_time(fn, reps=24)
            t2 = graph_time(lambda: A.qkv_attention(q, kv[0][..., :H * 64], kv[0][..., H * 64:], H, impl="tcgen05") if dtype == torch.bfloat16
                            else A.qkv_attention(torch.cat([q, q], 1), kv[0][..., :H * 64], kv[0][..., H * 64:], H, impl="simt"), reps=24)
        print(f"{str(dtype):15s} Tk {Tk:5d}: single-query kernel {t * 1e3:6.1f} us   query-tile kernel {t2 * 1e3:6.1f} us")
