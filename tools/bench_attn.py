"""Micro-benchmark of the attention kernels (CUDA events, L2 flushed between iterations)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import aga_b200 as A

def timeit(fn, iters=10, warm=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

def main():
    for (B, H, Tq, Tk) in [(16, 12, 1500, 1500), (16, 12, 64, 1500), (32, 16, 1500, 1500)]:
        q = torch.randn(B, Tq, H * 64, device="cuda").bfloat16()
        k = torch.randn(B, Tk, H * 64, device="cuda").bfloat16()
        v = torch.randn(B, Tk, H * 64, device="cuda").bfloat16()
        fl = 4.0 * B * H * Tq * Tk * 64
        for impl in (["tcgen05", "simt"] if B == 16 else ["tcgen05"]):
            ms = timeit(lambda: A.qkv_attention(q, k, v, H, impl=impl))
            print(f"fwd {impl:8s} B{B} H{H} {Tq}x{Tk}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
        if "--bwd" in sys.argv:
            qg, kg, vg = (x.clone().requires_grad_() for x in (q, k, v))
            out, _, _ = A.qkv_attention(qg, kg, vg, H)
            do = torch.randn_like(out)
            ms = timeit(lambda: torch.autograd.grad(out, (qg, kg, vg), do, retain_graph=True), iters=5, warm=2)
            print(f"bwd auto     B{B} H{H} {Tq}x{Tk}: {ms:.3f} ms  {2.5 * fl / ms / 1e9:.1f} TFLOP/s")
        try:
            import torch.nn.functional as F
            qh, kh, vh = (x.view(B, -1, H, 64).transpose(1, 2) for x in (q, k, v))
            ms = timeit(lambda: F.scaled_dot_product_attention(qh, kh, vh))
            print(f"fwd torch-sdpa B{B} H{H} {Tq}x{Tk}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (informational)")
        except Exception as e:
            print("sdpa failed", e)

if __name__ == "__main__":
    main()
