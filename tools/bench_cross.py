"""Cross-attention (Tq = 64, Tk = 1500) kernel timings: forward and backward, CUDA events, L2 flushed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import torch
import aga_b200 as A
if len(sys.argv) > 1:  # alternative build of the library
    from aga_b200 import _lib
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
from bench_attn import timeit


def graph_time(fn, reps=20):
    """GPU time per call with the CPU launch path taken out: `reps` calls captured in one CUDA graph."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    return sorted(ts)[2]

def main():
    for (B, H, Tq, Tk) in [(16, 12, 64, 1500), (16, 12, 128, 1500), (32, 16, 64, 1500), (16, 12, 64, 64)]:
        q = torch.randn(B, Tq, H * 64, device="cuda").bfloat16().requires_grad_()
        k = torch.randn(B, Tk, H * 64, device="cuda").bfloat16().requires_grad_()
        v = torch.randn(B, Tk, H * 64, device="cuda").bfloat16().requires_grad_()
        fl = 4.0 * B * H * Tq * Tk * 64
        ms = timeit(lambda: A.qkv_attention(q, k, v, H))
        with torch.no_grad():
            mg = graph_time(lambda: A.qkv_attention(q, k, v, H))
        print(f"fwd B{B} H{H} {Tq}x{Tk}: {ms * 1e3:.1f} us eager, {mg * 1e3:.1f} us in a graph  {fl / mg / 1e9:.1f} TFLOP/s")
        out, _, _ = A.qkv_attention(q, k, v, H)
        do = torch.randn_like(out)
        ms = timeit(lambda: torch.autograd.grad(out, (q, k, v), do, retain_graph=True), iters=7, warm=2)
        from aga_b200 import ops, _lib as L
        out, lse, _ = A.qkv_attention(q, k, v, H)
        qd, kd, vd, od = q.detach(), k.detach(), v.detach(), out.detach()
        dq, dk, dv = torch.empty_like(qd), torch.empty_like(kd), torch.empty_like(vd)
        cfg = (H, False, L.EXPORT_NONE, (0, 0), L.ATTN_AUTO, False)
        mg = graph_time(lambda: ops._attn_backward(qd, kd, vd, od, lse, None, torch.empty(0), cfg, do, None, dq, dk, dv))
        print(f"bwd B{B} H{H} {Tq}x{Tk}: {ms * 1e3:.1f} us eager, {mg * 1e3:.1f} us in a graph  {2.5 * fl / mg / 1e9:.1f} TFLOP/s")

if __name__ == "__main__":
    main()
