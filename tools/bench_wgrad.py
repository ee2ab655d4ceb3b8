"""Adapter weight gradients (a^T b, rows = 24000): the tcgen05 kernel of the library vs torch.mm (cuBLAS split-K + reduce),
CUDA-graph timed over rotating operand sets (> L2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import torch
import aga_b200  # noqa: F401
from aga_b200 import ops
from bench_cross import graph_time

for rows, M, N in ((24000, 768, 192), (48000, 1024, 256), (12000, 1280, 320), (1024, 768, 192)):
    g = torch.Generator().manual_seed(0)
    sets = [(torch.randn(rows, M, generator=g).bfloat16().cuda(), torch.randn(rows, N, generator=g).bfloat16().cuda()) for _ in range(6)]
    st = {"i": 0}
    def nxt():
        st["i"] = (st["i"] + 1) % len(sets)
        return sets[st["i"]]
    out = torch.zeros(M * N, device="cuda")
    def ours(tr):
        a, b = nxt()
        ops.L.torch_ops().wgrad(a, b, out, tr)
    def cublas(tr):
        a, b = nxt()
        return torch.mm(b.t(), a, out_dtype=torch.float32) if tr else torch.mm(a.t(), b, out_dtype=torch.float32)
    mb = rows * (M + N) * 2 / 1e6
    for tr in (False, True):
        t1 = graph_time(lambda: ours(tr), reps=12)
        t2 = graph_time(lambda: cublas(tr), reps=12)
        print(f"rows {rows:6d} M {M:5d} N {N:4d} {'(N,M) out' if tr else '(M,N) out'}: tcgen05 {t1 * 1e3:6.1f} us ({mb / t1 / 1e3:5.2f} TB/s)   "
              f"torch.mm {t2 * 1e3:6.1f} us")
