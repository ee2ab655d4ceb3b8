"""LayerNorm kernels at the encoder's shape (24000 rows x 768, bf16): every variant the training step launches, timed
inside a CUDA graph whose working set (8 rotating buffer sets, > 126 MB L2) keeps the data out of L2, with the
algorithmic bytes each moves."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import torch
import aga_b200  # noqa: F401
from aga_b200 import ops
from bench_cross import graph_time

def main():
    rows, D = 24000, 768
    nset = 8
    g = torch.Generator().manual_seed(0)
    mk = lambda: [torch.randn(rows, D, generator=g).bfloat16().cuda() for _ in range(nset)]
    x, r, dy, dres = mk(), mk(), mk(), mk()
    w = (1 + 0.1 * torch.randn(D, generator=g)).cuda()
    b = (0.1 * torch.randn(D, generator=g)).cuda()
    _, _, mean, rstd = ops._ln_fwd(x[0], None, w, b, 1e-5, False)
    state = {"i": 0}
    def rot():
        state["i"] = (state["i"] + 1) % nset
        return state["i"]
    row_bytes = rows * D * 2
    cases = [
        ("fwd plain            (x -> y)", 2, lambda i: ops._ln_fwd(x[i], None, w, b, 1e-5, False)),
        ("fwd residual + sum   (x, r -> y, s)", 4, lambda i: ops._ln_fwd(x[i], r[i], w, b, 1e-5, True)),
        ("bwd frozen           (dy, x -> dx)", 3, lambda i: ops._ln_bwd(dy[i], x[i], w, mean, rstd, False)),
        ("bwd frozen + dres    (dy, x, dres -> dx)", 4, lambda i: ops._ln_bwd(dy[i], x[i], w, mean, rstd, False, False, dres[i])),
        ("bwd params           (dy, x -> dx, dgamma, dbeta)", 3, lambda i: ops._ln_bwd(dy[i], x[i], w, mean, rstd, True)),
        ("bwd params + dxsum   (adapter LN)", 3, lambda i: ops._ln_bwd(dy[i], x[i], w, mean, rstd, True, True)),
        ("bwd params + dxsum + dres", 4, lambda i: ops._ln_bwd(dy[i], x[i], w, mean, rstd, True, True, dres[i])),
    ]
    with torch.no_grad():
        for name, n_t, fn in cases:
            t = graph_time(lambda: fn(rot()), reps=16)
            nbytes = n_t * row_bytes
            print(f"{name:52s} {t * 1e3:7.1f} us  {nbytes / 1e6:6.1f} MB  {nbytes / t / 1e9:7.2f} TB/s")

if __name__ == "__main__":
    main()
