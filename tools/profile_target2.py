"""Short program for ncu: the kernels round 2 added, at the Whisper-small training shapes (24000 rows x 768) / one decoding
step — adapter weight gradients, LayerNorm pair forward, LayerNorm ring backward, adapter GELU backward, the two optimizer
kernels, the single-query attention kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import aga_b200 as A
from aga_b200 import ops, _lib


def main():
    rows, D, Bn, H = 24000, 768, 192, 12
    g = torch.Generator().manual_seed(0)
    mk = lambda *s: torch.randn(*s, generator=g).bfloat16().cuda()
    x, r, dy, ds = mk(rows, D), mk(rows, D), mk(rows, D), mk(rows, D)
    gact, dg, h1 = mk(rows, Bn), mk(rows, Bn), mk(rows, Bn)
    w = (1 + 0.1 * torch.randn(D, generator=g)).cuda()
    b = (0.1 * torch.randn(D, generator=g)).cuda()
    t = _lib.torch_ops()
    n = 14_275_584
    p, gr, m, v = (torch.randn(n, device="cuda") * s for s in (1.0, 1e-3, 1e-3, 1e-3))
    v.abs_()
    step, norm, lr = torch.zeros((), device="cuda"), torch.zeros((), device="cuda"), torch.full((), 1e-3, device="cuda")
    ws = torch.zeros(2048, dtype=torch.int64, device="cuda")
    shadow = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    out = torch.zeros(D * Bn, device="cuda")
    q1 = mk(1, 1, H * 64)
    kv = mk(1, 1500, 2 * H * 64)
    for _ in range(3):
        t.wgrad(ds, gact, out, False)                    # dW2 = ds^T g
        t.wgrad(x, gact, out, True)                      # dW1 = (x^T dh1)^T
        z, s, mean, rstd, y2, mean2, rstd2 = t.layernorm_pair_fwd(x, r, w, b, 1e-5, True, w, b, 1e-5)
        ops._ln_bwd(dy, s, w, mean, rstd, True, True)    # parameter gradients + dxsum: the cp.async ring kernel
        ops.gelu_bwd_colsum(dg, h1)
        t.flat_grad_norm(gr, norm, step, None, ws)
        t.flat_adamw(p, gr, m, v, lr, 0.9, 0.99, 1e-6, 0.01, step, norm, 1.0, shadow)
        A.qkv_attention(q1, kv[..., :H * 64], kv[..., H * 64:], H)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
