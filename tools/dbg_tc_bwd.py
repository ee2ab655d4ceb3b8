"""Debug driver: tcgen05 attention backward vs the oracle (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import aga_oracle as O
import aga_b200 as A

def run(B, H, Tq, Tk, amp=1.0, seed=0, impl="tcgen05"):
    g = torch.Generator().manual_seed(seed)
    q = (amp * torch.randn(B, Tq, H * 64, generator=g)).bfloat16()
    k = (amp * torch.randn(B, Tk, H * 64, generator=g)).bfloat16()
    v = torch.randn(B, Tk, H * 64, generator=g).bfloat16()
    do = torch.randn(B, Tq, H * 64, generator=g).bfloat16()
    qd, kd, vd = (x.cuda().requires_grad_() for x in (q, k, v))
    try:
        out, lse, _ = A.qkv_attention(qd, kd, vd, H, impl=impl)
        out.backward(do.cuda())
        torch.cuda.synchronize()
    except Exception as e:
        print(f"B{B} H{H} Tq{Tq} Tk{Tk}: EXC {e}")
        return False
    dq, dk, dv = O.qkv_attention_bwd(q.float().numpy(), k.float().numpy(), v.float().numpy(), H, False, do.float().numpy())
    ok = True
    msg = []
    for name, got, ref in (("dq", qd.grad, dq), ("dk", kd.grad, dk), ("dv", vd.grad, dv)):
        gnp = got.float().cpu().numpy()
        err = np.abs(gnp - ref)
        scale = np.abs(ref).max()
        good = np.allclose(gnp, ref, rtol=2e-2, atol=2e-2 * scale)
        ok &= good
        msg.append(f"{name}: max|err|={err.max():.3g} (ref max {scale:.3g}) nan={np.isnan(gnp).sum()} {'ok' if good else 'BAD'}")
    print(f"B{B} H{H} Tq{Tq} Tk{Tk} amp{amp}: " + " | ".join(msg))
    return ok

if __name__ == "__main__":
    shapes = [(1, 1, 128, 128), (1, 1, 256, 128), (1, 1, 128, 256), (1, 2, 256, 384), (2, 3, 300, 200), (1, 2, 64, 1500),
              (1, 1, 1, 1), (1, 2, 1500, 1500)]
    res = [run(*s) for s in shapes]
    res.append(run(1, 2, 512, 640, amp=3.0, seed=1))
    print("ALL OK" if all(res) else "SOME FAILED")
