"""Debug driver: tcgen05 attention forward vs the oracle on a ladder of shapes (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import aga_oracle as O
import aga_b200 as A

def run(B, H, Tq, Tk, amp=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    q = (amp * torch.randn(B, Tq, H * 64, generator=g)).bfloat16()
    k = (amp * torch.randn(B, Tk, H * 64, generator=g)).bfloat16()
    v = torch.randn(B, Tk, H * 64, generator=g).bfloat16()
    try:
        out, lse, _ = A.qkv_attention(q.cuda(), k.cuda(), v.cuda(), H, impl="tcgen05")
        torch.cuda.synchronize()
    except Exception as e:
        print(f"B{B} H{H} Tq{Tq} Tk{Tk}: EXC {e}")
        return False
    ref, qk, _ = O.qkv_attention(q.float().numpy(), k.float().numpy(), v.float().numpy(), H)
    mx = qk.max(-1); lse_ref = mx + np.log(np.exp(qk - mx[..., None]).sum(-1))
    o = out.float().cpu().numpy()
    err = np.abs(o - ref)
    lerr = np.abs(lse.cpu().numpy() - lse_ref).max()
    ok = np.allclose(o, ref, rtol=2e-2, atol=2e-2) and lerr < 2e-2
    print(f"B{B} H{H} Tq{Tq} Tk{Tk} amp{amp}: max|err|={err.max():.4g} mean={err.mean():.3g} lse_err={lerr:.3g} nan={np.isnan(o).sum()} {'OK' if ok else 'FAIL'}")
    if not ok:
        bad = np.argwhere(err > 2e-2 + 2e-2 * np.abs(ref))
        print("   first bad idx:", bad[:5].tolist(), "n_bad", len(bad), "of", err.size)
        print("   out[0,0,:8]", o[0, 0, :8], "\n   ref[0,0,:8]", ref[0, 0, :8])
    return ok

if __name__ == "__main__":
    shapes = [(1, 1, 128, 128), (1, 1, 256, 128), (1, 1, 128, 256), (1, 2, 256, 384), (2, 3, 300, 200), (1, 2, 64, 1500),
              (1, 1, 1, 1), (1, 2, 1500, 1500)]
    res = [run(*s) for s in shapes]
    res.append(run(1, 2, 512, 640, amp=4.0, seed=1))  # peaky logits exercise the lazy rescale
    print("ALL OK" if all(res) else "SOME FAILED")
