"""Every kernel of the library once on small ragged shapes (quick end-to-end exercise of all entry points on a GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import aga_b200 as A
from aga_b200 import ops

def main():
    g = torch.Generator().manual_seed(0)
    dev = "cuda"
    r = lambda *s: torch.randn(*s, generator=g)
    # log-mel (two kernels), ragged length
    A.log_mel_spectrogram((0.1 * r(2, 16000 + 37)).to(dev))
    A.log_mel_spectrogram((0.1 * r(1, 4000)).to(dev), n_mels=128)
    H = 2
    for (Tq, Tk, causal, export) in [(300, 200, False, None), (64, 333, False, None), (70, 70, True, "logits"), (200, 200, True, "logits"),
                                     (1, 1, False, None), (129, 257, False, None)]:
        for dt in (torch.bfloat16, torch.float32):
            q = r(2, Tq, H * 64).to(dt).to(dev).requires_grad_()
            k = r(2, Tk, H * 64).to(dt).to(dev).requires_grad_()
            v = r(2, Tk, H * 64).to(dt).to(dev).requires_grad_()
            out, lse, slab = A.qkv_attention(q, k, v, H, causal=causal, export=export, export_cols=(1, 3) if export else None)
            loss = out.float().sum()
            if slab is not None:
                loss = loss + torch.where(torch.isfinite(slab), slab, torch.zeros_like(slab)).sum()
            loss.backward()
    # guided loss / pattern / vote
    toks = torch.full((2, 24), 50257, dtype=torch.long)
    toks[:, :5] = torch.tensor([50258, 50260, 50259, 50359, 50363])
    lid = torch.zeros(51865, dtype=torch.uint8)
    pat = A.attention_pattern(toks.to(dev), lid, 0.6)
    slab = r(3, 2, H, 24, 2).to(dev).requires_grad_()
    A.guided_loss(slab, pat, torch.ones(3, H), n_early=1).backward()
    ops.head_vote(torch.softmax(r(3, 2, H, 24, 24), -1).to(dev))
    # LayerNorm family, adapter, CE, GEMMs
    for dt in (torch.bfloat16, torch.float32):
        for D in (384, 768, 1280):
            x = r(37, D).to(dt).to(dev).requires_grad_()
            w, b = torch.ones(D, device=dev, requires_grad=True), torch.zeros(D, device=dev, requires_grad=True)
            y, xr = ops.layer_norm_residual(x, w, b)
            (y.float().sum() + 2 * xr.float().sum()).backward()
            w1 = (r(D // 4, D) / 30).to(dev).requires_grad_(); b1 = torch.zeros(D // 4, device=dev, requires_grad=True)
            w2 = (r(D, D // 4) / 15).to(dev).requires_grad_(); b2 = torch.zeros(D, device=dev, requires_grad=True)
            x2 = r(37, D).to(dt).to(dev).requires_grad_()
            ops.adapter_layer_norm(x2, w1, b1, w2, b2, w, b).float().sum().backward()
            wl = (r(D, D) / 30).to(dt).to(dev); bl = r(D).to(dt).to(dev)
            ops.linear_residual(x2.detach(), wl, bl, x2.detach())
        V = 1003
        lg = r(3, 7, 1024).to(dt).to(dev).requires_grad_()
        tg = torch.randint(0, V, (3, 7), generator=g); tg[0, 4:] = -1
        loss, acc = ops.ls_cross_entropy(ops.VocabLogits(lg, V), tg.to(dev), -1, 0.1)
        loss.backward()
    for (M, K, N) in [(300, 384, 1536), (129, 768, 3072), (5, 64, 256)]:
        x = r(M, K).bfloat16().to(dev); w1 = (r(N, K) / 20).bfloat16().to(dev); b1 = r(N).bfloat16().to(dev)
        h, act = ops.gemm_gelu_fwd(x, w1, b1)
        ops.gemm_gelu_bwd(r(M, K).bfloat16().to(dev), (r(N, K) / 40).bfloat16().to(dev), h)
    torch.cuda.synchronize()
    print("all kernels ran")

if __name__ == "__main__":
    main()
