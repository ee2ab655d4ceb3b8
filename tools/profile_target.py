"""Short program for ncu: a few launches of each hot kernel at the BASELINE shapes (B=16, Whisper-small)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import aga_b200 as A

def main():
    B, H, T = 16, 12, 1500
    g = torch.Generator().manual_seed(0)
    audio = (0.1 * torch.randn(B, 480000, generator=g)).cuda()
    q, k, v = (torch.randn(B, T, H * 64, generator=g).bfloat16().cuda().requires_grad_() for _ in range(3))
    do = torch.randn(B, T, H * 64, generator=g).bfloat16().cuda()
    # decoder cross attention (Tq = 64: the one-query-tile backward kernel)
    qc = torch.randn(B, 64, H * 64, generator=g).bfloat16().cuda().requires_grad_()
    doc = torch.randn(B, 64, H * 64, generator=g).bfloat16().cuda()
    from aga_b200 import ops
    xm = torch.randn(B * T, H * 64, generator=g).bfloat16().cuda()
    w1 = (torch.randn(4 * H * 64, H * 64, generator=g) / 28).bfloat16().cuda()
    b1 = torch.randn(4 * H * 64, generator=g).bfloat16().cuda()
    w2t = (torch.randn(4 * H * 64, H * 64, generator=g) / 55).bfloat16().cuda()
    for _ in range(3):
        hm, _ = ops.gemm_gelu_fwd(xm, w1, b1)          # the MLP GEMM with the GELU epilogue, forward ...
        ops.gemm_gelu_bwd(xm, w2t, hm)                 # ... and backward
        A.log_mel_spectrogram(audio)
        out, _, _ = A.qkv_attention(q, k, v, H)
        out.backward(do)
        oc, _, _ = A.qkv_attention(qc, k, v, H)
        oc.backward(doc)
    torch.cuda.synchronize()
    print("ok")

if __name__ == "__main__":
    main()
