"""Sustained-loop probe: SM clock / power while the attention forward (or backward) runs back to back."""
import os, sys, subprocess, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import aga_b200 as A

def sample(stop, rows):
    while not stop.is_set():
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
        rows.append(out)
        time.sleep(0.1)

def run(name, fn, secs=2.5):
    fn(); torch.cuda.synchronize()
    stop, rows = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, rows)); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0; t0 = time.time(); e0.record()
    while time.time() - t0 < secs:
        for _ in range(20): fn()
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    ms = e0.elapsed_time(e1) / n
    print(f"{name}: {ms:.4f} ms/iter sustained; samples (MHz, W, pwr_cap): {rows[3:-1][::3]}")
    return ms

B, H, T = 16, 12, 1500
q, k, v = (torch.randn(B, T, H * 64, device="cuda").bfloat16().requires_grad_() for _ in range(3))
fl = 4.0 * B * H * T * T * 64
ms = run("fwd", lambda: A.qkv_attention(q, k, v, H))
print(f"  fwd sustained {fl/ms/1e9:.1f} TFLOP/s")
out, _, _ = A.qkv_attention(q, k, v, H)
do = torch.randn_like(out)
ms = run("bwd", lambda: torch.autograd.grad(out, (q, k, v), do, retain_graph=True))
print(f"  bwd sustained {2.5*fl/ms/1e9:.1f} TFLOP/s")
import torch.nn.functional as F
qh, kh, vh = (x.detach().view(B, -1, H, 64).transpose(1, 2) for x in (q, k, v))
ms = run("sdpa", lambda: F.scaled_dot_product_attention(qh, kh, vh))
print(f"  sdpa sustained {fl/ms/1e9:.1f} TFLOP/s")
a = torch.randn(8192, 8192, device="cuda").bfloat16(); b = torch.randn(8192, 8192, device="cuda").bfloat16()
ms = run("gemm8192", lambda: a @ b)
print(f"  gemm sustained {2*8192**3/ms/1e9:.1f} TFLOP/s")
