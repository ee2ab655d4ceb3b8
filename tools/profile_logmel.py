"""ncu target: a few calls of the log-mel frontend (B x 30 s), kernel chosen by argv[1] (tc | simt), B by argv[2]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import aga_b200 as A

algo = sys.argv[1] if len(sys.argv) > 1 else "tc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
audio = (0.1 * torch.randn(B, 480000, device="cuda")).clamp_(-1, 1)
for _ in range(4):
    y, _ = A.log_mel_spectrogram(audio, algo=algo)
torch.cuda.synchronize()
print("ok", float(y.mean()))
