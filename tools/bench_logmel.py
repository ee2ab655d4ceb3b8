"""Log-mel frontend timing (BASELINE config 5 sweep): CUDA events, 20 calls captured in one CUDA graph, inputs > L2 at B >= 64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import torch
import aga_b200 as A
from bench_cross import graph_time

def main():
    for B in (1, 4, 16, 64, 256):
        audio = (0.1 * torch.randn(B, 480000, device="cuda")).clamp_(-1, 1)
        for n_mels in (80, 128):
            ms = graph_time(lambda: A.log_mel_spectrogram(audio, n_mels=n_mels), reps=10)
            nbytes = B * (480000 * 4 + n_mels * 3000 * 4)
            print(f"logmel B={B:3d} n_mels={n_mels:3d}: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.1f} GB/s  {B * 30 / ms * 1e3:10.0f} audio-s/s")

if __name__ == "__main__":
    main()
