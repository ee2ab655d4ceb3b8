"""Log-mel frontend alone: both kernels, B in {1, 16, 64, 256} x 30 s, graph-timed (10 calls per replay), GB/s of algorithmic bytes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import aga_b200 as A

peak = 6556.2
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def graph_time(fn, reps=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    return sorted(ts)[len(ts) // 2]


for n_mels in (80, 128):
    for B in (1, 16, 64, 256):
        audio = (0.1 * torch.randn(B, 480000, device="cuda")).clamp_(-1, 1)
        nbytes = B * (480000 * 4 + n_mels * 3000 * 4)
        for algo in ("tc", "simt"):
            ms = graph_time(lambda: A.log_mel_spectrogram(audio, n_mels=n_mels, algo=algo))
            print(f"logmel n_mels={n_mels} B={B:3d} {algo:4s}: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.1f} GB/s  "
                  f"{nbytes / ms / 1e6 / peak:.3f} of measured HBM peak", flush=True)
