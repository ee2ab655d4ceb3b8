"""Per-role clock64() timeline of CTA 0 of the tensor-core log-mel kernel (debug build: make -C csrc timeline)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import aga_b200 as A
from aga_b200 import _lib

_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libaga_b200_timeline.so")
lib = _lib.lib()
lib.aga_debug_set_logmel_timeline.argtypes = [C.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
audio = (0.1 * torch.randn(B, 480000, device="cuda")).clamp_(-1, 1)
A.log_mel_spectrogram(audio, algo="tc")
buf = torch.zeros(8 * 4096, dtype=torch.int64, device="cuda")
lib.aga_debug_set_logmel_timeline(C.c_void_p(buf.data_ptr()))
A.log_mel_spectrogram(audio, algo="tc")
torch.cuda.synchronize()
lib.aga_debug_set_logmel_timeline(C.c_void_p(0))
t = buf.cpu().view(-1, 4096)
rows = []
for r in range(6):
    for i in range(2047):
        tag, clk = int(t[r][2 * i]), int(t[r][2 * i + 1])
        if tag < 0 or (tag == 0 and clk == 0):
            break
        rows.append((clk, r, tag))
rows.sort()
t0 = rows[0][0]
prev = {}
names = {0: "prep0", 1: "prep1", 2: "mma", 3: "load", 4: "epiL", 5: "epiU"}
for clk, r, tag in rows[:400]:
    d = clk - prev.get(r, clk)
    prev[r] = clk
    print(f"{clk - t0:8d}  {names[r]:6s} tag{tag:4d}  (+{d})")
