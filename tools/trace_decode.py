"""Kernel timeline of graph-replayed decoding steps (torch.profiler / CUPTI): per-kernel start, duration and the idle
gap before it, to see whether a token's time is kernels or the gaps between them."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
os.environ.setdefault("AGA_ALLOW_RANDOM_INIT", "1")
import torch
from torch.profiler import profile, ProfilerActivity
import aga_b200  # noqa: F401
from aga_b200 import espnet_whisper as EW

dtype = torch.float32 if (len(sys.argv) > 1 and sys.argv[1] == "fp32") else torch.bfloat16
dec = EW.OpenAIWhisperDecoder(51865, 768, whisper_model="small", adapter=True, whisper_cs=True, src_layer=1).cuda().eval()
enc_out = torch.randn(1, 1500, 768, device="cuda", dtype=dtype)
ys = torch.tensor([[50258, 50260, 50259, 50359, 50363]], device="cuda")
gd = dec.greedy_decoder(enc_out, max_len=128)
gd.prefill(ys)
gd.decode(4)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    gd.graph.replay(); gd.graph.replay(); gd.graph.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
print("cuda events:", len(ev))
n = len(ev) // 3
step = ev[n:2 * n]
t0 = step[0].time_range.start
total = step[-1].time_range.end - t0
busy = sum(e.time_range.end - e.time_range.start for e in step)
print(f"middle replay: {len(step)} device activities, span {total:.1f} us, busy {busy:.1f} us, idle {total - busy:.1f} us")
g = collections.OrderedDict()
prev_end = t0
for e in step:
    c = g.setdefault(e.name[:70], [0, 0.0, 0.0])
    c[0] += 1
    c[1] += e.time_range.end - e.time_range.start
    c[2] += max(0.0, e.time_range.start - prev_end)
    prev_end = e.time_range.end
for k, v in sorted(g.items(), key=lambda kv: -(kv[1][1] + kv[1][2]))[:25]:
    print(f"{v[0]:4d}  busy {v[1]:8.1f} us  gap-before {v[2]:8.1f} us  {k}")
