"""Kernel timeline of ONE graph-replayed training step (torch.profiler / CUPTI): warm in-graph duration of every kernel
group and the idle gaps between kernels — the complement of the cold, serialised ncu launch list."""
import os, sys, collections, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
os.environ.setdefault("AGA_ALLOW_RANDOM_INIT", "1")
import torch
from torch.profiler import profile, ProfilerActivity
import bench
import aga_b200  # noqa: F401
from aga_b200.parallel import FlatGradBucket
from aga_b200.graphed import GraphedTrainStep

dev = torch.device("cuda", 0)
torch.manual_seed(2022)
model = bench.build_model("small", dev, specaug=True)
params = [p for p in model.parameters() if p.requires_grad]
bucket = FlatGradBucket(params, shadow_dtype=torch.bfloat16, n_chunks=1, overlap=False)
from aga_b200.optim import FlatAdamW
opt = FlatAdamW(bucket, lr=1e-3, betas=(0.9, 0.99), eps=1e-6, weight_decay=0.01)
resident = tuple(t.to(dev) for t in bench.synthetic_batch(16, 64, seed=2022))
model.static_shapes = True
step = GraphedTrainStep(model, opt, bucket, resident, max_grad_norm=1.0, warmup=3)
for _ in range(3):
    step(resident)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(resident); step(resident); step(resident)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
n = len(ev) // 3
one = ev[n:2 * n]
t0 = one[0].time_range.start
span = one[-1].time_range.end - t0
busy = sum(e.time_range.end - e.time_range.start for e in one)
print(f"middle replay: {len(one)} device activities, span {span / 1e3:.2f} ms, sum of durations {busy / 1e3:.2f} ms")
def short(name):
    name = re.sub(r"\(anonymous namespace\)::|at::native::|void ", "", name)
    m = re.match(r"([A-Za-z0-9_:]+)", name)
    base = m.group(1) if m else name[:50]
    for key in ("GeluCUDAKernelImpl", "bfloat16_copy", "FillFunctor", "CUDAFunctor_add", "MulFunctor", "multi_tensor_apply", "reduce_kernel",
                "FusedOptimizer", "CopyFunctor", "direct_copy"):
        if key in name:
            return base.split("::")[-1] + ":" + key
    return base
g = collections.OrderedDict()
prev_end = t0
gap_total = 0.0
gaps = []
prev_name = "(start)"
for e in one:
    gaps.append((max(0.0, e.time_range.start - prev_end), prev_name, short(e.name)))
    prev_name = short(e.name)
    c = g.setdefault(short(e.name), [0, 0.0, 0.0])
    c[0] += 1
    c[1] += e.time_range.end - e.time_range.start
    gap = max(0.0, e.time_range.start - prev_end)
    c[2] += gap
    gap_total += gap
    prev_end = max(prev_end, e.time_range.end)
print(f"idle gaps: {gap_total / 1e3:.2f} ms; the largest: " + "; ".join(f"{g_:.1f} us {a} -> {b}" for g_, a, b in sorted(gaps, reverse=True)[:8]))
print("| kernel | launches | in-graph ms | idle before (ms) |\n|---|---:|---:|---:|")
for k, v in sorted(g.items(), key=lambda kv: -(kv[1][1]))[:45]:
    print(f"| `{k}` | {v[0]} | {v[1] / 1e3:.3f} | {v[2] / 1e3:.3f} |")
