"""Which aten ops (with input shapes) the non-library part of an eager training step consists of: torch.profiler with
record_shapes, grouped by (op, shapes), sorted by device time.  Finds the elementwise leftovers worth fusing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
os.environ.setdefault("AGA_ALLOW_RANDOM_INIT", "1")
import torch
import bench
from torch.profiler import profile, ProfilerActivity
from aga_b200.parallel import FlatGradBucket
from aga_b200.optim import FlatAdamW
from aga_b200.graphed import EagerTrainStep

dev = torch.device("cuda")
torch.manual_seed(2022)
model = bench.build_model("small", dev, specaug=True)
model.static_shapes = True
params = [p for p in model.parameters() if p.requires_grad]
bucket = FlatGradBucket(params, shadow_dtype=torch.bfloat16, n_chunks=1, overlap=False)
opt = FlatAdamW(bucket, lr=1e-3, betas=(0.9, 0.99), eps=1e-6, weight_decay=0.01)
data = tuple(t.to(dev) for t in bench.synthetic_batch(16, 64, 2022))
step = EagerTrainStep(model, opt, bucket, max_grad_norm=1.0)
for _ in range(3):
    step(data)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step(data)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.self_device_time_total > 0 and e.key.startswith("aten::"):
        rows.append((e.self_device_time_total, e.count, e.key, str(e.input_shapes)[:110]))
rows.sort(reverse=True)
print("self device us | calls | op | input shapes")
for t, n, k, sh in rows[:60]:
    print(f"{t:9.0f} {n:5d}  {k:28s} {sh}")
