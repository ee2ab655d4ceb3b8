"""Per-shape time of the cuBLAS GEMMs (aten::mm / addmm / our linear_residual) in one training step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import bench
from torch.profiler import profile, ProfilerActivity

def main():
    dev = torch.device("cuda")
    model = bench.build_model("small", dev)
    from aga_b200.parallel import FlatGradBucket
    params = [p for p in model.parameters() if p.requires_grad]
    bucket = FlatGradBucket(params, shadow_dtype=torch.bfloat16)
    data = tuple(t.to(dev) for t in bench.synthetic_batch(16, 64, 2022))
    model.static_shapes = True
    def step():
        bucket.begin_step()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss, stats, w = model(*data)
        loss.backward()
        bucket.finish_backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages(group_by_input_shape=True):
        if e.key in ("aten::mm", "aten::addmm", "_LinearResidualFn") and e.device_time_total > 0:
            shp = [s for s in e.input_shapes if len(s) == 2]
            rows.append((e.device_time_total, e.count, e.key, str(e.input_shapes)[:110]))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"GEMM time per step: {tot / 1e3:.2f} ms")
    for t, n, k, s in rows[:28]:
        print(f"{t / 1e3:7.3f} ms  x{n:3d}  {t / n:7.1f} us  {k:18s} {s}")

if __name__ == "__main__":
    main()
