"""torch.profiler breakdown of the AGA training step (GPU box): where the non-attention time goes."""
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
    opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.99), eps=1e-6, weight_decay=0.01, fused=True)
    data = tuple(t.to(dev) for t in bench.synthetic_batch(16, 64, 2022))
    def step():
        bucket.begin_step()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss, stats, w = model(*data, static_text=True)
        loss.backward()
        bucket.finish_backward()
        bucket.clip_grad_norm_(1.0)
        opt.step()
    for _ in range(3): step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2): step()
        torch.cuda.synchronize()
    ka = prof.key_averages()
    tot_cuda = sum(e.self_device_time_total for e in ka) / 2e3
    print(f"total CUDA kernel time per step: {tot_cuda:.2f} ms; kernels per step: {sum(e.count for e in ka if e.self_device_time_total>0)/2:.0f}")
    print(ka.table(sort_by="self_cuda_time_total", row_limit=40, max_name_column_width=60))

if __name__ == "__main__":
    main()
