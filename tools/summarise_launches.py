"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step (the launches between
two consecutive logmel_frames_kernel launches), grouped by kernel, with each group's share of the step.
  python tools/summarise_launches.py gpurun_out/r1b_launches.csv profiles/r1_launches_step.csv > profiles/r1_launches.md
"""
import csv, re, sys, collections

def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|at::native::|void ", "", name)
    m = re.match(r"([A-Za-z0-9_:]+)", name)
    base = m.group(1) if m else name[:60]
    for key in ("GeluCUDAKernelImpl", "GeluBackwardCUDAKernelImpl", "bfloat16_copy_kernel_cuda", "float_copy", "direct_copy_kernel_cuda",
                "FillFunctor", "CUDAFunctor_add", "MulFunctor", "GammaBetaBackward", "layer_norm_grad_input",
                "vectorized_layer_norm", "reduce_kernel", "multi_tensor_apply"):
        if key in name:
            return base.split("::")[-1] + ":" + key
    return base

def main():
    src, step_out = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) >= 15 and r[0].isdigit()]
    marks = [i for i, r in enumerate(rows) if "logmel_frames_kernel" in r[4]]
    # last complete step in the capture
    lo, hi = marks[-2], marks[-1]
    step = rows[lo:hi]
    if step_out:
        w = csv.writer(open(step_out, "w"))
        w.writerow(["id", "kernel", "block", "grid", "gpu__time_duration.sum [ns]"])
        for r in step:
            w.writerow([r[0], r[4][:160], r[7], r[8], r[14]])
    tot = sum(float(r[14]) for r in step)
    g = collections.OrderedDict()
    for r in step:
        k = short(r[4])
        c = g.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += float(r[14])
    print(f"launches in the step: {len(step)}; sum of gpu__time_duration: {tot / 1e6:.2f} ms (ncu-serialised, cold cache)\n")
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for k, (n, t) in sorted(g.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| `{k[:90]}` | {n} | {t / 1e6:.3f} | {100 * t / tot:.1f} % |")
    ours = sum(t for k, (n, t) in g.items() if k.startswith("aga::"))
    print(f"\nkernels of this library (aga::*): {100 * ours / tot:.1f} % of the step's kernel time")

if __name__ == "__main__":
    main()
