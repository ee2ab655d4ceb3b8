"""Per-role clock64() timeline of CTA (0,0,0) of the tcgen05 attention kernels (debug build with -DAGA_TIMELINE)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import aga_b200 as A
from aga_b200 import _lib

_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libaga_b200_timeline.so")
lib = _lib.lib()
lib.aga_debug_set_timeline.argtypes = [C.c_void_p]

def dump(buf, roles, max_rows=70):
    t = buf.cpu().view(-1, 4096)
    t0 = None
    rows = []
    for r in roles:
        row = t[r]
        for i in range(2047):
            tag, clk = int(row[2 * i]), int(row[2 * i + 1])
            if tag < 0 or (tag == 0 and clk == 0): break
            rows.append((clk, r, tag))
    rows.sort()
    t0 = rows[0][0]
    # per-role period of the first tag of its loop body (20 for the softmax roles): tile periods over the whole CTA
    for r in roles:
        first = [c for c, rr, t in rows if rr == r and t in (20, 30)]
        if len(first) > 2:
            d = [b - a for a, b in zip(first, first[1:])]
            print(f"role{r}: {len(first)} loop iterations; periods: " + " ".join(str(x) for x in d[:60]))
    prev = {}
    for clk, r, tag in rows[:max_rows]:
        d = clk - prev.get(r, clk)
        prev[r] = clk
        print(f"{clk - t0:8d}  role{r}  tag{tag:3d}  (+{d})")
    return rows

def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    B, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2, 2)  # 16 12 = every SM busy
    T = 1500
    Tq = 64 if which == "cross" else T  # "cross": the Q-resident backward kernel (decoder cross attention)
    q = torch.randn(B, Tq, H * 64, device="cuda").bfloat16().requires_grad_()
    k, v = (torch.randn(B, T, H * 64, device="cuda").bfloat16().requires_grad_() for _ in range(2))
    buf = torch.zeros(8 * 4096, dtype=torch.int64, device="cuda")
    out, _, _ = A.qkv_attention(q, k, v, H)  # warm
    do = torch.randn_like(out)
    if which == "fwd":
        lib.aga_debug_set_timeline(C.c_void_p(buf.data_ptr()))
        out, _, _ = A.qkv_attention(q, k, v, H)
        torch.cuda.synchronize()
        lib.aga_debug_set_timeline(C.c_void_p(0))
        rows = dump(buf, [0, 1, 2, 3], 560)
    else:
        out.backward(do, retain_graph=True)
        lib.aga_debug_set_timeline(C.c_void_p(buf.data_ptr()))
        out.backward(do)
        torch.cuda.synchronize()
        lib.aga_debug_set_timeline(C.c_void_p(0))
        rows = dump(buf, [0, 1, 2, 3], 560)

if __name__ == "__main__":
    main()
