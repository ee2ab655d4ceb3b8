"""Where the warps of the persistent attention backward kernel wait: per-role shares of CTA 0's life spent inside each
barrier wait, from the clock64() timeline of the debug build (make -C csrc timeline; tags in csrc/attn_tc.cu).
  LD_PRELOAD=$PWD/<pkg>/libaga_b200_timeline.so python tools/bwd_waits.py [B H]        (default 16 12: every SM busy)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import aga_b200 as A
from aga_b200 import _lib

lib = C.CDLL(os.path.join(os.path.dirname(_lib.LIB_PATH), "libaga_b200_timeline.so"))
lib.aga_debug_set_timeline.argtypes = [C.c_void_p]
B, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 12)
T = 1500
q, k, v = (torch.randn(B, T, H * 64, device="cuda").bfloat16().requires_grad_() for _ in range(3))
out, _, _ = A.qkv_attention(q, k, v, H)
do = torch.randn_like(out)
out.backward(do, retain_graph=True)
buf = torch.zeros(8 * 4096, dtype=torch.int64, device="cuda")
lib.aga_debug_set_timeline(C.c_void_p(buf.data_ptr()))
out.backward(do)
torch.cuda.synchronize()
lib.aga_debug_set_timeline(C.c_void_p(0))
t = buf.cpu().view(-1, 4096)


def events(role):
    ev = []
    for i in range(2047):
        tag, clk = int(t[role][2 * i]), int(t[role][2 * i + 1])
        if tag < 0 or (tag == 0 and clk == 0):
            break
        ev.append((tag, clk))
    return ev


def spans(ev, pairs):
    """sum of clk(b) - clk(a) over consecutive events a -> b for each (a, b) in pairs; total = last - first"""
    acc = {p: 0 for p in pairs}
    for (ta, ca), (tb, cb) in zip(ev, ev[1:]):
        if (ta, tb) in acc:
            acc[(ta, tb)] += cb - ca
    return acc, ev[-1][1] - ev[0][1]


names = {
    0: ("MMA stream of half 0 (S^T, dV, dP^T, dK)", {(10, 11): "waits for P (softmax phase 1) + next Q/dO", (12, 13): "waits for dS (softmax phase 2)",
                                                    (11, 12): "issues S^T(next) / dV", (13, 14): "issues dP^T(next) / dK"}),
    1: ("softmax warps, half 0", {(20, 26): "waits for the tile's statistics", (26, 21): "waits for S^T (MMA)", (21, 22): "waits for its turn at phase 1 (halves alternate)",
                                 (22, 23): "phase 1: exp2, P -> TMEM", (23, 24): "waits for dP^T / a free dS panel", (24, 25): "phase 2: dS -> TMEM + smem"}),
    2: ("softmax warps, half 1", {(20, 26): "waits for the tile's statistics", (26, 21): "waits for S^T (MMA)", (21, 22): "waits for its turn at phase 1 (halves alternate)",
                                 (22, 23): "phase 1: exp2, P -> TMEM", (23, 24): "waits for dP^T / a free dS panel", (24, 25): "phase 2: dS -> TMEM + smem"}),
    3: ("dQ drain warp", {(30, 31): "waits for dQ (MMA)", (31, 32): "TMEM -> staging -> bulk reduce-add"}),
}
print(f"persistent attention backward, B {B} H {H} T {T}: CTA 0, clock64 cycles\n")
for role, (title, pairs) in names.items():
    ev = events(role)
    if len(ev) < 4:
        continue
    acc, total = spans(ev, list(pairs))
    n_iter = sum(1 for tg, _ in ev if tg == min(a for a, _ in pairs))
    print(f"{title}: {n_iter} tiles, {total} cycles ({total / max(n_iter, 1):.0f} per tile)")
    for p, label in pairs.items():
        print(f"    {100.0 * acc[p] / total:5.1f} %  {label}")
