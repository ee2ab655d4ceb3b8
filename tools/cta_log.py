"""Per-CTA (clock64, globaltimer, smid) log of the tcgen05 attention kernels (debug build, -DAGA_TIMELINE): SM clock under
load, CTA durations, gaps between consecutive CTAs on an SM."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import aga_b200 as A
from aga_b200 import _lib

_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libaga_b200_timeline.so")
lib = _lib.lib()
lib.aga_debug_set_cta_log.argtypes = [C.c_void_p]

def analyse(buf, n, name):
    t = buf[: n * 8].view(n, 8).cpu()
    c0, g0, c1, g1, sm = t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]
    dur_c, dur_ns = (c1 - c0).double(), (g1 - g0).double()
    pro, loop, epi = (t[:, 5] - c0).double(), (t[:, 6] - t[:, 5]).double(), (c1 - t[:, 6]).double()
    print(f"   per CTA (cycles): prologue {pro.mean():.0f}  main loop {loop.mean():.0f}  epilogue {epi.mean():.0f}")
    span_ns = float(g1.max() - g0.min())
    print(f"{name}: {n} CTAs, kernel span {span_ns / 1e3:.1f} us; CTA duration {dur_c.mean():.0f} cycles = {dur_ns.mean() / 1e3:.2f} us "
          f"(min {dur_ns.min() / 1e3:.2f}, max {dur_ns.max() / 1e3:.2f}); SM clock {1e3 * dur_c.sum() / dur_ns.sum():.0f} MHz")
    gaps, busy = [], 0.0
    for s in sm.unique():
        idx = (sm == s).nonzero().flatten()
        order = idx[g0[idx].argsort()]
        st, en = g0[order], g1[order]
        gaps += (st[1:] - en[:-1]).tolist()
        busy += float((en - st).sum())
    gaps = torch.tensor(gaps if gaps else [0.0], dtype=torch.double)
    print(f"   CTAs per SM {n / len(sm.unique()):.2f}; gap between consecutive CTAs on an SM: mean {gaps.mean() / 1e3:.2f} us, "
          f"max {gaps.max() / 1e3:.2f} us; SM busy fraction {busy / (span_ns * len(sm.unique())):.3f}")

def main():
    B, H, T = 16, 12, 1500
    q, k, v = (torch.randn(B, T, H * 64, device="cuda").bfloat16().requires_grad_() for _ in range(3))
    out, _, _ = A.qkv_attention(q, k, v, H)
    do = torch.randn_like(out)
    out.backward(do, retain_graph=True)
    n_fwd = 12 * H * B  # one 128-row query tile per CTA
    n_bwd = min(12 * H * B, torch.cuda.get_device_properties(0).multi_processor_count)  # persistent: one CTA per SM
    buf = torch.zeros(max(n_fwd, n_bwd) * 8, dtype=torch.int64, device="cuda")
    for _ in range(3):  # a few back-to-back launches first: clocks settle
        A.qkv_attention(q, k, v, H)
    lib.aga_debug_set_cta_log(C.c_void_p(buf.data_ptr()))
    A.qkv_attention(q, k, v, H)
    torch.cuda.synchronize()
    lib.aga_debug_set_cta_log(C.c_void_p(0))
    analyse(buf, n_fwd, "fwd")
    for _ in range(3):
        out.backward(do, retain_graph=True)
    buf.zero_()
    torch.cuda.synchronize()
    lib.aga_debug_set_cta_log(C.c_void_p(buf.data_ptr()))
    out.backward(do, retain_graph=True)
    torch.cuda.synchronize()
    lib.aga_debug_set_cta_log(C.c_void_p(0))
    analyse(buf, n_bwd, "bwd")

if __name__ == "__main__":
    main()
