"""Greedy decoding speed of the drop-in decoder: whole-prefix recompute (the reference's forward_one_step) vs KV-cached
incremental steps.  Whisper-small, random weights, 1 hypothesis, 60 steps, encoder output of 30 s."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import aga_b200  # noqa: F401
os.environ.setdefault("AGA_ALLOW_RANDOM_INIT", "1")
from aga_b200 import espnet_whisper as EW

def run(dec, enc_out, steps, cached):
    dec.kv_cache = cached
    ys = torch.tensor([[50258, 50260, 50259, 50359, 50363]], device="cuda")
    cache = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(steps):
            logp, cache = dec.forward_one_step(ys, torch.empty(0), enc_out, cache=cache)
            ys = torch.cat([ys, logp.argmax(-1, keepdim=True)], dim=1)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, ys

def run_graphed(dec, enc_out, steps):
    ys = torch.tensor([[50258, 50260, 50259, 50359, 50363]], device="cuda")
    gd = dec.greedy_decoder(enc_out, max_len=128)
    gd.prefill(ys)
    gd.decode(2)  # capture
    gd.prefill(ys)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ids, _ = gd.decode(steps)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, ids

def main():
    dec = EW.OpenAIWhisperDecoder(51865, 768, whisper_model="small", adapter=True, whisper_cs=True, src_layer=1).cuda().eval()
    for dtype in (torch.float32, torch.bfloat16):
        enc_out = torch.randn(1, 1500, 768, device="cuda", dtype=dtype)
        for cached in (False, True):
            run(dec, enc_out, 5, cached)
            dt, ys = run(dec, enc_out, 60, cached)
            print(f"{str(dtype):16s} {'kv-cached ' if cached else 'recompute '} 60 steps: {dt * 1e3:7.1f} ms  {60 / dt:7.1f} tokens/s")
        dt, ids = run_graphed(dec, enc_out, 60)
        print(f"{str(dtype):16s} cuda-graph  60 steps: {dt * 1e3:7.1f} ms  {60 / dt:7.1f} tokens/s")

if __name__ == "__main__":
    main()
