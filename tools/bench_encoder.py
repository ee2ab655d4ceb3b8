"""BASELINE configs[4]: log-mel frontend + encoder-only throughput sweep, batches of 30 s synthetic audio, forward pass
(inference: no SpecAug, no gradients), bf16 autocast as the recipe's `use_amp`, against the same modules driven through
the reference's eager op sequence (oracle/torch_port.py) on the same GPU and — for B = 1 — on the host CPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")]
os.environ.setdefault("AGA_ALLOW_RANDOM_INIT", "1")
import torch
import aga_b200  # noqa: F401
from aga_b200 import espnet_whisper as EW


def gpu_time(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    import torch_port
    enc = EW.OpenAIWhisperEncoder(1, whisper_model="small", adapter=True).cuda().eval()
    print("| batch x 30 s | this library, bf16 | audio-s/s | reference op sequence on the same B200, bf16 | speed-up |")
    print("|---:|---:|---:|---:|---:|")
    for B in (1, 4, 16, 64, 256):
        audio = (0.1 * torch.randn(B, 480000, device="cuda")).clamp_(-1, 1)
        lens = torch.full((B,), 480000, device="cuda")
        def ours():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return enc(audio, lens)
        t = gpu_time(ours, 5 if B >= 64 else 10)
        t_ref = None
        if B <= 64:
            def ref():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16), torch_port.patched_ops():
                    return enc(audio, lens)
            try:
                t_ref = gpu_time(ref, 3)
            except torch.cuda.OutOfMemoryError:
                t_ref = None
            torch.cuda.empty_cache()
        print(f"| {B} | {t:8.2f} ms | {B * 30 / t * 1e3:9.0f} | " + (f"{t_ref:8.2f} ms | {t_ref / t:5.2f}x |" if t_ref else "— (out of memory / skipped) | — |"))
    # host CPU, B = 1, fp32 (the reference's CPU-runnable case)
    enc_cpu = EW.OpenAIWhisperEncoder(1, whisper_model="small", adapter=True).eval()
    audio = (0.1 * torch.randn(1, 480000)).clamp_(-1, 1)
    lens = torch.full((1,), 480000)
    with torch.no_grad(), torch_port.patched_ops():
        enc_cpu(audio, lens)
        t0 = time.perf_counter()
        enc_cpu(audio, lens)
        dt = time.perf_counter() - t0
    print(f"\nhost CPU ({torch.get_num_threads()} threads), reference op sequence, fp32, B = 1: {dt * 1e3:.0f} ms = {30 / dt:.0f} audio-s/s")


if __name__ == "__main__":
    main()
