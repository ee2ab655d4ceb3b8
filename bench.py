#!/usr/bin/env python
"""bench.py — train audio-sec/s of the Whisper attention-guided-adaptation step on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's eager-PyTorch path on the host CPU cores

A "step" is one pass of the hot path over one batch of synthetic input: log-mel -> Whisper-small encoder (adapters)
-> decoder (adapters, self-attention columns 1:3 exported) -> label-smoothing CE + attention-guided loss -> backward
-> all-reduce of the adapter gradients (N > 1) -> grad clip -> AdamW on the adapters, under bf16 autocast.
Workload: BASELINE.json configs[1] ("Whisper-small AGA training step, synthetic 30 s 16 kHz audio, batch 16, bf16").

Prints ONE JSON line (rank 0).  `value` = whole-job audio-seconds per second with inputs resident in HBM;
`e2e` = the same through the public module API with pinned-host inputs copied every step and the loss read back.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

AUDIO_SECONDS = 30
N_SAMPLES = 480000
# BASELINE.json configs[1..3]
CONFIGS = {
    "small16": {"model": "small", "batch": 16, "baseline_config": "configs[1]"},
    "medium32": {"model": "medium", "batch": 32, "baseline_config": "configs[2]"},
    "large8": {"model": "large-v2-mel128", "batch": 8, "baseline_config": "configs[3]"},
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="aga_b200", choices=["aga_b200", "reference"])
    ap.add_argument("--config", default=os.environ.get("AGA_BENCH_CONFIG", "small16"), choices=sorted(CONFIGS),
                    help="BASELINE.json workload: small16 = configs[1] (default), medium32 = configs[2], large8 = configs[3] "
                         "(large-v2 body, 128-bin mel stem)")
    ap.add_argument("--model", default=None, help="override the config's model")
    ap.add_argument("--batch", type=int, default=None, help="override utterances per GPU per step (weak scaling)")
    ap.add_argument("--text-len", type=int, default=64, help="decoder input length T (ys_in)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="drive the step from Python instead of replaying CUDA graphs")
    ap.add_argument("--no-specaug", dest="specaug", action="store_false",
                    help="leave out the recipe's SpecAug (graph-safe device variant; on by default, both arms)")
    ap.add_argument("--cpu-batch", type=int, default=2, help="utterances per CPU-baseline step (bounded sample)")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the eager-PyTorch (reference op sequence) leg on the GPU")
    ap.add_argument("--torch-adamw", action="store_true", help="use torch.optim.AdamW(fused) + clip_grad_norm_ instead of optim.FlatAdamW")
    ap.add_argument("--accum-grad", type=int, default=1, help="micro-batches per optimizer step (recipe: 4); a 'step' stays one micro-batch")
    ap.add_argument("--comm-chunks", type=int, default=1, help="gradient all-reduce chunks")
    ap.add_argument("--comm-overlap", action="store_true",
                    help="launch each chunk's all-reduce from inside the backward pass (measured SLOWER on this step: the "
                         "persistent attention kernels own all 148 SMs and wait for the SMs NCCL occupies; default: one "
                         "all-reduce after the backward pass, inside the same CUDA graph)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    args.model = args.model or cfg["model"]
    args.batch = args.batch or cfg["batch"]
    return args


# ---------------------------------------------------------------------------------------------- data / model
def synthetic_batch(batch, text_len, seed):
    """0.1*randn audio clipped to [-1,1]; tokens = [zh,en,transcribe,notimestamps] + random ids + EOT."""
    g = torch.Generator().manual_seed(seed)
    speech = (0.1 * torch.randn(batch, N_SAMPLES, generator=g)).clamp_(-1, 1)
    speech_lengths = torch.full((batch,), N_SAMPLES, dtype=torch.long)
    n_words = text_len - 6
    body = torch.randint(0, 50257, (batch, n_words), generator=g)
    prompt = torch.tensor([50260, 50259, 50359, 50363]).expand(batch, 4)
    text = torch.cat([prompt, body, torch.full((batch, 1), 50257)], dim=1)
    text_lengths = torch.full((batch,), text.shape[1], dtype=torch.long)
    return speech, speech_lengths, text, text_lengths


# specaug_conf of espnet/egs2/seame/asr1/conf/whisper/train_asr_whisper_small_adapter_csloss_2stage_check.yaml:7-24
RECIPE_SPECAUG = dict(apply_time_warp=True, time_warp_window=5, time_warp_mode="bicubic", apply_freq_mask=True,
                      freq_mask_width_range=(0, 30), num_freq_mask=2, apply_time_mask=True, time_mask_width_range=(0, 40),
                      num_time_mask=2)


def build_model(name, device, export_mode="fused", specaug=False):
    import aga_b200  # noqa: F401
    from aga_b200 import espnet_model as EM, espnet_whisper as EW, whisper_model as W

    W.allow_random_init()  # synthetic benchmark: random-init weights of the named architecture (no checkpoints offline)
    d = W.MODEL_DIMS[name]
    enc = EW.OpenAIWhisperEncoder(whisper_model=name, adapter=True, use_specaug=specaug,
                                  specaug_conf=dict(RECIPE_SPECAUG, graph_safe=True) if specaug else None)
    dec = EW.OpenAIWhisperDecoder(d.n_vocab, d.n_text_state, whisper_model=name, adapter=True, whisper_cs=True,
                                  src_layer=1, export_mode=export_mode, fused_loss=str(device) != "cpu")
    toks = [str(i) for i in range(d.n_vocab)]
    toks[50258], toks[50257] = "<|startoftranscript|>", "<|endoftext|>"
    if (d.n_text_layer, d.n_text_head) == (12, 12):
        kw = {}
    else:  # documented extrapolation: seeded ~50 % mask, first three layers off
        g = torch.Generator().manual_seed(2022)
        mask = (torch.rand(d.n_text_layer, d.n_text_head, generator=g) < 0.5).float()
        mask[:3] = 0
        kw = {"use_literal_head_mask": False}
    model = EM.ESPnetASRModel(d.n_vocab, toks, encoder=enc, decoder=dec, cs_weight=0.01, lsm_weight=0.1,
                              c_val_attention=0.6, sym_sos="<|startoftranscript|>", sym_eos="<|endoftext|>", **kw)
    if kw:
        model.cs_head_mask = mask
    for n, p in model.named_parameters():  # --freeze_param policy "adapter" (espnet2/tasks/abs_task.py:1170-1177)
        p.requires_grad_("adapter" in n)
    return model.to(device).train()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------- reference arm
def cpu_reference_rate(model_name, batch, text_len, steps, warmup, threads, specaug=True):
    """The reference's eager-PyTorch path (oracle/torch_port.py inside the mirror modules) on the host cores, fp32."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_port

    torch.set_num_threads(threads)
    with torch_port.patched_ops():
        model = build_model(model_name, "cpu", export_mode="full", specaug=specaug)
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.99), eps=1e-6, weight_decay=0.01)
        data = synthetic_batch(batch, text_len, seed=2022)
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            loss, stats, _ = model(*data)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * AUDIO_SECONDS * len(times) / total, 1e3 * total / len(times)


def gpu_eager_rate(args, dev, batch, steps=3, warmup=2):
    """G-eager: the reference's eager-PyTorch op sequence on the SAME B200 under bf16 autocast, as
    espnet2/train/trainer.py:41-50,567-576 would run it — materialised (B,H,T,T) scores, fp32 softmax, separate q/k/v
    Linears, nn.Conv1d stem, torch.stft frontend, full (L,B,H,T,T) map export, unfused loss, stock AdamW.  The real
    "before" number on this box; oracle/torch_port.py supplies the op sequence (a port: /root/reference is not on the box)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_port

    with torch_port.patched_ops():
        model = build_model(args.model, dev, export_mode="full", specaug=args.specaug)
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.99), eps=1e-6, weight_decay=0.01)
        data = tuple(t.to(dev) for t in synthetic_batch(batch, args.text_len, seed=2022))
        torch.cuda.reset_peak_memory_stats(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(warmup + steps):
            if i == warmup:
                torch.cuda.synchronize()
                e0.record()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss, stats, _ = model(*data)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        peak_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    del model, opt, params, data, loss, stats
    torch.cuda.empty_cache()
    return batch * AUDIO_SECONDS / (ms / 1e3), ms, peak_gb


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rate, ms = cpu_reference_rate(args.model, args.cpu_batch, args.text_len, args.steps, args.warmup, cores, args.specaug)
    line = {
        "impl": "reference", "metric": "train audio-sec/s, Whisper AGA step", "value": rate, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # what THIS arm timed: a bounded sample (cpu_batch utterances, full maps exported as the reference does), one CPU
        # process whatever --gpus says — not the GPU arm's batch
        "config": workload_config(args, batch=args.cpu_batch, world=1, optimizer="AdamW(adapters): torch.optim.AdamW", export="decoder self-attn full (L,B,H,T,T) maps "
                                  "(the reference's export)", note=f"bounded sample of the GPU arm's workload "
                                  f"(batch_per_gpu {args.batch}); CPU port of the reference (oracle/torch_port.py inside the "
                                  "repo's mirror modules), 1 process"),
        "same_config_as_gpu_arm": False,
        "cpu_baseline": {"value": rate, "unit": "audio-s/s", "cores": cores, "kind": "port",
                         "sample": f"{args.cpu_batch} x 30 s utterance(s) per step, eager-PyTorch port of the reference "
                                   f"step (oracle/torch_port.py) on {cores} host threads, fp32"},
        "e2e": {"value": rate, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, batch=None, world=None, optimizer=None, export="none: guided-loss reduction fused into the decoder self-attention "
                    "epilogue (cols 1:3 of the scaled logits)", note=None):
    batch = args.batch if batch is None else batch
    world = args.gpus if world is None else world
    cfg = {"workload": f"Whisper-{args.model} attention-guided adaptation training step "
                       f"({CONFIGS[args.config]['baseline_config']})",
           "batch_per_gpu": batch, "global_batch": batch * world, "audio_seconds": AUDIO_SECONDS,
           "text_len": args.text_len, "adapters": True, "export": export,
           "optimizer": optimizer or ("AdamW(adapters): " + ("torch fused" if args.torch_adamw else "flat buffers, clip + update + bf16 copies in one pass")), "parallelism": f"dp{world}", "specaug": bool(args.specaug),
           "accum_grad": args.accum_grad,
           "l2_policy": "inputs+activations per step (>3 GB) exceed the 126 MB L2; no explicit flush"}
    if note:
        cfg["note"] = note
    return cfg


# ---------------------------------------------------------------------------------------------- main arm
_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract goes to the process's original stdout."""
    print(json.dumps(line), file=_REAL_STDOUT or sys.stdout, flush=True)


def main():
    global _REAL_STDOUT
    args = parse_args()
    # stdout carries exactly one JSON line: whatever else writes to fd 1 (NCCL's version banner, library chatter) is
    # sent to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: aga_b200 has no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    import aga_b200 as A
    from aga_b200 import ops
    from aga_b200.parallel import FlatGradBucket, all_reduce_stats

    torch.manual_seed(2022)
    model = build_model(args.model, dev, specaug=args.specaug)
    params = [p for p in model.parameters() if p.requires_grad]
    bucket = FlatGradBucket(params, shadow_dtype=torch.bfloat16, n_chunks=args.comm_chunks, overlap=args.comm_overlap)
    if args.torch_adamw:  # stock optimizer chain (multi-tensor norm / clip / fused AdamW / bf16 re-casts), for comparison
        opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.99), eps=1e-6, weight_decay=0.01, fused=True,
                                capturable=not args.no_graph)
    else:  # the same update on flat buffers: two launches of the library (csrc/flat_adamw.cu)
        from aga_b200.optim import FlatAdamW
        opt = FlatAdamW(bucket, lr=1e-3, betas=(0.9, 0.99), eps=1e-6, weight_decay=0.01)

    host = synthetic_batch(args.batch, args.text_len, seed=2022 + rank)
    host = tuple(t.pin_memory() for t in host)
    resident = tuple(t.to(dev) for t in host)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)

    model.static_shapes = True  # synthetic batches are already cut to the longest target: no host syncs in the step

    from aga_b200.graphed import EagerTrainStep, GraphedTrainStep
    eager_step = EagerTrainStep(model, opt, bucket, max_grad_norm=1.0, accum_grad=args.accum_grad)
    if args.no_graph:
        step = eager_step
    else:
        step = GraphedTrainStep(model, opt, bucket, resident, max_grad_norm=1.0, warmup=3, accum_grad=args.accum_grad)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(3, args.warmup)):
        step(resident)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    total_ms = timed(args.steps, lambda: step(resident))

    # per-kernel timing (roofline) and launch counting need Python-driven launches: an eager pass of the same K steps
    eager_step(resident)
    n0 = A.launch_count()
    ops.PROFILE = {}
    eager_ms = timed(args.steps, lambda: eager_step(resident))
    prof, ops.PROFILE = ops.PROFILE, None
    launches = A.launch_count() - n0

    def graph_time_ms(fn, reps=10):
        """Device time per call with the CPU launch path taken out: `reps` calls captured in one CUDA graph (the eager
        CUDA-event brackets above over-state kernels shorter than the Python launch path, ~60 us)."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / reps)
        return sorted(ts)[len(ts) // 2]

    # the frontend alone (metric 3, "mel GB/s"): its two kernels are ~50 us, far below the eager launch path
    n_mels = model.encoder.n_mels
    with torch.no_grad():
        mel_ms = graph_time_ms(lambda: ops.log_mel_spectrogram(resident[0], n_mels=n_mels))
        mel_tc_ms = graph_time_ms(lambda: ops.log_mel_spectrogram(resident[0], n_mels=n_mels, algo="tc"))
    mel_bytes = args.batch * (N_SAMPLES * 4 + n_mels * (N_SAMPLES // 160) * 4)

    if args.no_graph:
        def e2e_step():
            batch = tuple(t.to(dev, non_blocking=True) for t in host)
            loss = step(batch)
            return loss.item()  # device -> host read of the step's result
    else:
        # pipelined input path of the public API (GraphedTrainStep.prefetch / run_prefetched): every step's batch is
        # copied from pinned host memory inside the timed region — on a copy stream, under the previous step's kernels —
        # and every step's loss is read back to the host
        step.prefetch(host)

        def e2e_step():
            loss = step.run_prefetched()
            step.prefetch(host)       # the next step's batch starts crossing PCIe now
            return loss.item()        # device -> host read of the step's result

    e2e_step()
    e2e_ms = timed(args.steps, e2e_step)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        leave(world)
        return

    ms_per_step = total_ms / args.steps
    audio_s = args.batch * world * AUDIO_SECONDS
    value = audio_s / (ms_per_step / 1e3)
    e2e_value = audio_s / (e2e_ms / args.steps / 1e3)

    # per-kernel-family device time inside the timed region (CUDA events on the launching stream)
    kern = {}
    for tag, recs in prof.items():
        ms = [e0.elapsed_time(e1) for (_, e0, e1) in recs]
        work = sum(w for (w, _, _) in recs)
        kern[tag] = {"launches": len(recs), "ms_total": sum(ms), "ms_avg": sum(ms) / len(ms),
                     "rate": work / (sum(ms) / 1e3)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    attn = {k: v for k, v in kern.items() if k.startswith("attn_")}
    dom = max(attn, key=lambda k: attn[k]["ms_total"]) if attn else None
    roofline = None
    if dom:
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        ach = kern[dom]["rate"] / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(dom)
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach / peak, "frac_of_burst_peak": ach / peaks.get("bf16_tflops", 1590.0), "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                    if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)",
                    "share_of_step": kern[dom]["ms_total"] / total_ms,
                    "timed_in": "eager pass of the same K steps (CUDA events around each launch)"}
    summary = {k: {"launches": v["launches"], "ms_per_step": v["ms_total"] / args.steps,
                   ("GB/s" if k == "logmel" else "TFLOP/s"): v["rate"] / (1e9 if k == "logmel" else 1e12)}
               for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms_total"])}
    hbm_peak = peaks.get("hbm_gbs", 6556.2)
    summary["logmel"] = {"launches": kern["logmel"]["launches"] if "logmel" in kern else 0, "us_per_call": mel_ms * 1e3,
                         "GB/s": mel_bytes / mel_ms / 1e6, "frac_of_hbm_peak": mel_bytes / mel_ms / 1e6 / hbm_peak,
                         "bound": "fp32 issue (400-point DFT on CUDA cores), not HBM: see DESIGN.md",
                         "timed_in": "10 calls captured in one CUDA graph"}
    summary["logmel_tc"] = {"us_per_call": mel_tc_ms * 1e3, "GB/s": mel_bytes / mel_tc_ms / 1e6,
                            "frac_of_hbm_peak": mel_bytes / mel_tc_ms / 1e6 / hbm_peak,
                            "what": "the tensor-core frontend (csrc/logmel_tc.cu: fp16 split-precision DFT GEMMs on tcgen05), not "
                                    "the default: its epilogue (mel projection out of single-buffered TMEM accumulators) bounds it",
                            "timed_in": "10 calls captured in one CUDA graph"}

    gpu_eager = None
    if world == 1 and not args.no_gpu_eager:
        # free the product model's activations/graph pools first? they stay: 180 GB holds both at these sizes
        eb = args.batch if args.model in ("tiny", "base", "small") else min(args.batch, 4)
        try:
            rate, ms, peak_gb = gpu_eager_rate(args, dev, eb)
            gpu_eager = {"value": rate, "unit": "audio-s/s", "ms_per_step": ms, "batch": eb, "dtype": "bf16 autocast",
                         "kind": "port", "peak_mem_gib": peak_gb,
                         "what": "the reference's eager-PyTorch op sequence (oracle/torch_port.py in the mirror modules) on this "
                                 "B200: materialised scores, full-map export, unfused loss, stock AdamW; 3 timed steps after 2"}
        except torch.cuda.OutOfMemoryError as e:  # reported, not hidden
            gpu_eager = {"unavailable": f"out of memory at batch {eb}: {str(e)[:120]}"}
            torch.cuda.empty_cache()

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        t0 = time.time()
        rate, ms = cpu_reference_rate(args.model, args.cpu_batch, args.text_len, steps=2, warmup=1, threads=cores,
                                      specaug=args.specaug)
        cpu_baseline = {"value": rate, "unit": "audio-s/s", "cores": cores, "kind": "port",
                        "sample": f"2 timed steps (+1 warm-up) of {args.cpu_batch} x 30 s utterance(s): eager-PyTorch port "
                                  f"of the reference step (oracle/torch_port.py), fp32, {cores} host threads, "
                                  f"{time.time() - t0:.0f} s wall"}

    line = {
        "metric": "train audio-sec/s, Whisper AGA step", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": summary,
        "cpu_baseline": cpu_baseline, "gpu_eager_baseline": gpu_eager,
        "comm": {"chunks": bucket.n_chunks, "overlap_with_backward": bool(args.comm_overlap),
                 "chunks_all_reduced_inside_backward": bucket.chunks_reduced_in_backward > 0,
                 "one_graph": not args.no_graph},
        "skipped_steps": float(step.skipped_steps), "allreduce_bytes_per_step": bucket.nbytes if world > 1 else 0,
        "step_driver": "python-eager" if args.no_graph else "cuda-graph replay", "eager_ms_per_step": eager_ms / args.steps,
    }
    emit(line)
    leave(world)


def leave(world):
    """End of a rank.  The captured step holds NCCL kernels: tearing the communicator down under live CUDA graphs (or
    letting interpreter shutdown order their destructors) hung the 2-GPU run after the result line had been printed, so
    multi-rank processes synchronise, flush and leave without running destructors."""
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        if _REAL_STDOUT is not None:
            _REAL_STDOUT.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
