"""CUDA-graph capture of the adaptation step (streams and graphs instead of a tracing compiler).

The step launches ~1 300 kernels (cuBLAS GEMMs, the aga_b200 kernels …) and is launch-bound when driven from Python.
``GraphedTrainStep`` captures forward + backward + the chunked gradient all-reduces (launched from inside the backward
pass on a communication stream, ``parallel.FlatGradBucket``) + clip + AdamW ONCE on static input buffers and replays
it; inputs are copied into the static buffers before each replay.  One graph covers the whole step at any world size
(NCCL collectives are capturable).

Semantics of the reference trainer that live here (espnet2/train/trainer.py):
* ``accum_grad`` (:622-625, :649): ``accum_grad`` micro-batches per optimizer step, loss / accum_grad, gradients of all
  but the last micro-step accumulate locally without any collective (DDP ``no_sync``).  Two graphs are captured: the
  accumulating micro-step and the closing micro-step (backward + all-reduce + update).
* non-finite gradient norm => the update is skipped (:677).  Here that is decided ON THE DEVICE (the captured graph
  cannot branch on the host): ``isfinite(norm)`` is handed to the fused AdamW as ``found_inf`` (what GradScaler does),
  which leaves parameters, moments and step counters untouched; ``skipped_steps`` counts them.

``BucketedTrainStep`` is the same for the recipe's variable shapes (``batch_type: numel``: B, audio length and text
length change every step): one captured step per (B, frames-bucket, text-bucket), batches padded up into their bucket.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch


def _is_fused(opt: torch.optim.Optimizer) -> bool:
    return bool(opt.defaults.get("fused", False))


class GuardedUpdate:
    """clip_grad_norm_ + optimizer.step() with the reference's "skip the update when the norm is not finite"
    (trainer.py:666-677) decided on the device.  Shared by the graphed and the eager step."""

    def __init__(self, optimizer: torch.optim.Optimizer, bucket, max_grad_norm: float):
        self.opt, self.bucket, self.max_grad_norm = optimizer, bucket, max_grad_norm
        dev = bucket.flat.device
        self.found_inf = torch.zeros((), dtype=torch.float32, device=dev)
        self.skipped_steps = torch.zeros((), dtype=torch.float32, device=dev)
        self.last_norm = torch.zeros((), dtype=torch.float32, device=dev)
        self.fused = _is_fused(optimizer)
        from .optim import FlatAdamW
        self.flat = isinstance(optimizer, FlatAdamW)
        if self.flat:  # norm, skip rule, clip, update and bf16 copies are two launches of the library; it owns the counters
            self.skipped_steps, self.last_norm = optimizer.skipped, optimizer.grad_norm

    def __call__(self) -> None:
        if self.flat:
            self.opt.step(max_norm=self.max_grad_norm)
            return
        total = self.bucket.clip_grad_norm_(self.max_grad_norm)
        self.last_norm.copy_(total)
        bad = (~torch.isfinite(total)).float()
        self.found_inf.copy_(bad)
        self.skipped_steps.add_(bad)
        if self.fused:
            # fused (capturable) Adam/AdamW/SGD honour `found_inf` exactly like under GradScaler: no parameter, moment or
            # step-count change when it is 1 — the NaN gradients a non-finite norm leaves behind never reach the weights
            self.opt.found_inf = self.found_inf
            self.opt.step()
        elif not bool(bad):  # unfused optimizers (CPU tests) ignore found_inf: decide on the host
            self.opt.step()


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, bucket, example_batch: Sequence[torch.Tensor],
                 max_grad_norm: float = 1.0, amp_dtype: Optional[torch.dtype] = torch.bfloat16, warmup: int = 3,
                 accum_grad: int = 1, warmup_updates: bool = True):
        """``warmup_updates=False``: the warm-up passes before capture run forward + backward only — no optimizer step, so
        capturing a graph in the middle of training (BucketedTrainStep) does not change the training state."""
        self.model, self.opt, self.bucket = model, optimizer, bucket
        self.max_grad_norm, self.amp_dtype = max_grad_norm, amp_dtype
        self.accum_grad = max(1, int(accum_grad))
        self.update = GuardedUpdate(optimizer, bucket, max_grad_norm)
        self.static_in = tuple(t.clone() for t in example_batch)
        self._stage = self._copy_stream = self._staged = self._consumed = None
        self._micro = 0
        model.static_shapes = True
        # warm up on a side stream (allocator pools, cuBLAS handles, NCCL channels, packed-filter cache, lazy CUDA modules)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                for m in range(self.accum_grad):
                    self._fwd_bwd(m)
                if warmup_updates:
                    self.update()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # the closing micro-step: backward (+ all-reduce under it) + clip + optimizer, ONE graph
        self.g_last = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_last):
            self.loss, self.stats = self._fwd_bwd(self.accum_grad - 1)
            self.update()
        # accumulating micro-steps (accum_grad > 1): first (assigns) and middle (adds); no collective, no update
        self.g_first = self.g_mid = None
        if self.accum_grad > 1:
            self.g_first = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_first):
                self.loss_first, _ = self._fwd_bwd(0)
            if self.accum_grad > 2:
                self.g_mid = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.g_mid):
                    self.loss_mid, _ = self._fwd_bwd(1)

    @property
    def skipped_steps(self) -> torch.Tensor:
        return self.update.skipped_steps

    def _fwd_bwd(self, micro: int):
        last = micro == self.accum_grad - 1
        self.bucket.begin_step(accumulate=micro > 0, sync=last)
        with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
            loss, stats, _ = self.model(*self.static_in)
        (loss / self.accum_grad if self.accum_grad > 1 else loss).backward()
        self.bucket.finish_backward()
        return loss.detach(), {k: v for k, v in stats.items() if v is not None}

    # ---- pipelined input path: the NEXT batch crosses PCIe on a copy stream while the current step runs; at the start of
    #      its own step it is moved into the graph's static input buffers with a device-to-device copy (microseconds)
    def prefetch(self, host_batch: Sequence[torch.Tensor]) -> None:
        """Start the host->device copy of the batch the next ``run_prefetched()`` will train on (pinned host memory)."""
        if self._stage is None:
            self._copy_stream = torch.cuda.Stream()
            self._stage = tuple(torch.empty_like(t) for t in self.static_in)
            self._staged, self._consumed = torch.cuda.Event(), torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream())
        self._copy_stream.wait_event(self._consumed)  # the previous staged batch has been moved into the static buffers
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(self._stage, host_batch):
                dst.copy_(src, non_blocking=True)
            self._staged.record(self._copy_stream)

    def run_prefetched(self) -> torch.Tensor:
        main = torch.cuda.current_stream()
        main.wait_event(self._staged)
        for dst, src in zip(self.static_in, self._stage):
            dst.copy_(src, non_blocking=True)
        self._consumed.record(main)
        return self._replay()

    def _replay(self) -> torch.Tensor:
        """One MICRO-step; the optimizer steps on every ``accum_grad``-th call."""
        m, self._micro = self._micro, (self._micro + 1) % self.accum_grad
        if m == self.accum_grad - 1:
            self.g_last.replay()
            return self.loss
        if m == 0:
            self.g_first.replay()
            return self.loss_first
        self.g_mid.replay()
        return self.loss_mid

    def __call__(self, batch: Sequence[torch.Tensor]) -> torch.Tensor:
        for dst, src in zip(self.static_in, batch):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        return self._replay()


class EagerTrainStep:
    """The same step driven from Python (any shape, no capture): what ``BucketedTrainStep`` is checked against."""

    def __init__(self, model, optimizer, bucket, max_grad_norm: float = 1.0, amp_dtype: Optional[torch.dtype] = torch.bfloat16,
                 accum_grad: int = 1):
        self.model, self.bucket, self.amp_dtype = model, bucket, amp_dtype
        self.accum_grad = max(1, int(accum_grad))
        self.update = GuardedUpdate(optimizer, bucket, max_grad_norm)
        self._micro = 0

    @property
    def skipped_steps(self) -> torch.Tensor:
        return self.update.skipped_steps

    def __call__(self, batch: Sequence[torch.Tensor]) -> torch.Tensor:
        m, self._micro = self._micro, (self._micro + 1) % self.accum_grad
        last = m == self.accum_grad - 1
        self.bucket.begin_step(accumulate=m > 0, sync=last)
        dev_type = batch[0].device.type
        with torch.autocast(dev_type, dtype=self.amp_dtype, enabled=self.amp_dtype is not None and dev_type == "cuda"):
            loss, stats, _ = self.model(*batch)
        (loss / self.accum_grad if self.accum_grad > 1 else loss).backward()
        self.bucket.finish_backward()
        if last:
            self.update()
        self.stats = stats
        return loss.detach()


class BucketedTrainStep:
    """Graph-speed training on the recipe's VARIABLE shapes (``batch_type: numel``, no pad-trim:
    egs2/seame/asr1/conf/whisper/train_asr_whisper_small_adapter_csloss_2stage_check.yaml:43,58-59 — B, the audio
    length and the text length change with every mini-batch).

    A batch (speech (B,N), speech_lengths, text (B,L), text_lengths) is padded UP into a bucket and run through that
    bucket's captured step (captured on first use, kept in a small cache):
      * text: L -> the next multiple of ``text_multiple`` with ``ignore_id`` (-1).  Exact: padded target positions are
        ignored by the label-smoothing loss, carry +inf pattern rows in the guided loss (zeroed, not counted), and the
        causal decoder self-attention never lets a real row see a later (padded) key.
      * audio: N -> the bucket's sample count with zeros, and the TRUE common length ``N`` is passed to the model as
        ``valid_samples`` so that the padded frames are exactly what the reference never computes: the log-mel reflects
        at N and writes zeros past N//160, the encoder / cross attention mask keys past the true frame count
        (``kv_len``), and padded query rows feed nothing.  Equal to the unpadded step up to summation order.
      * B is part of the bucket key (a batch is never padded with fake utterances: the guided loss is a batch mean).
    """

    def __init__(self, model, optimizer, bucket, frame_buckets: Sequence[int] = (500, 1000, 1500, 2000, 2500, 3000),
                 text_multiple: int = 16, max_grad_norm: float = 1.0, amp_dtype: Optional[torch.dtype] = torch.bfloat16,
                 warmup: int = 2, max_graphs: int = 16, ignore_id: int = -1):
        self.model, self.opt, self.bucket = model, optimizer, bucket
        self.frame_buckets = tuple(sorted(int(f) for f in frame_buckets))
        self.text_multiple, self.ignore_id = int(text_multiple), ignore_id
        self.kw = dict(max_grad_norm=max_grad_norm, amp_dtype=amp_dtype, warmup=max(1, warmup), warmup_updates=False)
        self.max_graphs = max_graphs
        self.cache: Dict[Tuple[int, int, int], GraphedTrainStep] = {}
        self.captures = 0
        # the optimizer's lazily created state must exist before the first capture: one SKIPPED step (found_inf = 1, what
        # a non-finite gradient norm triggers) creates it without touching parameters, moments or step counts
        if _is_fused(optimizer) and not optimizer.state:
            for p in bucket.params:
                p.grad = bucket.views[bucket._index[id(p)]]
            bucket.flat.zero_()
            optimizer.found_inf = torch.ones((), dtype=torch.float32, device=bucket.flat.device)
            optimizer.step()
            optimizer.found_inf = None

    def bucket_of(self, B: int, N: int, L: int) -> Tuple[int, int, int]:
        frames = N // 160
        fb = next((f for f in self.frame_buckets if f >= frames), None)
        if fb is None:
            raise ValueError(f"{frames} mel frames exceed the largest bucket {self.frame_buckets[-1]}")
        tb = -(-L // self.text_multiple) * self.text_multiple
        return int(B), fb * 160, tb

    def pad(self, batch: Sequence[torch.Tensor]):
        speech, speech_lengths, text, text_lengths = batch
        B, N = speech.shape
        key = self.bucket_of(B, N, text.shape[1])
        _, Nb, Lb = key
        sp = speech if N == Nb else torch.nn.functional.pad(speech, (0, Nb - N))
        tx = text if text.shape[1] == Lb else torch.nn.functional.pad(text, (0, Lb - text.shape[1]), value=self.ignore_id)
        valid = torch.full((), N, dtype=torch.int32, device=speech.device)
        return key, (sp, speech_lengths, tx, text_lengths, valid)

    def __call__(self, batch: Sequence[torch.Tensor]) -> torch.Tensor:
        key, padded = self.pad(batch)
        step = self.cache.get(key)
        if step is None:
            if len(self.cache) >= self.max_graphs:
                self.cache.pop(next(iter(self.cache)))
            step = self.cache[key] = _ValidLenGraphedStep(self.model, self.opt, self.bucket, padded, **self.kw)
            self.captures += 1
        return step(padded)


class _ValidLenGraphedStep(GraphedTrainStep):
    """GraphedTrainStep whose fifth static input is the device scalar ``valid_samples`` (see BucketedTrainStep)."""

    def _fwd_bwd(self, micro: int):
        last = micro == self.accum_grad - 1
        self.bucket.begin_step(accumulate=micro > 0, sync=last)
        speech, speech_lengths, text, text_lengths, valid = self.static_in
        with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
            loss, stats, _ = self.model(speech, speech_lengths, text, text_lengths, valid_samples=valid)
        loss.backward()
        self.bucket.finish_backward()
        return loss.detach(), {k: v for k, v in stats.items() if v is not None}
