"""CUDA-graph capture of the adaptation step (streams and graphs instead of a tracing compiler).

The step launches ~3 400 kernels (cuBLAS GEMMs, LayerNorm, GELU, the aga_b200 kernels …) and is launch-bound when
driven from Python.  ``GraphedTrainStep`` captures forward + backward (+ clip + AdamW when single-GPU) once on static
input buffers and replays it; inputs are copied into the static buffers before each replay.  With more than one rank
the gradient all-reduce runs between two captured halves (forward/backward | clip/optimizer)."""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, bucket, example_batch: Sequence[torch.Tensor],
                 max_grad_norm: float = 1.0, amp_dtype: Optional[torch.dtype] = torch.bfloat16, warmup: int = 3):
        self.model, self.opt, self.bucket = model, optimizer, bucket
        self.max_grad_norm, self.amp_dtype = max_grad_norm, amp_dtype
        self.static_in = tuple(t.clone() for t in example_batch)
        self._stage = self._copy_stream = self._staged = self._consumed = None
        self.multi = bucket.world_size() > 1
        model.static_shapes = True
        # warm up on a side stream (allocator pools, cuBLAS handles, packed-filter cache, lazy CUDA modules)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                self._reduce()
                self._update()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fb):
            self.loss, self.stats = self._fwd_bwd()
            if not self.multi:
                self._update()
        self.g_up = None
        if self.multi:
            self.g_up = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_up):
                self._update()

    def _fwd_bwd(self):
        self.bucket.begin_step()
        with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
            loss, stats, _ = self.model(*self.static_in)
        loss.backward()
        self.bucket.gather_()
        return loss.detach(), {k: v for k, v in stats.items() if v is not None}

    def _reduce(self):
        if self.multi:
            self.bucket.all_reduce_mean_async()
            self.bucket.wait()

    def _update(self):
        self.bucket.clip_grad_norm_(self.max_grad_norm)
        self.opt.step()

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
        self.g_fb.replay()
        if self.multi:
            self._reduce()
            self.g_up.replay()
        return self.loss

    def __call__(self, batch: Sequence[torch.Tensor]) -> torch.Tensor:
        for dst, src in zip(self.static_in, batch):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        return self._replay()
