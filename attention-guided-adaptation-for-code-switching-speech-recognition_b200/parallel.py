"""Batch-sharded data parallelism for the adaptation step: one process per GPU, NCCL over NVLink/NVSwitch.

The reference wraps the model in DistributedDataParallel (espnet2/train/trainer.py:229-244); with
``--freeze_param`` only the adapter Linear/LayerNorm parameters carry gradients (14.3 M fp32 values for
Whisper-small), so the one collective of a step is a SUM all-reduce of those gradients.  Here the trainable
gradients live in ONE contiguous fp32 buffer (``p.grad`` are views into it), laid out in the order the backward
pass PRODUCES them (decoder block L-1 … 0, then encoder block L-1 … 0) and cut into a few chunks.  A
post-accumulate-grad hook per parameter counts a chunk's gradients in; the moment the last one arrives the chunk is
gathered with one multi-tensor copy and its all-reduce is launched on a communication stream — under the rest of the
backward pass, like DDP's buckets, but without DDP's per-bucket copies (the buffer IS the gradient storage).  NCCL
collectives are capturable, so the whole step (forward, backward, chunked all-reduces, clip, AdamW) is one CUDA graph.
``accum_grad`` (trainer.py:622-625, 649): micro-steps before the last one accumulate locally and launch nothing
(DDP's ``no_sync``).  The per-iteration scalar all-reduces of the reference (trainer.py:523, recursive_op.py:18,44)
are packed into one small tensor.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatGradBucket:
    """Contiguous gradient storage for the trainable parameters + its chunked, backward-overlapped all-reduce."""

    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None,
                 shadow_dtype: Optional[torch.dtype] = None, n_chunks: int = 4, overlap: bool = True):
        given = [p for p in params if p.requires_grad]
        if not given:
            raise ValueError("no trainable parameters")
        # module order is encoder 0..L-1, decoder 0..L-1 and autograd finishes in the reverse order: laying the buffer
        # out back to front makes "a prefix of the buffer is complete" true early in the backward pass
        self.params: List[torch.nn.Parameter] = list(reversed(given))
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views: List[torch.Tensor] = []
        self.offsets: List[int] = []
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise ValueError("trainable parameters are kept in fp32 (AMP master weights)")
            self.views.append(self.flat[off: off + p.numel()].view_as(p))
            self.offsets.append(off)
            p.grad = self.views[-1]
            off += p.numel()
        # chunks: contiguous runs of parameters of about numel / n_chunks elements each
        n_chunks = max(1, min(int(n_chunks), len(self.params)))
        target = self.numel / n_chunks
        self.chunk_of: List[int] = []
        self.chunk_bounds: List[List[int]] = []  # [first param, last param + 1, first element, last element + 1]
        c, start_p, start_e = 0, 0, 0
        for i, p in enumerate(self.params):
            self.chunk_of.append(c)
            end_e = self.offsets[i] + p.numel()
            if (end_e >= target * (c + 1) and c < n_chunks - 1) or i == len(self.params) - 1:
                self.chunk_bounds.append([start_p, i + 1, start_e, end_e])
                c, start_p, start_e = c + 1, i + 1, end_e
        self.n_chunks = len(self.chunk_bounds)
        self._pending = [0] * self.n_chunks
        # low-precision shadows of the trainable parameters (what autocast would re-cast on every use), refreshed by ONE
        # multi-tensor copy per step instead of one cast kernel per parameter per use (ops.cast_trainable reads them)
        self.shadow_flat = None
        self.shadow_views: List[torch.Tensor] = []
        if shadow_dtype is not None and dev.type == "cuda":
            self.shadow_flat = torch.empty(self.numel, dtype=shadow_dtype, device=dev)
            for p, off in zip(self.params, self.offsets):
                self.shadow_views.append(self.shadow_flat[off: off + p.numel()].view_as(p))
        if dev.type == "cuda":
            from . import ops
            ops.ZeroArena.of(dev).reserve(self.numel + (1 << 20))  # weight gradients + the small atomically-accumulated outputs
        self.group = process_group
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.overlap = bool(overlap)
        self._accumulate = False   # this backward ADDS to the buffer (micro-step > 0 of an accum_grad group)
        self._sync = True          # this backward ends an accum_grad group: all-reduce its chunks
        self._armed = False        # hooks are live between begin_step() and finish_backward()
        self._launched = [False] * self.n_chunks
        self.chunks_reduced_in_backward = 0  # evidence for tests / bench: chunks whose all-reduce started under backward
        self._index = {id(p): i for i, p in enumerate(self.params)}
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._on_grad)

    @property
    def nbytes(self) -> int:
        return self.numel * 4

    def zero_(self) -> None:
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    # ------------------------------------------------------------------ one backward pass
    def begin_step(self, accumulate: bool = False, sync: bool = True) -> None:
        """Call before the forward pass.  Drops ``p.grad`` so that autograd ASSIGNS this pass's gradients instead of
        launching one ``add_`` per parameter, arms the chunk hooks and (first micro-step only) refreshes the
        low-precision parameter shadows.  ``accumulate``: add this pass to what the buffer holds (micro-step > 0);
        ``sync``: this pass completes the accumulation group, so each finished chunk is all-reduced."""
        self._accumulate, self._sync = bool(accumulate), bool(sync)
        for p in self.params:
            p.grad = None
        if self.flat.is_cuda:
            from . import ops
            ops.ZeroArena.of(self.flat.device).reset()  # one clear for all the small atomically-accumulated outputs of this pass
        for c, (p0, p1, _, _) in enumerate(self.chunk_bounds):
            self._pending[c] = p1 - p0
            self._launched[c] = False
        self._armed = True
        if self.shadow_flat is not None and not accumulate:
            owner = getattr(self, "shadow_owner", None)  # optim.FlatAdamW writes the copies inside its update kernel
            if owner is None or not owner.shadow_is_fresh():
                with torch.no_grad():
                    torch._foreach_copy_(self.shadow_views, [p.detach() for p in self.params])
            for p, v in zip(self.params, self.shadow_views):
                p._aga_shadow = (p._version, v)

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if not self._armed:
            return
        c = self.chunk_of[self._index[id(p)]]
        self._pending[c] -= 1
        if self._pending[c] == 0 and self.overlap:
            self._flush_chunk(c, in_backward=True)

    def _flush_chunk(self, c: int, in_backward: bool = False) -> None:
        """Chunk c's gradients -> the flat buffer (one multi-tensor copy / add), then its all-reduce on the comm stream."""
        if self._launched[c]:
            return
        self._launched[c] = True
        p0, p1, e0, e1 = self.chunk_bounds[c]
        src, dst = [], []
        for i in range(p0, p1):
            p, v = self.params[i], self.views[i]
            g = p.grad
            if g is None:
                if not self._accumulate:
                    v.zero_()
            elif g.data_ptr() != v.data_ptr():
                src.append(g)
                dst.append(v)
            p.grad = v
        if src:
            with torch.no_grad():
                if self._accumulate:
                    torch._foreach_add_(dst, src)
                else:
                    torch._foreach_copy_(dst, src)
        if self._sync and self.world_size() > 1:
            self._all_reduce_range(e0, e1)
            if in_backward:
                self.chunks_reduced_in_backward += 1

    def _all_reduce_range(self, e0: int, e1: int) -> None:
        ws = self.world_size()
        chunk = self.flat[e0:e1]
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self.comm_stream):
                # AVG == SUM then / world (DDP's gradient averaging) inside the collective: no extra pass over the chunk
                dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group)
        else:  # gloo / CPU tests
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
            chunk.div_(ws)

    def finish_backward(self) -> None:
        """After ``loss.backward()``: gathers / reduces whatever the hooks have not (parameters without a gradient this
        pass, ``overlap=False``) and makes the current stream wait for the communication stream."""
        self._armed = False
        for c in range(self.n_chunks):
            self._flush_chunk(c)
        self.wait()
        if self.flat.is_cuda:
            from . import ops
            ops.ZeroArena.of(self.flat.device).disarm()

    # kept for callers of the round-1 interface: one gather, one all-reduce
    def gather_(self) -> None:
        sync, self._sync = self._sync, False
        self._armed = False
        for c in range(self.n_chunks):
            self._flush_chunk(c)
        self._sync = sync
        if self.flat.is_cuda:
            from . import ops
            ops.ZeroArena.of(self.flat.device).disarm()

    def all_reduce_mean_async(self) -> None:
        """SUM over ranks then / world of the whole buffer (== DDP's gradient averaging).  Returns immediately."""
        if self.world_size() > 1:
            self._all_reduce_range(0, self.numel)

    def wait(self) -> None:
        if self.comm_stream is not None and self.world_size() > 1:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """torch.nn.utils.clip_grad_norm_ on the flat buffer: one norm, one scale, no host sync.  Returns the norm
        (non-finite when any gradient is: the caller skips the update then, trainer.py:677)."""
        total = torch.linalg.vector_norm(self.flat)
        scale = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        self.flat.mul_(scale)
        return total


def all_reduce_stats(stats: Dict[str, Optional[torch.Tensor]], weight: torch.Tensor, group=None
                     ) -> Dict[str, Optional[torch.Tensor]]:
    """Weighted average of the scalar stats over ranks in ONE collective (reference: one per key)."""
    keys = [k for k, v in stats.items() if v is not None]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    packed = torch.stack([stats[k].detach().float().reshape(()) * weight for k in keys] + [weight.float().reshape(())])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    out = dict(stats)
    for i, k in enumerate(keys):
        out[k] = packed[i] / packed[-1]
    return out


def shard_batch(n_items: int, rank: int, world_size: int) -> slice:
    """The reference's ``batch[rank::world_size]`` partition (espnet2/tasks/abs_task.py:1632)."""
    return slice(rank, n_items, world_size)
