"""Batch-sharded data parallelism for the adaptation step: one process per GPU, NCCL over NVLink/NVSwitch.

The reference wraps the model in DistributedDataParallel (espnet2/train/trainer.py:229-244); with
``--freeze_param`` only the adapter Linear/LayerNorm parameters carry gradients (14.3 M fp32 values for
Whisper-small), so the one collective of a step is a SUM all-reduce of those gradients.  Here the trainable
gradients live in ONE contiguous fp32 buffer (``p.grad`` are views into it) that is reduced with a single
``all_reduce`` per optimizer step — launched on a side stream so that it overlaps whatever the main stream still
has queued — and the per-iteration scalar all-reduces of the reference (trainer.py:523, recursive_op.py:18,44)
are packed into one small tensor.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatGradBucket:
    """Contiguous gradient storage for the trainable parameters + its all-reduce."""

    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None,
                 shadow_dtype: Optional[torch.dtype] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise ValueError("trainable parameters are kept in fp32 (AMP master weights)")
            self.views.append(self.flat[off: off + p.numel()].view_as(p))
            p.grad = self.views[-1]
            off += p.numel()
        # low-precision shadows of the trainable parameters (what autocast would re-cast on every use), refreshed by ONE
        # multi-tensor copy per step instead of one cast kernel per parameter per use (ops.cast_trainable reads them)
        self.shadow_flat = None
        self.shadow_views: List[torch.Tensor] = []
        if shadow_dtype is not None and dev.type == "cuda":
            self.shadow_flat = torch.empty(self.numel, dtype=shadow_dtype, device=dev)
            off = 0
            for p in self.params:
                self.shadow_views.append(self.shadow_flat[off: off + p.numel()].view_as(p))
                off += p.numel()
        self.group = process_group
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._work = None

    @property
    def nbytes(self) -> int:
        return self.numel * 4

    def zero_(self) -> None:
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def begin_step(self) -> None:
        """Call before the forward pass of a (non-accumulating) step: drops ``p.grad`` so that autograd ASSIGNS this
        step's gradients instead of launching one ``add_`` per parameter into the zeroed views, and refreshes the
        low-precision parameter shadows.  ``gather_()`` after backward brings the gradients into the flat buffer."""
        for p in self.params:
            p.grad = None
        if self.shadow_flat is not None:
            with torch.no_grad():
                torch._foreach_copy_(self.shadow_views, [p.detach() for p in self.params])
            for p, v in zip(self.params, self.shadow_views):
                p._aga_shadow = (p._version, v)

    def gather_(self) -> None:
        """p.grad (whatever autograd produced) -> the flat buffer, with one multi-tensor copy; p.grad become views again."""
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                src.append(g)
                dst.append(v)
            p.grad = v
        if src:
            with torch.no_grad():
                torch._foreach_copy_(dst, src)

    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def all_reduce_mean_async(self) -> None:
        """SUM over ranks then / world (== DDP's gradient averaging).  Returns immediately."""
        ws = self.world_size()
        if ws == 1:
            return
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self.comm_stream):
                self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self._work.wait()  # orders the comm stream after NCCL's internal stream
                self.flat.div_(ws)
        else:  # gloo / CPU tests
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(ws)

    def wait(self) -> None:
        if self.comm_stream is not None and self.world_size() > 1:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """torch.nn.utils.clip_grad_norm_ on the flat buffer: one norm, one scale, no host sync."""
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
