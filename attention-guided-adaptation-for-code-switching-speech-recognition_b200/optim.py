"""AdamW for the trainable (adapter) parameters on the flat buffers of ``parallel.FlatGradBucket``.

Reference: the tail of ``Trainer.train_one_epoch``'s inner step (espnet2/train/trainer.py:649-716) — ``clip_grad_norm_``, the
"norm not finite -> skip the update" rule, ``optimizer.step()`` with the recipe's ``optim: adamw`` — and the fp32 -> bf16
re-cast of the updated parameters that autocast performs on their next use.  Stock PyTorch spends ~25 small launches on
that chain per step (multi-tensor norm, clip, 8 chunked fused-AdamW launches over ~200 tensors, 5 multi-tensor casts);
here it is two launches of the library (csrc/flat_adamw.cu): the gradient norm, then ONE pass that clips, updates and
writes the bf16 copies.  Same arithmetic as ``torch.optim.AdamW(fused=True)`` statement for statement
(tests/test_gpu_e2e.py::test_flat_adamw_matches_torch_adamw).

The parameters are moved into one flat fp32 buffer laid out like the bucket's gradient buffer (``p.data`` becomes a view:
modules, ``state_dict`` keys and values are unchanged).  ``state_dict`` / ``load_state_dict`` use torch's per-parameter
format (``step``, ``exp_avg``, ``exp_avg_sq``), so a checkpoint written by ``torch.optim.AdamW`` resumes here and back.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, bucket, lr=1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        params = list(bucket.params)
        dev = bucket.flat.device
        if dev.type != "cuda":
            raise L.AgaError("FlatAdamW runs on the CUDA library only (no CPU path); use torch.optim.AdamW on the CPU")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or eps < 0.0 or weight_decay < 0.0:
            raise ValueError("invalid AdamW hyper-parameters")
        # the parameter group lists them in MODEL order (the bucket stores them back to front), so that state_dict indices
        # agree with a torch.optim.AdamW built from the same `[p for p in model.parameters() if p.requires_grad]`
        super().__init__(list(reversed(params)), dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.bucket = bucket
        n = bucket.numel
        self.pflat = torch.empty(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_t = torch.zeros((), dtype=torch.float32, device=dev)       # updates applied so far
        self.grad_norm = torch.zeros((), dtype=torch.float32, device=dev)    # of the last step() call
        self.skipped = torch.zeros((), dtype=torch.float32, device=dev)      # updates skipped (norm not finite)
        self._lr_t = torch.zeros((), dtype=torch.float32, device=dev)
        import ctypes as C
        need = C.c_size_t()
        L.check(L.lib().aga_flat_grad_norm_workspace_bytes(C.byref(need)), "aga_flat_grad_norm_workspace_bytes")
        self._ws = torch.zeros((need.value + 7) // 8, dtype=torch.int64, device=dev)  # zero-initialised once, 8-byte aligned
        # parameters -> views of one flat buffer (bucket order); values, shapes and the Parameter objects stay
        with torch.no_grad():
            for p, off in zip(params, bucket.offsets):
                view = self.pflat[off: off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                self.state[p] = {"step": self.step_t, "exp_avg": self.exp_avg[off: off + p.numel()].view_as(p),
                                 "exp_avg_sq": self.exp_avg_sq[off: off + p.numel()].view_as(p)}
        self._refresh_shadow()
        bucket.shadow_owner = self  # begin_step() leaves the bf16 copies to this optimizer (see shadow_is_fresh)

    # ---- bf16 copies of the parameters (what ops.cast_trainable serves during the forward / backward pass)
    def _refresh_shadow(self) -> None:
        b = self.bucket
        if b.shadow_flat is not None:
            b.shadow_flat.copy_(self.pflat)
        self._versions = [p._version for p in b.params]
        self._ptrs = [p.data_ptr() for p in b.params]

    def shadow_is_fresh(self) -> bool:
        """Called by the bucket at the start of every step.  The update kernel keeps the bf16 copies current (it writes
        through raw pointers and does not bump tensor versions); if somebody else wrote the parameters since (checkpoint
        load, manual edit: their version moved) the copies are re-made here.  Returns True: they are valid now."""
        b = self.bucket
        ok = all(p._version == v and p.data_ptr() == q for p, v, q in zip(b.params, self._versions, self._ptrs))
        if not ok:
            if any(p.data_ptr() != q for p, q in zip(b.params, self._ptrs)):
                raise L.AgaError("a trainable parameter was re-allocated after FlatAdamW took it over (p.data = ...): "
                                 "write into it in place (p.copy_) instead")
            self._refresh_shadow()
        return True

    # ---- the update
    def _lr_tensor(self) -> torch.Tensor:
        lr = self.param_groups[0]["lr"]
        if isinstance(lr, torch.Tensor):
            if lr.device == self._lr_t.device and lr.dtype == torch.float32:
                return lr  # an LR scheduler that fills it in place is followed by a captured graph too
            self._lr_t.copy_(lr)
        else:
            self._lr_t.fill_(float(lr))
        return self._lr_t

    @torch.no_grad()
    def step(self, closure=None, max_norm: Optional[float] = None):
        """Gradient norm of the bucket's flat buffer -> (optional) clip to ``max_norm`` -> AdamW, skipped entirely when the
        norm is not finite.  Returns the norm (device scalar, valid after the stream reaches this point)."""
        if closure is not None:
            raise L.AgaError("FlatAdamW does not re-evaluate closures")
        if len(self.param_groups) != 1:
            raise L.AgaError("FlatAdamW keeps one parameter group (the adapters of the recipe share one)")
        g = self.param_groups[0]
        ops = L.torch_ops()
        ops.flat_grad_norm(self.bucket.flat, self.grad_norm, self.step_t, self.skipped, self._ws)
        ops.flat_adamw(self.pflat, self.bucket.flat, self.exp_avg, self.exp_avg_sq, self._lr_tensor(), float(g["betas"][0]),
                       float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), self.step_t, self.grad_norm,
                       float(max_norm) if max_norm is not None else 0.0, self.bucket.shadow_flat)
        return self.grad_norm

    # ---- checkpoints in torch.optim.AdamW's format
    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)  # replaces the per-parameter tensors by copies: pull them back into the flat buffers
        with torch.no_grad():
            step = None
            for p, off in zip(self.bucket.params, self.bucket.offsets):
                st = self.state.get(p, {})
                if "exp_avg" in st:
                    self.exp_avg[off: off + p.numel()].view_as(p).copy_(st["exp_avg"])
                    self.exp_avg_sq[off: off + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
                    step = st["step"] if step is None else step
                self.state[p] = {"step": self.step_t, "exp_avg": self.exp_avg[off: off + p.numel()].view_as(p),
                                 "exp_avg_sq": self.exp_avg_sq[off: off + p.numel()].view_as(p)}
            if step is not None:
                self.step_t.copy_(torch.as_tensor(step, dtype=torch.float32))
