"""Host-side mirror of the reference's Whisper nn modules (whisper/whisper/model.py), with the attention
core replaced by the aga_b200 CUDA kernels.

Only the module tree that the shipped SEAME recipes instantiate is mirrored (``pe_whisper: false``, no
side network — SURVEY.md §2a #1); parameter names and shapes are identical to the reference so that
``state_dict`` keys (``blocks.N.attn.query.weight``, ``blocks.N.adapter_attn.model.0.weight`` …) and therefore
checkpoints are interchangeable.  Everything except ``MultiHeadAttention.qkv_attention`` stays plain PyTorch
(cuBLAS GEMMs on frozen weights, LayerNorm, GELU): plumbing around the hot path, not the product.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import ops


@dataclass
class ModelDimensions:  # whisper/model.py:16-27
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int


def _dims(state, heads, layers, n_mels=80, n_vocab=51865):
    return ModelDimensions(n_mels, 1500, state, heads, layers, n_vocab, 448, state, heads, layers)


# SURVEY.md Appendix A.1 (checkpoint `dims`; head dim is 64 for every size)
MODEL_DIMS: Dict[str, ModelDimensions] = {
    "tiny": _dims(384, 6, 4),
    "base": _dims(512, 8, 6),
    "small": _dims(768, 12, 12),
    "medium": _dims(1024, 16, 24),
    "large-v1": _dims(1280, 20, 32),
    "large-v2": _dims(1280, 20, 32),
    "large": _dims(1280, 20, 32),
    # documented extrapolation (BASELINE.json config 4): large-v2 body with a 128-bin mel stem
    "large-v2-mel128": _dims(1280, 20, 32, n_mels=128),
}


def available_models():
    return list(MODEL_DIMS)


def _native(x: Tensor) -> bool:
    """True when ``x`` takes the library's fused CUDA paths.  oracle/torch_port.patched_ops() (the eager-PyTorch restatement
    of the reference, timed as the CPU baseline and as the eager-GPU comparator) replaces this with ``False`` so that the
    mirror modules run the reference's own op sequence (separate q/k/v Linears, nn.Conv1d stem, unfused residual adds, fp32
    logits) on either device; the product never does."""
    return x.is_cuda


class LayerNorm(nn.LayerNorm):
    """fp32 statistics whatever the activation dtype (whisper/model.py:30-32), one CUDA kernel each way."""

    def forward(self, x: Tensor) -> Tensor:
        return ops.layer_norm(x, self.weight, self.bias, self.eps)

    def with_residual(self, x: Tensor):
        """(LN(x), x') with x' == x: feed x' to the residual add of the branch (`x = x + f(ln(x))`) and the backward
        pass adds the residual path's gradient inside the LayerNorm-backward kernel."""
        if _native(x) and torch.is_grad_enabled() and x.requires_grad:
            return ops.layer_norm_residual(x, self.weight, self.bias, self.eps)
        return self.forward(x), x


def cast_param(owner: nn.Module, slot: str, p: Optional[Tensor], dtype: torch.dtype) -> Optional[Tensor]:
    """``p.to(dtype)`` as the reference writes it (whisper/model.py:35-49), except that the cast of a FROZEN
    parameter is done once and kept: under ``--freeze_param`` every Whisper weight is constant, and re-casting
    all of them fp32->bf16 on every call was 10 % of the training step.  Same values, no graph edge needed."""
    if p is None or p.dtype == dtype:
        return p
    if p.requires_grad and torch.is_grad_enabled():
        return p.to(dtype)
    c = owner.__dict__.get(slot)
    if c is None or c[0] != p._version or c[1] != p.data_ptr() or c[2].dtype != dtype or c[2].device != p.device:
        c = (p._version, p.data_ptr(), p.detach().to(dtype))
        owner.__dict__[slot] = c
    return c[2]


class Linear(nn.Linear):
    """Weights follow the activation dtype (whisper/model.py:35-41)."""

    def forward(self, x: Tensor) -> Tensor:
        return F.linear(x, cast_param(self, "_w_cast", self.weight, x.dtype), cast_param(self, "_b_cast", self.bias, x.dtype))


def _no_grad_needed(*params) -> bool:
    """True when none of ``params`` will receive a gradient from this call: frozen (--freeze_param), or autograd is off
    (inference / decoding under ``torch.no_grad()``)."""
    return (not torch.is_grad_enabled()) or not any(p is not None and p.requires_grad for p in params)


def linear_plus_residual(lin: "Linear", x: Tensor, residual: Tensor) -> Tensor:
    """``residual + lin(x)``; for a frozen Linear on a CUDA device the add rides the GEMM (ops.linear_residual)."""
    frozen = _no_grad_needed(lin.weight, lin.bias)
    if not (_native(x) and frozen):
        return residual + lin(x)
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    if dt not in (torch.float32, torch.bfloat16):
        return residual + lin(x)
    return ops.linear_residual(x.to(dt), cast_param(lin, "_w_cast", lin.weight, dt), cast_param(lin, "_b_cast", lin.bias, dt),
                               residual.to(dt))


class Conv1d(nn.Conv1d):
    def _conv_forward(self, x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
        return super()._conv_forward(x, cast_param(self, "_w_cast", weight, x.dtype),
                                     cast_param(self, "_b_cast", bias, x.dtype))


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> Tensor:
    """Fixed positional table of the audio encoder (whisper/model.py:53-59)."""
    half = channels // 2
    inv = torch.exp(-(math.log(max_timescale) / (half - 1)) * torch.arange(half))
    ang = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([ang.sin(), ang.cos()], dim=1)


class MultiHeadAttention(nn.Module):
    """whisper/model.py:62-109 with ``qkv_attention`` served by the CUDA library.

    ``forward`` keeps the reference signature and its ``(out, second)`` return.  The reference always
    materialises the full fp32 ``qk`` as ``second`` (and throws it away everywhere except the decoder
    self-attention); here ``second`` is what the export policy asks for:

      export = None                     -> second is None (encoder, cross attention)
      export = ("logits" | "probs", None)      -> full (B,H,Tq,Tk) map, same values as the reference's qk / w
      export = ("logits" | "probs", (lo, hi))  -> only key columns [lo,hi): (B,H,Tq,hi-lo)
    """

    def __init__(self, n_state: int, n_head: int):
        super().__init__()
        self.n_head = n_head
        self.query = Linear(n_state, n_state)
        self.key = Linear(n_state, n_state, bias=False)
        self.value = Linear(n_state, n_state)
        self.out = Linear(n_state, n_state)
        self.export: Optional[Tuple[str, Optional[Tuple[int, int]]]] = None
        self.guided: Optional[Tuple[Tensor, bool]] = None  # (pattern (B,T,2), early layer): fuse the guided loss' reduction
        self.head_sel: Optional[Tensor] = None
        self.impl = "auto"

    def forward(self, x: Tensor, xa: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                kv_cache: Optional[dict] = None, residual: Optional[Tensor] = None, kv_len: Optional[Tensor] = None):
        """``residual``: when given, the first return value is ``residual + out`` (the block's `x = x + attn(...)`, with
        the add folded into the output projection's GEMM).  ``kv_len`` (device int32 scalar): number of keys that really
        exist when the key sequence is zero-padded to a static length (graphed.BucketedTrainStep); non-causal only."""
        if kv_cache is None and _native(x) and self._frozen():
            return self._forward_packed(x, xa, mask, residual, kv_len)
        if kv_len is not None:
            raise ops.L.AgaError("kv_len needs the fused CUDA path (frozen projections, CUDA tensors)")
        q = self.query(x)
        if kv_cache is None or xa is None or self.key not in kv_cache:
            src = x if xa is None else xa
            k = self.key(src)
            v = self.value(src)
            if kv_cache is not None and xa is not None:
                kv_cache[self.key], kv_cache[self.value] = k, v  # cross-attention K/V are computed once
        else:
            k, v = kv_cache[self.key], kv_cache[self.value]
        wv, second = self.qkv_attention(q, k, v, mask)
        return (self.out(wv) if residual is None else linear_plus_residual(self.out, wv, residual)), second

    # ---- frozen projections (the --freeze_param adapter policy): one [q|k|v] (or [k|v]) GEMM, attention on the packed
    #      result, one packed gradient -> one dgrad GEMM.  Same arithmetic as three F.linear calls on row blocks of
    #      the concatenated weight.
    def _frozen(self) -> bool:
        return _no_grad_needed(self.query.weight, self.query.bias, self.key.weight, self.value.weight, self.value.bias)

    def _packed_weights(self, dtype: torch.dtype, with_q: bool):
        mods = ([self.query] if with_q else []) + [self.key, self.value]
        sig = (dtype, with_q) + tuple((m.weight._version, m.weight.data_ptr(), 0 if m.bias is None else m.bias._version)
                                      for m in mods)
        c = self.__dict__.get("_packed_cache", {}).get(with_q)
        if c is None or c[0] != sig:
            w = torch.cat([m.weight.detach() for m in mods], dim=0).to(dtype)
            b = torch.cat([torch.zeros_like(m.weight[:, 0]) if m.bias is None else m.bias.detach() for m in mods]).to(dtype)
            c = (sig, w, b)
            self.__dict__.setdefault("_packed_cache", {})[with_q] = c
        return c[1], c[2]

    def _forward_packed(self, x: Tensor, xa: Optional[Tensor], mask: Optional[Tensor], residual: Optional[Tensor] = None,
                        kv_len: Optional[Tensor] = None):
        kind, cols = self.export if self.export is not None else (None, None)
        if xa is None:
            w, b = self._packed_weights(x.dtype, True)
            qkv = F.linear(x, w, b)
            causal = mask is not None
            out, _lse, second = ops.qkv_attention_packed(qkv, self.n_head, causal=causal, export=kind, export_cols=cols,
                                                         head_sel=self.head_sel, impl=self.impl,
                                                         kv_len=None if causal else kv_len,
                                                         guided=self.guided if causal else None)
        else:
            w, b = self._packed_weights(x.dtype, False)
            q = self.query(x)
            kv = F.linear(xa.to(x.dtype), w, b)
            out, _lse, second = ops.qkv_attention_packed(kv, self.n_head, q=q, causal=False, export=kind, export_cols=cols,
                                                         head_sel=self.head_sel, impl=self.impl, kv_len=kv_len)
        return (self.out(out) if residual is None else linear_plus_residual(self.out, out, residual)), second

    def step(self, x: Tensor, past_k: Optional[Tensor] = None, past_v: Optional[Tensor] = None,
             cross_kv: Optional[Tuple[Tensor, Tensor]] = None, residual: Optional[Tensor] = None):
        """Incremental attention for KV-cached decoding (the reference recomputes the whole prefix
        on every step, whisper_decoder.py:172-244).  Self attention (``cross_kv`` None): the keys / values of the new
        tokens ``x`` (n, t_new, D) are appended to ``past_k`` / ``past_v`` and returned; cross attention: the K / V of
        the encoder output, projected once, are passed in.  Same arithmetic as ``forward`` on the full prefix."""
        q = self.query(x)
        if cross_kv is not None:
            k, v = cross_kv
            out, _, _ = ops.qkv_attention(q, k, v, self.n_head, causal=False, impl=self.impl)
            return self.out(out) if residual is None else linear_plus_residual(self.out, out, residual)
        k, v = self.key(x), self.value(x)
        if past_k is not None:
            k, v = torch.cat([past_k, k], dim=1), torch.cat([past_v, v], dim=1)
        if q.shape[1] > 1 and q.shape[1] != k.shape[1]:
            raise ValueError("chunked prefill is not supported: feed the whole prefix first, then one token per step")
        out, _, _ = ops.qkv_attention(q, k, v, self.n_head, causal=q.shape[1] > 1, impl=self.impl)
        return self.out(out), k, v

    def step_static(self, x: Tensor, kv_buf: Tensor, pos: Tensor, kv_len: Tensor, residual: Optional[Tensor] = None) -> Tensor:
        """One decoding step with STATIC shapes (graph-capturable): x (n, 1, D) is the new token's activation, ``kv_buf``
        (n, max_len, 2D) = [K | V] the preallocated self-attention cache, ``pos`` (1,) int64 the row the new key / value go
        to, ``kv_len`` = pos + 1 as the device scalar the attention kernel masks with.  Same arithmetic as ``step``, in
        four launches: one [q|k|v] GEMM, one cache-row write, the attention kernel on the packed cache, the output
        projection with ``residual`` (if given) added inside the GEMM."""
        D = x.shape[-1]
        w, b = self._packed_weights(x.dtype, True)
        qkv = F.linear(x, w, b)
        kv_buf.index_copy_(1, pos, qkv[..., D:])
        out, _, _ = ops.qkv_attention_packed(kv_buf, self.n_head, q=qkv[..., :D], causal=False, impl=self.impl, kv_len=kv_len)
        return self.out(out) if residual is None else linear_plus_residual(self.out, out, residual)

    def qkv_attention(self, q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor] = None):
        # The only mask the reference ever passes is TextDecoder.mask = triu(-inf) (whisper/model.py:322,103),
        # i.e. "mask is not None" <=> causal over equal-length q/k.
        causal = mask is not None and q.shape[1] == k.shape[1]
        kind, cols = self.export if self.export is not None else (None, None)
        out, _lse, second = ops.qkv_attention(q, k, v, self.n_head, causal=causal, export=kind, export_cols=cols,
                                              head_sel=self.head_sel, impl=self.impl)
        return out, second


class Adapter(nn.Module):
    """Bottleneck adapter x + W2 gelu(W1 x), bottleneck idim//4 (whisper/model.py:181-194)."""

    def __init__(self, idim: int, bottleneck_dim: Optional[int] = None) -> None:
        super().__init__()
        bottleneck_dim = bottleneck_dim or idim // 4
        self.model = nn.Sequential(nn.Linear(idim, bottleneck_dim), nn.GELU(), nn.Linear(bottleneck_dim, idim))

    def forward(self, x: Tensor) -> Tensor:
        return x + self.model(x)


class ResidualAttentionBlock(nn.Module):
    """whisper/model.py:195-248: returns (x, second output of the SELF attention)."""

    def __init__(self, n_state: int, n_head: int, adapter: bool = False, pe_whisper: bool = False,
                 cross_attention: bool = False):
        super().__init__()
        if pe_whisper:
            raise NotImplementedError("pe_whisper (MultiHeadAttentionPE ablation) is outside the hot path")
        self.adapter_flag = adapter
        self.attn = MultiHeadAttention(n_state, n_head)
        self.attn_ln = LayerNorm(n_state)
        if adapter:
            self.adapter_attn = Adapter(n_state)
            self.adapter_attn_ln = LayerNorm(n_state)
        self.cross_attn = MultiHeadAttention(n_state, n_head) if cross_attention else None
        self.cross_attn_ln = LayerNorm(n_state) if cross_attention else None
        self.mlp = nn.Sequential(Linear(n_state, 4 * n_state), nn.GELU(), Linear(4 * n_state, n_state))
        self.mlp_ln = LayerNorm(n_state)
        if adapter:
            self.adapter_mlp = Adapter(n_state)
            self.adapter_mlp_ln = LayerNorm(n_state)

    def forward(self, x: Tensor, xa: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                kv_cache: Optional[dict] = None, kv_len: Optional[Tensor] = None, xa_len: Optional[Tensor] = None):
        """``kv_len`` / ``xa_len``: true lengths of x / xa when they are zero-padded to static shapes (device scalars)."""
        x, second, _ = self.forward_chain(x, xa, mask, kv_cache, kv_len, xa_len)
        return x, second

    def forward_chain(self, x: Tensor, xa: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                      kv_cache: Optional[dict] = None, kv_len: Optional[Tensor] = None, xa_len: Optional[Tensor] = None,
                      pre: Optional[Tensor] = None, next_ln: Optional["LayerNorm"] = None):
        """``forward`` for a chain of blocks.  ``pre``: ``self.attn_ln(x)`` already computed by the previous block's last
        kernel; ``next_ln``: the LayerNorm the caller applies to this block's output next (the next block's ``attn_ln``,
        ``ln_post`` / ``ln`` after the last one) — with adapters its result comes out of the adapter post-LN kernel and
        is returned as the third value (else None).  Every adapter post-LN is followed by a pre-LN of the same tensor
        (whisper/model.py:231-246): each such pair is one kernel and one read of the row."""
        if pre is None:
            y, x = self.attn_ln.with_residual(x)
        else:
            y = pre
        x, second = self.attn(y, mask=mask, kv_cache=kv_cache, residual=x, kv_len=kv_len)  # x + attn(...)
        y = None
        if self.adapter_flag:  # post-LN replaces x (:234-236)
            x, y = self._adapter_ln(self.adapter_attn, self.adapter_attn_ln, x,
                                    next_ln=self.cross_attn_ln if self.cross_attn is not None else self.mlp_ln)
        if self.cross_attn is not None:
            if y is None:
                y, x = self.cross_attn_ln.with_residual(x)
            x = self.cross_attn(y, xa, kv_cache=kv_cache, residual=x, kv_len=xa_len)[0]
            y = None
        if y is None:
            y, x = self.mlp_ln.with_residual(x)
        x = self._mlp_residual(y, x)  # x + mlp(...)
        y = None
        if self.adapter_flag:
            x, y = self._adapter_ln(self.adapter_mlp, self.adapter_mlp_ln, x, next_ln=next_ln)
        return x, second, y

    def step(self, x: Tensor, xa: Tensor, cache: Optional[Tuple[Tensor, Tensor, Tensor, Tensor]] = None):
        """``forward`` for the new tokens only.  ``cache`` = (self K, self V, cross K, cross V) of the prefix, or None on
        the first call (prefill: x is the whole prefix); returns (x, updated cache)."""
        a, k, v = self.attn.step(self.attn_ln(x), *((cache[0], cache[1]) if cache is not None else (None, None)))
        x = x + a
        if self.adapter_flag:
            x = self._adapter_ln(self.adapter_attn, self.adapter_attn_ln, x)[0]
        if cache is not None:
            kc, vc = cache[2], cache[3]
        else:
            src = xa.to(x.dtype)
            kc, vc = self.cross_attn.key(src), self.cross_attn.value(src)
        x = x + self.cross_attn.step(self.cross_attn_ln(x), cross_kv=(kc, vc))
        x = x + self.mlp(self.mlp_ln(x))
        if self.adapter_flag:
            x = self._adapter_ln(self.adapter_mlp, self.adapter_mlp_ln, x)[0]
        return x, (k, v, kc, vc)

    def step_static(self, x: Tensor, kv_buf: Tensor, pos: Tensor, kv_len: Tensor, cross_kv: Tuple[Tensor, Tensor]):
        """``step`` on a preallocated [K | V] cache (see MultiHeadAttention.step_static): no shape depends on the position,
        and every residual add rides a GEMM (17 launches per block instead of 27)."""
        x = self.attn.step_static(self.attn_ln(x), kv_buf, pos, kv_len, residual=x)
        if self.adapter_flag:
            x = self._adapter_ln(self.adapter_attn, self.adapter_attn_ln, x)[0]
        x = self.cross_attn.step(self.cross_attn_ln(x), cross_kv=cross_kv, residual=x)
        x = self._mlp_residual(self.mlp_ln(x), x)
        if self.adapter_flag:
            x = self._adapter_ln(self.adapter_mlp, self.adapter_mlp_ln, x)[0]
        return x

    def _mlp_residual(self, y: Tensor, x: Tensor) -> Tensor:
        """``x + self.mlp(y)`` (whisper/model.py:242).  Frozen bf16 MLP on a CUDA device: two GEMMs and nothing else — the
        exact GELU in the first GEMM's epilogue (tcgen05, csrc/gemm_gelu.cu), the residual add in the second's."""
        l1, l2 = self.mlp[0], self.mlp[2]
        frozen = _no_grad_needed(l1.weight, l1.bias, l2.weight, l2.bias)
        dt = torch.get_autocast_dtype("cuda") if (y.is_cuda and torch.is_autocast_enabled("cuda")) else y.dtype
        rows = y.numel() // y.shape[-1]
        # (a handful of rows — a decoding step — is latency-bound: 12 CTAs of the tiled kernel walk K serially, 13.8 us
        # against 8 us for cuBLAS's GEMV-shaped kernel + GELU)
        if not (_native(y) and frozen and dt == torch.bfloat16 and rows >= 64):
            return linear_plus_residual(l2, self.mlp[1](l1(y)), x)
        w2 = cast_param(l2, "_w_cast", l2.weight, dt)
        c = self.__dict__.get("_w2t")
        if c is None or c[0] is not w2:
            c = (w2, w2.t().contiguous())
            self.__dict__["_w2t"] = c
        return ops.mlp_residual(y.to(dt), cast_param(l1, "_w_cast", l1.weight, dt), cast_param(l1, "_b_cast", l1.bias, dt),
                                w2, c[1], cast_param(l2, "_b_cast", l2.bias, dt), x.to(dt))

    @staticmethod
    def _adapter_ln(adapter: "Adapter", ln: "LayerNorm", x: Tensor, next_ln: Optional["LayerNorm"] = None):
        """``ln(adapter(x))`` = LN(x + W2 gelu(W1 x)) as one fused autograd node.  With ``next_ln`` (the FROZEN pre-LayerNorm of
        the residual branch that consumes the result next) returns ``(z, next_ln(z))`` from one kernel
        (ops.adapter_layer_norm_pair), else ``(z, None)``."""
        m = adapter.model
        ps = (m[0].weight, m[0].bias, m[2].weight, m[2].bias)
        if not torch.is_grad_enabled() and x.dtype != m[0].weight.dtype:
            # inference / decoding: the low-precision copies of the adapter weights are made once, not per call
            # (8 cast kernels per decoder block and token otherwise)
            ps = tuple(cast_param(lin, slot, t, x.dtype) for lin, slot, t in
                       ((m[0], "_w_cast", m[0].weight), (m[0], "_b_cast", m[0].bias),
                        (m[2], "_w_cast", m[2].weight), (m[2], "_b_cast", m[2].bias)))
        if next_ln is not None and _native(x) and _no_grad_needed(next_ln.weight, next_ln.bias):
            return ops.adapter_layer_norm_pair(x, *ps, ln.weight, ln.bias, ln.eps, next_ln.weight, next_ln.bias, next_ln.eps)
        return ops.adapter_layer_norm(x, *ps, ln.weight, ln.bias, ln.eps), None


def run_blocks(blocks, x: Tensor, final_ln: "LayerNorm", between=None, between_changes_x: bool = True, on_block=None, **kw) -> Tensor:
    """``for block in blocks: x, second = block(x, **kw)`` then ``final_ln(x)``, with each block's leading LayerNorm (and
    the final one) taken from the previous block's last kernel when that block ends in an adapter post-LN
    (ResidualAttentionBlock.forward_chain).  ``between(x)``: applied to x between blocks (ESPnet's dropout); unless the
    caller states that it leaves x as it is (``between_changes_x=False``: p = 0, or eval mode) the chaining is off;
    ``on_block(layer, second)``: receives every block's second output."""
    chain = between is None or not between_changes_x
    pre = None
    n = len(blocks)
    for layer, block in enumerate(blocks):
        nxt = (blocks[layer + 1].attn_ln if layer + 1 < n else final_ln) if chain else None
        x, second, pre = block.forward_chain(x, pre=pre, next_ln=nxt, **kw)
        if on_block is not None:
            on_block(layer, second)
        if between is not None and layer + 1 < n:
            x = between(x)
    return pre if pre is not None else final_ln(x)


class AudioEncoder(nn.Module):
    """whisper/model.py:251-290."""

    def __init__(self, n_mels: int, n_ctx: int, n_state: int, n_head: int, n_layer: int, adapter: bool = False,
                 pe_whisper: bool = False):
        super().__init__()
        self.conv1 = Conv1d(n_mels, n_state, kernel_size=3, padding=1)
        self.conv2 = Conv1d(n_state, n_state, kernel_size=3, stride=2, padding=1)
        self.register_buffer("positional_embedding", sinusoids(n_ctx, n_state))
        self.blocks = nn.ModuleList(
            [ResidualAttentionBlock(n_state, n_head, adapter=adapter, pe_whisper=pe_whisper) for _ in range(n_layer)])
        self.ln_post = LayerNorm(n_state)
        self.n_layer = n_layer

    def _stem_weights(self, dtype: torch.dtype):
        """conv1 / conv2 kernels as GEMM operands: (D, n_mels*3) with columns (c, k), (D, 3*D) with columns (k, c)."""
        w1, w2 = self.conv1.weight, self.conv2.weight
        sig = (dtype, w1._version, w1.data_ptr(), w2._version, w2.data_ptr(), self.conv1.bias._version, self.conv2.bias._version)
        c = self.__dict__.get("_stem_cache")
        frozen = not (w1.requires_grad or w2.requires_grad or self.conv1.bias.requires_grad or self.conv2.bias.requires_grad)
        if c is not None and c[0] == sig and (frozen or not torch.is_grad_enabled()):
            return c[1]
        ws = (w1.reshape(w1.shape[0], -1).to(dtype), self.conv1.bias.to(dtype),
              w2.permute(0, 2, 1).reshape(w2.shape[0], -1).to(dtype), self.conv2.bias.to(dtype))
        if frozen or not torch.is_grad_enabled():
            ws = tuple(t.detach() for t in ws)
            self.__dict__["_stem_cache"] = (sig, ws)
        return ws

    def stem(self, x: Tensor, valid_frames: Optional[Tensor] = None) -> Tensor:
        """gelu(conv2(gelu(conv1(x)))).permute(0, 2, 1) (whisper/model.py:277-279): (B, n_mels, T) -> (B, T', D).

        On a CUDA device both convolutions run as ONE cuBLAS GEMM each on token-major activations (k=3 windows are
        gathered once; conv2's stride-2 windows are three row-strided slices): cuDNN serves these shapes with a
        legacy sm_75 implicit-GEMM kernel that is 8x slower, and the token-major result needs no permute."""
        if not _native(x):
            x = F.gelu(self.conv1(x))
            return F.gelu(self.conv2(x)).permute(0, 2, 1)
        dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
        w1, b1, w2, b2 = self._stem_weights(dt)
        B, C, T = x.shape
        cols = F.pad(x.to(dt), (1, 1)).unfold(2, 3, 1).permute(0, 2, 1, 3).reshape(B, T, C * 3)
        with torch.autocast("cuda", enabled=False):
            h = F.gelu(F.linear(cols, w1, b1))                                   # (B, T, D)
            T2 = (T - 1) // 2 + 1
            hp = F.pad(h, (0, 0, 1, 1))                                          # zero rows at t = -1 and t = T
            if valid_frames is not None:
                # the batch is zero-padded past `valid_frames` mel frames: conv2's window of the last real output reads
                # conv1-output frame `valid_frames`, which the reference's own zero padding supplies as 0
                hp.index_fill_(1, (valid_frames.to(torch.int64) + 1).reshape(1), 0.0)
            # window t = rows 2t, 2t+1, 2t+2 of hp = 3 D CONTIGUOUS elements: one strided copy of overlapping windows
            # (a cat of three row-strided slices ran at a sixth of the copy bandwidth); columns (k, c)
            D = hp.shape[2]
            cols2 = hp.as_strided((B, T2, 3 * D), (hp.stride(0), 2 * D, 1)).contiguous()
            return F.gelu(F.linear(cols2, w2, b2))

    def forward(self, x: Tensor) -> Tensor:
        x = self.stem(x)
        assert x.shape[1:] == self.positional_embedding.shape, "incorrect audio shape"
        x = (x + self.positional_embedding).to(x.dtype)
        return run_blocks(self.blocks, x, self.ln_post)


class TextDecoder(nn.Module):
    """whisper/model.py:293-347 (the upstream forward is broken by the fork's tuple-returning blocks, :339-340;
    this one unpacks the tuple)."""

    def __init__(self, n_vocab: int, n_ctx: int, n_state: int, n_head: int, n_layer: int, pe_whisper: bool = False,
                 adapter: bool = False):
        super().__init__()
        self.token_embedding = nn.Embedding(n_vocab, n_state)
        self.positional_embedding = nn.Parameter(torch.empty(n_ctx, n_state))
        self.blocks = nn.ModuleList(
            [ResidualAttentionBlock(n_state, n_head, adapter=adapter, pe_whisper=pe_whisper, cross_attention=True)
             for _ in range(n_layer)])
        self.ln = LayerNorm(n_state)
        mask = torch.empty(n_ctx, n_ctx).fill_(float("-inf")).triu_(1)
        self.register_buffer("mask", mask, persistent=False)
        self.n_layer = n_layer

    def forward(self, x: Tensor, xa: Tensor, kv_cache: Optional[dict] = None) -> Tensor:
        offset = next(iter(kv_cache.values())).shape[1] if kv_cache else 0
        x = self.token_embedding(x) + self.positional_embedding[offset: offset + x.shape[-1]]
        x = x.to(xa.dtype)
        x = run_blocks(self.blocks, x, self.ln, xa=xa, mask=self.mask, kv_cache=kv_cache)
        return self.vocab_logits(x)

    def vocab_logits(self, x: Tensor, lazy: bool = False):
        """``(x @ token_embedding.weight.T).float()`` (whisper/model.py:343-345); ``lazy``: an ``ops.VocabLogits`` handle
        on the padded GEMM output instead of the fp32 tensor (consumed by ``ops.ls_cross_entropy``).

        n_vocab = 51865 is odd: with a leading dimension that is not a multiple of 16 bytes cuBLAS has no TMA / vector
        path and falls back to a legacy sm_75 kernel (650 us forward, 550 us backward at B*T = 1024, 8x slower than the
        aligned GEMM).  A frozen embedding is therefore multiplied as a copy padded with zero rows to a multiple of 64;
        the extra logit columns are sliced away before anyone sees them (their gradient is zero)."""
        w = self.token_embedding.weight
        n_vocab = w.shape[0]
        pad = (-n_vocab) % 64
        if not _native(x) or pad == 0 or (w.requires_grad and torch.is_grad_enabled()):
            full = x @ cast_param(self, "_emb_cast", w, x.dtype).t()
            return ops.VocabLogits(full, n_vocab) if (lazy and _native(x)) else full.float()
        c = self.__dict__.get("_emb_pad")
        if c is None or c[0] != (w._version, w.data_ptr(), x.dtype):
            wp = torch.zeros(n_vocab + pad, w.shape[1], dtype=x.dtype, device=w.device)
            wp[:n_vocab] = w.detach()
            c = ((w._version, w.data_ptr(), x.dtype), wp)
            self.__dict__["_emb_pad"] = c
        padded = F.linear(x, c[1])
        return ops.VocabLogits(padded, n_vocab) if lazy else padded[..., :n_vocab].float()


class Whisper(nn.Module):
    """Container with the reference's constructor signature (whisper/model.py:485-506)."""

    def __init__(self, dims: ModelDimensions, pe_whisper: bool = False, adapter: bool = False,
                 side_network: bool = False, side_network_conf: Optional[dict] = None):
        super().__init__()
        if side_network:
            raise NotImplementedError("side networks are an unused ablation of the reference (SURVEY.md §2a #1)")
        self.dims = dims
        self.encoder = AudioEncoder(dims.n_mels, dims.n_audio_ctx, dims.n_audio_state, dims.n_audio_head,
                                    dims.n_audio_layer, adapter=adapter, pe_whisper=pe_whisper)
        self.decoder = TextDecoder(dims.n_vocab, dims.n_text_ctx, dims.n_text_state, dims.n_text_head,
                                   dims.n_text_layer, pe_whisper=pe_whisper, adapter=adapter)

    @property
    def is_multilingual(self) -> bool:
        return self.dims.n_vocab == 51865


def seeded_init_(module: nn.Module, seed: int = 0) -> nn.Module:
    """Deterministic init that depends only on parameter NAMES and SHAPES (so the reference module tree and this
    mirror, which share both, get identical weights): used where no checkpoint is available offline."""
    for name, p in sorted(module.named_parameters(), key=lambda kv: kv[0]):
        g = torch.Generator().manual_seed((hash_name(name) + seed) % (2 ** 31))
        with torch.no_grad():
            if p.dim() >= 2:
                fan_in = p.shape[1] * (p.shape[2] if p.dim() > 2 else 1)
                vals = torch.randn(p.shape, generator=g) / math.sqrt(max(1, fan_in))
                if name.endswith("token_embedding.weight"):
                    vals = vals * math.sqrt(fan_in) * 0.02
            elif name.endswith("ln.weight") or name.endswith("ln_post.weight") or "_ln.weight" in name:
                vals = 1.0 + 0.02 * torch.randn(p.shape, generator=g)
            else:
                vals = 0.02 * torch.randn(p.shape, generator=g)
            p.copy_(vals.to(p.dtype))
    return module


def hash_name(name: str) -> int:
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


_ALLOW_RANDOM_INIT = False


def allow_random_init(flag: bool = True) -> None:
    """Explicit opt-in to name-seeded random weights when no checkpoint exists (tests, bench.py, oracle/make_golden.py)."""
    global _ALLOW_RANDOM_INIT
    _ALLOW_RANDOM_INIT = bool(flag)


def random_init_allowed() -> bool:
    import os
    return _ALLOW_RANDOM_INIT or os.environ.get("AGA_ALLOW_RANDOM_INIT", "0") == "1"


def load_model(name: str, adapter: bool = False, pe_whisper: bool = False, side_network: bool = False,
               side_network_conf: Optional[dict] = None, device=None, download_root: Optional[str] = None,
               in_memory: bool = False, seed: int = 0) -> Whisper:
    """Signature of the fork's whisper.load_model (whisper/__init__.py:182-268).

    ``name`` is a model size or a path to an OpenAI-format checkpoint ({"dims", "model_state_dict"}); a size name
    is looked up as ``<download_root>/<name>.pt`` (default root ``~/.cache/whisper``, as upstream).  There is no
    network here, so nothing is downloaded: when no checkpoint file exists this RAISES, like the reference does when
    its download fails — a drop-in that silently trained a random Whisper would be worse than an error.  Tests, the
    benchmark and the golden generator opt in to the deterministic name-seeded initialisation explicitly
    (``allow_random_init()`` / ``AGA_ALLOW_RANDOM_INIT=1``), and it is logged whenever it is taken.
    """
    import os

    path = None
    root = download_root if download_root is not None else os.path.join(os.path.expanduser("~"), ".cache", "whisper")
    if os.path.isfile(name):
        path = name
    elif os.path.isfile(os.path.join(root, f"{name}.pt")):
        path = os.path.join(root, f"{name}.pt")
    if path is not None:
        ckpt = torch.load(path, map_location="cpu")
        dims = ModelDimensions(**ckpt["dims"])
        model = Whisper(dims, pe_whisper, adapter, side_network, side_network_conf)
        seeded_init_(model, seed)  # adapter params are absent from stock checkpoints
        model.load_state_dict(ckpt["model_state_dict"], strict=not adapter)
    else:
        if name not in MODEL_DIMS:
            raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
        if not random_init_allowed():
            raise RuntimeError(
                f"no Whisper checkpoint for '{name}' (looked for {os.path.join(root, name + '.pt')}); this build cannot "
                "download one.  Put the OpenAI checkpoint there / pass download_dir, or opt in to seeded random weights "
                "with aga_b200.whisper_model.allow_random_init() or AGA_ALLOW_RANDOM_INIT=1 (tests and benchmarks only)")
        import logging
        logging.getLogger("aga_b200").warning(
            "whisper '%s': no checkpoint found, using SEEDED RANDOM weights (seed %d) — explicit opt-in", name, seed)
        model = Whisper(MODEL_DIMS[name], pe_whisper, adapter, side_network, side_network_conf)
        seeded_init_(model, seed)
    return model.to(device) if device is not None else model
