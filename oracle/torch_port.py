"""Eager-PyTorch port of the reference's hot-path functions (TEST INFRASTRUCTURE / CPU baseline only).

Each function restates, op for op, what the reference executes in PyTorch eager mode, so that timing it on the
host cores is "the reference's CPU path", and running it on the GPU is the reference's eager-CUDA path ("G-eager",
BASELINE.md §4).  ``patched_ops()`` swaps these functions into the product's host-side mirror modules for that
purpose; the product itself never imports this file.
"""
from __future__ import annotations

import contextlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import aga_oracle as O  # noqa: E402


def log_mel_spectrogram(audio, ilens=None, n_mels=80, filters=None, valid_samples=None, algo=None):
    """espnet2/asr/encoder/whisper_encoder.py:105-135, verbatim sequence of torch ops."""
    assert valid_samples is None, "the reference has no static-shape padding"
    window = torch.hann_window(400).to(audio.device)
    stft = torch.stft(audio, 400, 160, window=window, return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    if filters is None:
        filters = torch.from_numpy(O.mel_filterbank(n_mels)).to(audio.device)
    mel_spec = filters @ magnitudes
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()
    olens = None if ilens is None else ilens // 160
    log_spec = torch.maximum(log_spec, log_spec.view(audio.size(0), -1).max(dim=-1)[0][:, None, None] - 8.0)
    return (log_spec + 4.0) / 4.0, olens


def qkv_attention(q, k, v, n_head, causal=False, export=None, export_cols=None, head_sel=None, impl="auto", kv_len=None):
    """whisper/whisper/model.py:93-109: materialised scores, fp32 softmax, always returns the full qk."""
    n_batch, n_ctx, n_state = q.shape
    scale = (n_state // n_head) ** -0.25
    qh = q.view(*q.shape[:2], n_head, -1).permute(0, 2, 1, 3) * scale
    kh = k.view(*k.shape[:2], n_head, -1).permute(0, 2, 3, 1) * scale
    vh = v.view(*v.shape[:2], n_head, -1).permute(0, 2, 1, 3)
    qk = qh @ kh
    if causal:
        mask = torch.empty(n_ctx, n_ctx, device=q.device).fill_(-np.inf).triu_(1)
        qk = qk + mask
    qk = qk.float()
    w = F.softmax(qk, dim=-1).to(q.dtype)
    out = (w @ vh).permute(0, 2, 1, 3).flatten(start_dim=2)
    second = None
    if export is not None:
        second = qk if export == "logits" else w.float()
        if export_cols is not None:
            second = second[..., export_cols[0]:export_cols[1]]
    return out, None, second


def qkv_attention_packed(x, n_head, q=None, causal=False, export=None, export_cols=None, head_sel=None, impl="auto", kv_len=None):
    """The same attention on a packed [q|k|v] (or [k|v] + q) projection: slices, then the reference op sequence."""
    D = n_head * 64
    if q is None:
        q, k, v = x[..., :D], x[..., D:2 * D], x[..., 2 * D:]
    else:
        k, v = x[..., :D], x[..., D:]
    return qkv_attention(q.contiguous(), k.contiguous(), v.contiguous(), n_head, causal, export, export_cols, head_sel, impl)


def attention_pattern(tokens, lid_table, c=0.6):
    """espnet2/asr/espnet_model.py:236-275: a Python loop per utterance (the tokenizer strings are replaced by the
    LID table, the per-token loop and the host round trip are kept)."""
    lid = lid_table.cpu().numpy()
    rows = [torch.from_numpy(O.create_attention_pattern(t, lid, c)) for t in tokens.cpu().numpy()]
    return torch.stack(rows).to(tokens.device)


def guided_loss(slab, pattern, head_mask, n_early=2):
    """espnet2/asr/espnet_model.py:463-530 on the two columns it reads, same op sequence (full-size temporaries
    of the reference collapse to the slab; autograd does the backward)."""
    L, B, H, T, _ = slab.shape
    A = slab.permute(1, 0, 2, 3, 4).clone()  # (B,L,H,T,2)
    pat = torch.zeros(B, L, T, 2, device=slab.device)
    pat[:, n_early:] = pattern.to(slab.device)[:, None]
    rep = pat.unsqueeze(2).repeat(1, 1, H, 1, 1)
    A[torch.isinf(rep)] = 0.0
    A[torch.isinf(A)] = 0.0
    rep[torch.isinf(rep)] = 0.0
    e = F.mse_loss(A, rep, reduction="none")
    r = torch.sum(e, dim=-1)
    m = torch.sum(r, dim=-1) / torch.count_nonzero(r, dim=-1)
    masked = head_mask.to(slab.device) * m
    return torch.mean(torch.sum(masked, dim=[-1, -2]))


def head_vote(probs, counts=None):
    c = torch.from_numpy(O.new_check_attention_language(probs.detach().cpu().numpy())).to(torch.int32)
    s1, s2 = O.head_vote_sums(probs.detach().cpu().numpy())
    return torch.from_numpy((s1 > s2).astype(np.uint8)), c if counts is None else counts + c.to(counts.device)


def layer_norm(x, weight, bias, eps=1e-5):
    """whisper/whisper/model.py:30-32: up-cast, F.layer_norm, down-cast."""
    return F.layer_norm(x.float(), (x.shape[-1],), weight, bias, eps).type(x.dtype)


def adapter_layer_norm(x, w1, b1, w2, b2, gamma, beta, eps=1e-5):
    """whisper/whisper/model.py:193 (Adapter.forward) followed by the post-LayerNorm of :234-236 / :244-246."""
    y = x + F.linear(F.gelu(F.linear(x, w1.to(x.dtype), b1.to(x.dtype))), w2.to(x.dtype), b2.to(x.dtype))
    return layer_norm(y, gamma, beta, eps)


@contextlib.contextmanager
def patched_ops():
    """Route the mirror modules' hot-path calls to the eager port (CPU baseline / eager-GPU comparator only)."""
    import aga_b200
    from aga_b200 import ops, whisper_model
    saved_native = whisper_model._native
    whisper_model._native = lambda x: False  # the reference's own op sequence on either device
    saved = {n: getattr(ops, n) for n in ("log_mel_spectrogram", "qkv_attention", "qkv_attention_packed", "attention_pattern", "guided_loss",
                                          "head_vote", "layer_norm", "adapter_layer_norm")}
    try:
        ops.log_mel_spectrogram = log_mel_spectrogram
        ops.qkv_attention = qkv_attention
        ops.qkv_attention_packed = qkv_attention_packed
        ops.attention_pattern = attention_pattern
        ops.guided_loss = guided_loss
        ops.head_vote = head_vote
        ops.layer_norm = layer_norm
        ops.adapter_layer_norm = adapter_layer_norm
        yield
    finally:
        whisper_model._native = saved_native
        for n, f in saved.items():
            setattr(ops, n, f)
