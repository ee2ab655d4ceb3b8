"""torch-facing operators of the hot path; each one is a thin call into the C ABI (``_lib``).

PyTorch is used for device memory, streams and autograd bookkeeping only.  Tensors must live on a CUDA
device: there is no CPU implementation in this package (the CPU restatement lives in ``oracle/`` and is
test infrastructure).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

HOP_LENGTH = 160
N_FFT = 400
N_FREQ = 201


# Optional per-launch timing (bench.py): when set to a dict, C-ABI calls that launch the hot kernels are bracketed
# with CUDA events on the launching stream and appended as (tag, flops_or_bytes, start_event, end_event).
PROFILE = None


class _Timed:
    __slots__ = ("tag", "work", "e0", "e1")

    def __init__(self, tag, work, device):
        self.tag, self.work = tag, work
        self.e0 = self.e1 = None
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(device))

    def done(self, device):
        if self.e0 is not None:
            self.e1.record(torch.cuda.current_stream(device))
            PROFILE.setdefault(self.tag, []).append((self.work, self.e0, self.e1))


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise L.AgaError(f"{name} must be a CUDA tensor: aga_b200 has no CPU path")


# ------------------------------------------------------------------------------------------------
# a2. mel filterbank  (reference: whisper/audio.py:92-107 loads librosa.filters.mel from an npz)
# ------------------------------------------------------------------------------------------------
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank_numpy(n_mels: int = 80, sr: int = 16000, n_fft: int = N_FFT) -> np.ndarray:
    """Slaney-scale, area-normalised triangular filterbank == librosa.filters.mel(sr, n_fft, n_mels)."""
    n_freq = n_fft // 2 + 1
    fftfreqs = np.linspace(0.0, sr / 2.0, n_freq)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


_FILTER_CACHE = {}
_PACKED_CACHE = {}


def mel_filters(device, n_mels: int = 80) -> torch.Tensor:
    """Drop-in for whisper.audio.mel_filters(device, n_mels) (cached per device); also serves 128 bins."""
    key = (str(torch.device(device)), int(n_mels))
    if key not in _FILTER_CACHE:
        _FILTER_CACHE[key] = torch.from_numpy(mel_filterbank_numpy(n_mels)).to(device)
    return _FILTER_CACHE[key]


def _packed_filters(filters: torch.Tensor, algo: str) -> torch.Tensor:
    """The filterbank packed for one of the two frontends (cached per tensor): "simt" — the banded form of
    csrc/logmel.cu, any matrix; "tc" — mel bands + DFT tables of csrc/logmel_tc.cu, banded filterbanks only (the stock
    Slaney triangles): AgaError otherwise."""
    key = (filters.data_ptr(), filters._version, tuple(filters.shape), str(filters.device), algo)
    hit = _PACKED_CACHE.get(key)
    if hit is not None:
        return hit
    lib = L.lib()
    n_mels = filters.shape[0]
    nbytes = C.c_size_t()
    if algo == "tc":
        host = np.ascontiguousarray(filters.detach().cpu().numpy(), dtype=np.float32)  # once per filterbank
        hptr = C.c_void_p(host.ctypes.data)
        if lib.aga_logmel_filters_banded(hptr, n_mels) != 1:
            raise L.AgaError("the tensor-core frontend needs a banded filterbank (at most two adjacent filters per bin)")
        L.check(lib.aga_logmel_tc_packed_bytes(n_mels, C.byref(nbytes)), "aga_logmel_tc_packed_bytes")
        packed = torch.empty(nbytes.value, dtype=torch.uint8, device=filters.device)
        L.check(lib.aga_logmel_tc_pack(hptr, n_mels, _ptr(packed), nbytes.value, _stream_ptr(filters.device)),
                "aga_logmel_tc_pack")
    else:
        L.check(lib.aga_logmel_packed_filter_bytes(n_mels, C.byref(nbytes)), "aga_logmel_packed_filter_bytes")
        packed = torch.empty(nbytes.value, dtype=torch.uint8, device=filters.device)
        L.check(lib.aga_logmel_pack_filters(_ptr(filters), n_mels, _ptr(packed), nbytes.value,
                                            _stream_ptr(filters.device)), "aga_logmel_pack_filters")
    if len(_PACKED_CACHE) > 16:
        _PACKED_CACHE.clear()
    _PACKED_CACHE[key] = packed
    packed._aga_keepalive = filters
    return packed


# ------------------------------------------------------------------------------------------------
# a1. log-mel
# ------------------------------------------------------------------------------------------------
def log_mel_spectrogram(audio: torch.Tensor, ilens: Optional[torch.Tensor] = None, n_mels: int = 80,
                        filters: Optional[torch.Tensor] = None, valid_samples: Optional[torch.Tensor] = None,
                        algo: Optional[str] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """(B, N) fp32 -> ((B, n_mels, N//160) fp32, ilens // 160).

    Same contract as OpenAIWhisperEncoder.log_mel_spectrogram (espnet2/asr/encoder/whisper_encoder.py:105-135).
    ``valid_samples`` (device int32 scalar): the batch's true common length when ``audio`` has been zero-padded to a
    static bucket length (graphed.BucketedTrainStep) — frames past it come back as zeros.
    ``algo``: "simt" = the CUDA-core kernel (csrc/logmel.cu), "tc" = the tensor-core kernel (csrc/logmel_tc.cu; banded
    filterbanks only; the one that serves ``valid_samples``).  Default (None): the FASTER one as measured on B200 —
    the CUDA-core kernel (54 us vs 72 us at B = 16 x 30 s inside the training step: the tensor-core kernel's epilogue,
    not its GEMMs, bounds it — DESIGN.md §4); "tc" when ``valid_samples`` is given.
    """
    if algo is None:
        algo = "tc" if valid_samples is not None else "simt"
    _require_cuda(audio, "audio")
    if audio.dim() != 2:
        raise L.AgaError("audio must be (B, N)")
    if audio.dtype != torch.float32:
        audio = audio.float()
    if audio.stride(1) != 1:
        audio = audio.contiguous()
    B, N = audio.shape
    if N <= N_FFT // 2:
        raise L.AgaError("log_mel_spectrogram needs N > 200 samples (reflect padding)")
    if filters is None:
        filters = mel_filters(audio.device, n_mels)
    else:
        filters = filters.to(device=audio.device, dtype=torch.float32).contiguous()
        n_mels = filters.shape[0]
    if filters.shape != (n_mels, N_FREQ):
        raise L.AgaError(f"filters must be ({n_mels}, {N_FREQ})")
    if algo not in ("tc", "simt"):
        raise L.AgaError("algo must be 'tc', 'simt' or None")
    if valid_samples is not None and algo != "tc":
        raise L.AgaError("valid_samples is served by the tensor-core frontend (banded filterbanks) only")
    kind = algo
    packed = _packed_filters(filters, algo)
    F = N // HOP_LENGTH
    if valid_samples is not None:
        valid_samples = valid_samples.to(device=audio.device, dtype=torch.int32).reshape(())
    tm = _Timed("logmel", float(B) * (N * 4 + n_mels * F * 4), audio.device)
    out = L.torch_ops().logmel(audio, packed, n_mels, valid_samples, kind == "tc")
    tm.done(audio.device)
    olens = None if ilens is None else ilens // HOP_LENGTH
    return out, olens


# ------------------------------------------------------------------------------------------------
# a3. attention core
# ------------------------------------------------------------------------------------------------
_DTYPES = {torch.float32: L.AGA_F32, torch.bfloat16: L.AGA_BF16}
_IMPLS = {"auto": L.ATTN_AUTO, "simt": L.ATTN_SIMT, "tcgen05": L.ATTN_TCGEN05}
_KINDS = {None: L.EXPORT_NONE, "none": L.EXPORT_NONE, "logits": L.EXPORT_LOGITS, "probs": L.EXPORT_PROBS}


TC_BWD_AVAILABLE = True  # the tcgen05 backward exists (attn_tc.cu)


def _impl_name(q, causal, kind, impl, bwd=False, cols=(0, 0)) -> str:
    """Which implementation the C ABI will pick (mirrors attn_tc.cu::attn_tc_supported / attn_tc_bwd_supported) — for
    profiling tags only."""
    narrow = kind == L.EXPORT_NONE or (kind == L.EXPORT_LOGITS and cols[1] - cols[0] <= 16)
    tc = impl != L.ATTN_SIMT and q.dtype == torch.bfloat16 and narrow
    return "tc" if tc else "simt"


def _prep(t: torch.Tensor) -> torch.Tensor:
    vec = 8 if t.dtype == torch.bfloat16 else 4
    ok = t.stride(2) == 1 and t.stride(0) % vec == 0 and t.stride(1) % vec == 0 and t.data_ptr() % 16 == 0
    return t if ok else t.contiguous()


def _attn_forward(q, k, v, n_head, causal, kind, cols, head_sel, impl, kv_len=None, guided=None):
    """One aga_attn_fwd call on (possibly strided) q (B,Tq,D), k, v (B,Tk,D) -> (out, lse, export_buf or None, part or None)."""
    B, Tq, D = q.shape
    if D != n_head * 64:
        raise L.AgaError("head dim must be 64 (every Whisper size)")
    lo, hi = cols if kind != L.EXPORT_NONE else (0, 0)
    flops = 4.0 * B * n_head * Tq * k.shape[1] * 64 * (0.5 if causal else 1.0)
    tm = _Timed(f"attn_fwd_{_impl_name(q, causal, kind, impl, cols=cols)}_{Tq}x{k.shape[1]}", flops, q.device)
    g_pat, g_early = guided if guided is not None else (None, False)
    out, lse, export_buf, part = L.torch_ops().attn_fwd(q, k, v, n_head, causal, kind, lo, hi, head_sel, impl, kv_len, g_pat,
                                                        bool(g_early))
    tm.done(q.device)
    return out, lse, (export_buf if kind != L.EXPORT_NONE else None), (part if guided is not None else None)


def _attn_backward(q, k, v, out, lse, head_sel, probs, cfg, dout, dexport, dq, dk, dv, kv_len=None, guided=None, dpart=None):
    """One aga_attn_bwd call; dq / dk / dv are caller-allocated and share the strides of q / k / v."""
    n_head, causal, kind, cols, impl, has_sel = cfg
    if dout is None:
        dout = torch.zeros_like(out)
    dout = dout.to(out.dtype)
    if dout.stride() != out.stride():
        dout = dout.contiguous()
    if dexport is not None:
        dexport = dexport.float().contiguous()
    assert dq.stride() == q.stride() and dk.stride() == k.stride() and dv.stride() == v.stride()
    flops = 10.0 * q.shape[0] * n_head * q.shape[1] * k.shape[1] * 64 * (0.5 if causal else 1.0)
    tm = _Timed(f"attn_bwd_{_impl_name(q, causal, kind if dexport is not None else L.EXPORT_NONE, impl, bwd=True, cols=cols)}"
                f"_{q.shape[1]}x{k.shape[1]}", flops, q.device)
    lo, hi = cols if kind != L.EXPORT_NONE else (0, 0)
    g_pat, g_early = guided if (guided is not None and dpart is not None) else (None, False)
    L.torch_ops().attn_bwd(q, k, v, out, lse, dout, dexport, probs if probs.numel() else None, dq, dk, dv, n_head, causal, kind,
                           lo, hi, head_sel if has_sel else None, impl, kv_len, g_pat,
                           None if g_pat is None else dpart.float().contiguous(), bool(g_early))
    tm.done(q.device)


def _save(ctx, tensors, out, lse, head_sel, export_buf, n_head, causal, kind, cols, impl):
    ctx.save_for_backward(*tensors, out, lse, head_sel if head_sel is not None else torch.empty(0),
                          export_buf if (export_buf is not None and kind == L.EXPORT_PROBS) else torch.empty(0))
    ctx.cfg = (n_head, causal, kind, cols, impl, head_sel is not None)
    ctx.mark_non_differentiable(lse)


class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, n_head, causal, kind, cols, head_sel, impl, kv_len=None):
        for name, t in (("q", q), ("k", k), ("v", v)):
            _require_cuda(t, name)
        if q.dtype not in _DTYPES:
            raise L.AgaError(f"attention supports fp32 and bf16, got {q.dtype}")
        k = k.to(q.dtype)
        v = v.to(q.dtype)
        q, k, v = _prep(q), _prep(k), _prep(v)
        out, lse, export_buf, _ = _attn_forward(q, k, v, n_head, causal, kind, cols, head_sel, impl, kv_len)
        _save(ctx, (q, k, v), out, lse, head_sel, export_buf, n_head, causal, kind, cols, impl)
        ctx.kv_len = kv_len
        return out, lse, export_buf

    @staticmethod
    def backward(ctx, dout, _dlse, dexport):
        q, k, v, out, lse, head_sel, probs = ctx.saved_tensors
        # gradients share the strides of q / k / v (strided inputs were made dense by _prep or are dense slices' parents)
        # (with kv_len the kernel never visits key tiles past the true length: those dk / dv rows are zero)
        alloc = torch.empty_strided if ctx.kv_len is None else (lambda sh, st, dtype, device: torch.zeros(sh, dtype=dtype, device=device))
        dq, dk, dv = (alloc(t.shape, t.stride(), dtype=t.dtype, device=t.device) for t in (q, k, v))
        _attn_backward(q, k, v, out, lse, head_sel, probs, ctx.cfg, dout, dexport, dq, dk, dv, ctx.kv_len)
        return dq, dk, dv, None, None, None, None, None, None, None


class _AttnPackedFn(torch.autograd.Function):
    """Attention on ONE packed projection output.  ``n_q`` = 1: x is (B,T,3D) = [q | k | v] (self-attention after a
    single fused QKV GEMM); ``n_q`` = 0: x is (B,Tk,2D) = [k | v] and q comes separately (cross-attention).  The kernels
    read the column slices in place (they take token / batch strides) and the backward writes dq / dk / dv straight into
    one packed gradient, so the projection's backward is a single GEMM with no gradient-accumulation adds."""

    @staticmethod
    def forward(ctx, q, x, n_head, causal, kind, cols, head_sel, impl, kv_len=None, guided_pattern=None, guided_early=False):
        _require_cuda(x, "packed projection")
        if x.dtype not in _DTYPES:
            raise L.AgaError(f"attention supports fp32 and bf16, got {x.dtype}")
        if not x.is_contiguous():
            x = x.contiguous()
        D = n_head * 64
        if q is None:
            qv, kv, vv = x[..., :D], x[..., D:2 * D], x[..., 2 * D:]
        else:
            qv, kv, vv = _prep(q.to(x.dtype)), x[..., :D], x[..., D:]
        guided = None if guided_pattern is None else (guided_pattern, guided_early)
        out, lse, export_buf, part = _attn_forward(qv, kv, vv, n_head, causal, kind, cols, head_sel, impl, kv_len, guided)
        _save(ctx, (qv if q is not None else torch.empty(0), x), out, lse, head_sel, export_buf, n_head, causal, kind, cols, impl)
        ctx.packed_q = q is None
        ctx.kv_len = kv_len
        ctx.guided = guided
        return out, lse, export_buf, part

    @staticmethod
    def backward(ctx, dout, _dlse, dexport, dpart):
        qs, x, out, lse, head_sel, probs = ctx.saved_tensors
        D = ctx.cfg[0] * 64
        dx = torch.empty_like(x) if ctx.kv_len is None else torch.zeros_like(x)  # rows past kv_len are never visited
        if ctx.packed_q:
            q, k, v = x[..., :D], x[..., D:2 * D], x[..., 2 * D:]
            dq, dk, dv = dx[..., :D], dx[..., D:2 * D], dx[..., 2 * D:]
        else:
            q, k, v = qs, x[..., :D], x[..., D:]
            dq = torch.empty_strided(q.shape, q.stride(), dtype=q.dtype, device=q.device)
            dk, dv = dx[..., :D], dx[..., D:]
        _attn_backward(q, k, v, out, lse, head_sel, probs, ctx.cfg, dout, dexport, dq, dk, dv, ctx.kv_len, ctx.guided, dpart)
        return (None if ctx.packed_q else dq), dx, None, None, None, None, None, None, None, None, None


def qkv_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, n_head: int, causal: bool = False,
                  export: Optional[str] = None, export_cols: Optional[Tuple[int, int]] = None,
                  head_sel: Optional[torch.Tensor] = None, impl: str = "auto", kv_len: Optional[torch.Tensor] = None):
    """Fused MultiHeadAttention.qkv_attention (whisper/model.py:93-109).

    q (B,Tq,D), k,v (B,Tk,D), D = n_head*64.  Returns (out (B,Tq,D), lse (B,H,Tq), exported) where
    ``exported`` is None or the fp32 (B,H,Tq,hi-lo) side buffer holding key columns [lo,hi) of the
    scaled, masked logits (``export="logits"``, what the reference returns at HEAD) or of the softmax
    (``export="probs"``).  Differentiable in q, k, v through ``out`` and through ``exported``.
    ``kv_len`` (device int32 scalar, non-causal bf16 only): keys at or past it do not exist — a batch zero-padded to a
    static key length for CUDA-graph replay attends exactly like the unpadded one.
    """
    if kv_len is not None:
        kv_len = kv_len.to(device=q.device, dtype=torch.int32).reshape(())
    kind = _KINDS[export]
    if kind != L.EXPORT_NONE and export_cols is None:
        export_cols = (0, k.shape[1])
    if head_sel is not None:
        head_sel = head_sel.to(device=q.device, dtype=torch.uint8).contiguous()
    empty_cols = export_cols if export_cols is not None else (0, 0)
    return _AttnFn.apply(q, k, v, int(n_head), bool(causal), kind, tuple(empty_cols), head_sel, _IMPLS[impl], kv_len)


def qkv_attention_packed(x: torch.Tensor, n_head: int, q: Optional[torch.Tensor] = None, causal: bool = False,
                         export: Optional[str] = None, export_cols: Optional[Tuple[int, int]] = None,
                         head_sel: Optional[torch.Tensor] = None, impl: str = "auto", kv_len: Optional[torch.Tensor] = None,
                         guided: Optional[Tuple[torch.Tensor, bool]] = None):
    """``qkv_attention`` on a packed projection: x = [q | k | v] (B,T,3D), or x = [k | v] (B,Tk,2D) with ``q`` given.
    Same results as :func:`qkv_attention` on the three column slices; one packed gradient comes back.

    ``guided = (pattern (B,T,2), early)`` (decoder self attention: causal, bf16, T <= 128): the per-(utterance, head) part
    of the guided loss (espnet_model.py:496-512) is reduced inside the attention kernel's epilogue; the third return value
    is then ``GuidedParts`` (B,H,4,2) — per 32-row group [sum_t r_t, #{t: r_t != 0}] — instead of an exported slab, and
    nothing of size (T, T) or (T, 2) is written."""
    kind = _KINDS[export]
    Tk = x.shape[1]
    if kind != L.EXPORT_NONE and export_cols is None:
        export_cols = (0, Tk)
    if head_sel is not None:
        head_sel = head_sel.to(device=x.device, dtype=torch.uint8).contiguous()
    empty_cols = export_cols if export_cols is not None else (0, 0)
    if kv_len is not None:
        kv_len = kv_len.to(device=x.device, dtype=torch.int32).reshape(())
    if guided is not None:
        pat = guided[0].to(device=x.device, dtype=torch.float32).contiguous()
        out, lse, _, part = _AttnPackedFn.apply(q, x, int(n_head), bool(causal), L.EXPORT_NONE, (0, 0), None, _IMPLS[impl], kv_len,
                                                pat, bool(guided[1]))
        return out, lse, GuidedParts(part)
    out, lse, exported, _ = _AttnPackedFn.apply(q, x, int(n_head), bool(causal), kind, tuple(empty_cols), head_sel, _IMPLS[impl],
                                                kv_len)
    return out, lse, exported


# ------------------------------------------------------------------------------------------------
# 8f #2. LayerNorm with fp32 statistics (whisper/model.py:30-32)
# ------------------------------------------------------------------------------------------------
def _rows(t: torch.Tensor, D: int) -> torch.Tensor:
    t2 = t.reshape(-1, D)
    return t2 if (t2.is_contiguous() and t2.data_ptr() % 16 == 0) else t2.contiguous()


def _ln_fwd(x2, residual2, w32, b32, eps, want_sum):
    """x2 (rows, D) [+ residual2] -> (y, s, mean, rstd); s = x2 + residual2 in x2.dtype (or x2 itself)."""
    rows, D = x2.shape
    has_sum = residual2 is not None and want_sum
    tm = _Timed("layernorm_fwd", (2.0 + (residual2 is not None) + has_sum) * rows * D * x2.element_size(), x2.device)
    y, s, mean, rstd = L.torch_ops().layernorm_fwd(x2, residual2, w32, b32, float(eps), bool(want_sum))
    tm.done(x2.device)
    return y, s, mean, rstd


class ZeroArena:
    """One fp32 buffer per device that is cleared ONCE per training micro-step (``reset``: a single fill kernel) and
    handed out in slices to the kernels whose small outputs are accumulated with atomics — LayerNorm dgamma / dbeta /
    dxsum rows, the adapter's bias-gradient column sums.  Each of the ~100 such outputs per step otherwise costs its own
    memset node plus the dependency bubble behind it in the captured step (0.4 ms of a 25 ms step, tools/trace_step.py).
    A slice is handed out once per reset; without a preceding ``reset`` (plain eager use) or past the capacity ``take``
    returns None and the kernel's own clearing entry point is used."""

    _by_device = {}

    def __init__(self, device, n_floats: int = 1 << 20):
        self.buf = torch.zeros(n_floats, dtype=torch.float32, device=device)
        self.off = 0      # next free element of this pass
        self.high = 0     # everything at or past it has never been handed out and is still zero
        self.armed = False

    @classmethod
    def of(cls, device) -> "ZeroArena":
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        a = cls._by_device.get(key)
        if a is None:
            a = cls._by_device[key] = cls(device)
        return a

    def reserve(self, n_floats: int) -> None:
        """Make room for ``n_floats`` per pass (FlatGradBucket asks for the size of its gradient buffer: the adapter weight
        gradients live here until the bucket gathers them).  Only before any CUDA graph has captured slices of it."""
        if n_floats > self.buf.numel():
            self._retired = getattr(self, "_retired", []) + [self.buf]  # a graph captured earlier may still point into it
            self.buf = torch.zeros(int(n_floats), dtype=torch.float32, device=self.buf.device)
            self.off = self.high = 0
            self.armed = False

    def reset(self) -> None:
        """Start of a backward-producing pass (FlatGradBucket.begin_step): clear everything any earlier pass used — its
        consumers (the gradient bucket's copy, the optimizer) are done by now — and start handing out slices.  Captured
        into a CUDA graph this is one fill over the high-water mark, which covers what the graph itself uses because
        the warm-up passes before the capture used the same amount."""
        if self.high:
            self.buf[:self.high].zero_()
        self.off = 0
        self.armed = True

    def disarm(self) -> None:
        """End of the pass (FlatGradBucket.finish_backward): later backward passes that are not bracketed by a reset get
        None from ``take`` and clear their outputs themselves."""
        self.armed = False

    def take(self, n: int) -> Optional[torch.Tensor]:
        n_al = (n + 63) & ~63  # 256-byte granules
        if not self.armed or self.off + n_al > self.buf.numel():
            return None
        out = self.buf[self.off:self.off + n]
        self.off += n_al
        self.high = max(self.high, self.off)
        return out


def _ln_bwd(dy2, s2, w32, mean, rstd, need_params, need_dxsum=False, dres2=None):
    rows, D = s2.shape
    tm = _Timed("layernorm_bwd", 3.0 * rows * D * s2.element_size(), s2.device)
    pgz = None
    if need_params:
        pgz = ZeroArena.of(s2.device).take((3 if need_dxsum else 2) * D)
        pgz = None if pgz is None else pgz.view(3 if need_dxsum else 2, D)
    dx, pg = L.torch_ops().layernorm_bwd(dy2, s2, w32, mean, rstd, bool(need_params), bool(need_dxsum), dres2, pgz)
    tm.done(s2.device)
    dgamma = dbeta = dxsum = None
    if need_params:  # rows of one buffer: the library clears them with a single memset
        dgamma, dbeta = pg[0], pg[1]
        dxsum = pg[2] if need_dxsum else None
    return dx, dgamma, dbeta, dxsum


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        _require_cuda(x, "x")
        if x.dtype not in _DTYPES:
            raise L.AgaError(f"layer_norm supports fp32 and bf16 rows, got {x.dtype}")
        D = x.shape[-1]
        x2 = _rows(x, D)
        w32 = weight.detach().float().contiguous()
        b32 = bias.detach().float().contiguous()
        y, _, mean, rstd = _ln_fwd(x2, None, w32, b32, eps, False)
        ctx.save_for_backward(x2, w32, mean, rstd)
        ctx.shape = x.shape
        ctx.param_dtypes = (weight.dtype, bias.dtype)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, w32, mean, rstd = ctx.saved_tensors
        rows, D = x2.shape
        dy2 = _rows(dy.to(x2.dtype), D)
        need_params = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx, dgamma, dbeta, _ = _ln_bwd(dy2, x2, w32, mean, rstd, need_params)
        wd, bd = ctx.param_dtypes
        return (dx.view(ctx.shape), dgamma.to(wd) if ctx.needs_input_grad[1] else None,
                dbeta.to(bd) if ctx.needs_input_grad[2] else None, None)


class _LayerNormResidualFn(torch.autograd.Function):
    """(LN(x), x): the pre-LayerNorm of a residual branch together with the tensor the branch is later added to
    (`x = x + f(ln(x))`, whisper/model.py:231-242).  Owning both uses of x lets the backward add the gradient arriving
    through the residual connection inside the LayerNorm-backward kernel instead of a separate full-size add."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        _require_cuda(x, "x")
        if x.dtype not in _DTYPES:
            raise L.AgaError(f"layer_norm supports fp32 and bf16 rows, got {x.dtype}")
        D = x.shape[-1]
        x2 = _rows(x, D)
        w32 = weight.detach().float().contiguous()
        b32 = bias.detach().float().contiguous()
        y, _, mean, rstd = _ln_fwd(x2, None, w32, b32, eps, False)
        ctx.save_for_backward(x2, w32, mean, rstd)
        ctx.shape = x.shape
        ctx.param_dtypes = (weight.dtype, bias.dtype)
        return y.view(x.shape), x2.view(x.shape)

    @staticmethod
    def backward(ctx, dy, dres):
        x2, w32, mean, rstd = ctx.saved_tensors
        rows, D = x2.shape
        dy2 = _rows(dy.to(x2.dtype), D)
        dres2 = None if dres is None else _rows(dres.to(x2.dtype), D)
        need_params = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx, dgamma, dbeta, _ = _ln_bwd(dy2, x2, w32, mean, rstd, need_params, dres2=dres2)
        wd, bd = ctx.param_dtypes
        return (dx.view(ctx.shape), dgamma.to(wd) if ctx.needs_input_grad[1] else None,
                dbeta.to(bd) if ctx.needs_input_grad[2] else None, None)


def layer_norm_residual(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5):
    """``(layer_norm(x), x)``: use the second value as the residual the branch output is added to."""
    return _LayerNormResidualFn.apply(x, weight, bias, float(eps))


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """whisper.model.LayerNorm.forward (whisper/model.py:30-32): statistics in fp32, result in x.dtype, one kernel.
    Differentiable in x, weight and bias."""
    return _LayerNormFn.apply(x, weight, bias, float(eps))


def cast_trainable(p: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """``p.detach().to(dtype)``; served from the per-step low-precision shadow when the parameter has a fresh one
    (``parallel.FlatGradBucket.begin_step`` refreshes all of them with one multi-tensor copy)."""
    if p.dtype == dtype:
        return p.detach()
    sh = getattr(p, "_aga_shadow", None)
    if sh is not None and sh[0] == p._version and sh[1].dtype == dtype and sh[1].device == p.device:
        return sh[1]
    return p.detach().to(dtype)


def _wgrad(a_t: torch.Tensor, b: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """a_t @ b with the result in the PARAMETER's dtype straight out of the GEMM (fp32 accumulators are written as they
    are: no bf16 rounding of the weight gradient and no cast kernel)."""
    if dtype == a_t.dtype:
        return a_t @ b
    return torch.mm(a_t, b, out_dtype=dtype)


def adapter_wgrad(a: torch.Tensor, b: torch.Tensor, dtype: torch.dtype, transpose_out: bool = False) -> torch.Tensor:
    """``a^T @ b`` (or its transpose) in the PARAMETER's dtype: the weight gradient of an adapter Linear, a (rows, M) and
    b (rows, N) being activations / upstream gradients with rows = batch x frames.  bf16 operands of the adapter shapes
    go to the tcgen05 kernel of the library (csrc/wgrad_tc.cu: MN-major operands straight from the row-major tensors,
    K splits combined with atomics into a slice of the step's cleared arena) where it beats cuBLAS's split-K + reduce
    pair (measured, profiles/r2_wgrad.txt: long contractions, Whisper-small / medium adapter shapes: 16 vs 23 us);
    everything else is one GEMM whose fp32 accumulators are written as they are."""
    rows, M = a.shape
    N = b.shape[1]
    if (a.is_cuda and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and dtype == torch.float32 and rows >= 8192
            and M % 64 == 0 and N % 64 == 0 and max(M, N) <= 1024 and min(M, N) <= 256 and a.is_contiguous() and b.is_contiguous()
            and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0):
        out = ZeroArena.of(a.device).take(M * N)
        if out is None:
            out = torch.zeros(M * N, dtype=torch.float32, device=a.device)
        tm = _Timed("adapter_wgrad", 2.0 * rows * M * N, a.device)
        L.torch_ops().wgrad(a, b, out, bool(transpose_out))
        tm.done(a.device)
        return out.view(N, M) if transpose_out else out.view(M, N)
    return _wgrad(b.t(), a, dtype) if transpose_out else _wgrad(a.t(), b, dtype)


def gelu_bwd_colsum(dg: torch.Tensor, h: torch.Tensor):
    """(dg * gelu'(h), column sums of that product) in one pass — at::gelu_backward + sum(0) of the Adapter backward."""
    _require_cuda(dg, "dg")
    dg = dg if dg.is_contiguous() else dg.contiguous()
    return L.torch_ops().gelu_bwd_colsum(dg, h, ZeroArena.of(h.device).take(h.shape[1]))


class _AdapterLayerNormFn(torch.autograd.Function):
    """LN(x + W2 gelu(W1 x + b1) + b2): the Adapter (whisper/model.py:181-194) and the post-LayerNorm that replaces x
    (whisper/model.py:234-236, 244-246) as ONE autograd node.  The two small GEMMs stay cuBLAS; the residual add is
    folded into the LayerNorm kernel, the gradient of b2 falls out of the LayerNorm backward (column sums of dx), the
    GELU backward and the gradient of b1 are one kernel, the weight gradients leave their GEMMs in fp32, and
    dx = dx_ln + dh1 W1 is one addmm."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, gamma, beta, eps):
        _require_cuda(x, "x")
        if x.dtype not in _DTYPES:
            raise L.AgaError(f"adapter_layer_norm supports fp32 and bf16 rows, got {x.dtype}")
        D = x.shape[-1]
        x2 = _rows(x, D)
        dt = x2.dtype
        w1c, w2c = cast_trainable(w1, dt), cast_trainable(w2, dt)
        h1 = torch.addmm(cast_trainable(b1, dt), x2, w1c.t())
        g = torch.nn.functional.gelu(h1)
        y = torch.addmm(cast_trainable(b2, dt), g, w2c.t())
        g32 = gamma.detach().float().contiguous()
        be32 = beta.detach().float().contiguous()
        z, s, mean, rstd = _ln_fwd(x2, y, g32, be32, eps, True)
        ctx.save_for_backward(x2, h1, g, s, mean, rstd, w1c, w2c, g32)
        ctx.shape = x.shape
        ctx.param_dtypes = tuple(t.dtype for t in (w1, b1, w2, b2, gamma, beta))
        return z.view(x.shape)

    @staticmethod
    def backward(ctx, dz):
        x2, h1, g, s, mean, rstd, w1c, w2c, g32 = ctx.saved_tensors
        rows, D = x2.shape
        dts = ctx.param_dtypes
        dz2 = _rows(dz.to(x2.dtype), D)
        return _adapter_ln_backward(ctx, x2, h1, g, s, mean, rstd, w1c, w2c, g32, dz2)


class _AdapterLayerNormPairFn(torch.autograd.Function):
    """(z, y2) = (LN_a(x + adapter(x)), LN_b(z)): the adapter's post-LayerNorm, whose output replaces the residual stream,
    together with the FROZEN pre-LayerNorm of the residual branch that follows (whisper/model.py:234-246; across blocks:
    the next block's ``attn_ln`` / the encoder's ``ln_post``).  Forward: both normalisations in ONE kernel (the row is
    read once); backward: LN_b's backward with the residual-path gradient of z added in the same pass, then the adapter /
    LN_a backward of ``_AdapterLayerNormFn``.  Same values as the two nodes it replaces."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, gamma, beta, eps, gamma2, beta2, eps2):
        _require_cuda(x, "x")
        if x.dtype not in _DTYPES:
            raise L.AgaError(f"adapter_layer_norm supports fp32 and bf16 rows, got {x.dtype}")
        D = x.shape[-1]
        x2 = _rows(x, D)
        dt = x2.dtype
        w1c, w2c = cast_trainable(w1, dt), cast_trainable(w2, dt)
        h1 = torch.addmm(cast_trainable(b1, dt), x2, w1c.t())
        g = torch.nn.functional.gelu(h1)
        y = torch.addmm(cast_trainable(b2, dt), g, w2c.t())
        g32 = gamma.detach().float().contiguous()
        be32 = beta.detach().float().contiguous()
        g2_32 = gamma2.detach().float().contiguous()
        be2_32 = beta2.detach().float().contiguous()
        tm = _Timed("layernorm_fwd", 5.0 * x2.numel() * x2.element_size(), x2.device)
        z, s, mean, rstd, y2, mean2, rstd2 = L.torch_ops().layernorm_pair_fwd(x2, y, g32, be32, float(eps), True, g2_32, be2_32,
                                                                            float(eps2))
        tm.done(x2.device)
        ctx.save_for_backward(x2, h1, g, s, mean, rstd, w1c, w2c, g32, z, g2_32, mean2, rstd2)
        ctx.shape = x.shape
        ctx.param_dtypes = tuple(t.dtype for t in (w1, b1, w2, b2, gamma, beta))
        return z.view(x.shape), y2.view(x.shape)

    @staticmethod
    def backward(ctx, dz, dy2):
        x2, h1, g, s, mean, rstd, w1c, w2c, g32, z, g2_32, mean2, rstd2 = ctx.saved_tensors
        rows, D = x2.shape
        if dy2 is not None:  # d z = LN_b'(dy2) + (gradient that reached z through the residual connection)
            dres = None if dz is None else _rows(dz.to(x2.dtype), D)
            dz2, _, _, _ = _ln_bwd(_rows(dy2.to(x2.dtype), D), z, g2_32, mean2, rstd2, False, dres2=dres)
        else:
            dz2 = _rows(dz.to(x2.dtype), D)
        # (one kernel for both backwards — z recomputed from s, dz kept in shared memory — was built and measured: three
        # passes over the staged row with the recomputation made it ALU-bound, 79 us against 50 us for these two launches)
        return _adapter_ln_backward(ctx, x2, h1, g, s, mean, rstd, w1c, w2c, g32, dz2) + (None, None, None)


def _adapter_ln_backward(ctx, x2, h1, g, s, mean, rstd, w1c, w2c, g32, dz2):
    dts = ctx.param_dtypes
    ds, dgamma, dbeta, db2 = _ln_bwd(dz2, s, g32, mean, rstd, True, need_dxsum=True)
    dg = ds @ w2c                                  # (rows, bottleneck)
    dw2 = adapter_wgrad(ds, g, dts[2])                          # ds^T g: (D, bottleneck)
    dh1, db1 = gelu_bwd_colsum(dg, h1)
    dw1 = adapter_wgrad(x2, dh1, dts[0], transpose_out=True)   # (x^T dh1)^T = dh1^T x: (bottleneck, D)
    dx = _linear_residual_raw(dh1, w1c, None, ds, w_kn=True)  # ds + dh1 @ W1: the residual branch's gradient rides the GEMM
    return (dx.view(ctx.shape), dw1, db1.to(dts[1]), dw2, db2.to(dts[3]), dgamma.to(dts[4]), dbeta.to(dts[5]), None)


def adapter_layer_norm_pair(x, w1, b1, w2, b2, gamma, beta, eps, gamma2, beta2, eps2):
    """``z = ln_a(x + adapter(x))`` and ``ln_b(z)`` (ln_b frozen) as one node: see ``_AdapterLayerNormPairFn``.  Returns
    (z, ln_b(z)); use z as the residual stream and the second value as the next branch's input."""
    return _AdapterLayerNormPairFn.apply(x, w1, b1, w2, b2, gamma, beta, float(eps), gamma2, beta2, float(eps2))


def adapter_layer_norm(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
                       gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """``ln(x + adapter(x))`` of ResidualAttentionBlock.forward (whisper/model.py:234-236, 244-246) as one node;
    w1 (bottleneck, D), w2 (D, bottleneck) are the Adapter's ``model.0`` / ``model.2`` parameters."""
    return _AdapterLayerNormFn.apply(x, w1, b1, w2, b2, gamma, beta, float(eps))


# ------------------------------------------------------------------------------------------------
# 8f #2. Linear with the residual add folded into the GEMM
# ------------------------------------------------------------------------------------------------
_LR_WORKSPACE = {}


def _lr_workspace(device) -> torch.Tensor:
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _LR_WORKSPACE.get(key)
    if ws is None:
        n = C.c_size_t()
        L.check(L.lib().aga_linear_residual_workspace_bytes(C.byref(n)), "aga_linear_residual_workspace_bytes")
        ws = _LR_WORKSPACE[key] = torch.empty(n.value, dtype=torch.uint8, device=device)
    return ws


def _linear_residual_raw(x2: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], r2: torch.Tensor, w_kn: bool = False):
    """r2 + x2 @ W (+ b) on contiguous 2-D operands; W = w^T for an (N, K) weight, W = w for a (K, N) one (``w_kn``)."""
    # no fallback: AGA_ERR_UNSUPPORTED (cuBLASLt not loadable in the process, operands not 16-byte aligned) raises
    return L.torch_ops().linear_residual(x2, w, bool(w_kn), b, r2, _lr_workspace(x2.device))


class _LinearResidualFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, residual):
        N, K = w.shape
        x2 = x.reshape(-1, K)
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        r2 = residual.reshape(-1, N)
        r2 = r2 if r2.is_contiguous() else r2.contiguous()
        wc = w if w.is_contiguous() else w.contiguous()
        out = _linear_residual_raw(x2, wc, b, r2)
        ctx.save_for_backward(x2 if ctx.needs_input_grad[1] else torch.empty(0), wc)
        ctx.shapes = (x.shape, residual.shape)
        return out.view(residual.shape)

    @staticmethod
    def backward(ctx, dout):
        x2, wc = ctx.saved_tensors
        N, K = wc.shape
        d2 = dout.reshape(-1, N)
        dx = (d2 @ wc).view(ctx.shapes[0]) if ctx.needs_input_grad[0] else None
        dw = (d2.t() @ x2) if ctx.needs_input_grad[1] else None
        db = d2.sum(0) if ctx.needs_input_grad[2] else None
        return dx, dw, db, (dout if ctx.needs_input_grad[3] else None)


def linear_residual(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], residual: torch.Tensor) -> torch.Tensor:
    """``residual + F.linear(x, weight, bias)`` in one GEMM (cuBLASLt reads the residual as C with beta = 1 and applies the
    bias in the epilogue): the `x = x + self.attn(...)` / `x = x + self.mlp(...)` adds of ResidualAttentionBlock.forward
    (whisper/model.py:231-242) without their own pass over the activations.  x, weight, bias, residual share one dtype
    (fp32 or bf16)."""
    _require_cuda(x, "x")
    if not (x.dtype == weight.dtype == residual.dtype and (bias is None or bias.dtype == x.dtype)) or x.dtype not in _DTYPES:
        raise L.AgaError("linear_residual needs x, weight, bias and residual in one dtype (fp32 or bf16)")
    return _LinearResidualFn.apply(x, weight, bias, residual)


# ------------------------------------------------------------------------------------------------
# 8f #2. the block's MLP: GEMM with the exact GELU in its epilogue (tcgen05), second GEMM with the residual add
# ------------------------------------------------------------------------------------------------
def gemm_gelu_fwd(x2: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]):
    """(h, gelu(h)) with h = x2 @ w^T + b, bf16, in one tcgen05 GEMM (csrc/gemm_gelu.cu, mode 0)."""
    return L.torch_ops().gemm_gelu_fwd(x2, w, b)


def gemm_gelu_bwd(dy2: torch.Tensor, w_t: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """dh = (dy2 @ w_t^T) * gelu'(h): the second Linear's dgrad GEMM with the GELU backward in its epilogue (mode 1);
    ``w_t`` (N, K) is the transpose of that Linear's (K, N) weight."""
    return L.torch_ops().gemm_gelu_bwd(dy2, w_t, h)


class _MlpResidualFn(torch.autograd.Function):
    """residual + W2 gelu(W1 x + b1) + b2 for FROZEN bf16 weights: `x = x + self.mlp(self.mlp_ln(x))`
    (whisper/model.py:213,242) as two GEMMs and nothing else — the GELU lives in the first GEMM's epilogue, the residual
    add in the second's, the GELU backward in the epilogue of the second Linear's dgrad GEMM."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, w2_t, b2, residual):
        K = w1.shape[1]
        x2 = x.reshape(-1, K)
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        r2 = residual.reshape(-1, w2.shape[0])
        r2 = r2 if r2.is_contiguous() else r2.contiguous()
        h, g = gemm_gelu_fwd(x2, w1, b1)
        out = _linear_residual_raw(g, w2, b2, r2)
        ctx.save_for_backward(h, w1, w2_t)
        ctx.shapes = (x.shape, residual.shape)
        return out.view(residual.shape)

    @staticmethod
    def backward(ctx, dout):
        h, w1, w2_t = ctx.saved_tensors
        d2 = dout.reshape(-1, w2_t.shape[1])
        d2 = d2 if d2.is_contiguous() else d2.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dh = gemm_gelu_bwd(d2, w2_t, h)
            dx = (dh @ w1).view(ctx.shapes[0])
        return dx, None, None, None, None, None, (dout if ctx.needs_input_grad[6] else None)


def mlp_residual(x: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor], w2: torch.Tensor, w2_t: torch.Tensor,
                 b2: Optional[torch.Tensor], residual: torch.Tensor) -> torch.Tensor:
    """``residual + F.linear(gelu(F.linear(x, w1, b1)), w2, b2)`` for frozen bf16 weights (``w2_t`` = ``w2.t().contiguous()``,
    prepared once by the caller)."""
    _require_cuda(x, "x")
    for t in (x, w1, w2, w2_t, residual):
        if t.dtype != torch.bfloat16:
            raise L.AgaError("mlp_residual runs in bf16")
    return _MlpResidualFn.apply(x, w1, b1, w2, w2_t, b2, residual)


# ------------------------------------------------------------------------------------------------
# 8f #4. vocabulary logits -> label-smoothing KL loss + accuracy, without fp32 logits
# ------------------------------------------------------------------------------------------------
class VocabLogits:
    """The decoder's logits as the aligned GEMM left them: ``padded`` (..., ld) in the activation dtype, of which the
    first ``n_vocab`` columns are real (ld = n_vocab rounded up to 64).  ``materialize()`` gives the reference's tensor
    ``(x @ tok_emb.T).float()`` (whisper_decoder.py:164-166); ``ops.ls_cross_entropy`` consumes the handle directly."""

    def __init__(self, padded: torch.Tensor, n_vocab: int):
        self.padded, self.n_vocab = padded, int(n_vocab)

    def materialize(self) -> torch.Tensor:
        return self.padded[..., : self.n_vocab].float()

    @property
    def shape(self):
        return tuple(self.padded.shape[:-1]) + (self.n_vocab,)


class _LsCeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits2, target, n_vocab, padding_idx, smoothing, denom):
        _require_cuda(logits2, "logits")
        row_loss, row_lse, row_correct = L.torch_ops().ls_ce_fwd(logits2, target, n_vocab, padding_idx, float(smoothing))
        ctx.save_for_backward(logits2, target, row_lse, denom)
        ctx.cfg = (n_vocab, padding_idx, float(smoothing))
        ctx.mark_non_differentiable(row_correct)
        return row_loss.sum() / denom, row_correct

    @staticmethod
    def backward(ctx, g, _):
        logits2, target, row_lse, denom = ctx.saved_tensors
        n_vocab, padding_idx, smoothing = ctx.cfg
        gscale = (g.float() / denom).reshape(1).contiguous()  # upstream gradient / denominator, as a device scalar
        dlogits = L.torch_ops().ls_ce_bwd(logits2, target, n_vocab, padding_idx, smoothing, row_lse, gscale)
        return dlogits, None, None, None, None, None


def ls_cross_entropy(logits, target: torch.Tensor, padding_idx: int, smoothing: float, normalize_length: bool = False,
                     n_vocab: Optional[int] = None):
    """LabelSmoothingLoss.forward (label_smoothing_loss.py:41-63) + th_accuracy (nets_utils.py:304-324) in two kernels.

    ``logits``: a :class:`VocabLogits` handle or a plain (B, T, V) tensor (fp32 / bf16); ``target`` (B, T) int64 with
    ``padding_idx`` on ignored positions.  Returns (loss, accuracy): loss = sum of the per-token KL / (batch size, or
    the number of real tokens when ``normalize_length``)."""
    if isinstance(logits, VocabLogits):
        padded, n_vocab = logits.padded, logits.n_vocab
    else:
        padded = logits
        n_vocab = n_vocab or logits.shape[-1]
    if padded.dtype not in _DTYPES:
        padded = padded.float()
    vec = 16 // padded.element_size()
    if padded.shape[-1] % vec != 0:  # unaligned rows (a plain 51865-wide tensor): pad once
        padded = torch.nn.functional.pad(padded, (0, (-padded.shape[-1]) % vec))
    batch = padded.shape[0]
    logits2 = padded.reshape(-1, padded.shape[-1])
    if not logits2.is_contiguous():
        logits2 = logits2.contiguous()
    tgt = target.reshape(-1).to(torch.int64).contiguous()
    real = (tgt != padding_idx).sum()
    denom = real.float() if normalize_length else torch.full((), float(batch), dtype=torch.float32, device=logits2.device)
    loss, correct = _LsCeFn.apply(logits2, tgt, int(n_vocab), int(padding_idx), float(smoothing), denom)
    acc = correct.sum().float() / real.float()
    return loss, acc


# ------------------------------------------------------------------------------------------------
# a11 / a12 / a10
# ------------------------------------------------------------------------------------------------
def attention_pattern(tokens: torch.Tensor, lid_table: torch.Tensor, c: float = 0.6) -> torch.Tensor:
    """(B,T) int64 ys_in_pad -> (B,T,2) fp32 pattern (+inf on pad rows); espnet_model.py:236-275."""
    _require_cuda(tokens, "tokens")
    tokens = tokens.to(torch.int64).contiguous()
    B, T = tokens.shape
    lid_table = lid_table.to(device=tokens.device, dtype=torch.uint8).contiguous()
    return L.torch_ops().attention_pattern(tokens, lid_table, float(c))


class _GuidedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, slab, pattern, head_mask, n_early):
        _require_cuda(slab, "slab")
        if slab.dtype != torch.float32 or slab.dim() != 5 or slab.shape[-1] != 2 or slab.stride(-1) != 1:
            raise L.AgaError("slab must be fp32 (L,B,H,T,2) with unit last stride")
        Lyr, B, H, T, _ = slab.shape
        pattern = pattern.to(device=slab.device, dtype=torch.float32).contiguous()
        head_mask = head_mask.to(device=slab.device, dtype=torch.float32).contiguous()
        if pattern.shape != (B, T, 2) or head_mask.shape != (Lyr, H):
            raise L.AgaError("pattern must be (B,T,2) and head_mask (L,H)")
        need_grad = ctx.needs_input_grad[0]
        # a strided view (columns 1:3 of full maps) is gathered once: the slab is only L*B*H*T*2 floats,
        # and autograd scatters the dense gradient back through the view
        src = slab.contiguous()
        loss, d_slab = L.torch_ops().guided_loss(src, pattern, head_mask, int(n_early), bool(need_grad))
        d_slab = d_slab if need_grad else None
        ctx.d_slab = d_slab
        return loss

    @staticmethod
    def backward(ctx, g):
        d = ctx.d_slab
        return (None if d is None else d * g), None, None, None


class GuidedParts:
    """Per-(utterance, head, 32-row group) partial sums of the guided loss as the attention epilogue left them:
    ``t`` (..., B, H, 4, 2) = [sum_t r_t, #{t: r_t != 0}].  ``stack`` joins the layers; ``guided_loss_from_parts``
    finishes the loss."""

    def __init__(self, t: torch.Tensor):
        self.t = t

    @staticmethod
    def stack(parts):
        return GuidedParts(torch.stack([p.t for p in parts]))


def guided_loss_from_parts(parts: GuidedParts, head_mask: torch.Tensor) -> torch.Tensor:
    """The tail of calculate_cs_loss (espnet_model.py:509-530) on the fused partial sums (L,B,H,4,2): mean over the
    non-zero rows (0/0 = NaN as in the reference), selected-head mask, sum over (layer, head), mean over the batch."""
    t = parts.t
    m = t[..., 0].sum(-1) / t[..., 1].sum(-1)                      # (L,B,H)
    masked = head_mask.to(device=t.device, dtype=torch.float32)[:, None, :] * m
    return masked.sum(dim=(0, 2)).mean()


def guided_loss(slab: torch.Tensor, pattern: torch.Tensor, head_mask: torch.Tensor, n_early: int = 2) -> torch.Tensor:
    """Attention-guided loss on exported columns 1:3 — ESPnetASRModel.calculate_cs_loss (espnet_model.py:463-530)."""
    return _GuidedLossFn.apply(slab, pattern, head_mask, n_early)


def head_vote(probs: torch.Tensor, counts: Optional[torch.Tensor] = None):
    """probs (L,B,H,T,T) fp32 -> (decisions (L,B,H) uint8, counts (L,H) int32); espnet_model.py:285-310."""
    _require_cuda(probs, "probs")
    probs = probs.float().contiguous()
    Lyr, B, H, T, T2 = probs.shape
    assert T == T2
    if counts is None:
        counts = torch.zeros((Lyr, H), dtype=torch.int32, device=probs.device)
    dec = L.torch_ops().head_vote(probs, counts)
    return dec, counts
