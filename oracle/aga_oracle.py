"""CPU oracle for the AGA hot path (TEST INFRASTRUCTURE — not shipped, not measured as product).

A plain numpy restatement of the reference algorithms on the hot path
(SURVEY.md §8a).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module.  The product path (``aga_b200``) never does: it calls the CUDA library
through the C ABI and fails loudly when that library is missing.

Pinning status ("parity pinned by reference outputs"): the reference's own
tests hold no numeric golden vectors for this path, so the
oracle is pinned against outputs of the *reference code itself*, run in the
build container by ``oracle/make_golden.py`` (which imports
``/root/reference``) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every function below against those
fixtures, plus the one implicit golden the reference holds: the literal
72-head mask (espnet2/asr/espnet_model.py:514-525) == top-72 of
``attention_count_whispernoft_new.pkl`` (espnet_model.py:198-219).

All citations are relative to /root/reference/espnet unless noted:
  W/  = whisper/whisper/      E2/ = espnet2/
"""
from __future__ import annotations

import math
import string
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# ----------------------------------------------------------------------------
# Whisper audio constants — W/audio.py:13-23
# ----------------------------------------------------------------------------
SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
N_FREQ = N_FFT // 2 + 1  # 201

# Token layout the loss hard-codes — E2/text/whisper_token_id_converter.py:57-64,
# E2/bin/asr_inference.py:324 ; ids confirmed against W/tokenizer.py
TOK_EOT = 50257
TOK_SOT = 50258
TOK_EN = 50259
TOK_ZH = 50260
TOK_TRANSCRIBE = 50359
TOK_NOTIMESTAMPS = 50363
PROMPT_LEN = 5  # espnet2/asr/espnet_model.py:241

# LID classes of the (vocab,) lookup table that replaces the per-step tokenizer loop
LID_OTHER = 0    # Mandarin / anything not pure ASCII letters -> [c, 0]
LID_ENGLISH = 1  # ASCII letters only (after removing the space marker) -> [0, c]
LID_BOTH = 2     # token is only the space marker(s) -> [c, c]
LID_EOT = 3      # id 50257 -> [c, c] then stop


# ----------------------------------------------------------------------------
# a2. mel filterbank — W/audio.py:92-107 loads librosa.filters.mel(sr=16000,
#     n_fft=400, n_mels=80) from assets/mel_filters.npz.  librosa is a
#     third-party dependency absent from /root/reference; its published
#     algorithm (Slaney scale + Slaney area normalisation) is restated here and
#     pinned against the stored npz (tests/golden/mel_80_ref.npy).
# ----------------------------------------------------------------------------
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(n_mels: int = 80, sr: int = SAMPLE_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """(n_mels, n_fft//2+1) float32 Slaney mel filterbank."""
    n_freq = n_fft // 2 + 1
    fftfreqs = np.linspace(0.0, sr / 2.0, n_freq)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, n_freq), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


# ----------------------------------------------------------------------------
# a1. log-mel — E2/asr/encoder/whisper_encoder.py:105-135 (twin:
#     E2/asr/frontend/whisper.py:54-83).  torch.stft defaults: center=True,
#     pad_mode="reflect", onesided, unnormalised; periodic Hann window.
# ----------------------------------------------------------------------------
def hann_window(n: int = N_FFT) -> np.ndarray:
    """torch.hann_window(n) (periodic) — whisper_encoder.py:110."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def stft_power(audio: np.ndarray) -> np.ndarray:
    """|STFT|^2 with the last frame dropped — whisper_encoder.py:111-117.

    audio (B, N) -> (B, 201, N//160) float64.
    """
    audio = np.asarray(audio, dtype=np.float64)
    B, N = audio.shape
    pad = N_FFT // 2
    xp = np.pad(audio, ((0, 0), (pad, pad)), mode="reflect")
    n_frames = 1 + N // HOP_LENGTH
    keep = n_frames - 1  # stft[..., :-1]
    idx = (np.arange(keep) * HOP_LENGTH)[:, None] + np.arange(N_FFT)[None, :]
    frames = xp[:, idx] * hann_window()[None, None, :]  # (B, keep, 400)
    spec = np.fft.rfft(frames, n=N_FFT, axis=-1)  # (B, keep, 201)
    return (spec.real ** 2 + spec.imag ** 2).transpose(0, 2, 1)


def log_mel_spectrogram(
    audio: np.ndarray, ilens: Optional[np.ndarray] = None, n_mels: int = 80,
    filters: Optional[np.ndarray] = None,
) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """(B, N) -> ((B, n_mels, N//160) float64, olens) — whisper_encoder.py:105-135."""
    power = stft_power(audio)
    if filters is None:
        filters = mel_filterbank(n_mels)
    mel = np.einsum("mk,bkt->bmt", filters.astype(np.float64), power)  # :119-120
    log_spec = np.log10(np.maximum(mel, 1e-10))  # :122
    olens = None if ilens is None else np.asarray(ilens) // HOP_LENGTH  # :124-127
    mx = log_spec.reshape(log_spec.shape[0], -1).max(axis=-1)[:, None, None]  # :129-132
    log_spec = np.maximum(log_spec, mx - 8.0)
    log_spec = (log_spec + 4.0) / 4.0  # :133
    return log_spec, olens


# ----------------------------------------------------------------------------
# a3. attention core — W/model.py:93-109
# ----------------------------------------------------------------------------
def qkv_attention(
    q: np.ndarray, k: np.ndarray, v: np.ndarray, n_head: int, causal: bool = False,
    dtype=np.float64,
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """q (B,Tq,D), k,v (B,Tk,D) -> (out (B,Tq,D), qk (B,H,Tq,Tk), w (B,H,Tq,Tk)).

    ``qk`` is the scaled, causally masked pre-softmax logits the reference
    returns at HEAD (W/model.py:102-109); ``w`` is softmax(qk), the quantity the
    ``#modify here qk to w`` switch exports instead (W/model.py:108).
    """
    q = np.asarray(q, dtype=dtype); k = np.asarray(k, dtype=dtype); v = np.asarray(v, dtype=dtype)
    B, Tq, D = q.shape
    Tk = k.shape[1]
    d = D // n_head
    scale = d ** -0.25  # :96
    qh = q.reshape(B, Tq, n_head, d).transpose(0, 2, 1, 3) * scale  # :97
    kh = k.reshape(B, Tk, n_head, d).transpose(0, 2, 3, 1) * scale  # :98
    vh = v.reshape(B, Tk, n_head, d).transpose(0, 2, 1, 3)  # :99
    qk = qh @ kh  # :101
    if causal:  # mask = full(-inf).triu_(1), W/model.py:322 ; added at :103
        assert Tq == Tk
        mask = np.triu(np.full((Tq, Tk), -np.inf), 1)
        qk = qk + mask
    mx = qk.max(axis=-1, keepdims=True)
    e = np.exp(qk - mx)
    w = e / e.sum(axis=-1, keepdims=True)  # :106
    out = (w @ vh).transpose(0, 2, 1, 3).reshape(B, Tq, D)  # :109
    return out, qk, w


def qkv_attention_bwd(
    q, k, v, n_head: int, causal: bool, dout: np.ndarray,
    d_qk: Optional[np.ndarray] = None, d_w: Optional[np.ndarray] = None,
):
    """Analytic backward of qkv_attention (what autograd replays for W/model.py:96-109).

    dout (B,Tq,D); d_qk / d_w: optional upstream gradients (B,H,Tq,Tk) on the
    exported logits / probabilities.  Returns dq, dk, dv.
    """
    q = np.asarray(q, np.float64); k = np.asarray(k, np.float64); v = np.asarray(v, np.float64)
    B, Tq, D = q.shape
    Tk = k.shape[1]
    d = D // n_head
    scale = d ** -0.25
    _, qk, w = qkv_attention(q, k, v, n_head, causal)
    qh = q.reshape(B, Tq, n_head, d).transpose(0, 2, 1, 3) * scale
    kh = k.reshape(B, Tk, n_head, d).transpose(0, 2, 1, 3) * scale
    vh = v.reshape(B, Tk, n_head, d).transpose(0, 2, 1, 3)
    doh = np.asarray(dout, np.float64).reshape(B, Tq, n_head, d).transpose(0, 2, 1, 3)
    dv = w.transpose(0, 1, 3, 2) @ doh
    dw = doh @ vh.transpose(0, 1, 3, 2)
    if d_w is not None:
        dw = dw + d_w
    ds = w * (dw - (w * dw).sum(axis=-1, keepdims=True))
    if d_qk is not None:
        g = np.where(np.isfinite(qk), d_qk, 0.0)
        ds = ds + g
    dq = (ds @ kh) * scale
    dk = (ds.transpose(0, 1, 3, 2) @ qh) * scale
    back = lambda x, T: x.transpose(0, 2, 1, 3).reshape(B, T, D)
    return back(dq, Tq), back(dk, Tk), back(dv, Tk)


# ----------------------------------------------------------------------------
# a9. head-mask construction — E2/asr/espnet_model.py:186-219
# ----------------------------------------------------------------------------
def select_heads(attention_count: Dict[int, Dict[int, int]], head_percentage: float,
                 n_layers: int = 12, n_heads: int = 12, base: int = 110) -> np.ndarray:
    """{layer:{head:count}} (1-based) -> (n_layers, n_heads) float32 0/1 mask.

    Flatten in dict order (:204-207), *stable* sort by count descending (:210),
    take the first int(110*head_percentage/100) (:213-214), keep count>0 (:216).
    """
    freq = []
    for a, inner in attention_count.items():
        for b, c in inner.items():
            freq.append((a, b, c))
    srt = sorted(freq, key=lambda x: x[2], reverse=True)
    out = np.zeros((n_layers, n_heads), dtype=np.float32)
    n_sel = int(base * head_percentage / 100)
    for a, b, c in srt[:n_sel]:
        if c > 0:
            out[a - 1][b - 1] = 1
    return out


# The literal mask the loss actually uses — E2/asr/espnet_model.py:514-525.
# Stored as one bit-string per layer (derived, and equal to select_heads(pkl, 72/110*100)).
LITERAL_HEAD_MASK_ROWS = [
    "000000000000", "000000000000", "000000000000", "011101100111", "001111011111",
    "111111011101", "111111111110", "011111111111", "100110101010", "111100100010",
    "111110101001", "010010000001",
]


def literal_head_mask() -> np.ndarray:
    return np.array([[float(c) for c in row] for row in LITERAL_HEAD_MASK_ROWS], dtype=np.float32)


# ----------------------------------------------------------------------------
# a10. head selection — E2/asr/espnet_model.py:285-310
# ----------------------------------------------------------------------------
def check_attention_language(maps: np.ndarray, k: int = 2) -> np.ndarray:
    """The OLDER head vote, E2/asr/espnet_model.py:312-363: maps (L,B,H,T,T) -> int64 (L,H) count increments.

    Per utterance and (layer, head): every row is arg-sorted in descending order and its first ``k`` = 2 key indices
    are kept (:325-337); the indices are counted over all rows (:340-343); the ``k`` most frequent indices — equal counts
    resolved towards the SMALLER index, because ``torch.unique`` lists them ascending and Python's ``sorted`` is stable
    (:340-349) — must contain both 1 and 2 (<|zh|>, <|en|>) for the head to be selected (:352-357).
    Ties INSIDE a row (e.g. the zeros above the diagonal of a causal probability map, which decide the second index
    of row 0) are resolved towards the smaller index here (a stable sort); ``torch.argsort`` leaves them unspecified.
    """
    maps = np.asarray(maps)
    L, B, H, T, _ = maps.shape
    counts = np.zeros((L, H), dtype=np.int64)
    for b in range(B):
        for l in range(L):
            for h in range(H):
                order = np.argsort(-maps[l, b, h], axis=-1, kind="stable")[:, :k]
                idx, cnt = np.unique(order.reshape(-1), return_counts=True)
                top = [int(i) for i, _ in sorted(zip(idx, cnt), key=lambda x: x[1], reverse=True)[:k]]
                if 1 in top and 2 in top:
                    counts[l, h] += 1
    return counts



def new_check_attention_language(maps: np.ndarray) -> np.ndarray:
    """maps (L,B,H,T,T) *probabilities* -> int64 (L,H) count increments.

    Per utterance and (layer, head): sum_1 = sum over rows of cols 1:3,
    sum_2 = sum of col 0 + sum of cols 3: ; selected iff sum_1 > sum_2
    (:297-299).  Accumulated in float32 in the reference's order (Python
    ``sum`` over rows of row-vectors, then over columns).
    """
    maps = np.asarray(maps, dtype=np.float32)
    L, B, H, T, _ = maps.shape
    counts = np.zeros((L, H), dtype=np.int64)
    for b in range(B):
        for l in range(L):
            for h in range(H):
                a = maps[l, b, h]
                # sum(sum(x)) : inner sum adds the rows (vector adds), outer sums the columns
                col = np.zeros(2, dtype=np.float32)
                for t in range(T):
                    col = col + a[t, 1:3]
                s1 = np.float32(0)
                for c in col:
                    s1 = np.float32(s1 + c)
                s2a = np.float32(0)
                for t in range(T):
                    s2a = np.float32(s2a + a[t, 0])
                if T > 3:
                    colr = np.zeros(T - 3, dtype=np.float32)
                    for t in range(T):
                        colr = colr + a[t, 3:]
                    s2b = np.float32(0)
                    for c in colr:
                        s2b = np.float32(s2b + c)
                else:
                    s2b = np.float32(0)
                s2 = np.float32(s2a + s2b)
                if s1 > s2:
                    counts[l, h] += 1
    return counts


def head_vote_sums(maps: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """float64 (sum_1, sum_2) per (L,B,H) — for margin checks of the decisions."""
    maps = np.asarray(maps, dtype=np.float64)
    s1 = maps[..., 1:3].sum(axis=(-1, -2))
    s2 = maps[..., 0].sum(axis=-1) + maps[..., 3:].sum(axis=(-1, -2))
    return s1, s2


# ----------------------------------------------------------------------------
# a11. attention pattern — E2/asr/espnet_model.py:234-275
# ----------------------------------------------------------------------------
def bytes_to_unicode() -> Dict[int, str]:
    """GPT-2 byte -> printable unicode map (what HF convert_ids_to_tokens shows)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def lid_class_of_token_string(token: str) -> int:
    """Classify one HF-style token string exactly as espnet_model.py:246-258 does."""
    if token == "<|endoftext|>":  # :247
        return LID_EOT
    stripped = token.replace("Ġ", "")  # 'Ġ' ; :251,254
    if stripped == "":
        return LID_BOTH
    if all(ch in string.ascii_letters for ch in stripped):  # :234-235
        return LID_ENGLISH
    return LID_OTHER


def lid_class_of_token_bytes(raw: bytes) -> int:
    """Same classification from the raw BPE bytes (space byte 0x20 <-> 'Ġ')."""
    b2u = bytes_to_unicode()
    return lid_class_of_token_string("".join(b2u[x] for x in raw))


def create_attention_pattern(tokens: Sequence[int], lid_table: np.ndarray, c: float = 0.6) -> np.ndarray:
    """(T,) token ids -> (T,2) float32 target, inf on pad rows — espnet_model.py:236-275."""
    T = len(tokens)
    rows: List[List[float]] = [[0.0, 0.0], [c, 0.0], [0.0, c], [0.0, 0.0], [0.0, 0.0]]  # :261-265
    n_lid = 0
    for tok in list(tokens)[PROMPT_LEN:]:  # :246
        cls = int(lid_table[int(tok)])
        if cls == LID_EOT:
            rows.append([c, c]); n_lid += 1
            break
        elif cls == LID_BOTH:
            rows.append([c, c])
        elif cls == LID_ENGLISH:
            rows.append([0.0, c])
        else:
            rows.append([c, 0.0])
        n_lid += 1
    pad = T - PROMPT_LEN - n_lid  # :267
    rows.extend([[np.inf, np.inf]] * pad)
    return np.asarray(rows, dtype=np.float32).reshape(-1, 2)


# ----------------------------------------------------------------------------
# a12. guided ("cs") loss — E2/asr/espnet_model.py:463-530
# ----------------------------------------------------------------------------
def calculate_cs_loss(
    slab: np.ndarray, pattern: np.ndarray, head_mask: np.ndarray, n_early: int = 2,
    want_grad: bool = False, dtype=np.float64,
):
    """Guided loss on the compact slab.

    slab    (L,B,H,T,2) = maps[..., 1:3] (logits with -inf above the diagonal,
            or probabilities) — the only columns the reference reads (:506)
    pattern (B,T,2) from create_attention_pattern, inf on pad rows
    head_mask (L,H) 0/1 (the reference uses the literal at :514-525)
    n_early  layers [0,n_early) use the "early" pattern whose cols 1:3 are all
            zero and — unlike later layers — whose pad rows are NOT zeroed in
            the maps (:479-481,496)
    Returns loss (and d loss / d slab when want_grad).
    """
    A = np.array(slab, dtype=dtype)  # (L,B,H,T,2)
    L, B, H, T, _ = A.shape
    P = np.asarray(pattern, dtype=dtype)  # (B,T,2)
    pad = np.isinf(P)  # (B,T,2)
    tgt = np.zeros((L, B, 1, T, 2), dtype=dtype)
    late = np.zeros((L, 1, 1, 1, 1), dtype=bool)
    late[n_early:] = True
    Pz = np.where(pad, 0.0, P)
    tgt = np.where(late, Pz[None, :, None], 0.0)  # (L,B,1,T,2)
    zero_a = (late & pad[None, :, None]) | np.isinf(A)  # :496-497
    Az = np.where(zero_a, 0.0, A)
    diff = Az - tgt
    e = diff * diff  # MSELoss(reduction='none') :501,506
    r = e.sum(axis=-1)  # :509  (L,B,H,T)
    cnt = np.count_nonzero(r, axis=-1)  # :512
    with np.errstate(invalid="ignore", divide="ignore"):
        m = r.sum(axis=-1) / cnt  # (L,B,H)
    masked = np.asarray(head_mask, dtype=dtype)[:, None, :] * m  # :527
    loss = masked.sum(axis=(0, 2)).mean()  # :529
    if not want_grad:
        return loss
    with np.errstate(invalid="ignore", divide="ignore"):
        g = (np.asarray(head_mask, dtype=dtype)[:, None, :, None, None] * 2.0 * diff
             / (B * cnt[..., None, None]))
    g = np.where(zero_a, 0.0, g)
    return loss, g


# ----------------------------------------------------------------------------
# a13. label-smoothing loss + accuracy — espnet/nets/pytorch_backend/transformer/
#      label_smoothing_loss.py:41-63 ; nets_utils.py:304-324 ; add_sos_eos.py:12-31
# ----------------------------------------------------------------------------
def layer_norm(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """whisper/whisper/model.py:30-32 (LayerNorm.forward = F.layer_norm(x.float()).type(x.dtype)), in float64."""
    x = np.asarray(x, dtype=np.float64)
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)  # biased, as torch
    return (x - mu) / np.sqrt(var + eps) * np.asarray(gamma, np.float64) + np.asarray(beta, np.float64)


def layer_norm_bwd(dy: np.ndarray, x: np.ndarray, gamma: np.ndarray, eps: float = 1e-5):
    """Gradient of layer_norm w.r.t. (x, gamma, beta) — what autograd derives for whisper/model.py:30-32."""
    x = np.asarray(x, np.float64)
    dy = np.asarray(dy, np.float64)
    D = x.shape[-1]
    mu = x.mean(axis=-1, keepdims=True)
    rstd = 1.0 / np.sqrt(((x - mu) ** 2).mean(axis=-1, keepdims=True) + eps)
    xh = (x - mu) * rstd
    g = dy * np.asarray(gamma, np.float64)
    dx = rstd * (g - g.mean(axis=-1, keepdims=True) - xh * (g * xh).mean(axis=-1, keepdims=True))
    return dx, (dy * xh).reshape(-1, D).sum(0), dy.reshape(-1, D).sum(0)


def add_sos_eos(ys_pad: np.ndarray, sos: int, eos: int, ignore_id: int):
    ys = [y[y != ignore_id] for y in ys_pad]
    T = max(len(y) for y in ys) + 1
    ys_in = np.full((len(ys), T), eos, dtype=np.int64)
    ys_out = np.full((len(ys), T), ignore_id, dtype=np.int64)
    for i, y in enumerate(ys):
        ys_in[i, 0] = sos
        ys_in[i, 1 : 1 + len(y)] = y
        ys_out[i, : len(y)] = y
        ys_out[i, len(y)] = eos
    return ys_in, ys_out


def label_smoothing_loss(logits: np.ndarray, target: np.ndarray, smoothing: float = 0.1,
                         padding_idx: int = -1, normalize_length: bool = False) -> float:
    x = np.asarray(logits, np.float64)
    B, T, V = x.shape
    x = x.reshape(-1, V)
    t = np.asarray(target).reshape(-1)
    ignore = t == padding_idx
    total = len(t) - int(ignore.sum())
    tt = np.where(ignore, 0, t)
    true = np.full_like(x, smoothing / (V - 1))
    true[np.arange(len(tt)), tt] = 1.0 - smoothing
    lse = np.log(np.exp(x - x.max(1, keepdims=True)).sum(1, keepdims=True)) + x.max(1, keepdims=True)
    logp = x - lse
    with np.errstate(divide="ignore", invalid="ignore"):
        kl = np.where(true > 0, true * (np.log(true) - logp), 0.0)  # nn.KLDivLoss: 0 * log 0 = 0
    kl[ignore] = 0.0
    return float(kl.sum() / (total if normalize_length else B))
