"""CPU checks of the tensor-core log-mel frontend's HOST side (csrc/logmel_tc.cu): the band analysis of the filterbank and
the DFT tables, by emulating the kernel's arithmetic in numpy — folded frame, fp16 (hi, lo) operand pairs, three products
per GEMM, fp32 accumulation, the (Ce, Co, Se, So) -> |X_k|^2, |X_{200-k}|^2 combination and the streamed banded mel
projection — against the oracle's float64 log-mel (espnet2/asr/encoder/whisper_encoder.py:105-135)."""
import ctypes as C

import numpy as np
import pytest

import aga_oracle as O


@pytest.fixture(scope="module")
def lib():
    import aga_b200
    return aga_b200._lib.lib()


def _build(lib, fb):
    n = C.c_size_t()
    assert lib.aga_logmel_tc_packed_bytes(fb.shape[0], C.byref(n)) == 0
    buf = np.zeros(n.value, dtype=np.uint8)
    st = lib.aga_logmel_tc_build_host(C.c_void_p(fb.ctypes.data), fb.shape[0], C.c_void_p(buf.ctypes.data), n.value)
    return st, buf


HDR = 32
STREAM = 2 * 112 * 8 + 2 * 7 * 2 * 4


def _streams(buf):
    """-> wt (2, 112, 2) fp32, mask (2, 7, 2) uint32: the epilogue's stream tables."""
    o = HDR + 201 * 16
    wt = buf[o:o + 2 * 112 * 8].view(np.float32).reshape(2, 112, 2)
    mask = buf[o + 2 * 112 * 8:o + STREAM].view(np.uint32).reshape(2, 7, 2)
    return wt, mask


def _stream_projection(P, hdr, wt, mask, n_mels):
    """The kernel's epilogue in numpy: two streams with a rotating pair of running sums, merged where they meet."""
    F = P.shape[0]
    mel = np.zeros((F, n_mels), dtype=np.float32)
    written = np.zeros(n_mels, dtype=int)

    def emit(m, acc):
        if 0 <= m < n_mels:
            mel[:, m] = acc
            written[m] += 1

    final = []
    for role in (0, 1):
        accA = np.zeros(F, np.float32)
        accB = np.zeros(F, np.float32)
        m_id, step = int(hdr[4 + role]), (-1 if role else 1)
        for pos in range(112):
            blk, i = divmod(pos, 16)
            if (mask[role, blk, 0] >> i) & 1:
                emit(m_id, accA); accA, accB = accB, np.zeros(F, np.float32); m_id += step
                if (mask[role, blk, 1] >> i) & 1:
                    emit(m_id, accA); accA = accB; m_id += step
            pw = P[:, pos] if role == 0 else (P[:, 200 - pos] if pos <= 99 else np.zeros(F, np.float32))
            if role == 0 and pos > 100:
                pw = np.zeros(F, np.float32)
            accA = accA + wt[role, pos, 0] * pw
            accB = accB + wt[role, pos, 1] * pw
        final.append((accA, accB))
    L100, L101 = int(hdr[2]), int(hdr[3])
    (aA, aB), (dA, dB) = final
    u0, u1 = dB, dA
    d = L101 - L100
    if d == 0:
        emit(L100, aA + u0); emit(L100 + 1, aB + u1)
    elif d == 1:
        emit(L100, aA); emit(L100 + 1, aB + u0); emit(L100 + 2, u1)
    else:
        emit(L100, aA); emit(L100 + 1, aB); emit(L100 + 2, u0); emit(L100 + 3, u1)
    assert np.all(written == 1), written  # every filter is written exactly once
    return mel


def _tables(buf):
    """-> hdr (8 int32), bins (201, 4) fp32 view, T[table 0..3][hi/lo] as (112 bins, 112 K) float32."""
    hdr = buf[:HDR].view(np.int32)
    bins = buf[HDR:HDR + 201 * 16].view(np.float32).reshape(201, 4)
    off = (HDR + 201 * 16 + STREAM + 255) // 256 * 256
    raw = buf[off:off + 7 * 8 * 3584].view(np.float16).reshape(7, 4, 2, 14, 2, 8, 8)  # [j][tb][hl][k/8][kk/8][k%8][kk%8]
    T = np.zeros((4, 2, 112, 112), dtype=np.float32)
    for j in range(7):
        blk = raw[j].transpose(0, 1, 2, 4, 3, 5).reshape(4, 2, 112, 16)  # [tb][hl][k][kk]
        T[:, :, :, 16 * j:16 * j + 16] = blk.astype(np.float32)
    return hdr, bins, T


def _split(v):
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def _emulate(audio, bins, T, n_mels, buf=None):
    """One utterance, kernel arithmetic in numpy (fp32 accumulation)."""
    N = audio.shape[0]
    F = N // 160
    xp = np.pad(audio.astype(np.float32), 200, mode="reflect")
    frames = np.stack([xp[160 * f:160 * f + 400] for f in range(F)])  # (F, 400)
    m = np.abs(frames).max()
    e = int(np.floor(np.log2(m))) if m > 0 else 12
    S = np.float32(2.0 ** (12 - e))
    x = frames * S
    K = np.arange(100)
    ne, no = 2 * (K + 1), 2 * K + 1
    seqs = [x[:, ne] + x[:, 400 - ne], x[:, no] + x[:, 400 - no], x[:, ne] - x[:, 400 - ne],
            (x[:, no] - x[:, 400 - no]) * np.where(K % 2 == 1, -1.0, 1.0).astype(np.float32)]
    acc = []
    for sq, tb in zip(seqs, range(4)):
        a = np.zeros((F, 112), dtype=np.float32)
        a[:, :100] = sq
        ahi, alo = _split(a)
        thi, tlo = T[tb, 0], T[tb, 1]
        acc.append((ahi @ thi.T + ahi @ tlo.T + alo @ thi.T).astype(np.float32))
    ce, co, se, so = acc
    P = np.zeros((F, 201), dtype=np.float32)
    k = np.arange(101)
    P[:, k] = (ce[:, :101] + co[:, :101]) ** 2 + (se[:, :101] + so[:, :101]) ** 2
    kk = np.arange(100)
    P[:, 200 - kk] = (ce[:, kk] - co[:, kk]) ** 2 + (se[:, kk] - so[:, kk]) ** 2
    P *= np.float32(1.0) / (S * S)
    if buf is not None:  # the epilogue's two streams
        wt, mask = _streams(buf)
        mel = _stream_projection(P, buf[:HDR].view(np.int32), wt, mask, n_mels)
    else:  # banded projection: bin k adds w(L) to filter L[k] and w(L+1) to filter L[k] + 1
        mel = np.zeros((F, n_mels), dtype=np.float32)
        L = bins[:, 2].view(np.int32)
        for b in range(201):
            if 0 <= L[b] < n_mels:
                mel[:, L[b]] += bins[b, 0] * P[:, b]
            if 0 <= L[b] + 1 < n_mels:
                mel[:, L[b] + 1] += bins[b, 1] * P[:, b]
    lg = np.log10(np.maximum(mel, 1e-10)).T
    lg = np.maximum(lg, lg.max() - 8.0)
    return (lg + 4.0) / 4.0


@pytest.mark.parametrize("n_mels", [80, 128])
def test_band_analysis_reproduces_the_filterbank(lib, n_mels):
    fb = O.mel_filterbank(n_mels)
    assert lib.aga_logmel_filters_banded(C.c_void_p(fb.ctypes.data), n_mels) == 1
    st, buf = _build(lib, fb)
    assert st == 0
    hdr, bins, _ = _tables(buf)
    L = bins[:, 2].view(np.int32)
    assert np.all(np.diff(L) >= 0) and L.min() >= -1 and L.max() <= n_mels - 1
    assert hdr[1] == n_mels and hdr[2] == L[100] and hdr[3] == L[101]
    dense = np.zeros_like(fb)
    for b in range(201):
        if 0 <= L[b] < n_mels:
            dense[L[b], b] += bins[b, 0]
        if L[b] + 1 < n_mels:
            dense[L[b] + 1, b] += bins[b, 1]
    assert np.array_equal(dense, fb)  # every non-zero of the filterbank, bit for bit, nothing else


def test_dense_filterbank_is_not_banded(lib):
    rng = np.random.default_rng(0)
    fb = np.abs(rng.standard_normal((40, 201))).astype(np.float32)
    assert lib.aga_logmel_filters_banded(C.c_void_p(fb.ctypes.data), 40) == 0
    st, _ = _build(lib, fb)
    assert st == -2  # AGA_ERR_UNSUPPORTED


@pytest.mark.parametrize("kind", ["noise", "tonal", "loud"])
def test_folded_split_precision_dft_matches_oracle(lib, kind):
    rng = np.random.default_rng(3)
    N = 4000
    t = np.arange(N) / 16000.0
    if kind == "noise":
        audio = 0.1 * rng.standard_normal(N)
    elif kind == "tonal":  # strong tone + weak tone 70 dB down: dynamic range inside one frame
        audio = 0.5 * np.sin(2 * np.pi * 440 * t) + 1.5e-4 * np.sin(2 * np.pi * 3000 * t)
    else:  # int16-scale waveform: the per-tile power-of-two scale keeps the fp16 operands finite
        audio = 20000.0 * rng.standard_normal(N)
    audio = audio.astype(np.float32)
    fb = O.mel_filterbank(80)
    st, buf = _build(lib, fb)
    assert st == 0
    _, bins, T = _tables(buf)
    got = _emulate(audio, bins, T, 80, buf)
    ref, _ = O.log_mel_spectrogram(audio[None])
    if kind == "tonal":
        strong = ref[0] > ref[0].max() - 1.0
        np.testing.assert_allclose(got[strong], ref[0][strong], rtol=1e-4, atol=1e-5)
        assert np.abs(got - ref[0]).max() < 3e-4  # bins at the max-8 clamp sit at any fp32 transform's noise floor
    else:
        np.testing.assert_allclose(got, ref[0], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("n_mels", [80, 128])
def test_stream_tables_reproduce_the_projection(lib, n_mels):
    """The two-stream form of the mel projection (what the epilogue warps execute) == filterbank @ power, every filter
    written exactly once, for both stock banks."""
    fb = O.mel_filterbank(n_mels)
    st, buf = _build(lib, fb)
    assert st == 0
    rng = np.random.default_rng(n_mels)
    P = rng.random((5, 201)).astype(np.float32)
    wt, mask = _streams(buf)
    mel = _stream_projection(P, buf[:HDR].view(np.int32), wt, mask, n_mels)
    np.testing.assert_allclose(mel, P @ fb.T, rtol=2e-6, atol=1e-9)
