"""GPU parity: log-mel frontend (C ABI -> CUDA) vs the oracle, the reference goldens and size-independent properties."""
import os

import numpy as np
import pytest
import torch

import aga_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    import aga_b200
    return aga_b200


def _run(A, audio, ilens=None, **kw):
    y, ol = A.log_mel_spectrogram(torch.from_numpy(np.ascontiguousarray(audio)).cuda(),
                                  None if ilens is None else torch.from_numpy(ilens).cuda(), **kw)
    torch.cuda.synchronize()
    return y.cpu().numpy(), None if ol is None else ol.cpu().numpy()


ALGOS = ["tc", "simt"]  # tensor-core frontend (csrc/logmel_tc.cu, the default for the stock banks) | general kernel


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("case", ["noise", "short"])
def test_logmel_vs_reference_golden(A, golden_dir, case, algo):
    g = np.load(os.path.join(golden_dir, "logmel.npz"))
    y, ol = _run(A, g[f"{case}_audio"], g[f"{case}_ilens"], algo=algo)
    assert np.array_equal(ol, g[f"{case}_olens"])
    # north-star tolerance: 1e-4 relative in fp32 (atol floor: outputs cross zero)
    np.testing.assert_allclose(y, g[f"{case}_logmel"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(y, g[f"{case}_logmel_f64"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("case", ["chirp", "chirp_padded"])
def test_logmel_tonal_vs_f64(A, golden_dir, case, algo):
    """Tonal input: bins near the max-8 clamp sit at the fp32 noise floor of ANY fp32 STFT; the reference's own
    fp32 run is 8e-5 abs from the float64 evaluation there, so the gate is 'no worse than 3x the reference'."""
    g = np.load(os.path.join(golden_dir, "logmel.npz"))
    y, _ = _run(A, g[f"{case}_audio"], g[f"{case}_ilens"], algo=algo)
    f64 = g[f"{case}_logmel_f64"]
    ref_err = np.abs(g[f"{case}_logmel"] - f64)
    err = np.abs(y - f64)
    assert err.max() <= max(3 * ref_err.max(), 1e-4), (err.max(), ref_err.max())
    assert np.median(err) <= max(3 * np.median(ref_err), 1e-6)
    # bins well above the floor obey the strict tolerance
    strong = f64 > (f64.max() - 1.0)
    np.testing.assert_allclose(y[strong], f64[strong], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("B,N", [(1, 201), (1, 479), (3, 4000), (2, 41760), (2, 48000 + 37), (1, 128 * 160), (2, 129 * 160 + 1)])
def test_logmel_vs_oracle_shapes(A, B, N, algo):
    """Ragged / minimal sizes: N=201 is the smallest torch.stft accepts; 41760 is the bundled SEAME utterance;
    N not a multiple of 160 and batch rows that are not 16-byte aligned (odd N)."""
    rng = np.random.default_rng(N)
    audio = (0.1 * rng.standard_normal((B, N))).astype(np.float32)
    y, _ = _run(A, audio, algo=algo)
    ref, _ = O.log_mel_spectrogram(audio)
    assert y.shape == ref.shape == (B, 80, N // 160)
    if ref.size:
        np.testing.assert_allclose(y, ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("algo", ALGOS)
def test_logmel_zero_and_silence_tail(A, algo):
    """All-zero input hits the clamp floor: log10(1e-10) = -10 -> (−10+4)/4 = −1.5 everywhere."""
    audio = np.zeros((2, 3200), dtype=np.float32)
    y, _ = _run(A, audio, algo=algo)
    assert np.all(y == -1.5)
    rng = np.random.default_rng(0)
    audio[0, :1600] = 0.1 * rng.standard_normal(1600)
    y, _ = _run(A, audio, algo=algo)
    ref, _ = O.log_mel_spectrogram(audio)
    np.testing.assert_allclose(y, ref, rtol=1e-4, atol=1e-5)


def test_logmel_128_bins_and_custom_filters(A):
    rng = np.random.default_rng(1)
    audio = (0.1 * rng.standard_normal((2, 8000))).astype(np.float32)
    ref, _ = O.log_mel_spectrogram(audio, n_mels=128)
    for algo in ALGOS:
        y, _ = _run(A, audio, n_mels=128, algo=algo)
        assert y.shape == (2, 128, 50)
        np.testing.assert_allclose(y, ref, rtol=1e-4, atol=1e-5)
    # a dense random (non-banded) filterbank exercises the generic packed path
    fb = np.abs(rng.standard_normal((40, 201))).astype(np.float32)
    y2, _ = A.log_mel_spectrogram(torch.from_numpy(audio).cuda(), filters=torch.from_numpy(fb).cuda())
    ref2, _ = O.log_mel_spectrogram(audio, filters=fb)
    np.testing.assert_allclose(y2.cpu().numpy(), ref2, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("algo", ALGOS)
def test_logmel_full_size_properties(A, algo):
    """BASELINE size (B=16, 30 s): batch independence, per-utterance max == 1 - ... dynamic range <= 2.0
    (whisper/tests/test_audio.py:19), and strided rows."""
    import functools
    g = torch.Generator().manual_seed(2022)
    audio = (0.1 * torch.randn(16, 480000, generator=g)).clamp(-1, 1).cuda()
    real = A.log_mel_spectrogram

    class _A:  # the same checks on the chosen kernel
        log_mel_spectrogram = staticmethod(functools.partial(real, algo=algo))
    A = _A
    y, _ = A.log_mel_spectrogram(audio)
    assert y.shape == (16, 80, 3000)
    assert float((y.amax(dim=(1, 2)) - y.amin(dim=(1, 2))).max()) <= 2.0 + 1e-6
    # each utterance is independent of its batch neighbours and of the row stride
    wide = torch.zeros(4, 480000 + 64, device="cuda")
    wide[:, :480000] = audio[3:7]
    y2, _ = A.log_mel_spectrogram(wide[:, :480000])
    assert torch.equal(y2, y[3:7])
    ref, _ = O.log_mel_spectrogram(audio[5:6, :].cpu().numpy())
    np.testing.assert_allclose(y[5:6].cpu().numpy(), ref, rtol=1e-4, atol=1e-5)
    # scaling the waveform by 10 shifts the un-clamped log-mel by exactly log10(100)/4 = 0.5
    y3, _ = A.log_mel_spectrogram(audio[:2] * 10.0)
    d = (y3 - y[:2])
    assert float((d - 0.5).abs().max()) < 2e-4


def test_logmel_loud_and_quiet_waveforms_tc(A):
    """The tensor-core frontend scales every tile by a power of two before the fp16 split: int16-range and very quiet
    waveforms, and a tile that mixes a loud and a quiet second, keep the fp32-transform accuracy."""
    rng = np.random.default_rng(5)
    base = rng.standard_normal((3, 32000)).astype(np.float32)
    audio = base.copy()
    audio[0] *= 20000.0
    audio[1] *= 3e-6
    audio[2, :16000] *= 0.5
    audio[2, 16000:] *= 2e-3
    y, _ = _run(A, audio, algo="tc")
    ref, _ = O.log_mel_spectrogram(audio)
    np.testing.assert_allclose(y, ref, rtol=1e-4, atol=2e-5)


def test_logmel_valid_samples_equals_unpadded(A):
    """Static launch shape for a shorter batch (graphed.BucketedTrainStep): audio zero-padded to the bucket length with the
    true common length as a DEVICE scalar — identical frames below it (reflect padding at the true end), exact zeros past."""
    rng = np.random.default_rng(6)
    n_true, n_bucket = 30000 + 77, 48000
    audio = (0.1 * rng.standard_normal((2, n_true))).astype(np.float32)
    y, _ = _run(A, audio, algo="tc")
    padded = np.zeros((2, n_bucket), dtype=np.float32)
    padded[:, :n_true] = audio
    yp, _ = A.log_mel_spectrogram(torch.from_numpy(padded).cuda(), valid_samples=torch.tensor(n_true, dtype=torch.int32).cuda())
    yp = yp.cpu().numpy()
    Fv = n_true // 160
    assert yp.shape == (2, 80, n_bucket // 160)
    assert np.array_equal(yp[:, :, :Fv], y)
    assert not yp[:, :, Fv:].any()
