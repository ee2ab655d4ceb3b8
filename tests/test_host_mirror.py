"""CPU checks of the host-side mirror: state_dict compatibility with the reference module tree, head-mask
construction, and the plain-PyTorch loss plumbing against the oracle."""
import json
import os

import numpy as np
import pytest
import torch

import aga_oracle as O


@pytest.fixture(scope="module")
def pkg():
    import aga_b200  # noqa: F401
    from aga_b200 import espnet_model, espnet_whisper, whisper_model
    return whisper_model, espnet_whisper, espnet_model


def _register_e2e_dims(W):
    W.MODEL_DIMS["e2e-12x12"] = W.ModelDimensions(80, 1500, 768, 12, 2, 51865, 448, 768, 12, 12)


def test_state_dict_keys_match_reference(pkg, golden_dir):
    W, EW, _ = pkg
    _register_e2e_dims(W)
    meta = json.load(open(os.path.join(golden_dir, "e2e_small_meta.json")))
    enc = EW.OpenAIWhisperEncoder(whisper_model="e2e-12x12", adapter=True)
    dec = EW.OpenAIWhisperDecoder(51865, 768, whisper_model="e2e-12x12", adapter=True, whisper_cs=True, src_layer=1)
    mine = {k: list(v.shape) for k, v in list(enc.state_dict().items()) + list(dec.state_dict().items())}
    assert mine == meta["state_dict"]  # same keys, same shapes, same order: checkpoints are interchangeable


def test_adapter_param_counts(pkg):
    W, _, _ = pkg
    m = W.load_model("tiny", adapter=True)
    n = sum(p.numel() for k, p in m.named_parameters() if "adapter" in k)
    assert n == 1199616
    D, L = 768, 12
    assert (D * D + 6.5 * D) * 2 * L == 14275584


def test_load_model_needs_checkpoint_or_opt_in(pkg, monkeypatch, tmp_path):
    """a registry swap with `whisper_model: small` must not silently train a random Whisper.  Without a
    checkpoint file load_model raises (as the reference does when its download fails); seeded weights need an opt-in."""
    W, EW, _ = pkg
    monkeypatch.setenv("AGA_ALLOW_RANDOM_INIT", "0")
    W.allow_random_init(False)
    with pytest.raises(RuntimeError, match="no Whisper checkpoint"):
        W.load_model("tiny", download_root=str(tmp_path))
    with pytest.raises(RuntimeError, match="no Whisper checkpoint"):
        EW.OpenAIWhisperEncoder(whisper_model="tiny", download_dir=str(tmp_path))
    # a checkpoint in OpenAI format under download_root is found and loaded (adapter run: strict=False)
    W.allow_random_init(True)
    ref = W.load_model("tiny", download_root=str(tmp_path), seed=3)
    W.allow_random_init(False)
    torch.save({"dims": ref.dims.__dict__, "model_state_dict": ref.state_dict()}, tmp_path / "tiny.pt")
    got = W.load_model("tiny", adapter=True, download_root=str(tmp_path))
    assert torch.equal(got.encoder.conv1.weight, ref.encoder.conv1.weight)
    assert torch.equal(got.decoder.token_embedding.weight, ref.decoder.token_embedding.weight)


def test_select_heads_and_literal_mask(pkg, golden_dir):
    _, _, EM = pkg
    meta = json.load(open(os.path.join(golden_dir, "meta.json")))
    sel = EM.select_heads(EM._load_attention_count(meta["attention_count"]), 72 / 110 * 100 + 1e-9, 12, 12)
    assert torch.equal(sel, EM.literal_head_mask())
    assert np.array_equal(EM.literal_head_mask().numpy(), O.literal_head_mask())
    full = EM.select_heads(EM._load_attention_count(meta["attention_count"]), 100.0, 12, 12)
    assert int(full.sum()) == 110
    # ties keep (layer, head) insertion order: python's sort is stable
    tie = {1: {1: 5, 2: 5}, 2: {1: 5, 2: 0}}
    got = EM.select_heads(tie, 100.0 * 2 / 110 + 1e-9, 2, 2)
    assert got.tolist() == [[1.0, 1.0], [0.0, 0.0]]


def test_label_smoothing_and_sos_eos_vs_oracle(pkg):
    _, _, EM = pkg
    rng = np.random.default_rng(0)
    B, T, V = 3, 7, 50
    logits = rng.standard_normal((B, T, V)).astype(np.float32) * 3
    ys = np.full((B, T - 1), -1, dtype=np.int64)
    for b, n in enumerate([6, 3, 1]):
        ys[b, :n] = rng.integers(0, V - 2, n)
    yin, yout = EM.add_sos_eos(torch.from_numpy(ys), V - 2, V - 1, -1)
    yin_r, yout_r = O.add_sos_eos(ys, V - 2, V - 1, -1)
    assert np.array_equal(yin.numpy(), yin_r) and np.array_equal(yout.numpy(), yout_r)
    crit = EM.LabelSmoothingLoss(V, -1, 0.1, False)
    got = crit(torch.from_numpy(logits), yout).item()
    np.testing.assert_allclose(got, O.label_smoothing_loss(logits, yout_r, 0.1, -1, False), rtol=1e-5)
    crit0 = EM.LabelSmoothingLoss(V, -1, 0.0, True)
    np.testing.assert_allclose(crit0(torch.from_numpy(logits), yout).item(),
                               O.label_smoothing_loss(logits, yout_r, 0.0, -1, True), rtol=1e-5)


def test_whisper_frontend_constructor_and_keys(pkg):
    """WhisperFrontend (espnet2/asr/frontend/whisper.py:17-52): keywords, output_size, the `whisper.*` state_dict layout
    of a stock (adapter-free) Whisper, frozen-weights flag; unknown model names are rejected like the reference does."""
    W, EW, _ = pkg
    fe = EW.WhisperFrontend(whisper_model="tiny", freeze_weights=True, download_dir=None)
    assert fe.output_size() == 384 and fe.n_mels == 80 and (fe.n_fft, fe.hop_length, fe.win_length) == (400, 160, 400)
    keys = list(fe.state_dict().keys())
    assert keys[0] == "whisper.encoder.positional_embedding" and "whisper.decoder.token_embedding.weight" in keys
    assert not any("adapter" in k for k in keys)
    assert "whisper.encoder.blocks.3.mlp.2.bias" in keys and "whisper.decoder.blocks.3.cross_attn.key.weight" in keys
    assert not fe.whisper.training
    with pytest.raises(AssertionError):
        EW.WhisperFrontend("no-such-model")


def test_specaug_and_kv_cache_flags_do_not_change_state_dict(pkg):
    W, EW, _ = pkg
    a = EW.OpenAIWhisperDecoder(51865, 384, whisper_model="tiny", adapter=True, whisper_cs=True, src_layer=1)
    b = EW.OpenAIWhisperDecoder(51865, 384, whisper_model="tiny", adapter=True, whisper_cs=True, src_layer=1, kv_cache=True,
                                fused_loss=True)
    assert list(a.state_dict().keys()) == list(b.state_dict().keys())
    e = EW.OpenAIWhisperEncoder(whisper_model="tiny", adapter=True, use_specaug=True,
                                specaug_conf=dict(apply_time_warp=True, time_warp_window=5, time_warp_mode="bicubic",
                                                  apply_freq_mask=True, freq_mask_width_range=[0, 30], num_freq_mask=2,
                                                  apply_time_mask=True, time_mask_width_range=[0, 40], num_time_mask=2))
    assert not any("specaug" in k for k in e.state_dict().keys())
