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
