"""BASELINE configs[0]: greedy decode of the reference's bundled SEAME utterance (decoded PCM fixture) through the
drop-in modules in fp32, against the reference modules' own run (tests/golden/decode_seame.npz, same name-seeded
weights): identical token ids, log-probabilities and dumped self-attention maps."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_greedy_decode_matches_reference(golden_dir):
    import aga_b200  # noqa: F401
    from aga_b200 import espnet_whisper as EW

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(golden_dir, "decode_seame.npz"))
    enc = EW.OpenAIWhisperEncoder(whisper_model="small", adapter=True).cuda().eval()
    dec = EW.OpenAIWhisperDecoder(51865, 768, whisper_model="small", adapter=True, whisper_cs=True, src_layer=1).cuda().eval()
    speech = torch.from_numpy(g["pcm"].astype(np.float32) / 32768.0)[None].cuda()
    assert speech.shape == (1, 41760)
    with torch.no_grad():
        enc_out, enc_lens, _ = enc(speech, torch.tensor([41760], device="cuda"))
        assert enc_out.shape == (1, 131, 768) and int(enc_lens[0]) == int(g["enc_out_lens"][0]) == 131
        np.testing.assert_allclose(enc_out[0, :8, :16].cpu().numpy(), g["enc_out_slice"], rtol=2e-3, atol=2e-4)
        ys = torch.tensor([[50258, 50260, 50259, 50359, 50363]], device="cuda")
        ids, logps = [], []
        for step in range(len(g["token_ids"])):
            last = step == len(g["token_ids"]) - 1
            logp, _ = dec.forward_one_step(ys, torch.empty(0), enc_out, return_maps=last)
            nxt = int(logp[0].argmax())
            ids.append(nxt)
            logps.append(float(logp[0, nxt]))
            ys = torch.cat([ys, torch.tensor([[nxt]], device="cuda")], dim=1)
    assert ids == g["token_ids"].tolist()  # margins of the reference run are >= 0.04, far above fp32 noise
    np.testing.assert_allclose(np.array(logps), g["logp"], rtol=1e-3, atol=1e-3)
    maps = torch.stack([m[0] for m in dec.att_map]).cpu().numpy()  # (12, H, t, t) logits, -inf above the diagonal
    ref = g["last_maps"]
    assert maps.shape == ref.shape
    assert np.array_equal(np.isinf(maps), np.isinf(ref))
    fin = np.isfinite(ref)
    np.testing.assert_allclose(maps[fin], ref[fin], rtol=2e-3, atol=2e-3)


def test_kv_cached_decode_matches_recompute_and_reference(golden_dir):
    """incremental decoding (self K/V appended, cross K/V projected once) gives the reference's greedy
    hypothesis on the bundled utterance and the recompute path's log-probabilities; batch_score / score carry the
    per-hypothesis states ESPnet's beam search expects."""
    import aga_b200  # noqa: F401
    from aga_b200 import espnet_whisper as EW

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(golden_dir, "decode_seame.npz"))
    enc = EW.OpenAIWhisperEncoder(whisper_model="small", adapter=True).cuda().eval()
    dec = EW.OpenAIWhisperDecoder(51865, 768, whisper_model="small", adapter=True, whisper_cs=True, src_layer=1,
                                  kv_cache=True).cuda().eval()
    speech = torch.from_numpy(g["pcm"].astype(np.float32) / 32768.0)[None].cuda()
    n_steps = len(g["token_ids"])
    with torch.no_grad():
        enc_out, _, _ = enc(speech, torch.tensor([41760], device="cuda"))
        ys = torch.tensor([[50258, 50260, 50259, 50359, 50363]], device="cuda")
        ids, logps, cache = [], [], None
        for step in range(n_steps):
            logp, cache = dec.forward_one_step(ys, torch.empty(0), enc_out, cache=cache)
            assert cache[0][0].shape == (1, ys.size(1), 768) and cache[0][2].shape == (1, 131, 768)
            nxt = int(logp[0].argmax())
            ids.append(nxt)
            logps.append(float(logp[0, nxt]))
            ys = torch.cat([ys, torch.tensor([[nxt]], device="cuda")], dim=1)
        assert ids == g["token_ids"].tolist()
        np.testing.assert_allclose(np.array(logps), g["logp"], rtol=1e-3, atol=1e-3)
        # the same prefix through the recompute path (kv_cache off): same distribution over the whole vocabulary
        dec.kv_cache = False
        ref_logp, none_state = dec.forward_one_step(ys[:, :-1], torch.empty(0), enc_out)
        assert none_state is None
        dec.kv_cache = True
        torch.testing.assert_close(logp, ref_logp, rtol=1e-3, atol=1e-3)
        # ESPnet scorer interface: two hypotheses of equal length, states stacked / split per hypothesis
        h1, h2 = ys[0, :7], torch.cat([ys[0, :6], torch.tensor([11], device="cuda")])
        s1 = s2 = None
        for t in range(5, 8):
            lp1, s1 = dec.score(h1[:t], s1, enc_out[0])
            lp2, s2 = dec.score(h2[:t], s2, enc_out[0])
        lpb, sb = dec.batch_score(torch.stack([h1, h2]), [None, None], enc_out.expand(2, -1, -1))
        torch.testing.assert_close(lpb, torch.stack([lp1, lp2]), rtol=1e-3, atol=1e-3)
        nxt = torch.stack([torch.cat([h1, lp1.argmax()[None]]), torch.cat([h2, lp2.argmax()[None]])])
        lpc, sc = dec.batch_score(nxt, sb, enc_out.expand(2, -1, -1))  # one incremental step from the batched states
        lpd, _ = dec.batch_score(nxt, [None, None], enc_out.expand(2, -1, -1))
        torch.testing.assert_close(lpc, lpd, rtol=1e-3, atol=1e-3)
        assert len(sc) == 2 and sc[0][0][0].shape == (8, 768)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_graphed_greedy_decode_matches_forward_one_step(golden_dir, dtype):
    """GraphedGreedyDecoder: one captured CUDA graph per token on static shapes (preallocated K / V, position and key
    count as device scalars).  fp32: the reference's greedy hypothesis and log-probabilities on the bundled utterance
    (decode_seame.npz); bf16: the same tokens as the KV-cached forward_one_step in bf16."""
    import aga_b200  # noqa: F401
    from aga_b200 import espnet_whisper as EW

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(golden_dir, "decode_seame.npz"))
    enc = EW.OpenAIWhisperEncoder(whisper_model="small", adapter=True).cuda().eval()
    dec = EW.OpenAIWhisperDecoder(51865, 768, whisper_model="small", adapter=True, whisper_cs=True, src_layer=1,
                                  kv_cache=True).cuda().eval()
    speech = torch.from_numpy(g["pcm"].astype(np.float32) / 32768.0)[None].cuda()
    n_tok = len(g["token_ids"])
    prompt = torch.tensor([[50258, 50260, 50259, 50359, 50363]], device="cuda")
    with torch.no_grad():
        enc_out, _, _ = enc(speech, torch.tensor([41760], device="cuda"))
        enc_out = enc_out.to(dtype)
        gd = dec.greedy_decoder(enc_out, max_len=64)
        gd.prefill(prompt)
        ids, logp = gd.decode(n_tok)
        # a second call continues where the first stopped, replaying the same graph
        ids2, _ = gd.decode(3)
        assert ids2.shape == (1, n_tok + 3) and torch.equal(ids2[:, :n_tok], ids)
        # the Python-driven KV-cached path in the same dtype
        ys, cache, ref_ids, ref_logp = prompt, None, [], []
        for _ in range(n_tok + 3):
            lp, cache = dec.forward_one_step(ys, torch.empty(0), enc_out, cache=cache)
            nxt = lp.argmax(-1, keepdim=True)
            ref_ids.append(int(nxt))
            ref_logp.append(float(lp[0, int(nxt)]))
            ys = torch.cat([ys, nxt], dim=1)
    assert ids2[0].tolist() == ref_ids
    if dtype == torch.float32:
        assert ids[0].tolist() == g["token_ids"].tolist()
        np.testing.assert_allclose(logp[0].cpu().numpy(), g["logp"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(logp[0].cpu().numpy(), np.array(ref_logp[:n_tok]), rtol=2e-3, atol=2e-3)
