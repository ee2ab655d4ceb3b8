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
