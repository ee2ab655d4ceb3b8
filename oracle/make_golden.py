#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE CODE (build container only).

TEST INFRASTRUCTURE.  Imports the unmodified reference modules from
/root/reference with the three shims of SURVEY.md §8c (typeguard no-op, no
pretrained weights, offline tokenizer built from the reference's own
``whisper/assets/multilingual.tiktoken``) and stores small input/output
vectors.  /root/reference does not exist on the GPU box, so the committed
fixtures are what travels.  Run:  python oracle/make_golden.py
"""
import base64
import json
import os
import pickle
import sys
import types

import numpy as np

REF = "/root/reference"
sys.path[:0] = [f"{REF}/espnet", f"{REF}/espnet/whisper"]
import typeguard  # noqa: E402

typeguard.check_argument_types = lambda *a, **k: True
typeguard.check_return_type = lambda *a, **k: True

import torch  # noqa: E402
import whisper  # noqa: E402
from whisper.model import MultiHeadAttention  # noqa: E402
from espnet2.asr.encoder.whisper_encoder import OpenAIWhisperEncoder  # noqa: E402
from espnet2.asr.espnet_model import ESPnetASRModel  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import aga_oracle as O  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
PKG_DATA = os.path.join(HERE, "..", "attention-guided-adaptation-for-code-switching-speech-recognition_b200", "data")
os.makedirs(OUT, exist_ok=True)
os.makedirs(PKG_DATA, exist_ok=True)
torch.set_num_threads(4)


# ----------------------------------------------------------------------------- tokenizer shim
def load_tiktoken_ranks(path):
    ranks = {}
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            tok, rank = line.split()
            ranks[int(rank)] = base64.b64decode(tok)
    return ranks


class OfflineTokenizer:
    """convert_ids_to_tokens in HF style from the reference's own BPE asset."""

    def __init__(self):
        self.ranks = load_tiktoken_ranks(f"{REF}/espnet/whisper/whisper/assets/multilingual.tiktoken")
        self.b2u = O.bytes_to_unicode()
        self.n_base = len(self.ranks)
        assert self.n_base == 50257, self.n_base

    def token_string(self, i):
        i = int(i)
        if i < self.n_base:
            return "".join(self.b2u[b] for b in self.ranks[i])
        if i == 50257:
            return "<|endoftext|>"
        return f"<|special_{i}|>"

    def convert_ids_to_tokens(self, ids):
        return [self.token_string(i) for i in (ids.tolist() if hasattr(ids, "tolist") else ids)]


TOK = OfflineTokenizer()


def make_lid_table(vocab=51865):
    tab = np.zeros(vocab, dtype=np.uint8)
    for i in range(vocab):
        tab[i] = O.lid_class_of_token_string(TOK.token_string(i))
    return tab


# ----------------------------------------------------------------------------- model shims
def bare_asr_model():
    m = object.__new__(ESPnetASRModel)
    torch.nn.Module.__init__(m)
    m.tokenizer = TOK
    m.attention_count = {l: {h: 0 for h in range(1, 13)} for l in range(1, 13)}
    return m


def fake_encoder_self():
    s = types.SimpleNamespace()
    s.n_fft, s.win_length, s.hop_length, s.n_mels = 400, 400, 160, 80
    s.mel_filters = whisper.audio.mel_filters
    return s


def synth_audio(B, N, seed, kind):
    g = torch.Generator().manual_seed(seed)
    if kind == "noise":
        return (0.1 * torch.randn(B, N, generator=g)).clamp(-1, 1)
    t = torch.arange(N, dtype=torch.float64) / 16000.0
    out = []
    for b in range(B):
        x = 0.0
        for (f0, f1, a) in [(200.0, 3000.0, 0.5), (7000.0, 500.0, 0.2), (50.0, 120.0, 0.3)]:
            ph = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / (N / 16000.0) * (1 + 0.1 * b))
            x = x + a * torch.sin(ph)
        x = x.float() + 1e-3 * torch.randn(N, generator=g)
        out.append(x)
    x = torch.stack(out).clamp(-1, 1)
    if kind == "chirp_padded":
        x[-1, N // 2:] = 0.0  # ESPnet zero-pads shorter utterances
    return x


def main():
    meta = {}
    # ---- a2 mel filters
    mel80 = whisper.audio.mel_filters("cpu", 80).numpy()
    np.save(os.path.join(OUT, "mel_80_ref.npy"), mel80)

    # ---- LID table (product data + fixture of sample strings)
    lid = make_lid_table()
    lid.tofile(os.path.join(PKG_DATA, "lid_table_multilingual.u8"))
    sample_ids = list(range(0, 50257, 997)) + [220, 50256, 50257, 50258, 50259, 50260, 50359, 50363, 51864]
    meta["lid_samples"] = [[i, TOK.token_string(i), int(lid[i])] for i in sample_ids]
    meta["lid_hist"] = np.bincount(lid, minlength=4).tolist()

    # ---- a1 log-mel
    enc_self = fake_encoder_self()
    lm = {}
    for name, (B, N, seed, kind) in {
        "noise": (2, 16000, 2022, "noise"),
        "chirp": (2, 12800, 7, "chirp"),
        "chirp_padded": (3, 8000, 11, "chirp_padded"),
        "short": (1, 480, 3, "noise"),
    }.items():
        x = synth_audio(B, N, seed, kind)
        ilens = torch.full((B,), N, dtype=torch.long)
        if kind == "chirp_padded":
            ilens[-1] = N // 2
        y, ol = OpenAIWhisperEncoder.log_mel_spectrogram(enc_self, x, ilens)
        lm[f"{name}_audio"] = x.numpy()
        lm[f"{name}_ilens"] = ilens.numpy()
        lm[f"{name}_logmel"] = y.numpy()
        lm[f"{name}_olens"] = ol.numpy()
        # fp64 run of the same reference code (torch.stft in double): accuracy yardstick
        enc64 = fake_encoder_self()
        enc64.mel_filters = lambda dev, n: whisper.audio.mel_filters(dev, n).double()
        y64, _ = _logmel64(x)
        lm[f"{name}_logmel_f64"] = y64.numpy()
    np.savez_compressed(os.path.join(OUT, "logmel.npz"), **lm)

    # ---- a3 attention fwd + bwd (fp32 reference autograd)
    at = {}
    for name, (B, H, Tq, Tk, causal, seed, amp) in {
        "self_causal": (2, 2, 37, 37, True, 1, 1.0),
        "cross": (2, 3, 5, 50, False, 2, 1.0),
        "enc_self": (1, 2, 70, 70, False, 3, 3.0),
    }.items():
        d = 64
        D = H * d
        g = torch.Generator().manual_seed(seed)
        q = (amp * torch.randn(B, Tq, D, generator=g)).requires_grad_()
        k = (amp * torch.randn(B, Tk, D, generator=g)).requires_grad_()
        v = torch.randn(B, Tk, D, generator=g).requires_grad_()
        mha = MultiHeadAttention(D, H)
        mask = torch.empty(Tk, Tk).fill_(-np.inf).triu_(1) if causal else None
        out, qk = mha.qkv_attention(q, k, v, mask)
        dout = torch.randn(B, Tq, D, generator=g)
        dqk = torch.zeros_like(qk)
        dqk[..., 1:3] = torch.randn(B, H, Tq, 2, generator=g)
        fin = torch.isfinite(qk)
        loss = (out * dout).sum() + (torch.where(fin, qk, torch.zeros_like(qk)) * dqk).sum()
        loss.backward()
        for nm, t in dict(q=q, k=k, v=v, out=out, qk=qk, dout=dout, dqk=dqk, dq=q.grad, dk=k.grad, dv=v.grad).items():
            at[f"{name}_{nm}"] = t.detach().numpy()
        at[f"{name}_cfg"] = np.array([B, H, Tq, Tk, int(causal)])
    np.savez_compressed(os.path.join(OUT, "attention.npz"), **at)

    # ---- a9 head mask
    with open(f"{REF}/espnet/egs2/seame/asr1/attention_count_whispernoft_new.pkl", "rb") as f:
        counts = pickle.load(f)
    meta["attention_count"] = {str(l): {str(h): int(c) for h, c in d.items()} for l, d in counts.items()}

    # ---- a11/a12 pattern + cs loss, a10 head selection
    m = bare_asr_model()
    cs = {}
    L, B, H, T = 12, 3, 12, 20
    g = torch.Generator().manual_seed(5)
    toks = []
    lens = [T, 14, 9]
    # a few real vocabulary ids: english words, chinese bytes, space, digits
    eng = [i for i in range(1000, 50257) if lid[i] == O.LID_ENGLISH][:200]
    oth = [i for i in range(1000, 50257) if lid[i] == O.LID_OTHER][:200]
    both = [i for i in range(0, 50257) if lid[i] == O.LID_BOTH]
    meta["lid_both_ids"] = both[:10]
    for b in range(B):
        n_words = lens[b] - 6
        pool = eng + oth + both
        idx = torch.randint(0, len(pool), (n_words,), generator=g).tolist()
        body = [pool[i] for i in idx]
        seq = [50258, 50260, 50259, 50359, 50363] + body + [50257]
        seq = seq + [50257] * (T - len(seq))
        toks.append(seq[:T])
    toks = torch.tensor(toks, dtype=torch.long)
    maps = torch.randn(L, B, H, T, T, generator=g)
    cmask = torch.empty(T, T).fill_(-np.inf).triu_(1)
    maps = (maps + cmask).requires_grad_()
    pats = torch.stack([m.create_attention_pattern(t, 0.6).detach() for t in toks])
    loss = m.calculate_cs_loss(maps * 1.0, toks, 0.6)
    loss.backward()
    cs["tokens"] = toks.numpy()
    cs["maps"] = maps.detach().numpy()
    cs["pattern"] = pats.numpy()
    cs["loss"] = np.array(loss.item(), dtype=np.float64)
    cs["dmaps"] = maps.grad.numpy()
    # probabilities for head selection
    probs = torch.softmax(maps.detach() * 2.0, dim=-1)
    m.new_check_attention_language(probs)
    cnt = np.array([[m.attention_count[l + 1][h + 1] for h in range(12)] for l in range(12)], dtype=np.int64)
    cs["probs"] = probs.numpy()
    cs["vote_counts"] = cnt
    np.savez_compressed(os.path.join(OUT, "cs_loss.npz"), **cs)

    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, ensure_ascii=True, indent=0)
    print("wrote", sorted(os.listdir(OUT)))


def _logmel64(x):
    """The reference recipe (whisper_encoder.py:105-135) evaluated in float64 by torch."""
    x = x.double()
    window = torch.hann_window(400, dtype=torch.float64)
    stft = torch.stft(x, 400, 160, window=window, return_complex=True)
    mag = stft[..., :-1].abs() ** 2
    filt = whisper.audio.mel_filters("cpu", 80).double()
    mel = filt @ mag
    ls = torch.clamp(mel, min=1e-10).log10()
    ls = torch.maximum(ls, ls.view(x.size(0), -1).max(dim=-1)[0][:, None, None] - 8.0)
    return (ls + 4.0) / 4.0, None


def old_vote_golden():
    """ESPnetASRModel.check_attention_language (espnet_model.py:312-363, the older per-row top-2 vote) run on the
    probabilities of cs_loss.npz plus a tie-free random case -> tests/golden/head_vote_old.npz."""
    m = bare_asr_model()
    out = {}
    for name, probs in (("cs", torch.from_numpy(np.load(os.path.join(OUT, "cs_loss.npz"))["probs"])),
                        ("dense", torch.softmax(3.0 * torch.randn(12, 2, 12, 9, 9, generator=torch.Generator().manual_seed(3)), -1))):
        m.attention_count = {l: {h: 0 for h in range(1, 13)} for l in range(1, 13)}
        m.check_attention_language(probs)
        out[name + "_counts"] = np.array([[m.attention_count[l + 1][h + 1] for h in range(12)] for l in range(12)], dtype=np.int64)
        if name == "dense":
            out["dense_probs"] = probs.numpy()
    np.savez_compressed(os.path.join(OUT, "head_vote_old.npz"), **out)
    print("wrote head_vote_old.npz", {k: (v.shape, int(v.sum())) for k, v in out.items() if k.endswith("counts")})


if __name__ == "__main__" and "--oldvote" in sys.argv:
    old_vote_golden()
    sys.exit(0)

if __name__ == "__main__" and not ({"--e2e", "--e2e12", "--decode"} & set(sys.argv)):
    main()


# ----------------------------------------------------------------------------- end-to-end module golden
def e2e_golden(n_audio_layer=2, N=32000, tag="e2e_small", with_bf16=False):
    """Reference OpenAIWhisperEncoder + OpenAIWhisperDecoder + ESPnetASRModel.forward (fp32, CPU) on a 12x12-head
    decoder (the guided loss hard-codes 12 layers x 12 heads) with weights from the name-seeded initialiser that the
    product mirror shares (aga_b200.whisper_model.seeded_init_).  Stores inputs, losses and gradient summaries.

    ``--e2e12`` (n_audio_layer=12, 1 s of audio, tag e2e_full12) is the full Whisper-small depth: every one of the 12
    encoder attention layers of BASELINE configs[1] is in the parity check; it also stores a strided sample of EVERY
    adapter gradient and the reference's own bf16-autocast run (trainer.py:569 semantics on the CPU) so that the bf16 gate
    can be stated relative to the reference's own bf16-vs-fp32 deviation."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    import aga_b200  # noqa: F401  (product package: only its deterministic initialiser is used here)
    from aga_b200.whisper_model import seeded_init_
    from whisper.model import ModelDimensions, Whisper
    from espnet2.asr.decoder.whisper_decoder import OpenAIWhisperDecoder
    from espnet.nets.pytorch_backend.transformer.label_smoothing_loss import LabelSmoothingLoss

    dims = ModelDimensions(n_mels=80, n_audio_ctx=1500, n_audio_state=768, n_audio_head=12, n_audio_layer=n_audio_layer,
                           n_vocab=51865, n_text_ctx=448, n_text_state=768, n_text_head=12, n_text_layer=12)

    def fake_load_model(name, adapter=False, pe_whisper=False, side_network=False, side_network_conf=None, **kw):
        m = Whisper(dims, pe_whisper, adapter, side_network, side_network_conf)
        return seeded_init_(m, seed=0)

    whisper.load_model = fake_load_model
    whisper.available_models = lambda: ["small"]
    enc = OpenAIWhisperEncoder(whisper_model="small", adapter=True)
    dec = OpenAIWhisperDecoder(51865, 768, whisper_model="small", adapter=True, whisper_cs=True, src_layer=1)
    keys = {k: list(v.shape) for k, v in list(enc.state_dict().items()) + list(dec.state_dict().items())}

    m = bare_asr_model()
    m.vocab_size, m.ignore_id, m.sos, m.eos = 51865, -1, 50258, 50257
    m.ctc_weight, m.cs_weight, m.interctc_weight = 0.0, 0.01, 0.0
    m.encoder, m.decoder = enc, dec
    m.encoder.interctc_use_conditioning = False
    m.use_transducer_decoder = False
    m.frontend = m.specaug = m.normalize = m.preencoder = m.postencoder = None
    m.error_calculator = None
    m.criterion_att = LabelSmoothingLoss(size=51865, padding_idx=-1, smoothing=0.1, normalize_length=False)
    m.is_encoder_whisper = True
    m.c_val_attention = 0.6
    m.lang_token_id = None
    m.train()
    for n, p in m.named_parameters():
        p.requires_grad_("adapter" in n)

    g = torch.Generator().manual_seed(2022)
    B = 2
    speech = (0.1 * torch.randn(B, N, generator=g)).clamp(-1, 1)
    speech_lengths = torch.tensor([N, N - 4000])
    speech[1, N - 4000:] = 0.0
    lid = make_lid_table()
    eng = [i for i in range(1000, 50257) if lid[i] == O.LID_ENGLISH][:300]
    oth = [i for i in range(1000, 50257) if lid[i] == O.LID_OTHER][:300]
    text = torch.full((B, 16), -1, dtype=torch.long)
    tl = [16, 11]
    for b in range(B):
        pool = eng + oth
        body = [pool[i] for i in torch.randint(0, len(pool), (tl[b] - 5,), generator=g).tolist()]
        text[b, : tl[b]] = torch.tensor([50260, 50259, 50359, 50363] + body + [50257])
    text_lengths = torch.tensor(tl)
    loss, stats, weight = m(speech, speech_lengths, text.clone(), text_lengths)
    loss.backward()
    out = dict(speech=speech.numpy(), speech_lengths=speech_lengths.numpy(), text=text.numpy(),
               text_lengths=text_lengths.numpy(), loss=np.float64(loss.item()),
               loss_att=np.float64(stats["loss_att"].item()), loss_cs=np.float64(stats["loss_cs"].item()),
               acc=np.float64(float(stats["acc"])))
    with torch.no_grad():
        eo, el, _ = enc(speech, speech_lengths)
    out["encoder_out_slice"] = eo[:, :8, :16].numpy()
    out["encoder_out_lens"] = el.numpy()
    gnames, gnorms = [], []
    for n, p in sorted(m.named_parameters()):
        if p.grad is not None:
            gnames.append(n)
            gnorms.append(float(p.grad.double().norm()))
    out["grad_norms"] = np.array(gnorms)
    pick = "decoder.decoders.blocks.5.adapter_attn.model.2.bias"
    out["grad_pick"] = dict(m.named_parameters())[pick].grad.numpy()
    pick2 = "encoder.encoders.blocks.0.adapter_mlp.model.0.bias"
    out["grad_pick2"] = dict(m.named_parameters())[pick2].grad.numpy()
    if with_bf16:
        # a strided sample (<= 512 elements) of EVERY adapter gradient, and the reference's own bf16-autocast step
        params = dict(m.named_parameters())
        samples = []
        for n in gnames:
            flat = params[n].grad.reshape(-1)
            idx = torch.linspace(0, flat.numel() - 1, min(512, flat.numel())).long()
            samples.append(flat[idx].numpy())
        out["grad_samples"] = np.concatenate(samples)
        out["grad_sample_sizes"] = np.array([len(x) for x in samples])
        out["grad_absmax"] = np.array([float(params[n].grad.abs().max()) for n in gnames])
        for p in m.parameters():
            p.grad = None
        with torch.autocast("cpu", dtype=torch.bfloat16):
            loss16, stats16, _ = m(speech, speech_lengths, text.clone(), text_lengths)
        loss16.backward()
        out["bf16_loss_att"] = np.float64(stats16["loss_att"].item())
        out["bf16_loss_cs"] = np.float64(stats16["loss_cs"].item())
        out["bf16_grad_norms"] = np.array([float(params[n].grad.double().norm()) for n in gnames])
        s16 = []
        for n in gnames:
            flat = params[n].grad.reshape(-1).float()
            idx = torch.linspace(0, flat.numel() - 1, min(512, flat.numel())).long()
            s16.append(flat[idx].numpy())
        out["bf16_grad_samples"] = np.concatenate(s16)
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    with open(os.path.join(OUT, f"{tag}_meta.json"), "w") as f:
        json.dump({"state_dict": keys, "grad_names": gnames, "grad_pick": pick, "grad_pick2": pick2}, f, indent=0)
    print("e2e golden: loss", loss.item(), {k: (float(v) if v is not None else None) for k, v in stats.items()})


if __name__ == "__main__" and "--e2e" in sys.argv:
    e2e_golden()
if __name__ == "__main__" and "--e2e12" in sys.argv:
    e2e_golden(n_audio_layer=12, N=16000, tag="e2e_full12", with_bf16=True)


# ----------------------------------------------------------------------------- config 1: bundled utterance
def decode_golden():
    """BASELINE.json configs[0]: Whisper-small greedy decode of the bundled SEAME utterance on CPU with the decoder
    self-attention maps dumped (code_util/whisper_check.py, attention_map.md) — run through the REFERENCE modules
    (OpenAIWhisperEncoder.forward, OpenAIWhisperDecoder.batch_score) with the name-seeded weights (no checkpoint
    exists offline).  Stores the decoded PCM, the greedy token ids and the last step's maps."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    sys.path.insert(0, os.path.join(HERE, "..", "tools"))
    import aga_b200  # noqa: F401
    from aga_b200.whisper_model import seeded_init_
    from flac_decode import decode_flac
    from whisper.model import ModelDimensions, Whisper
    from espnet2.asr.decoder.whisper_decoder import OpenAIWhisperDecoder

    pcm, sr, bps = decode_flac(f"{REF}/code_util/nc41m-46nc41mbp_0101-047421-047682.flac")
    assert sr == 16000 and bps == 16 and pcm.shape == (1, 41760)
    dims = ModelDimensions(80, 1500, 768, 12, 12, 51865, 448, 768, 12, 12)
    whisper.load_model = lambda name, adapter=False, pe_whisper=False, side_network=False, side_network_conf=None, **kw: \
        seeded_init_(Whisper(dims, pe_whisper, adapter, side_network, side_network_conf), seed=0)
    whisper.available_models = lambda: ["small"]
    enc = OpenAIWhisperEncoder(whisper_model="small", adapter=True).eval()
    dec = OpenAIWhisperDecoder(51865, 768, whisper_model="small", adapter=True, whisper_cs=True, src_layer=1).eval()
    speech = torch.from_numpy(pcm[0].astype(np.float32) / 32768.0)[None]  # librosa.load scaling
    maps = {}
    hooks = [blk.register_forward_hook(lambda m, i, o, l=l: maps.__setitem__(l, o[1].detach().clone()))
             for l, blk in enumerate(dec.decoders.blocks)]
    with torch.no_grad():
        enc_out, enc_lens, _ = enc(speech, torch.tensor([speech.shape[1]]))
        ys = torch.tensor([[50258, 50260, 50259, 50359, 50363]])  # asr_inference.py:324
        ids, margins, chosen = [], [], []
        for step in range(12):
            logp, _ = dec.batch_score(ys, [None], enc_out)
            top2 = logp[0].topk(2)
            ids.append(int(top2.indices[0]))
            margins.append(float(top2.values[0] - top2.values[1]))
            chosen.append(float(top2.values[0]))
            ys = torch.cat([ys, top2.indices[:1][None]], dim=1)
    for h in hooks:
        h.remove()
    last_maps = torch.stack([maps[l][0] for l in range(12)]).numpy()  # (12, H, t, t) of the final step
    np.savez_compressed(os.path.join(OUT, "decode_seame.npz"), pcm=pcm[0].astype(np.int16),
                        enc_out_lens=enc_lens.numpy(), enc_out_slice=enc_out[0, :8, :16].numpy(),
                        token_ids=np.array(ids), margins=np.array(margins), logp=np.array(chosen), last_maps=last_maps)
    print("decode golden:", ids, "min margin", min(margins), "enc frames", int(enc_lens[0]))


if __name__ == "__main__" and "--decode" in sys.argv:
    decode_golden()
