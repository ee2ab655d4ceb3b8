"""Mirror of the Whisper / attention-decoder branch of ESPnetASRModel (espnet2/asr/espnet_model.py) with the
attention-guided adaptation additions: head-mask construction (:186-219), language pattern (:236-275), head vote
(:285-310), guided loss (:463-530) and the loss plumbing (:534-710, :900-961).

CTC / transducer / inter-CTC branches are out of scope (``ctc_weight: 0.0`` in every AGA recipe) and raise.
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import ops

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

# The literal 12x12 matrix the reference loss multiplies with (espnet_model.py:514-525), one bit-string per layer.
# It equals the top-72 selection of SEAME/attention_count_whispernoft_new.pkl (tests/test_oracle_golden.py).
LITERAL_HEAD_MASK_ROWS = (
    "000000000000", "000000000000", "000000000000", "011101100111", "001111011111", "111111011101",
    "111111111110", "011111111111", "100110101010", "111100100010", "111110101001", "010010000001")


def literal_head_mask() -> torch.Tensor:
    return torch.tensor([[float(c) for c in row] for row in LITERAL_HEAD_MASK_ROWS], dtype=torch.float32)


def load_lid_table(path: Optional[str] = None) -> torch.Tensor:
    """(51865,) uint8 language-id class per multilingual Whisper token id (0 other, 1 English, 2 space-only, 3 EOT).
    Generated once from the reference's ``whisper/assets/multilingual.tiktoken`` by oracle/make_golden.py with the
    exact string tests of espnet_model.py:234-258."""
    path = path or os.path.join(_DATA, "lid_table_multilingual.u8")
    return torch.from_numpy(np.fromfile(path, dtype=np.uint8))


def build_lid_table(tiktoken_path: str, vocab_size: int = 51865) -> torch.Tensor:
    """The product's own generator of the language-id table (no test infrastructure involved): reads a Whisper BPE asset
    (``whisper/assets/multilingual.tiktoken``: one ``base64(token bytes) rank`` pair per line) and classifies every token
    id with the string tests of ESPnetASRModel.is_english / create_attention_pattern (espnet_model.py:234-258) on the
    token's GPT-2 "bytes to unicode" spelling — what the reference gets from HF ``convert_ids_to_tokens``:
    3 = <|endoftext|> (id = number of BPE ranks), 2 = only the space marker, 1 = ASCII letters only (English), 0 = other.
    ``load_lid_table()`` reads the table shipped in ``data/`` (generated from the reference's own asset)."""
    import base64
    import string
    ranks = {}
    with open(tiktoken_path) as f:
        for line in f:
            if line.strip():
                tok, rank = line.split()
                ranks[int(rank)] = base64.b64decode(tok)
    # GPT-2 byte -> printable unicode map: the space byte 0x20 becomes 'Ġ' (U+0120)
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    cs, n = bs[:], 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    b2u = dict(zip(bs, (chr(c) for c in cs)))
    letters = set(string.ascii_letters)
    table = np.zeros(vocab_size, dtype=np.uint8)
    for i in range(vocab_size):
        if i == len(ranks):          # the first special id: <|endoftext|> (50257 for the multilingual vocabulary)
            table[i] = 3
            continue
        if i > len(ranks):           # other specials ("<|...|>"): not letters-only -> class 0
            continue
        word = "".join(b2u[x] for x in ranks[i]).replace("\u0120", "")
        table[i] = 2 if word == "" else (1 if all(ch in letters for ch in word) else 0)
    return torch.from_numpy(table)


def select_heads(attention_count: Dict[int, Dict[int, int]], head_percentage: float, n_layers: int, n_heads: int,
                 base: int = 110) -> torch.Tensor:
    """espnet_model.py:202-216: flatten in dict order, STABLE sort by count (desc), keep the first
    int(110*head_percentage/100) entries with count > 0 -> (n_layers, n_heads) 0/1 mask."""
    flat = [(a, b, c) for a, inner in attention_count.items() for b, c in inner.items()]
    flat.sort(key=lambda x: x[2], reverse=True)  # list.sort is stable like sorted()
    out = torch.zeros((n_layers, n_heads), dtype=torch.float32)
    for a, b, c in flat[: int(base * head_percentage / 100)]:
        if c > 0:
            out[int(a) - 1][int(b) - 1] = 1
    return out


def _load_attention_count(src) -> Dict[int, Dict[int, int]]:
    if isinstance(src, dict):
        return {int(l): {int(h): int(c) for h, c in d.items()} for l, d in src.items()}
    with open(src, "rb") as f:
        head = f.read(1)
        f.seek(0)
        raw = json.load(f) if head in (b"{", b"[") else pickle.load(f)
    return {int(l): {int(h): int(c) for h, c in d.items()} for l, d in raw.items()}


def pad_list(xs: List[torch.Tensor], pad_value):
    n, m = len(xs), max(x.size(0) for x in xs)
    out = xs[0].new_full((n, m) + tuple(xs[0].shape[1:]), pad_value)
    for i, x in enumerate(xs):
        out[i, : x.size(0)] = x
    return out


def add_sos_eos(ys_pad: torch.Tensor, sos: int, eos: int, ignore_id: int):
    """espnet/nets/pytorch_backend/transformer/add_sos_eos.py:12-31."""
    ys = [y[y != ignore_id] for y in ys_pad]
    _sos, _eos = ys_pad.new_tensor([sos]), ys_pad.new_tensor([eos])
    return (pad_list([torch.cat([_sos, y]) for y in ys], eos), pad_list([torch.cat([y, _eos]) for y in ys], ignore_id))


def add_sos_eos_static(ys_pad: torch.Tensor, ys_lens: torch.Tensor, sos: int, eos: int, ignore_id: int):
    """Same result as add_sos_eos for the batches ESPnet's collate_fn produces (ignore_id only as trailing padding,
    ys_pad already cut to the longest target) without the per-utterance boolean indexing, i.e. without a
    device->host sync: usable inside a CUDA graph."""
    B, Lmax = ys_pad.shape
    pad = ys_pad == ignore_id
    ys_in = torch.cat([ys_pad.new_full((B, 1), sos), ys_pad.masked_fill(pad, eos)], dim=1)
    ys_out = torch.cat([ys_pad, ys_pad.new_full((B, 1), ignore_id)], dim=1)
    ys_out = ys_out.scatter(1, ys_lens.view(B, 1).to(torch.int64), eos)
    return ys_in, ys_out


class LabelSmoothingLoss(torch.nn.Module):
    """espnet/nets/pytorch_backend/transformer/label_smoothing_loss.py:14-63 (KL to the smoothed one-hot)."""

    def __init__(self, size: int, padding_idx: int, smoothing: float, normalize_length: bool = False):
        super().__init__()
        self.size, self.padding_idx, self.smoothing, self.normalize_length = size, padding_idx, smoothing, normalize_length
        self.confidence = 1.0 - smoothing

    def forward(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        batch_size = x.size(0)
        x = x.view(-1, self.size)
        target = target.reshape(-1)
        ignore = target == self.padding_idx
        total = ignore.numel() - ignore.sum()
        tgt = target.masked_fill(ignore, 0)
        logp = torch.log_softmax(x, dim=1)
        # sum_v t_v (log t_v - logp_v) with t = eps/(V-1) off the target and 1-eps on it, without materialising t
        eps = self.smoothing / (self.size - 1)
        const = self.confidence * float(np.log(self.confidence)) if self.confidence > 0 else 0.0
        if eps > 0:
            const += (self.size - 1) * eps * float(np.log(eps))
        picked = logp.gather(1, tgt.unsqueeze(1)).squeeze(1)
        kl = const - (eps * (logp.sum(dim=1) - picked) + self.confidence * picked)
        denom = total if self.normalize_length else batch_size
        return kl.masked_fill(ignore, 0.0).sum() / denom


def th_accuracy(pad_outputs: torch.Tensor, pad_targets: torch.Tensor, ignore_label: int) -> torch.Tensor:
    """espnet/nets/pytorch_backend/nets_utils.py:304-324, kept on the device (no .item() sync in the step)."""
    pred = pad_outputs.view(pad_targets.size(0), pad_targets.size(1), -1).argmax(2)
    mask = pad_targets != ignore_label
    return ((pred == pad_targets) & mask).sum().float() / mask.sum().float()


class ESPnetASRModel(torch.nn.Module):
    def __init__(self, vocab_size: int, token_list: Union[Tuple[str, ...], List[str], None], frontend=None,
                 specaug=None, normalize=None, preencoder=None, encoder=None, postencoder=None, decoder=None, ctc=None,
                 joint_network=None, aux_ctc: dict = None, ctc_weight: float = 0.0, cs_weight: float = 0.0,
                 interctc_weight: float = 0.0, ignore_id: int = -1, lsm_weight: float = 0.0,
                 length_normalized_loss: bool = False, report_cer: bool = False, report_wer: bool = False,
                 sym_space: str = "<space>", sym_blank: str = "<blank>", sym_sos: str = "<sos/eos>",
                 sym_eos: str = "<sos/eos>", extract_feats_in_collect_stats: bool = True, lang_token_id: int = -1,
                 c_val_attention: float = 0.6, head_percentage: float = 100.0,
                 # additions that replace paths / network access hard-coded in the reference (:200, :226)
                 attention_count=None, lid_table: Optional[torch.Tensor] = None, use_literal_head_mask: bool = True,
                 n_early_layers: int = 2):
        super().__init__()
        if ctc_weight != 0.0 or joint_network is not None or interctc_weight != 0.0:
            raise NotImplementedError("only the attention-decoder branch (ctc_weight 0.0) is on the AGA hot path")
        assert frontend is None, "frontend should be None when using full Whisper model"  # :221-223
        token_list = list(token_list) if token_list is not None else []
        self.sos = token_list.index(sym_sos) if sym_sos in token_list else vocab_size - 1
        self.eos = token_list.index(sym_eos) if sym_eos in token_list else vocab_size - 1
        self.vocab_size, self.ignore_id = vocab_size, ignore_id
        self.ctc_weight, self.cs_weight, self.interctc_weight = ctc_weight, cs_weight, interctc_weight
        self.token_list = token_list
        self.frontend, self.specaug, self.normalize = frontend, specaug, normalize
        self.preencoder, self.postencoder = preencoder, postencoder
        self.encoder, self.decoder, self.ctc = encoder, decoder, None
        self.criterion_att = LabelSmoothingLoss(vocab_size, ignore_id, lsm_weight, length_normalized_loss)
        self.is_encoder_whisper = "Whisper" in type(self.encoder).__name__
        self.c_val_attention = c_val_attention
        self.n_early_layers = n_early_layers
        n_layers = len(decoder.decoders.blocks)
        n_heads = decoder.decoders.blocks[0].attn.n_head
        self.attention_count = {l: {h: 0 for h in range(1, n_heads + 1)} for l in range(1, n_layers + 1)}  # :184-196
        self.head_percentage = head_percentage
        sel = None
        if cs_weight and attention_count is not None:
            sel = select_heads(_load_attention_count(attention_count), head_percentage, n_layers, n_heads)
        self.register_buffer("selected_heads", sel if sel is not None else torch.zeros(n_layers, n_heads), persistent=False)
        # The reference multiplies with the literal (:527); `self.selected_heads` is commented out (:528).
        if use_literal_head_mask and (n_layers, n_heads) == (12, 12):
            mask = literal_head_mask()
        elif sel is not None:
            mask = sel
        else:
            mask = torch.zeros(n_layers, n_heads)
        self.register_buffer("cs_head_mask", mask, persistent=False)
        self.register_buffer("lid_table", lid_table if lid_table is not None else load_lid_table(), persistent=False)
        self.lang_token_id = torch.tensor([[lang_token_id]]) if lang_token_id != -1 else None
        # static_shapes=True: the caller guarantees `text` is already cut to the longest target and padded only at the
        # end (what ESPnet's collate_fn produces); the step then runs without device->host syncs (CUDA-graph capturable)
        self.static_shapes = False

    # ------------------------------------------------------------------ a11
    def create_attention_pattern(self, ground_truth_token: torch.Tensor, attention_default: float = 0.6) -> torch.Tensor:
        """(T,) or (B,T) token ids -> (T,2) / (B,T,2) target pattern, inf on pad rows (:236-275)."""
        single = ground_truth_token.dim() == 1
        toks = ground_truth_token[None] if single else ground_truth_token
        pat = ops.attention_pattern(toks, self.lid_table, attention_default)
        return pat[0] if single else pat

    # ------------------------------------------------------------------ a10
    def new_check_attention_language(self, attention_maps: torch.Tensor) -> None:
        """attention_maps (L,B,H,T,T) softmax probabilities; updates self.attention_count (:285-310)."""
        _, counts = ops.head_vote(attention_maps)
        for (l, h) in torch.nonzero(counts).tolist():
            self.attention_count[l + 1][h + 1] += int(counts[l, h])

    # ------------------------------------------------------------------ a12
    def check_attention_language(self, attention_maps: torch.Tensor, k: int = 2) -> None:
        """The older head vote (:312-363, "old formulation"): per (utterance, layer, head) the two key indices that
        appear most often among the rows' top-2 entries must be {1, 2}.  Runs where the maps live (no per-head Python
        loop, no host copies of the maps); ties are resolved towards the smaller index (stable sorts), which is also
        what the reference's ``torch.unique`` + stable ``sorted`` do for equal counts."""
        L, B, H, T, _ = attention_maps.shape
        top = torch.argsort(attention_maps, dim=-1, descending=True, stable=True)[..., :k]        # (L,B,H,T,k)
        votes = torch.zeros((L, B, H, T), dtype=torch.int64, device=attention_maps.device)
        votes.scatter_add_(-1, top.reshape(L, B, H, T * k), torch.ones((L, B, H, T * k), dtype=torch.int64, device=votes.device))
        best = torch.argsort(votes, dim=-1, descending=True, stable=True)[..., :k]               # most frequent indices
        picked = ((best == 1).any(-1) & (best == 2).any(-1)).sum(dim=1).cpu()                    # (L,H) utterance counts
        for layer in range(L):
            for head in range(H):
                self.attention_count[layer + 1][head + 1] += int(picked[layer, head])

    def calculate_cs_loss(self, attention_maps: torch.Tensor, ground_truth_token: torch.Tensor,
                          attention_default: float = 0.6) -> torch.Tensor:
        """attention_maps: (L,B,H,T,T) full maps (reference layout) or the compact (L,B,H,T,2) export of key
        columns 1:3; ground_truth_token = ys_in_pad (B,T).  (:463-530)"""
        if isinstance(attention_maps, ops.GuidedParts):  # decoder(export_mode="fused"): the reduction over t is already done
            return ops.guided_loss_from_parts(attention_maps, self.cs_head_mask)
        slab = attention_maps if attention_maps.shape[-1] == 2 and attention_maps.shape[-2] != 2 \
            else attention_maps[..., 1:3]
        pattern = self.create_attention_pattern(ground_truth_token, attention_default)
        return ops.guided_loss(slab, pattern, self.cs_head_mask, self.n_early_layers)

    # ------------------------------------------------------------------ plumbing
    def encode(self, speech: torch.Tensor, speech_lengths: torch.Tensor, valid_samples: Optional[torch.Tensor] = None):
        """:723-788 for frontend=None / Whisper: the encoder consumes raw audio."""
        if valid_samples is None:
            encoder_out, encoder_out_lens, _ = self.encoder(speech, speech_lengths)
        else:
            encoder_out, encoder_out_lens, _ = self.encoder(speech, speech_lengths, valid_samples=valid_samples)
        assert encoder_out.size(0) == speech.size(0)
        return encoder_out, encoder_out_lens

    def _calc_att_loss(self, encoder_out, encoder_out_lens, ys_pad, ys_pad_lens, memory_len=None):
        """:900-961."""
        if self.lang_token_id is not None:
            ys_pad = torch.cat([self.lang_token_id.repeat(ys_pad.size(0), 1).to(ys_pad.device), ys_pad], dim=1)
            ys_pad_lens = ys_pad_lens + 1
        if self.static_shapes:
            ys_in_pad, ys_out_pad = add_sos_eos_static(ys_pad, ys_pad_lens, self.sos, self.eos, self.ignore_id)
        else:
            ys_in_pad, ys_out_pad = add_sos_eos(ys_pad, self.sos, self.eos, self.ignore_id)
        ys_in_lens = ys_pad_lens + 1
        if getattr(self.decoder, "export_mode", None) == "fused" and self.cs_weight != 0:
            # the attention epilogue reduces the guided loss: it needs the target pattern of this batch
            self.decoder.guided_pattern = self.create_attention_pattern(ys_in_pad, self.c_val_attention)
            self.decoder.n_early_layers = self.n_early_layers
        if memory_len is None:
            decoder_out, att_map = self.decoder(encoder_out, encoder_out_lens, ys_in_pad, ys_in_lens)
        else:
            decoder_out, att_map = self.decoder(encoder_out, encoder_out_lens, ys_in_pad, ys_in_lens, memory_len=memory_len)
        fused = isinstance(decoder_out, ops.VocabLogits)
        if fused:  # decoder(fused_loss=True): KL loss + accuracy straight from the padded logits, no fp32 (B,T,V) tensor
            c = self.criterion_att
            loss_att, acc_att = ops.ls_cross_entropy(decoder_out, ys_out_pad, c.padding_idx, c.smoothing, c.normalize_length)
        else:
            loss_att = self.criterion_att(decoder_out, ys_out_pad)
        loss_cs = None
        if self.is_encoder_whisper and self.cs_weight != 0:
            loss_cs = self.calculate_cs_loss(att_map, ys_in_pad, self.c_val_attention)
        if not fused:
            acc_att = th_accuracy(decoder_out.view(-1, self.vocab_size), ys_out_pad, ignore_label=self.ignore_id)
        return loss_att, acc_att, None, None, loss_cs

    def forward(self, speech: torch.Tensor, speech_lengths: torch.Tensor, text: torch.Tensor,
                text_lengths: torch.Tensor, valid_samples: Optional[torch.Tensor] = None, **kwargs):
        """:534-710 (attention-decoder branch): returns (loss, stats, weight).

        ``valid_samples`` (device int32 scalar, not in the reference): ``speech`` is zero-padded to a static bucket length
        and this is the batch's true padded length — what the reference's collate_fn would have produced.  The log-mel
        frontend, the conv stem, the encoder self attention and the decoder cross attention then compute exactly the
        unpadded batch (graphed.BucketedTrainStep)."""
        assert text_lengths.dim() == 1, text_lengths.shape
        assert speech.shape[0] == speech_lengths.shape[0] == text.shape[0] == text_lengths.shape[0]
        batch_size = speech.shape[0]
        if not self.static_shapes:
            text = text[:, : text_lengths.max()]  # "for data-parallel" (:566); needs a host sync
        encoder_out, encoder_out_lens = self.encode(speech, speech_lengths, valid_samples)
        memory_len = None if valid_samples is None else self.encoder.encoder_frames(valid_samples)
        loss_att, acc_att, cer_att, wer_att, loss_cs = self._calc_att_loss(encoder_out, encoder_out_lens, text, text_lengths,
                                                                           memory_len=memory_len)
        loss = loss_att
        stats = dict()
        if self.cs_weight != 0.0:
            loss = self.cs_weight * loss_cs + loss_att  # :694
            stats["loss_cs"] = loss_cs.detach()
        stats["loss_att"] = loss_att.detach()
        stats["acc"] = acc_att
        stats["cer"], stats["wer"] = cer_att, wer_att
        stats["loss"] = loss.detach()
        weight = loss.new_full((), float(batch_size))  # no host->device copy (CUDA-graph capturable)
        return loss, stats, weight
