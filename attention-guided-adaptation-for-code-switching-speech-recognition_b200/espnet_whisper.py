"""Drop-in mirrors of the reference's ESPnet2 plugin modules for the Whisper path:

  OpenAIWhisperEncoder  <- espnet2/asr/encoder/whisper_encoder.py:12-243
  OpenAIWhisperDecoder  <- espnet2/asr/decoder/whisper_decoder.py:12-273

Same constructor keywords, same forward signatures and return conventions, same ``state_dict`` keys
(``encoders.*`` / ``decoders.*``); log-mel and attention run in the aga_b200 CUDA library.
"""
from __future__ import annotations

import contextlib
import copy
from typing import Any, List, Optional, Tuple, Union

import torch
import torch.nn.functional as F

from . import ops
from . import whisper_model as W
from .specaug import SpecAug

N_SAMPLES = 480000  # whisper/audio.py:18


class OpenAIWhisperEncoder(torch.nn.Module):
    def __init__(self, input_size: int = 1, dropout_rate: float = 0.0, whisper_model: str = "small",
                 download_dir: Optional[str] = None, use_specaug: bool = False,
                 specaug_conf: Union[dict, None] = None, do_pad_trim: bool = False, pe_whisper: bool = False,
                 adapter: bool = False, side_network: bool = False, side_network_conf=None, seed: int = 0):
        super().__init__()
        self.n_fft, self.win_length, self.hop_length = ops.N_FFT, ops.N_FFT, ops.HOP_LENGTH
        self.mel_filters = ops.mel_filters
        self.dropout = torch.nn.Dropout(dropout_rate)
        if whisper_model not in W.available_models():
            import os
            assert os.path.isfile(whisper_model), f"unknown whisper model {whisper_model}"
        _model = W.load_model(whisper_model, adapter, pe_whisper, side_network, side_network_conf,
                              download_root=download_dir, seed=seed)
        self.sidenetwork = side_network
        self.encoders = copy.deepcopy(_model.encoder)
        self.n_mels = _model.dims.n_mels
        self.encoders.train()
        del _model
        self.specaug = SpecAug(**specaug_conf) if use_specaug else None
        self.do_pad_trim = do_pad_trim
        self.pad_samples = N_SAMPLES
        self.interctc_use_conditioning = False

    def output_size(self) -> int:
        return self.encoders.ln_post.normalized_shape[-1]

    def pad_or_trim(self, array: torch.Tensor, length: int, axis: int = -1) -> torch.Tensor:
        """Zero-pad or cut to ``length`` samples (whisper_encoder.py:84-103)."""
        n = array.shape[axis]
        if n > length:
            array = array.narrow(axis, 0, length)
        elif n < length:
            pad = [0, 0] * array.ndim
            pad[2 * (array.ndim - 1 - (axis % array.ndim)) + 1] = length - n
            array = F.pad(array, pad)
        return array

    def log_mel_spectrogram(self, audio: torch.Tensor, ilens: torch.Tensor = None, valid_samples: torch.Tensor = None):
        """whisper_encoder.py:105-135 -> one fused CUDA pass (csrc/logmel_tc.cu; csrc/logmel.cu for non-banded filterbanks)."""
        return ops.log_mel_spectrogram(audio, ilens, n_mels=self.n_mels, valid_samples=valid_samples)

    @staticmethod
    def encoder_frames(valid_samples: torch.Tensor) -> torch.Tensor:
        """Encoder output frames of ``valid_samples`` audio samples (conv2: k 3, stride 2, pad 1), as a device int32 scalar."""
        mel = valid_samples.to(torch.int64) // ops.HOP_LENGTH
        return ((mel - 1) // 2 + 1).to(torch.int32)

    def whisper_encode(self, input: torch.Tensor, ilens: torch.Tensor = None, valid_samples: torch.Tensor = None):
        """whisper_encoder.py:137-222 (no side network).  ``valid_samples``: the batch is zero-padded to a static bucket
        length (graphed.BucketedTrainStep); frames past the true length are never attended to."""
        enc = self.encoders
        kv_len = None
        if valid_samples is None:
            x = enc.stem(input)  # conv1 + GELU + conv2 + GELU, token-major
        else:
            x = enc.stem(input, valid_frames=valid_samples.to(torch.int64) // ops.HOP_LENGTH)
            kv_len = self.encoder_frames(valid_samples)
        n_frames, max_pos = x.size(1), enc.positional_embedding.size(0)
        if n_frames <= max_pos:
            x = (x + enc.positional_embedding[:n_frames, :]).to(x.dtype)
        else:  # audio > 30 s: truncated to the positional table (:163-165)
            x = x[:, :max_pos, :] + enc.positional_embedding
        x = self.dropout(x)
        x = W.run_blocks(enc.blocks, x, enc.ln_post, between=self.dropout,
                         between_changes_x=self.training and self.dropout.p > 0, kv_len=kv_len)
        if ilens is not None:
            olens = 1 + (ilens - enc.conv2.kernel_size[0] + 2 * enc.conv2.padding[0]) // enc.conv2.stride[0]
            olens = torch.clamp(olens, max=max_pos)
        else:
            olens = None
        return x, olens

    def forward(self, xs_pad: torch.Tensor, ilens: torch.Tensor, prev_states: torch.Tensor = None,
                valid_samples: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
        if self.do_pad_trim:
            xs_pad = self.pad_or_trim(xs_pad, self.pad_samples)
        feats, feats_lens = self.log_mel_spectrogram(xs_pad, ilens, valid_samples)
        if self.specaug is not None and self.encoders.training:
            feats, feats_lens = self.specaug(feats, feats_lens)
        xs_pad, olens = self.whisper_encode(feats, feats_lens, valid_samples)
        return xs_pad, olens, None


class WhisperFrontend(torch.nn.Module):
    """Mirror of espnet2/asr/frontend/whisper.py:11-135 (``frontend: whisper`` in the task registry, tasks/asr.py:95): the
    Whisper log-mel + audio encoder used as a feature extractor.  Same constructor keywords, ``output_size()``,
    ``log_mel_spectrogram`` (the twin of the encoder's, :54-83 — the same fused CUDA pass here), ``whisper_encode`` and
    ``forward(input, input_lengths) -> (feats (B, T', D), feats_lens)``; the whole Whisper model hangs under ``.whisper``
    as in the reference, so ``state_dict`` keys are ``whisper.encoder.*`` / ``whisper.decoder.*``.  (The reference's
    ``x = block(x)`` at :104 no longer works with the fork's tuple-returning blocks; the tuple is unpacked here.)"""

    def __init__(self, whisper_model: str = "small", freeze_weights: bool = True, download_dir: Optional[str] = None,
                 seed: int = 0):
        super().__init__()
        self.n_fft, self.win_length, self.hop_length = ops.N_FFT, ops.N_FFT, ops.HOP_LENGTH
        assert whisper_model in W.available_models(), f"unknown whisper model {whisper_model}"
        self.whisper = W.load_model(whisper_model, download_root=download_dir, seed=seed)
        self.whisper.eval()
        self.n_mels = self.whisper.dims.n_mels
        self.mel_filters = ops.mel_filters
        self.freeze_weights = freeze_weights

    def output_size(self) -> int:
        return self.whisper.encoder.ln_post.normalized_shape[-1]

    def pad_or_trim(self, array: torch.Tensor, length: int = N_SAMPLES, axis: int = -1) -> torch.Tensor:
        return OpenAIWhisperEncoder.pad_or_trim(self, array, length, axis)

    def log_mel_spectrogram(self, audio: torch.Tensor, ilens: torch.Tensor = None):
        return ops.log_mel_spectrogram(audio, ilens, n_mels=self.n_mels)

    def whisper_encode(self, input: torch.Tensor, ilens: torch.Tensor = None):
        enc = self.whisper.encoder
        x = enc.stem(input)
        n_frames, max_pos = x.size(1), enc.positional_embedding.size(0)
        if n_frames <= max_pos:
            x = (x + enc.positional_embedding[:n_frames, :]).to(x.dtype)
        else:
            x = x[:, :max_pos, :] + enc.positional_embedding
        for block in enc.blocks:
            x, _ = block(x)
        x = enc.ln_post(x)
        olens = None
        if ilens is not None:
            olens = 1 + (ilens - enc.conv2.kernel_size[0] + 2 * enc.conv2.padding[0]) // enc.conv2.stride[0]
            olens = torch.clamp(olens, max=max_pos)
        return x, olens

    def forward(self, input: torch.Tensor, input_lengths: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        feats, feats_lens = self.log_mel_spectrogram(input, input_lengths)
        with torch.no_grad() if self.freeze_weights else contextlib.nullcontext():
            feats, feats_lens = self.whisper_encode(feats, feats_lens)
        return feats, feats_lens


class GraphedGreedyDecoder:
    """Greedy decoding as ONE captured CUDA graph per token (the next step after the KV cache).

    The reference recomputes the whole prefix on every step and copies every layer's map to the host
    (whisper_decoder.py:172-244); the KV-cached ``forward_one_step`` of this package is O(t) per token but still ~200
    kernel launches from Python (8-9 ms per token).  Here every shape is static: the self-attention K / V live in
    preallocated (n, max_len, 2D) [K | V] buffers, the position is a DEVICE scalar (it indexes the positional embedding, the cache
    row to write and — as ``kv_len`` — the number of keys the attention kernel may see), the arg-max token is fed back
    on the device.  ``prefill`` runs the prompt through the normal KV-cached path; ``decode(n)`` replays the captured
    step n times with no host round trip and returns (token ids (n_hyp, n), their log-probabilities).
    Same tokens and log-probabilities as ``forward_one_step`` (tests/test_gpu_decode.py)."""

    def __init__(self, decoder: "OpenAIWhisperDecoder", memory: torch.Tensor, max_len: int = 448):
        self.dec, self.memory, self.max_len = decoder, memory, max_len
        self.n = memory.size(0)
        self.graph = None

    @torch.no_grad()
    def prefill(self, prompt: torch.Tensor) -> None:
        """prompt (n, t0) int64: fills the caches with the prompt's keys / values and leaves the first generated token in
        ``self.tok`` (as greedy ``forward_one_step`` on the prompt would)."""
        d, dec = self.dec, self.dec.decoders
        was = d.kv_cache
        d.kv_cache = True
        try:
            logp, cache = d._forward_one_step_cached(prompt, self.memory, None)
        finally:
            d.kv_cache = was
        dev, dt = self.memory.device, self.memory.dtype
        D = dec.token_embedding.embedding_dim
        t0 = prompt.size(1)
        best = logp.max(dim=-1)
        if self.graph is None:
            self.kv = [torch.zeros(self.n, self.max_len, 2 * D, device=dev, dtype=dt) for _ in dec.blocks]  # [K | V] rows
            self.cross = [(kc.contiguous(), vc.contiguous()) for (_, _, kc, vc) in cache]
            self.tok = torch.zeros(self.n, 1, dtype=torch.int64, device=dev)
            self.tok_logp = torch.zeros(self.n, dtype=torch.float32, device=dev)
            self.pos = torch.zeros((1,), dtype=torch.int64, device=dev)
            self.out_tok = torch.zeros(self.n, self.max_len, dtype=torch.int64, device=dev)
            self.out_logp = torch.zeros(self.n, self.max_len, dtype=torch.float32, device=dev)
            self.n_out = torch.zeros((1,), dtype=torch.int64, device=dev)
        # a captured step stays valid across prompts: every buffer it reads is refilled IN PLACE
        for layer, (k, v, kc, vc) in enumerate(cache):
            self.kv[layer].zero_()
            self.kv[layer][:, :t0, :D] = k
            self.kv[layer][:, :t0, D:] = v
            if self.graph is not None:
                self.cross[layer][0].copy_(kc)
                self.cross[layer][1].copy_(vc)
        self.tok.copy_(best.indices.view(self.n, 1))          # token at position t0 (not yet in the cache)
        self.tok_logp.copy_(best.values)
        self.pos.fill_(t0)
        self.out_tok.zero_()
        self.out_logp.zero_()
        self.n_out.zero_()

    def _step(self) -> None:
        """Consume ``self.tok`` at position ``self.pos``; leave the next token in ``self.tok``."""
        dec = self.dec.decoders
        self.out_tok.index_copy_(1, self.n_out, self.tok)
        self.out_logp.index_copy_(1, self.n_out, self.tok_logp.view(self.n, 1))
        x = dec.token_embedding(self.tok) + dec.positional_embedding.index_select(0, self.pos)
        x = x.to(self.memory.dtype)
        kv_len = (self.pos + 1).to(torch.int32)
        for layer, block in enumerate(dec.blocks):
            x = block.step_static(x, self.kv[layer], self.pos, kv_len, self.cross[layer])
        logp = torch.log_softmax(dec.vocab_logits(dec.ln(x[:, -1])), dim=-1)
        best = logp.max(dim=-1)
        self.tok.copy_(best.indices.view(self.n, 1))
        self.tok_logp.copy_(best.values)
        self.pos.add_(1)
        self.n_out.add_(1)

    @torch.no_grad()
    def decode(self, n_tokens: int):
        """Generate ``n_tokens`` more tokens (the first is the one ``prefill`` left pending).  Returns (ids, logp) of all
        tokens generated so far, shapes (n_hyp, total)."""
        if self.graph is None:
            self.dec._set_export(False)
            state = [t.clone() for t in (self.tok, self.tok_logp, self.pos, self.n_out, self.out_tok, self.out_logp)]
            kv = [t.clone() for t in self.kv]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._step()  # warm-up (lazy initialisation outside the capture)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._step()
            # the warm-up step and the capture changed nothing that matters, but restore the exact pre-step state anyway
            for dst, src in zip((self.tok, self.tok_logp, self.pos, self.n_out, self.out_tok, self.out_logp), state):
                dst.copy_(src)
            for dst, src in zip(self.kv, kv):
                dst.copy_(src)
        total = int(self.n_out.item()) + n_tokens
        if int(self.pos.item()) + n_tokens > self.max_len:
            raise ValueError("decoding past max_len")
        for _ in range(n_tokens):
            self.graph.replay()
        return self.out_tok[:, :total].clone(), self.out_logp[:, :total].clone()


class OpenAIWhisperDecoder(torch.nn.Module):
    """``forward`` returns ``(logits, att_maps)`` like the reference.  ``att_maps`` holds, per decoder layer
    ``>= src_layer-1``, the SELF-attention map the block returns (whisper/model.py:232,248):

      export_mode="full"    (B,H,T,T) per layer, stacked to (L_sel,B,H,T,T) — the reference's tensor, fp32 logits
                            with -inf above the diagonal (or probabilities with ``export_kind="probs"``)
      export_mode="compact" only key columns 1:3 (the <|zh|>, <|en|> prompt tokens, the only ones the guided loss
                            reads, espnet_model.py:506): (L_sel,B,H,T,2)
      export_mode="fused"   nothing is exported: the guided loss' reduction over the target positions runs in the
                            attention kernel's epilogue and ``att_maps`` is an ``ops.GuidedParts`` (L_sel,B,H,4,2) of
                            partial sums that ``ESPnetASRModel.calculate_cs_loss`` finishes.  Needs ``guided_pattern``
                            (set by ESPnetASRModel before the call), bf16 activations and T <= 128; longer targets and
                            fp32 runs use the compact export for that call.
    """

    def __init__(self, vocab_size: int, encoder_output_size: int, dropout_rate: float = 0.0,
                 whisper_model: str = "small", download_dir: Optional[str] = None, src_layer: int = 12,
                 whisper_cs: bool = False, pe_whisper: bool = False, adapter: bool = False, side_network: bool = False,
                 side_network_conf=None, c_val_attention: float = 0.6, estimate_c: bool = False,
                 export_mode: str = "full", export_kind: str = "logits", seed: int = 0, kv_cache: bool = False,
                 fused_loss: bool = False):
        super().__init__()
        if whisper_model not in W.available_models():  # whisper_decoder.py:55 asserts the same
            import os
            assert os.path.isfile(whisper_model), f"unknown whisper model {whisper_model}"
        _model = W.load_model(whisper_model, adapter, pe_whisper, side_network, side_network_conf,
                              download_root=download_dir, seed=seed)
        self.sidenetwork = side_network
        # forward() returns an ops.VocabLogits handle (padded logits in the GEMM's dtype) instead of the fp32 (B,T,V)
        # tensor; ESPnetASRModel feeds it to the fused label-smoothing loss / accuracy
        self.fused_loss = fused_loss
        self.kv_cache = kv_cache  # forward_one_step / score / batch_score keep per-hypothesis K/V states
        self.decoders = copy.deepcopy(_model.decoder)
        self.decoders.train()
        del _model
        attention_dim = self.decoders.token_embedding.embedding_dim
        self.dropout = torch.nn.Dropout(dropout_rate)
        if vocab_size != self.decoders.token_embedding.num_embeddings:  # whisper_decoder.py:66-79
            std, mean = torch.std_mean(self.decoders.token_embedding.weight)
            self.decoders.token_embedding = torch.nn.Embedding(vocab_size, attention_dim)
            torch.nn.init.normal_(self.decoders.token_embedding.weight, mean.item(), std.item())
        self.whisper_cs = whisper_cs
        self.src_layer = src_layer - 1
        self.att_map = None
        self.estimate_c = estimate_c
        self.c_val_attention = c_val_attention
        if estimate_c:
            self.decoders.estimated_c_val = torch.nn.Parameter(torch.Tensor([c_val_attention]))
        assert export_mode in ("full", "compact", "fused") and export_kind in ("logits", "probs")
        self.export_mode, self.export_kind = export_mode, export_kind
        self.guided_pattern: Optional[torch.Tensor] = None  # (B,T,2) target pattern of the current batch (export_mode "fused")
        self.n_early_layers = 2

    def _set_export(self, enabled: bool, mode: Optional[str] = None, guided: Optional[torch.Tensor] = None) -> None:
        mode = mode or self.export_mode
        for layer, block in enumerate(self.decoders.blocks):
            on = enabled and layer >= self.src_layer
            if on and mode == "fused":
                block.attn.export = None
                block.attn.guided = (guided, layer < self.n_early_layers)
            else:
                block.attn.guided = None
                block.attn.export = (self.export_kind, (1, 3) if mode == "compact" else None) if on else None

    def forward(self, hs_pad: torch.Tensor, hlens: torch.Tensor, ys_in_pad: torch.Tensor, ys_in_lens: torch.Tensor,
                side_encoder_output: torch.Tensor = None, memory_len: torch.Tensor = None) -> Tuple[torch.Tensor, Any]:
        """whisper_decoder.py:89-170.  hlens / ys_in_lens are ignored exactly as in the reference (no key padding).
        ``memory_len`` (device int32 scalar): encoder frames that really exist when ``hs_pad`` is zero-padded to a static
        bucket length (graphed.BucketedTrainStep) — the cross attention then sees exactly the unpadded memory."""
        dec = self.decoders
        tgt = dec.token_embedding(ys_in_pad) + dec.positional_embedding[: ys_in_pad.size(1)]
        x = self.dropout(tgt).to(hs_pad.dtype)
        mode = self.export_mode
        if mode == "fused":
            act = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
            usable = (self.guided_pattern is not None and act == torch.bfloat16 and 3 <= ys_in_pad.size(1) <= 128
                      and self.export_kind == "logits" and x.is_cuda and dec.blocks[0].attn._frozen())
            mode = "fused" if usable else "compact"
        self._set_export(self.whisper_cs, mode, self.guided_pattern)
        attention_scores: List[torch.Tensor] = []
        def collect(layer, attention_map):
            if self.whisper_cs and layer >= self.src_layer:
                attention_scores.append(attention_map)
        x = W.run_blocks(dec.blocks, x, dec.ln, between=self.dropout, between_changes_x=self.training and self.dropout.p > 0,
                         on_block=collect, xa=hs_pad, mask=dec.mask, xa_len=memory_len)
        logits = dec.vocab_logits(x, lazy=self.fused_loss)
        if self.whisper_cs:
            if mode == "fused":
                return logits, ops.GuidedParts.stack(attention_scores)
            return logits, torch.stack(attention_scores)
        return logits, attention_scores

    def forward_one_step(self, tgt: torch.Tensor, tgt_mask: torch.Tensor, memory: torch.Tensor,
                         cache: List[Any] = None, side_encoder_output: torch.Tensor = None,
                         return_maps: bool = False):
        """whisper_decoder.py:172-244: whole-prefix recompute, last position -> log_softmax; returns (logp, None).

        The reference dumps every layer's map to the CPU each step (:230); here maps stay on the device and are
        only produced when ``return_maps`` is set (then stored in ``self.att_map`` as a list of (n,H,t,t)).

        With ``kv_cache=True`` (constructor) the call is incremental instead: ``cache`` is None on the first call
        (the whole prefix is processed, the cross-attention K/V are projected once) or the state returned by the
        previous call, a list over layers of (self K, self V, cross K, cross V) with a leading batch dimension; only
        the last token of ``tgt`` is run and (logp, new state) is returned — O(t) per step instead of O(t^2), no
        device->host copies.  Same logits as the recompute path."""
        dec = self.decoders
        if tgt.size(1) > 448:
            tgt = tgt[:, :448]
        if self.kv_cache and not return_maps:
            return self._forward_one_step_cached(tgt, memory, cache)
        x = dec.token_embedding(tgt) + dec.positional_embedding[: tgt.size(1)]
        x = self.dropout(x).to(memory.dtype)
        self._set_export(return_maps, mode="full")
        maps = []
        last = len(dec.blocks) - 1
        for layer, block in enumerate(dec.blocks):
            x, att_map = block(x, memory, mask=dec.mask)
            if return_maps:
                maps.append(att_map)
            if layer < last:
                x = self.dropout(x)
        if return_maps:
            self.att_map = maps
        x = dec.ln(x)
        y = dec.vocab_logits(x[:, -1])
        return torch.log_softmax(y, dim=-1), None

    def _forward_one_step_cached(self, tgt: torch.Tensor, memory: torch.Tensor, cache: Optional[List[Any]]):
        dec = self.decoders
        t = tgt.size(1)
        self._set_export(False)
        if cache is None:  # prefill
            x = dec.token_embedding(tgt) + dec.positional_embedding[:t]
        else:
            if cache[0][0].size(1) != t - 1:
                raise ValueError("kv cache holds %d positions, the prefix has %d" % (cache[0][0].size(1), t))
            x = dec.token_embedding(tgt[:, -1:]) + dec.positional_embedding[t - 1: t]
        x = self.dropout(x).to(memory.dtype)
        new_cache = []
        last = len(dec.blocks) - 1
        for layer, block in enumerate(dec.blocks):
            x, st = block.step(x, memory, None if cache is None else cache[layer])
            new_cache.append(st)
            if layer < last:
                x = self.dropout(x)
        x = dec.ln(x[:, -1])
        return torch.log_softmax(dec.vocab_logits(x), dim=-1), new_cache

    def greedy_decoder(self, memory: torch.Tensor, max_len: int = 448) -> "GraphedGreedyDecoder":
        """A CUDA-graph greedy decoder over this module for the encoder output ``memory`` (n, Ta, D)."""
        return GraphedGreedyDecoder(self, memory, max_len)

    def score(self, ys, state, x):
        if self.kv_cache and state is not None:
            state = [tuple(t.unsqueeze(0) for t in layer) for layer in state]
        logp, state = self.forward_one_step(ys.unsqueeze(0), torch.empty(0), x.unsqueeze(0), cache=state)
        if state is not None:
            state = [tuple(t.squeeze(0) for t in layer) for layer in state]
        return logp.squeeze(0), state

    def batch_score(self, ys: torch.Tensor, states: List[Any], xs: torch.Tensor, x_enc: torch.Tensor = None):
        """ESPnet's BatchScorerInterface: per-hypothesis states (no batch dimension) are stacked, run, and split again
        (the reference keeps no state and returns None, whisper_decoder.py:246-273)."""
        cache = None
        if self.kv_cache and states and states[0] is not None:
            n_layer = len(states[0])
            cache = [tuple(torch.stack([st[layer][i] for st in states]) for i in range(4)) for layer in range(n_layer)]
        logp, new = self.forward_one_step(ys, torch.empty(0), xs, cache=cache, side_encoder_output=x_enc)
        if new is None:
            return logp, None
        return logp, [[tuple(t[b] for t in layer) for layer in new] for b in range(ys.size(0))]
