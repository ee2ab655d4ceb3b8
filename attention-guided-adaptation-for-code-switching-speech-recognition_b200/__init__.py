"""aga_b200 — B200-native hot path for Attention-Guided Adaptation (Whisper code-switching ASR).

Import name: ``aga_b200`` (the repo-root shim ``aga_b200.py`` maps it onto this directory, whose on-disk
name carries hyphens).  See DESIGN.md for the path, its boundary and the kernels.
"""
from . import _lib
from ._lib import AgaError, launch_count
from .ops import (adapter_layer_norm, attention_pattern, guided_loss, head_vote, layer_norm, log_mel_spectrogram, mel_filterbank_numpy, mel_filters,
                  qkv_attention, qkv_attention_packed)

__all__ = ["AgaError", "adapter_layer_norm", "launch_count", "attention_pattern", "guided_loss", "head_vote", "layer_norm", "log_mel_spectrogram",
           "mel_filterbank_numpy", "mel_filters", "qkv_attention", "qkv_attention_packed"]
