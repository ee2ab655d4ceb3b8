"""SpecAug as the Whisper encoder applies it (adjacent to the hot path): plain PyTorch.

Mirrors espnet2/asr/specaug/specaug.py + espnet2/layers/{time_warp,mask_along_axis}.py for the options the SEAME
recipe uses (bicubic time warp, width-range masks).  NB: the encoder calls it on (B, n_mels, T) without a
transpose (whisper_encoder.py:235-236), so what SpecAug calls "time" (dim 1) is the mel axis and "freq" (dim 2)
is the frame axis — reproduced, not fixed.
"""
from typing import Optional, Sequence, Union

import torch


def time_warp(x: torch.Tensor, window: int = 80, mode: str = "bicubic") -> torch.Tensor:
    """x (B, T, F): warp along dim 1 around a random centre (espnet2/layers/time_warp.py:12-50)."""
    org = x.size()
    if x.dim() == 3:
        x = x[:, None]
    t = x.shape[2]
    if t - window <= window:
        return x.view(*org)
    center = torch.randint(window, t - window, (1,))[0]
    warped = torch.randint(center - window, center + window, (1,))[0] + 1
    left = torch.nn.functional.interpolate(x[:, :, :center], (warped, x.shape[3]), mode=mode, align_corners=False)
    right = torch.nn.functional.interpolate(x[:, :, center:], (t - warped, x.shape[3]), mode=mode, align_corners=False)
    if x.requires_grad:
        x = torch.cat([left, right], dim=-2)
    else:
        x[:, :, :warped] = left
        x[:, :, warped:] = right
    return x.view(*org)


def mask_along_axis(spec: torch.Tensor, mask_width_range: Sequence[int], dim: int, num_mask: int) -> torch.Tensor:
    """Zero `num_mask` random bands per utterance along `dim` (espnet2/layers/mask_along_axis.py:10-60)."""
    org = spec.size()
    if spec.dim() == 4:
        spec = spec.view(-1, spec.size(2), spec.size(3))
    B, D = spec.shape[0], spec.shape[dim]
    length = torch.randint(mask_width_range[0], mask_width_range[1], (B, num_mask), device=spec.device).unsqueeze(2)
    pos = torch.randint(0, max(1, D - int(length.max())), (B, num_mask), device=spec.device).unsqueeze(2)
    ar = torch.arange(D, device=spec.device)[None, None, :]
    mask = ((pos <= ar) * (ar < (pos + length))).any(dim=1)
    mask = mask.unsqueeze(2) if dim == 1 else mask.unsqueeze(1)
    value = 0.0
    spec = spec.masked_fill(mask, value)
    return spec.view(*org)


def _cubic_weights(t: torch.Tensor, A: float = -0.75):
    """The four cubic-convolution coefficients torch's bicubic interpolation uses (A = -0.75) for a fractional offset t."""
    def near(x):   # |x| <= 1
        return ((A + 2) * x - (A + 3)) * x * x + 1
    def far(x):    # 1 < |x| < 2
        return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A
    return torch.stack([far(t + 1), near(t), near(1 - t), far(2 - t)], dim=-1)


def time_warp_device(x: torch.Tensor, center: torch.Tensor, warped: torch.Tensor) -> torch.Tensor:
    """``time_warp`` with the random centre / warped position given as 0-d DEVICE integer tensors: no host
    synchronisation and static shapes, so the op is CUDA-graph capturable.  Same values as the two
    ``interpolate(mode="bicubic", align_corners=False)`` calls of espnet2/layers/time_warp.py:31-46 — rows [0, center)
    resampled to [0, warped), rows [center, t) to [warped, t) — written as ONE four-tap gather along dim 1."""
    t = x.shape[1]
    dst = torch.arange(t, device=x.device)
    c = center.to(torch.float32)
    w = warped.to(torch.float32)
    left = dst < warped
    # source coordinate inside the own part (left: [0, center), right: [center, t)), align_corners=False
    in_size = torch.where(left, c, t - c)
    out_size = torch.where(left, w, t - w)
    local = torch.where(left, dst, dst - warped).to(torch.float32)
    src = (local + 0.5) * (in_size / out_size) - 0.5
    x0 = torch.floor(src)
    frac = src - x0
    taps = x0.to(torch.int64)[:, None] + torch.arange(-1, 3, device=x.device)[None, :]      # (t, 4), local indices
    hi = (in_size - 1).to(torch.int64)[:, None]
    taps = torch.minimum(torch.clamp(taps, min=0), hi) + torch.where(left, 0, center)[:, None]  # clamp inside the part
    wts = _cubic_weights(frac).to(x.dtype)                                                   # (t, 4)
    g = x[:, taps.reshape(-1)].reshape(x.shape[0], t, 4, *x.shape[2:])                        # (B, t, 4, F)
    shape = (1, t, 4) + (1,) * (x.dim() - 2)
    return (g * wts.reshape(shape)).sum(dim=2)


def mask_along_axis_device(spec: torch.Tensor, mask_width_range: Sequence[int], dim: int, num_mask: int) -> torch.Tensor:
    """``mask_along_axis`` without the host read of ``length.max()`` (mask_along_axis.py:33): the position range is
    computed on the device, so the op is CUDA-graph capturable.  Same distribution of widths and positions."""
    B, D = spec.shape[0], spec.shape[dim]
    length = torch.randint(mask_width_range[0], mask_width_range[1], (B, num_mask), device=spec.device)
    span = torch.clamp(D - length.max(), min=1)
    pos = torch.floor(torch.rand((B, num_mask), device=spec.device) * span).to(torch.int64)
    ar = torch.arange(D, device=spec.device)[None, None, :]
    mask = ((pos.unsqueeze(2) <= ar) & (ar < (pos + length).unsqueeze(2))).any(dim=1)
    mask = mask.unsqueeze(2) if dim == 1 else mask.unsqueeze(1)
    return spec.masked_fill(mask, 0.0)


class SpecAug(torch.nn.Module):
    def __init__(self, apply_time_warp: bool = True, time_warp_window: int = 5, time_warp_mode: str = "bicubic",
                 apply_freq_mask: bool = True, freq_mask_width_range: Union[int, Sequence[int]] = (0, 20),
                 num_freq_mask: int = 2, apply_time_mask: bool = True,
                 time_mask_width_range: Optional[Union[int, Sequence[int]]] = None,
                 time_mask_width_ratio_range=None, num_time_mask: int = 2, graph_safe: bool = False):
        super().__init__()
        # graph_safe: random parameters are drawn and consumed on the device (no host sync, static shapes), so the
        # augmentation can be captured in the CUDA graph of the training step
        self.graph_safe = graph_safe
        if not (apply_time_warp or apply_time_mask or apply_freq_mask):
            raise ValueError("Either one of time_warp, time_mask, or freq_mask should be applied")
        if time_mask_width_ratio_range is not None:
            raise NotImplementedError("ratio-range time masks are not used by the SEAME recipe")
        as_range = lambda r: (0, r) if isinstance(r, int) else tuple(r)
        self.apply_time_warp, self.window, self.mode = apply_time_warp, time_warp_window, time_warp_mode
        self.freq = (as_range(freq_mask_width_range), num_freq_mask) if apply_freq_mask else None
        self.time = (as_range(time_mask_width_range), num_time_mask) if apply_time_mask else None

    def forward(self, x, x_lengths=None):
        if self.graph_safe:
            return self._forward_device(x, x_lengths)
        if self.apply_time_warp:
            x = time_warp(x, self.window, self.mode)
        if self.freq is not None:
            x = mask_along_axis(x, self.freq[0], dim=2, num_mask=self.freq[1])
        if self.time is not None:
            x = mask_along_axis(x, self.time[0], dim=1, num_mask=self.time[1])
        return x, x_lengths

    def _forward_device(self, x, x_lengths=None):
        t = x.shape[1]
        if self.apply_time_warp and t - self.window > self.window:
            center = torch.randint(self.window, t - self.window, (), device=x.device)
            warped = center - self.window + 1 + torch.randint(0, 2 * self.window, (), device=x.device)
            x = time_warp_device(x, center, warped)
        if self.freq is not None:
            x = mask_along_axis_device(x, self.freq[0], dim=2, num_mask=self.freq[1])
        if self.time is not None:
            x = mask_along_axis_device(x, self.time[0], dim=1, num_mask=self.time[1])
        return x, x_lengths
