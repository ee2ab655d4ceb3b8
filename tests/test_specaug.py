"""SpecAug mirror (espnet2/asr/specaug/specaug.py, layers/time_warp.py, layers/mask_along_axis.py): the graph-safe
device-side variant gives the reference's interpolation for given random parameters and never synchronises."""
import torch
import torch.nn.functional as F

import aga_b200  # noqa: F401
from aga_b200 import specaug as S


def _reference_time_warp(x, center, warped):
    """espnet2/layers/time_warp.py:31-46 with the two random integers fixed."""
    xx = x[:, None]
    t = xx.shape[2]
    left = F.interpolate(xx[:, :, :center], (warped, xx.shape[3]), mode="bicubic", align_corners=False)
    right = F.interpolate(xx[:, :, center:], (t - warped, xx.shape[3]), mode="bicubic", align_corners=False)
    return torch.cat([left, right], dim=-2)[:, 0]


def test_time_warp_device_matches_bicubic_interpolate():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 80, 37, generator=g)
    window = 5
    for center in (window, 17, 40, 80 - window - 1):
        for warped in (center - window + 1, center, center + 1, center + window):
            got = S.time_warp_device(x, torch.tensor(center), torch.tensor(warped))
            torch.testing.assert_close(got, _reference_time_warp(x, center, warped), rtol=1e-4, atol=1e-4)


def test_graph_safe_specaug_shapes_and_masks():
    torch.manual_seed(1)
    conf = dict(apply_time_warp=True, time_warp_window=5, time_warp_mode="bicubic", apply_freq_mask=True,
                freq_mask_width_range=(0, 30), num_freq_mask=2, apply_time_mask=True, time_mask_width_range=(0, 40),
                num_time_mask=2)
    sa = S.SpecAug(graph_safe=True, **conf)
    x = torch.randn(4, 80, 3000) + 3.0
    y, lens = sa(x, torch.full((4,), 3000))
    assert y.shape == x.shape and lens.shape == (4,)
    zero_rows = (y == 0).all(dim=2).sum(dim=1)   # masked mel rows ("time" axis of the un-transposed call, width < 40 each)
    zero_cols = (y == 0).all(dim=1).sum(dim=1)   # masked frame columns (width < 30 each)
    assert int(zero_rows.max()) <= 2 * 39 and int(zero_cols.max()) <= 2 * 29
    # un-masked entries are the warped input: identical across the two variants for the same parameters
    m = S.mask_along_axis_device(x, (0, 30), dim=2, num_mask=2)
    assert m.shape == x.shape and ((m == 0) | (m == x)).all()
