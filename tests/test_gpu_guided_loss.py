"""GPU parity: language pattern, guided (cs) loss + gradient, head vote — against reference goldens and the oracle."""
import os

import numpy as np
import pytest
import torch

import aga_oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LID = os.path.join(ROOT, "attention-guided-adaptation-for-code-switching-speech-recognition_b200", "data",
                   "lid_table_multilingual.u8")


@pytest.fixture(scope="module")
def A():
    import aga_b200
    return aga_b200


def test_pattern_loss_grad_vs_reference_golden(A, golden_dir):
    g = np.load(os.path.join(golden_dir, "cs_loss.npz"))
    lid = torch.from_numpy(np.fromfile(LID, dtype=np.uint8))
    toks = torch.from_numpy(g["tokens"]).cuda()
    pat = A.attention_pattern(toks, lid, 0.6)
    assert np.array_equal(pat.cpu().numpy(), g["pattern"])  # bit-exact
    maps = torch.from_numpy(g["maps"]).cuda().requires_grad_()
    mask = torch.from_numpy(O.literal_head_mask())
    loss = A.guided_loss(maps[..., 1:3], pat, mask, n_early=2)
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=1e-5)
    loss.backward()
    np.testing.assert_allclose(maps.grad.cpu().numpy(), g["dmaps"], rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize("L,B,H,T,n_early", [(12, 16, 12, 64, 2), (4, 3, 6, 448, 2), (24, 2, 16, 17, 3), (2, 1, 1, 6, 0)])
def test_guided_loss_vs_oracle(A, L, B, H, T, n_early):
    rng = np.random.default_rng(L * 100 + T)
    lid_np = np.fromfile(LID, dtype=np.uint8)
    toks = np.full((B, T), 50257, dtype=np.int64)
    toks[:, :5] = [50258, 50260, 50259, 50359, 50363]
    for b in range(B):
        n = int(rng.integers(0, max(1, T - 5)))
        toks[b, 5:5 + n] = rng.integers(0, 50257, n)
    pat_ref = np.stack([O.create_attention_pattern(t, lid_np, 0.6) for t in toks])
    pat = A.attention_pattern(torch.from_numpy(toks).cuda(), torch.from_numpy(lid_np), 0.6)
    assert np.array_equal(pat.cpu().numpy(), pat_ref)
    slab = rng.standard_normal((L, B, H, T, 2)).astype(np.float32)
    slab[:, :, :, 0, :] = -np.inf
    slab[:, :, :, 1, 1] = -np.inf  # causal mask on columns 1,2
    mask = (rng.random((L, H)) < 0.5).astype(np.float32)
    s = torch.from_numpy(slab).cuda().requires_grad_()
    loss = A.guided_loss(s, pat, torch.from_numpy(mask), n_early=n_early)
    ref, gref = O.calculate_cs_loss(slab, pat_ref, mask, n_early=n_early, want_grad=True)
    if np.isnan(ref):
        assert np.isnan(loss.item())  # 0/0 mirrors the reference's count_nonzero division
        return
    np.testing.assert_allclose(loss.item(), ref, rtol=2e-5)
    (loss * 3.0).backward()
    np.testing.assert_allclose(s.grad.cpu().numpy(), 3.0 * gref, rtol=1e-4, atol=1e-8)


def test_head_vote_vs_reference_golden(A, golden_dir):
    g = np.load(os.path.join(golden_dir, "cs_loss.npz"))
    dec, cnt = A.head_vote(torch.from_numpy(g["probs"]).cuda())
    assert np.array_equal(cnt.cpu().numpy(), g["vote_counts"])  # bit-exact decisions
    s1, s2 = O.head_vote_sums(g["probs"])
    assert np.array_equal(dec.cpu().numpy().astype(bool), s1 > s2)


def test_head_selection_end_to_end(A):
    """probs export of the attention kernel -> vote -> counts -> selected-head mask (top-K, stable)."""
    rng = np.random.default_rng(0)
    L, B, H, T = 3, 4, 4, 20
    maps = []
    for l in range(L):
        q, k, v = (torch.from_numpy(rng.standard_normal((B, T, H * 64)).astype(np.float32)).cuda() for _ in range(3))
        k[:, 1:3] *= 1.0 + l  # make columns 1,2 progressively more attractive
        _, _, w = A.qkv_attention(q, k * 2.0, v, H, causal=True, export="probs", impl="simt")
        maps.append(w)
    probs = torch.stack(maps)
    dec, cnt = A.head_vote(probs)
    ref = O.new_check_attention_language(probs.cpu().numpy())
    assert np.array_equal(cnt.cpu().numpy(), ref)


@pytest.mark.parametrize("B,H,T,early", [(2, 3, 64, False), (3, 2, 37, False), (2, 2, 128, True), (1, 12, 5, False)])
def test_guided_loss_fused_into_attention_epilogue(A, B, H, T, early):
    """N1 (north star: "the guided-loss reduction is fused into the same epilogue"): the decoder self-attention kernel
    reduces sum_t r_t and the non-zero count per (utterance, head) itself — no slab is exported.  Loss and the gradient
    that reaches the packed QKV projection equal the two-kernel path (export columns 1:3 -> aga_guided_loss_fwd_bwd) and
    the oracle (espnet_model.py:463-530), pad rows / causal zeros / early layers included."""
    from aga_b200 import ops
    g = torch.Generator().manual_seed(T * 7 + H)
    x = torch.randn(B, T, 3 * H * 64, generator=g).bfloat16().cuda()
    toks = torch.full((B, T), 50257, dtype=torch.long)
    toks[:, :5] = torch.tensor([50258, 50260, 50259, 50359, 50363])
    for b in range(B):
        n = min(T - 5, 3 + 7 * b)
        toks[b, 5:5 + n] = torch.randint(0, 50257, (n,), generator=g)
    lid = torch.from_numpy(np.fromfile(LID, dtype=np.uint8))
    pat = A.attention_pattern(toks.cuda(), lid, 0.6)
    mask = (torch.rand(1, H, generator=g) < 0.7).float().cuda()
    n_early = 1 if early else 0
    do = torch.randn(B, T, H * 64, generator=g).bfloat16().cuda()
    res = {}
    for mode in ("fused", "slab"):
        xd = x.clone().requires_grad_()
        if mode == "fused":
            out, _, parts = ops.qkv_attention_packed(xd, H, causal=True, guided=(pat, early), impl="tcgen05")
            assert parts.t.shape == (B, H, 4, 2)
            loss = ops.guided_loss_from_parts(ops.GuidedParts(parts.t[None]), mask)
        else:
            out, _, slab = ops.qkv_attention_packed(xd, H, causal=True, export="logits", export_cols=(1, 3), impl="tcgen05")
            loss = A.guided_loss(slab[None], pat, mask, n_early=n_early)
        torch.autograd.backward([out, loss * 3.0], [do, torch.ones((), device="cuda")])
        res[mode] = (float(loss), xd.grad.float(), slab.detach() if mode == "slab" else None)
    want = O.calculate_cs_loss(res["slab"][2].cpu().numpy()[None], pat.cpu().numpy(), mask.cpu().numpy(), n_early=n_early)
    np.testing.assert_allclose(res["fused"][0], want, rtol=1e-5)
    np.testing.assert_allclose(res["fused"][0], res["slab"][0], rtol=1e-5)
    scale = float(res["slab"][1].abs().max())
    torch.testing.assert_close(res["fused"][1], res["slab"][1], rtol=2e-2, atol=2e-3 * scale)
