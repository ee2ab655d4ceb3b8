"""GPU parity: tcgen05 attention (bf16) vs the fp64 oracle on the same bf16-rounded inputs (north-star: 2e-2)."""
import numpy as np
import pytest
import torch

import aga_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    import aga_b200
    return aga_b200


def _mk(B, Tq, Tk, H, amp, seed):
    g = torch.Generator().manual_seed(seed)
    q = (amp * torch.randn(B, Tq, H * 64, generator=g)).bfloat16()
    k = (amp * torch.randn(B, Tk, H * 64, generator=g)).bfloat16()
    v = torch.randn(B, Tk, H * 64, generator=g).bfloat16()
    return q, k, v


def _assert_close_scaled(got, ref, name, tol=2e-2):
    """North-star bf16 gate: 2e-2 RELATIVE TO THE TENSOR'S OWN SCALE.  With 1500 N(0,1) keys the attention output is an
    average of ~1500 values (|out| ~ 0.03): an absolute 2e-2 would accept a 60 % error there."""
    scale = max(float(np.abs(ref).max()), 1e-6)
    np.testing.assert_allclose(got.float().cpu().numpy(), ref, rtol=tol, atol=tol * scale, err_msg=name)


@pytest.mark.parametrize("B,H,Tq,Tk,amp", [
    (1, 1, 128, 128, 1.0), (1, 1, 1, 1, 1.0), (1, 2, 257, 129, 1.0), (2, 3, 300, 200, 1.0), (1, 2, 64, 1500, 1.0),
    (1, 12, 1500, 1500, 1.0), (1, 2, 512, 640, 4.0), (2, 2, 131, 131, 2.0), (1, 1, 448, 1500, 1.0),
    # the BASELINE cross-attention shape, and low-entropy (peaked) softmaxes whose outputs are O(1)
    (16, 12, 64, 1500, 1.0), (1, 2, 1500, 1500, 3.0), (2, 2, 64, 1500, 3.0)])
def test_tc_forward_vs_oracle(A, B, H, Tq, Tk, amp):
    q, k, v = _mk(B, Tq, Tk, H, amp, seed=Tq + Tk)
    out, lse, _ = A.qkv_attention(q.cuda(), k.cuda(), v.cuda(), H, impl="tcgen05")
    ref, qk, _ = O.qkv_attention(q.float().numpy(), k.float().numpy(), v.float().numpy(), H)
    _assert_close_scaled(out, ref, "out")
    mx = qk.max(-1)
    lse_ref = mx + np.log(np.exp(qk - mx[..., None]).sum(-1))
    np.testing.assert_allclose(lse.cpu().numpy(), lse_ref, rtol=1e-3, atol=1e-3)


def test_tc_matches_simt_and_auto_dispatch(A):
    """AUTO picks tcgen05 for bf16 non-causal; it must agree with the CUDA-core path to bf16 rounding."""
    q, k, v = _mk(2, 200, 333, 4, 1.0, seed=7)
    o_auto, l_auto, _ = A.qkv_attention(q.cuda(), k.cuda(), v.cuda(), 4)
    o_tc, l_tc, _ = A.qkv_attention(q.cuda(), k.cuda(), v.cuda(), 4, impl="tcgen05")
    o_si, l_si, _ = A.qkv_attention(q.cuda(), k.cuda(), v.cuda(), 4, impl="simt")
    assert torch.equal(o_auto, o_tc) and torch.equal(l_auto, l_tc)
    _assert_close_scaled(o_tc, o_si.float().cpu().numpy(), "tc vs simt")
    np.testing.assert_allclose(l_tc.cpu().numpy(), l_si.cpu().numpy(), rtol=1e-4, atol=1e-4)


def test_tc_full_size_properties(A):
    """BASELINE shape (B=16, H=12, 1500x1500): rows are convex combinations of V (bounded by V's range per
    column), constant V gives that constant, and batch entries are independent."""
    g = torch.Generator().manual_seed(0)
    B, H, T = 16, 12, 1500
    q = torch.randn(B, T, H * 64, generator=g).bfloat16().cuda()
    k = torch.randn(B, T, H * 64, generator=g).bfloat16().cuda()
    v = torch.randn(B, T, H * 64, generator=g).bfloat16().cuda()
    out, lse, _ = A.qkv_attention(q, k, v, H)
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    vmax = v.float().amax(dim=1, keepdim=True)
    vmin = v.float().amin(dim=1, keepdim=True)
    assert (out.float() <= vmax + 2e-2).all() and (out.float() >= vmin - 2e-2).all()
    ones = torch.full_like(v, 0.5)
    o1, _, _ = A.qkv_attention(q, k, ones, H)
    assert float((o1.float() - 0.5).abs().max()) < 1e-2
    o2, l2, _ = A.qkv_attention(q[3:5], k[3:5], v[3:5], H)
    assert torch.equal(o2, out[3:5]) and torch.equal(l2, lse[3:5])
    ref, _, _ = O.qkv_attention(q[7:8, :, :128].float().cpu().numpy(), k[7:8, :, :128].float().cpu().numpy(),
                                v[7:8, :, :128].float().cpu().numpy(), 2)
    _assert_close_scaled(out[7:8, :, :128], ref, "out[7]")


@pytest.mark.parametrize("B,H,Tq,Tk,amp", [
    (1, 1, 128, 128, 1.0), (1, 1, 1, 1, 1.0), (1, 2, 257, 129, 1.0), (2, 3, 300, 200, 1.0), (1, 2, 64, 1500, 1.0),
    (1, 2, 1500, 1500, 1.0), (1, 2, 512, 640, 3.0),
    # one query tile -> the Q-resident kernel (decoder cross attention): ragged rows / keys, several chunk CTAs
    (2, 3, 100, 1500, 1.0), (1, 2, 128, 700, 2.0), (3, 2, 5, 64, 1.0), (16, 12, 64, 1500, 1.0), (1, 1, 33, 129, 1.0)])
def test_tc_backward_vs_oracle(A, B, H, Tq, Tk, amp):
    """dQ/dK/dV of the 5-GEMM tcgen05 backward (dQ via fp32 red.add accumulation) vs the fp64 oracle: 2e-2 of the
    gradient's scale (north-star bf16 tolerance)."""
    q, k, v = _mk(B, Tq, Tk, H, amp, seed=Tq * 3 + Tk)
    g = torch.Generator().manual_seed(1)
    do = torch.randn(B, Tq, H * 64, generator=g).bfloat16()
    qd, kd, vd = (x.cuda().requires_grad_() for x in (q, k, v))
    out, _, _ = A.qkv_attention(qd, kd, vd, H, impl="tcgen05")
    out.backward(do.cuda())
    dq, dk, dv = O.qkv_attention_bwd(q.float().numpy(), k.float().numpy(), v.float().numpy(), H, False,
                                     do.float().numpy())
    for name, got, ref in (("dq", qd.grad, dq), ("dk", kd.grad, dk), ("dv", vd.grad, dv)):
        scale = max(float(np.abs(ref).max()), 1e-6)
        np.testing.assert_allclose(got.float().cpu().numpy(), ref, rtol=2e-2, atol=2e-2 * scale, err_msg=name)


def test_tc_backward_full_size_linearity(A):
    """BASELINE shape: the backward is linear in dO (size-independent property), and repeated runs agree to the
    fp32 atomics' reordering noise."""
    g = torch.Generator().manual_seed(0)
    B, H, T = 4, 12, 1500
    q, k, v = (torch.randn(B, T, H * 64, generator=g).bfloat16().cuda().requires_grad_() for _ in range(3))
    out, _, _ = A.qkv_attention(q, k, v, H)
    d1 = torch.randn(B, T, H * 64, generator=g).bfloat16().cuda()
    g1 = torch.autograd.grad(out, (q, k, v), d1, retain_graph=True)
    g2 = torch.autograd.grad(out, (q, k, v), d1 * 2, retain_graph=True)
    g1b = torch.autograd.grad(out, (q, k, v), d1, retain_graph=True)
    for a, b2, a2 in zip(g1, g2, g1b):
        assert torch.isfinite(a.float()).all()
        scale = float(a.float().abs().max())
        assert float((b2.float() - 2 * a.float()).abs().max()) <= 2e-2 * scale
        assert float((a2.float() - a.float()).abs().max()) <= 1e-2 * scale


@pytest.mark.parametrize("B,H,T,cols,sel", [(2, 3, 64, (1, 3), None), (1, 2, 128, (1, 3), [1, 0]), (2, 2, 37, (0, 5), None),
                                            (16, 12, 64, (1, 3), None), (1, 2, 300, (1, 3), None), (1, 1, 448, (120, 131), None)])
def test_tc_causal_export_forward_vs_oracle(A, B, H, T, cols, sel):
    """Decoder self attention on the tcgen05 path: causal mask (key tiles above the diagonal skipped) and the scaled,
    masked logits of the selected key columns written from the S registers (whisper/model.py:103-109)."""
    q, k, v = _mk(B, T, T, H, 1.0, seed=T + 11)
    head_sel = None if sel is None else torch.tensor(sel, dtype=torch.uint8)
    out, lse, slab = A.qkv_attention(q.cuda(), k.cuda(), v.cuda(), H, causal=True, export="logits", export_cols=cols,
                                     head_sel=head_sel, impl="tcgen05")
    ref, qk, _ = O.qkv_attention(q.float().numpy(), k.float().numpy(), v.float().numpy(), H, causal=True)
    _assert_close_scaled(out, ref, "out")
    want = qk[..., cols[0]:cols[1]]
    got = slab.cpu().numpy()
    for h in range(H):
        if sel is not None and not sel[h]:
            assert not got[:, h].any()  # unselected heads are never written (buffer starts as zeros)
            continue
        assert np.array_equal(np.isinf(got[:, h]), np.isinf(want[:, h]))
        fin = np.isfinite(want[:, h])
        np.testing.assert_allclose(got[:, h][fin], want[:, h][fin], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("B,H,T,cols,sel", [(2, 3, 64, (1, 3), None), (1, 2, 128, (1, 3), [0, 1]), (2, 2, 37, (0, 5), None),
                                            (16, 12, 64, (1, 3), None),
                                            # more than one query tile (the decoder's context is 448, whisper/model.py:322):
                                            # the persistent kernel with the causal mask and the exported columns' gradient
                                            (1, 2, 300, (1, 3), None), (1, 1, 448, (1, 3), None), (2, 2, 200, (0, 5), [1, 0]),
                                            (1, 2, 129, (1, 3), None)])
def test_tc_causal_export_backward_vs_oracle(A, B, H, T, cols, sel):
    """dQ/dK/dV of the causal tcgen05 kernels (one query tile: Q-resident kernel; more: persistent kernel) with a gradient
    on the exported columns (the guided loss' gradient, espnet_model.py:463-530) vs the fp64 oracle, and agreement with the
    CUDA-core path."""
    q, k, v = _mk(B, T, T, H, 1.0, seed=T * 5 + 1)
    g = torch.Generator().manual_seed(3)
    do = torch.randn(B, T, H * 64, generator=g).bfloat16()
    dE = torch.randn(B, H, T, cols[1] - cols[0], generator=g)
    head_sel = None if sel is None else torch.tensor(sel, dtype=torch.uint8)
    grads = {}
    for impl in ("tcgen05", "simt"):
        qd, kd, vd = (x.cuda().requires_grad_() for x in (q, k, v))
        out, _, slab = A.qkv_attention(qd, kd, vd, H, causal=True, export="logits", export_cols=cols, head_sel=head_sel,
                                       impl=impl)
        fin = torch.isfinite(slab)
        loss_e = (torch.where(fin, slab, torch.zeros_like(slab)) * dE.cuda()).sum()
        torch.autograd.backward([out, loss_e], [do.cuda(), torch.ones((), device="cuda")])
        grads[impl] = (qd.grad, kd.grad, vd.grad)
    d_qk = np.zeros((B, H, T, T))
    d_qk[..., cols[0]:cols[1]] = dE.numpy()
    if sel is not None:
        d_qk *= np.asarray(sel, dtype=np.float64)[None, :, None, None]
    d_qk *= np.tril(np.ones((T, T)))[None, None]
    dq, dk, dv = O.qkv_attention_bwd(q.float().numpy(), k.float().numpy(), v.float().numpy(), H, True, do.float().numpy(),
                                     d_qk=d_qk)
    for name, got, alt, ref in zip(("dq", "dk", "dv"), grads["tcgen05"], grads["simt"], (dq, dk, dv)):
        scale = max(float(np.abs(ref).max()), 1e-6)
        np.testing.assert_allclose(got.float().cpu().numpy(), ref, rtol=2e-2, atol=2e-2 * scale, err_msg=name)
        np.testing.assert_allclose(got.float().cpu().numpy(), alt.float().cpu().numpy(), rtol=2e-2, atol=2e-2 * scale,
                                   err_msg=name + " vs simt")


@pytest.mark.parametrize("B,H,Tq,Tk_true,Tk_pad", [(2, 3, 300, 700, 1000), (2, 2, 64, 900, 1500), (1, 2, 128, 129, 512),
                                                  (2, 2, 1000, 650, 1000), (1, 1, 64, 64, 1500)])
def test_tc_kv_len_equals_unpadded(A, B, H, Tq, Tk_true, Tk_pad):
    """Static launch shape for a shorter batch (graphed.BucketedTrainStep): K / V zero-padded... no, GARBAGE-padded to Tk_pad
    with the true key count as a DEVICE scalar.  Forward output / lse and dQ equal the unpadded call, dK / dV equal it on
    the real keys and are exactly zero on the padded ones (encoder self attention, decoder cross attention incl. the
    one-query-tile kernel, a partial last key tile, a single key tile)."""
    q, k, v = _mk(B, Tq, Tk_true, H, 1.0, seed=Tq + Tk_true)
    g = torch.Generator().manual_seed(2)
    kp = torch.randn(B, Tk_pad, H * 64, generator=g).bfloat16()
    vp = torch.randn(B, Tk_pad, H * 64, generator=g).bfloat16()
    kp[:, :Tk_true], vp[:, :Tk_true] = k, v
    do = torch.randn(B, Tq, H * 64, generator=g).bfloat16().cuda()
    kv = torch.tensor(Tk_true, dtype=torch.int32).cuda()
    res = {}
    for name, (kk, vv, kw) in {"ref": (k, v, {}), "pad": (kp, vp, {"kv_len": kv})}.items():
        qd, kd, vd = (x.cuda().requires_grad_() for x in (q, kk, vv))
        out, lse, _ = A.qkv_attention(qd, kd, vd, H, impl="tcgen05", **kw)
        out.backward(do)
        res[name] = (out.detach(), lse, qd.grad, kd.grad, vd.grad)
    o_r, l_r, dq_r, dk_r, dv_r = res["ref"]
    o_p, l_p, dq_p, dk_p, dv_p = res["pad"]
    assert torch.equal(o_p, o_r) and torch.equal(l_p, l_r)
    for a, b2 in ((dq_p, dq_r), (dk_p[:, :Tk_true], dk_r), (dv_p[:, :Tk_true], dv_r)):
        scale = float(b2.float().abs().max())
        assert float((a.float() - b2.float()).abs().max()) <= 1e-2 * scale  # fp32 atomics order only
    assert not dk_p[:, Tk_true:].any() and not dv_p[:, Tk_true:].any()
