"""GPU parity: CUDA-core attention (fp32 and bf16 inputs), column export, backward incl. export gradient."""
import os

import numpy as np
import pytest
import torch

import aga_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    import aga_b200
    return aga_b200


def _t(x, dtype=torch.float32, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda().to(dtype)
    return t.requires_grad_() if grad else t


@pytest.mark.parametrize("case", ["self_causal", "cross", "enc_self"])
def test_attention_fp32_vs_reference_golden(A, golden_dir, case):
    g = np.load(os.path.join(golden_dir, "attention.npz"))
    B, H, Tq, Tk, causal = [int(x) for x in g[f"{case}_cfg"]]
    q, k, v = (_t(g[f"{case}_{n}"], grad=True) for n in "qkv")
    out, lse, qk = A.qkv_attention(q, k, v, H, causal=bool(causal), export="logits", impl="simt")
    ref_qk = g[f"{case}_qk"]
    got_qk = qk.detach().cpu().numpy()
    assert got_qk.shape == ref_qk.shape
    assert np.array_equal(np.isinf(got_qk), np.isinf(ref_qk))
    fin = np.isfinite(ref_qk)
    np.testing.assert_allclose(got_qk[fin], ref_qk[fin], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g[f"{case}_out"], rtol=1e-4, atol=1e-5)
    # lse vs oracle
    _, qk64, _ = O.qkv_attention(g[f"{case}_q"], g[f"{case}_k"], g[f"{case}_v"], H, bool(causal))
    mx = qk64.max(-1)
    lse_ref = mx + np.log(np.exp(qk64 - mx[..., None]).sum(-1))
    np.testing.assert_allclose(lse.cpu().numpy(), lse_ref, rtol=1e-5, atol=1e-5)
    # backward with the same upstream gradients the reference saw (dout and a gradient on columns 1:3 of qk)
    dqk = _t(g[f"{case}_dqk"])
    fin_t = torch.isfinite(qk)
    loss = (out * _t(g[f"{case}_dout"])).sum() + (torch.where(fin_t, qk, torch.zeros_like(qk)) * dqk).sum()
    loss.backward()
    for n, t in (("dq", q), ("dk", k), ("dv", v)):
        np.testing.assert_allclose(t.grad.cpu().numpy(), g[f"{case}_{n}"], rtol=2e-4, atol=2e-5, err_msg=n)


@pytest.mark.parametrize("B,H,Tq,Tk,causal", [(2, 2, 64, 64, True), (1, 3, 130, 130, True), (2, 2, 7, 131, False),
                                              (1, 1, 1, 1, True), (1, 2, 65, 200, False), (1, 1, 448, 448, True)])
def test_attention_fp32_compact_export_and_grads(A, B, H, Tq, Tk, causal):
    rng = np.random.default_rng(Tq * 1000 + Tk)
    D = H * 64
    qn, kn, vn = (rng.standard_normal((B, t, D)).astype(np.float32) for t in (Tq, Tk, Tk))
    lo, hi = (1, 3) if Tk >= 3 else (0, 1)
    sel = np.array([(h % 2) == 0 for h in range(H)], dtype=np.uint8)
    q, k, v = (_t(x, grad=True) for x in (qn, kn, vn))
    out, lse, slab = A.qkv_attention(q, k, v, H, causal=causal, export="logits", export_cols=(lo, hi),
                                     head_sel=torch.from_numpy(sel), impl="simt")
    o_ref, qk_ref, _ = O.qkv_attention(qn, kn, vn, H, causal)
    np.testing.assert_allclose(out.detach().cpu().numpy(), o_ref, rtol=1e-4, atol=2e-5)
    s_ref = qk_ref[..., lo:hi].copy()
    s_ref[:, sel == 0] = 0.0  # unselected heads are not exported (buffer defined as zero)
    s = slab.detach().cpu().numpy()
    assert np.array_equal(np.isinf(s), np.isinf(s_ref))
    np.testing.assert_allclose(np.where(np.isinf(s_ref), 0, s), np.where(np.isinf(s_ref), 0, s_ref), rtol=1e-4, atol=2e-5)
    dout = rng.standard_normal((B, Tq, D)).astype(np.float32)
    dsl = rng.standard_normal(s.shape).astype(np.float32)
    fin = torch.isfinite(slab)
    ((out * _t(dout)).sum() + (torch.where(fin, slab, torch.zeros_like(slab)) * _t(dsl)).sum()).backward()
    d_qk = np.zeros_like(qk_ref)
    d_qk[..., lo:hi] = dsl * (sel[None, :, None, None] != 0)
    dq, dk, dv = O.qkv_attention_bwd(qn, kn, vn, H, causal, dout, d_qk=d_qk)
    np.testing.assert_allclose(q.grad.cpu().numpy(), dq, rtol=3e-4, atol=5e-5)
    np.testing.assert_allclose(k.grad.cpu().numpy(), dk, rtol=3e-4, atol=5e-5)
    np.testing.assert_allclose(v.grad.cpu().numpy(), dv, rtol=3e-4, atol=5e-5)


def test_attention_probs_export_and_grads(A):
    rng = np.random.default_rng(5)
    B, H, T = 2, 2, 40
    qn, kn, vn = (rng.standard_normal((B, T, H * 64)).astype(np.float32) for _ in range(3))
    q, k, v = (_t(x, grad=True) for x in (qn, kn, vn))
    out, _, w = A.qkv_attention(q, k, v, H, causal=True, export="probs", impl="simt")
    o_ref, _, w_ref = O.qkv_attention(qn, kn, vn, H, True)
    np.testing.assert_allclose(w.detach().cpu().numpy(), w_ref, rtol=1e-4, atol=1e-6)
    dout = rng.standard_normal((B, T, H * 64)).astype(np.float32)
    dw = rng.standard_normal(w_ref.shape).astype(np.float32)
    ((out * _t(dout)).sum() + (w * _t(dw)).sum()).backward()
    dq, dk, dv = O.qkv_attention_bwd(qn, kn, vn, H, True, dout, d_w=dw)
    np.testing.assert_allclose(q.grad.cpu().numpy(), dq, rtol=3e-4, atol=5e-5)
    np.testing.assert_allclose(k.grad.cpu().numpy(), dk, rtol=3e-4, atol=5e-5)
    np.testing.assert_allclose(v.grad.cpu().numpy(), dv, rtol=3e-4, atol=5e-5)


def test_attention_bf16_inputs_simt(A):
    """bf16 in/out on the CUDA-core path: within the north-star 2e-2 of the fp32 oracle on the same rounded inputs."""
    rng = np.random.default_rng(9)
    B, H, Tq, Tk = 2, 3, 70, 150
    qn, kn, vn = (rng.standard_normal((B, t, H * 64)).astype(np.float32) for t in (Tq, Tk, Tk))
    q, k, v = (_t(x, torch.bfloat16, grad=True) for x in (qn, kn, vn))
    out, _, _ = A.qkv_attention(q, k, v, H, impl="simt")
    r = lambda t: t.detach().float().cpu().numpy()
    o_ref, _, _ = O.qkv_attention(r(q), r(k), r(v), H, False)
    np.testing.assert_allclose(r(out), o_ref, rtol=2e-2, atol=2e-2 * float(np.abs(o_ref).max()))
    dout = rng.standard_normal((B, Tq, H * 64)).astype(np.float32)
    (out.float() * _t(dout)).sum().backward()
    dq, dk, dv = O.qkv_attention_bwd(r(q), r(k), r(v), H, False, dout)
    for got, ref in ((q.grad, dq), (k.grad, dk), (v.grad, dv)):
        np.testing.assert_allclose(r(got), ref, rtol=2e-2, atol=2e-2 * float(np.abs(ref).max()))


def test_attention_strided_qkv_views(A):
    """q/k/v as column slices of one fused (B,T,3D) projection: strides, not copies."""
    rng = np.random.default_rng(3)
    B, H, T = 2, 2, 33
    D = H * 64
    qkv = _t(rng.standard_normal((B, T, 3 * D)).astype(np.float32))
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    out, _, _ = A.qkv_attention(q, k, v, H, impl="simt")
    o_ref, _, _ = O.qkv_attention(q.cpu().numpy(), k.cpu().numpy(), v.cpu().numpy(), H, False)
    np.testing.assert_allclose(out.cpu().numpy(), o_ref, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("B,H,Tk,kv_len", [(1, 12, 1500, None), (3, 2, 1, None), (2, 3, 33, None), (1, 12, 128, 7),
                                           (2, 2, 448, 300), (1, 1, 200, 200), (5, 6, 131, 1)])
def test_single_query_decode_kernel_vs_oracle(A, dtype, tol, B, H, Tk, kv_len):
    """Tq = 1 (a KV-cached decoding step) takes the cluster-split single-query kernel: same out / lse as the oracle on the
    keys that exist, through packed [K | V] cache rows and a q that is a column slice of a packed projection; its
    gradients (the ordinary backward kernels work from out and lse) match as well."""
    rng = np.random.default_rng(Tk * 7 + B)
    D = H * 64
    qkv = torch.from_numpy(rng.standard_normal((B, 1, 3 * D)).astype(np.float32)).cuda().to(dtype)
    cache = torch.from_numpy(rng.standard_normal((B, Tk, 2 * D)).astype(np.float32)).cuda().to(dtype)
    q, k, v = qkv[..., :D], cache[..., :D], cache[..., D:]
    n = Tk if kv_len is None else kv_len
    kl = None if kv_len is None else torch.tensor(kv_len, dtype=torch.int32, device="cuda")
    out, lse, _ = A.qkv_attention(q, k, v, H, kv_len=kl)
    ref_out, ref_qk, _ = O.qkv_attention(q.float().cpu().numpy(), k[:, :n].float().cpu().numpy(), v[:, :n].float().cpu().numpy(), H, False)
    mx = ref_qk.max(-1)
    ref_lse = mx + np.log(np.exp(ref_qk - mx[..., None]).sum(-1))
    np.testing.assert_allclose(out.float().cpu().numpy(), ref_out, rtol=tol, atol=tol)
    np.testing.assert_allclose(lse.cpu().numpy(), ref_lse, rtol=1e-5, atol=1e-5)
    if kv_len is None:  # same numbers as the query-tile kernel of the same dtype on a 2-row problem's first row
        q2 = torch.cat([q, q], dim=1).contiguous()
        out2, lse2, _ = A.qkv_attention(q2, k, v, H, impl="simt")
        torch.testing.assert_close(out.float(), out2[:, :1].float(), rtol=tol, atol=tol)
        torch.testing.assert_close(lse, lse2[..., :1], rtol=1e-5, atol=1e-5)
        qg, kg, vg = (t.detach().clone().contiguous().requires_grad_() for t in (q, k, v))
        o, _, _ = A.qkv_attention(qg, kg, vg, H)
        do = torch.from_numpy(rng.standard_normal((B, 1, D)).astype(np.float32)).cuda().to(dtype)
        o.backward(do)
        q2g, k2g, v2g = (t.detach().clone().contiguous().requires_grad_() for t in (q2, k, v))
        o2, _, _ = A.qkv_attention(q2g, k2g, v2g, H, impl="simt")
        o2.backward(torch.cat([do, torch.zeros_like(do)], dim=1))
        gtol = 1e-4 if dtype == torch.float32 else 2e-2
        torch.testing.assert_close(qg.grad.float(), q2g.grad[:, :1].float(), rtol=gtol, atol=gtol)
        torch.testing.assert_close(kg.grad.float(), k2g.grad.float(), rtol=gtol, atol=gtol)
        torch.testing.assert_close(vg.grad.float(), v2g.grad.float(), rtol=gtol, atol=gtol)
