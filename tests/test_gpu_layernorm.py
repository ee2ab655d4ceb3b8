"""GPU parity of the fused LayerNorm (whisper/whisper/model.py:30-32) against the oracle and the reference expression."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import aga_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    import aga_b200
    return aga_b200


def _case(rows_shape, D, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*rows_shape, D, generator=g) * 1.7 + 0.3
    w = 1.0 + 0.1 * torch.randn(D, generator=g)
    b = 0.1 * torch.randn(D, generator=g)
    dy = torch.randn(*rows_shape, D, generator=g)
    return x, w, b, dy


@pytest.mark.parametrize("shape,D", [((2, 37), 384), ((3, 50), 512), ((16, 1500), 768), ((2, 64), 1024), ((1, 9), 1280)])
def test_layernorm_fp32_vs_oracle(A, shape, D):
    x, w, b, dy = _case(shape, D, D)
    xd = x.cuda().requires_grad_()
    wd, bd = w.cuda().requires_grad_(), b.cuda().requires_grad_()
    y = A.layer_norm(xd, wd, bd, 1e-5)
    y.backward(dy.cuda())
    ref = O.layer_norm(x.numpy(), w.numpy(), b.numpy())
    dx, dg, db = O.layer_norm_bwd(dy.numpy(), x.numpy(), w.numpy())
    np.testing.assert_allclose(y.detach().cpu().numpy(), ref, rtol=1e-4, atol=1e-5)  # fp32 tolerance of the north star
    np.testing.assert_allclose(xd.grad.cpu().numpy(), dx, rtol=1e-4, atol=1e-5)
    scale = max(1.0, float(np.abs(dg).max()))
    np.testing.assert_allclose(wd.grad.cpu().numpy(), dg, rtol=1e-4, atol=1e-4 * scale)
    np.testing.assert_allclose(bd.grad.cpu().numpy(), db, rtol=1e-4, atol=1e-4 * scale)


@pytest.mark.parametrize("shape,D", [((16, 1500), 768), ((4, 64), 768), ((2, 100), 1280)])
def test_layernorm_bf16_matches_reference_expression(A, shape, D):
    """bf16 rows: same result as the reference's up-cast / F.layer_norm / down-cast chain on the GPU (one bf16 rounding)."""
    x, w, b, dy = _case(shape, D, 7 + D)
    xb = x.bfloat16().cuda()
    x1 = xb.clone().requires_grad_()
    w1, b1 = w.cuda().requires_grad_(), b.cuda().requires_grad_()
    y1 = A.layer_norm(x1, w1, b1, 1e-5)
    assert y1.dtype == torch.bfloat16
    y1.backward(dy.bfloat16().cuda())
    x2 = xb.clone().requires_grad_()
    w2, b2 = w.cuda().requires_grad_(), b.cuda().requires_grad_()
    y2 = F.layer_norm(x2.float(), (D,), w2, b2, 1e-5).type(torch.bfloat16)
    y2.backward(dy.bfloat16().cuda())
    torch.testing.assert_close(y1.float(), y2.float(), rtol=2e-2, atol=2e-2)
    assert (y1 == y2).float().mean() > 0.99  # identical up to rare last-bit ties of the final rounding
    torch.testing.assert_close(x1.grad.float(), x2.grad.float(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(w1.grad, w2.grad, rtol=2e-3, atol=2e-3 * float(w2.grad.abs().max()))
    torch.testing.assert_close(b1.grad, b2.grad, rtol=2e-3, atol=2e-3 * float(b2.grad.abs().max()))


def test_layernorm_frozen_params_and_errors(A):
    x, w, b, dy = _case((4, 10), 768, 1)
    xd = x.cuda().requires_grad_()
    y = A.layer_norm(xd, w.cuda(), b.cuda(), 1e-5)  # frozen gamma / beta: no parameter-gradient pass
    y.backward(dy.cuda())
    dx, _, _ = O.layer_norm_bwd(dy.numpy(), x.numpy(), w.numpy())
    np.testing.assert_allclose(xd.grad.cpu().numpy(), dx, rtol=1e-4, atol=1e-5)
    with pytest.raises(A.AgaError):
        A.layer_norm(torch.zeros(2, 768), w, b)  # CPU tensors are rejected: no fallback
    with pytest.raises(A.AgaError):
        A.layer_norm(torch.zeros(2, 100, device="cuda"), torch.ones(100, device="cuda"), torch.zeros(100, device="cuda"))


def _adapter_ref(x, w1, b1, w2, b2, gamma, beta):
    """whisper/whisper/model.py:193 + :234-236: LN(x + W2 gelu(W1 x + b1) + b2), eager PyTorch."""
    y = x + F.linear(F.gelu(F.linear(x, w1.to(x.dtype), b1.to(x.dtype))), w2.to(x.dtype), b2.to(x.dtype))
    return F.layer_norm(y.float(), (x.shape[-1],), gamma, beta, 1e-5).type(x.dtype)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("shape,D", [((4, 300), 768), ((2, 64), 1024), ((6, 1000), 768)])  # the last one runs the ring kernel
def test_adapter_layer_norm_matches_reference_expression(A, dtype, tol, shape, D):
    g = torch.Generator().manual_seed(D + shape[1])
    Bn = D // 4
    x = torch.randn(*shape, D, generator=g).to(dtype).cuda()
    params = [torch.randn(Bn, D, generator=g) / D ** 0.5, 0.02 * torch.randn(Bn, generator=g),
              torch.randn(D, Bn, generator=g) / Bn ** 0.5, 0.02 * torch.randn(D, generator=g),
              1.0 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)]
    dz = torch.randn(*shape, D, generator=g).to(dtype).cuda()
    outs = []
    for fn in (A.adapter_layer_norm, _adapter_ref):
        xs = x.clone().requires_grad_()
        ps = [p.cuda().requires_grad_() for p in params]
        z = fn(xs, *ps)
        z.backward(dz)
        outs.append((z, xs.grad, [p.grad for p in ps]))
    (z1, dx1, g1), (z2, dx2, g2) = outs
    assert z1.dtype == dtype
    torch.testing.assert_close(z1.float(), z2.float(), rtol=tol, atol=tol)
    torch.testing.assert_close(dx1.float(), dx2.float(), rtol=tol, atol=tol * float(dx2.float().abs().max()))
    for a, b in zip(g1, g2):
        assert a.dtype == torch.float32
        torch.testing.assert_close(a, b, rtol=tol, atol=tol * float(b.abs().max()))


@pytest.mark.parametrize("dtype,rows,cols", [(torch.float32, 1000, 96), (torch.bfloat16, 24000, 192), (torch.bfloat16, 333, 256),
                                             (torch.bfloat16, 7, 320), (torch.float32, 5, 192)])
def test_gelu_bwd_colsum_matches_aten(A, dtype, rows, cols):
    """dh = dg * gelu'(h) and its column sums (the Adapter's first-bias gradient, whisper/model.py:181-194) in one pass."""
    from aga_b200 import ops
    g = torch.Generator().manual_seed(rows + cols)
    h = (2.0 * torch.randn(rows, cols, generator=g)).to(dtype).cuda()
    dg = torch.randn(rows, cols, generator=g).to(dtype).cuda()
    dh, colsum = ops.gelu_bwd_colsum(dg, h)
    ref = torch.ops.aten.gelu_backward(dg, h)
    # fp32 oracle of the exact (erf) derivative
    x64 = h.double()
    exact = dg.double() * (0.5 * (1 + torch.erf(x64 / 2 ** 0.5)) + x64 * torch.exp(-0.5 * x64 * x64) / (2 * np.pi) ** 0.5)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    torch.testing.assert_close(dh.double(), exact, rtol=tol, atol=tol)
    torch.testing.assert_close(dh.float(), ref.float(), rtol=tol, atol=tol)
    torch.testing.assert_close(colsum, dh.float().sum(0), rtol=1e-4, atol=1e-3)


def test_conv_stem_gemm_matches_conv1d(A):
    """AudioEncoder.stem (two GEMMs on token-major activations) == gelu(conv2(gelu(conv1(x)))).permute(0,2,1)
    (whisper/model.py:277-279)."""
    from aga_b200 import whisper_model as W
    enc = W.AudioEncoder(80, 1500, 384, 6, 1).cuda()
    W.seeded_init_(enc, 1)
    for T in (3000, 261, 100):
        x = torch.randn(2, 80, T, generator=torch.Generator().manual_seed(T)).cuda()
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            ref = F.gelu(F.conv1d(F.gelu(F.conv1d(x, enc.conv1.weight, enc.conv1.bias, padding=1)), enc.conv2.weight,
                                  enc.conv2.bias, stride=2, padding=1)).permute(0, 2, 1)
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        got = enc.stem(x)
        assert got.shape == ref.shape
        torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            got16 = enc.stem(x)
        assert got16.dtype == torch.bfloat16
        torch.testing.assert_close(got16.float(), ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("rows,N,K,with_bias", [(24000, 768, 768, True), (1024, 768, 3072, True), (333, 1280, 1280, False),
                                                (7, 384, 1536, True)])
def test_linear_residual_matches_linear_plus_add(A, dtype, tol, rows, N, K, with_bias):
    """ops.linear_residual == residual + F.linear(x, w, b) (`x = x + self.attn(...)`, whisper/model.py:231-242), forward
    and the gradients w.r.t. x and the residual."""
    from aga_b200 import ops
    g = torch.Generator().manual_seed(rows + N)
    x = torch.randn(rows, K, generator=g).to(dtype).cuda().requires_grad_()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dtype).cuda()
    b = torch.randn(N, generator=g).to(dtype).cuda() if with_bias else None
    r = torch.randn(rows, N, generator=g).to(dtype).cuda().requires_grad_()
    do = torch.randn(rows, N, generator=g).to(dtype).cuda()
    out = ops.linear_residual(x, w, b, r)
    out.backward(do)
    x64, r64 = x.detach().double(), r.detach().double()
    ref = r64 + x64 @ w.double().t() + (b.double() if with_bias else 0.0)
    torch.testing.assert_close(out.double(), ref, rtol=tol, atol=tol * float(ref.abs().max()))
    dx_ref = do.double() @ w.double()
    torch.testing.assert_close(x.grad.double(), dx_ref, rtol=tol, atol=tol * float(dx_ref.abs().max()))
    assert torch.equal(r.grad, do)


@pytest.mark.parametrize("rows_shape", [(3, 50), (5, 1000)])  # 5000 rows: the cp.async ring kernel with the residual gradient
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_layer_norm_residual_backward_adds_residual_gradient(A, dtype, tol, rows_shape):
    """ops.layer_norm_residual: (LN(x), x) whose backward adds the residual path's gradient inside the LayerNorm-backward
    kernel == layer_norm(x) used next to x itself (`x = x + f(ln(x))`, whisper/model.py:231-242)."""
    from aga_b200 import ops
    g = torch.Generator().manual_seed(5)
    D = 768
    x0 = torch.randn(*rows_shape, D, generator=g).to(dtype).cuda()
    w = (1 + 0.1 * torch.randn(D, generator=g)).cuda().requires_grad_()
    b = (0.1 * torch.randn(D, generator=g)).cuda().requires_grad_()
    m = torch.randn(D, D, generator=g).to(dtype).cuda() / D ** 0.5
    do = torch.randn(*rows_shape, D, generator=g).to(dtype).cuda()
    res = {}
    for fused in (True, False):
        x = x0.clone().requires_grad_()
        if fused:
            y, xr = ops.layer_norm_residual(x, w, b)
        else:
            y, xr = A.layer_norm(x, w, b, 1e-5), x
        out = xr + torch.tanh(y @ m)
        gx, gw, gb = torch.autograd.grad(out, (x, w, b), do)
        res[fused] = (out.detach(), gx, gw, gb)
    assert torch.equal(res[True][0], res[False][0])
    for a, r in zip(res[True][1:], res[False][1:]):
        torch.testing.assert_close(a.float(), r.float(), rtol=tol, atol=tol * float(r.float().abs().max()))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("shape,D", [((4, 300), 768), ((6, 1000), 768), ((2, 64), 1280), ((5, 1000), 1280), ((16, 1500), 768)])  # last: the BASELINE size
def test_adapter_layer_norm_pair_equals_the_two_nodes(A, dtype, tol, shape, D):
    """ops.adapter_layer_norm_pair — the adapter post-LN and the frozen pre-LN that follows it, one forward kernel — gives
    the values and gradients of adapter_layer_norm followed by layer_norm_residual (z feeds both the next LN and the
    residual stream, whisper/model.py:234-246): outputs bit-identical, gradients equal up to the order of the atomics."""
    from aga_b200 import ops
    g = torch.Generator().manual_seed(D + shape[1])
    Bn = D // 4
    x = torch.randn(*shape, D, generator=g).to(dtype).cuda()
    params = [torch.randn(Bn, D, generator=g) / D ** 0.5, 0.02 * torch.randn(Bn, generator=g),
              torch.randn(D, Bn, generator=g) / Bn ** 0.5, 0.02 * torch.randn(D, generator=g),
              1.0 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)]
    g2 = (1.0 + 0.1 * torch.randn(D, generator=g)).cuda()
    b2 = (0.1 * torch.randn(D, generator=g)).cuda()
    m = torch.randn(D, D, generator=g).to(dtype).cuda() / D ** 0.5
    do = torch.randn(*shape, D, generator=g).to(dtype).cuda()
    res = {}
    for fused in (True, False):
        xs = x.clone().requires_grad_()
        ps = [p.cuda().requires_grad_() for p in params]
        if fused:
            z, y2 = ops.adapter_layer_norm_pair(xs, *ps, 1e-5, g2, b2, 1e-5)
        else:
            z = ops.adapter_layer_norm(xs, *ps, 1e-5)
            y2, z = ops.layer_norm_residual(z, g2, b2, 1e-5)
        out = z + torch.tanh(y2 @ m)  # z is used twice, as in `x = x + f(ln(x))`
        grads = torch.autograd.grad(out, [xs] + ps, do)
        res[fused] = (z.detach(), y2.detach(), grads)
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])
    for a, r in zip(res[True][2], res[False][2]):
        torch.testing.assert_close(a.float(), r.float(), rtol=tol, atol=tol * float(r.float().abs().max()))
