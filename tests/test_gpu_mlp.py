"""GPU parity of the tcgen05 GEMM with the erf-GELU epilogue (csrc/gemm_gelu.cu) and of the fused MLP node against the
reference expression `x + Linear(GELU(Linear(x)))` (whisper/whisper/model.py:213,242) in bf16."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mk(M, K, N, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g).bfloat16().cuda()
    w1 = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().cuda()
    b1 = (0.5 * torch.randn(N, generator=g)).bfloat16().cuda()
    return x, w1, b1, g


@pytest.mark.parametrize("M,K,N", [(24000, 768, 3072), (1024, 768, 3072), (300, 1024, 4096), (129, 384, 1536), (7, 1280, 5120)])
def test_gemm_gelu_forward_and_backward_epilogues(M, K, N):
    from aga_b200 import ops
    x, w1, b1, g = _mk(M, K, N, M + N)
    h, act = ops.gemm_gelu_fwd(x, w1, b1)
    h_ref = F.linear(x, w1, b1)
    torch.testing.assert_close(h.float(), h_ref.float(), rtol=2e-2, atol=2e-2)
    # gelu of OUR h (the epilogue rounds h to bf16 first, as the reference's separate kernel sees it)
    torch.testing.assert_close(act.float(), F.gelu(h).float(), rtol=1e-2, atol=1e-3)
    # backward epilogue: dh = (dy @ w2) * gelu'(h), w2 (Kout, N); pass w2^T (N, Kout)
    Kout = K
    w2 = (torch.randn(Kout, N, generator=g) / N ** 0.5).bfloat16().cuda()
    dy = torch.randn(M, Kout, generator=g).bfloat16().cuda()
    dh = ops.gemm_gelu_bwd(dy, w2.t().contiguous(), h)
    dh_ref = torch.ops.aten.gelu_backward(dy @ w2, h)
    scale = float(dh_ref.float().abs().max())
    torch.testing.assert_close(dh.float(), dh_ref.float(), rtol=2e-2, atol=2e-2 * scale)


@pytest.mark.parametrize("M,K,N", [(24000, 768, 3072), (1100, 768, 3072 - 256), (300, 1024, 4096 + 8)])
def test_gemm_gelu_cluster_variants_agree_bit_for_bit(M, K, N):
    """The clustered variants of the kernel (multicast pair, two-SM MMA, multicast quad) differ only in how operand tiles
    reach shared memory: every accumulator sees the same products in the same order, so h, gelu(h) and dh are identical.
    Odd tile counts along M and N exercise the zero-filled / clipped mates of a cluster."""
    import ctypes
    from aga_b200 import ops, _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    x, w1, b1, g = _mk(M, K, N, M + N + 1)
    w2t = (torch.randn(N, K, generator=g) / N ** 0.5).bfloat16().cuda()
    dy = torch.randn(M, K, generator=g).bfloat16().cuda()
    prev = lib.aga_debug_set_gemm_variant(1)
    try:
        h1, a1 = ops.gemm_gelu_fwd(x, w1, b1)
        d1 = ops.gemm_gelu_bwd(dy, w2t, h1)
        for variant in (2, 3, 0):
            lib.aga_debug_set_gemm_variant(variant)
            h, act = ops.gemm_gelu_fwd(x, w1, b1)
            d = ops.gemm_gelu_bwd(dy, w2t, h1)
            assert torch.equal(h, h1) and torch.equal(act, a1) and torch.equal(d, d1), f"variant {variant}"
    finally:
        lib.aga_debug_set_gemm_variant(prev)


def test_mlp_residual_node_matches_reference_expression():
    from aga_b200 import ops
    M, D = 3000, 768
    x, w1, b1, g = _mk(M, D, 4 * D, 11)
    w2 = (torch.randn(D, 4 * D, generator=g) / (4 * D) ** 0.5).bfloat16().cuda()
    b2 = (0.5 * torch.randn(D, generator=g)).bfloat16().cuda()
    res = torch.randn(M, D, generator=g).bfloat16().cuda()
    do = torch.randn(M, D, generator=g).bfloat16().cuda()
    xa, ra = x.clone().requires_grad_(), res.clone().requires_grad_()
    out = ops.mlp_residual(xa, w1, b1, w2, w2.t().contiguous(), b2, ra)
    out.backward(do)
    xb, rb = x.clone().requires_grad_(), res.clone().requires_grad_()
    ref = rb + F.linear(F.gelu(F.linear(xb, w1, b1)), w2, b2)
    ref.backward(do)
    torch.testing.assert_close(out.float(), ref.float(), rtol=2e-2, atol=2e-2 * float(ref.detach().float().abs().max()))
    torch.testing.assert_close(xa.grad.float(), xb.grad.float(), rtol=2e-2, atol=2e-2 * float(xb.grad.float().abs().max()))
    assert torch.equal(ra.grad, rb.grad)


@pytest.mark.parametrize("rows,M,N", [(24000, 768, 192), (1024, 768, 192), (5000, 1024, 256), (3000, 1280, 320), (513, 128, 64)])
@pytest.mark.parametrize("transpose_out", [False, True])
def test_adapter_wgrad_kernel_vs_fp32_matmul(rows, M, N, transpose_out):
    """csrc/wgrad_tc.cu (a^T b over rows = batch x frames, MN-major tcgen05 operands, K splits combined with atomics)
    against the fp32 product of the same bf16 operands, in both output layouts; row counts that are not a multiple of
    the 64-row stage, N = 320 (two MMAs per step: 256 + 64 columns), and accumulation into a non-zero output."""
    from aga_b200 import ops, _lib
    g = torch.Generator().manual_seed(rows + M + N)
    a = torch.randn(rows, M, generator=g).bfloat16().cuda()
    b = torch.randn(rows, N, generator=g).bfloat16().cuda()
    ref = a.float().t() @ b.float()
    ref = ref.t().contiguous() if transpose_out else ref
    got = ops.adapter_wgrad(a, b, torch.float32, transpose_out=transpose_out)  # (rows < 8192: one cuBLAS GEMM instead)
    assert got.shape == ref.shape and got.dtype == torch.float32
    scale = float(ref.abs().max())
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4 * scale)
    got = torch.zeros_like(ref).contiguous()
    _lib.torch_ops().wgrad(a, b, got.view(-1), transpose_out)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4 * scale)
    # the C ABI adds into what is there
    out = torch.full_like(ref, 3.0).contiguous()
    _lib.torch_ops().wgrad(a, b, out.view(-1), transpose_out)
    torch.testing.assert_close(out, ref + 3.0, rtol=1e-4, atol=1e-4 * scale)
