// Thin PyTorch C++ extension over the C ABI of libaga_b200.so (include/aga_b200.h): TORCH_LIBRARY(aga, ...) operators
// that take at::Tensor arguments, allocate outputs / scratch with PyTorch's caching allocator, pass raw pointers, sizes
// and the CURRENT CUDA stream to the library and turn its status codes into exceptions (TORCH_CHECK).  No kernels here and
// no torch types below this file: the drop-in boundary stays the C ABI.  Replaces the ctypes marshalling
// of round 1 on every per-step call (ctypes remains for the host-pointer setup calls and the symbol tests).
//
// Every operator is CUDA-only: a CPU tensor is an error (there is no CPU implementation in this package).
#include <torch/library.h>
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <tuple>

#include "aga_b200.h"

namespace {

using at::Tensor;
using c10::optional;

void check(int status, const char* what) {
  TORCH_CHECK(status == AGA_OK, what, " failed: ", aga_status_str(status),
              status == AGA_ERR_CUDA ? " (cudaError " + std::to_string(aga_last_cuda_error()) + ")" : std::string());
}
void* stream_of(const Tensor& t) { return at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }
int dtype_of(const Tensor& t) {
  TORCH_CHECK(t.scalar_type() == at::kFloat || t.scalar_type() == at::kBFloat16, "aga: fp32 or bf16 tensors only, got ", t.scalar_type());
  return t.scalar_type() == at::kFloat ? AGA_F32 : AGA_BF16;
}
void need_cuda(const Tensor& t, const char* name) { TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor: aga_b200 has no CPU path"); }
const void* ptr(const optional<Tensor>& t) { return t.has_value() && t->defined() ? t->data_ptr() : nullptr; }
Tensor scratch(size_t bytes, const Tensor& like) {
  return at::empty({int64_t(bytes < 16 ? 16 : bytes)}, like.options().dtype(at::kByte));
}

// ------------------------------------------------------------------------------------------------ log-mel
Tensor logmel(const Tensor& audio, const Tensor& packed, int64_t n_mels, const optional<Tensor>& valid, bool tensor_core) {
  need_cuda(audio, "audio");
  c10::cuda::CUDAGuard guard(audio.device());
  TORCH_CHECK(audio.dim() == 2 && audio.scalar_type() == at::kFloat && audio.stride(1) == 1, "audio must be (B, N) fp32 rows");
  const int64_t B = audio.size(0), N = audio.size(1);
  Tensor out = at::empty({B, n_mels, N / 160}, audio.options());
  size_t nbytes = 0;
  check(aga_logmel_workspace_bytes(B, N, int(n_mels), &nbytes), "aga_logmel_workspace_bytes");
  Tensor ws = scratch(nbytes, audio);
  if (tensor_core) {
    check(aga_logmel_tc_fwd(audio.data_ptr<float>(), B, N, audio.stride(0), packed.data_ptr(), int(n_mels), out.data_ptr<float>(),
                            static_cast<const int32_t*>(ptr(valid)), ws.data_ptr(), nbytes, stream_of(audio)),
          "aga_logmel_tc_fwd");
  } else {
    TORCH_CHECK(!valid.has_value(), "valid_samples needs the tensor-core frontend");
    check(aga_logmel_fwd(audio.data_ptr<float>(), B, N, audio.stride(0), packed.data_ptr(), int(n_mels), out.data_ptr<float>(),
                         ws.data_ptr(), nbytes, stream_of(audio)),
          "aga_logmel_fwd");
  }
  return out;
}

// ------------------------------------------------------------------------------------------------ attention
void fill(aga_attn_params& p, const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& out, const Tensor& lse,
          int64_t n_head, bool causal, int64_t kind, int64_t lo, int64_t hi, const optional<Tensor>& head_sel,
          const void* export_buf, int64_t impl, const optional<Tensor>& kv_len, const optional<Tensor>& guided_pattern = c10::nullopt,
          void* guided_part = nullptr, bool guided_early = false) {
  p.dtype = dtype_of(q);
  p.impl = int(impl);
  p.B = int(q.size(0));
  p.H = int(n_head);
  p.Tq = int(q.size(1));
  p.Tk = int(k.size(1));
  p.causal = causal ? 1 : 0;
  p.export_kind = int(kind);
  p.export_lo = kind != AGA_EXPORT_NONE ? int(lo) : 0;
  p.export_hi = kind != AGA_EXPORT_NONE ? int(hi) : 0;
  p.q_stride_b = q.stride(0); p.q_stride_t = q.stride(1);
  p.k_stride_b = k.stride(0); p.k_stride_t = k.stride(1);
  p.v_stride_b = v.stride(0); p.v_stride_t = v.stride(1);
  p.o_stride_b = out.stride(0); p.o_stride_t = out.stride(1);
  p.q = q.data_ptr(); p.k = k.data_ptr(); p.v = v.data_ptr(); p.out = out.data_ptr();
  p.lse = lse.data_ptr<float>();
  p.head_sel = static_cast<const uint8_t*>(ptr(head_sel));
  p.export_buf = static_cast<float*>(const_cast<void*>(export_buf));
  p.kv_len = static_cast<const int32_t*>(ptr(kv_len));
  p.guided_pattern = static_cast<const float*>(ptr(guided_pattern));
  p.guided_part = static_cast<float*>(guided_part);
  p.guided_early = guided_early ? 1 : 0;
}

std::tuple<Tensor, Tensor, Tensor, Tensor> attn_fwd(const Tensor& q, const Tensor& k, const Tensor& v, int64_t n_head, bool causal,
                                                    int64_t kind, int64_t lo, int64_t hi, const optional<Tensor>& head_sel,
                                                    int64_t impl, const optional<Tensor>& kv_len,
                                                    const optional<Tensor>& guided_pattern, bool guided_early) {
  need_cuda(q, "q");
  c10::cuda::CUDAGuard guard(q.device());
  TORCH_CHECK(q.dim() == 3 && q.size(2) == n_head * 64, "head dim must be 64 (every Whisper size)");
  const int64_t B = q.size(0), Tq = q.size(1);
  Tensor out = at::empty({B, Tq, q.size(2)}, q.options());
  Tensor lse = at::empty({B, n_head, Tq}, q.options().dtype(at::kFloat));
  Tensor exp;
  if (kind != AGA_EXPORT_NONE) {
    // rows of unselected heads are never written by the kernel: define them as zero
    auto o = q.options().dtype(at::kFloat);
    exp = head_sel.has_value() ? at::zeros({B, n_head, Tq, hi - lo}, o) : at::empty({B, n_head, Tq, hi - lo}, o);
  }
  const bool guided = guided_pattern.has_value() && guided_pattern->defined();
  // per-(b, h, 32-row group) partial sums of the guided loss, written (not accumulated) by the attention epilogue
  Tensor part = guided ? at::zeros({B, n_head, 4, 2}, q.options().dtype(at::kFloat)) : at::empty({0}, q.options().dtype(at::kFloat));
  aga_attn_params p{};
  fill(p, q, k, v, out, lse, n_head, causal, kind, lo, hi, head_sel, exp.defined() ? exp.data_ptr() : nullptr, impl, kv_len,
       guided_pattern, guided ? part.data_ptr() : nullptr, guided_early);
  size_t nbytes = 0;
  check(aga_attn_fwd_workspace_bytes(&p, &nbytes), "aga_attn_fwd_workspace_bytes");
  Tensor ws = scratch(nbytes, q);
  check(aga_attn_fwd(&p, ws.data_ptr(), nbytes < 16 ? 16 : nbytes, stream_of(q)), "aga_attn_fwd");
  return {out, lse, exp.defined() ? exp : at::empty({0}, q.options().dtype(at::kFloat)), part};
}

void attn_bwd(const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& out, const Tensor& lse, const Tensor& dout,
              const optional<Tensor>& dexport, const optional<Tensor>& probs, Tensor dq, Tensor dk, Tensor dv, int64_t n_head,
              bool causal, int64_t kind, int64_t lo, int64_t hi, const optional<Tensor>& head_sel, int64_t impl,
              const optional<Tensor>& kv_len, const optional<Tensor>& guided_pattern, const optional<Tensor>& d_part,
              bool guided_early) {
  need_cuda(q, "q");
  c10::cuda::CUDAGuard guard(q.device());
  aga_attn_bwd_params bp{};
  const bool has_de = dexport.has_value() && dexport->defined();
  const void* ebuf = probs.has_value() && probs->defined() && probs->numel() ? probs->data_ptr() : nullptr;
  if (has_de && kind == AGA_EXPORT_LOGITS) ebuf = dexport->data_ptr();  // logits export: only the gradient is needed (non-null marker)
  const bool guided = guided_pattern.has_value() && guided_pattern->defined() && d_part.has_value() && d_part->defined();
  // (guided_part itself is not read by the backward; any non-null pointer satisfies the "both or none" check)
  fill(bp.fwd, q, k, v, out, lse, n_head, causal, has_de ? kind : int64_t(AGA_EXPORT_NONE), lo, hi, head_sel, ebuf, impl, kv_len,
       guided ? guided_pattern : c10::nullopt, guided ? d_part->data_ptr() : nullptr, guided_early);
  bp.d_guided_part = guided ? d_part->data_ptr<float>() : nullptr;
  bp.dout = dout.data_ptr();
  bp.d_export = has_de ? dexport->data_ptr<float>() : nullptr;
  bp.dq = dq.data_ptr(); bp.dk = dk.data_ptr(); bp.dv = dv.data_ptr();
  size_t nbytes = 0;
  check(aga_attn_bwd_workspace_bytes(&bp, &nbytes), "aga_attn_bwd_workspace_bytes");
  Tensor ws = scratch(nbytes, q);
  check(aga_attn_bwd(&bp, ws.data_ptr(), nbytes < 16 ? 16 : nbytes, stream_of(q)), "aga_attn_bwd");
}

// ------------------------------------------------------------------------------------------------ LayerNorm
std::tuple<Tensor, Tensor, Tensor, Tensor> layernorm_fwd(const Tensor& x2, const optional<Tensor>& residual2, const Tensor& gamma,
                                                         const Tensor& beta, double eps, bool want_sum) {
  need_cuda(x2, "x");
  c10::cuda::CUDAGuard guard(x2.device());
  const int64_t rows = x2.size(0), D = x2.size(1);
  Tensor y = at::empty_like(x2);
  const bool has_res = residual2.has_value() && residual2->defined();
  Tensor s = (has_res && want_sum) ? at::empty_like(x2) : Tensor();
  Tensor mean = at::empty({rows}, x2.options().dtype(at::kFloat)), rstd = at::empty({rows}, x2.options().dtype(at::kFloat));
  check(aga_layernorm_fwd(x2.data_ptr(), ptr(residual2), dtype_of(x2), rows, int(D), gamma.data_ptr<float>(), beta.data_ptr<float>(),
                          float(eps), y.data_ptr(), s.defined() ? s.data_ptr() : nullptr, mean.data_ptr<float>(),
                          rstd.data_ptr<float>(), stream_of(x2)),
        "aga_layernorm_fwd");
  return {y, s.defined() ? s : x2, mean, rstd};
}

std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> layernorm_pair_fwd(
    const Tensor& x2, const optional<Tensor>& residual2, const Tensor& gamma, const Tensor& beta, double eps, bool want_sum,
    const Tensor& gamma2, const Tensor& beta2, double eps2) {
  need_cuda(x2, "x");
  c10::cuda::CUDAGuard guard(x2.device());
  const int64_t rows = x2.size(0), D = x2.size(1);
  Tensor y = at::empty_like(x2), y2 = at::empty_like(x2);
  const bool has_res = residual2.has_value() && residual2->defined();
  Tensor s = (has_res && want_sum) ? at::empty_like(x2) : Tensor();
  auto fopt = x2.options().dtype(at::kFloat);
  Tensor mean = at::empty({rows}, fopt), rstd = at::empty({rows}, fopt), mean2 = at::empty({rows}, fopt), rstd2 = at::empty({rows}, fopt);
  check(aga_layernorm_pair_fwd(x2.data_ptr(), ptr(residual2), dtype_of(x2), rows, int(D), gamma.data_ptr<float>(),
                               beta.data_ptr<float>(), float(eps), y.data_ptr(), s.defined() ? s.data_ptr() : nullptr,
                               mean.data_ptr<float>(), rstd.data_ptr<float>(), gamma2.data_ptr<float>(), beta2.data_ptr<float>(),
                               float(eps2), y2.data_ptr(), mean2.data_ptr<float>(), rstd2.data_ptr<float>(), stream_of(x2)),
        "aga_layernorm_pair_fwd");
  return {y, s.defined() ? s : x2, mean, rstd, y2, mean2, rstd2};
}

std::tuple<Tensor, Tensor> layernorm_bwd(const Tensor& dy2, const Tensor& s2, const Tensor& gamma, const Tensor& mean,
                                         const Tensor& rstd, bool need_params, bool need_dxsum, const optional<Tensor>& dres2,
                                         const optional<Tensor>& pg_zeroed) {
  need_cuda(s2, "x");
  c10::cuda::CUDAGuard guard(s2.device());
  const int64_t rows = s2.size(0), D = s2.size(1);
  Tensor dx = at::empty_like(s2);
  Tensor pg;  // rows: dgamma, dbeta (, dxsum): one buffer, cleared by the library with a single memset
  float *dg = nullptr, *db = nullptr, *dxs = nullptr;
  // pg_zeroed: a (2 or 3, D) fp32 slice of the caller's once-per-step cleared arena -> the _acc entry point, no memset node
  const bool acc = need_params && pg_zeroed.has_value() && pg_zeroed->defined();
  if (need_params) {
    pg = acc ? *pg_zeroed : at::empty({need_dxsum ? 3 : 2, D}, s2.options().dtype(at::kFloat));
    TORCH_CHECK(pg.is_contiguous() && pg.scalar_type() == at::kFloat && pg.numel() == (need_dxsum ? 3 : 2) * D, "bad pg_zeroed");
    dg = pg.data_ptr<float>();
    db = dg + D;
    dxs = need_dxsum ? dg + 2 * D : nullptr;
  }
  check((acc ? aga_layernorm_bwd_acc : aga_layernorm_bwd)(dy2.data_ptr(), s2.data_ptr(), dtype_of(s2), rows, int(D), gamma.data_ptr<float>(), mean.data_ptr<float>(),
                          rstd.data_ptr<float>(), ptr(dres2), dx.data_ptr(), dg, db, dxs, stream_of(s2)),
        "aga_layernorm_bwd");
  return {dx, pg.defined() ? pg : at::empty({0}, s2.options().dtype(at::kFloat))};
}

std::tuple<Tensor, Tensor> gelu_bwd_colsum(const Tensor& dg, const Tensor& h, const optional<Tensor>& colsum_zeroed) {
  need_cuda(h, "h");
  c10::cuda::CUDAGuard guard(h.device());
  Tensor dh = at::empty_like(h);
  const bool acc = colsum_zeroed.has_value() && colsum_zeroed->defined();
  Tensor colsum = acc ? *colsum_zeroed : at::empty({h.size(1)}, h.options().dtype(at::kFloat));
  TORCH_CHECK(colsum.is_contiguous() && colsum.scalar_type() == at::kFloat && colsum.numel() == h.size(1), "bad colsum_zeroed");
  check((acc ? aga_gelu_bwd_colsum_acc : aga_gelu_bwd_colsum)(dg.data_ptr(), h.data_ptr(), dtype_of(h), h.size(0), int(h.size(1)), dh.data_ptr(),
                            colsum.data_ptr<float>(), stream_of(h)),
        "aga_gelu_bwd_colsum");
  return {dh, colsum};
}

// ------------------------------------------------------------------------------------------------ adapter weight gradients
void wgrad(const Tensor& a, const Tensor& b, Tensor out_zeroed, bool transpose_out) {
  need_cuda(a, "a");
  c10::cuda::CUDAGuard guard(a.device());
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && out_zeroed.scalar_type() == at::kFloat,
              "wgrad: bf16 operands, fp32 output");
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && a.is_contiguous() && b.is_contiguous() && out_zeroed.is_contiguous() &&
                  a.size(0) == b.size(0) && out_zeroed.numel() == a.size(1) * b.size(1), "wgrad: a (rows, M), b (rows, N), out M*N");
  check(aga_wgrad_bf16(a.data_ptr(), b.data_ptr(), a.size(0), int(a.size(1)), int(b.size(1)), out_zeroed.data_ptr<float>(),
                       transpose_out ? 1 : 0, stream_of(a)),
        "aga_wgrad_bf16");
}

// ------------------------------------------------------------------------------------------------ flat optimizer update
void flat_grad_norm(const Tensor& g, Tensor norm_out, const optional<Tensor>& step, const optional<Tensor>& skipped, Tensor ws) {
  need_cuda(g, "g");
  c10::cuda::CUDAGuard guard(g.device());
  TORCH_CHECK(g.scalar_type() == at::kFloat && g.is_contiguous() && norm_out.scalar_type() == at::kFloat, "flat_grad_norm: fp32");
  check(aga_flat_grad_norm(g.data_ptr<float>(), g.numel(), norm_out.data_ptr<float>(),
                           step.has_value() && step->defined() ? step->data_ptr<float>() : nullptr,
                           skipped.has_value() && skipped->defined() ? skipped->data_ptr<float>() : nullptr, ws.data_ptr(),
                           size_t(ws.numel()) * ws.element_size(), stream_of(g)),
        "aga_flat_grad_norm");
}

void flat_adamw(Tensor p, const Tensor& g, Tensor m, Tensor v, const Tensor& lr, double beta1, double beta2, double eps,
                double weight_decay, const Tensor& step, const optional<Tensor>& grad_norm, double max_norm,
                const optional<Tensor>& shadow) {
  need_cuda(p, "p");
  c10::cuda::CUDAGuard guard(p.device());
  TORCH_CHECK(p.scalar_type() == at::kFloat && g.scalar_type() == at::kFloat && m.scalar_type() == at::kFloat &&
                  v.scalar_type() == at::kFloat && lr.scalar_type() == at::kFloat && step.scalar_type() == at::kFloat,
              "flat_adamw: fp32 buffers and scalars");
  TORCH_CHECK(p.is_contiguous() && g.is_contiguous() && m.is_contiguous() && v.is_contiguous() && g.numel() == p.numel() &&
                  m.numel() == p.numel() && v.numel() == p.numel(), "flat_adamw: flat buffers of one length");
  const bool has_shadow = shadow.has_value() && shadow->defined();
  if (has_shadow) TORCH_CHECK(shadow->scalar_type() == at::kBFloat16 && shadow->numel() == p.numel() && shadow->is_contiguous(), "flat_adamw: bf16 shadow");
  check(aga_flat_adamw(p.data_ptr<float>(), g.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), p.numel(),
                       lr.data_ptr<float>(), beta1, beta2, eps, weight_decay, step.data_ptr<float>(),
                       grad_norm.has_value() && grad_norm->defined() ? grad_norm->data_ptr<float>() : nullptr, float(max_norm),
                       has_shadow ? shadow->data_ptr() : nullptr, stream_of(p)),
        "aga_flat_adamw");
}

// ------------------------------------------------------------------------------------------------ GEMMs with epilogues
Tensor linear_residual(const Tensor& x2, const Tensor& w, bool w_kn, const optional<Tensor>& bias, const Tensor& r2, Tensor ws) {
  need_cuda(x2, "x");
  c10::cuda::CUDAGuard guard(x2.device());
  const int64_t K = x2.size(1), N = w_kn ? w.size(1) : w.size(0);
  Tensor out = at::empty_like(r2);
  check(aga_linear_residual(x2.data_ptr(), w.data_ptr(), w_kn ? 1 : 0, ptr(bias), r2.data_ptr(), out.data_ptr(), dtype_of(x2),
                            x2.size(0), int(N), int(K), ws.data_ptr(), size_t(ws.numel()), stream_of(x2)),
        "aga_linear_residual");
  return out;
}

std::tuple<Tensor, Tensor> gemm_gelu_fwd(const Tensor& x2, const Tensor& w, const optional<Tensor>& bias) {
  need_cuda(x2, "x");
  c10::cuda::CUDAGuard guard(x2.device());
  const int64_t M = x2.size(0), K = x2.size(1), N = w.size(0);
  Tensor h = at::empty({M, N}, x2.options()), g = at::empty({M, N}, x2.options());
  check(aga_gemm_gelu(x2.data_ptr(), w.data_ptr(), ptr(bias), h.data_ptr(), g.data_ptr(), 0, M, int(N), int(K), stream_of(x2)),
        "aga_gemm_gelu");
  return {h, g};
}

Tensor gemm_gelu_bwd(const Tensor& dy2, const Tensor& w_t, const Tensor& h) {
  need_cuda(dy2, "dy");
  c10::cuda::CUDAGuard guard(dy2.device());
  const int64_t M = dy2.size(0), K = dy2.size(1), N = w_t.size(0);
  Tensor dh = at::empty({M, N}, dy2.options());
  check(aga_gemm_gelu(dy2.data_ptr(), w_t.data_ptr(), nullptr, const_cast<void*>(h.data_ptr()), dh.data_ptr(), 1, M, int(N), int(K),
                      stream_of(dy2)),
        "aga_gemm_gelu");
  return dh;
}

// ------------------------------------------------------------------------------------------------ vocabulary CE
std::tuple<Tensor, Tensor, Tensor> ls_ce_fwd(const Tensor& logits2, const Tensor& target, int64_t n_vocab, int64_t padding_idx,
                                             double smoothing) {
  need_cuda(logits2, "logits");
  c10::cuda::CUDAGuard guard(logits2.device());
  const int64_t rows = logits2.size(0), ld = logits2.size(1);
  auto f = logits2.options().dtype(at::kFloat);
  Tensor row_loss = at::empty({rows}, f), row_lse = at::empty({rows}, f), row_correct = at::empty({rows}, f.dtype(at::kInt));
  check(aga_ls_ce_fwd(logits2.data_ptr(), dtype_of(logits2), rows, int(n_vocab), ld, target.data_ptr<int64_t>(), padding_idx,
                      float(smoothing), row_loss.data_ptr<float>(), row_lse.data_ptr<float>(), row_correct.data_ptr<int32_t>(),
                      stream_of(logits2)),
        "aga_ls_ce_fwd");
  return {row_loss, row_lse, row_correct};
}

Tensor ls_ce_bwd(const Tensor& logits2, const Tensor& target, int64_t n_vocab, int64_t padding_idx, double smoothing,
                 const Tensor& row_lse, const Tensor& gscale) {
  need_cuda(logits2, "logits");
  c10::cuda::CUDAGuard guard(logits2.device());
  Tensor dlogits = at::empty_like(logits2);
  check(aga_ls_ce_bwd(logits2.data_ptr(), dtype_of(logits2), logits2.size(0), int(n_vocab), logits2.size(1),
                      target.data_ptr<int64_t>(), padding_idx, float(smoothing), row_lse.data_ptr<float>(),
                      gscale.data_ptr<float>(), 1.0f, dlogits.data_ptr(), stream_of(logits2)),
        "aga_ls_ce_bwd");
  return dlogits;
}

// ------------------------------------------------------------------------------------------------ guided loss / pattern / vote
Tensor attention_pattern(const Tensor& tokens, const Tensor& lid_table, double c) {
  need_cuda(tokens, "tokens");
  c10::cuda::CUDAGuard guard(tokens.device());
  const int64_t B = tokens.size(0), T = tokens.size(1);
  Tensor out = at::empty({B, T, 2}, tokens.options().dtype(at::kFloat));
  check(aga_attention_pattern(tokens.data_ptr<int64_t>(), lid_table.data_ptr<uint8_t>(), int(lid_table.numel()), int(B), int(T),
                              float(c), out.data_ptr<float>(), stream_of(tokens)),
        "aga_attention_pattern");
  return out;
}

std::tuple<Tensor, Tensor> guided_loss(const Tensor& slab, const Tensor& pattern, const Tensor& head_mask, int64_t n_early,
                                       bool need_grad) {
  need_cuda(slab, "slab");
  c10::cuda::CUDAGuard guard(slab.device());
  const int64_t L = slab.size(0), B = slab.size(1), H = slab.size(2), T = slab.size(3);
  Tensor loss = at::empty({}, slab.options());
  Tensor d_slab = need_grad ? at::empty_like(slab) : Tensor();
  size_t nbytes = 0;
  check(aga_guided_loss_workspace_bytes(int(L), int(B), int(H), &nbytes), "aga_guided_loss_workspace_bytes");
  Tensor ws = scratch(nbytes, slab);
  check(aga_guided_loss_fwd_bwd(slab.data_ptr<float>(), slab.stride(0), slab.stride(1), slab.stride(2), slab.stride(3),
                                pattern.data_ptr<float>(), head_mask.data_ptr<float>(), int(L), int(B), int(H), int(T), int(n_early),
                                loss.data_ptr<float>(), d_slab.defined() ? d_slab.data_ptr<float>() : nullptr, ws.data_ptr(),
                                nbytes, stream_of(slab)),
        "aga_guided_loss_fwd_bwd");
  return {loss, d_slab.defined() ? d_slab : at::empty({0}, slab.options())};
}

Tensor head_vote(const Tensor& probs, Tensor counts) {
  need_cuda(probs, "probs");
  c10::cuda::CUDAGuard guard(probs.device());
  const int64_t L = probs.size(0), B = probs.size(1), H = probs.size(2), T = probs.size(3);
  Tensor dec = at::empty({L, B, H}, probs.options().dtype(at::kByte));
  check(aga_head_vote(probs.data_ptr<float>(), int(L), int(B), int(H), int(T), dec.data_ptr<uint8_t>(), counts.data_ptr<int32_t>(),
                      stream_of(probs)),
        "aga_head_vote");
  return dec;
}

int64_t launch_count() { return int64_t(aga_launch_count()); }

}  // namespace

TORCH_LIBRARY(aga, m) {
  m.def("logmel(Tensor audio, Tensor packed, int n_mels, Tensor? valid, bool tensor_core) -> Tensor");
  m.def("attn_fwd(Tensor q, Tensor k, Tensor v, int n_head, bool causal, int kind, int lo, int hi, Tensor? head_sel, int impl, "
        "Tensor? kv_len, Tensor? guided_pattern, bool guided_early) -> (Tensor, Tensor, Tensor, Tensor)");
  m.def("attn_bwd(Tensor q, Tensor k, Tensor v, Tensor out, Tensor lse, Tensor dout, Tensor? dexport, Tensor? probs, Tensor(a!) dq, "
        "Tensor(b!) dk, Tensor(c!) dv, int n_head, bool causal, int kind, int lo, int hi, Tensor? head_sel, int impl, Tensor? kv_len, "
        "Tensor? guided_pattern, Tensor? d_part, bool guided_early) -> ()");
  m.def("layernorm_fwd(Tensor x, Tensor? residual, Tensor gamma, Tensor beta, float eps, bool want_sum) -> (Tensor, Tensor, Tensor, Tensor)");
  m.def("layernorm_pair_fwd(Tensor x, Tensor? residual, Tensor gamma, Tensor beta, float eps, bool want_sum, Tensor gamma2, "
        "Tensor beta2, float eps2) -> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)");
  m.def("layernorm_bwd(Tensor dy, Tensor x, Tensor gamma, Tensor mean, Tensor rstd, bool need_params, bool need_dxsum, Tensor? dres, "
        "Tensor? pg_zeroed) -> (Tensor, Tensor)");
  m.def("gelu_bwd_colsum(Tensor dg, Tensor h, Tensor? colsum_zeroed) -> (Tensor, Tensor)");
  m.def("wgrad(Tensor a, Tensor b, Tensor(a!) out_zeroed, bool transpose_out) -> ()");
  m.def("flat_grad_norm(Tensor g, Tensor(a!) norm_out, Tensor(b!)? step, Tensor(c!)? skipped, Tensor(d!) ws) -> ()");
  m.def("flat_adamw(Tensor(a!) p, Tensor g, Tensor(b!) m, Tensor(c!) v, Tensor lr, float beta1, float beta2, float eps, float weight_decay, "
        "Tensor step, Tensor? grad_norm, float max_norm, Tensor(d!)? shadow) -> ()");
  m.def("linear_residual(Tensor x, Tensor w, bool w_kn, Tensor? bias, Tensor residual, Tensor ws) -> Tensor");
  m.def("gemm_gelu_fwd(Tensor x, Tensor w, Tensor? bias) -> (Tensor, Tensor)");
  m.def("gemm_gelu_bwd(Tensor dy, Tensor w_t, Tensor h) -> Tensor");
  m.def("ls_ce_fwd(Tensor logits, Tensor target, int n_vocab, int padding_idx, float smoothing) -> (Tensor, Tensor, Tensor)");
  m.def("ls_ce_bwd(Tensor logits, Tensor target, int n_vocab, int padding_idx, float smoothing, Tensor row_lse, Tensor gscale) -> Tensor");
  m.def("attention_pattern(Tensor tokens, Tensor lid_table, float c) -> Tensor");
  m.def("guided_loss(Tensor slab, Tensor pattern, Tensor head_mask, int n_early, bool need_grad) -> (Tensor, Tensor)");
  m.def("head_vote(Tensor probs, Tensor(a!) counts) -> Tensor");
  m.def("launch_count() -> int");
}

TORCH_LIBRARY_IMPL(aga, CUDA, m) {
  m.impl("logmel", logmel);
  m.impl("attn_fwd", attn_fwd);
  m.impl("attn_bwd", attn_bwd);
  m.impl("layernorm_fwd", layernorm_fwd);
  m.impl("layernorm_pair_fwd", layernorm_pair_fwd);
  m.impl("layernorm_bwd", layernorm_bwd);
  m.impl("gelu_bwd_colsum", gelu_bwd_colsum);
  m.impl("wgrad", wgrad);
  m.impl("flat_grad_norm", flat_grad_norm);
  m.impl("flat_adamw", flat_adamw);
  m.impl("linear_residual", linear_residual);
  m.impl("gemm_gelu_fwd", gemm_gelu_fwd);
  m.impl("gemm_gelu_bwd", gemm_gelu_bwd);
  m.impl("ls_ce_fwd", ls_ce_fwd);
  m.impl("ls_ce_bwd", ls_ce_bwd);
  m.impl("attention_pattern", attention_pattern);
  m.impl("guided_loss", guided_loss);
  m.impl("head_vote", head_vote);
}

TORCH_LIBRARY_IMPL(aga, CompositeExplicitAutograd, m) { m.impl("launch_count", launch_count); }
