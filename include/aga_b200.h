/*
 * aga_b200.h — C ABI of the B200-native hot path for Attention-Guided Adaptation
 * (Whisper attention fwd/bwd with selected-head map export, guided "cs" loss,
 * Whisper log-mel frontend).
 *
 * Drop-in boundary.  Every entry point names the reference
 * interface it replaces; paths are relative to /root/reference/espnet, with
 *   W/  = whisper/whisper/   and   E2/ = espnet2/ .
 *
 * Contract shared by all functions
 *   - plain pointers and sizes only; all device buffers are owned by the caller
 *     (PyTorch's caching allocator in the reference integration);
 *   - the library never allocates or frees device memory, never synchronises
 *     the device and enqueues work only on the stream passed in
 *     (a cudaStream_t, passed as void*);
 *   - scratch space is caller-provided: query the size, pass the buffer;
 *   - return value: AGA_OK (0) or a negative aga_status; no C++ exceptions
 *     cross the ABI; aga_status_str() names the code;
 *   - re-entrant: no mutable global state beyond immutable per-device tables;
 *   - there is no CPU fallback and no other backend: the kernels are sm_100a.
 */
#ifndef AGA_B200_H_
#define AGA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGA_B200_VERSION 100

#if defined(__GNUC__)
#define AGA_API __attribute__((visibility("default")))
#else
#define AGA_API
#endif

typedef enum aga_status {
  AGA_OK = 0,
  AGA_ERR_INVALID_ARGUMENT = -1,
  AGA_ERR_UNSUPPORTED = -2,
  AGA_ERR_CUDA = -3,
  AGA_ERR_WORKSPACE_TOO_SMALL = -4,
  AGA_ERR_NO_SM100 = -5
} aga_status;

typedef enum aga_dtype { AGA_F32 = 0, AGA_BF16 = 1 } aga_dtype;

/* What the attention kernel writes to the side buffer (W/model.py:108-109):
 * LOGITS = the scaled, causally masked pre-softmax scores `qk` the reference
 * returns at HEAD; PROBS = softmax `w`, the "#modify here qk to w" variant that
 * head selection and plotting need (code_util/head_selection.md:5-9). */
typedef enum aga_export_kind { AGA_EXPORT_NONE = 0, AGA_EXPORT_LOGITS = 1, AGA_EXPORT_PROBS = 2 } aga_export_kind;

/* Which implementation ran (for tests / gpu_launches accounting). */
typedef enum aga_attn_impl { AGA_ATTN_AUTO = 0, AGA_ATTN_SIMT = 1, AGA_ATTN_TCGEN05 = 2 } aga_attn_impl;

AGA_API int aga_version(void);
AGA_API const char* aga_status_str(int status);
/* cudaError_t of the last failing CUDA call made by this library on the calling thread (0 if none). */
AGA_API int aga_last_cuda_error(void);
/* Number of kernels this library has launched in this process (all threads). */
AGA_API uint64_t aga_launch_count(void);
/* AGA_OK when device `dev` is compute capability 10.x (tcgen05/TMEM/TMA path usable). */
AGA_API int aga_device_is_sm100(int dev);

/* ------------------------------------------------------------------------------------------
 * Log-mel frontend.
 * Replaces OpenAIWhisperEncoder.log_mel_spectrogram (E2/asr/encoder/whisper_encoder.py:105-135),
 * its twin WhisperFrontend.log_mel_spectrogram (E2/asr/frontend/whisper.py:54-83) and
 * whisper.audio.log_mel_spectrogram (W/audio.py:110-157): reflect-pad 200, periodic Hann(400),
 * 400-point real DFT at hop 160, drop last frame, |.|^2, mel projection, log10(clamp 1e-10),
 * max(x, per-utterance max - 8), (x + 4) / 4.
 * ------------------------------------------------------------------------------------------ */

/* Bytes for the packed (banded) form of an (n_mels x 201) filterbank. */
AGA_API int aga_logmel_packed_filter_bytes(int n_mels, size_t* bytes);
/* Pack a dense device filterbank `melfb` (n_mels x 201 fp32, row-major — the tensor
 * whisper.audio.mel_filters returns, W/audio.py:92-107) into `packed`. Done once per device. */
AGA_API int aga_logmel_pack_filters(const float* melfb, int n_mels, void* packed, size_t packed_bytes, void* stream);
/* Scratch bytes for aga_logmel_fwd. */
AGA_API int aga_logmel_workspace_bytes(int64_t B, int64_t N, int n_mels, size_t* bytes);
/* audio: (B, N) fp32, row stride `ld` elements (16-byte aligned rows give 128-bit loads).
 * out:   (B, n_mels, N / 160) fp32 contiguous.  Requires N > 200 (torch.stft reflect pad). */
AGA_API int aga_logmel_fwd(const float* audio, int64_t B, int64_t N, int64_t ld, const void* packed_filters, int n_mels,
                   float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Tensor-core frontend (csrc/logmel_tc.cu): the same function for BANDED filterbanks — every bin feeds at most two
 * adjacent filters and the filter index does not decrease with the bin (the Slaney triangles of W/audio.py:92-107 and
 * their 128-bin sibling).  The 400-point DFT runs as split-precision fp16 GEMMs on tcgen05 (fp32-FFT accuracy).
 *   aga_logmel_filters_banded: 1 when the HOST filterbank (n_mels x 201, row-major) has that structure.
 *   aga_logmel_tc_pack: builds mel bands + DFT tables on the host and enqueues their upload into `packed`
 *     (aga_logmel_tc_packed_bytes bytes of device memory, 16-byte aligned); AGA_ERR_UNSUPPORTED if not banded.
 *   aga_logmel_tc_fwd: as aga_logmel_fwd.  `n_valid` (device int32 scalar, may be NULL): the batch's true common length
 *     when `audio` is zero-padded to N for a static launch shape — reflect padding happens at *n_valid, frames past
 *     *n_valid / 160 are written as zeros (what the reference's conv stem would pad with) and do not enter the maximum. */
AGA_API int aga_logmel_filters_banded(const float* melfb_host, int n_mels);
AGA_API int aga_logmel_tc_packed_bytes(int n_mels, size_t* bytes);
/* the packed image in HOST memory (what aga_logmel_tc_pack uploads): header | float4 bin[201] = {w(L), w(L+1), L, 0} | DFT tables */
AGA_API int aga_logmel_tc_build_host(const float* melfb_host, int n_mels, void* packed_host, size_t packed_bytes);
AGA_API int aga_logmel_tc_pack(const float* melfb_host, int n_mels, void* packed, size_t packed_bytes, void* stream);
AGA_API int aga_logmel_tc_fwd(const float* audio, int64_t B, int64_t N, int64_t ld, const void* packed_tc, int n_mels,
                      float* out, const int32_t* n_valid, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-head attention core, head dim 64 (every Whisper size).
 * Replaces MultiHeadAttention.qkv_attention (W/model.py:93-109) and what autograd replays for it.
 *   S = (q d^-1/4)(k d^-1/4)^T (+ causal -inf mask, W/model.py:322), P = softmax_fp32(S), out = P v.
 * Tensors are addressed as x[b, t, h*64 + c] = base + b*stride_b + t*stride_t + h*64 + c (elements),
 * i.e. the (B, T, D) activations of the q/k/v Linear layers with heads interleaved along D.
 * ------------------------------------------------------------------------------------------ */
typedef struct aga_attn_params {
  int32_t dtype;  /* aga_dtype of q, k, v, out, dout, dq, dk, dv */
  int32_t impl;   /* aga_attn_impl; AUTO = tcgen05 for bf16, SIMT for fp32 */
  int32_t B, H, Tq, Tk;
  int32_t causal; /* 1: key j visible to query i iff j <= i (requires Tq == Tk) */
  int32_t export_kind;           /* aga_export_kind */
  int32_t export_lo, export_hi;  /* exported key columns [lo, hi) */
  int64_t q_stride_b, q_stride_t;
  int64_t k_stride_b, k_stride_t;
  int64_t v_stride_b, v_stride_t;
  int64_t o_stride_b, o_stride_t;  /* out and dout */
  const void* q;
  const void* k;
  const void* v;
  void* out;                /* (B, Tq, H*64) */
  float* lse;               /* (B, H, Tq) fp32 natural-log row log-sum-exp of S (needed by backward) */
  const uint8_t* head_sel;  /* (H) 0/1: heads whose columns are exported; NULL = all heads */
  float* export_buf;        /* (B, H, Tq, hi-lo) fp32; rows of unselected heads are left untouched */
  const int32_t* kv_len;    /* NULL, or a DEVICE scalar: only keys [0, min(*kv_len, Tk)) exist (non-causal attention on a batch
                             * zero-padded to a static Tk for CUDA-graph replay; the reference never computes the padded keys).
                             * Backward (tcgen05 path only): dk / dv rows of the partial key tile at or past *kv_len are written
                             * as zeros, whole key tiles past it are not touched (zero-initialise dk / dv). */
  /* Guided-loss reduction fused into the attention epilogue (decoder self attention: causal, Tq <= 128, tcgen05 path).
   * Replaces the per-(utterance, layer, head) part of ESPnetASRModel.calculate_cs_loss (E2/asr/espnet_model.py:496-512):
   * r_t = sum_{j in {1,2}} (S~[t,j] - c[t,j])^2 with the reference's zeroing rules (-inf -> 0; pad rows -> 0 unless
   * guided_early), then guided_part[b,h,g,0] = sum_t r_t and guided_part[b,h,g,1] = #{t : r_t != 0} over the 32 query
   * rows t of row group g = t / 32 (no atomics: deterministic).  The (L,B,H,T,2) slab need not be exported at all. */
  const float* guided_pattern; /* NULL, or (B, Tq, 2) fp32 from aga_attention_pattern (+inf rows = padding) */
  float* guided_part;          /* (B, H, 4, 2) fp32, zero-initialised by the caller */
  int32_t guided_early;        /* 1: layer uses the all-zero "early" target and keeps pad rows (espnet_model.py:479-481) */
} aga_attn_params;

AGA_API int aga_attn_fwd_workspace_bytes(const aga_attn_params* p, size_t* bytes);
AGA_API int aga_attn_fwd(const aga_attn_params* p, void* workspace, size_t workspace_bytes, void* stream);

typedef struct aga_attn_bwd_params {
  aga_attn_params fwd;      /* same problem as forward; fwd.out / fwd.lse / fwd.export_buf are inputs here */
  const void* dout;         /* (B, Tq, H*64), strides o_stride_* */
  const float* d_export;    /* gradient w.r.t. export_buf, same layout, or NULL */
  void* dq;                 /* strides as q */
  void* dk;                 /* strides as k */
  void* dv;                 /* strides as v */
  const float* d_guided_part; /* NULL, or (B, H, 4, 2): gradient w.r.t. fwd.guided_part ([...,0] is used); the kernel adds
                               * 2 (S~ - c) d_guided_part[b,h,t/32,0] to dS on the visible, non-zeroed entries of columns 1, 2 */
} aga_attn_bwd_params;

AGA_API int aga_attn_bwd_workspace_bytes(const aga_attn_bwd_params* p, size_t* bytes);
AGA_API int aga_attn_bwd(const aga_attn_bwd_params* p, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Attention-guided ("cs") loss on the exported columns.
 * Replaces ESPnetASRModel.calculate_cs_loss (E2/asr/espnet_model.py:463-530): MSE between the
 * decoder self-attention columns 1:3 (<|zh|>, <|en|> prompt tokens) and the per-token language
 * pattern, mean over non-zero rows, masked by the selected-head matrix, mean over the batch.
 *   slab element (l,b,h,t,j), j in {0,1}  = slab[l*stride_l + b*stride_b + h*stride_h + t*stride_t + j]
 *   (compact (L,B,H,T,2) export, or a view of columns 1:3 of full (L,B,H,T,T) maps)
 *   pattern (B,T,2) fp32 from aga_attention_pattern; +inf marks pad rows
 *   head_mask (L,H) fp32; layers [0,n_early) use the all-zero "early" target and keep pad rows
 * loss: 1 fp32 on device.  d_slab (may be NULL): d loss / d slab, same strides as slab.
 * ------------------------------------------------------------------------------------------ */
AGA_API int aga_guided_loss_workspace_bytes(int L, int B, int H, size_t* bytes);
AGA_API int aga_guided_loss_fwd_bwd(const float* slab, int64_t stride_l, int64_t stride_b, int64_t stride_h, int64_t stride_t,
                            const float* pattern, const float* head_mask, int L, int B, int H, int T, int n_early,
                            float* loss, float* d_slab, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces ESPnetASRModel.create_attention_pattern (E2/asr/espnet_model.py:236-275), with the
 * per-step HF-tokenizer string loop turned into a (vocab) uint8 lookup:
 * lid_table[id]: 0 other/Mandarin, 1 ASCII-letters-only (English), 2 space-only, 3 end-of-text.
 * tokens (B,T) int64 = ys_in_pad; pattern (B,T,2) fp32. */
AGA_API int aga_attention_pattern(const int64_t* tokens, const uint8_t* lid_table, int vocab, int B, int T, float c,
                          float* pattern, void* stream);

/* Replaces ESPnetASRModel.new_check_attention_language (E2/asr/espnet_model.py:285-310).
 * probs: full maps (L,B,H,T,T) fp32 contiguous.  decisions (L,B,H) uint8 (sum_1 > sum_2);
 * counts (L,H) int32, incremented per selected (utterance, layer, head). Sums are accumulated in
 * fp32 in the reference's order so near-ties decide identically. */
AGA_API int aga_head_vote(const float* probs, int L, int B, int H, int T, uint8_t* decisions, int32_t* counts, void* stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm with fp32 statistics on bf16 / fp32 rows (the first "next" row).
 * Replaces whisper.model.LayerNorm.forward (W/model.py:30-32: `super().forward(x.float()).type(x.dtype)`),
 * called four times per ResidualAttentionBlock (attn_ln, adapter_attn_ln, mlp_ln, adapter_mlp_ln; the two
 * adapter LNs are trainable) — one kernel instead of up-cast + normalise + down-cast.
 *   x, y, dy, dx : (rows, D) contiguous, dtype AGA_F32 or AGA_BF16, 16-byte aligned; D in {384,512,768,1024,1280}
 *   residual     : NULL, or (rows, D): the kernel normalises s = dtype(x + residual) — the Adapter's `x + self.model(x)`
 *                  (W/model.py:193) feeding its post-LayerNorm (W/model.py:234-236) — and writes s to sum_out if non-NULL
 *   gamma, beta  : (D) fp32;  mean, rstd : (rows) fp32, written by fwd and read by bwd (x of bwd = the normalised tensor)
 *   dgamma, dbeta: (D) fp32, OVERWRITTEN with this call's parameter gradients, or both NULL (frozen LN)
 *   dxsum        : NULL, or (D) fp32 OVERWRITTEN with sum_rows dx — the bias gradient of the Linear that produced `residual`
 *   dres         : NULL, or (rows, D): the gradient that reached the normalised tensor through the block's residual
 *                  connection (`x = x + f(ln(x))`, W/model.py:231-242); dx = dtype(dx_ln) + dres in the same pass
 * ------------------------------------------------------------------------------------------ */
AGA_API int aga_layernorm_fwd(const void* x, const void* residual, int dtype, int64_t rows, int D, const float* gamma,
                      const float* beta, float eps, void* y, void* sum_out, float* mean, float* rstd, void* stream);
/* aga_layernorm_fwd followed, on the row just produced (as stored: rounded to dtype), by a SECOND LayerNorm (gamma2, beta2,
 * eps2 -> y2, mean2, rstd2): ResidualAttentionBlock.forward applies the adapter's post-LN and then the next branch's pre-LN to
 * the same tensor (W/model.py:231-246); one kernel, one read of the row.  Same results as two aga_layernorm_fwd calls. */
AGA_API int aga_layernorm_pair_fwd(const void* x, const void* residual, int dtype, int64_t rows, int D, const float* gamma,
                      const float* beta, float eps, void* y, void* sum_out, float* mean, float* rstd, const float* gamma2,
                      const float* beta2, float eps2, void* y2, float* mean2, float* rstd2, void* stream);
AGA_API int aga_layernorm_bwd(const void* dy, const void* x, int dtype, int64_t rows, int D, const float* gamma,
                      const float* mean, const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                      float* dxsum, void* stream);
/* The same, but dgamma / dbeta / dxsum are ADDED TO: the caller zero-initialises them.  A training step has ~100 of these
 * small outputs; carving them out of one buffer that is cleared once per step removes one memset node (and the
 * dependency bubble behind it) per call from the captured step. */
AGA_API int aga_layernorm_bwd_acc(const void* dy, const void* x, int dtype, int64_t rows, int D, const float* gamma,
                      const float* mean, const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                      float* dxsum, void* stream);

/* Backward of the Adapter's GELU fused with the bias gradient of its first Linear (W/model.py:181-194,
 * Adapter.model = Linear -> GELU -> Linear).  dh = dg * gelu'(h) (exact erf form, fp32 math,
 * rounded to dtype) and colsum[c] = sum_rows dh[r, c] (of the rounded values), in one pass.
 *   dg, h, dh : (rows, cols) contiguous, dtype AGA_F32 or AGA_BF16, 16-byte aligned; cols a multiple of 16 B / sizeof(dtype)
 *   colsum    : (cols) fp32, OVERWRITTEN */
AGA_API int aga_gelu_bwd_colsum(const void* dg, const void* h, int dtype, int64_t rows, int cols, void* dh, float* colsum,
                        void* stream);
/* The same with colsum ADDED TO (zero-initialised by the caller; see aga_layernorm_bwd_acc). */
AGA_API int aga_gelu_bwd_colsum_acc(const void* dg, const void* h, int dtype, int64_t rows, int cols, void* dh, float* colsum,
                        void* stream);

/* Weight gradient of an adapter Linear (W/model.py:181-194; autograd of `Linear -> GELU -> Linear` with rows = batch x
 * frames): out += a^T b on tcgen05, a (rows, M) and b (rows, N) bf16 row-major (16-byte aligned), out fp32 — (M, N) row-major,
 * or (N, M) when transpose_out — ADDED TO with atomics: the caller zero-initialises it.  M % 64 == 0, N % 64 == 0;
 * anything else returns AGA_ERR_UNSUPPORTED.  Replaces cuBLAS's split-K kernel + reduce pass. */
AGA_API int aga_wgrad_bf16(const void* a, const void* b, int64_t rows, int M, int N, float* out, int transpose_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer update of the trainable (adapter) parameters on FLAT buffers.
 * Replaces the tail of Trainer.train_one_epoch's inner step (espnet2/train/trainer.py:649-716): clip_grad_norm_ over all
 * trainable parameters, "skip the update when the norm is not finite", torch.optim.AdamW.step() (recipe: optim adamw),
 * plus the fp32 -> bf16 re-cast of the updated parameters that autocast performs on their next use.
 *   aga_flat_grad_norm: norm_out[0] = ||g||_2 (fp32, deterministic); then, by the last block, step[0] += 1 when the norm is
 *     finite, skipped[0] += 1 otherwise (either may be NULL).  workspace: zero-initialised ONCE by the caller, 8-byte aligned,
 *     aga_flat_grad_norm_workspace_bytes() bytes; left ready for the next call.
 *   aga_flat_adamw: when grad_norm[0] is finite (or grad_norm is NULL): g' = g * min(1, max_norm / (grad_norm + 1e-6))
 *     (max_norm <= 0 or grad_norm NULL: no clipping), then torch.optim.AdamW's update with bias corrections for step[0]
 *     (already advanced), statement for statement the arithmetic of torch's fused kernel; shadow_bf16 (may be NULL) receives
 *     bf16(p).  A non-finite norm leaves p, m, v, shadow untouched.  p, g, m, v: n fp32, 16-byte aligned; lr, step,
 *     grad_norm: DEVICE scalars (a captured CUDA graph keeps following an LR schedule that writes lr in place).
 * ------------------------------------------------------------------------------------------ */
AGA_API int aga_flat_grad_norm_workspace_bytes(size_t* bytes);
AGA_API int aga_flat_grad_norm(const float* g, int64_t n, float* norm_out, float* step, float* skipped, void* workspace,
                       size_t workspace_bytes, void* stream);
AGA_API int aga_flat_adamw(float* p, const float* g, float* m, float* v, int64_t n, const float* lr, double beta1, double beta2,
                   double eps, double weight_decay, const float* step, const float* grad_norm, float max_norm,
                   void* shadow_bf16, void* stream);

/* Label-smoothing cross entropy + accuracy on the vocabulary logits, kept in the GEMM's dtype.
 * Replaces `.float()` of the logits (E2/asr/decoder/whisper_decoder.py:164-166), LabelSmoothingLoss.forward
 * (espnet/nets/pytorch_backend/transformer/label_smoothing_loss.py:41-63) and th_accuracy
 * (espnet/nets/pytorch_backend/nets_utils.py:304-324).
 *   logits  : (rows, ld) AGA_F32 / AGA_BF16, 16-byte aligned rows (ld a multiple of 16 B / sizeof(dtype)); columns [V, ld)
 *             are padding of the aligned GEMM and ignored
 *   target  : (rows) int64; rows with target == padding_idx contribute nothing
 *   fwd     : row_loss[r] = KL(smoothed one-hot || softmax) (0 for ignored rows), row_lse[r] = log-sum-exp,
 *             row_correct[r] = 1 iff arg-max (lowest index) == target (0 for ignored rows)
 *   bwd     : dlogits = (*gscale) * inv_denom * (softmax - smoothed one-hot), zero in padding columns / ignored rows;
 *             gscale = device scalar with the upstream gradient (NULL = 1) */
AGA_API int aga_ls_ce_fwd(const void* logits, int dtype, int64_t rows, int V, int64_t ld, const int64_t* target,
                  int64_t padding_idx, float smoothing, float* row_loss, float* row_lse, int32_t* row_correct, void* stream);
AGA_API int aga_ls_ce_bwd(const void* logits, int dtype, int64_t rows, int V, int64_t ld, const int64_t* target,
                  int64_t padding_idx, float smoothing, const float* row_lse, const float* gscale, float inv_denom,
                  void* dlogits, void* stream);

/* out (rows, N) = x (rows, K) @ W + bias (N) + residual (rows, N), W = w^T for w_layout 0 (w is an (N, K) nn.Linear weight)
 * or W = w for w_layout 1 (w is (K, N): the dgrad form dY @ weight): a Linear with the residual add folded into the
 * GEMM (cuBLASLt: C = residual, beta = 1, bias epilogue) — `x = x + self.attn(...)`, `x = x + self.mlp(...)` of
 * ResidualAttentionBlock.forward (W/model.py:231-242).  All row-major contiguous, one dtype, 16-byte
 * aligned; bias may be NULL; workspace of aga_linear_residual_workspace_bytes() bytes.  AGA_ERR_UNSUPPORTED when cuBLASLt
 * is not available in the process. */
AGA_API int aga_linear_residual_workspace_bytes(size_t* bytes);
AGA_API int aga_linear_residual(const void* x, const void* w, int w_layout, const void* bias, const void* residual, void* out,
                        int dtype, int64_t rows, int N, int K, void* workspace, size_t workspace_bytes, void* stream);

/* bf16 GEMM with the exact (erf) GELU in its epilogue (tcgen05): the MLP of ResidualAttentionBlock, W/model.py:213,242
 * (Linear -> GELU -> Linear).  a (M, K), w (N, K) row-major bf16, K and N multiples of 8, 16-byte aligned.
 *   mode 0: h = bf16(a w^T + bias), out = bf16(gelu(h)) — h (M, N) is written (kept for the backward pass), bias (N) bf16 or NULL
 *   mode 1: out = bf16(bf16(a w^T) * gelu'(h))          — h (M, N) is read; pass the transposed second-Linear weight as w */
AGA_API int aga_gemm_gelu(const void* a, const void* w, const void* bias, void* h, void* out, int mode, int64_t M, int N, int K,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AGA_B200_H_ */
