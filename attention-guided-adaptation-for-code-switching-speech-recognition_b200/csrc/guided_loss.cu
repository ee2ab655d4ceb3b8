// Attention-guided ("cs") loss, language pattern and head-vote kernels (tiny, latency-bound).
//
// Replaces, in espnet2/asr/espnet_model.py:
//   calculate_cs_loss            :463-530  -> guided_loss_head_kernel + guided_loss_finalize_kernel
//   create_attention_pattern     :236-275  -> attention_pattern_kernel (HF tokenizer loop -> LID table gather)
//   new_check_attention_language :285-310  -> head_vote_kernel (same fp32 summation order => same decisions)
#include "aga_common.cuh"

namespace aga {
namespace {

constexpr int kLossThreads = 128;

// One CTA per (head, utterance, layer): r[t] = sum_j (A~[t,j] - P~[t,j])^2 over the two exported columns,
// m = sum_t r / count_nonzero_t(r)  (0/0 -> NaN exactly like the reference), then the gradient.
__global__ void __launch_bounds__(kLossThreads)
guided_loss_head_kernel(const float* __restrict__ slab, int64_t sl, int64_t sb, int64_t sh, int64_t st,
                        const float* __restrict__ pattern, const float* __restrict__ head_mask,
                        int L, int B, int H, int T, int n_early, float* __restrict__ per_head,
                        float* __restrict__ d_slab) {
  const int h = blockIdx.x, b = blockIdx.y, l = blockIdx.z;
  const bool late = l >= n_early;
  const float* a = slab + l * sl + b * sb + h * sh;
  const float* p = pattern + int64_t(b) * T * 2;
  float sum = 0.0f;
  int cnt = 0;
  for (int t = threadIdx.x; t < T; t += kLossThreads) {
    float r = 0.0f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float av = a[t * st + j];
      const float pv = p[t * 2 + j];
      const bool pad = isinf(pv);
      const bool zero_a = (late && pad) || isinf(av);  // :496-497
      const float az = zero_a ? 0.0f : av;
      const float tg = (late && !pad) ? pv : 0.0f;     // early layers: target 0 in cols 1:3 (:479-481)
      const float d = az - tg;
      r += d * d;
    }
    sum += r;
    cnt += (r != 0.0f) ? 1 : 0;
  }
  __shared__ float s_sum[kLossThreads / 32];
  __shared__ int s_cnt[kLossThreads / 32];
  __shared__ float s_m;
  __shared__ int s_n;
  sum = warp_sum(sum);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) {
    s_sum[threadIdx.x >> 5] = sum;
    s_cnt[threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.0f;
    int n = 0;
    for (int w = 0; w < kLossThreads / 32; ++w) {
      tot += s_sum[w];
      n += s_cnt[w];
    }
    const float m = tot / float(n);  // :512 — count_nonzero denominator, NaN when n == 0
    s_m = m;
    s_n = n;
    per_head[(int64_t(l) * B + b) * H + h] = head_mask[l * H + h] * m;  // :527
  }
  if (d_slab == nullptr) return;
  __syncthreads();
  const float g = head_mask[l * H + h] * 2.0f / (float(B) * float(s_n));
  float* da = d_slab + l * sl + b * sb + h * sh;
  for (int t = threadIdx.x; t < T; t += kLossThreads) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float av = a[t * st + j];
      const float pv = p[t * 2 + j];
      const bool pad = isinf(pv);
      const bool zero_a = (late && pad) || isinf(av);
      const float tg = (late && !pad) ? pv : 0.0f;
      da[t * st + j] = zero_a ? 0.0f : g * (av - tg);
    }
  }
}

// loss = mean_b sum_{l,h} per_head[l,b,h]   (:529), fixed summation order.
__global__ void __launch_bounds__(256)
guided_loss_finalize_kernel(const float* __restrict__ per_head, int L, int B, int H, float* __restrict__ loss) {
  __shared__ float s[256];
  float acc = 0.0f;
  const int n = L * B * H;
  for (int i = threadIdx.x; i < n; i += 256) acc += per_head[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = s[0] / float(B);
}

// One warp per utterance.  Rows 0-4: fixed prompt rows; rows 5..first EOT: by LID class; after: +inf.
__global__ void __launch_bounds__(32)
attention_pattern_kernel(const int64_t* __restrict__ tokens, const uint8_t* __restrict__ lid, int vocab, int T,
                         float c, float* __restrict__ pattern) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int64_t* tk = tokens + int64_t(b) * T;
  float* out = pattern + int64_t(b) * T * 2;
  int first_eot = T;  // index of the first end-of-text at or after the prompt
  for (int base = 5; base < T && first_eot == T; base += 32) {
    const int t = base + lane;
    bool is_eot = false;
    if (t < T) {
      const int64_t id = tk[t];
      is_eot = (id >= 0 && id < vocab) && lid[id] == 3;
    }
    const unsigned m = __ballot_sync(0xffffffffu, is_eot);
    if (m) first_eot = base + __ffs(m) - 1;
  }
  for (int t = lane; t < T; t += 32) {
    float p0, p1;
    if (t < 5) {
      p0 = (t == 1) ? c : 0.0f;  // :261-265
      p1 = (t == 2) ? c : 0.0f;
    } else if (t > first_eot) {
      p0 = p1 = INFINITY;  // :268
    } else {
      const int64_t id = tk[t];
      const int cls = (id >= 0 && id < vocab) ? lid[id] : 0;
      if (cls >= 2) {  // space-only token or end-of-text: [c, c]  (:247-252)
        p0 = p1 = c;
      } else if (cls == 1) {  // English: column 2 (<|en|>)
        p0 = 0.0f;
        p1 = c;
      } else {  // Mandarin / other: column 1 (<|zh|>)
        p0 = c;
        p1 = 0.0f;
      }
    }
    out[t * 2] = p0;
    out[t * 2 + 1] = p1;
  }
}

// One CTA per (head, utterance, layer); thread j owns key column j and adds the rows in order, exactly the
// fp32 sequence of the reference's Python sum(); thread 0 then adds the column sums left to right.
__global__ void __launch_bounds__(256)
head_vote_kernel(const float* __restrict__ probs, int L, int B, int H, int T, uint8_t* __restrict__ decisions,
                 int32_t* __restrict__ counts) {
  extern __shared__ float s_col[];
  const int h = blockIdx.x, b = blockIdx.y, l = blockIdx.z;
  const float* a = probs + ((int64_t(l) * B + b) * H + h) * int64_t(T) * T;
  for (int j = threadIdx.x; j < T; j += blockDim.x) {
    float acc = 0.0f;
    for (int t = 0; t < T; ++t) acc = __fadd_rn(acc, a[int64_t(t) * T + j]);
    s_col[j] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s1 = 0.0f;
    if (T > 1) s1 = __fadd_rn(s1, s_col[1]);
    if (T > 2) s1 = __fadd_rn(s1, s_col[2]);
    float s2b = 0.0f;
    for (int j = 3; j < T; ++j) s2b = __fadd_rn(s2b, s_col[j]);
    const float s2 = __fadd_rn(s_col[0], s2b);
    const bool sel = s1 > s2;  // :299
    if (decisions) decisions[(int64_t(l) * B + b) * H + h] = sel ? 1 : 0;
    if (sel && counts) atomicAdd(counts + l * H + h, 1);
  }
}

}  // namespace
}  // namespace aga

using namespace aga;

extern "C" int aga_guided_loss_workspace_bytes(int L, int B, int H, size_t* bytes) {
  if (!bytes || L <= 0 || B <= 0 || H <= 0) return AGA_ERR_INVALID_ARGUMENT;
  *bytes = align_up(size_t(L) * B * H * sizeof(float), 256);
  return AGA_OK;
}

extern "C" int aga_guided_loss_fwd_bwd(const float* slab, int64_t stride_l, int64_t stride_b, int64_t stride_h,
                                       int64_t stride_t, const float* pattern, const float* head_mask, int L, int B,
                                       int H, int T, int n_early, float* loss, float* d_slab, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  size_t need = 0;
  int st = aga_guided_loss_workspace_bytes(L, B, H, &need);
  if (st != AGA_OK) return st;
  if (!slab || !pattern || !head_mask || !loss || !workspace || T <= 0 || n_early < 0) return AGA_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < need) return AGA_ERR_WORKSPACE_TOO_SMALL;
  if (B > 65535 || L > 65535) return AGA_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* per_head = static_cast<float*>(workspace);
  guided_loss_head_kernel<<<dim3(H, B, L), kLossThreads, 0, s>>>(slab, stride_l, stride_b, stride_h, stride_t, pattern,
                                                                  head_mask, L, B, H, T, n_early, per_head, d_slab);
  AGA_AFTER_LAUNCH();
  guided_loss_finalize_kernel<<<1, 256, 0, s>>>(per_head, L, B, H, loss);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

extern "C" int aga_attention_pattern(const int64_t* tokens, const uint8_t* lid_table, int vocab, int B, int T, float c,
                                     float* pattern, void* stream) {
  if (!tokens || !lid_table || !pattern || vocab <= 0 || B <= 0 || T <= 0) return AGA_ERR_INVALID_ARGUMENT;
  attention_pattern_kernel<<<B, 32, 0, static_cast<cudaStream_t>(stream)>>>(tokens, lid_table, vocab, T, c, pattern);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}

extern "C" int aga_head_vote(const float* probs, int L, int B, int H, int T, uint8_t* decisions, int32_t* counts,
                             void* stream) {
  if (!probs || L <= 0 || B <= 0 || H <= 0 || T <= 0 || (!decisions && !counts)) return AGA_ERR_INVALID_ARGUMENT;
  if (B > 65535 || L > 65535 || T > 8192) return AGA_ERR_UNSUPPORTED;
  head_vote_kernel<<<dim3(H, B, L), 256, size_t(T) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      probs, L, B, H, T, decisions, counts);
  AGA_AFTER_LAUNCH();
  return AGA_OK;
}
