// Loss + optimiser kernels of the training step (see include/dfol_b200.h).
#include "dfol_common.cuh"

namespace dfol {

// VQATrainer._compute_loss (trainer.py:181-262) and d loss / d lp, scaled by 1/total questions (:434-435).
// BINARY follows torch.nn.functional.binary_cross_entropy on p = exp(lp): log terms clamped at -100,
// derivative (p - y) / max(p (1 - p), 1e-12).
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ lp, const float* __restrict__ target,
                                                   const int32_t* __restrict__ seg, int n_seg, int n_lp, int kind,
                                                   float scale, float* __restrict__ loss_out,
                                                   float* __restrict__ d_lp) {
  __shared__ float red[8];
  float local = 0.f;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (kind == 0) {
    if (i < n_lp) {
      const float x = lp[i], y = target[i];
      const float p = expf(x);
      const float l1 = fmaxf(logf(p), -100.0f), l0 = fmaxf(log1pf(-p), -100.0f);
      local = (y - 1.0f) * l0 - y * l1;
      if (d_lp) d_lp[i] = scale * ((p - y) / fmaxf((1.0f - p) * p, 1e-12f)) * p;
    }
  } else if (kind == 1) {
    // one warp per option list (query-type questions have ~100 options): coalesced reads, warp sums
    const int w = i >> 5, lane = threadIdx.x & 31;
    if (w < n_seg) {
      const int a = seg[w], b = seg[w + 1];
      float s = 0.f, dot = 0.f;
      for (int k = a + lane; k < b; k += 32) { s += expf(lp[k]); dot += target[k] * lp[k]; }
      s = warp_sum(s);
      dot = warp_sum(dot);
      if (lane == 0) local = slog(s) - dot;
      if (d_lp) {
        const float inv = (s >= kLogEps) ? 1.0f / s : 0.0f;
        for (int k = a + lane; k < b; k += 32) d_lp[k] = scale * (expf(lp[k]) * inv - target[k]);
      }
    }
  } else {
    if (i < n_lp) {
      local = -lp[i];
      if (d_lp) d_lp[i] = -scale;
    }
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(loss_out, scale * t);
  }
}

// Answers on the device (eval / predict path; reference: the per-op .cpu().numpy().tolist() of batch_gqa_ops.py:222-225,
// :404-407, :744-748 and util.find_max_ind util.py:64-66).  One warp per question:
//   mode 0 (binary)  : best = exp(lp) > 0.5 (1 = yes), count = 1
//   mode 1 (options) : p = exp(lp) over the question's option segment; the answer set is {k : p_k == max p and
//                      p_k > threshold} (exact ties, as find_max_ind); first = first member, count = size of the set,
//                      sel[k] = membership
//   mode 2 (compare) : argmax of the two entries (first wins a tie, as numpy.argmax)
// so the host reads 12 bytes per question instead of the whole log-probability vector (5 KB for a 1356-noun query) and
// only touches the membership bytes of questions whose answer is not a single option.
__global__ void __launch_bounds__(256) answers_kernel(const float* __restrict__ lp, const int32_t* __restrict__ seg,
                                                      int n_q, int mode, float threshold, int32_t* __restrict__ first,
                                                      int32_t* __restrict__ count, float* __restrict__ best_lp,
                                                      uint8_t* __restrict__ sel) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (q >= n_q) return;
  if (mode == 0) {
    if (lane == 0) {
      const float x = lp[q];
      first[q] = expf(x) > 0.5f ? 1 : 0;
      count[q] = 1;
      best_lp[q] = x;
    }
    return;
  }
  const int a = seg[q], b = seg[q + 1];
  if (mode == 2) {
    if (lane == 0) {
      const int k = lp[a + 1] > lp[a] ? 1 : 0;
      first[q] = k;
      count[q] = 1;
      best_lp[q] = lp[a + k];
    }
    return;
  }
  float mx = -1.0f;
  for (int k = a + lane; k < b; k += 32) mx = fmaxf(mx, expf(lp[k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  int cnt = 0, fst = 0x7fffffff;
  for (int k = a + lane; k < b; k += 32) {
    const float pk = expf(lp[k]);
    const bool in = (pk == mx) && (pk > threshold);
    if (sel != nullptr) sel[k] = in ? 1 : 0;
    if (in) { ++cnt; fst = min(fst, k - a); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    fst = min(fst, __shfl_xor_sync(0xffffffffu, fst, o));
  }
  if (lane == 0) {
    first[q] = cnt > 0 ? fst : -1;
    count[q] = cnt;
    best_lp[q] = cnt > 0 ? lp[a + fst] : 0.0f;
  }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* out) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i];
    acc += v * v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(out, t);
  }
}

// L1 regularisation of VQATrainer._compute_loss (trainer.py:257-259): loss += lambda * ||theta||_1 / numel.
// g += coef * sign(p) (torch's subgradient: sign(0) = 0); loss_out[0] += loss_coef * sum |p|.
__global__ void __launch_bounds__(256) l1_kernel(const float* __restrict__ p, float* __restrict__ g, long long n,
                                                 float coef, float loss_coef, float* loss_out) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = p[i];
    acc += fabsf(v);
    g[i] += (v > 0.0f) ? coef : (v < 0.0f ? -coef : 0.0f);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && loss_out != nullptr) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(loss_out, loss_coef * t);
  }
}

// torch.nn.utils.clip_grad_norm_ (coef = clip / (norm + 1e-6), clamped to 1) followed by torch.optim.Adam
// (L2 weight decay folded into the gradient, bias-corrected moments, eps added after the sqrt).
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n,
                                                   const float* __restrict__ sumsq, float clip_norm, float lr,
                                                   float beta1, float beta2, float eps, float wd, float bc1,
                                                   float bc2_sqrt) {
  float coef = 1.0f;
  if (clip_norm > 0.0f) coef = fminf(1.0f, clip_norm / (sqrtf(sumsq[0]) + 1e-6f));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pi = p[i];
    float gi = g[i] * coef + wd * pi;
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

}  // namespace dfol

using namespace dfol;

extern "C" int dfol_loss_fwd_bwd(const float* lp, const float* target, const int32_t* seg, int n_seg, int n_lp,
                                 int kind, float scale, float* loss_out, float* d_lp, void* stream) {
  DFOL_REQUIRE(lp && loss_out, "dfol_loss_fwd_bwd: null pointer");
  DFOL_REQUIRE(kind == 2 || target, "dfol_loss_fwd_bwd: target missing");
  DFOL_REQUIRE(kind != 1 || seg, "dfol_loss_fwd_bwd: segments missing");
  const long long work = (kind == 1) ? 32ll * n_seg : n_lp;
  if (work == 0) return 0;
  loss_kernel<<<(unsigned)((work + 255) / 256), 256, 0, (cudaStream_t)stream>>>(lp, target, seg, n_seg, n_lp, kind, scale,
                                                                   loss_out, d_lp);
  return finish_launch("dfol_loss_fwd_bwd");
}

extern "C" int dfol_sumsq(const float* g, int64_t n, float* out, void* stream) {
  DFOL_REQUIRE(g && out, "dfol_sumsq: null pointer");
  if (n == 0) return 0;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  sumsq_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, n, out);
  return finish_launch("dfol_sumsq");
}

extern "C" int dfol_l1_regularize(const float* p, float* g, int64_t n, float coef, float loss_coef, float* loss_out,
                                  void* stream) {
  DFOL_REQUIRE(p && g, "dfol_l1_regularize: null pointer");
  if (n == 0) return 0;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  l1_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, n, coef, loss_coef, loss_out);
  return finish_launch("dfol_l1_regularize");
}

extern "C" int dfol_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* sumsq,
                              float clip_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                              int step, void* stream) {
  DFOL_REQUIRE(p && g && m && v, "dfol_adam_step: null pointer");
  DFOL_REQUIRE(clip_norm <= 0.0f || sumsq, "dfol_adam_step: sumsq missing");
  DFOL_REQUIRE(step >= 1, "dfol_adam_step: step must start at 1");
  if (n == 0) return 0;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2 = 1.0f - powf(beta2, (float)step);
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, sumsq, clip_norm, lr, beta1, beta2, eps,
                                                       weight_decay, bc1, sqrtf(bc2));
  return finish_launch("dfol_adam_step");
}

extern "C" int dfol_answers(const float* lp, const int32_t* seg, int question_num, int mode, float threshold,
                            int32_t* first, int32_t* count, float* best_lp, uint8_t* sel, void* stream) {
  DFOL_REQUIRE(lp && first && count && best_lp && (mode == 0 || seg), "dfol_answers: null pointer");
  DFOL_REQUIRE(mode >= 0 && mode <= 2, "dfol_answers: mode must be 0 (binary), 1 (options) or 2 (compare)");
  if (question_num == 0) return 0;
  const int blocks = (question_num * 32 + 255) / 256;
  dfol::answers_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(lp, seg, question_num, mode, threshold, first, count,
                                                                best_lp, sel);
  return dfol::finish_launch("dfol_answers");
}
