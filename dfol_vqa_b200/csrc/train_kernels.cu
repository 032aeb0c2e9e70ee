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
