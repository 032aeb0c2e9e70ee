// Shared device/host helpers for libdfol_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "dfol_b200.h"

namespace dfol {

constexpr float kDefaultLL = -30.0f;  // default_log_likelihood of the reference (base_oracle.py:18)
constexpr float kLogEps = 1e-20f;     // util.safe_log clamp (util.py:25)

void set_error(const char* fmt, ...);
int finish_launch(const char* what);

#define DFOL_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      dfol::set_error(__VA_ARGS__);    \
      return -1;                       \
    }                                  \
  } while (0)

// ---- log-space primitives, written to round like the reference's fp32 torch ops (util.py:22-47) ----
__device__ __forceinline__ float slog(float x) { return logf(fmaxf(x, kLogEps)); }
__device__ __forceinline__ float lnot(float x) { return slog(1.0f - expf(x)); }
// d/dx slog(1 - e^x): zero where the clamp is active (torch clamp backward)
__device__ __forceinline__ float lnot_grad(float x) {
  float e = expf(x);
  float u = 1.0f - e;
  return (u >= kLogEps) ? (-e / u) : 0.0f;
}
// slog(exp(x)) "round trip" (log_parametric_not with alpha = 0) and its derivative
__device__ __forceinline__ float roundtrip(float x) { return slog(expf(x)); }
__device__ __forceinline__ float roundtrip_grad(float x) { return (expf(x) >= kLogEps) ? 1.0f : 0.0f; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case DFOL_ACT_ELU: return x > 0.0f ? x : expm1f(x);
    case DFOL_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case DFOL_ACT_LOGSIGMOID: return fminf(x, 0.0f) - log1pf(expf(-fabsf(x)));
    default: return x;
  }
}
// act'(z) expressed with the saved output h = act(z)
__device__ __forceinline__ float act_grad_from_output(float h, int act) {
  switch (act) {
    case DFOL_ACT_ELU: return h > 0.0f ? 1.0f : h + 1.0f;
    case DFOL_ACT_SIGMOID: return h * (1.0f - h);
    case DFOL_ACT_LOGSIGMOID: return 1.0f - expf(h);
    default: return 1.0f;
  }
}

}  // namespace dfol
