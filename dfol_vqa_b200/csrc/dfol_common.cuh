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
// A translation unit compiled with DFOL_PROGRAM_FAST (the tensor-core mode's interpreter) uses the MUFU
// approximations instead of the accurate libdevice functions.
#ifdef DFOL_PROGRAM_FAST
#define DFOL_EXPF __expf
#define DFOL_LOGF __logf
#define DFOL_DIVF __fdividef
#else
#define DFOL_EXPF expf
#define DFOL_LOGF logf
#define DFOL_DIVF(a, b) ((a) / (b))
#endif
constexpr float kLnLogEps = -46.051702f;  // ln(1e-20)
__device__ __forceinline__ float slog(float x) { return DFOL_LOGF(fmaxf(x, kLogEps)); }
#ifdef DFOL_PROGRAM_FAST
// log(1 - e^x): __logf has an ABSOLUTE error of ~4e-7 near 1, far too coarse for the log of a probability close to
// one (sums of ~50 terms of size 1e-7 feed the quantifiers).  u = 1 - p is rounded to fp32 exactly as the
// reference's torch ops round it (p < 6e-8 flushes to log 1 = 0), then log(u) = log1p(u - 1) by its series.
__device__ __forceinline__ float lnot(float x) {
  const float u = 1.0f - __expf(x);
  const float d = u - 1.0f;  // exact
  if (d > -0.03125f) return d * (1.0f + d * (-0.5f + d * (0.33333334f + d * (-0.25f + d * 0.2f))));
  return __logf(fmaxf(u, kLogEps));
}
#else
__device__ __forceinline__ float lnot(float x) { return slog(1.0f - expf(x)); }
#endif
// d/dx slog(1 - e^x): zero where the clamp is active (torch clamp backward)
__device__ __forceinline__ float lnot_grad(float x) {
  float e = DFOL_EXPF(x);
  float u = 1.0f - e;
  return (u >= kLogEps) ? DFOL_DIVF(-e, u) : 0.0f;
}
// slog(exp(x)) "round trip" (log_parametric_not with alpha = 0) and its derivative
#ifdef DFOL_PROGRAM_FAST
__device__ __forceinline__ float roundtrip(float x) { return fmaxf(x, kLnLogEps); }
__device__ __forceinline__ float roundtrip_grad(float x) { return (x >= kLnLogEps) ? 1.0f : 0.0f; }
#else
__device__ __forceinline__ float roundtrip(float x) { return slog(expf(x)); }
__device__ __forceinline__ float roundtrip_grad(float x) { return (expf(x) >= kLogEps) ? 1.0f : 0.0f; }
#endif

// ---- shared-memory addresses and mbarriers (TMA / bulk-async completion) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (long long spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spin > (1ll << 26)) __trap();  // never hang the device: a lost arrival becomes a launch error
  }
}
// 1-D bulk-async copy global -> shared (TMA engine, no tensor map): bytes %% 16 == 0, both addresses 16-byte aligned;
// completion is signalled on `bar` as transaction bytes.
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_prod(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
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
