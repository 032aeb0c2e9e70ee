// Device building blocks shared by the forward and backward program interpreter kernels.
#pragma once

#include "dfol_common.cuh"

namespace dfol {

#ifdef DFOL_PROGRAM_FAST
constexpr int PROG_THREADS = 512;  // 16 warps per question: twice the rows of a relation tile in flight
#else
constexpr int PROG_THREADS = 256;
#endif
constexpr int PROG_WARPS = PROG_THREADS / 32;
constexpr int MAXN = 128;  // objects per image supported by the interpreter kernels (GQA: <= 100)
constexpr int NCHUNK = MAXN / 32;

struct Instr {
  int op, flags, a0, a1, a2, out, ga0, ga1, gr, mod, mod2;
};

__device__ __forceinline__ Instr load_instr(const int32_t* instr, int ip) {
  const int32_t* w = instr + (long long)ip * DFOL_INSTR_WORDS;
  Instr I;
  I.op = w[DFOL_I_OP]; I.flags = w[DFOL_I_FLAGS]; I.a0 = w[DFOL_I_A0]; I.a1 = w[DFOL_I_A1]; I.a2 = w[DFOL_I_A2];
  I.out = w[DFOL_I_OUT]; I.ga0 = w[DFOL_I_GA0]; I.ga1 = w[DFOL_I_GA1]; I.gr = w[DFOL_I_GR];
  I.mod = w[DFOL_I_MOD]; I.mod2 = w[DFOL_I_MOD2];
  return I;
}

// Attention-transfer modulation of one predicate row (BatchVariableSet.apply_modulations, batch_base_types.py:170-187):
//   out = temp - slog(e^{beta * lnot(L) + slog(1 - d)} + e^{temp}),   temp = alpha * L + slog(c) + slog(d)
// with (alpha, beta, c) = 10 * raw[0..2] and d = raw[3] (the 4-output network of gqa_interpreter_experiments.py:119-131).
struct Mod {
  float alpha, beta, c, d;  // scaled
  float lcd, l1d;           // slog(c) + slog(d), slog(1 - d)
  bool on;
};
__device__ __forceinline__ Mod load_mod(const float* __restrict__ mods, int row) {
  Mod m;
  m.on = (mods != nullptr) && (row >= 0);
  m.alpha = m.beta = m.c = 1.0f; m.d = 0.5f; m.lcd = m.l1d = 0.0f;
  if (m.on) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(mods) + row);
    m.alpha = 10.0f * r.x; m.beta = 10.0f * r.y; m.c = 10.0f * r.z; m.d = r.w;
    m.lcd = slog(m.c) + slog(m.d);
    m.l1d = slog(1.0f - m.d);
  }
  return m;
}
// log(1 + t) for t >= 0 without the cancellation of log(1.0f + t) for small t
__device__ __forceinline__ float mod_log1p(float t) {
#ifdef DFOL_PROGRAM_FAST
  if (t < 0.03125f) return t * (1.0f + t * (-0.5f + t * (0.33333334f + t * (-0.25f + t * 0.2f))));
  return __logf(1.0f + t);
#else
  return log1pf(t);
#endif
}
// out = temp - log(e^u + e^temp) = -log(1 + e^{u - temp}): evaluated in this form, which has no cancellation when the
// result is close to 0 (the reference's fp32 form loses ~6e-8 absolute there, which log(1 - e^x) then amplifies)
__device__ __forceinline__ float mod_apply(const Mod& m, float L) {
  if (!m.on) return L;
  const float temp = m.alpha * L + m.lcd;
  const float u = m.beta * lnot(L) + m.l1d;
  const float z = u - temp;
  if (z > 30.0f) return (u < kLnLogEps && temp < kLnLogEps) ? temp - kLnLogEps : -z;
  return -mod_log1p(DFOL_EXPF(z));
}
// Backward of mod_apply: returns d loss / d L given g = d loss / d out and adds g * d out / d raw[i] to dm[i].
__device__ __forceinline__ float mod_grad(const Mod& m, float L, float g, float dm[4]) {
  if (!m.on) return g;
  const float nl = lnot(L);
  const float temp = m.alpha * L + m.lcd;
  const float u = m.beta * nl + m.l1d;
  // d out / d temp = sigmoid(u - temp) = -d out / d u
  const float z = u - temp;
  const float wu = (z > 30.0f) ? 1.0f : DFOL_DIVF(1.0f, 1.0f + DFOL_EXPF(-z));
  const float dtemp = g * wu;
  const float du = -dtemp;
  dm[0] += 10.0f * dtemp * L;
  dm[1] += 10.0f * du * nl;
  dm[2] += (m.c >= kLogEps) ? 10.0f * dtemp / m.c : 0.0f;
  dm[3] += ((m.d >= kLogEps) ? dtemp / m.d : 0.0f) - ((1.0f - m.d >= kLogEps) ? du / (1.0f - m.d) : 0.0f);
  return dtemp * m.alpha + du * m.beta * lnot_grad(L);
}

// One question's image: table slices of its own image only.
struct Image {
  int n;              // objects
  const float* attr;  // [C][astride]
  int astride;
  const float* rel;   // [nR][rstride], tile [s*n + o]
  int rstride;
};

__device__ __forceinline__ float attr_raw(const Image& im, int col, int t) {
  return __ldg(im.attr + (long long)col * im.astride + t);
}
__device__ __forceinline__ float rel_raw(const Image& im, int col, int s, int o) {
  return __ldg(im.rel + (long long)col * im.rstride + s * im.n + o);
}

// BatchBayesianLogicCell.forward: ll <- min(ll, 0) (batch_base_ops.py:194), then log_parametric_not(ll, neg, 1)
// when any predicate of the op slot is negated (:212-213).
__device__ __forceinline__ float post_ll(float raw, bool neg, bool rt) {
  const float c = fminf(raw, 0.0f);
  if (neg) return lnot(c);
  if (rt) return roundtrip(c);
  return c;
}
// d post_ll / d raw
__device__ __forceinline__ float post_ll_grad(float raw, bool neg, bool rt) {
  if (!(raw < 0.0f)) return 0.0f;
  if (neg) return lnot_grad(raw);
  if (rt) return roundtrip_grad(raw);
  return 1.0f;
}

struct BlockScratch {
  float red[PROG_WARPS];
  __align__(16) float colacc[PROG_WARPS][MAXN];
};

// Sum over all threads of the block; every thread gets the result. Deterministic order.
__device__ __forceinline__ float block_sum(float v, BlockScratch& sc) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sc.red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < PROG_WARPS; ++i) t += sc.red[i];
  return t;
}
__device__ __forceinline__ float block_min(float v, BlockScratch& sc) {
  v = warp_min(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sc.red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = sc.red[0];
#pragma unroll
  for (int i = 1; i < PROG_WARPS; ++i) t = fminf(t, sc.red[i]);
  return t;
}

// Reduce the per-warp column accumulators acc[j] (object lane + 32 j) across warps into dst[0..n).
__device__ __forceinline__ void reduce_columns(const float acc[NCHUNK], int n, float* dst, BlockScratch& sc,
                                               bool accumulate) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NCHUNK; ++j) sc.colacc[w][lane + 32 * j] = acc[j];
  __syncthreads();
  if (threadIdx.x < n) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < PROG_WARPS; ++i) t += sc.colacc[i][threadIdx.x];
    dst[threadIdx.x] = accumulate ? dst[threadIdx.x] + t : t;
  }
  __syncthreads();
}

// log P(exists) over att[0..n): lnot(sum_t lnot(att_t))  (BatchVariableSet.log_probability,
// batch_base_types.py:113-123); hard mode: lnot(min_t lnot(att_t)) (:104-112).  Also returns S.
__device__ __forceinline__ float exists_block(const float* att, int n, bool hard, BlockScratch& sc, float* s_out) {
  const int t = threadIdx.x;
  float s;
  if (hard) s = block_min(t < n ? lnot(att[t]) : 0.0f, sc);
  else s = block_sum(t < n ? lnot(att[t]) : 0.0f, sc);
  if (s_out) *s_out = s;
  return lnot(s);
}

// Likelihood of relation option k of a (possibly normalised) option list at pair (s,o).
// classifier_oracle.py:114-135: ll_k - slog(sum_j exp(ll_j)) per pair when the op slot is normalised.
struct RelOption {
  const Image* im;
  const int32_t* opts;  // option words (column | DFOL_OPT_NEG)
  int count;
  int k;
  bool normalise, roundtrip;
  int single_col;       // >= 0: plain relate on this column (opts unused)
  bool single_neg;

  __device__ __forceinline__ float raw_nrm(int s, int o) const {
    if (single_col >= 0) return rel_raw(*im, single_col, s, o);
    const float r = rel_raw(*im, opts[k] & ~DFOL_OPT_NEG, s, o);
    if (!normalise) return r;
    float den = 0.f;
    for (int j = 0; j < count; ++j) den += DFOL_EXPF(rel_raw(*im, opts[j] & ~DFOL_OPT_NEG, s, o));
    return r - slog(den);
  }
  __device__ __forceinline__ bool neg() const {
    return single_col >= 0 ? single_neg : ((opts[k] & DFOL_OPT_NEG) != 0);
  }
  __device__ __forceinline__ float ll(int s, int o) const { return post_ll(raw_nrm(s, o), neg(), roundtrip); }
};

// Both-role relate posterior restricted to the role that is kept (BatchBayesianLogicCell._forward_core,
// arity 2, batch_base_ops.py:90-149; GQARelateBatch.forward, batch_gqa_ops.py:364-371):
//   subject role: res[s] = a_subj[s] + lnot( sum_{o != s} lnot( ll[s,o] + a_obj[o] ) )
//   object role : res[o] = a_obj[o]  + lnot( sum_{s != o} lnot( ll[s,o] + a_subj[s] ) )
// inner[] receives the inner sums S (needed by the backward pass). Warps stride over rows s, lanes over o.
template <class LL>
__device__ __forceinline__ void relate_forward(int n, const LL& L, const float* a_subj, const float* a_obj,
                                               bool subject_role, float* res, float* inner, BlockScratch& sc) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float acc[NCHUNK];
#pragma unroll
  for (int j = 0; j < NCHUNK; ++j) acc[j] = 0.f;
  for (int s = w; s < n; s += PROG_WARPS) {
    const float as = a_subj[s];
    float rowsum = 0.f;
#pragma unroll
    for (int j = 0; j < NCHUNK; ++j) {
      const int o = lane + 32 * j;
      if (o < n && o != s) {
        const float l = L.ll(s, o);
        if (subject_role) rowsum += lnot(l + a_obj[o]);
        else acc[j] += lnot(l + as);
      }
    }
    if (subject_role) {
      rowsum = warp_sum(rowsum);
      if (lane == 0) { inner[s] = rowsum; res[s] = as + lnot(rowsum); }
    }
  }
  if (!subject_role) {
    reduce_columns(acc, n, inner, sc, false);
    if (threadIdx.x < n) res[threadIdx.x] = a_obj[threadIdx.x] + lnot(inner[threadIdx.x]);
  }
  __syncthreads();
}

#ifdef DFOL_PROGRAM_FAST
// ---------------------------------------------------------------------------------------------------------
// Tensor-core-mode interpreter: the N x N relation tile of every relate hop is streamed into shared memory by a
// bulk-async copy (one elected thread, mbarrier completion) through a ring of tile buffers, so the tile of hop h+1 ..
// h+nbuf-1 is in flight while hop h computes: the table loads never sit on the dependent chain of the attention
// vector.  The hop itself is evaluated in probability space,
//   res[s] = a[s] + slog(1 - prod_{o != s} max(1 - e^{ll[s,o]} e^{a'[o]}, eps))
// (one MUFU.EX2 per pair instead of an exp and a log; identical to sum_o slog(1 - e^{ll+a'}) up to fp32 rounding).
constexpr int MAX_CODE = 48;  // instructions per program staged in shared memory
constexpr int MAX_REL = 64;  // relate hops per program served by the ring (longer programs fall back to direct loads)

struct TileRing {
  float* buf;
  int nbuf, tile_floats;
  uint64_t* full;
};

// post-processed likelihood of a raw tile entry and its derivative w.r.t. the raw entry
__device__ __forceinline__ float tile_post(float raw, bool neg, bool rt) {
  const float c = fminf(raw, 0.0f);
  return neg ? lnot(c) : (rt ? roundtrip(c) : c);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// t = max(1 - e^{post(raw)} * e_other, eps) for the four pairs of one float4 of a tile row
__device__ __forceinline__ float4 tile_terms(float4 r, float4 eo, bool neg, bool rt) {
  float4 t;
  t.x = fmaxf(1.0f - __expf(tile_post(r.x, neg, rt)) * eo.x, kLogEps);
  t.y = fmaxf(1.0f - __expf(tile_post(r.y, neg, rt)) * eo.y, kLogEps);
  t.z = fmaxf(1.0f - __expf(tile_post(r.z, neg, rt)) * eo.z, kLogEps);
  t.w = fmaxf(1.0f - __expf(tile_post(r.w, neg, rt)) * eo.w, kLogEps);
  return t;
}

// Lean forward hop: the products Q_x = prod_{y != x} max(1 - e^{post(ll)} e^{a_other[y]}, eps) of the kept role, left
// in inner[] (subject role) or as per-warp partial products in sc.colacc (object role: relate_kept_q multiplies
// them).  ea[] = e^{a_other} must be visible to the block; ONE barrier, at the end.
__device__ __forceinline__ void relate_tile_products(int n, const float* __restrict__ tile, bool neg, bool rt,
                                                     const float* ea, bool subject_role, float* inner,
                                                     BlockScratch& sc) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if ((n & 3) == 0 && !neg && !rt) {
    constexpr float kLog2e = 1.4426950408889634f;
    const int o4 = 4 * lane;
    const bool act = o4 < n;
    const float* rowp = tile + w * n + o4;
    const int step = PROG_WARPS * n;
    if (subject_role) {
      const float4 eo = act ? *reinterpret_cast<const float4*>(ea + o4) : make_float4(1.f, 1.f, 1.f, 1.f);
      for (int s = w; s < n; s += 2 * PROG_WARPS, rowp += 2 * step) {
        const bool two = s + PROG_WARPS < n;
        float q0 = 1.f, q1 = 1.f;
        if (act) {
          const float4 r0 = *reinterpret_cast<const float4*>(rowp);
          const float4 r1 = two ? *reinterpret_cast<const float4*>(rowp + step)
                                : make_float4(-100.f, -100.f, -100.f, -100.f);
          q0 = (fmaf(-ex2_approx(r0.x * kLog2e), eo.x, 1.0f) *
                fmaf(-ex2_approx(r0.y * kLog2e), eo.y, 1.0f)) *
               (fmaf(-ex2_approx(r0.z * kLog2e), eo.z, 1.0f) *
                fmaf(-ex2_approx(r0.w * kLog2e), eo.w, 1.0f));
          q1 = (fmaf(-ex2_approx(r1.x * kLog2e), eo.x, 1.0f) *
                fmaf(-ex2_approx(r1.y * kLog2e), eo.y, 1.0f)) *
               (fmaf(-ex2_approx(r1.z * kLog2e), eo.z, 1.0f) *
                fmaf(-ex2_approx(r1.w * kLog2e), eo.w, 1.0f));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          q0 *= __shfl_xor_sync(0xffffffffu, q0, o);
          q1 *= __shfl_xor_sync(0xffffffffu, q1, o);
        }
        if (lane == 0) {
          inner[s] = q0;
          if (two) inner[s + PROG_WARPS] = q1;
        }
      }
    } else {
      float4 acc = make_float4(1.f, 1.f, 1.f, 1.f);
      for (int s = w; s < n; s += PROG_WARPS, rowp += step) {
        if (act) {
          const float4 r = *reinterpret_cast<const float4*>(rowp);
          const float es = ea[s];
          acc.x *= fmaf(-ex2_approx(r.x * kLog2e), es, 1.0f);
          acc.y *= fmaf(-ex2_approx(r.y * kLog2e), es, 1.0f);
          acc.z *= fmaf(-ex2_approx(r.z * kLog2e), es, 1.0f);
          acc.w *= fmaf(-ex2_approx(r.w * kLog2e), es, 1.0f);
        }
      }
      if (o4 < MAXN) *reinterpret_cast<float4*>(&sc.colacc[w][o4]) = acc;
    }
  } else {
    float acc[NCHUNK];
#pragma unroll
    for (int j = 0; j < NCHUNK; ++j) acc[j] = 1.f;
    for (int s = w; s < n; s += PROG_WARPS) {
      const float es = ea[s];
      const float* row = tile + s * n;
      float rowprod = 1.f;
#pragma unroll
      for (int j = 0; j < NCHUNK; ++j) {
        const int o = lane + 32 * j;
        if (o < n && o != s) {
          const float t = fmaxf(1.0f - __expf(tile_post(row[o], neg, rt)) * (subject_role ? ea[o] : es), kLogEps);
          if (subject_role) rowprod *= t;
          else acc[j] *= t;
        }
      }
      if (subject_role) {
        rowprod = warp_prod(rowprod);
        if (lane == 0) inner[s] = rowprod;
      }
    }
    if (!subject_role) {
#pragma unroll
      for (int j = 0; j < NCHUNK; ++j) sc.colacc[w][lane + 32 * j] = acc[j];
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float relate_kept_q(int x, bool subject_role, const float* inner, const BlockScratch& sc) {
  if (subject_role) return inner[x];
  float q = 1.f;
#pragma unroll
  for (int i = 0; i < PROG_WARPS; ++i) q *= sc.colacc[i][x];
  return q;
}

// inner[] receives the products Q (NOT their logarithm); ea[] (n floats of scratch) the exponentials of the other
// role's attention, both re-used by relate_backward_tile.
// Fast path (n % 4 == 0): a warp owns a row, lane l the four objects 4l..4l+3 (one LDS.128 per row and lane); the
// self pair needs no test unless the relation is negated: its raw entry is -30, so 1 - e^{-30} e^{a} == 1.0f.
__device__ __forceinline__ void relate_forward_tile(int n, const float* __restrict__ tile, bool neg, bool rt,
                                                    const float* a_subj, const float* a_obj, bool subject_role,
                                                    float* res, float* inner, float* ea, BlockScratch& sc) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x < n) ea[threadIdx.x] = __expf(subject_role ? a_obj[threadIdx.x] : a_subj[threadIdx.x]);
  __syncthreads();
  if ((n & 3) == 0) {
    const int o4 = 4 * lane;
    const bool act = o4 < n;
    const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 eo = (act && subject_role) ? *reinterpret_cast<const float4*>(ea + o4) : one;
    float4 acc = one;
    if (!neg && !rt) {
      // plain relation (the common case): raw <= 0 by construction (log-sigmoid table entries, -30 on self pairs), so
      // min(raw, 0) is the identity and t = 1 - e^raw * e_other needs no clamp in
      // the product (a zero factor gives slog(1 - 0) = 0 exactly as the clamped one does).  Two rows per iteration
      // for instruction-level parallelism; res[] is finished for all rows at once after the loop.
      constexpr float kLog2e = 1.4426950408889634f;
      const float* rowp = tile + w * n + o4;
      const int step = PROG_WARPS * n;
      if (subject_role) {
        for (int s = w; s < n; s += 2 * PROG_WARPS, rowp += 2 * step) {
          const bool two = s + PROG_WARPS < n;
          float q0 = 1.f, q1 = 1.f;
          if (act) {
            const float4 r0 = *reinterpret_cast<const float4*>(rowp);
            const float4 r1 = two ? *reinterpret_cast<const float4*>(rowp + step) : make_float4(-100.f, -100.f, -100.f, -100.f);
            q0 = (fmaf(-ex2_approx(r0.x * kLog2e), eo.x, 1.0f) *
                  fmaf(-ex2_approx(r0.y * kLog2e), eo.y, 1.0f)) *
                 (fmaf(-ex2_approx(r0.z * kLog2e), eo.z, 1.0f) *
                  fmaf(-ex2_approx(r0.w * kLog2e), eo.w, 1.0f));
            q1 = (fmaf(-ex2_approx(r1.x * kLog2e), eo.x, 1.0f) *
                  fmaf(-ex2_approx(r1.y * kLog2e), eo.y, 1.0f)) *
                 (fmaf(-ex2_approx(r1.z * kLog2e), eo.z, 1.0f) *
                  fmaf(-ex2_approx(r1.w * kLog2e), eo.w, 1.0f));
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            q0 *= __shfl_xor_sync(0xffffffffu, q0, o);
            q1 *= __shfl_xor_sync(0xffffffffu, q1, o);
          }
          if (lane == 0) {
            inner[s] = q0;
            if (two) inner[s + PROG_WARPS] = q1;
          }
        }
        __syncthreads();
        if (threadIdx.x < n) res[threadIdx.x] = a_subj[threadIdx.x] + slog(1.0f - inner[threadIdx.x]);
        __syncthreads();
        return;
      }
      for (int s = w; s < n; s += PROG_WARPS, rowp += step) {
        if (act) {
          const float4 r = *reinterpret_cast<const float4*>(rowp);
          const float es = ea[s];
          acc.x *= fmaf(-ex2_approx(r.x * kLog2e), es, 1.0f);
          acc.y *= fmaf(-ex2_approx(r.y * kLog2e), es, 1.0f);
          acc.z *= fmaf(-ex2_approx(r.z * kLog2e), es, 1.0f);
          acc.w *= fmaf(-ex2_approx(r.w * kLog2e), es, 1.0f);
        }
      }
    } else {
      for (int s = w; s < n; s += PROG_WARPS) {
        float4 t = one;
        if (act) {
          const float4 r = *reinterpret_cast<const float4*>(tile + s * n + o4);
          const float es = ea[s];
          t = tile_terms(r, subject_role ? eo : make_float4(es, es, es, es), neg, rt);
          if (neg && (s >> 2) == lane) {  // negated relation: the self pair must be skipped explicitly
            const int d = s & 3;
            if (d == 0) t.x = 1.f; else if (d == 1) t.y = 1.f; else if (d == 2) t.z = 1.f; else t.w = 1.f;
          }
        }
        if (subject_role) {
          const float q = warp_prod((t.x * t.y) * (t.z * t.w));
          if (lane == 0) { inner[s] = q; res[s] = a_subj[s] + slog(1.0f - q); }
        } else {
          acc.x *= t.x; acc.y *= t.y; acc.z *= t.z; acc.w *= t.w;
        }
      }
    }
    if (!subject_role) {
      __syncthreads();
      *reinterpret_cast<float4*>(&sc.colacc[w][o4]) = acc;
      __syncthreads();
      if (threadIdx.x < n) {
        float q = 1.f;
#pragma unroll
        for (int i = 0; i < PROG_WARPS; ++i) q *= sc.colacc[i][threadIdx.x];
        inner[threadIdx.x] = q;
        res[threadIdx.x] = a_obj[threadIdx.x] + slog(1.0f - q);
      }
    }
    __syncthreads();
    return;
  }
  float acc[NCHUNK];
#pragma unroll
  for (int j = 0; j < NCHUNK; ++j) acc[j] = 1.f;
  for (int s = w; s < n; s += PROG_WARPS) {
    const float es = ea[s];
    const float* row = tile + s * n;
    float rowprod = 1.f;
#pragma unroll
    for (int j = 0; j < NCHUNK; ++j) {
      const int o = lane + 32 * j;
      if (o < n && o != s) {
        const float p = __expf(tile_post(row[o], neg, rt)) * (subject_role ? ea[o] : es);
        const float t = fmaxf(1.0f - p, kLogEps);
        if (subject_role) rowprod *= t;
        else acc[j] *= t;
      }
    }
    if (subject_role) {
      rowprod = warp_prod(rowprod);
      if (lane == 0) { inner[s] = rowprod; res[s] = a_subj[s] + slog(1.0f - rowprod); }
    }
  }
  if (!subject_role) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NCHUNK; ++j) sc.colacc[w][lane + 32 * j] = acc[j];
    __syncthreads();
    if (threadIdx.x < n) {
      float q = 1.f;
#pragma unroll
      for (int i = 0; i < PROG_WARPS; ++i) q *= sc.colacc[i][threadIdx.x];
      inner[threadIdx.x] = q;
      res[threadIdx.x] = a_obj[threadIdx.x] + slog(1.0f - q);
    }
  }
  __syncthreads();
}

// d slog(1 - Q) / d log Q given the product Q
__device__ __forceinline__ float lnot_grad_q(float q) {
  const float u = 1.0f - q;
  return (u >= kLogEps) ? __fdividef(-q, u) : 0.0f;
}

// du/d(l+a) factor -p/(1-p) (zero where the clamp is active) of one pair, p = e^{post(raw)} * e_other
__device__ __forceinline__ float tile_lgrad(float raw, float eo, bool neg, bool rt) {
  const float p = __expf(tile_post(raw, neg, rt)) * eo;
  const float u = 1.0f - p;
  return (u >= kLogEps) ? __fdividef(-p, u) : 0.0f;
}

// Backward of relate_forward_tile (same contract as relate_backward); dq[] is n floats of scratch.
__device__ __forceinline__ void relate_backward_tile(int n, const float* __restrict__ tile, bool neg, bool rt,
                                                     const float* a_subj, bool subject_role, const float* dres,
                                                     const float* inner, const float* ea, float* dq, float* g_other,
                                                     float* __restrict__ gslice, BlockScratch& sc) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x < n) dq[threadIdx.x] = dres[threadIdx.x] * lnot_grad_q(inner[threadIdx.x]);
  __syncthreads();
  if ((n & 3) == 0) {
    const int o4 = 4 * lane;
    const bool act = o4 < n;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 eo = (act && subject_role) ? *reinterpret_cast<const float4*>(ea + o4) : zero;
    const float4 dqo = (act && !subject_role) ? *reinterpret_cast<const float4*>(dq + o4) : zero;
    float4 acc = zero;
    for (int s = w; s < n; s += PROG_WARPS) {
      float rowsum = 0.f;
      if (act) {
        const float4 r = *reinterpret_cast<const float4*>(tile + s * n + o4);
        const float es = ea[s], dS_row = dq[s];
        float4 du;
        if (subject_role) {
          du.x = dS_row * tile_lgrad(r.x, eo.x, neg, rt);
          du.y = dS_row * tile_lgrad(r.y, eo.y, neg, rt);
          du.z = dS_row * tile_lgrad(r.z, eo.z, neg, rt);
          du.w = dS_row * tile_lgrad(r.w, eo.w, neg, rt);
        } else {
          du.x = dqo.x * tile_lgrad(r.x, es, neg, rt);
          du.y = dqo.y * tile_lgrad(r.y, es, neg, rt);
          du.z = dqo.z * tile_lgrad(r.z, es, neg, rt);
          du.w = dqo.w * tile_lgrad(r.w, es, neg, rt);
        }
        if ((s >> 2) == lane) {  // self pair: no gradient
          const int d = s & 3;
          if (d == 0) du.x = 0.f; else if (d == 1) du.y = 0.f; else if (d == 2) du.z = 0.f; else du.w = 0.f;
        }
        if (subject_role) { acc.x += du.x; acc.y += du.y; acc.z += du.z; acc.w += du.w; }
        else rowsum = (du.x + du.y) + (du.z + du.w);
        float4 dn;
        dn.x = du.x * post_ll_grad(r.x, neg, rt);
        dn.y = du.y * post_ll_grad(r.y, neg, rt);
        dn.z = du.z * post_ll_grad(r.z, neg, rt);
        dn.w = du.w * post_ll_grad(r.w, neg, rt);
        *reinterpret_cast<float4*>(gslice + s * n + o4) = dn;
      }
      if (!subject_role) {
        rowsum = warp_sum(rowsum);
        if (lane == 0) g_other[s] += rowsum;
      }
    }
    if (subject_role) {
      __syncthreads();
      *reinterpret_cast<float4*>(&sc.colacc[w][o4]) = acc;
      __syncthreads();
      if (threadIdx.x < n) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < PROG_WARPS; ++i) t += sc.colacc[i][threadIdx.x];
        g_other[threadIdx.x] += t;
      }
    }
    __syncthreads();
    return;
  }
  float acc[NCHUNK];
#pragma unroll
  for (int j = 0; j < NCHUNK; ++j) acc[j] = 0.f;
  for (int s = w; s < n; s += PROG_WARPS) {
    const float es = ea[s];
    const float dS_row = dq[s];
    const float* row = tile + s * n;
    float rowsum = 0.f;
#pragma unroll
    for (int j = 0; j < NCHUNK; ++j) {
      const int o = lane + 32 * j;
      if (o < n) {
        float dn = 0.0f;
        if (o != s) {
          const float raw = row[o];
          const float lg = tile_lgrad(raw, subject_role ? ea[o] : es, neg, rt);
          float du;
          if (subject_role) {
            du = dS_row * lg;
            acc[j] += du;
          } else {
            du = dq[o] * lg;
            rowsum += du;
          }
          dn = du * post_ll_grad(raw, neg, rt);
        }
        gslice[s * n + o] = dn;
      }
    }
    if (!subject_role) {
      rowsum = warp_sum(rowsum);
      if (lane == 0) g_other[s] += rowsum;
    }
  }
  if (subject_role) reduce_columns(acc, n, g_other, sc, true);
  __syncthreads();
}
#endif  // DFOL_PROGRAM_FAST

}  // namespace dfol
