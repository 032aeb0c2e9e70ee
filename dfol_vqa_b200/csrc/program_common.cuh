// Device building blocks shared by the forward and backward program interpreter kernels.
#pragma once

#include "dfol_common.cuh"

namespace dfol {

#ifdef DFOL_PROGRAM_FAST
#ifndef DFOL_PROG_THREADS
#define DFOL_PROG_THREADS 512
#endif
constexpr int PROG_THREADS = DFOL_PROG_THREADS;  // 16 warps per question: twice the rows of a relation tile in flight
#ifndef DFOL_PROG_MIN_BLOCKS
#define DFOL_PROG_MIN_BLOCKS (1024 / DFOL_PROG_THREADS)
#endif
constexpr int PROG_MIN_BLOCKS = DFOL_PROG_MIN_BLOCKS;  // two questions per SM (64 registers per thread)
#else
constexpr int PROG_THREADS = 256;
constexpr int PROG_MIN_BLOCKS = 1;
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

// Option lists (query / choose / verify / same over up to ~1750 attribute options): a warp takes U options per pass and
// issues all U table rows before it touches any of them -- U independent global loads in flight per lane instead of a
// chain of (word -> row) round trips per option (a 1749-option query spent ~110 dependent HBM latencies per warp and
// pass; the probability-space query path below additionally keeps the rows of pass i+1 in flight under pass i).  The option words come from
// `op`, which the tensor-core build points at a shared-memory copy of the list (stage_options), so that no row address
// waits on a global load.  NC = 32-object chunks per row (2 for images of <= 64 objects).  Options are visited in the
// same order as a plain strided loop, so per-warp accumulations round identically.
//   body(k, word, raw): warp-uniform call for option k; raw[j] = table entry of object lane + 32 j (0 beyond n).
template <int U, int NC>
__device__ __forceinline__ void options_load(const Image& im, const int32_t* op, int count, int k0, float (&dst)[U][NC]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int k = k0 + u * PROG_WARPS;
    const bool ok = k < count;
    const int word = ok ? op[k] : 0;
    const float* row = im.attr + (long long)(word & ~DFOL_OPT_NEG) * im.astride;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int t = lane + 32 * j;
      dst[u][j] = (ok && t < im.n) ? __ldg(row + t) : 0.0f;
    }
  }
}
template <int U, int NC, class Body>
__device__ __forceinline__ void for_options_nc(const Image& im, const int32_t* op, int count, Body body) {
  const int w = threadIdx.x >> 5;
  for (int k0 = w; k0 < count; k0 += PROG_WARPS * U) {
    float raw[U][NC];
    options_load<U, NC>(im, op, count, k0, raw);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + u * PROG_WARPS;
      if (k < count) body(k, op[k], raw[u]);
    }
  }
}
template <int U, class Body>
__device__ __forceinline__ void for_options(const Image& im, const int32_t* op, int count, Body body) {
  if (im.n <= 64) for_options_nc<U, 2>(im, op, count, body);
  else for_options_nc<(U > 4 ? 4 : U), NCHUNK>(im, op, count, body);
}
#define DFOL_NC_OF(raw) ((int)(sizeof(raw) / sizeof(float)))

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
// bulk-async copy (one elected thread, mbarrier completion) through a ring of tile buffers, so the tiles of hops
// h+1 .. h+nbuf-1 are in flight while hop h computes: the table loads never sit on the dependent chain of the
// attention vector.  The hop itself is evaluated in probability space,
//   res[x] = prior[x] + slog(1 - prod_{y != x} (1 - p[x,y] e^{a[y]}))
// (identical to sum_y slog(1 - e^{ll + a}) up to fp32 rounding).  When the scene supplies the PROBABILITY table
// (PTAB: p = e^{ll} written next to ll by the slot kernels, 0 on self pairs) a pair costs one FFMA and one FMUL and no
// MUFU; with the log table alone it costs one MUFU.EX2 more.
constexpr int MAX_CODE = 48;  // instructions per program staged in shared memory
constexpr int MAX_REL = 64;   // relate hops per program served by the ring (longer programs fall back to direct loads)

struct TileRing {
  float* buf;
  int nbuf, tile_floats;
  uint64_t* full;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Geometry of a hop: LPR lanes share a tile row (lane l of the group owns the objects 4l .. 4l+3), a warp holds
// RS = 32 / LPR row groups, the block NS = 16 RS of them; group g owns the rows g, g + NS, ... (at most NR).
//   n <= 32: LPR 8, NR 1;  n <= 64: LPR 16, NR 2;  n <= 128: LPR 32, NR 8.
// The NR per-row partial results of a lane are reduced over the LPR lanes of the group by a TRANSPOSED butterfly: each
// of the first log2(NR) steps halves the number of values a lane carries, so 8 rows cost 9 shuffles instead of 40.
// On return every lane holds the complete result of row index `ri` of its group.
template <int LPR, int NR, bool PROD>
__device__ __forceinline__ float rows_reduce(float (&v)[NR], int l, int& ri) {
  constexpr unsigned FULL = 0xffffffffu;
  ri = 0;
  int cnt = NR;
#pragma unroll
  for (int d = LPR / 2; d >= 1; d >>= 1) {
    if (cnt > 1) {
      const bool hi = (l & d) != 0;
      cnt >>= 1;
#pragma unroll
      for (int i = 0; i < NR / 2; ++i) {
        if (i < cnt) {
          const float keep = hi ? v[i + cnt] : v[i];
          const float send = hi ? v[i] : v[i + cnt];
          const float got = __shfl_xor_sync(FULL, send, d);
          v[i] = PROD ? keep * got : keep + got;
        }
      }
      ri += hi ? cnt : 0;
    } else {
      const float got = __shfl_xor_sync(FULL, v[0], d);
      v[0] = PROD ? v[0] * got : v[0] + got;
    }
  }
  return v[0];
}

// per-column accumulators of the row groups of one warp combined (lanes with the same l), then parked in colacc[w]
template <int LPR, bool PROD>
__device__ __forceinline__ void park_columns(float4 acc, int lane, int w, BlockScratch& sc) {
  constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
  for (int d = LPR; d < 32; d <<= 1) {
    const float gx = __shfl_xor_sync(FULL, acc.x, d), gy = __shfl_xor_sync(FULL, acc.y, d);
    const float gz = __shfl_xor_sync(FULL, acc.z, d), gw = __shfl_xor_sync(FULL, acc.w, d);
    if (PROD) { acc.x *= gx; acc.y *= gy; acc.z *= gz; acc.w *= gw; }
    else { acc.x += gx; acc.y += gy; acc.z += gz; acc.w += gw; }
  }
  if (lane < LPR) *reinterpret_cast<float4*>(&sc.colacc[w][4 * lane]) = acc;
}

// Probabilities p'[s, 4l .. 4l+3] of one tile row after the logic cell's post-processing: p (plain / round trip, which
// differs from the identity only below 1e-20) or max(1 - p, eps) (negated relation); zero outside the image and on the
// self pair, so that neither the products nor the gradients need a test.  `orig` receives the un-negated p (NEG only).
template <bool PTAB, bool NEG, bool ALIGNED>
__device__ __forceinline__ float4 hop_row(const float* __restrict__ tile, int n, int s, int l, bool row_ok,
                                          float4* orig) {
  constexpr float kLog2e = 1.4426950408889634f;
  const int o0 = 4 * l;
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (orig) *orig = r;
  if (!(row_ok && o0 < n)) return r;
  const float* src = tile + s * n + o0;
  bool v1 = true, v2 = true, v3 = true;
  if (ALIGNED) {
    r = *reinterpret_cast<const float4*>(src);
  } else {
    v1 = o0 + 1 < n; v2 = o0 + 2 < n; v3 = o0 + 3 < n;
    r.x = src[0];
    r.y = v1 ? src[1] : 0.f;
    r.z = v2 ? src[2] : 0.f;
    r.w = v3 ? src[3] : 0.f;
  }
  if (PTAB && !NEG) return r;  // zeros outside the row and on the diagonal already
  if (!PTAB) {
    r.x = ex2_approx(r.x * kLog2e); r.y = ex2_approx(r.y * kLog2e);
    r.z = ex2_approx(r.z * kLog2e); r.w = ex2_approx(r.w * kLog2e);
  }
  if (NEG) {
    if (orig) *orig = r;
    r.x = fmaxf(1.0f - r.x, kLogEps); r.y = fmaxf(1.0f - r.y, kLogEps);
    r.z = fmaxf(1.0f - r.z, kLogEps); r.w = fmaxf(1.0f - r.w, kLogEps);
  }
  if (!ALIGNED) {
    if (!v1) r.y = 0.f;
    if (!v2) r.z = 0.f;
    if (!v3) r.w = 0.f;
  }
  if ((s >> 2) == l) {  // self pair
    const int d = s & 3;
    if (d == 0) r.x = 0.f; else if (d == 1) r.y = 0.f; else if (d == 2) r.z = 0.f; else r.w = 0.f;
  }
  return r;
}

template <bool ALIGNED>
__device__ __forceinline__ void hop_store4(float* __restrict__ dst, int n, int o0, float4 v) {
  if (ALIGNED) {
    *reinterpret_cast<float4*>(dst) = v;
  } else {
    dst[0] = v.x;
    if (o0 + 1 < n) dst[1] = v.y;
    if (o0 + 2 < n) dst[2] = v.z;
    if (o0 + 3 < n) dst[3] = v.w;
  }
}

// Forward hop.  ea[0..MAXN) = e^{attention of the other role} (zero beyond n), prior[] = the kept role's prior.
// finish(x, Q) is called once per object x of the kept role with Q = prod_y (1 - p'[x,y] ea[y]), by the lane that
// ends up holding the row (subject role) or by thread x (object role); it must not read anything another thread's
// finish writes.  Ends with a block barrier (all tile reads and all finish calls done).
template <int LPR, int NR, bool PTAB, bool NEG, bool ALIGNED, class Finish>
__device__ __forceinline__ void hop_forward_t(int n, const float* __restrict__ tile, const float* ea, bool subject_role,
                                              BlockScratch& sc, Finish finish) {
  constexpr int RS = 32 / LPR, NS = PROG_WARPS * RS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int l = lane % LPR, g = w * RS + lane / LPR;
  // Common case (probability tiles, plain relation, n % 4 == 0): NO predication.  Rows / columns outside the image read
  // a clamped (valid, finite) address and are neutralised by the zero entries of ea[] beyond n (factor 1 - p * 0 = 1);
  // the products of row slots >= n are computed and dropped.
  constexpr bool LEAN = PTAB && !NEG && ALIGNED;
  const float* lean_col = tile + min(4 * l, n - 4);
  if (subject_role) {
    const float4 eo = *reinterpret_cast<const float4*>(ea + 4 * l);
    float q[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int s = g + NS * i;
      q[i] = 1.0f;
      if (LPR == 32 && s >= n) continue;   // (a warp = one row group: warp-uniform)
      const float4 r = LEAN ? *reinterpret_cast<const float4*>(lean_col + min(s, n - 1) * n)
                            : hop_row<PTAB, NEG, ALIGNED>(tile, n, s, l, s < n, nullptr);
      q[i] = (fmaf(-r.x, eo.x, 1.0f) * fmaf(-r.y, eo.y, 1.0f)) * (fmaf(-r.z, eo.z, 1.0f) * fmaf(-r.w, eo.w, 1.0f));
    }
    int ri;
    const float Q = rows_reduce<LPR, NR, true>(q, l, ri);
    const int s = g + NS * ri;
    if ((l & (LPR / NR - 1)) == 0 && s < n) finish(s, Q);
  } else {
    float4 acc = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int s = g + NS * i;   // s < MAXN always: ea[s] is zero for s >= n
      if (LPR == 32 && s >= n) continue;
      const float4 r = LEAN ? *reinterpret_cast<const float4*>(lean_col + min(s, n - 1) * n)
                            : hop_row<PTAB, NEG, ALIGNED>(tile, n, s, l, s < n, nullptr);
      const float es = LEAN ? ea[s] : ((s < n) ? ea[s] : 0.0f);
      acc.x *= fmaf(-r.x, es, 1.0f); acc.y *= fmaf(-r.y, es, 1.0f);
      acc.z *= fmaf(-r.z, es, 1.0f); acc.w *= fmaf(-r.w, es, 1.0f);
    }
    park_columns<LPR, true>(acc, lane, w, sc);
    __syncthreads();
    if (threadIdx.x < n) {
      float Q = 1.f;
#pragma unroll
      for (int i = 0; i < PROG_WARPS; ++i) Q *= sc.colacc[i][threadIdx.x];
      finish((int)threadIdx.x, Q);
    }
  }
  __syncthreads();
}

template <bool PTAB, class Finish>
__device__ __forceinline__ void hop_forward(int n, const float* __restrict__ tile, bool neg, const float* ea,
                                            bool subject_role, BlockScratch& sc, Finish finish) {
#define DFOL_HOP_FWD(LPR, NR)                                                                              \
  {                                                                                                        \
    if ((n & 3) == 0) {                                                                                    \
      if (!neg) hop_forward_t<LPR, NR, PTAB, false, true>(n, tile, ea, subject_role, sc, finish);         \
      else hop_forward_t<LPR, NR, PTAB, true, true>(n, tile, ea, subject_role, sc, finish);               \
    } else {                                                                                               \
      if (!neg) hop_forward_t<LPR, NR, PTAB, false, false>(n, tile, ea, subject_role, sc, finish);        \
      else hop_forward_t<LPR, NR, PTAB, true, false>(n, tile, ea, subject_role, sc, finish);              \
    }                                                                                                      \
  }
  if (n <= 32) DFOL_HOP_FWD(8, (32 + 4 * PROG_WARPS - 1) / (4 * PROG_WARPS))
  else if (n <= 64) DFOL_HOP_FWD(16, 64 / (2 * PROG_WARPS))
  else DFOL_HOP_FWD(32, MAXN / PROG_WARPS)
#undef DFOL_HOP_FWD
}

// c = dres * Q / (1 - Q): d loss / d(-log Q)... the common factor of the pair gradients of one kept object
// (du[x,y] = c[x] * m / (1 - m), m = p'[x,y] ea[y]); zero where the clamp of slog(1 - Q) is active.
__device__ __forceinline__ float hop_cfactor(float dres, float Q) {
  const float u = 1.0f - Q;
  return (u >= kLogEps) ? dres * __fdividef(Q, u) : 0.0f;
}

// d loss / d(l + a) of the four pairs of a lane: ce = c * ea per pair; a vanished factor (1 - m == 0) made Q and c zero,
// so the clamped reciprocal yields 0 exactly as the reference's clamp backward does
__device__ __forceinline__ float4 hop_du(float4 r, float4 ce, float4 e) {
  float4 du;
  du.x = (ce.x * r.x) * rcp_approx(fmaxf(fmaf(-r.x, e.x, 1.0f), kLogEps));
  du.y = (ce.y * r.y) * rcp_approx(fmaxf(fmaf(-r.y, e.y, 1.0f), kLogEps));
  du.z = (ce.z * r.z) * rcp_approx(fmaxf(fmaf(-r.z, e.z, 1.0f), kLogEps));
  du.w = (ce.w * r.w) * rcp_approx(fmaxf(fmaf(-r.w, e.w, 1.0f), kLogEps));
  return du;
}
// gradient w.r.t. the RAW table entry: identity for a plain relation, -p / (1 - p) through the negation
template <bool NEG>
__device__ __forceinline__ float4 hop_draw(float4 du, float4 p) {
  if (!NEG) return du;
  float4 dn;
  dn.x = (1.0f - p.x >= kLogEps) ? du.x * __fdividef(-p.x, 1.0f - p.x) : 0.0f;
  dn.y = (1.0f - p.y >= kLogEps) ? du.y * __fdividef(-p.y, 1.0f - p.y) : 0.0f;
  dn.z = (1.0f - p.z >= kLogEps) ? du.z * __fdividef(-p.z, 1.0f - p.z) : 0.0f;
  dn.w = (1.0f - p.w >= kLogEps) ? du.w * __fdividef(-p.w, 1.0f - p.w) : 0.0f;
  return dn;
}

// Backward hop (re-evaluates the products).  cfac(x, Q) -> c[x] is called once per kept object (same calling convention
// as `finish` above; it may write per-object results for the caller).  cbuf: MAXN floats of scratch; g_other[0..n)
// receives d loss / d(attention of the other role); gslice (n x n, global) d loss / d raw table entry.
// Ends with a block barrier.
template <int LPR, int NR, bool PTAB, bool NEG, bool ALIGNED, class CFac>
__device__ __forceinline__ void hop_backward_t(int n, const float* __restrict__ tile, const float* ea,
                                               bool subject_role, float* cbuf, float* g_other,
                                               float* __restrict__ gslice, BlockScratch& sc, CFac cfac) {
  constexpr int RS = 32 / LPR, NS = PROG_WARPS * RS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int l = lane % LPR, g = w * RS + lane / LPR;
  const int o0 = 4 * l;
  // (LEAN: see hop_forward_t -- clamped addresses instead of predication; cbuf[] and ea[] are zero beyond n)
  constexpr bool LEAN = PTAB && !NEG && ALIGNED;
  const float* lean_col = tile + min(o0, n - 4);
  if (subject_role) {
    const float4 eo = *reinterpret_cast<const float4*>(ea + o0);
    {
      float q[NR];
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int s = g + NS * i;
        q[i] = 1.0f;
        if (LPR == 32 && s >= n) continue;   // warp-uniform
        const float4 r = LEAN ? *reinterpret_cast<const float4*>(lean_col + min(s, n - 1) * n)
                              : hop_row<PTAB, NEG, ALIGNED>(tile, n, s, l, s < n, nullptr);
        q[i] = (fmaf(-r.x, eo.x, 1.0f) * fmaf(-r.y, eo.y, 1.0f)) * (fmaf(-r.z, eo.z, 1.0f) * fmaf(-r.w, eo.w, 1.0f));
      }
      int ri;
      const float Q = rows_reduce<LPR, NR, true>(q, l, ri);
      const int s = g + NS * ri;
      if ((l & (LPR / NR - 1)) == 0 && s < n) cbuf[s] = cfac(s, Q);
    }
    __syncwarp();  // the rows of a group are produced and consumed inside one warp
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int s = g + NS * i;
      if (LPR == 32 && s >= n) continue;   // warp-uniform
      if (LEAN) {
        const float4 r = *reinterpret_cast<const float4*>(lean_col + min(s, n - 1) * n);
        const float c = cbuf[s];   // zero for s >= n
        const float4 du = hop_du(r, make_float4(c * eo.x, c * eo.y, c * eo.z, c * eo.w), eo);
        acc.x += du.x; acc.y += du.y; acc.z += du.z; acc.w += du.w;
        if (s < n && o0 < n) *reinterpret_cast<float4*>(gslice + s * n + o0) = du;
      } else if (s < n && o0 < n) {
        float4 p0;
        const float4 r = hop_row<PTAB, NEG, ALIGNED>(tile, n, s, l, true, &p0);
        const float c = cbuf[s];
        const float4 du = hop_du(r, make_float4(c * eo.x, c * eo.y, c * eo.z, c * eo.w), eo);
        acc.x += du.x; acc.y += du.y; acc.z += du.z; acc.w += du.w;
        hop_store4<ALIGNED>(gslice + s * n + o0, n, o0, hop_draw<NEG>(du, p0));
      }
    }
    park_columns<LPR, false>(acc, lane, w, sc);
    __syncthreads();
    if (threadIdx.x < n) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < PROG_WARPS; ++i) t += sc.colacc[i][threadIdx.x];
      g_other[threadIdx.x] = t;
    }
  } else {
    {
      float4 acc = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int s = g + NS * i;
        if (LPR == 32 && s >= n) continue;   // warp-uniform
        const float4 r = LEAN ? *reinterpret_cast<const float4*>(lean_col + min(s, n - 1) * n)
                              : hop_row<PTAB, NEG, ALIGNED>(tile, n, s, l, s < n, nullptr);
        const float es = LEAN ? ea[s] : ((s < n) ? ea[s] : 0.0f);
        acc.x *= fmaf(-r.x, es, 1.0f); acc.y *= fmaf(-r.y, es, 1.0f);
        acc.z *= fmaf(-r.z, es, 1.0f); acc.w *= fmaf(-r.w, es, 1.0f);
      }
      park_columns<LPR, true>(acc, lane, w, sc);
    }
    __syncthreads();
    if (threadIdx.x < MAXN) {
      float c = 0.f;
      if (threadIdx.x < n) {
        float Q = 1.f;
#pragma unroll
        for (int i = 0; i < PROG_WARPS; ++i) Q *= sc.colacc[i][threadIdx.x];
        c = cfac((int)threadIdx.x, Q);
      }
      cbuf[threadIdx.x] = c;
    }
    __syncthreads();
    const float4 c4 = *reinterpret_cast<const float4*>(cbuf + o0);
    float rs[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int s = g + NS * i;
      rs[i] = 0.f;
      if (LPR == 32 && s >= n) continue;   // warp-uniform
      if (LEAN) {   // c4 is zero for objects >= n, ea[s] for rows >= n
        const float4 r = *reinterpret_cast<const float4*>(lean_col + min(s, n - 1) * n);
        const float es = ea[s];
        const float4 du = hop_du(r, make_float4(c4.x * es, c4.y * es, c4.z * es, c4.w * es),
                                 make_float4(es, es, es, es));
        rs[i] = (du.x + du.y) + (du.z + du.w);
        if (s < n && o0 < n) *reinterpret_cast<float4*>(gslice + s * n + o0) = du;
      } else if (s < n && o0 < n) {
        float4 p0;
        const float4 r = hop_row<PTAB, NEG, ALIGNED>(tile, n, s, l, true, &p0);
        const float es = ea[s];
        const float4 du = hop_du(r, make_float4(c4.x * es, c4.y * es, c4.z * es, c4.w * es),
                                 make_float4(es, es, es, es));
        rs[i] = (du.x + du.y) + (du.z + du.w);
        hop_store4<ALIGNED>(gslice + s * n + o0, n, o0, hop_draw<NEG>(du, p0));
      }
    }
    int ri;
    const float t = rows_reduce<LPR, NR, false>(rs, l, ri);
    const int s = g + NS * ri;
    if ((l & (LPR / NR - 1)) == 0 && s < n) g_other[s] = t;
  }
  __syncthreads();
}

template <bool PTAB, class CFac>
__device__ __forceinline__ void hop_backward(int n, const float* __restrict__ tile, bool neg, const float* ea,
                                             bool subject_role, float* cbuf, float* g_other,
                                             float* __restrict__ gslice, BlockScratch& sc, CFac cfac) {
#define DFOL_HOP_BWD(LPR, NR)                                                                                        \
  {                                                                                                                  \
    if ((n & 3) == 0) {                                                                                              \
      if (!neg) hop_backward_t<LPR, NR, PTAB, false, true>(n, tile, ea, subject_role, cbuf, g_other, gslice, sc, cfac); \
      else hop_backward_t<LPR, NR, PTAB, true, true>(n, tile, ea, subject_role, cbuf, g_other, gslice, sc, cfac);    \
    } else {                                                                                                         \
      if (!neg) hop_backward_t<LPR, NR, PTAB, false, false>(n, tile, ea, subject_role, cbuf, g_other, gslice, sc, cfac); \
      else hop_backward_t<LPR, NR, PTAB, true, false>(n, tile, ea, subject_role, cbuf, g_other, gslice, sc, cfac);   \
    }                                                                                                                \
  }
  if (n <= 32) DFOL_HOP_BWD(8, (32 + 4 * PROG_WARPS - 1) / (4 * PROG_WARPS))
  else if (n <= 64) DFOL_HOP_BWD(16, 64 / (2 * PROG_WARPS))
  else DFOL_HOP_BWD(32, MAXN / PROG_WARPS)
#undef DFOL_HOP_BWD
}

// ---- option lists in probability space (choose_attr / query_attr, soft quantifier, no modulation) ----
//   lp_k = slog(1 - Q_k),  Q_k = prod_t (1 - term_kt),  term_kt = a_t e^{raw_kt} / den_t   (negated option: a_t - that)
// at_s[] = e^{cur}, wt_s[] = at / max(den, eps) (at itself when the list is not normalised), both ZERO beyond n.
// A warp takes 8 options per pass (rows of the next pass already in flight), every lane multiplies its objects' factors
// and ONE transposed butterfly reduces the 8 products.  BWD: the lane that ends up with option k turns d loss / d lp_k
// into c_k = d_lp Q / (1 - Q), the eight c are broadcast back and every lane emits the gradients of its (k, t) entries:
//   d loss / d cur_t += c_k term / (1 - term),   d loss / d nrm_kt = +- c_k x / (1 - term),  x = a_t e^{raw} / den_t
// (the softmax correction of the normalisation is the caller's second pass, as in the log-space path).
constexpr int OPT_STAGE = 2048;  // option words of one instruction staged in shared memory

__device__ __forceinline__ const int32_t* stage_options(const int32_t* __restrict__ op, int count, int32_t* stage) {
  if (count > OPT_STAGE) return op;
  __syncthreads();  // previous users of the staging area
  for (int i = threadIdx.x; i < count; i += PROG_THREADS) stage[i] = __ldg(op + i);
  __syncthreads();
  return stage;
}

// lane that holds row index u after rows_reduce<32, U, .> (U = 8 or 4)
template <int U>
__device__ __forceinline__ int reduce_src_lane(int u) {
  return U == 8 ? (((u & 4) ? 16 : 0) | ((u & 2) ? 8 : 0) | ((u & 1) ? 4 : 0)) : (((u & 2) ? 16 : 0) | ((u & 1) ? 8 : 0));
}

template <int U, int NC, bool BWD, bool PREFETCH, class LpSink>
__device__ __forceinline__ void options_pspace_nc(const Image& im, const int32_t* op, int count, const float* at_s,
                                                  const float* wt_s, const float* __restrict__ d_lp,
                                                  float* __restrict__ gslice, float (&acc_g)[NCHUNK],
                                                  float (&acc_tot)[NCHUNK], LpSink lp_sink) {
  static_assert(U == 8 || U == 4, "options per warp pass");
  constexpr float kLog2e = 1.4426950408889634f;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float at[NC], wt[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) { at[j] = at_s[lane + 32 * j]; wt[j] = wt_s[lane + 32 * j]; }
  float raw[U][NC];
  if constexpr (PREFETCH) {
    if (w < count) options_load<U, NC>(im, op, count, w, raw);
  }
  for (int k0 = w; k0 < count; k0 += PROG_WARPS * U) {
    float nxt[PREFETCH ? U : 1][NC];
    const int k1 = k0 + PROG_WARPS * U;
    if constexpr (PREFETCH) {
      if (k1 < count) options_load<U, NC>(im, op, count, k1, nxt);
    } else {
      options_load<U, NC>(im, op, count, k0, raw);
    }
    float v[U];
    unsigned negmask = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + u * PROG_WARPS;
      const bool ok = k < count;
      const bool neg = ok && (op[k] & DFOL_OPT_NEG) != 0;
      negmask |= neg ? (1u << u) : 0u;
      float prod = 1.0f;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const float e = ex2_approx(raw[u][j] * kLog2e);
        raw[u][j] = e;  // kept for the gradient pass
        const float x = wt[j] * e;
        prod *= 1.0f - (neg ? at[j] - x : x);
      }
      v[u] = ok ? prod : 1.0f;
    }
    int ri;
    const float Q = rows_reduce<32, U, true>(v, lane, ri);
    const int kq = k0 + ri * PROG_WARPS;
    const bool writer = (lane & (32 / U - 1)) == 0 && kq < count;
    if (!BWD) {
      if (writer) lp_sink(kq, Q);
    } else {
      float c = 0.0f;
      if (writer) {
        const float uu = 1.0f - Q;
        c = (uu >= kLogEps) ? d_lp[kq] * __fdividef(Q, uu) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float cu = __shfl_sync(0xffffffffu, c, reduce_src_lane<U>(u));
        const int k = k0 + u * PROG_WARPS;
        if (k < count) {  // warp-uniform
          const bool neg = (negmask >> u) & 1u;
#pragma unroll
          for (int j = 0; j < NC; ++j) {
            const int t = lane + 32 * j;
            const float x = wt[j] * raw[u][j];
            const float term = neg ? at[j] - x : x;
            const float cf = cu * rcp_approx(fmaxf(1.0f - term, kLogEps));
            const float dn = neg ? -(cf * x) : cf * x;
            acc_g[j] += cf * term;
            acc_tot[j] += dn;
            if (t < im.n) gslice[(long long)k * im.astride + t] = dn;
          }
        }
      }
    }
    if constexpr (PREFETCH) {
      if (k1 < count) {
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < NC; ++j) raw[u][j] = nxt[u][j];
      }
    }
  }
}
#endif  // DFOL_PROGRAM_FAST

}  // namespace dfol
