// Device building blocks shared by the forward and backward program interpreter kernels.
#pragma once

#include "dfol_common.cuh"

namespace dfol {

constexpr int PROG_THREADS = 256;
constexpr int PROG_WARPS = PROG_THREADS / 32;
constexpr int MAXN = 128;  // objects per image supported by the interpreter kernels (GQA: <= 100)
constexpr int NCHUNK = MAXN / 32;

struct Instr {
  int op, flags, a0, a1, a2, out, ga0, ga1, gr;
};

__device__ __forceinline__ Instr load_instr(const int32_t* __restrict__ instr, int ip) {
  const int32_t* w = instr + (long long)ip * DFOL_INSTR_WORDS;
  Instr I;
  I.op = w[DFOL_I_OP]; I.flags = w[DFOL_I_FLAGS]; I.a0 = w[DFOL_I_A0]; I.a1 = w[DFOL_I_A1]; I.a2 = w[DFOL_I_A2];
  I.out = w[DFOL_I_OUT]; I.ga0 = w[DFOL_I_GA0]; I.ga1 = w[DFOL_I_GA1]; I.gr = w[DFOL_I_GR];
  return I;
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
  float colacc[PROG_WARPS][MAXN];
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
    for (int j = 0; j < count; ++j) den += expf(rel_raw(*im, opts[j] & ~DFOL_OPT_NEG, s, o));
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

}  // namespace dfol
