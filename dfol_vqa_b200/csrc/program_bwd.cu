// Backward program interpreter: one thread block per question walks its program in reverse, re-reading the table
// slices and the attention tape written by the forward kernel, and emits one compact gradient slice per
// (instruction, table operand) w.r.t. the RAW table entries (see include/dfol_b200.h).
// Closed forms: SURVEY.md Appendix B, verified against the reference's autograd through the CPU oracle.
#include "program_common.cuh"

namespace dfol {

// modulation row, compiled out of the unmodulated instantiation
#define DFOL_LOAD_MOD(row) (MOD ? load_mod(mods, (row)) : load_mod(nullptr, -1))

struct BwdShared {
  float cur[MAXN];    // attention before the current instruction (tape)
  float saved[MAXN];  // final attention of the first branch
  float g[MAXN];      // d loss / d cur, flowing backwards
  float gs[MAXN];     // d loss / d saved
  float nw[MAXN];
  float res[MAXN];
  float inner[MAXN];
  float den[MAXN];
  float tot[MAXN];
  float dres[MAXN];
  float tmp[MAXN];
  BlockScratch sc;
};

__device__ __forceinline__ void bwd_option_denominators(const Image& im, const int32_t* opts, int count, float* den,
                                                        BlockScratch& sc) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float acc[NCHUNK];
#pragma unroll
  for (int j = 0; j < NCHUNK; ++j) acc[j] = 0.f;
  for (int k = w; k < count; k += PROG_WARPS) {
    const int col = opts[k] & ~DFOL_OPT_NEG;
#pragma unroll
    for (int j = 0; j < NCHUNK; ++j) {
      const int t = lane + 32 * j;
      if (t < im.n) acc[j] += DFOL_EXPF(attr_raw(im, col, t));
    }
  }
  reduce_columns(acc, im.n, den, sc, false);
}

// normalised raw value of attribute option `word` at object t
__device__ __forceinline__ float option_nrm(const Image& im, int word, int t, bool normalise, const float* den) {
  float r = attr_raw(im, word & ~DFOL_OPT_NEG, t);
  if (normalise) r -= slog(den[t]);
  return r;
}

// Softmax correction of the option normalisation: d raw_j = d nrm_j - (sum_k d nrm_k) * exp(raw_j) / den.
// g slices currently hold d nrm_k; tot[t] = sum_k d nrm_k[t].
__device__ __forceinline__ void attr_softmax_correction(const Image& im, const int32_t* op, int count, float* gslice,
                                                        const float* den, const float* tot) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = w; k < count; k += PROG_WARPS) {
    const int col = op[k] & ~DFOL_OPT_NEG;
#pragma unroll
    for (int j = 0; j < NCHUNK; ++j) {
      const int t = lane + 32 * j;
      if (t < im.n) {
        const float d = den[t];
        const float inv = (d >= kLogEps) ? 1.0f / d : 0.0f;
        gslice[(long long)k * im.astride + t] -= tot[t] * DFOL_EXPF(attr_raw(im, col, t)) * inv;
      }
    }
  }
  __syncthreads();
}

// d loss / d modulation row: block-wide (every thread calls) or warp-wide sum of the per-thread partials; one writer.
__device__ __forceinline__ void block_write_dm(const float dm[4], float* __restrict__ d_mods, int row, BlockScratch& sc) {
  const float r0 = block_sum(dm[0], sc), r1 = block_sum(dm[1], sc), r2 = block_sum(dm[2], sc), r3 = block_sum(dm[3], sc);
  if (threadIdx.x == 0 && d_mods != nullptr) reinterpret_cast<float4*>(d_mods)[row] = make_float4(r0, r1, r2, r3);
}
__device__ __forceinline__ void warp_write_dm(const float dm[4], float* __restrict__ d_mods, int row) {
  const float r0 = warp_sum(dm[0]), r1 = warp_sum(dm[1]), r2 = warp_sum(dm[2]), r3 = warp_sum(dm[3]);
  if ((threadIdx.x & 31) == 0 && d_mods != nullptr) reinterpret_cast<float4*>(d_mods)[row] = make_float4(r0, r1, r2, r3);
}

// Backward of relate_forward for the kept role. dres = d loss / d res. Adds the gradient of the OTHER role's
// prior into g_other (accumulating) and writes d loss / d nrm[s,o] into gslice (n x n tile, diagonal zero).
template <class LL>
__device__ __forceinline__ void relate_backward(int n, const LL& L, const float* a_subj, const float* a_obj,
                                                bool subject_role, const float* dres, const float* inner,
                                                float* g_other, float* gslice, BlockScratch& sc) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool neg = L.neg(), rt = L.roundtrip;
  float acc[NCHUNK];
#pragma unroll
  for (int j = 0; j < NCHUNK; ++j) acc[j] = 0.f;
  for (int s = w; s < n; s += PROG_WARPS) {
    const float as = a_subj[s];
    const float dS_row = subject_role ? dres[s] * lnot_grad(inner[s]) : 0.0f;
    float rowsum = 0.f;
#pragma unroll
    for (int j = 0; j < NCHUNK; ++j) {
      const int o = lane + 32 * j;
      if (o < n) {
        float dn = 0.0f;
        if (o != s) {
          const float nrm = L.raw_nrm(s, o);
          const float l = post_ll(nrm, neg, rt);
          float du;
          if (subject_role) {
            du = dS_row * lnot_grad(l + a_obj[o]);
            acc[j] += du;
          } else {
            du = dres[o] * lnot_grad(inner[o]) * lnot_grad(l + as);
            rowsum += du;
          }
          dn = du * post_ll_grad(nrm, neg, rt);
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

template <bool MOD>
static __global__ void __launch_bounds__(PROG_THREADS) program_bwd_kernel(
    const int32_t* __restrict__ instr, const int32_t* __restrict__ q_instr, const int32_t* __restrict__ opts,
    const float* __restrict__ attr_ll, const int64_t* __restrict__ attr_blk, const int32_t* __restrict__ attr_stride,
    const float* __restrict__ rel_ll, const int64_t* __restrict__ rel_blk, const int32_t* __restrict__ rel_stride,
    const int32_t* __restrict__ img_n, const float* __restrict__ mods, const float* __restrict__ d_lp,
    const float* __restrict__ tape, int tape_stride, float* __restrict__ g_attr, float* __restrict__ g_rel,
    float* __restrict__ d_mods
#ifdef DFOL_PROGRAM_FAST
    , int ring_nbuf, int ring_tile_floats
#endif
    ) {
  __shared__ __align__(16) BwdShared sm;
  const int q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  Image im;
  im.n = img_n[q];
  im.attr = attr_ll + attr_blk[q];
  im.astride = attr_stride[q];
  im.rel = rel_ll + rel_blk[q];
  im.rstride = rel_stride[q];
  const int n = im.n;
  const int ip0 = q_instr[q], ip1 = q_instr[q + 1];

  if (tid < MAXN) { sm.g[tid] = 0.f; sm.gs[tid] = 0.f; sm.saved[tid] = 0.f; sm.cur[tid] = 0.f; }
#ifdef DFOL_PROGRAM_FAST
  // relate tiles of this program, streamed through the shared-memory ring in REVERSE execution order
  extern __shared__ __align__(128) float ring_mem[];
  __shared__ __align__(8) uint64_t ring_full[8];
  __shared__ int rel_ip[MAX_REL];
  __shared__ int rel_count;
  TileRing ring{ring_mem, ring_nbuf, ring_tile_floats, ring_full};
  auto issue_tile = [&](int j) {  // elected thread: tile of the j-th relate counted from the END -> slot j % nbuf
    const int col = instr[(long long)rel_ip[rel_count - 1 - j] * DFOL_INSTR_WORDS + DFOL_I_A0];
    const int b = j % ring.nbuf;
    const uint32_t bytes = (uint32_t)im.rstride * 4u;
    mbar_expect_tx(&ring.full[b], bytes);
    bulk_load(ring.buf + (size_t)b * ring.tile_floats, im.rel + (long long)col * im.rstride, bytes, &ring.full[b]);
  };
  if (tid == 0) {
    int c = 0;
    for (int ip = ip0; ip < ip1 && c < MAX_REL; ++ip)
      if (instr[(long long)ip * DFOL_INSTR_WORDS + DFOL_I_OP] == DFOL_OP_RELATE) rel_ip[c++] = ip;
    rel_count = c;
    for (int b = 0; b < ring.nbuf; ++b) mbar_init(&ring.full[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int j = 0; j < c && j < ring.nbuf; ++j) issue_tile(j);
  }
  __syncthreads();
  // relates beyond the ring capacity (the LAST ones in execution order, met first here) use direct loads
  int total_rel = 0;
  for (int ip = ip0; ip < ip1; ++ip)
    total_rel += instr[(long long)ip * DFOL_INSTR_WORDS + DFOL_I_OP] == DFOL_OP_RELATE;
  int rel_seen = 0;  // relates met so far walking backwards
#endif
  // the first branch's final attention is the tape row of the PUSH instruction
  for (int ip = ip0; ip < ip1; ++ip)
    if (instr[(long long)ip * DFOL_INSTR_WORDS + DFOL_I_OP] == DFOL_OP_PUSH && tid < n)
      sm.saved[tid] = tape[(long long)ip * tape_stride + tid];
  __syncthreads();

  for (int ip = ip1 - 1; ip >= ip0; --ip) {
    const Instr I = load_instr(instr, ip);
    const bool neg = I.flags & DFOL_F_NEG, rt = I.flags & DFOL_F_ROUNDTRIP;
    const bool normalise = I.flags & DFOL_F_NORMALISE;
    __syncthreads();
    if (tid < n) sm.cur[tid] = tape[(long long)ip * tape_stride + tid];
    __syncthreads();

    switch (I.op) {
      case DFOL_OP_SELECT:
      case DFOL_OP_FILTER: {
        // x = (filter: cur +) post(ll); out = M(x)
        const Mod m = DFOL_LOAD_MOD(I.mod);
        float dm[4] = {0.f, 0.f, 0.f, 0.f};
        if (tid < n) {
          const float raw = (I.a0 >= 0) ? attr_raw(im, I.a0, tid) : 0.0f;
          const float l = (I.a0 >= 0) ? post_ll(raw, neg, rt) : 0.0f;
          const float x = (I.op == DFOL_OP_FILTER) ? sm.cur[tid] + l : l;
          const float dx = mod_grad(m, x, sm.g[tid], dm);
          if (I.a0 >= 0) g_attr[I.ga0 + tid] = dx * post_ll_grad(raw, neg, rt);
          sm.g[tid] = (I.op == DFOL_OP_FILTER) ? dx : 0.f;
        }
        if (m.on) block_write_dm(dm, d_mods, I.mod, sm.sc);
        break;
      }

      case DFOL_OP_PUSH:
        if (tid < n) { sm.g[tid] = sm.gs[tid]; sm.gs[tid] = 0.f; }
        break;

      case DFOL_OP_RELATE: {
        const bool nneg = I.flags & DFOL_F_NAME_NEG, nrt = I.flags & DFOL_F_NAME_ROUNDTRIP;
        const Mod mr = DFOL_LOAD_MOD(I.mod), ms = DFOL_LOAD_MOD(I.mod2);
        float nw0 = 0.0f;  // prior of the new object before its modulation
        if (tid < n) {
          nw0 = (I.a1 >= 0) ? post_ll(attr_raw(im, I.a1, tid), nneg, nrt) : 0.0f;
          sm.nw[tid] = mod_apply(ms, nw0);
          sm.dres[tid] = sm.g[tid];
          sm.tmp[tid] = 0.f;
        }
        __syncthreads();
        const bool subj = I.flags & DFOL_F_SUBJECT;
        const float* a_s = subj ? sm.nw : sm.cur;
        const float* a_o = subj ? sm.cur : sm.nw;
#ifdef DFOL_PROGRAM_FAST
        const int j = rel_seen - (total_rel - rel_count);  // position among the ring-served relates, from the end
        ++rel_seen;
        if (j >= 0) {
          const int b = j % ring.nbuf;
          mbar_wait(&ring.full[b], (uint32_t)(j / ring.nbuf) & 1u);
          const float* tile = ring.buf + (size_t)b * ring.tile_floats;
          relate_forward_tile(n, tile, neg, rt, a_s, a_o, subj, sm.res, sm.inner, sm.den, sm.sc);
          if (mr.on) {  // gradient through the modulation of the posterior: sm.g (d out) -> sm.dres (d res)
            float dm[4] = {0.f, 0.f, 0.f, 0.f};
            if (tid < n) sm.dres[tid] = mod_grad(mr, sm.res[tid], sm.g[tid], dm);
            block_write_dm(dm, d_mods, I.mod, sm.sc);
          }
          relate_backward_tile(n, tile, neg, rt, a_s, subj, sm.dres, sm.inner, sm.den, sm.tot, sm.tmp, g_rel + I.gr,
                               sm.sc);
          if (tid == 0 && j + ring.nbuf < rel_count) issue_tile(j + ring.nbuf);
        } else
#endif
        {
          RelOption L{&im, nullptr, 1, 0, false, rt, I.a0, neg};
          relate_forward(n, L, a_s, a_o, subj, sm.res, sm.inner, sm.sc);
          if (mr.on) {
            float dm[4] = {0.f, 0.f, 0.f, 0.f};
            if (tid < n) sm.dres[tid] = mod_grad(mr, sm.res[tid], sm.g[tid], dm);
            block_write_dm(dm, d_mods, I.mod, sm.sc);
          }
          relate_backward(n, L, a_s, a_o, subj, sm.dres, sm.inner, sm.tmp, g_rel + I.gr, sm.sc);
        }
        {
          float dm[4] = {0.f, 0.f, 0.f, 0.f};
          if (tid < n) {
            // the kept role's own prior is the (modulated) new object: its gradient feeds the name select
            const float dnw0 = mod_grad(ms, nw0, sm.dres[tid], dm);
            if (I.a1 >= 0) g_attr[I.ga1 + tid] = dnw0 * post_ll_grad(attr_raw(im, I.a1, tid), nneg, nrt);
            sm.g[tid] = sm.tmp[tid];
          }
          if (ms.on) block_write_dm(dm, d_mods, I.mod2, sm.sc);
        }
        break;
      }

      case DFOL_OP_EXIST: {
        float S;
        exists_block(sm.cur, n, false, sm.sc, &S);
        const float d = d_lp[I.out] * lnot_grad(S);
        if (tid < n) sm.g[tid] = d * lnot_grad(sm.cur[tid]);
        break;
      }

      case DFOL_OP_AND:
      case DFOL_OP_OR: {
        float S1, S2;
        const float e1 = exists_block(sm.saved, n, false, sm.sc, &S1);
        const float e2 = exists_block(sm.cur, n, false, sm.sc, &S2);
        const float dlp = d_lp[I.out];
        float de1 = dlp, de2 = dlp;
        if (I.op == DFOL_OP_OR) {
          const float E1 = DFOL_EXPF(e1), E2 = DFOL_EXPF(e2);
          const float v = 1.0f - (1.0f - E1) * (1.0f - E2);
          const float dv = (v >= kLogEps) ? dlp / v : 0.0f;
          de1 = dv * (1.0f - E2) * E1;
          de2 = dv * (1.0f - E1) * E2;
        }
        if (tid < n) {
          sm.gs[tid] = de1 * lnot_grad(S1) * lnot_grad(sm.saved[tid]);
          sm.g[tid] = de2 * lnot_grad(S2) * lnot_grad(sm.cur[tid]);
        }
        break;
      }

      case DFOL_OP_VERIFY_ATTRS: {
        const int32_t* op = opts + I.a0;
        float acc[NCHUNK];
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) acc[j] = 0.f;
        const bool modulated = MOD && I.mod >= 0;
        for (int k = w; k < I.a1; k += PROG_WARPS) {
          const Mod mk = DFOL_LOAD_MOD(modulated ? I.mod + k : -1);
#pragma unroll
          for (int j = 0; j < NCHUNK; ++j) {
            const int t = lane + 32 * j;
            if (t < n)
              acc[j] += mod_apply(mk, sm.cur[t] + post_ll(attr_raw(im, op[k] & ~DFOL_OPT_NEG, t), op[k] & DFOL_OPT_NEG, rt));
          }
        }
        reduce_columns(acc, n, sm.res, sm.sc, false);
        float S;
        exists_block(sm.res, n, false, sm.sc, &S);
        const float d = d_lp[I.out] * lnot_grad(S);
        if (tid < n) {
          const float datt = d * lnot_grad(sm.res[tid]);
          sm.g[tid] = (float)I.a1 * datt;
          sm.tmp[tid] = datt;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) acc[j] = 0.f;
        for (int k = w; k < I.a1; k += PROG_WARPS) {
          const Mod mk = DFOL_LOAD_MOD(modulated ? I.mod + k : -1);
          float dm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < NCHUNK; ++j) {
            const int t = lane + 32 * j;
            if (t < n) {
              const float raw = attr_raw(im, op[k] & ~DFOL_OPT_NEG, t);
              const float dx = mod_grad(mk, sm.cur[t] + post_ll(raw, op[k] & DFOL_OPT_NEG, rt), sm.tmp[t], dm);
              acc[j] += dx;
              g_attr[I.ga0 + (long long)k * im.astride + t] = dx * post_ll_grad(raw, op[k] & DFOL_OPT_NEG, rt);
            }
          }
          if (mk.on) warp_write_dm(dm, d_mods, I.mod + k);
        }
        if (modulated) reduce_columns(acc, n, sm.g, sm.sc, false);  // d cur = sum_k dx_k (unmodulated: a1 * datt, above)
        break;
      }

      case DFOL_OP_CHOOSE_ATTR:
      case DFOL_OP_ALL_SAME:
      case DFOL_OP_TWO_SAME: {
        const int32_t* op = opts + I.a0;
        float* gslice = g_attr + I.ga0;
        if (normalise) bwd_option_denominators(im, op, I.a1, sm.den, sm.sc);
        float dQ = 0.f;
        if (I.op != DFOL_OP_CHOOSE_ATTR) {
          // first pass: Q = sum_k lnot(v_k) exactly as in the forward kernel
          float part = 0.f;
          for (int k = w; k < I.a1; k += PROG_WARPS) {
            const bool kneg = op[k] & DFOL_OPT_NEG;
            const Mod m1 = DFOL_LOAD_MOD(I.mod >= 0 ? I.mod + k : -1), m2 = DFOL_LOAD_MOD(I.mod2 >= 0 ? I.mod2 + k : -1);
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < NCHUNK; ++j) {
              const int t = lane + 32 * j;
              if (t < n) {
                const float l = post_ll(option_nrm(im, op[k], t, normalise, sm.den), kneg, rt);
                if (I.op == DFOL_OP_ALL_SAME) {
                  const float a = sm.cur[t];
                  s1 += roundtrip(lnot(a + lnot(mod_apply(m1, a + l))));
                } else {
                  s1 += lnot(mod_apply(m1, sm.saved[t] + l));
                  s2 += lnot(mod_apply(m2, sm.cur[t] + l));
                }
              }
            }
            s1 = warp_sum(s1);
            if (I.op == DFOL_OP_ALL_SAME) part += lnot(roundtrip(s1));
            else { s2 = warp_sum(s2); part += lnot(lnot(s1) + lnot(s2)); }
          }
          const float Q = block_sum(lane == 0 ? part : 0.f, sm.sc);
          float dlp = d_lp[I.out];
          if (I.flags & DFOL_F_NEGATE_RESULT) dlp *= lnot_grad(lnot(Q));
          dQ = dlp * lnot_grad(Q);
        }
        float acc_g[NCHUNK], acc_gs[NCHUNK], acc_tot[NCHUNK];
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) { acc_g[j] = 0.f; acc_gs[j] = 0.f; acc_tot[j] = 0.f; }
        for (int k = w; k < I.a1; k += PROG_WARPS) {
          const bool kneg = op[k] & DFOL_OPT_NEG;
          const Mod m1 = DFOL_LOAD_MOD(I.mod >= 0 ? I.mod + k : -1), m2 = DFOL_LOAD_MOD(I.mod2 >= 0 ? I.mod2 + k : -1);
          float dm1[4] = {0.f, 0.f, 0.f, 0.f}, dm2[4] = {0.f, 0.f, 0.f, 0.f};
          float nrm[NCHUNK], l[NCHUNK], xm1[NCHUNK], xm2[NCHUNK];  // xm: modulated filter outputs
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int j = 0; j < NCHUNK; ++j) {
            const int t = lane + 32 * j;
            nrm[j] = 0.f; l[j] = 0.f; xm1[j] = 0.f; xm2[j] = 0.f;
            if (t < n) {
              nrm[j] = option_nrm(im, op[k], t, normalise, sm.den);
              l[j] = post_ll(nrm[j], kneg, rt);
              if (I.op == DFOL_OP_CHOOSE_ATTR) { xm1[j] = mod_apply(m1, sm.cur[t] + l[j]); s1 += lnot(xm1[j]); }
              else if (I.op == DFOL_OP_ALL_SAME) {
                const float a = sm.cur[t];
                xm1[j] = mod_apply(m1, a + l[j]);
                s1 += roundtrip(lnot(a + lnot(xm1[j])));
              } else {
                xm1[j] = mod_apply(m1, sm.saved[t] + l[j]);
                xm2[j] = mod_apply(m2, sm.cur[t] + l[j]);
                s1 += lnot(xm1[j]);
                s2 += lnot(xm2[j]);
              }
            }
          }
          s1 = warp_sum(s1);
          s2 = warp_sum(s2);
          float d1 = 0.f, d2 = 0.f;  // d loss / d s1, d s2
          if (I.op == DFOL_OP_CHOOSE_ATTR) d1 = d_lp[I.out + k] * lnot_grad(s1);
          else if (I.op == DFOL_OP_ALL_SAME) d1 = dQ * lnot_grad(roundtrip(s1)) * roundtrip_grad(s1);
          else {
            const float dv = dQ * lnot_grad(lnot(s1) + lnot(s2));
            d1 = dv * lnot_grad(s1);
            d2 = dv * lnot_grad(s2);
          }
#pragma unroll
          for (int j = 0; j < NCHUNK; ++j) {
            const int t = lane + 32 * j;
            if (t < n) {
              float dl;
              if (I.op == DFOL_OP_CHOOSE_ATTR) {
                const float dx = mod_grad(m1, sm.cur[t] + l[j], d1 * lnot_grad(xm1[j]), dm1);
                acc_g[j] += dx;
                dl = dx;
              } else if (I.op == DFOL_OP_ALL_SAME) {
                const float a = sm.cur[t];
                const float wv = a + lnot(xm1[j]);
                const float y = lnot(wv);
                const float dw = d1 * roundtrip_grad(y) * lnot_grad(wv);
                const float dx = mod_grad(m1, a + l[j], dw * lnot_grad(xm1[j]), dm1);
                acc_g[j] += dw + dx;
                dl = dx;
              } else {
                const float dx1 = mod_grad(m1, sm.saved[t] + l[j], d1 * lnot_grad(xm1[j]), dm1);
                const float dx2 = mod_grad(m2, sm.cur[t] + l[j], d2 * lnot_grad(xm2[j]), dm2);
                acc_gs[j] += dx1;
                acc_g[j] += dx2;
                dl = dx1 + dx2;
              }
              const float dn = dl * post_ll_grad(nrm[j], kneg, rt);
              acc_tot[j] += dn;
              gslice[(long long)k * im.astride + t] = dn;
            }
          }
          if (m1.on) warp_write_dm(dm1, d_mods, I.mod + k);
          if (m2.on) warp_write_dm(dm2, d_mods, I.mod2 + k);
        }
        reduce_columns(acc_g, n, sm.g, sm.sc, false);
        if (I.op == DFOL_OP_TWO_SAME) reduce_columns(acc_gs, n, sm.gs, sm.sc, false);
        if (normalise) {
          reduce_columns(acc_tot, n, sm.tot, sm.sc, false);
          attr_softmax_correction(im, op, I.a1, gslice, sm.den, sm.tot);
        }
        break;
      }

      case DFOL_OP_COMPARE: {
        const Mod m1 = DFOL_LOAD_MOD(I.mod), m2 = DFOL_LOAD_MOD(I.mod2);
        if (tid < n) {
          const float l = (I.a0 >= 0) ? post_ll(attr_raw(im, I.a0, tid), neg, rt) : 0.0f;
          sm.res[tid] = mod_apply(m1, sm.saved[tid] + l);
          sm.nw[tid] = mod_apply(m2, sm.cur[tid] + l);
        }
        __syncthreads();
        float S1, S2;
        const float e1 = exists_block(sm.res, n, false, sm.sc, &S1);
        const float e2 = exists_block(sm.nw, n, false, sm.sc, &S2);
        const float mx = fmaxf(e1, e2);
        const float lse = DFOL_LOGF(DFOL_EXPF(e1 - mx) + DFOL_EXPF(e2 - mx));
        const float z1 = e1 - mx - lse, z2 = e2 - mx - lse;
        const float alpha = (I.flags & DFOL_F_IS_LESS) ? 1.0f : 0.0f;
        const float c = 1.0f - 2.0f * alpha;
        const float v1 = alpha + c * DFOL_EXPF(z1), v2 = alpha + c * DFOL_EXPF(z2);
        const float dz1 = (v1 >= kLogEps) ? d_lp[I.out] * c * DFOL_EXPF(z1) / v1 : 0.0f;
        const float dz2 = (v2 >= kLogEps) ? d_lp[I.out + 1] * c * DFOL_EXPF(z2) / v2 : 0.0f;
        const float de1 = dz1 - DFOL_EXPF(z1) * (dz1 + dz2);
        const float de2 = dz2 - DFOL_EXPF(z2) * (dz1 + dz2);
        float dm1[4] = {0.f, 0.f, 0.f, 0.f}, dm2[4] = {0.f, 0.f, 0.f, 0.f};
        if (tid < n) {
          const float raw = (I.a0 >= 0) ? attr_raw(im, I.a0, tid) : 0.0f;
          const float l = (I.a0 >= 0) ? post_ll(raw, neg, rt) : 0.0f;
          const float dx1 = mod_grad(m1, sm.saved[tid] + l, de1 * lnot_grad(S1) * lnot_grad(sm.res[tid]), dm1);
          const float dx2 = mod_grad(m2, sm.cur[tid] + l, de2 * lnot_grad(S2) * lnot_grad(sm.nw[tid]), dm2);
          sm.gs[tid] = dx1;
          sm.g[tid] = dx2;
          if (I.a0 >= 0) g_attr[I.ga0 + tid] = (dx1 + dx2) * post_ll_grad(raw, neg, rt);
        }
        if (m1.on) block_write_dm(dm1, d_mods, I.mod, sm.sc);
        if (m2.on) block_write_dm(dm2, d_mods, I.mod2, sm.sc);
        break;
      }

      case DFOL_OP_CHOOSE_REL: {
        const bool nneg = I.flags & DFOL_F_NAME_NEG, nrt = I.flags & DFOL_F_NAME_ROUNDTRIP;
        const Mod ms = DFOL_LOAD_MOD(I.mod2);
        float nw0 = 0.0f;
        if (tid < n) {
          nw0 = (I.a2 >= 0) ? post_ll(attr_raw(im, I.a2, tid), nneg, nrt) : 0.0f;
          sm.nw[tid] = mod_apply(ms, nw0);
          sm.tmp[tid] = 0.f;  // gradient of the incoming attention
          sm.tot[tid] = 0.f;  // gradient of the new object's (modulated) prior
        }
        __syncthreads();
        const bool subj = I.flags & DFOL_F_SUBJECT;
        const float* a_s = subj ? sm.nw : sm.cur;
        const float* a_o = subj ? sm.cur : sm.nw;
        for (int k = 0; k < I.a1; ++k) {
          RelOption L{&im, opts + I.a0, I.a1, k, normalise, rt, -1, false};
          relate_forward(n, L, a_s, a_o, subj, sm.res, sm.inner, sm.sc);
          const Mod mk = DFOL_LOAD_MOD(I.mod >= 0 ? I.mod + k : -1);
          if (tid < n) sm.den[tid] = mod_apply(mk, sm.res[tid]);  // modulated posterior of option k
          __syncthreads();
          float S;
          exists_block(sm.den, n, false, sm.sc, &S);
          const float d = d_lp[I.out + k] * lnot_grad(S);
          float dm[4] = {0.f, 0.f, 0.f, 0.f};
          if (tid < n) {
            sm.dres[tid] = mod_grad(mk, sm.res[tid], d * lnot_grad(sm.den[tid]), dm);
            sm.tot[tid] += sm.dres[tid];
          }
          if (mk.on) block_write_dm(dm, d_mods, I.mod + k, sm.sc);
          __syncthreads();
          relate_backward(n, L, a_s, a_o, subj, sm.dres, sm.inner, sm.tmp, g_rel + I.gr + (long long)k * im.rstride,
                          sm.sc);
        }
        if (normalise) {
          // softmax correction per pair: d raw_j = d nrm_j - (sum_k d nrm_k) exp(raw_j) / den
          const int32_t* op = opts + I.a0;
          float* gs = g_rel + I.gr;
          for (int e = tid; e < n * n; e += PROG_THREADS) {
            const int s = e / n, o = e - s * n;
            if (s == o) continue;
            float den = 0.f, tot = 0.f;
            for (int j = 0; j < I.a1; ++j) {
              den += DFOL_EXPF(rel_raw(im, op[j] & ~DFOL_OPT_NEG, s, o));
              tot += gs[(long long)j * im.rstride + e];
            }
            const float inv = (den >= kLogEps) ? 1.0f / den : 0.0f;
            for (int j = 0; j < I.a1; ++j)
              gs[(long long)j * im.rstride + e] -= tot * DFOL_EXPF(rel_raw(im, op[j] & ~DFOL_OPT_NEG, s, o)) * inv;
          }
        }
        __syncthreads();
        {
          float dm[4] = {0.f, 0.f, 0.f, 0.f};
          if (tid < n) {
            const float dnw0 = mod_grad(ms, nw0, sm.tot[tid], dm);
            if (I.a2 >= 0) g_attr[I.ga1 + tid] = dnw0 * post_ll_grad(attr_raw(im, I.a2, tid), nneg, nrt);
            sm.g[tid] = sm.tmp[tid];
          }
          if (ms.on) block_write_dm(dm, d_mods, I.mod2, sm.sc);
        }
        break;
      }

      default:
        break;
    }
  }
}

}  // namespace dfol

using namespace dfol;

#ifdef DFOL_PROGRAM_FAST
#define DFOL_PROGRAM_BWD_ENTRY dfol_program_bwd_fast
#else
#define DFOL_PROGRAM_BWD_ENTRY dfol_program_bwd
#endif

extern "C" int DFOL_PROGRAM_BWD_ENTRY(const int32_t* instr, const int32_t* q_instr, const int32_t* opts,
                                      int question_num, const float* attr_ll, const int64_t* attr_blk,
                                      const int32_t* attr_stride, const float* rel_ll, const int64_t* rel_blk,
                                      const int32_t* rel_stride, const int32_t* img_n, const float* mods,
                                      const float* d_lp, const float* tape, int tape_stride, float* g_attr,
                                      float* g_rel, float* d_mods, void* stream) {
  DFOL_REQUIRE(instr && q_instr && attr_ll && attr_blk && attr_stride && rel_ll && rel_blk && rel_stride && img_n &&
                   d_lp && tape && g_attr && g_rel,
               "dfol_program_bwd: null pointer");
  DFOL_REQUIRE((mods == nullptr) == (d_mods == nullptr), "dfol_program_bwd: mods and d_mods go together");
  if (question_num == 0) return 0;
#ifdef DFOL_PROGRAM_FAST
  DFOL_REQUIRE(tape_stride >= 1 && tape_stride <= MAXN, "dfol_program_bwd_fast: tape_stride = max objects rounded to 4");
  const int tile_floats = (tape_stride * tape_stride + 31) / 32 * 32;
  int nbuf = (96 * 1024) / (tile_floats * 4);
  nbuf = nbuf < 1 ? 1 : (nbuf > 4 ? 4 : nbuf);
  const size_t smem = (size_t)nbuf * tile_floats * 4;
  cudaFuncSetAttribute(mods ? program_bwd_kernel<true> : program_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  (mods ? program_bwd_kernel<true> : program_bwd_kernel<false>)<<<question_num, PROG_THREADS, smem, (cudaStream_t)stream>>>(
      instr, q_instr, opts, attr_ll, attr_blk, attr_stride, rel_ll, rel_blk, rel_stride, img_n, mods, d_lp, tape,
      tape_stride, g_attr, g_rel, d_mods, nbuf, tile_floats);
  return finish_launch("dfol_program_bwd_fast");
#else
  (mods ? program_bwd_kernel<true> : program_bwd_kernel<false>)<<<question_num, PROG_THREADS, 0, (cudaStream_t)stream>>>(
      instr, q_instr, opts, attr_ll, attr_blk, attr_stride, rel_ll, rel_blk, rel_stride, img_n, mods, d_lp, tape,
      tape_stride, g_attr, g_rel, d_mods);
  return finish_launch("dfol_program_bwd");
#endif
}
