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
  float acc[NCHUNK];
#pragma unroll
  for (int j = 0; j < NCHUNK; ++j) acc[j] = 0.f;
  const int lane = threadIdx.x & 31;
  for_options<8>(im, opts, count, [&](int, int, const auto& raw) {
#pragma unroll
    for (int j = 0; j < DFOL_NC_OF(raw); ++j)
      if (lane + 32 * j < im.n) acc[j] += DFOL_EXPF(raw[j]);
  });
  reduce_columns(acc, im.n, den, sc, false);
}

// normalised raw value of attribute option `word` at object t
__device__ __forceinline__ float option_nrm(const Image& im, int word, int t, bool normalise, const float* den) {
  float r = attr_raw(im, word & ~DFOL_OPT_NEG, t);
  if (normalise) r -= slog(den[t]);
  return r;
}

__device__ __forceinline__ float option_nrm_of(float r, int t, bool normalise, const float* den) {
  if (normalise) r -= slog(den[t]);
  return r;
}

// Softmax correction of the option normalisation: d raw_j = d nrm_j - (sum_k d nrm_k) * exp(raw_j) / den.
// g slices currently hold d nrm_k; tot[t] = sum_k d nrm_k[t].
__device__ __forceinline__ void attr_softmax_correction(const Image& im, const int32_t* op, int count, float* gslice,
                                                        const float* den, const float* tot) {
  const int lane = threadIdx.x & 31;
  for_options<8>(im, op, count, [&](int k, int, const auto& raw) {
#pragma unroll
    for (int j = 0; j < DFOL_NC_OF(raw); ++j) {
      const int t = lane + 32 * j;
      if (t < im.n) {
        const float d = den[t];
        const float inv = (d >= kLogEps) ? 1.0f / d : 0.0f;
        gslice[(long long)k * im.astride + t] -= tot[t] * DFOL_EXPF(raw[j]) * inv;
      }
    }
  });
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

template <bool MOD, bool PTAB>
static __global__ void __launch_bounds__(PROG_THREADS, MOD ? 1 : PROG_MIN_BLOCKS) program_bwd_kernel(
    const int32_t* __restrict__ instr, const int32_t* __restrict__ q_instr, const int32_t* __restrict__ opts,
    const float* __restrict__ attr_ll, const int64_t* __restrict__ attr_blk, const int32_t* __restrict__ attr_stride,
    const float* __restrict__ rel_ll, const int64_t* __restrict__ rel_blk, const int32_t* __restrict__ rel_stride,
    const int32_t* __restrict__ img_n, const float* __restrict__ mods, const float* __restrict__ d_lp,
    const float* __restrict__ tape, int tape_stride, float* __restrict__ g_attr, float* __restrict__ g_rel,
    float* __restrict__ d_mods
#ifdef DFOL_PROGRAM_FAST
    , const float* __restrict__ rel_p, int ring_nbuf, int ring_tile_floats
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

  if (tid < MAXN) {
    sm.g[tid] = 0.f; sm.gs[tid] = 0.f; sm.saved[tid] = 0.f; sm.cur[tid] = 0.f; sm.den[tid] = 0.f; sm.tot[tid] = 0.f;
  }
#ifdef DFOL_PROGRAM_FAST
  // relate tiles of this program, streamed through the shared-memory ring in REVERSE execution order.  The bytecode is
  // staged in shared memory first (one parallel read instead of a chain of dependent global loads per instruction).
  extern __shared__ __align__(128) float ring_mem[];
  __shared__ __align__(8) uint64_t ring_full[8];
  __shared__ int rel_ip[MAX_REL];
  __shared__ int rel_count, rel_beyond, push_ip;
  __shared__ int32_t code_s[MAX_CODE * DFOL_INSTR_WORDS];
  __shared__ int32_t opt_s[OPT_STAGE];  // option words of the instruction being executed
  TileRing ring{ring_mem, ring_nbuf, ring_tile_floats, ring_full};
  const float* ring_src = PTAB ? rel_p + rel_blk[q] : im.rel;
  const int code_n = min(ip1 - ip0, MAX_CODE);
  for (int i = tid; i < code_n * DFOL_INSTR_WORDS; i += PROG_THREADS)
    code_s[i] = instr[(long long)ip0 * DFOL_INSTR_WORDS + i];
  if (tid == 0) {
    for (int b = 0; b < ring.nbuf; ++b) mbar_init(&ring.full[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    push_ip = -1;
  }
  __syncthreads();
  auto issue_tile = [&](int j, int b) {  // elected thread: tile of the j-th ring-served relate counted from the END
    const int col = code_s[(rel_ip[rel_count - 1 - j] - ip0) * DFOL_INSTR_WORDS + DFOL_I_A0];
    const uint32_t bytes = (uint32_t)im.rstride * 4u;
    mbar_expect_tx(&ring.full[b], bytes);
    bulk_load(ring.buf + (size_t)b * ring.tile_floats, ring_src + (long long)col * im.rstride, bytes, &ring.full[b]);
  };
  if (w == 0) {  // warp 0 compacts the relate hops of the staged instructions (ballot + prefix popcount)
    int base = 0;
    for (int c0 = 0; c0 < code_n; c0 += 32) {
      const int idx = c0 + lane;
      const int op = idx < code_n ? code_s[idx * DFOL_INSTR_WORDS + DFOL_I_OP] : 0;
      const unsigned m = __ballot_sync(0xffffffffu, op == DFOL_OP_RELATE);
      if (op == DFOL_OP_RELATE) rel_ip[base + __popc(m & ((1u << lane) - 1u))] = ip0 + idx;
      if (op == DFOL_OP_PUSH) push_ip = ip0 + idx;
      base += __popc(m);
    }
    __syncwarp();
    if (lane == 0) {
      int beyond = 0;  // programs longer than the staged window: their last relates use direct loads
      for (int ip = ip0 + code_n; ip < ip1; ++ip) {
        const int op = instr[(long long)ip * DFOL_INSTR_WORDS + DFOL_I_OP];
        beyond += op == DFOL_OP_RELATE;
        if (op == DFOL_OP_PUSH) push_ip = ip;
      }
      rel_count = base;
      rel_beyond = beyond;
      for (int j = 0; j < base && j < ring.nbuf; ++j) issue_tile(j, j);
    }
  }
  __syncthreads();
  int rel_seen = 0;  // relates met so far walking backwards
  int ring_slot = 0;
  uint32_t ring_phase = 0;
  // the first branch's final attention is the tape row of the PUSH instruction
  if (push_ip >= 0 && tid < n) sm.saved[tid] = tape[(long long)push_ip * tape_stride + tid];
  auto fetch_instr = [&](int ip) -> Instr {
    if (ip - ip0 < MAX_CODE) {
      const int32_t* c = code_s + (ip - ip0) * DFOL_INSTR_WORDS;
      Instr J;
      J.op = c[DFOL_I_OP]; J.flags = c[DFOL_I_FLAGS]; J.a0 = c[DFOL_I_A0]; J.a1 = c[DFOL_I_A1]; J.a2 = c[DFOL_I_A2];
      J.out = c[DFOL_I_OUT]; J.ga0 = c[DFOL_I_GA0]; J.ga1 = c[DFOL_I_GA1]; J.gr = c[DFOL_I_GR];
      J.mod = c[DFOL_I_MOD]; J.mod2 = c[DFOL_I_MOD2];
      return J;
    }
    return load_instr(instr, ip);
  };
#else
  // the first branch's final attention is the tape row of the PUSH instruction
  for (int ip = ip0; ip < ip1; ++ip)
    if (instr[(long long)ip * DFOL_INSTR_WORDS + DFOL_I_OP] == DFOL_OP_PUSH && tid < n)
      sm.saved[tid] = tape[(long long)ip * tape_stride + tid];
#endif
  __syncthreads();

  for (int ip = ip1 - 1; ip >= ip0; --ip) {
#ifdef DFOL_PROGRAM_FAST
    const Instr I = fetch_instr(ip);
#else
    const Instr I = load_instr(instr, ip);
#endif
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
        const bool subj = I.flags & DFOL_F_SUBJECT;
        float nw0 = 0.0f;  // prior of the new object before its modulation
#ifdef DFOL_PROGRAM_FAST
        const int j = rel_seen - rel_beyond;  // position among the ring-served relates, from the end
        ++rel_seen;
        if (j >= 0) {
          if (tid < n) {
            nw0 = (I.a1 >= 0) ? post_ll(attr_raw(im, I.a1, tid), nneg, nrt) : 0.0f;
            sm.nw[tid] = mod_apply(ms, nw0);
            sm.dres[tid] = sm.g[tid];
            sm.den[tid] = __expf(sm.cur[tid]);
          }
          mbar_wait(&ring.full[ring_slot], ring_phase);
          __syncthreads();
          float dm[4] = {0.f, 0.f, 0.f, 0.f};
          // the thread that ends up with object x's product turns d loss / d out[x] into the common factor of x's pair
          // gradients (through the modulation of the posterior when there is one)
          hop_backward<PTAB>(n, ring.buf + (size_t)ring_slot * ring.tile_floats, neg, sm.den, subj, sm.tot, sm.tmp,
                             g_rel + I.gr, sm.sc, [&](int x, float Q) -> float {
                               float d = sm.g[x];
                               if (mr.on) {
                                 d = mod_grad(mr, sm.nw[x] + slog(1.0f - Q), d, dm);
                                 sm.dres[x] = d;
                               }
                               return hop_cfactor(d, Q);
                             });
          if (tid == 0 && j + ring.nbuf < rel_count) issue_tile(j + ring.nbuf, ring_slot);
          if (++ring_slot == ring.nbuf) { ring_slot = 0; ring_phase ^= 1u; }
          if (mr.on) block_write_dm(dm, d_mods, I.mod, sm.sc);
        } else
#endif
        {
          if (tid < n) {
            nw0 = (I.a1 >= 0) ? post_ll(attr_raw(im, I.a1, tid), nneg, nrt) : 0.0f;
            sm.nw[tid] = mod_apply(ms, nw0);
            sm.dres[tid] = sm.g[tid];
            sm.tmp[tid] = 0.f;
          }
          __syncthreads();
          const float* a_s = subj ? sm.nw : sm.cur;
          const float* a_o = subj ? sm.cur : sm.nw;
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
#ifdef DFOL_PROGRAM_FAST
        const int32_t* op = stage_options(opts + I.a0, I.a1, opt_s);
#else
        const int32_t* op = opts + I.a0;
#endif
        float acc[NCHUNK];
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) acc[j] = 0.f;
        const bool modulated = MOD && I.mod >= 0;
        for_options<4>(im, op, I.a1, [&](int k, int word, const auto& raw) {
          const Mod mk = DFOL_LOAD_MOD(modulated ? I.mod + k : -1);
#pragma unroll
          for (int j = 0; j < DFOL_NC_OF(raw); ++j) {
            const int t = lane + 32 * j;
            if (t < n) acc[j] += mod_apply(mk, sm.cur[t] + post_ll(raw[j], word & DFOL_OPT_NEG, rt));
          }
        });
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
        for_options<4>(im, op, I.a1, [&](int k, int word, const auto& raws) {
          const Mod mk = DFOL_LOAD_MOD(modulated ? I.mod + k : -1);
          float dm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < DFOL_NC_OF(raws); ++j) {
            const int t = lane + 32 * j;
            if (t < n) {
              const float raw = raws[j];
              const float dx = mod_grad(mk, sm.cur[t] + post_ll(raw, word & DFOL_OPT_NEG, rt), sm.tmp[t], dm);
              acc[j] += dx;
              g_attr[I.ga0 + (long long)k * im.astride + t] = dx * post_ll_grad(raw, word & DFOL_OPT_NEG, rt);
            }
          }
          if (mk.on) warp_write_dm(dm, d_mods, I.mod + k);
        });
        if (modulated) reduce_columns(acc, n, sm.g, sm.sc, false);  // d cur = sum_k dx_k (unmodulated: a1 * datt, above)
        break;
      }

      case DFOL_OP_CHOOSE_ATTR:
      case DFOL_OP_ALL_SAME:
      case DFOL_OP_TWO_SAME: {
#ifdef DFOL_PROGRAM_FAST
        const int32_t* op = stage_options(opts + I.a0, I.a1, opt_s);
#else
        const int32_t* op = opts + I.a0;
#endif
        float* gslice = g_attr + I.ga0;
        if (normalise) bwd_option_denominators(im, op, I.a1, sm.den, sm.sc);
#ifdef DFOL_PROGRAM_FAST
        if (I.op == DFOL_OP_CHOOSE_ATTR && !(MOD && I.mod >= 0)) {
          // probability space (see options_pspace_nc): products, c_k, gradients of the (k, t) entries in one pass
          if (tid < MAXN) {
            float a = 0.f, wv = 0.f;
            if (tid < n) {
              a = __expf(sm.cur[tid]);
              wv = normalise ? __fdividef(a, fmaxf(sm.den[tid], kLogEps)) : a;
            }
            sm.inner[tid] = a;
            sm.res[tid] = wv;
          }
          __syncthreads();
          float acc_g[NCHUNK], acc_tot[NCHUNK];
#pragma unroll
          for (int j = 0; j < NCHUNK; ++j) { acc_g[j] = 0.f; acc_tot[j] = 0.f; }
          auto no_sink = [](int, float) {};
          if (n <= 64)
            options_pspace_nc<8, 2, true, false>(im, op, I.a1, sm.inner, sm.res, d_lp + I.out, gslice, acc_g, acc_tot, no_sink);
          else
            options_pspace_nc<4, NCHUNK, true, false>(im, op, I.a1, sm.inner, sm.res, d_lp + I.out, gslice, acc_g, acc_tot, no_sink);
          reduce_columns(acc_g, n, sm.g, sm.sc, false);
          if (normalise) {
            reduce_columns(acc_tot, n, sm.tot, sm.sc, false);
            attr_softmax_correction(im, op, I.a1, gslice, sm.den, sm.tot);
          }
          break;
        }
#endif
        float dQ = 0.f;
        if (I.op != DFOL_OP_CHOOSE_ATTR) {
          // first pass: Q = sum_k lnot(v_k) exactly as in the forward kernel
          float part = 0.f;
          for_options<4>(im, op, I.a1, [&](int k, int word, const auto& raw) {
            const bool kneg = word & DFOL_OPT_NEG;
            const Mod m1 = DFOL_LOAD_MOD(I.mod >= 0 ? I.mod + k : -1), m2 = DFOL_LOAD_MOD(I.mod2 >= 0 ? I.mod2 + k : -1);
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < DFOL_NC_OF(raw); ++j) {
              const int t = lane + 32 * j;
              if (t < n) {
                const float l = post_ll(option_nrm_of(raw[j], t, normalise, sm.den), kneg, rt);
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
          });
          const float Q = block_sum(lane == 0 ? part : 0.f, sm.sc);
          float dlp = d_lp[I.out];
          if (I.flags & DFOL_F_NEGATE_RESULT) dlp *= lnot_grad(lnot(Q));
          dQ = dlp * lnot_grad(Q);
        }
        float acc_g[NCHUNK], acc_gs[NCHUNK], acc_tot[NCHUNK];
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) { acc_g[j] = 0.f; acc_gs[j] = 0.f; acc_tot[j] = 0.f; }
        for_options<4>(im, op, I.a1, [&](int k, int word, const auto& raw) {
          constexpr int NC = DFOL_NC_OF(raw);
          const bool kneg = word & DFOL_OPT_NEG;
          const Mod m1 = DFOL_LOAD_MOD(I.mod >= 0 ? I.mod + k : -1), m2 = DFOL_LOAD_MOD(I.mod2 >= 0 ? I.mod2 + k : -1);
          float dm1[4] = {0.f, 0.f, 0.f, 0.f}, dm2[4] = {0.f, 0.f, 0.f, 0.f};
          float nrm[NC], l[NC], xm1[NC], xm2[NC];  // xm: modulated filter outputs
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int j = 0; j < NC; ++j) {
            const int t = lane + 32 * j;
            nrm[j] = 0.f; l[j] = 0.f; xm1[j] = 0.f; xm2[j] = 0.f;
            if (t < n) {
              nrm[j] = option_nrm_of(raw[j], t, normalise, sm.den);
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
          for (int j = 0; j < NC; ++j) {
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
        });
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
  #ifdef DFOL_PROGRAM_FAST
        const int32_t* op = stage_options(opts + I.a0, I.a1, opt_s);
#else
        const int32_t* op = opts + I.a0;
#endif
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
extern "C" int dfol_program_bwd_fast(const int32_t* instr, const int32_t* q_instr, const int32_t* opts,
                                     int question_num, const float* attr_ll, const int64_t* attr_blk,
                                     const int32_t* attr_stride, const float* rel_ll, const float* rel_p,
                                     const int64_t* rel_blk, const int32_t* rel_stride, const int32_t* img_n,
                                     const float* mods, const float* d_lp, const float* tape, int tape_stride,
                                     float* g_attr, float* g_rel, float* d_mods, void* stream) {
  DFOL_REQUIRE(instr && q_instr && attr_ll && attr_blk && attr_stride && rel_ll && rel_blk && rel_stride && img_n &&
                   d_lp && tape && g_attr && g_rel,
               "dfol_program_bwd_fast: null pointer");
  DFOL_REQUIRE((mods == nullptr) == (d_mods == nullptr), "dfol_program_bwd_fast: mods and d_mods go together");
  if (question_num == 0) return 0;
  DFOL_REQUIRE(tape_stride >= 1 && tape_stride <= MAXN, "dfol_program_bwd_fast: tape_stride = max objects rounded to 4");
  const int tile_floats = (tape_stride * tape_stride + 31) / 32 * 32;
  int nbuf = (96 * 1024) / (tile_floats * 4);
  nbuf = nbuf < 1 ? 1 : (nbuf > 4 ? 4 : nbuf);
  const size_t smem = (size_t)nbuf * tile_floats * 4;
  auto kern = mods ? (rel_p ? program_bwd_kernel<true, true> : program_bwd_kernel<true, false>)
                   : (rel_p ? program_bwd_kernel<false, true> : program_bwd_kernel<false, false>);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<question_num, PROG_THREADS, smem, (cudaStream_t)stream>>>(
      instr, q_instr, opts, attr_ll, attr_blk, attr_stride, rel_ll, rel_blk, rel_stride, img_n, mods, d_lp, tape,
      tape_stride, g_attr, g_rel, d_mods, rel_p, nbuf, tile_floats);
  return finish_launch("dfol_program_bwd_fast");
}
#else
extern "C" int dfol_program_bwd(const int32_t* instr, const int32_t* q_instr, const int32_t* opts, int question_num,
                                const float* attr_ll, const int64_t* attr_blk, const int32_t* attr_stride,
                                const float* rel_ll, const int64_t* rel_blk, const int32_t* rel_stride,
                                const int32_t* img_n, const float* mods, const float* d_lp, const float* tape,
                                int tape_stride, float* g_attr, float* g_rel, float* d_mods, void* stream) {
  DFOL_REQUIRE(instr && q_instr && attr_ll && attr_blk && attr_stride && rel_ll && rel_blk && rel_stride && img_n &&
                   d_lp && tape && g_attr && g_rel,
               "dfol_program_bwd: null pointer");
  DFOL_REQUIRE((mods == nullptr) == (d_mods == nullptr), "dfol_program_bwd: mods and d_mods go together");
  if (question_num == 0) return 0;
  (mods ? program_bwd_kernel<true, false> : program_bwd_kernel<false, false>)<<<question_num, PROG_THREADS, 0,
                                                                                 (cudaStream_t)stream>>>(
      instr, q_instr, opts, attr_ll, attr_blk, attr_stride, rel_ll, rel_blk, rel_stride, img_n, mods, d_lp, tape,
      tape_stride, g_attr, g_rel, d_mods);
  return finish_launch("dfol_program_bwd");
}
#endif
