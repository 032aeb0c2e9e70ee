// Forward program interpreter: one thread block per question walks its compiled FOL program; the attention
// vector stays in shared memory for the whole program, every table slice is read exactly once from its
// per-image contiguous block (see include/dfol_b200.h for the contract and reference citations).
#include "program_common.cuh"

#ifdef DFOL_PROG_TIMING  // timing experiment only: per-phase clock64 stamps of block 0 / thread 0 of the relate hops
__device__ long long dfol_tstamps[64 * 8];
__device__ int dfol_tcount;
#define DFOL_TSTAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && krel < 64) dfol_tstamps[krel * 8 + (i)] = clock64(); } while (0)
#else
#define DFOL_TSTAMP(i) do { } while (0)
#endif

namespace dfol {

// modulation row, compiled out of the unmodulated instantiation
#define DFOL_LOAD_MOD(row) (MOD ? load_mod(mods, (row)) : load_mod(nullptr, -1))

struct FwdShared {
  float cur[MAXN];
  float saved[MAXN];
  float nw[MAXN];
  float res[MAXN];
  float inner[MAXN];
  float den[MAXN];
  BlockScratch sc;
};

// cur-like vector of a select: zeros, plus the name predicate (GQASelectBatch, batch_gqa_ops.py:168-183)
__device__ __forceinline__ void select_into(float* dst, const Image& im, int col, bool neg, bool rt) {
  const int t = threadIdx.x;
  if (t < im.n) dst[t] = (col >= 0) ? post_ll(attr_raw(im, col, t), neg, rt) : 0.0f;
  __syncthreads();
}

// Attribute option lists (FilterBatch with a predicate->question map + ClassifierOracle normalisation,
// classifier_oracle.py:61-80): den[t] = sum_k exp(raw_k[t]).
__device__ __forceinline__ void option_denominators(const Image& im, const int32_t* opts, int count, float* den,
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

// post-processed likelihood of option `word` at object t from its already loaded raw table entry
__device__ __forceinline__ float option_ll_of(float r, int word, int t, bool normalise, bool rt, const float* den) {
  if (normalise) r -= slog(den[t]);
  return post_ll(r, (word & DFOL_OPT_NEG) != 0, rt);
}

__device__ __forceinline__ float option_ll(const Image& im, int word, int t, bool normalise, bool rt,
                                           const float* den) {
  float r = attr_raw(im, word & ~DFOL_OPT_NEG, t);
  if (normalise) r -= slog(den[t]);
  return post_ll(r, (word & DFOL_OPT_NEG) != 0, rt);
}

// warp-level exists over x(t) = base[t] + ll_k[t]; lanes own t = lane + 32 j. Returns lnot(S) to all lanes.
__device__ __forceinline__ float warp_exists(const float x[NCHUNK], int n, bool hard, float* s_out) {
  const int lane = threadIdx.x & 31;
  float s = hard ? 0.0f : 0.0f;
  bool first = true;
#pragma unroll
  for (int j = 0; j < NCHUNK; ++j) {
    const int t = lane + 32 * j;
    if (t < n) {
      const float v = lnot(x[j]);
      if (hard) { s = first ? v : fminf(s, v); first = false; }
      else s += v;
    }
  }
  s = hard ? warp_min(s) : warp_sum(s);
  if (s_out) *s_out = s;
  return lnot(s);
}

template <bool MOD, bool PTAB>
static __global__ void __launch_bounds__(PROG_THREADS, MOD ? 1 : PROG_MIN_BLOCKS) program_fwd_kernel(
    const int32_t* __restrict__ instr, const int32_t* __restrict__ q_instr, const int32_t* __restrict__ opts,
    const float* __restrict__ attr_ll, const int64_t* __restrict__ attr_blk, const int32_t* __restrict__ attr_stride,
    const float* __restrict__ rel_ll, const int64_t* __restrict__ rel_blk, const int32_t* __restrict__ rel_stride,
    const int32_t* __restrict__ img_n, const float* __restrict__ mods, float* __restrict__ lp_out,
    float* __restrict__ tape, int tape_stride
#ifdef DFOL_PROGRAM_FAST
    , const float* __restrict__ rel_p, int ring_nbuf, int ring_tile_floats
#endif
    ) {
  __shared__ __align__(16) FwdShared sm;
#ifdef DFOL_PROG_TIMING
  if (blockIdx.x == 0 && threadIdx.x == 0) dfol_tstamps[6] = clock64();
#endif
  const int q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  Image im;
  im.n = img_n[q];
  im.attr = attr_ll + attr_blk[q];
  im.astride = attr_stride[q];
  im.rel = rel_ll + rel_blk[q];
  im.rstride = rel_stride[q];
  const int n = im.n;

  if (tid < MAXN) { sm.cur[tid] = 0.f; sm.saved[tid] = 0.f; sm.den[tid] = 0.f; }
#ifdef DFOL_PROGRAM_FAST
  // the ring streams the probability tiles when the scene supplies them (same blocks and strides as the log table)
  const float* ring_src = PTAB ? rel_p + rel_blk[q] : im.rel;
  // relate tiles of this program, in execution order, streamed through the shared-memory ring
  extern __shared__ __align__(128) float ring_mem[];
  __shared__ __align__(8) uint64_t ring_full[8];
  __shared__ int rel_ip[MAX_REL];
  __shared__ int rel_count;
  TileRing ring{ring_mem, ring_nbuf, ring_tile_floats, ring_full};
  // the question's bytecode is read once into shared memory (no global load on the per-instruction critical path),
  // together with the attribute column each instruction will want prefetched and the list of its relate hops
  __shared__ int32_t code_s[MAX_CODE * DFOL_INSTR_WORDS];
  __shared__ int pre_col_s[MAX_CODE];
  __shared__ int32_t opt_s[OPT_STAGE];  // option words of the instruction being executed
  const int ip_first = q_instr[q];
  const int ip_last = q_instr[q + 1];
  const int code_n = min(ip_last - ip_first, MAX_CODE);
  for (int i = tid; i < code_n * DFOL_INSTR_WORDS; i += PROG_THREADS)
    code_s[i] = instr[(long long)ip_first * DFOL_INSTR_WORDS + i];
  if (tid == 0) {
    for (int b = 0; b < ring.nbuf; ++b) mbar_init(&ring.full[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue_tile = [&](int k, int b) {  // elected thread: tile of the k-th relate -> ring slot b = k % nbuf
    const int col = code_s[(rel_ip[k] - ip_first) * DFOL_INSTR_WORDS + DFOL_I_A0];
    const uint32_t bytes = (uint32_t)im.rstride * 4u;
    mbar_expect_tx(&ring.full[b], bytes);
    bulk_load(ring.buf + (size_t)b * ring.tile_floats, ring_src + (long long)col * im.rstride, bytes, &ring.full[b]);
  };
  if (tid < code_n) {
    const int op = code_s[tid * DFOL_INSTR_WORDS + DFOL_I_OP];
    pre_col_s[tid] = (op == DFOL_OP_SELECT || op == DFOL_OP_FILTER) ? code_s[tid * DFOL_INSTR_WORDS + DFOL_I_A0]
                     : (op == DFOL_OP_RELATE ? code_s[tid * DFOL_INSTR_WORDS + DFOL_I_A1] : -1);
  }
  if (w == 0) {  // warp 0 compacts the relate hops of the staged instructions (ballot + prefix popcount)
    int base = 0;
    for (int c0 = 0; c0 < code_n; c0 += 32) {
      const int idx = c0 + lane;
      const bool is_rel = idx < code_n && code_s[idx * DFOL_INSTR_WORDS + DFOL_I_OP] == DFOL_OP_RELATE;
      const unsigned m = __ballot_sync(0xffffffffu, is_rel);
      if (is_rel) rel_ip[base + __popc(m & ((1u << lane) - 1u))] = ip_first + idx;
      base += __popc(m);
    }
    __syncwarp();
    if (lane == 0) {
      rel_count = base;
      for (int k = 0; k < base && k < ring.nbuf; ++k) issue_tile(k, k);
    }
  }
  int krel = 0;
  int ring_slot = 0;          // slot of the krel-th relate and the parity of its mbarrier phase (no division per hop)
  uint32_t ring_phase = 0;
#endif
  __syncthreads();

#ifdef DFOL_PROGRAM_FAST
  // software prefetch of the single attribute row the NEXT instruction reads (select / filter / relate name):
  // the load is issued one instruction early, so its latency hides behind the current instruction
  auto fetch_instr = [&](int ip) -> Instr {
    if (ip - ip_first < MAX_CODE) {  // shared-memory copy, indexed directly (LDS); the forward pass needs six words
      const int o = (ip - ip_first) * DFOL_INSTR_WORDS;
      Instr J;
      J.op = code_s[o + DFOL_I_OP]; J.flags = code_s[o + DFOL_I_FLAGS]; J.a0 = code_s[o + DFOL_I_A0];
      J.a1 = code_s[o + DFOL_I_A1]; J.a2 = code_s[o + DFOL_I_A2]; J.out = code_s[o + DFOL_I_OUT];
      J.ga0 = J.ga1 = J.gr = -1;
      J.mod = code_s[o + DFOL_I_MOD]; J.mod2 = code_s[o + DFOL_I_MOD2];
      return J;
    }
    return load_instr(instr, ip);
  };
  auto operand_col_at = [&](int ip) -> int {
    if (ip - ip_first < MAX_CODE) return pre_col_s[ip - ip_first];
    const Instr J = load_instr(instr, ip);
    if (J.op == DFOL_OP_SELECT || J.op == DFOL_OP_FILTER) return J.a0;
    return J.op == DFOL_OP_RELATE ? J.a1 : -1;
  };
  float pre_raw = 0.f;
  if (ip_first < ip_last) {
    const int col = operand_col_at(ip_first);
    if (col >= 0 && tid < n) pre_raw = attr_raw(im, col, tid);
  }
#endif
#ifdef DFOL_PROGRAM_FAST
  const int ip_begin = ip_first, ip_end = ip_last;
#else
  const int ip_begin = q_instr[q], ip_end = q_instr[q + 1];
#endif
#ifdef DFOL_PROG_TIMING
  if (blockIdx.x == 0 && threadIdx.x == 0) dfol_tstamps[7] = clock64();
#endif
  for (int ip = ip_begin; ip < ip_end; ++ip) {
#ifdef DFOL_PROGRAM_FAST
    const Instr I = fetch_instr(ip);
    const float cur_raw = pre_raw;  // operand row of THIS instruction (valid where operand_col(I) >= 0)
    if (ip + 1 < ip_end) {
      const int col = operand_col_at(ip + 1);
      if (col >= 0 && tid < n) pre_raw = attr_raw(im, col, tid);
    }
#else
    const Instr I = load_instr(instr, ip);
#endif
    const bool neg = I.flags & DFOL_F_NEG, rt = I.flags & DFOL_F_ROUNDTRIP;
    const bool hard = I.flags & DFOL_F_HARD, normalise = I.flags & DFOL_F_NORMALISE;
    if (tape != nullptr && tid < n) tape[(long long)ip * tape_stride + tid] = sm.cur[tid];

    switch (I.op) {
      case DFOL_OP_SELECT: {
        const Mod m = DFOL_LOAD_MOD(I.mod);
#ifdef DFOL_PROGRAM_FAST
        if (tid < n) sm.cur[tid] = mod_apply(m, (I.a0 >= 0) ? post_ll(cur_raw, neg, rt) : 0.0f);
        __syncthreads();
#else
        select_into(sm.cur, im, I.a0, neg, rt);
        if (m.on) {
          if (tid < n) sm.cur[tid] = mod_apply(m, sm.cur[tid]);
          __syncthreads();
        }
#endif
        break;
      }

      case DFOL_OP_FILTER: {  // a'[t] = a[t] + ll[t]  (_forward_core arity 1); a0 < 0: modulated pass-through
        const Mod m = DFOL_LOAD_MOD(I.mod);
#ifdef DFOL_PROGRAM_FAST
        if (tid < n) sm.cur[tid] = mod_apply(m, sm.cur[tid] + ((I.a0 >= 0) ? post_ll(cur_raw, neg, rt) : 0.0f));
#else
        if (tid < n)
          sm.cur[tid] = mod_apply(m, sm.cur[tid] + ((I.a0 >= 0) ? post_ll(attr_raw(im, I.a0, tid), neg, rt) : 0.0f));
#endif
        __syncthreads();
        break;
      }

      case DFOL_OP_PUSH:
        if (tid < n) sm.saved[tid] = sm.cur[tid];
        __syncthreads();
        break;

      case DFOL_OP_RELATE: {
        const bool subj = I.flags & DFOL_F_SUBJECT;
#ifdef DFOL_PROGRAM_FAST
        if (krel < rel_count) {
          // phase A (thread-local): prior of the new object from the prefetched name row, e^{cur} of the other role
          DFOL_TSTAMP(0);
          const Mod mr = DFOL_LOAD_MOD(I.mod), ms = DFOL_LOAD_MOD(I.mod2);
          if (tid < n) {
            float nwv = 0.0f;
            if (I.a1 >= 0) nwv = post_ll(cur_raw, I.flags & DFOL_F_NAME_NEG, I.flags & DFOL_F_NAME_ROUNDTRIP);
            sm.nw[tid] = mod_apply(ms, nwv);
            sm.den[tid] = __expf(sm.cur[tid]);
          }
          DFOL_TSTAMP(1);
          mbar_wait(&ring.full[ring_slot], ring_phase);
          DFOL_TSTAMP(2);
          __syncthreads();
          DFOL_TSTAMP(3);
          // phase B: products over the tile; the thread that ends up holding object x's product writes its posterior,
          // which replaces the attention (the products read e^{cur} in sm.den, never sm.cur)
          hop_forward<PTAB>(n, ring.buf + (size_t)ring_slot * ring.tile_floats, neg, sm.den, subj, sm.sc,
                            [&](int x, float Q) { sm.cur[x] = mod_apply(mr, sm.nw[x] + slog(1.0f - Q)); });
          DFOL_TSTAMP(4);
          // every thread is past its last read of the slot: refill it with the tile nbuf hops ahead
          if (tid == 0 && krel + ring.nbuf < rel_count) issue_tile(krel + ring.nbuf, ring_slot);
          DFOL_TSTAMP(5);
          ++krel;
          if (++ring_slot == ring.nbuf) { ring_slot = 0; ring_phase ^= 1u; }
          break;
        }
        ++krel;
        if (tid < n)
          sm.nw[tid] = mod_apply(DFOL_LOAD_MOD(I.mod2), (I.a1 >= 0) ? post_ll(cur_raw, I.flags & DFOL_F_NAME_NEG,
                                                                               I.flags & DFOL_F_NAME_ROUNDTRIP) : 0.0f);
        __syncthreads();
#else
        select_into(sm.nw, im, I.a1, I.flags & DFOL_F_NAME_NEG, I.flags & DFOL_F_NAME_ROUNDTRIP);
        {
          const Mod ms = DFOL_LOAD_MOD(I.mod2);
          if (ms.on) {
            if (tid < n) sm.nw[tid] = mod_apply(ms, sm.nw[tid]);
            __syncthreads();
          }
        }
#endif
        {
          RelOption L{&im, nullptr, 1, 0, false, rt, I.a0, neg};
          relate_forward(n, L, subj ? sm.nw : sm.cur, subj ? sm.cur : sm.nw, subj, sm.res, sm.inner, sm.sc);
        }
        if (tid < n) sm.cur[tid] = mod_apply(DFOL_LOAD_MOD(I.mod), sm.res[tid]);
        __syncthreads();
        break;
      }

      case DFOL_OP_EXIST: {
        const float lp = exists_block(sm.cur, n, hard, sm.sc, nullptr);
        if (tid == 0) lp_out[I.out] = lp;
        break;
      }

      case DFOL_OP_AND:
      case DFOL_OP_OR: {
        const float e1 = exists_block(sm.saved, n, hard, sm.sc, nullptr);
        const float e2 = exists_block(sm.cur, n, hard, sm.sc, nullptr);
        if (tid == 0)
          lp_out[I.out] = (I.op == DFOL_OP_AND) ? e1 + e2 : slog(1.0f - (1.0f - DFOL_EXPF(e1)) * (1.0f - DFOL_EXPF(e2)));
        break;
      }

      case DFOL_OP_VERIFY_ATTRS: {
        // sum over the question's attributes of (a + ll_k), then exists (GQAVerifyAttrsBatch :452-473)
#ifdef DFOL_PROGRAM_FAST
        const int32_t* op = stage_options(opts + I.a0, I.a1, opt_s);
#else
        const int32_t* op = opts + I.a0;
#endif
        float acc[NCHUNK];
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) acc[j] = 0.f;
        for_options<4>(im, op, I.a1, [&](int k, int word, const auto& raw) {
          const Mod mk = DFOL_LOAD_MOD(I.mod >= 0 ? I.mod + k : -1);
#pragma unroll
          for (int j = 0; j < DFOL_NC_OF(raw); ++j) {
            const int t = lane + 32 * j;
            if (t < n) acc[j] += mod_apply(mk, sm.cur[t] + option_ll_of(raw[j], word, t, false, rt, nullptr));
          }
        });
        reduce_columns(acc, n, sm.res, sm.sc, false);
        const float lp = exists_block(sm.res, n, hard, sm.sc, nullptr);
        if (tid == 0) lp_out[I.out] = lp;
        break;
      }

      case DFOL_OP_CHOOSE_ATTR: {
#ifdef DFOL_PROGRAM_FAST
        const int32_t* op = stage_options(opts + I.a0, I.a1, opt_s);
#else
        const int32_t* op = opts + I.a0;
#endif
        if (normalise) option_denominators(im, op, I.a1, sm.den, sm.sc);
#ifdef DFOL_PROGRAM_FAST
        if (!hard && !(MOD && I.mod >= 0)) {
          // probability space: a_t = e^{cur}, w_t = a_t / den_t; 8 options per warp pass, one transposed reduction
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
          float unused_g[NCHUNK], unused_t[NCHUNK];
          auto sink = [&](int k, float Q) { lp_out[I.out + k] = slog(1.0f - Q); };
          if (n <= 64)
            options_pspace_nc<8, 2, false, true>(im, op, I.a1, sm.inner, sm.res, nullptr, nullptr, unused_g, unused_t, sink);
          else
            options_pspace_nc<4, NCHUNK, false, true>(im, op, I.a1, sm.inner, sm.res, nullptr, nullptr, unused_g, unused_t, sink);
          __syncthreads();
          break;
        }
#endif
        for_options<8>(im, op, I.a1, [&](int k, int word, const auto& raw) {
          const Mod mk = DFOL_LOAD_MOD(I.mod >= 0 ? I.mod + k : -1);
          float x[NCHUNK];
#pragma unroll
          for (int j = 0; j < NCHUNK; ++j) {
            const int t = lane + 32 * j;
            x[j] = 0.f;
            if (j < DFOL_NC_OF(raw) && t < n)
              x[j] = mod_apply(mk, sm.cur[t] + option_ll_of(raw[j < DFOL_NC_OF(raw) ? j : 0], word, t, normalise, rt, sm.den));
          }
          const float lp = warp_exists(x, n, hard, nullptr);
          if (lane == 0) lp_out[I.out + k] = lp;
        });
        __syncthreads();
        break;
      }

      case DFOL_OP_ALL_SAME: {
        // per option: q_k = forall_t lnot(a + lnot(a + ll_k)); lp = lnot(sum_k lnot(q_k))  (:582-608)
#ifdef DFOL_PROGRAM_FAST
        const int32_t* op = stage_options(opts + I.a0, I.a1, opt_s);
#else
        const int32_t* op = opts + I.a0;
#endif
        if (normalise) option_denominators(im, op, I.a1, sm.den, sm.sc);
        float part = 0.f;
        for_options<4>(im, op, I.a1, [&](int k, int word, const auto& raw) {
          const Mod mk = DFOL_LOAD_MOD(I.mod >= 0 ? I.mod + k : -1);
          float s = 0.f;
          bool first = true;
#pragma unroll
          for (int j = 0; j < DFOL_NC_OF(raw); ++j) {
            const int t = lane + 32 * j;
            if (t < n) {
              const float a = sm.cur[t];
              const float y = lnot(a + lnot(mod_apply(mk, a + option_ll_of(raw[j], word, t, normalise, rt, sm.den))));
              const float r = roundtrip(y);
              if (hard) { s = first ? r : fminf(s, r); first = false; }
              else s += r;
            }
          }
          s = hard ? warp_min(s) : warp_sum(s);
          part += lnot(roundtrip(s));
        });
        float Q = block_sum(lane == 0 ? part : 0.f, sm.sc);
        float lp = lnot(Q);
        if (I.flags & DFOL_F_NEGATE_RESULT) lp = lnot(lp);
        if (tid == 0) lp_out[I.out] = lp;
        break;
      }

      case DFOL_OP_TWO_SAME: {
#ifdef DFOL_PROGRAM_FAST
        const int32_t* op = stage_options(opts + I.a0, I.a1, opt_s);
#else
        const int32_t* op = opts + I.a0;
#endif
        if (normalise) option_denominators(im, op, I.a1, sm.den, sm.sc);
        float part = 0.f;
        for_options<4>(im, op, I.a1, [&](int k, int word, const auto& raw) {
          const Mod m1 = DFOL_LOAD_MOD(I.mod >= 0 ? I.mod + k : -1), m2 = DFOL_LOAD_MOD(I.mod2 >= 0 ? I.mod2 + k : -1);
          float x1[NCHUNK], x2[NCHUNK];
#pragma unroll
          for (int j = 0; j < NCHUNK; ++j) {
            const int t = lane + 32 * j;
            x1[j] = x2[j] = 0.f;
            if (j < DFOL_NC_OF(raw) && t < n) {
              const float l = option_ll_of(raw[j < DFOL_NC_OF(raw) ? j : 0], word, t, normalise, rt, sm.den);
              x1[j] = mod_apply(m1, sm.saved[t] + l);
              x2[j] = mod_apply(m2, sm.cur[t] + l);
            }
          }
          const float e1 = warp_exists(x1, n, hard, nullptr);
          const float e2 = warp_exists(x2, n, hard, nullptr);
          part += lnot(e1 + e2);
        });
        float Q = block_sum(lane == 0 ? part : 0.f, sm.sc);
        float lp = lnot(Q);
        if (I.flags & DFOL_F_NEGATE_RESULT) lp = lnot(lp);
        if (tid == 0) lp_out[I.out] = lp;
        break;
      }

      case DFOL_OP_COMPARE: {
        // filter both branches by the attribute, exists, log-softmax over the two, flip by is_less (:730-738)
        if (tid < n) {
          const float l = (I.a0 >= 0) ? post_ll(attr_raw(im, I.a0, tid), neg, rt) : 0.0f;
          sm.res[tid] = mod_apply(DFOL_LOAD_MOD(I.mod), sm.saved[tid] + l);
          sm.nw[tid] = mod_apply(DFOL_LOAD_MOD(I.mod2), sm.cur[tid] + l);
        }
        __syncthreads();
        const float e1 = exists_block(sm.res, n, hard, sm.sc, nullptr);
        const float e2 = exists_block(sm.nw, n, hard, sm.sc, nullptr);
        if (tid == 0) {
          const float mx = fmaxf(e1, e2);
          const float lse = DFOL_LOGF(DFOL_EXPF(e1 - mx) + DFOL_EXPF(e2 - mx));
          const float z1 = e1 - mx - lse, z2 = e2 - mx - lse;
          const float alpha = (I.flags & DFOL_F_IS_LESS) ? 1.0f : 0.0f;
          lp_out[I.out] = slog(alpha + (1.0f - 2.0f * alpha) * DFOL_EXPF(z1));
          lp_out[I.out + 1] = slog(alpha + (1.0f - 2.0f * alpha) * DFOL_EXPF(z2));
        }
        break;
      }

      case DFOL_OP_CHOOSE_REL: {
        select_into(sm.nw, im, I.a2, I.flags & DFOL_F_NAME_NEG, I.flags & DFOL_F_NAME_ROUNDTRIP);
        {
          const Mod ms = DFOL_LOAD_MOD(I.mod2);
          if (ms.on) {
            if (tid < n) sm.nw[tid] = mod_apply(ms, sm.nw[tid]);
            __syncthreads();
          }
        }
        const bool subj = I.flags & DFOL_F_SUBJECT;
        for (int k = 0; k < I.a1; ++k) {
          RelOption L{&im, opts + I.a0, I.a1, k, normalise, rt, -1, false};
          relate_forward(n, L, subj ? sm.nw : sm.cur, subj ? sm.cur : sm.nw, subj, sm.res, sm.inner, sm.sc);
          const Mod mk = DFOL_LOAD_MOD(I.mod >= 0 ? I.mod + k : -1);
          if (mk.on) {
            if (tid < n) sm.res[tid] = mod_apply(mk, sm.res[tid]);
            __syncthreads();
          }
          const float lp = exists_block(sm.res, n, hard, sm.sc, nullptr);
          if (tid == 0) lp_out[I.out + k] = lp;
          __syncthreads();
        }
        break;
      }

      default:
        break;
    }
  }
#ifdef DFOL_PROG_TIMING
  if (blockIdx.x == 0 && threadIdx.x == 0) dfol_tstamps[8 + 6] = clock64();
#endif
}

}  // namespace dfol

using namespace dfol;


#ifdef DFOL_PROG_TIMING
extern "C" int dfol_prog_timing_read(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, dfol_tstamps, sizeof(long long) * n);
}
#endif

#ifdef DFOL_PROGRAM_FAST
extern "C" int dfol_program_fwd_fast(const int32_t* instr, const int32_t* q_instr, const int32_t* opts,
                                     int question_num, const float* attr_ll, const int64_t* attr_blk,
                                     const int32_t* attr_stride, const float* rel_ll, const float* rel_p,
                                     const int64_t* rel_blk, const int32_t* rel_stride, const int32_t* img_n,
                                     const float* mods, float* lp_out, float* tape, int tape_stride, void* stream) {
  DFOL_REQUIRE(instr && q_instr && attr_ll && attr_blk && attr_stride && rel_ll && rel_blk && rel_stride && img_n &&
                   lp_out,
               "dfol_program_fwd_fast: null pointer");
  if (question_num == 0) return 0;
  // ring of relation-tile buffers: as many (up to 4) as fit in ~96 KB so that two blocks share an SM
  DFOL_REQUIRE(tape_stride >= 1 && tape_stride <= MAXN, "dfol_program_fwd_fast: tape_stride = max objects rounded to 4");
  const int tile_floats = (tape_stride * tape_stride + 31) / 32 * 32;
  int nbuf = (96 * 1024) / (tile_floats * 4);
  nbuf = nbuf < 1 ? 1 : (nbuf > 4 ? 4 : nbuf);
  const size_t smem = (size_t)nbuf * tile_floats * 4;
  auto kern = mods ? (rel_p ? program_fwd_kernel<true, true> : program_fwd_kernel<true, false>)
                   : (rel_p ? program_fwd_kernel<false, true> : program_fwd_kernel<false, false>);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<question_num, PROG_THREADS, smem, (cudaStream_t)stream>>>(
      instr, q_instr, opts, attr_ll, attr_blk, attr_stride, rel_ll, rel_blk, rel_stride, img_n, mods, lp_out, tape,
      tape_stride, rel_p, nbuf, tile_floats);
  return finish_launch("dfol_program_fwd_fast");
}
#else
extern "C" int dfol_program_fwd(const int32_t* instr, const int32_t* q_instr, const int32_t* opts, int question_num,
                                const float* attr_ll, const int64_t* attr_blk, const int32_t* attr_stride,
                                const float* rel_ll, const int64_t* rel_blk, const int32_t* rel_stride,
                                const int32_t* img_n, const float* mods, float* lp_out, float* tape, int tape_stride,
                                void* stream) {
  DFOL_REQUIRE(instr && q_instr && attr_ll && attr_blk && attr_stride && rel_ll && rel_blk && rel_stride && img_n &&
                   lp_out,
               "dfol_program_fwd: null pointer");
  DFOL_REQUIRE(tape == nullptr || tape_stride >= 1, "dfol_program_fwd: bad tape stride");
  if (question_num == 0) return 0;
  (mods ? program_fwd_kernel<true, false> : program_fwd_kernel<false, false>)<<<question_num, PROG_THREADS, 0,
                                                                                 (cudaStream_t)stream>>>(
      instr, q_instr, opts, attr_ll, attr_blk, attr_stride, rel_ll, rel_blk, rel_stride, img_n, mods, lp_out, tape,
      tape_stride);
  return finish_launch("dfol_program_fwd");
}
#endif
