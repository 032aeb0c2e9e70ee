// Attention-transfer calibrator, token side, as TWO persistent kernels (forward tape, backward tape).
//
// The LSTMCell passes over the op slots are a chain of ~100 small dependent steps per batch (one cell / output layer /
// gate / squeeze per slot and pass, each over a few hundred to a few thousand predicate rows of 50-wide states;
// reference: the modulator loops of BatchInterpreterBase.forward, batch_base_interpreter.py:87-140, and
// transform_attention of the operator modules, batch_base_ops.py:407-467, :598-684).  As one launch per step
// (csrc/modulator_kernels.cu) every step paid a launch, re-staged W_hh (40 KB) in shared memory and left the host to
// issue ~100 launches plus the allocations between them: ~1.2 ms of a 3.3 ms calibrator-training step.
//
// Here the host compiles the SAME sequence of steps into an array of records once per program batch (program-only
// data: pool offsets, row counts, owner / mask tables) and ONE kernel per pass walks it with both W_hh matrices and the
// output layer resident in shared memory.  The chain is sequential per QUESTION only -- no step mixes the rows of two
// questions -- so every block owns a contiguous range of questions and executes the whole tape for the rows of its
// own questions: a step's rows are either one per question or the (question-sorted) predicate rows of an option list,
// whose sub-range a block finds by binary search in the step's `part` map.  Dependent steps are separated by
// __syncthreads() only: no grid barrier, no cooperative launch.  The backward kernel walks the tape in reverse (BPTT)
// with the hand-derived cell / output-layer backward of modulator_kernels.cu; the parameter gradients stay
// GEMM-shaped reductions over d pre of all cell rows, done by the caller afterwards.
//
// States live in ONE fp32 pool (offsets in floats, -1 = the all-zero state); the gradient pool has the same layout.
#include "dfol_common.cuh"

namespace dfol {

constexpr int TAPE_MAXS = 64;
constexpr int TAPE_ROWS = 8;                       // rows per block and chunk
constexpr int TAPE_THREADS = TAPE_MAXS * TAPE_ROWS;

enum { TAPE_CELL = 0, TAPE_OUT = 1, TAPE_SQUEEZE = 2, TAPE_GATE = 3 };

// One step of the tape (mirrors the tuples NativeAttentionTransfer records; 112 bytes).
struct TapeRec {
  int32_t kind;      // TAPE_*
  int32_t net;       // CELL: 0 forward network, 1 backward network
  int32_t rows;      // rows of the step
  int32_t base;      // first row of the step in xproj / saved / dpre (CELL) or mods / cat / dzo (OUT)
  int32_t live;      // backward: 0 = nothing downstream depends on this step's output
  int32_t pad;
  int64_t in_h, in_c;      // CELL: incoming state; OUT: forward state; SQUEEZE: source; GATE: the new state
  int64_t add_h, add_c;    // CELL: state added to the incoming one; OUT: backward state; GATE: the old state
  int64_t fb_h, fb_c;      // CELL with a mask: state passed through where mask == 0
  int64_t out_h, out_c;    // result
  const int64_t* owner;    // CELL: row -> source row (expand); SQUEEZE: row -> destination row
  const float* mask;       // CELL / GATE: 0/1 per row
  const int64_t* part;     // row -> question (non-decreasing), NULL = row i belongs to question i
};

// rows [r0, r1) of a step that belong to the questions [q0, q1) of this block
__device__ __forceinline__ int tape_lower_bound(const int64_t* part, int rows, int q) {
  int lo = 0, hi = rows;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (part[mid] < q) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ void tape_row_range(const TapeRec& R, int q0, int q1, int* range) {
  // (called by every thread between two barriers; thread 0 publishes, the caller's next barrier makes it visible)
  if (threadIdx.x == 0) {
    if (R.part == nullptr) { range[0] = min(q0, R.rows); range[1] = min(q1, R.rows); }
    else { range[0] = tape_lower_bound(R.part, R.rows, q0); range[1] = tape_lower_bound(R.part, R.rows, q1); }
  }
}

struct TapeNets {
  const float* w_hh[2];   // [4S][S]
  const float* b_hh[2];
  const float* w_out;     // [n_out][2S]
  const float* b_out;
  int S, n_out;
};

__device__ __forceinline__ float tape_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ const float* at(const float* pool, int64_t off) { return off >= 0 ? pool + off : nullptr; }
__device__ __forceinline__ float* at(float* pool, int64_t off) { return off >= 0 ? pool + off : nullptr; }

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(TAPE_THREADS, 2) mod_tape_fwd_kernel(
    const TapeRec* __restrict__ recs, int n_rec, float* __restrict__ pool, const float* __restrict__ xproj_f,
    const float* __restrict__ xproj_b, TapeNets nets, float* __restrict__ saved_f, float* __restrict__ saved_b,
    float* __restrict__ mods, float* __restrict__ cat, int questions) {
  extern __shared__ float smem[];
  __shared__ int range[2];
  const int S = nets.S, n_out = nets.n_out;
  float* wt[2] = {smem, smem + 4 * S * S};          // W_hh transposed: wt[k * 4S + g * S + j]
  float* wo = smem + 8 * S * S;                     // [n_out][2S]
  float* bo = wo + n_out * 2 * S;
  float* bh[2] = {bo + n_out, bo + n_out + 4 * S};
  __shared__ float hin[TAPE_ROWS][TAPE_MAXS];
  for (int net = 0; net < 2; ++net) {
    for (int idx = threadIdx.x; idx < 4 * S * S; idx += TAPE_THREADS) {
      const int gj = idx / S, k = idx - gj * S;
      wt[net][k * 4 * S + gj] = nets.w_hh[net][idx];
    }
    for (int idx = threadIdx.x; idx < 4 * S; idx += TAPE_THREADS) bh[net][idx] = nets.b_hh[net][idx];
  }
  for (int idx = threadIdx.x; idx < n_out * 2 * S; idx += TAPE_THREADS) wo[idx] = nets.w_out[idx];
  for (int idx = threadIdx.x; idx < n_out; idx += TAPE_THREADS) bo[idx] = nets.b_out[idx];
  __syncthreads();
  const int j = threadIdx.x % TAPE_MAXS, lr = threadIdx.x / TAPE_MAXS;
  const int q0 = (int)(((long long)questions * blockIdx.x) / gridDim.x);
  const int q1 = (int)(((long long)questions * (blockIdx.x + 1)) / gridDim.x);

  for (int r = 0; r < n_rec; ++r) {
    const TapeRec R = recs[r];
    tape_row_range(R, q0, q1, range);
    __syncthreads();  // the previous step's results and this step's row range are visible to the whole block
    const int r0 = range[0], r1 = range[1];
    if (R.kind == TAPE_CELL) {
      const float* xproj = (R.net == 0 ? xproj_f : xproj_b) + (long long)R.base * 4 * S;
      float* saved = (R.net == 0 ? saved_f : saved_b) + (long long)R.base * 7 * S;
      const float* h_in = at(pool, R.in_h);
      const float* c_in = at(pool, R.in_c);
      const float* h_add = at(pool, R.add_h);
      const float* c_add = at(pool, R.add_c);
      const float* fb_h = at(pool, R.fb_h);
      const float* fb_c = at(pool, R.fb_c);
      float* h_out = at(pool, R.out_h);
      float* c_out = at(pool, R.out_c);
      const float* w = wt[R.net];
      const float* b = bh[R.net];
      for (int c0 = r0; c0 < r1; c0 += TAPE_ROWS) {
        const int row = c0 + lr;
        const bool act = row < r1 && j < S;
        long long src = 0;
        float cin = 0.f;
        __syncthreads();  // hin of the previous chunk has been consumed
        if (act) {
          src = R.owner ? R.owner[row] : row;
          float h = h_in ? h_in[src * S + j] : 0.f;
          cin = c_in ? c_in[src * S + j] : 0.f;
          if (h_add) { h += h_add[src * S + j]; cin += c_add[src * S + j]; }
          hin[lr][j] = h;
        }
        __syncthreads();
        if (act) {
          float pre[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) pre[g] = xproj[(long long)row * 4 * S + g * S + j] + b[g * S + j];
          for (int k = 0; k < S; ++k) {
            const float h = hin[lr][k];
            const float* wk = w + k * 4 * S + j;
#pragma unroll
            for (int g = 0; g < 4; ++g) pre[g] = fmaf(h, wk[g * S], pre[g]);
          }
          const float gi = tape_sigmoid(pre[0]), gf = tape_sigmoid(pre[1]), gg = tanhf(pre[2]), go = tape_sigmoid(pre[3]);
          const float cn = gf * cin + gi * gg;
          const float tc = tanhf(cn);
          float ho = go * tc, co = cn;
          if (R.mask && !(R.mask[row] > 0.f)) {
            ho = fb_h ? fb_h[(long long)row * S + j] : 0.f;
            co = fb_c ? fb_c[(long long)row * S + j] : 0.f;
          }
          h_out[(long long)row * S + j] = ho;
          c_out[(long long)row * S + j] = co;
          float* sv = saved + (long long)row * 7 * S;
          sv[j] = gi; sv[S + j] = gf; sv[2 * S + j] = gg; sv[3 * S + j] = go;
          sv[4 * S + j] = cin; sv[5 * S + j] = tc; sv[6 * S + j] = hin[lr][j];
        }
      }
    } else if (R.kind == TAPE_OUT) {
      const float* fh = at(pool, R.in_h);
      const float* bhs = at(pool, R.add_h);
      float* m = mods + (long long)R.base * n_out;
      float* c = cat + (long long)R.base * 2 * S;
      for (long long t = (long long)r0 * n_out + threadIdx.x; t < (long long)r1 * n_out; t += TAPE_THREADS) {
        const long long row = t / n_out;
        const int o = (int)(t - row * n_out);
        const float* w = wo + o * 2 * S;
        float acc = bo[o];
        for (int k = 0; k < S; ++k) acc = fmaf(w[k], fh[row * S + k], acc);
        if (bhs)
          for (int k = 0; k < S; ++k) acc = fmaf(w[S + k], bhs[row * S + k], acc);
        m[row * n_out + o] = tape_sigmoid(acc);
        if (o == 0) {
          float* cr = c + row * 2 * S;
          for (int k = 0; k < S; ++k) { cr[k] = fh[row * S + k]; cr[S + k] = bhs ? bhs[row * S + k] : 0.f; }
        }
      }
    } else if (R.kind == TAPE_SQUEEZE) {
      // out[owner[row]] += state[row]  (zeros(B).index_add_: the pool is zero-filled before the launch)
      const float* sh = at(pool, R.in_h);
      const float* sc = at(pool, R.in_c);
      float* oh = at(pool, R.out_h);
      float* oc = at(pool, R.out_c);
      for (long long t = (long long)r0 * S + threadIdx.x; t < (long long)r1 * S; t += TAPE_THREADS) {
        const long long row = t / S;
        const int k = (int)(t - row * S);
        const long long dst = R.owner[row];
        atomicAdd(oh + dst * S + k, sh[t]);
        atomicAdd(oc + dst * S + k, sc[t]);
      }
    } else {  // TAPE_GATE: mask ? new : old
      const float* nh = at(pool, R.in_h);
      const float* nc = at(pool, R.in_c);
      const float* oh = at(pool, R.add_h);
      const float* oc = at(pool, R.add_c);
      float* dh = at(pool, R.out_h);
      float* dc = at(pool, R.out_c);
      for (long long t = (long long)r0 * S + threadIdx.x; t < (long long)r1 * S; t += TAPE_THREADS) {
        const bool keep = R.mask[t / S] > 0.f;
        dh[t] = keep ? nh[t] : (oh ? oh[t] : 0.f);
        dc[t] = keep ? nc[t] : (oc ? oc[t] : 0.f);
      }
    }
    __syncthreads();  // range[] may be rewritten
  }
}

// ------------------------------------------------------------------------------------------------ backward
// gpool: gradients of every pool state (same offsets, zero-filled before the launch).  dpre_f / dpre_b (R x 4S) and dzo
// (R x n_out) are zero-filled by the caller; steps that are not live leave their rows zero.
__global__ void __launch_bounds__(TAPE_THREADS, 2) mod_tape_bwd_kernel(
    const TapeRec* __restrict__ recs, int n_rec, float* __restrict__ gpool, TapeNets nets,
    const float* __restrict__ saved_f, const float* __restrict__ saved_b, const float* __restrict__ mods,
    const float* __restrict__ d_mods, float* __restrict__ dpre_f, float* __restrict__ dpre_b, float* __restrict__ dzo,
    int questions) {
  extern __shared__ float smem[];
  __shared__ int range[2];
  const int S = nets.S, n_out = nets.n_out;
  float* w[2] = {smem, smem + 4 * S * S};   // W_hh row-major [4S][S]
  float* wo = smem + 8 * S * S;             // [n_out][2S]
  __shared__ float dp[TAPE_ROWS][4 * TAPE_MAXS];
  for (int net = 0; net < 2; ++net)
    for (int idx = threadIdx.x; idx < 4 * S * S; idx += TAPE_THREADS) w[net][idx] = nets.w_hh[net][idx];
  for (int idx = threadIdx.x; idx < n_out * 2 * S; idx += TAPE_THREADS) wo[idx] = nets.w_out[idx];
  __syncthreads();
  const int j = threadIdx.x % TAPE_MAXS, lr = threadIdx.x / TAPE_MAXS;
  const int q0 = (int)(((long long)questions * blockIdx.x) / gridDim.x);
  const int q1 = (int)(((long long)questions * (blockIdx.x + 1)) / gridDim.x);

  for (int r = n_rec - 1; r >= 0; --r) {
    const TapeRec R = recs[r];
    if (!R.live) continue;  // (block-uniform)
    tape_row_range(R, q0, q1, range);
    __syncthreads();
    const int r0 = range[0], r1 = range[1];
    if (R.kind == TAPE_OUT) {
      const float* m = mods + (long long)R.base * n_out;
      const float* dm = d_mods + (long long)R.base * n_out;
      float* dz = dzo + (long long)R.base * n_out;
      float* d_fh = at(gpool, R.in_h);
      float* d_bh = at(gpool, R.add_h);
      for (long long t = (long long)r0 * 2 * S + threadIdx.x; t < (long long)r1 * 2 * S; t += TAPE_THREADS) {
        const long long row = t / (2 * S);
        const int k = (int)(t - row * 2 * S);
        float acc = 0.f;
        for (int o = 0; o < n_out; ++o) {
          const float mv = m[row * n_out + o];
          const float dzv = dm[row * n_out + o] * mv * (1.0f - mv);
          if (k == 0) dz[row * n_out + o] = dzv;
          acc = fmaf(dzv, wo[o * 2 * S + k], acc);
        }
        if (k < S) d_fh[row * S + k] += acc;
        else if (d_bh) d_bh[row * S + (k - S)] += acc;
      }
    } else if (R.kind == TAPE_CELL) {
      const float* saved = (R.net == 0 ? saved_f : saved_b) + (long long)R.base * 7 * S;
      float* dpre = (R.net == 0 ? dpre_f : dpre_b) + (long long)R.base * 4 * S;
      const float* d_h_out = at(gpool, R.out_h);
      const float* d_c_out = at(gpool, R.out_c);
      float* d_h_in = at(gpool, R.in_h);
      float* d_c_in = at(gpool, R.in_c);
      float* d_h_add = at(gpool, R.add_h);
      float* d_c_add = at(gpool, R.add_c);
      float* d_fb_h = at(gpool, R.fb_h);
      float* d_fb_c = at(gpool, R.fb_c);
      const float* wn = w[R.net];
      for (int c0 = r0; c0 < r1; c0 += TAPE_ROWS) {
        const int row = c0 + lr;
        const bool act = row < r1 && j < S;
        float dcin = 0.f;
        __syncthreads();
        if (act) {
          float dh = d_h_out[(long long)row * S + j];
          float dc = d_c_out[(long long)row * S + j];
          const bool live = !(R.mask && !(R.mask[row] > 0.f));
          if (!live) {
            if (d_fb_h) { d_fb_h[(long long)row * S + j] += dh; d_fb_c[(long long)row * S + j] += dc; }
            dh = 0.f; dc = 0.f;
          }
          const float* sv = saved + (long long)row * 7 * S;
          const float gi = sv[j], gf = sv[S + j], gg = sv[2 * S + j], go = sv[3 * S + j], cin = sv[4 * S + j],
                      tc = sv[5 * S + j];
          const float dct = dc + dh * go * (1.0f - tc * tc);
          const float p_i = dct * gg * gi * (1.0f - gi);
          const float p_f = dct * cin * gf * (1.0f - gf);
          const float p_g = dct * gi * (1.0f - gg * gg);
          const float p_o = dh * tc * go * (1.0f - go);
          dcin = dct * gf;
          dp[lr][j] = p_i; dp[lr][S + j] = p_f; dp[lr][2 * S + j] = p_g; dp[lr][3 * S + j] = p_o;
          float* out = dpre + (long long)row * 4 * S;
          out[j] = p_i; out[S + j] = p_f; out[2 * S + j] = p_g; out[3 * S + j] = p_o;
        }
        __syncthreads();
        if (act) {
          float dhin = 0.f;  // thread j now owns input unit k = j
          for (int gj = 0; gj < 4 * S; ++gj) dhin = fmaf(dp[lr][gj], wn[gj * S + j], dhin);
          const long long src = R.owner ? R.owner[row] : row;
          if (R.owner) {
            if (d_h_in) { atomicAdd(d_h_in + src * S + j, dhin); atomicAdd(d_c_in + src * S + j, dcin); }
            if (d_h_add) { atomicAdd(d_h_add + src * S + j, dhin); atomicAdd(d_c_add + src * S + j, dcin); }
          } else {
            if (d_h_in) { d_h_in[src * S + j] += dhin; d_c_in[src * S + j] += dcin; }
            if (d_h_add) { d_h_add[src * S + j] += dhin; d_c_add[src * S + j] += dcin; }
          }
        }
      }
    } else if (R.kind == TAPE_SQUEEZE) {
      // state.d[row] += out.d[owner[row]]
      float* sh = at(gpool, R.in_h);
      float* sc = at(gpool, R.in_c);
      const float* oh = at(gpool, R.out_h);
      const float* oc = at(gpool, R.out_c);
      for (long long t = (long long)r0 * S + threadIdx.x; t < (long long)r1 * S; t += TAPE_THREADS) {
        const long long row = t / S;
        const int k = (int)(t - row * S);
        const long long src = R.owner[row];
        sh[t] += oh[src * S + k];
        sc[t] += oc[src * S + k];
      }
    } else {  // TAPE_GATE
      float* nh = at(gpool, R.in_h);
      float* nc = at(gpool, R.in_c);
      float* oh = at(gpool, R.add_h);
      float* oc = at(gpool, R.add_c);
      const float* dh = at(gpool, R.out_h);
      const float* dc = at(gpool, R.out_c);
      for (long long t = (long long)r0 * S + threadIdx.x; t < (long long)r1 * S; t += TAPE_THREADS) {
        const bool keep = R.mask[t / S] > 0.f;
        if (keep) { nh[t] += dh[t]; nc[t] += dc[t]; }
        else if (oh) { oh[t] += dh[t]; oc[t] += dc[t]; }
      }
    }
    __syncthreads();
  }
}

static int tape_grid(int questions) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int want = 2 * sms;  // two blocks per SM (80 KB of weights each)
  return questions < want ? questions : want;
}

}  // namespace dfol

using namespace dfol;

extern "C" int dfol_mod_tape_record_size(void) { return (int)sizeof(TapeRec); }

extern "C" int dfol_mod_tape_fwd(const void* records, int n_rec, int questions, float* pool, const float* xproj_f,
                                 const float* xproj_b, const float* w_hh_f, const float* b_hh_f, const float* w_hh_b,
                                 const float* b_hh_b, const float* w_out, const float* b_out, int S, int n_out,
                                 float* saved_f, float* saved_b, float* mods, float* cat, void* stream) {
  const char* who = "dfol_mod_tape_fwd";
  DFOL_REQUIRE(records && pool && xproj_f && xproj_b && w_hh_f && b_hh_f && w_hh_b && b_hh_b && w_out && b_out &&
                   saved_f && saved_b && mods && cat,
               "%s: null pointer", who);
  DFOL_REQUIRE(S >= 1 && S <= TAPE_MAXS && n_out >= 1 && n_out <= 16, "%s: state size 1..%d, n_out 1..16", who, TAPE_MAXS);
  if (n_rec <= 0) return 0;
  const size_t smem = (size_t)(8 * S * S + n_out * 2 * S + n_out + 8 * S) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(mod_tape_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return (int)e; }
  DFOL_REQUIRE(questions >= 1, "%s: no questions", who);
  const int blocks = tape_grid(questions);
  TapeNets nets;
  nets.w_hh[0] = w_hh_f; nets.w_hh[1] = w_hh_b; nets.b_hh[0] = b_hh_f; nets.b_hh[1] = b_hh_b;
  nets.w_out = w_out; nets.b_out = b_out; nets.S = S; nets.n_out = n_out;
  mod_tape_fwd_kernel<<<blocks, TAPE_THREADS, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const TapeRec*>(records), n_rec, pool, xproj_f, xproj_b, nets, saved_f, saved_b, mods, cat,
      questions);
  return finish_launch(who);
}

extern "C" int dfol_mod_tape_bwd(const void* records, int n_rec, int questions, float* grad_pool, const float* w_hh_f,
                                 const float* w_hh_b, const float* w_out, int S, int n_out, const float* saved_f,
                                 const float* saved_b, const float* mods, const float* d_mods, float* dpre_f,
                                 float* dpre_b, float* dzo, void* stream) {
  const char* who = "dfol_mod_tape_bwd";
  DFOL_REQUIRE(records && grad_pool && w_hh_f && w_hh_b && w_out && saved_f && saved_b && mods && d_mods && dpre_f &&
                   dpre_b && dzo,
               "%s: null pointer", who);
  DFOL_REQUIRE(S >= 1 && S <= TAPE_MAXS && n_out >= 1 && n_out <= 16, "%s: state size 1..%d, n_out 1..16", who, TAPE_MAXS);
  if (n_rec <= 0) return 0;
  const size_t smem = (size_t)(8 * S * S + n_out * 2 * S) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(mod_tape_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return (int)e; }
  DFOL_REQUIRE(questions >= 1, "%s: no questions", who);
  const int blocks = tape_grid(questions);
  TapeNets nets;
  nets.w_hh[0] = w_hh_f; nets.w_hh[1] = w_hh_b; nets.b_hh[0] = nullptr; nets.b_hh[1] = nullptr;
  nets.w_out = w_out; nets.b_out = nullptr; nets.S = S; nets.n_out = n_out;
  mod_tape_bwd_kernel<<<blocks, TAPE_THREADS, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const TapeRec*>(records), n_rec, grad_pool, nets, saved_f, saved_b, mods, d_mods, dpre_f, dpre_b,
      dzo, questions);
  return finish_launch(who);
}
