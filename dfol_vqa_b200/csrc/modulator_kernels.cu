// Attention-transfer calibrator, token side: the LSTMCell passes over the op slots and the modulation output layer
// (reference: BatchInterpreterBase.forward modulator loops, nsvqa/nn/interpreter/batch_base_interpreter.py:87-140;
// FilterBatch / RelateBatch.transform_attention, batch_base_ops.py:407-467, :598-684; _compute_attention_modulations
// :275-286; the networks are nn.LSTMCell(318, 50) x 2 and Linear(100, 4) + Sigmoid, gqa_interpreter_experiments.py:
// 119-131).  States are (rows x S) fp32 with S <= 64; the input projections features . W_ih^T of ALL cells of a batch
// are one GEMM (dfol_gemm_f32) done by the caller, so a cell here is the recurrent part only:
//     pre = xproj[row] + b_hh + W_hh . (h_in[src] (+ h_add[src]))        gate order i, f, g, o (torch.nn.LSTMCell)
//     c' = sigmoid(f) * c_in + sigmoid(i) * tanh(g),   h' = sigmoid(o) * tanh(c')
// with an optional row map src = owner[row] (expand of a question's state to its option predicates) and an optional
// 0/1 row mask (rows with mask 0 pass the fallback state through: the interpreter's gate of unaffected questions).
// Backward kernels write d pre of every cell row and keep h_in, so that dW_ih, dW_hh and the bias gradients are three
// GEMM-shaped reductions over all cell rows at the end (caller).
#include "dfol_common.cuh"

namespace dfol {

constexpr int LSTM_MAXS = 64;
constexpr int LSTM_ROWS = 8;  // rows per block

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// saved[row] = [i, f, g, o (post-activation, 4S) | c_in (S) | tanh(c') (S) | h_in total (S)]  (7S floats)
__global__ void __launch_bounds__(LSTM_MAXS* LSTM_ROWS) lstm_cell_fwd_kernel(
    const float* __restrict__ xproj, long long ldx, const float* __restrict__ b_hh, const float* __restrict__ w_hh,
    int S, const float* __restrict__ h_in, const float* __restrict__ c_in, const float* __restrict__ h_add,
    const float* __restrict__ c_add, const int64_t* __restrict__ owner, const float* __restrict__ mask,
    const float* __restrict__ fb_h, const float* __restrict__ fb_c, float* __restrict__ h_out, float* __restrict__ c_out,
    float* __restrict__ saved, int rows) {
  extern __shared__ float wt[];  // W_hh transposed: wt[k * 4S + g * S + j]
  __shared__ float hin[LSTM_ROWS][LSTM_MAXS];
  const int j = threadIdx.x % LSTM_MAXS, lr = threadIdx.x / LSTM_MAXS;
  for (int idx = threadIdx.x; idx < 4 * S * S; idx += blockDim.x) {
    const int gj = idx / S, k = idx - gj * S;
    wt[k * 4 * S + gj] = w_hh[idx];
  }
  const int row = blockIdx.x * LSTM_ROWS + lr;
  const bool act = row < rows && j < S;
  long long src = 0;
  float cin = 0.f;
  if (act) {
    src = owner ? owner[row] : row;
    float h = h_in ? h_in[src * S + j] : 0.f;
    cin = c_in ? c_in[src * S + j] : 0.f;
    if (h_add) { h += h_add[src * S + j]; cin += c_add[src * S + j]; }
    hin[lr][j] = h;
  }
  __syncthreads();
  if (!act) return;
  float pre[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) pre[g] = xproj[row * ldx + g * S + j] + b_hh[g * S + j];
  for (int k = 0; k < S; ++k) {
    const float h = hin[lr][k];
    const float* w = wt + k * 4 * S + j;
#pragma unroll
    for (int g = 0; g < 4; ++g) pre[g] = fmaf(h, w[g * S], pre[g]);
  }
  const float gi = sigmoidf_(pre[0]), gf = sigmoidf_(pre[1]), gg = tanhf(pre[2]), go = sigmoidf_(pre[3]);
  const float cn = gf * cin + gi * gg;
  const float tc = tanhf(cn);
  float ho = go * tc, co = cn;
  if (mask && !(mask[row] > 0.f)) { ho = fb_h[row * S + j]; co = fb_c[row * S + j]; }
  h_out[row * S + j] = ho;
  c_out[row * S + j] = co;
  float* sv = saved + (long long)row * 7 * S;
  sv[j] = gi; sv[S + j] = gf; sv[2 * S + j] = gg; sv[3 * S + j] = go;
  sv[4 * S + j] = cin; sv[5 * S + j] = tc; sv[6 * S + j] = hin[lr][j];
}

// d_h_out / d_c_out (rows x S, may be null = zero) -> dpre (rows x 4S, written), d_h_in / d_c_in (+=, at src rows, atomics
// when a row map is present), the same into d_h_add / d_c_add, and d_fb (+=) for rows whose mask is 0.
__global__ void __launch_bounds__(LSTM_MAXS* LSTM_ROWS) lstm_cell_bwd_kernel(
    const float* __restrict__ d_h_out, const float* __restrict__ d_c_out, const float* __restrict__ w_hh, int S,
    const float* __restrict__ saved, const int64_t* __restrict__ owner, const float* __restrict__ mask,
    float* __restrict__ dpre, long long lddp, float* __restrict__ d_h_in, float* __restrict__ d_c_in,
    float* __restrict__ d_h_add, float* __restrict__ d_c_add, float* __restrict__ d_fb_h, float* __restrict__ d_fb_c,
    int rows) {
  extern __shared__ float w[];  // W_hh row-major [4S][S]
  __shared__ float dp[LSTM_ROWS][4 * LSTM_MAXS];
  const int j = threadIdx.x % LSTM_MAXS, lr = threadIdx.x / LSTM_MAXS;
  for (int idx = threadIdx.x; idx < 4 * S * S; idx += blockDim.x) w[idx] = w_hh[idx];
  const int row = blockIdx.x * LSTM_ROWS + lr;
  const bool act = row < rows && j < S;
  bool live = false;
  float dcin = 0.f;
  if (act) {
    float dh = d_h_out ? d_h_out[row * S + j] : 0.f;
    float dc = d_c_out ? d_c_out[row * S + j] : 0.f;
    live = !(mask && !(mask[row] > 0.f));
    if (!live) {
      if (d_fb_h) { d_fb_h[row * S + j] += dh; d_fb_c[row * S + j] += dc; }
      dh = 0.f; dc = 0.f;
    }
    const float* sv = saved + (long long)row * 7 * S;
    const float gi = sv[j], gf = sv[S + j], gg = sv[2 * S + j], go = sv[3 * S + j], cin = sv[4 * S + j], tc = sv[5 * S + j];
    const float dct = dc + dh * go * (1.0f - tc * tc);
    const float p_i = dct * gg * gi * (1.0f - gi);
    const float p_f = dct * cin * gf * (1.0f - gf);
    const float p_g = dct * gi * (1.0f - gg * gg);
    const float p_o = dh * tc * go * (1.0f - go);
    dcin = dct * gf;
    dp[lr][j] = p_i; dp[lr][S + j] = p_f; dp[lr][2 * S + j] = p_g; dp[lr][3 * S + j] = p_o;
    float* out = dpre + row * lddp;
    out[j] = p_i; out[S + j] = p_f; out[2 * S + j] = p_g; out[3 * S + j] = p_o;
  }
  __syncthreads();
  if (!act) return;
  float dhin = 0.f;  // thread j now owns input unit k = j
  for (int gj = 0; gj < 4 * S; ++gj) dhin = fmaf(dp[lr][gj], w[gj * S + j], dhin);
  const long long src = owner ? owner[row] : row;
  if (owner) {
    if (d_h_in) { atomicAdd(d_h_in + src * S + j, dhin); atomicAdd(d_c_in + src * S + j, dcin); }
    if (d_h_add) { atomicAdd(d_h_add + src * S + j, dhin); atomicAdd(d_c_add + src * S + j, dcin); }
  } else {
    if (d_h_in) { d_h_in[src * S + j] += dhin; d_c_in[src * S + j] += dcin; }
    if (d_h_add) { d_h_add[src * S + j] += dhin; d_c_add[src * S + j] += dcin; }
  }
}

// mods[row] = sigmoid(W_out . [fh[row] | bh[src(row)]] + b_out)   (W_out: [n_out][2S]); cat[row] keeps the input
__global__ void __launch_bounds__(256) mod_out_fwd_kernel(const float* __restrict__ fh, const float* __restrict__ bh,
                                                          const int64_t* __restrict__ owner,
                                                          const float* __restrict__ w_out,
                                                          const float* __restrict__ b_out, int S, int n_out,
                                                          float* __restrict__ mods, float* __restrict__ cat, int rows) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = t / n_out, o = t - row * n_out;
  if (row >= rows) return;
  const long long src = owner ? owner[row] : row;
  const float* w = w_out + (long long)o * 2 * S;
  float acc = b_out[o];
  for (int k = 0; k < S; ++k) acc = fmaf(w[k], fh[(long long)row * S + k], acc);
  for (int k = 0; k < S; ++k) acc = fmaf(w[S + k], bh ? bh[src * S + k] : 0.f, acc);
  mods[(long long)row * n_out + o] = sigmoidf_(acc);
  if (o == 0) {
    float* c = cat + (long long)row * 2 * S;
    for (int k = 0; k < S; ++k) { c[k] = fh[(long long)row * S + k]; c[S + k] = bh ? bh[src * S + k] : 0.f; }
  }
}

// d_mods -> dzo[row] = d_mods * m (1 - m) (written), d_fh[row] += W_out[:, :S]^T dzo, d_bh[src] += W_out[:, S:]^T dzo
__global__ void __launch_bounds__(256) mod_out_bwd_kernel(const float* __restrict__ d_mods,
                                                          const float* __restrict__ mods,
                                                          const int64_t* __restrict__ owner,
                                                          const float* __restrict__ w_out, int S, int n_out,
                                                          float* __restrict__ dzo, float* __restrict__ d_fh,
                                                          float* __restrict__ d_bh, int rows) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = t / (2 * S), k = t - row * (2 * S);
  if (row >= rows) return;
  float acc = 0.f;
  for (int o = 0; o < n_out; ++o) {
    const float m = mods[(long long)row * n_out + o];
    const float dz = d_mods[(long long)row * n_out + o] * m * (1.0f - m);
    if (k == 0) dzo[(long long)row * n_out + o] = dz;
    acc = fmaf(dz, w_out[(long long)o * 2 * S + k], acc);
  }
  if (k < S) {
    d_fh[(long long)row * S + k] += acc;
  } else if (d_bh) {
    const long long src = owner ? owner[row] : row;
    if (owner) atomicAdd(d_bh + src * S + (k - S), acc);
    else d_bh[src * S + (k - S)] += acc;
  }
}

}  // namespace dfol

using namespace dfol;

extern "C" int dfol_lstm_cell_fwd(const float* xproj, int64_t ldx, const float* b_hh, const float* w_hh, int S,
                                  const float* h_in, const float* c_in, const float* h_add, const float* c_add,
                                  const int64_t* owner, const float* mask, const float* fb_h, const float* fb_c,
                                  float* h_out, float* c_out, float* saved, int rows, void* stream) {
  DFOL_REQUIRE(xproj && b_hh && w_hh && h_out && c_out && saved, "dfol_lstm_cell_fwd: null pointer");
  DFOL_REQUIRE(S >= 1 && S <= LSTM_MAXS, "dfol_lstm_cell_fwd: state size must be 1..%d", LSTM_MAXS);
  DFOL_REQUIRE((h_in == nullptr) == (c_in == nullptr) && (h_add == nullptr) == (c_add == nullptr),
               "dfol_lstm_cell_fwd: h and c go together");
  DFOL_REQUIRE(mask == nullptr || (fb_h && fb_c), "dfol_lstm_cell_fwd: a row mask needs the fallback state");
  if (rows <= 0) return 0;
  const size_t smem = (size_t)4 * S * S * sizeof(float);
  cudaFuncSetAttribute(lstm_cell_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lstm_cell_fwd_kernel<<<(rows + LSTM_ROWS - 1) / LSTM_ROWS, LSTM_MAXS * LSTM_ROWS, smem, (cudaStream_t)stream>>>(
      xproj, ldx, b_hh, w_hh, S, h_in, c_in, h_add, c_add, owner, mask, fb_h, fb_c, h_out, c_out, saved, rows);
  return finish_launch("dfol_lstm_cell_fwd");
}

extern "C" int dfol_lstm_cell_bwd(const float* d_h_out, const float* d_c_out, const float* w_hh, int S,
                                  const float* saved, const int64_t* owner, const float* mask, float* dpre,
                                  int64_t lddp, float* d_h_in, float* d_c_in, float* d_h_add, float* d_c_add,
                                  float* d_fb_h, float* d_fb_c, int rows, void* stream) {
  DFOL_REQUIRE(w_hh && saved && dpre, "dfol_lstm_cell_bwd: null pointer");
  DFOL_REQUIRE(S >= 1 && S <= LSTM_MAXS, "dfol_lstm_cell_bwd: state size must be 1..%d", LSTM_MAXS);
  DFOL_REQUIRE((d_h_in == nullptr) == (d_c_in == nullptr) && (d_h_add == nullptr) == (d_c_add == nullptr) &&
                   (d_fb_h == nullptr) == (d_fb_c == nullptr),
               "dfol_lstm_cell_bwd: h and c gradients go together");
  if (rows <= 0) return 0;
  const size_t smem = (size_t)4 * S * S * sizeof(float);
  cudaFuncSetAttribute(lstm_cell_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lstm_cell_bwd_kernel<<<(rows + LSTM_ROWS - 1) / LSTM_ROWS, LSTM_MAXS * LSTM_ROWS, smem, (cudaStream_t)stream>>>(
      d_h_out, d_c_out, w_hh, S, saved, owner, mask, dpre, lddp, d_h_in, d_c_in, d_h_add, d_c_add, d_fb_h, d_fb_c,
      rows);
  return finish_launch("dfol_lstm_cell_bwd");
}

extern "C" int dfol_mod_out_fwd(const float* fh, const float* bh, const int64_t* owner, const float* w_out,
                                const float* b_out, int S, int n_out, float* mods, float* cat, int rows,
                                void* stream) {
  DFOL_REQUIRE(fh && w_out && b_out && mods && cat, "dfol_mod_out_fwd: null pointer");
  DFOL_REQUIRE(S >= 1 && n_out >= 1, "dfol_mod_out_fwd: bad sizes");
  if (rows <= 0) return 0;
  const long long threads = (long long)rows * n_out;
  mod_out_fwd_kernel<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fh, bh, owner, w_out, b_out, S,
                                                                                      n_out, mods, cat, rows);
  return finish_launch("dfol_mod_out_fwd");
}

extern "C" int dfol_mod_out_bwd(const float* d_mods, const float* mods, const int64_t* owner, const float* w_out,
                                int S, int n_out, float* dzo, float* d_fh, float* d_bh, int rows, void* stream) {
  DFOL_REQUIRE(d_mods && mods && w_out && dzo && d_fh, "dfol_mod_out_bwd: null pointer");
  if (rows <= 0) return 0;
  const long long threads = (long long)rows * 2 * S;
  mod_out_bwd_kernel<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_mods, mods, owner, w_out, S,
                                                                                      n_out, dzo, d_fh, d_bh, rows);
  return finish_launch("dfol_mod_out_bwd");
}
