// Memory-bound pieces of the scene build and of its backward (see include/dfol_b200.h).
#include <cuda_bf16.h>

#include "dfol_common.cuh"

namespace dfol {

// ---------------------------------------------------------------------------------------------------------
// obj[t, out_col + j] = [x,y,w,h][j] / max([W,H,W,H][j], 1)
// (BatchGQABoxFeaturizer.featurize_scene, batch_gqa_boxfeatures_pipeline.py:208-211; row layout :70)
__global__ void box_position_kernel(const float* __restrict__ f, long long ldf, int D, float* __restrict__ obj,
                                    long long ldo, int out_col, long long rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 4) return;
  long long t = i >> 2;
  int j = (int)(i & 3);
  const float* r = f + t * ldf + D;  // [W, H, x, y, w, h]
  float denom = fmaxf(r[j & 1], 1.0f);
  obj[t * ldo + out_col + j] = r[2 + j] / denom;
}

// geometry features of an ordered pair (s,o): dist, asin(dy/dist), sign(x_o-x_s), sign(y_o-y_s)  (:260-279)
__device__ __forceinline__ void pair_geometry(const float* ps, const float* po, float g[4]) {
  const float x1 = ps[0], y1 = ps[1], w1 = ps[2], h1 = ps[3];
  const float x2 = po[0], y2 = po[1], w2 = po[2], h2 = po[3];
  const float dx = x1 + w1 / 2.0f - x2 - w2 / 2.0f;
  const float dy = y1 + h1 / 2.0f - y2 - h2 / 2.0f;
  const float dist = sqrtf(dx * dx + dy * dy);
  g[0] = dist;
  g[1] = asinf(dy / fmaxf(dist, 1e-10f));
  const float sx = x2 - x1, sy = y2 - y1;
  g[2] = (sx > 0.0f) ? 1.0f : (sx < 0.0f ? -1.0f : 0.0f);
  g[3] = (sy > 0.0f) ? 1.0f : (sy < 0.0f ? -1.0f : 0.0f);
}

// One warp per pair row, lanes over hidden units: h = act(U[s] + V[o] + Wg.geo + b).
__global__ void __launch_bounds__(256) pair_hidden_fwd_kernel(
    const float* __restrict__ uv, long long lduv, const float* __restrict__ pos, long long ldpos,
    const float* __restrict__ wg, long long ldw, const float* __restrict__ bias, void* __restrict__ hout,
    long long ldh, int H, int act, int out_bf16, const int32_t* __restrict__ pair_img, const int32_t* __restrict__ pair_row,
    const int32_t* __restrict__ obj_row, const int32_t* __restrict__ img_n, long long pair_rows) {
  const int lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= pair_rows) return;
  const int b = pair_img[row];
  const int n = img_n[b];
  const int l = (int)(row - pair_row[b]);
  const int s = l / n, o = l - s * n;
  const long long ts = obj_row[b] + s, to = obj_row[b] + o;
  float g[4];
  pair_geometry(pos + ts * ldpos, pos + to * ldpos, g);
  if (s == o) { g[0] = 0.f; g[1] = 0.f; g[2] = 0.f; g[3] = 0.f; }  // self pairs are not part of the path
  const float* u = uv + ts * lduv;
  const float* v = uv + to * lduv + H;
  float* out = reinterpret_cast<float*>(hout) + row * ldh;
  __nv_bfloat16* out16 = reinterpret_cast<__nv_bfloat16*>(hout) + row * ldh;
  if (out_bf16)
    for (int h = H + lane; h < ldh; h += 32) out16[h] = __float2bfloat16(0.0f);  // K padding of the next GEMM
  for (int h = lane; h < H; h += 32) {
    const float* w = wg + h * ldw;
    float z = u[h] + v[h];
    z += w[0] * g[0];
    z += w[1] * g[1];
    z += w[2] * g[2];
    z += w[3] * g[3];
    z += bias[h];
    const float a = act_apply(z, act);
    if (out_bf16) out16[h] = __float2bfloat16(a);
    else out[h] = a;
  }
}

// Backward of the pair hidden layer. One block per (image, subject s): loops over objects o, accumulates
// dU[s] in registers (block owns it), adds dV[o] / dWg / db with atomics.
__global__ void __launch_bounds__(256) pair_hidden_bwd_kernel(
    const float* __restrict__ dh, long long lddh, const float* __restrict__ hs, long long ldh,
    const float* __restrict__ pos, long long ldpos, float* __restrict__ duv, long long lduv,
    float* __restrict__ dwg, long long ldw, float* __restrict__ dbias, int H, int act,
    const int32_t* __restrict__ pair_row, const int32_t* __restrict__ obj_row, const int32_t* __restrict__ img_n,
    int max_n) {
  const int b = blockIdx.y;
  const int n = img_n[b];
  const int s = blockIdx.x;
  if (s >= n) return;
  const long long ts = obj_row[b] + s;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float du = 0.f, dw0 = 0.f, dw1 = 0.f, dw2 = 0.f, dw3 = 0.f;
    for (int o = 0; o < n; ++o) {
      if (o == s) continue;
      const long long row = (long long)pair_row[b] + (long long)s * n + o;
      const float hv = hs[row * ldh + h];
      const float dz = dh[row * lddh + h] * act_grad_from_output(hv, act);
      float g[4];
      pair_geometry(pos + ts * ldpos, pos + (obj_row[b] + o) * ldpos, g);
      du += dz;
      dw0 += dz * g[0]; dw1 += dz * g[1]; dw2 += dz * g[2]; dw3 += dz * g[3];
      atomicAdd(duv + (long long)(obj_row[b] + o) * lduv + H + h, dz);
    }
    duv[ts * lduv + h] += du;
    atomicAdd(dbias + h, du);
    atomicAdd(dwg + h * ldw + 0, dw0);
    atomicAdd(dwg + h * ldw + 1, dw1);
    atomicAdd(dwg + h * ldw + 2, dw2);
    atomicAdd(dwg + h * ldw + 3, dw3);
  }
}

// out[n] += sum_m X[m, n]: 32 column lanes x 8 row groups per block, grid-stride over row chunks.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, long long ldx, long long M, int N,
                                                     float* __restrict__ out, long long rows_per_block) {
  __shared__ float part[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float acc = 0.f;
  if (col < N)
    for (long long r = r0 + ry; r < r1; r += 8) acc += X[r * ldx + col];
  part[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][cx];
    atomicAdd(out + col, t);
  }
}

__global__ void act_grad_mul_kernel(float* __restrict__ dh, long long lddh, const float* __restrict__ h,
                                    long long ldh, long long rows, int cols, int act) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  long long r = i / cols;
  int c = (int)(i - r * cols);
  dh[r * lddh + c] *= act_grad_from_output(h[r * ldh + c], act);
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst,
                                 long long ldd, long long rows, int cols) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ldd) return;
  long long r = i / ldd;
  int c = (int)(i - r * ldd);
  dst[i] = __float2bfloat16(c < cols ? src[r * lds + c] : 0.0f);
}

// ---------------------------------------------------------------------------------------------------------
// Backward of the table layer LL = logsigmoid(H.W^T + b) for the compact gradient slices of the programs.
// One block per image: for each slice (column c of the table), dz[l] = g[l]*(1-exp(LL[l])).
//   dH[row0 + l, e] += dz[l] * W[wrow, e]      (image rows owned by this block: plain adds, thread owns e)
//   dW[wrow, e]     += sum_l dz[l] * H[row0 + l, e]   (atomic, one add per slice and e)
//   db[wrow]        += sum_l dz[l]
// Threads are laid out over e (coalesced on W, H, dH rows).
__global__ void __launch_bounds__(256) table_layer_bwd_kernel(
    const float* __restrict__ g, const int32_t* __restrict__ slice_goff, const int32_t* __restrict__ slice_col,
    const int32_t* __restrict__ slice_wrow, const int32_t* __restrict__ img_slice, const float* __restrict__ ll,
    const int64_t* __restrict__ blk, const int32_t* __restrict__ stride, const int32_t* __restrict__ row0,
    const int32_t* __restrict__ img_rows, const float* __restrict__ W, long long ldw,
    const float* __restrict__ hs, long long ldh, int E, float* __restrict__ dH, long long lddh,
    float* __restrict__ dW, float* __restrict__ db) {
  extern __shared__ float dz_s[];  // chunk of dz values
  const int b = blockIdx.x;
  const int j0 = img_slice[b], j1 = img_slice[b + 1];
  if (j0 == j1) return;
  const int rows = img_rows[b];
  const long long r0 = row0[b];
  const int st = stride[b];
  constexpr int CHUNK = 256;
  for (int j = j0; j < j1; ++j) {
    const float* gj = g + slice_goff[j];
    const float* lj = ll + blk[b] + (long long)slice_col[j] * st;
    const int wr = slice_wrow[j];
    float dbias = 0.f;
    for (int c0 = 0; c0 < rows; c0 += CHUNK) {
      const int cn = min(CHUNK, rows - c0);
      __syncthreads();
      for (int l = threadIdx.x; l < cn; l += blockDim.x) {
        const float gv = gj[c0 + l];
        dz_s[l] = (gv != 0.0f) ? gv * (1.0f - expf(lj[c0 + l])) : 0.0f;
      }
      __syncthreads();
      for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const float w = W[(long long)wr * ldw + e];
        float dw = 0.f;
        for (int l = 0; l < cn; ++l) {
          const float dz = dz_s[l];
          if (dz != 0.0f) {
            const long long r = r0 + c0 + l;
            dH[r * lddh + e] += dz * w;
            dw += dz * hs[r * ldh + e];
          }
        }
        atomicAdd(dW + (long long)wr * ldw + e, dw);
      }
      if (threadIdx.x == 0)
        for (int l = 0; l < cn; ++l) dbias += dz_s[l];
    }
    if (threadIdx.x == 0) atomicAdd(db + wr, dbias);
  }
}

}  // namespace dfol

using namespace dfol;

extern "C" int dfol_box_position(const float* features, int64_t ldf, int feature_dim, float* obj, int64_t ldo,
                                 int out_col, int64_t rows, void* stream) {
  DFOL_REQUIRE(features && obj, "dfol_box_position: null pointer");
  if (rows == 0) return 0;
  long long n = rows * 4;
  box_position_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(features, ldf, feature_dim, obj,
                                                                                     ldo, out_col, rows);
  return finish_launch("dfol_box_position");
}

extern "C" int dfol_pair_hidden_fwd(const float* uv, int64_t lduv, const float* obj_pos, int64_t ldpos,
                                    const float* wg, int64_t ldw, const float* bias, void* h_out, int64_t ldh, int H,
                                    int act, int out_bf16, const int32_t* pair_img, const int32_t* pair_row,
                                    const int32_t* obj_row, const int32_t* img_n, int64_t pair_rows, void* stream) {
  DFOL_REQUIRE(uv && obj_pos && wg && bias && h_out && pair_img && pair_row && obj_row && img_n,
               "dfol_pair_hidden_fwd: null pointer");
  if (pair_rows == 0) return 0;
  const int warps = 8;
  pair_hidden_fwd_kernel<<<(unsigned)((pair_rows + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(
      uv, lduv, obj_pos, ldpos, wg, ldw, bias, h_out, ldh, H, act, out_bf16, pair_img, pair_row, obj_row, img_n,
      pair_rows);
  return finish_launch("dfol_pair_hidden_fwd");
}

extern "C" int dfol_pair_hidden_bwd(const float* dh, int64_t lddh, const float* h_saved, int64_t ldh,
                                    const float* obj_pos, int64_t ldpos, float* duv, int64_t lduv, float* dwg,
                                    int64_t ldw, float* dbias, int H, int act, const int32_t* pair_row,
                                    const int32_t* obj_row, const int32_t* img_n, int image_num, void* stream) {
  DFOL_REQUIRE(dh && h_saved && obj_pos && duv && dwg && dbias && pair_row && obj_row && img_n,
               "dfol_pair_hidden_bwd: null pointer");
  if (image_num == 0) return 0;
  // grid.x covers the largest supported object count; blocks beyond an image's N_b exit immediately
  const int max_n = 128;
  dim3 grid(max_n, image_num);
  pair_hidden_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dh, lddh, h_saved, ldh, obj_pos, ldpos, duv, lduv,
                                                                dwg, ldw, dbias, H, act, pair_row, obj_row, img_n,
                                                                max_n);
  return finish_launch("dfol_pair_hidden_bwd");
}

extern "C" int dfol_colsum(const float* X, int64_t ldx, int64_t M, int N, float* out, void* stream) {
  DFOL_REQUIRE(X && out, "dfol_colsum: null pointer");
  if (M == 0 || N == 0) return 0;
  long long rows_per_block = 2048;
  dim3 grid((N + 31) / 32, (unsigned)((M + rows_per_block - 1) / rows_per_block));
  colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, M, N, out, rows_per_block);
  return finish_launch("dfol_colsum");
}

extern "C" int dfol_act_grad_mul(float* dH, int64_t lddh, const float* H, int64_t ldh, int64_t rows, int cols, int act,
                                 void* stream) {
  DFOL_REQUIRE(dH && H, "dfol_act_grad_mul: null pointer");
  long long n = rows * cols;
  if (n == 0) return 0;
  act_grad_mul_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dH, lddh, H, ldh, rows, cols, act);
  return finish_launch("dfol_act_grad_mul");
}

extern "C" int dfol_cast_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int cols,
                              void* stream) {
  DFOL_REQUIRE(src && dst && ldd >= cols, "dfol_cast_bf16: bad arguments");
  long long n = rows * ldd;
  if (n == 0) return 0;
  cast_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols);
  return finish_launch("dfol_cast_bf16");
}

extern "C" int dfol_table_layer_bwd(const float* g, const int32_t* slice_goff, const int32_t* slice_col,
                                    const int32_t* slice_wrow, const int32_t* img_slice, int image_num,
                                    const float* ll, const int64_t* blk, const int32_t* stride, const int32_t* row0,
                                    const int32_t* img_rows, const float* W, int64_t ldw, const float* h_saved,
                                    int64_t ldh, int E, float* dH, int64_t lddh, float* dW, float* db,
                                    void* stream) {
  DFOL_REQUIRE(g && slice_goff && slice_col && slice_wrow && img_slice && ll && blk && stride && row0 && img_rows &&
                   W && h_saved && dH && dW && db,
               "dfol_table_layer_bwd: null pointer");
  if (image_num == 0) return 0;
  table_layer_bwd_kernel<<<image_num, 256, 256 * sizeof(float), (cudaStream_t)stream>>>(
      g, slice_goff, slice_col, slice_wrow, img_slice, ll, blk, stride, row0, img_rows, W, ldw, h_saved, ldh, E, dH,
      lddh, dW, db);
  return finish_launch("dfol_table_layer_bwd");
}
