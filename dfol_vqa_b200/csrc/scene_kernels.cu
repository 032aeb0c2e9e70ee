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

// Pair hidden layer h[(s,o), :] = act(U[s] + V[o] + Wg.geo(s,o) + b), tiled so that every U/V row is fetched from
// L2 once per 8x32 pair tile instead of once per pair: one block per (image, 8 subjects, 32 objects); warp w owns
// subject s0+w and keeps its U row slice, the geometry weights and the bias of its hidden units in registers; the
// 32 V rows and the tile's pair geometry sit in shared memory. Lanes own hidden units 4*lane + 128*i (H <= 512).
// FAST selects the approximate intrinsics of the bf16 tensor-core mode; fp32 parity mode uses accurate functions.
constexpr int PH_TS = 8, PH_TO = 32, PH_MAXG = 4;  // H <= 128 * PH_MAXG

template <bool FAST, bool BF16OUT>
__global__ void __launch_bounds__(256) pair_hidden_fwd_kernel(
    const float* __restrict__ uv, long long lduv, const float* __restrict__ pos, long long ldpos,
    const float* __restrict__ wg, long long ldw, const float* __restrict__ bias, void* __restrict__ hout,
    long long ldh, int H, int act, const int32_t* __restrict__ pair_row, const int32_t* __restrict__ obj_row,
    const int32_t* __restrict__ img_n, int tiles_o) {
  extern __shared__ float4 vsm[];                       // [PH_TO][H/4] V rows
  float4* geo = vsm + PH_TO * (H / 4);                  // [PH_TS][PH_TO] pair geometry
  const int b = blockIdx.y;
  const int n = img_n[b];
  const int s0 = (blockIdx.x / tiles_o) * PH_TS, o0 = (blockIdx.x % tiles_o) * PH_TO;
  if (s0 >= n || o0 >= n) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long t0 = obj_row[b];
  const int H4 = H / 4;
  const int no = min(PH_TO, n - o0);
  for (int idx = threadIdx.x; idx < no * H4; idx += blockDim.x) {
    const int o = idx / H4, q = idx - o * H4;
    vsm[o * H4 + q] = __ldg(reinterpret_cast<const float4*>(uv + (t0 + o0 + o) * lduv + H) + q);
  }
  {
    const int si = threadIdx.x / PH_TO, oi = threadIdx.x % PH_TO;  // 256 threads = 8 x 32 pairs
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (s0 + si < n && oi < no && s0 + si != o0 + oi)
      pair_geometry(pos + (t0 + s0 + si) * ldpos, pos + (t0 + o0 + oi) * ldpos, g);
    geo[si * PH_TO + oi] = make_float4(g[0], g[1], g[2], g[3]);
  }
  __syncthreads();
  const int s = s0 + warp;
  if (s >= n) return;
  // per-lane constants: U[s], geometry weights and bias of hidden units 4*(lane + 32 i) .. +3
  float4 u[PH_MAXG], bz[PH_MAXG], w0[PH_MAXG], w1[PH_MAXG], w2[PH_MAXG], w3[PH_MAXG];
#pragma unroll
  for (int i = 0; i < PH_MAXG; ++i) {
    const int q = lane + 32 * i;
    if (q < H4) {
      u[i] = __ldg(reinterpret_cast<const float4*>(uv + (t0 + s) * lduv) + q);
      bz[i] = make_float4(bias[4 * q], bias[4 * q + 1], bias[4 * q + 2], bias[4 * q + 3]);
      const float* w = wg + (long long)(4 * q) * ldw;
      w0[i] = make_float4(w[0], w[1], w[2], w[3]);
      w1[i] = make_float4(w[ldw], w[ldw + 1], w[ldw + 2], w[ldw + 3]);
      w2[i] = make_float4(w[2 * ldw], w[2 * ldw + 1], w[2 * ldw + 2], w[2 * ldw + 3]);
      w3[i] = make_float4(w[3 * ldw], w[3 * ldw + 1], w[3 * ldw + 2], w[3 * ldw + 3]);
    }
  }
  auto activate = [&](float t) -> float {
    if (FAST) {
      return (act == DFOL_ACT_ELU) ? (t > 0.0f ? t : __expf(t) - 1.0f)
                                   : (act == DFOL_ACT_SIGMOID ? __fdividef(1.0f, 1.0f + __expf(-t)) : t);
    }
    return act_apply(t, act);
  };
  for (int oi = 0; oi < no; ++oi) {
    const float4 g = geo[warp * PH_TO + oi];
    const long long row = (long long)pair_row[b] + (long long)s * n + o0 + oi;
#pragma unroll
    for (int i = 0; i < PH_MAXG; ++i) {
      const int q = lane + 32 * i;
      if (q < H4) {
        const float4 v = vsm[oi * H4 + q];
        // same association as the unfused sum: ((((u + v) + w0 g0) + w1 g1) + w2 g2) + w3 g3) + b
        float a0 = u[i].x + v.x, a1 = u[i].y + v.y, a2 = u[i].z + v.z, a3 = u[i].w + v.w;
        a0 += w0[i].x * g.x; a0 += w0[i].y * g.y; a0 += w0[i].z * g.z; a0 += w0[i].w * g.w; a0 += bz[i].x;
        a1 += w1[i].x * g.x; a1 += w1[i].y * g.y; a1 += w1[i].z * g.z; a1 += w1[i].w * g.w; a1 += bz[i].y;
        a2 += w2[i].x * g.x; a2 += w2[i].y * g.y; a2 += w2[i].z * g.z; a2 += w2[i].w * g.w; a2 += bz[i].z;
        a3 += w3[i].x * g.x; a3 += w3[i].y * g.y; a3 += w3[i].z * g.z; a3 += w3[i].w * g.w; a3 += bz[i].w;
        a0 = activate(a0); a1 = activate(a1); a2 = activate(a2); a3 = activate(a3);
        if (BF16OUT) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(a0, a1), hi = __floats2bfloat162_rn(a2, a3);
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(hout) + row * ldh + 4 * q) =
              make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        } else {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(hout) + row * ldh + 4 * q) = make_float4(a0, a1, a2, a3);
        }
      }
    }
    if (BF16OUT)  // K padding of the next tensor-core GEMM
      for (int h = H + lane; h < ldh; h += 32)
        reinterpret_cast<__nv_bfloat16*>(hout)[row * ldh + h] = __float2bfloat16(0.0f);
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

// Same cast, 8 elements per thread (four 8-byte loads, one 16-byte store): ldd a multiple of 8, lds even, src 8-byte
// and dst 16-byte aligned.  Grid-stride over (row, 8-column group).
__global__ void __launch_bounds__(256) cast_bf16_vec8_kernel(const float* __restrict__ src, long long lds,
                                                             __nv_bfloat16* __restrict__ dst, long long ldd,
                                                             long long rows, int cols) {
  const int groups = (int)(ldd >> 3);
  const long long total = rows * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int c = (int)(i - r * groups) << 3;
    float v[8];
    if (c + 8 <= cols) {  // 8-byte loads: box-feature rows are (D + 6) floats long, i.e. only 8-byte aligned
      const float2* q = reinterpret_cast<const float2*>(src + r * lds + c);
      const float2 a = __ldg(q), b = __ldg(q + 1), d = __ldg(q + 2), e = __ldg(q + 3);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = d.x; v[5] = d.y; v[6] = e.x; v[7] = e.y;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c + j < cols) ? src[r * lds + c + j] : 0.0f;
    }
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      pk[j] = *reinterpret_cast<uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(dst + r * ldd + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
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


// Fused backward of the table layer for images with at most SMAX touched table columns (the relation table:
// a program touches a handful of relations). One block per (64-row chunk, image):
//   dz_j[l] = g_j[l] * (1 - exp(LL_j[l]))                                   (logsigmoid')
//   dZ[row, e] = (sum_j dz_j[l] W[wrow_j, e]) * act'(H[row, e])             written once, no read-modify-write
//   dW[wrow_j, e] += sum_l dz_j[l] H[row, e];  db[wrow_j] += sum_l dz_j[l]  (atomics, one per block and e)
// Rows of images without slices are written as zero, so dZ needs no memset.
__device__ __forceinline__ float ld_as_float(const float* p) { return *p; }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_from_float(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

template <int SMAX, typename HT>
__global__ void __launch_bounds__(256) table_layer_bwd_fused_kernel(
    const float* __restrict__ g, const int32_t* __restrict__ slice_goff, const int32_t* __restrict__ slice_col,
    const int32_t* __restrict__ slice_wrow, const int32_t* __restrict__ img_slice, const float* __restrict__ ll,
    const int64_t* __restrict__ blk, const int32_t* __restrict__ stride, const int32_t* __restrict__ row0,
    const int32_t* __restrict__ img_rows, const float* __restrict__ W, long long ldw,
    const HT* __restrict__ hs, long long ldh, int E, int act, HT* __restrict__ dZ, long long lddz, int out_cols,
    float* __restrict__ dW, float* __restrict__ db) {
  constexpr int R = 64;
  __shared__ float dz_s[SMAX][R];
  __shared__ int wrow_s[SMAX];
  const int b = blockIdx.y;
  const int rows = img_rows[b];
  const int c = blockIdx.x * R;
  if (c >= rows) return;
  const int cn = min(R, rows - c);
  const long long r0 = (long long)row0[b] + c;
  const int j0 = img_slice[b];
  const int S = min(img_slice[b + 1] - j0, SMAX);
  // columns E .. out_cols are the zero K-padding of the next (tensor-core) GEMM
  for (int idx = threadIdx.x; idx < cn * (out_cols - E); idx += blockDim.x) {
    const int l = idx / (out_cols - E), e = E + idx - l * (out_cols - E);
    st_from_float(dZ + (r0 + l) * lddz + e, 0.0f);
  }
  if (S == 0) {
    for (int idx = threadIdx.x; idx < cn * E; idx += blockDim.x) {
      const int l = idx / E, e = idx - l * E;
      st_from_float(dZ + (r0 + l) * lddz + e, 0.0f);
    }
    return;
  }
  const int st = stride[b];
  for (int idx = threadIdx.x; idx < SMAX * R; idx += blockDim.x) {  // unused slice rows must be zero (0 * garbage)
    const int j = idx / R, l = idx - j * R;
    float v = 0.0f;
    if (j < S && l < cn) {
      const float gv = g[slice_goff[j0 + j] + c + l];
      if (gv != 0.0f) v = gv * (1.0f - expf(ll[blk[b] + (long long)slice_col[j0 + j] * st + c + l]));
    }
    dz_s[j][l] = v;
  }
  if (threadIdx.x < S) wrow_s[threadIdx.x] = slice_wrow[j0 + threadIdx.x];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < S) {
    float t = dz_s[warp][lane] + dz_s[warp][lane + 32];
    t = warp_sum(t);
    if (lane == 0 && t != 0.0f) atomicAdd(db + wrow_s[warp], t);
  }
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float wj[SMAX], dwj[SMAX];
#pragma unroll
    for (int j = 0; j < SMAX; ++j) {
      wj[j] = (j < S) ? W[(long long)wrow_s[j] * ldw + e] : 0.0f;
      dwj[j] = 0.0f;
    }
    for (int l = 0; l < cn; ++l) {
      const float h = ld_as_float(hs + (r0 + l) * ldh + e);
      float out = 0.0f;
#pragma unroll
      for (int j = 0; j < SMAX; ++j) {
        const float dz = dz_s[j][l];
        out += dz * wj[j];
        dwj[j] += dz * h;
      }
      st_from_float(dZ + (r0 + l) * lddz + e, out * act_grad_from_output(h, act));
    }
#pragma unroll
    for (int j = 0; j < SMAX; ++j)
      if (j < S && dwj[j] != 0.0f) atomicAdd(dW + (long long)wrow_s[j] * ldw + e, dwj[j]);
  }
}

// Scatter the compact gradient slices into a dense (rows x columns) matrix with logsigmoid' applied:
// dZ[row0[b] + l, col] += g[l] * (1 - exp(LL[b][col][l])). One warp per slice.
__global__ void __launch_bounds__(256) table_grad_dense_kernel(
    const float* __restrict__ g, const int32_t* __restrict__ slice_goff, const int32_t* __restrict__ slice_col,
    const int32_t* __restrict__ slice_img, int slice_num, const float* __restrict__ ll,
    const int64_t* __restrict__ blk, const int32_t* __restrict__ stride, const int32_t* __restrict__ row0,
    const int32_t* __restrict__ img_rows, float* __restrict__ dZ, long long lddz) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= slice_num) return;
  const int b = slice_img[j], col = slice_col[j];
  const int rows = img_rows[b];
  const float* gj = g + slice_goff[j];
  const float* lj = ll + blk[b] + (long long)col * stride[b];
  for (int l = threadIdx.x & 31; l < rows; l += 32) {
    const float gv = gj[l];
    if (gv != 0.0f) atomicAdd(dZ + ((long long)row0[b] + l) * lddz + col, gv * (1.0f - expf(lj[l])));
  }
}


// Backward of the pair hidden layer from bf16 dZ (activation derivative already applied by the dgrad epilogue).
// One block per image, thread h owns hidden unit h: streams the image's N^2 rows (coalesced over h), keeps dU[s]
// in a register, dV[o][h] in shared memory (thread-private column: no conflicts), dWg / db in registers.
__global__ void __launch_bounds__(256) pair_hidden_bwd_bf16_kernel(
    const __nv_bfloat16* __restrict__ dz, long long lddz, const float* __restrict__ pos, long long ldpos,
    float* __restrict__ duv, long long lduv, float* __restrict__ dwg, long long ldw, float* __restrict__ dbias, int H,
    const int32_t* __restrict__ pair_row, const int32_t* __restrict__ obj_row, const int32_t* __restrict__ img_n) {
  extern __shared__ float dv_s[];  // [n][H]
  const int b = blockIdx.x;
  const int n = img_n[b];
  const long long t0 = obj_row[b];
  const long long p0 = pair_row[b];
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    for (int o = 0; o < n; ++o) dv_s[o * H + h] = 0.0f;
    float dw0 = 0.f, dw1 = 0.f, dw2 = 0.f, dw3 = 0.f, dbh = 0.f;
    for (int s = 0; s < n; ++s) {
      float du = 0.f;
      const __nv_bfloat16* row = dz + (p0 + (long long)s * n) * lddz + h;
      const float* ps = pos + (t0 + s) * ldpos;
      int o = 0;
      for (; o + 4 <= n; o += 4) {  // four independent loads in flight
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __bfloat162float(row[(long long)(o + i) * lddz]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (o + i == s) continue;
          float g[4];
          pair_geometry(ps, pos + (t0 + o + i) * ldpos, g);
          du += v[i];
          dv_s[(o + i) * H + h] += v[i];
          dw0 += v[i] * g[0]; dw1 += v[i] * g[1]; dw2 += v[i] * g[2]; dw3 += v[i] * g[3];
        }
      }
      for (; o < n; ++o) {
        if (o == s) continue;
        const float v = __bfloat162float(row[(long long)o * lddz]);
        float g[4];
        pair_geometry(ps, pos + (t0 + o) * ldpos, g);
        du += v;
        dv_s[o * H + h] += v;
        dw0 += v * g[0]; dw1 += v * g[1]; dw2 += v * g[2]; dw3 += v * g[3];
      }
      duv[(t0 + s) * lduv + h] = du;
      dbh += du;
    }
    for (int o = 0; o < n; ++o) duv[(t0 + o) * lduv + H + h] = dv_s[o * H + h];
    atomicAdd(dbias + h, dbh);
    atomicAdd(dwg + h * ldw + 0, dw0);
    atomicAdd(dwg + h * ldw + 1, dw1);
    atomicAdd(dwg + h * ldw + 2, dw2);
    atomicAdd(dwg + h * ldw + 3, dw3);
  }
}

// out[n] += sum_m X[m, n] for bf16 X: block = 64 columns (bf16x2 per lane) x 8 row groups, rows_per_block rows;
// four rows in flight per thread.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, long long ldx,
                                                          long long M, int N, float* __restrict__ out,
                                                          long long rows_per_block) {
  __shared__ float part[8][64];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 64 + 2 * cx;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float ax = 0.f, ay = 0.f;
  if (col + 1 < N || (col < N && (N & 1) == 0)) {
    const __nv_bfloat16* p = X + (r0 + ry) * ldx + col;
    const long long step = 8 * ldx;
    long long r = r0 + ry;
    for (; r + 24 < r1; r += 32, p += 4 * step) {
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + step));
      const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + 2 * step));
      const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + 3 * step));
      ax += (a.x + b.x) + (c.x + d.x);
      ay += (a.y + b.y) + (c.y + d.y);
    }
    for (; r < r1; r += 8, p += step) {
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
      ax += a.x;
      ay += a.y;
    }
  } else if (col < N) {  // odd trailing column
    for (long long r = r0 + ry; r < r1; r += 8) ax += __bfloat162float(X[r * ldx + col]);
  }
  part[ry][2 * cx] = ax;
  part[ry][2 * cx + 1] = ay;
  __syncthreads();
  if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    atomicAdd(out + blockIdx.x * 64 + threadIdx.x, t);
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
                                    int act, int out_bf16, const int32_t* pair_row, const int32_t* obj_row,
                                    const int32_t* img_n, int image_num, int max_n, void* stream) {
  DFOL_REQUIRE(uv && obj_pos && wg && bias && h_out && pair_row && obj_row && img_n,
               "dfol_pair_hidden_fwd: null pointer");
  if (image_num == 0) return 0;
  DFOL_REQUIRE((H % 4) == 0 && (lduv % 4) == 0 && (ldh % 4) == 0 && (reinterpret_cast<uintptr_t>(uv) % 16) == 0 &&
                   (reinterpret_cast<uintptr_t>(h_out) % 16) == 0,
               "dfol_pair_hidden_fwd: H, lduv, ldh must be multiples of 4 and buffers 16-byte aligned");
  DFOL_REQUIRE(H <= 128 * PH_MAXG, "dfol_pair_hidden_fwd: hidden width above %d is not supported", 128 * PH_MAXG);
  DFOL_REQUIRE(max_n >= 1 && image_num >= 1, "dfol_pair_hidden_fwd: empty batch");
  const int tiles_s = (max_n + PH_TS - 1) / PH_TS, tiles_o = (max_n + PH_TO - 1) / PH_TO;
  dim3 grid(tiles_s * tiles_o, image_num);
  const size_t smem = (size_t)(PH_TO * (H / 4) + PH_TS * PH_TO) * sizeof(float4);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_bf16) {
    auto kern = pair_hidden_fwd_kernel<true, true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, 256, smem, st>>>(uv, lduv, obj_pos, ldpos, wg, ldw, bias, h_out, ldh, H, act, pair_row, obj_row,
                                  img_n, tiles_o);
  } else {
    auto kern = pair_hidden_fwd_kernel<false, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, 256, smem, st>>>(uv, lduv, obj_pos, ldpos, wg, ldw, bias, h_out, ldh, H, act, pair_row, obj_row,
                                  img_n, tiles_o);
  }
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
  const bool vec = (ldd % 8) == 0 && (lds % 2) == 0 && (reinterpret_cast<uintptr_t>(src) % 8) == 0 &&
                   (reinterpret_cast<uintptr_t>(dst) % 16) == 0;
  if (vec) {
    long long blocks = (n / 8 + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    cast_bf16_vec8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols);
  } else {
    cast_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols);
  }
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

extern "C" int dfol_table_layer_bwd_fused(const float* g, const int32_t* slice_goff, const int32_t* slice_col,
                                          const int32_t* slice_wrow, const int32_t* img_slice, int image_num,
                                          int max_rows, const float* ll, const int64_t* blk, const int32_t* stride,
                                          const int32_t* row0, const int32_t* img_rows, const float* W, int64_t ldw,
                                          const void* h_saved, int64_t ldh, int E, int act, void* dZ, int64_t lddz,
                                          int out_cols, int bf16_io, float* dW, float* db, void* stream) {
  DFOL_REQUIRE(g && slice_goff && slice_col && slice_wrow && img_slice && ll && blk && stride && row0 && img_rows &&
                   W && h_saved && dZ && dW && db,
               "dfol_table_layer_bwd_fused: null pointer");
  if (image_num == 0 || max_rows == 0) return 0;
  dim3 grid((max_rows + 63) / 64, image_num);
  DFOL_REQUIRE(grid.y <= 65535, "dfol_table_layer_bwd_fused: too many images");
  DFOL_REQUIRE(out_cols >= E && out_cols <= lddz, "dfol_table_layer_bwd_fused: E <= out_cols <= lddz");
  if (bf16_io)
    table_layer_bwd_fused_kernel<8, __nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        g, slice_goff, slice_col, slice_wrow, img_slice, ll, blk, stride, row0, img_rows, W, ldw,
        reinterpret_cast<const __nv_bfloat16*>(h_saved), ldh, E, act, reinterpret_cast<__nv_bfloat16*>(dZ), lddz,
        out_cols, dW, db);
  else
    table_layer_bwd_fused_kernel<8, float><<<grid, 256, 0, (cudaStream_t)stream>>>(
        g, slice_goff, slice_col, slice_wrow, img_slice, ll, blk, stride, row0, img_rows, W, ldw,
        reinterpret_cast<const float*>(h_saved), ldh, E, act, reinterpret_cast<float*>(dZ), lddz, out_cols, dW, db);
  return finish_launch("dfol_table_layer_bwd_fused");
}

extern "C" int dfol_table_grad_dense(const float* g, const int32_t* slice_goff, const int32_t* slice_col,
                                     const int32_t* slice_img, int slice_num, const float* ll, const int64_t* blk,
                                     const int32_t* stride, const int32_t* row0, const int32_t* img_rows, float* dZ,
                                     int64_t lddz, void* stream) {
  DFOL_REQUIRE(g && slice_goff && slice_col && slice_img && ll && blk && stride && row0 && img_rows && dZ,
               "dfol_table_grad_dense: null pointer");
  if (slice_num == 0) return 0;
  table_grad_dense_kernel<<<(slice_num + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
      g, slice_goff, slice_col, slice_img, slice_num, ll, blk, stride, row0, img_rows, dZ, lddz);
  return finish_launch("dfol_table_grad_dense");
}

extern "C" int dfol_pair_hidden_bwd_bf16(const void* dz, int64_t lddz, const float* obj_pos, int64_t ldpos, float* duv,
                                         int64_t lduv, float* dwg, int64_t ldw, float* dbias, int H,
                                         const int32_t* pair_row, const int32_t* obj_row, const int32_t* img_n,
                                         int image_num, int max_n, void* stream) {
  DFOL_REQUIRE(dz && obj_pos && duv && dwg && dbias && pair_row && obj_row && img_n,
               "dfol_pair_hidden_bwd_bf16: null pointer");
  if (image_num == 0) return 0;
  const size_t smem = (size_t)max_n * H * sizeof(float);
  DFOL_REQUIRE(smem <= 200 * 1024, "dfol_pair_hidden_bwd_bf16: N * H too large for shared memory");
  cudaFuncSetAttribute(pair_hidden_bwd_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pair_hidden_bwd_bf16_kernel<<<image_num, 256, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dz), lddz, obj_pos, ldpos, duv, lduv, dwg, ldw, dbias, H, pair_row,
      obj_row, img_n);
  return finish_launch("dfol_pair_hidden_bwd_bf16");
}

extern "C" int dfol_colsum_bf16(const void* X, int64_t ldx, int64_t M, int N, float* out, void* stream) {
  DFOL_REQUIRE(X && out, "dfol_colsum_bf16: null pointer");
  if (M == 0 || N == 0) return 0;
  DFOL_REQUIRE((ldx % 2) == 0 && (reinterpret_cast<uintptr_t>(X) % 4) == 0, "dfol_colsum_bf16: ldx must be even");
  long long rows_per_block = 256;
  dim3 grid((N + 63) / 64, (unsigned)((M + rows_per_block - 1) / rows_per_block));
  colsum_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(X), ldx, M, N, out,
                                                            rows_per_block);
  return finish_launch("dfol_colsum_bf16");
}
