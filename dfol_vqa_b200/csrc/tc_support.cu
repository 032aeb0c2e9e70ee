// HBM-bound companions of the tensor-core (bf16) mode: operand preparation, the pair hidden layer (forward and
// backward) and the backward of the table layer.  Every kernel here streams its big operand exactly once with
// 128-byte warp transactions and keeps its reductions in registers / shared memory (see include/dfol_b200.h).
#include <cstdlib>
#include <type_traits>
#include <cuda_bf16.h>

#include "dfol_common.cuh"

namespace dfol {

// ---------------------------------------------------------------------------------------------------------
// Batched fp32 -> bf16 operand preparation: one launch casts (and optionally transposes) every weight matrix of
// the step.  dst is [out_rows][ldd] with zero padding beyond the source extent.
struct CastJob {
  const float* src; long long lds; int rows, cols;       // source view [rows][cols]
  __nv_bfloat16* dst; long long ldd; int out_rows;       // destination rows (>= rows, or >= cols if transposed)
  int transpose;                                         // dst[c][r] = src[r][c]
  int dcols;                                             // destination columns written per row (<= ldd)
  int pad;
};

__global__ void __launch_bounds__(256) cast_jobs_kernel(const CastJob* __restrict__ jobs) {
  const CastJob j = jobs[blockIdx.y];
  const long long total = (long long)j.out_rows * j.dcols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / j.dcols;
    const int c = (int)(i - r * j.dcols);
    float v = 0.0f;
    if (!j.transpose) {
      if (r < j.rows && c < j.cols) v = j.src[r * j.lds + c];
    } else {
      if (c < j.rows && r < j.cols) v = j.src[(long long)c * j.lds + r];
    }
    j.dst[r * j.ldd + c] = __float2bfloat16(v);
  }
}

// obj[t, F:F+4] = box position (batch_gqa_boxfeatures_pipeline.py:208-211) and obj16 = bf16(obj) with zero K-padding.
__global__ void __launch_bounds__(256) obj_finish_kernel(const float* __restrict__ f, long long ldf, int D,
                                                         float* __restrict__ obj, long long ldo, int F,
                                                         __nv_bfloat16* __restrict__ obj16, long long ld16,
                                                         long long rows) {
  const long long total = rows * ld16;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / ld16;
    const int c = (int)(i - t * ld16);
    float v = 0.0f;
    if (c < F) {
      v = obj[t * ldo + c];
    } else if (c < F + 4) {
      const float* r = f + t * ldf + D;  // [W, H, x, y, w, h]
      const int jj = c - F;
      v = r[2 + jj] / fmaxf(r[jj & 1], 1.0f);
      obj[t * ldo + c] = v;
    }
    obj16[i] = __float2bfloat16(v);
  }
}

// ---------------------------------------------------------------------------------------------------------
// geometry features of an ordered pair (s,o): dist, asin(dy/dist), sign(x_o-x_s), sign(y_o-y_s)
// (batch_gqa_boxfeatures_pipeline.py:260-279)
__device__ __forceinline__ float4 pair_geometry4(const float4 ps, const float4 po) {
  const float dx = ps.x + ps.z / 2.0f - po.x - po.z / 2.0f;
  const float dy = ps.y + ps.w / 2.0f - po.y - po.w / 2.0f;
  const float dist = sqrtf(dx * dx + dy * dy);
  float4 g;
  g.x = dist;
  g.y = asinf(dy / fmaxf(dist, 1e-10f));
  const float sx = po.x - ps.x, sy = po.y - ps.y;
  g.z = (sx > 0.0f) ? 1.0f : (sx < 0.0f ? -1.0f : 0.0f);
  g.w = (sy > 0.0f) ? 1.0f : (sy < 0.0f ? -1.0f : 0.0f);
  return g;
}

// Pair hidden layer (tensor-core mode): h[(s,o), :] = elu(U[s] + V[o] + Wg.geo(s,o) + b) in bf16, H = 256.
// One block per (image, 32-object tile): the V rows of the tile live in shared memory for all subjects, warp w walks
// the subjects s = w, w+8, ...; lane q owns hidden units 4q..4q+3 and 128+4q..; the geometry of (s, o0+lane) is
// computed by lane `lane` and broadcast with shuffles (no barrier in the loop).  Also writes the geometry table
// geo[pair] (float4) that the backward kernel re-uses.
// Packed fp32 pair arithmetic (sm_100a FFMA2 / FADD2): both halves are IEEE round-to-nearest, i.e. bit-identical to the
// scalar fmaf / add they replace, at half the issue slots.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// P = true: one packed instruction; P = false: the two scalar operations it stands for (same roundings).  FFMA2 is
// issued at half the FFMA rate per instruction on one half of the fma pipe: it saves issue slots, not pipe cycles, so
// which chains are packed is decided per kernel by measurement.
template <bool P>
__device__ __forceinline__ uint64_t fmaX(uint64_t a, uint64_t b, uint64_t c) {
  if (P) return fma2(a, b, c);
  float a0, a1, b0, b1, c0, c1;
  unpack2(a, a0, a1); unpack2(b, b0, b1); unpack2(c, c0, c1);
  return pack2(fmaf(a0, b0, c0), fmaf(a1, b1, c1));
}
template <bool P>
__device__ __forceinline__ uint64_t addX(uint64_t a, uint64_t b) {
  if (P) return add2(a, b);
  float a0, a1, b0, b1;
  unpack2(a, a0, a1); unpack2(b, b0, b1);
  return pack2(a0 + b0, a1 + b1);
}
template <bool P>
__device__ __forceinline__ uint64_t mulX(uint64_t a, uint64_t b) {
  if (P) return mul2(a, b);
  float a0, a1, b0, b1;
  unpack2(a, a0, a1); unpack2(b, b0, b1);
  return pack2(a0 * b0, a1 * b1);
}
// ELU with one MUFU: ex2.approx.ftz(a * log2 e) - 1 on the negative side.  Differs from __expf only where e^a is
// subnormal (a < -87), where both give -1 after the subtraction.
__device__ __forceinline__ float elu_fast(float a) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * 1.4426950408889634f));
  return a > 0.0f ? a : e - 1.0f;
}

// which chains run as packed pairs (measured on B200): 0 none, 1 all, 2 / 3 part of them
#define DFOL_PK_FWD_DEFAULT 1
#define DFOL_PK_PHB_RING_DEFAULT 1
#define DFOL_PK_TBL_DEFAULT 1

constexpr int PF_TO = 32;

template <int G, int PK>  // G = H / 128 (float4 groups per lane); PK: 0 scalar, 1 packed, 2 = half of the chains packed
__global__ void __launch_bounds__(256, G <= 2 ? 3 : 1) pair_hidden_fwd_tc_kernel(
    const float* __restrict__ uv, long long lduv, const float* __restrict__ pos, long long ldpos,
    const float* __restrict__ wg, long long ldw, const float* __restrict__ bias, __nv_bfloat16* __restrict__ hout,
    long long ldh, float4* __restrict__ geo_out, const int32_t* __restrict__ pair_row,
    const int32_t* __restrict__ obj_row, const int32_t* __restrict__ img_n, int subjects_per_block) {
  constexpr int H = 128 * G, H4 = H / 4;
  extern __shared__ float4 vsm[];  // [PF_TO][H4]
  const int b = blockIdx.y;
  const int n = img_n[b];
  const int o0 = blockIdx.x * PF_TO;
  const int s_begin = blockIdx.z * subjects_per_block;
  if (o0 >= n || s_begin >= n) return;
  const int s_end = min(n, s_begin + subjects_per_block);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long t0 = obj_row[b];
  const long long p0 = pair_row[b];
  const int no = min(PF_TO, n - o0);
  for (int idx = threadIdx.x; idx < no * H4; idx += 256) {
    const int o = idx / H4, q = idx - o * H4;
    vsm[o * H4 + q] = __ldg(reinterpret_cast<const float4*>(uv + (t0 + o0 + o) * lduv + H) + q);
  }
  // packed fp32 pairs (FFMA2 / FADD2 on sm_100a: two IEEE fma per issue slot): elements (4q, 4q+1) and (4q+2, 4q+3)
  // of a lane's float4 group; wk[i][c][h] = geometry weights of component c for pair h
  uint64_t bz[G][2], wk[G][4][2];
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const int q = lane + 32 * i;
    const float4 bq = __ldg(reinterpret_cast<const float4*>(bias) + q);
    bz[i][0] = pack2(bq.x, bq.y); bz[i][1] = pack2(bq.z, bq.w);
    const float* w = wg + (long long)(4 * q) * ldw;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      wk[i][c][0] = pack2(w[c], w[ldw + c]);
      wk[i][c][1] = pack2(w[2 * ldw + c], w[3 * ldw + c]);
    }
  }
  float4 po = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < no) po = __ldg(reinterpret_cast<const float4*>(pos + (t0 + o0 + lane) * ldpos));
  __syncthreads();
  const int zero_cols = (int)ldh - H;
  for (int s = s_begin + warp; s < s_end; s += 8) {
    const float4 ps = __ldg(reinterpret_cast<const float4*>(pos + (t0 + s) * ldpos));
    float4 gl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < no && o0 + lane != s) gl = pair_geometry4(ps, po);
    const long long prow = p0 + (long long)s * n + o0;
    if (geo_out != nullptr && lane < no) geo_out[prow + lane] = gl;
    uint64_t u[G][2];
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const float4 uq = __ldg(reinterpret_cast<const float4*>(uv + (t0 + s) * lduv) + lane + 32 * i);
      u[i][0] = addX<PK == 1>(pack2(uq.x, uq.y), bz[i][0]);
      u[i][1] = addX<PK == 1>(pack2(uq.z, uq.w), bz[i][1]);
    }
    for (int oi = 0; oi < no; ++oi) {
      uint64_t g[4];
      {
        const float gx = __shfl_sync(0xffffffffu, gl.x, oi), gy = __shfl_sync(0xffffffffu, gl.y, oi);
        const float gz = __shfl_sync(0xffffffffu, gl.z, oi), gw = __shfl_sync(0xffffffffu, gl.w, oi);
        g[0] = pack2(gx, gx); g[1] = pack2(gy, gy); g[2] = pack2(gz, gz); g[3] = pack2(gw, gw);
      }
      __nv_bfloat16* dst = hout + (prow + oi) * ldh;
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const int q = lane + 32 * i;
        const float4 v = vsm[oi * H4 + q];
        uint64_t a01 = addX<PK == 1>(u[i][0], pack2(v.x, v.y)), a23 = addX<PK == 1>(u[i][1], pack2(v.z, v.w));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          a01 = fmaX<PK >= 1>(wk[i][c][0], g[c], a01);
          a23 = fmaX<PK == 1>(wk[i][c][1], g[c], a23);
        }
        float a0, a1, a2, a3;
        unpack2(a01, a0, a1);
        unpack2(a23, a2, a3);
        a0 = elu_fast(a0); a1 = elu_fast(a1); a2 = elu_fast(a2); a3 = elu_fast(a3);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(a0, a1), hi = __floats2bfloat162_rn(a2, a3);
        *reinterpret_cast<uint2*>(dst + 4 * q) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
      }
      for (int h = lane; h < zero_cols; h += 32) dst[H + h] = __float2bfloat16(0.0f);
    }
  }
}

// Backward of the pair hidden layer from bf16 dZ1 (activation derivative already applied by the dgrad epilogue).
// One block per (image, CW hidden units), CW = 64 or 32: CW/2 consecutive threads form a row group that reads one row
// (s,o) of the chunk (thread l owns columns 2l, 2l+1); row group rg owns the objects o = rg, rg + NRG, ...  For every
// subject s a thread loads its NI rows, so
//   dV[o][h] = sum_s dz   is thread-private (registers, written once at the end),
//   dU[s][h] = sum_o dz   is reduced over the row groups through a double-buffered shared tile (one barrier per s),
//   dWg[h][k] = sum dz*geo_k and db[h] = sum dz stay in registers until the end (one atomic per block and column).
// dU / dV are written as bf16 (operands of the next tensor-core GEMMs).  CW = 32 halves the per-thread state (more
// resident blocks) and doubles the number of blocks: several waves instead of 1.4 on 148 SMs.
template <int THREADS, int NI, int CW, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) pair_hidden_bwd_tc_kernel(
    const __nv_bfloat16* __restrict__ dz, long long lddz, const float4* __restrict__ geo,
    __nv_bfloat16* __restrict__ du_out, __nv_bfloat16* __restrict__ dv_out, long long ldo, float* __restrict__ dwg,
    long long ldw, float* __restrict__ dbias, const int32_t* __restrict__ pair_row,
    const int32_t* __restrict__ obj_row, const int32_t* __restrict__ img_n) {
  constexpr int LPR = CW / 2;          // lanes per row
  constexpr int NRG = THREADS / LPR;   // row groups
  __shared__ __align__(16) float red[2][NRG][CW];
  __shared__ __align__(16) float redw[NRG][4][CW];
  __shared__ float4 gsm[2][NRG * NI];
  const int b = blockIdx.y;
  const int n = img_n[b];
  if (n == 0) return;  // image without pair rows (demand-driven: its program reads no relation)
  const long long t0 = obj_row[b];
  const long long p0 = pair_row[b];
  const int rg = threadIdx.x / LPR, l = threadIdx.x % LPR;
  const int c0 = blockIdx.x * CW + 2 * l;
  float dv[NI][2];
#pragma unroll
  for (int i = 0; i < NI; ++i) dv[i][0] = dv[i][1] = 0.0f;
  float dw[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  float dbv = 0.0f;  // threads < CW: column sum of dU
  // software pipeline: the NI row loads and the geometry row of subject s + 1 are issued before subject s is reduced, so
  // that loads stay in flight across the per-subject barrier (raw bf16x2 words: one register each); the geometry row
  // goes through a double-buffered shared tile (one float4 per object, broadcast reads)
  const long long srow = (long long)n * lddz;
  const __nv_bfloat16* sbase = dz + p0 * lddz + c0;
  const float4* gbase = geo + p0;
  uint32_t raw[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int o = rg + NRG * i;
    raw[i] = 0u;
    if (o < n && o != 0) raw[i] = *reinterpret_cast<const uint32_t*>(sbase + o * (int)lddz);
  }
  if ((int)threadIdx.x < n) gsm[0][threadIdx.x] = __ldg(gbase + threadIdx.x);
  __syncthreads();
  for (int s = 0; s < n; ++s) {
    uint32_t nxt[NI];
    float4 gn = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool more = s + 1 < n;
    sbase += srow; gbase += n;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int o = rg + NRG * i;
      nxt[i] = 0u;
      if (more && o < n && o != s + 1) nxt[i] = *reinterpret_cast<const uint32_t*>(sbase + o * (int)lddz);
    }
    if (more && (int)threadIdx.x < n) gn = __ldg(gbase + threadIdx.x);
    const float4* gs = gsm[s & 1];
    float du0 = 0.f, du1 = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int o = rg + NRG * i;
      if (o < n && o != s) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[i]));
        const float4 g = gs[o];
        du0 += v.x; du1 += v.y;
        dv[i][0] += v.x; dv[i][1] += v.y;
        dw[0][0] = fmaf(v.x, g.x, dw[0][0]); dw[0][1] = fmaf(v.y, g.x, dw[0][1]);
        dw[1][0] = fmaf(v.x, g.y, dw[1][0]); dw[1][1] = fmaf(v.y, g.y, dw[1][1]);
        dw[2][0] = fmaf(v.x, g.z, dw[2][0]); dw[2][1] = fmaf(v.y, g.z, dw[2][1]);
        dw[3][0] = fmaf(v.x, g.w, dw[3][0]); dw[3][1] = fmaf(v.y, g.w, dw[3][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) raw[i] = nxt[i];
    if (more && (int)threadIdx.x < n) gsm[(s + 1) & 1][threadIdx.x] = gn;
    const int buf = s & 1;
    *reinterpret_cast<float2*>(&red[buf][rg][2 * l]) = make_float2(du0, du1);
    __syncthreads();
    if (threadIdx.x < CW) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < NRG; ++r) t += red[buf][r][threadIdx.x];
      du_out[(t0 + s) * ldo + blockIdx.x * CW + threadIdx.x] = __float2bfloat16(t);
      dbv += t;
    }
  }
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int o = rg + NRG * i;
    if (o < n) {
      const __nv_bfloat162 x = __floats2bfloat162_rn(dv[i][0], dv[i][1]);
      *reinterpret_cast<__nv_bfloat162*>(dv_out + (t0 + o) * ldo + c0) = x;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) *reinterpret_cast<float2*>(&redw[rg][k][2 * l]) = make_float2(dw[k][0], dw[k][1]);
  __syncthreads();
  if (threadIdx.x < CW) {
    const int h = blockIdx.x * CW + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < NRG; ++r) t += redw[r][k][threadIdx.x];
      atomicAdd(dwg + (long long)h * ldw + k, t);
    }
    atomicAdd(dbias + h, dbv);
  }
}

// The same backward with the rows staged through shared memory by cp.async (16-byte LDGSTS, no registers held by
// loads in flight): a ring of STAGES subject tiles (n rows x 128 B of the block's 64 columns + the n geometry float4) is
// kept in flight per block, so HBM latency is covered by bytes in flight rather than by resident warps.  One barrier per
// subject: it publishes the tile of subject s, frees the slot of subject s - 1 for the next copy, and separates the
// dU partial sums of s - 1 (reduced by the first 64 threads while everybody works on s) from those of s + 1.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}

template <int THREADS, int NI, int STAGES, int PK>
__global__ void __launch_bounds__(THREADS) pair_hidden_bwd_async_kernel(
    const __nv_bfloat16* __restrict__ dz, long long lddz, const float4* __restrict__ geo,
    __nv_bfloat16* __restrict__ du_out, __nv_bfloat16* __restrict__ dv_out, long long ldo, float* __restrict__ dwg,
    long long ldw, float* __restrict__ dbias, const int32_t* __restrict__ pair_row,
    const int32_t* __restrict__ obj_row, const int32_t* __restrict__ img_n) {
  constexpr int RG = THREADS / 32, ROWS = RG * NI;
  constexpr int STAGE_BYTES = ROWS * 128 + ROWS * 16;
  extern __shared__ __align__(16) uint8_t ring[];
  __shared__ __align__(16) float red[2][RG][64];
  __shared__ __align__(16) float redw[RG][4][64];
  const int b = blockIdx.y;
  const int n = img_n[b];
  if (n == 0) return;  // image without pair rows
  const long long t0 = obj_row[b];
  const long long p0 = pair_row[b];
  const int rg = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 64 + 2 * lane;
  const __nv_bfloat16* src0 = dz + p0 * lddz + blockIdx.x * 64;
  const float4* geo0 = geo + p0;

  auto issue = [&](int s) {
    uint8_t* dst = ring + (s % STAGES) * STAGE_BYTES;
    const __nv_bfloat16* src = src0 + (long long)s * n * lddz;
    for (int c = threadIdx.x; c < n * 8; c += THREADS) {
      const int o = c >> 3, part = c & 7;
      cp_async16(dst + o * 128 + part * 16, src + (long long)o * lddz + part * 8);
    }
    if ((int)threadIdx.x < n) cp_async16(dst + ROWS * 128 + threadIdx.x * 16, geo0 + (long long)s * n + threadIdx.x);
  };
  auto reduce_du = [&](int s, float& dbv) {  // threads < 64
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < RG; ++r) t += red[s & 1][r][threadIdx.x];
    du_out[(t0 + s) * ldo + blockIdx.x * 64 + threadIdx.x] = __float2bfloat16(t);
    dbv += t;
  };

  uint64_t dv[NI], dw[4];  // packed fp32 pairs: the thread's two columns
#pragma unroll
  for (int i = 0; i < NI; ++i) dv[i] = pack2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) dw[k] = pack2(0.f, 0.f);
  float dbv = 0.0f;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < n) issue(s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int s = 0; s < n; ++s) {
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
    __syncthreads();
    if (s + STAGES - 1 < n) issue(s + STAGES - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (s > 0 && threadIdx.x < 64) reduce_du(s - 1, dbv);
    const uint8_t* tile = ring + (s % STAGES) * STAGE_BYTES;
    const float4* gs = reinterpret_cast<const float4*>(tile + ROWS * 128);
    uint64_t du = pack2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int o = rg + RG * i;
      if (o < n && o != s) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(tile + o * 128 + lane * 4));
        const float4 g = gs[o];
        const uint64_t v2 = pack2(v.x, v.y);
        du = addX<PK >= 1>(du, v2);
        dv[i] = addX<PK >= 1>(dv[i], v2);
        dw[0] = fmaX<PK >= 1>(v2, pack2(g.x, g.x), dw[0]);
        dw[1] = fmaX<PK >= 1>(v2, pack2(g.y, g.y), dw[1]);
        dw[2] = fmaX<PK == 1>(v2, pack2(g.z, g.z), dw[2]);
        dw[3] = fmaX<PK == 1>(v2, pack2(g.w, g.w), dw[3]);
      }
    }
    *reinterpret_cast<uint64_t*>(&red[s & 1][rg][2 * lane]) = du;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (n > 0 && threadIdx.x < 64) reduce_du(n - 1, dbv);
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int o = rg + RG * i;
    if (o < n) {
      float d0, d1;
      unpack2(dv[i], d0, d1);
      const __nv_bfloat162 x = __floats2bfloat162_rn(d0, d1);
      *reinterpret_cast<__nv_bfloat162*>(dv_out + (t0 + o) * ldo + c0) = x;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) *reinterpret_cast<uint64_t*>(&redw[rg][k][2 * lane]) = dw[k];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int h = blockIdx.x * 64 + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < RG; ++r) t += redw[r][k][threadIdx.x];
      atomicAdd(dwg + (long long)h * ldw + k, t);
    }
    atomicAdd(dbias + h, dbv);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Backward of the table layer LL = logsigmoid(H.W^T + b) in tensor-core mode (bf16 H in, bf16 dZ out) for up to
// S table columns per image and launch (slices [j0 + first, j0 + first + S) of image b):
//   dz_j[l]   = g_j[l] * (1 - exp(LL_j[l]))
//   dZ[row,e] (+)= (sum_j dz_j[l] W[wrow_j, e]) * h(1-h)          (sigmoid' of the layer below, h = H[row,e])
//   dW[wrow_j, e] += sum_l dz_j[l] h;  db[wrow_j] += sum_l dz_j[l];  dbelow[e] += sum_rows dZ[row,e]   (atomics)
// One block per (128-row chunk, image); warp w streams rows w, w+8, ... of the chunk, lane owns the column pairs
// 2*lane + 64*c (c < NC): every row is read and written with 128-byte warp transactions, exactly once.
constexpr int TB_ROWS = 128;
constexpr int TB_RP = 2;  // row-parallel warps per column chunk

template <int S, int NC, bool DROP, bool PK>  // PK: the FMA chains as packed fp32 pairs (FFMA2)
__global__ void __launch_bounds__(32 * NC * TB_RP) table_layer_bwd_tc_kernel(
    const float* __restrict__ g, const int32_t* __restrict__ slice_goff, const int32_t* __restrict__ slice_col,
    const int32_t* __restrict__ slice_wrow, const int32_t* __restrict__ img_slice, int first, int accumulate,
    const float* __restrict__ ll, const int64_t* __restrict__ blk, const int32_t* __restrict__ stride,
    const int32_t* __restrict__ row0, const int32_t* __restrict__ img_rows, const float* __restrict__ W,
    long long ldw, const __nv_bfloat16* __restrict__ hs, long long ldh, int E, __nv_bfloat16* __restrict__ dZ,
    long long lddz, int out_cols, float* __restrict__ dW, float* __restrict__ db, float* __restrict__ dbelow,
    float keep) {
  constexpr int THREADS = 32 * NC * TB_RP;
  // keep < 1: hs holds the activation AFTER dropout (0 where dropped, h / keep where kept): it is the operand of the dW
  // rows as it stands, sigmoid' is taken at h = hs * keep and the gradient carries the mask factor (hs != 0) / keep
  // (DROP is a template parameter: the instantiation without dropout is the kernel it was before)
  const float inv_keep = DROP ? 1.0f / keep : 1.0f;
  __shared__ float dz_s[S][TB_ROWS];
  __shared__ int wrow_s[S];
  __shared__ __align__(16) float red_s[TB_RP][NC * 64];
  const int b = blockIdx.y;
  const int rows = img_rows[b];
  const int c = blockIdx.x * TB_ROWS;
  if (c >= rows) return;
  const int cn = min(TB_ROWS, rows - c);
  const long long r0 = (long long)row0[b] + c;
  const int j0 = img_slice[b] + first;
  const int Sb = max(0, min(img_slice[b + 1] - j0, S));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = warp % NC, rp = warp / NC;   // column chunk of this warp, row phase
  const int e = 64 * k + 2 * lane;           // column pair of this thread
  if (Sb == 0) {
    if (!accumulate && e < out_cols)  // rows of images without (more) slices are zero
      for (int l = rp; l < cn; l += TB_RP) *reinterpret_cast<uint32_t*>(dZ + (r0 + l) * lddz + e) = 0u;
    return;
  }
  const int st = stride[b];
  for (int idx = threadIdx.x; idx < S * TB_ROWS; idx += THREADS) {
    const int j = idx / TB_ROWS, l = idx - j * TB_ROWS;
    float v = 0.0f;
    if (j < Sb && l < cn) {
      const float gv = g[slice_goff[j0 + j] + c + l];
      if (gv != 0.0f) v = gv * (1.0f - __expf(ll[blk[b] + (long long)slice_col[j0 + j] * st + c + l]));
    }
    dz_s[j][l] = v;
  }
  if (threadIdx.x < S) wrow_s[threadIdx.x] = threadIdx.x < Sb ? slice_wrow[j0 + threadIdx.x] : 0;
  __syncthreads();
  for (int j = warp; j < Sb; j += THREADS / 32) {  // db[wrow_j] += sum_l dz_j[l]
    float t = 0.f;
    for (int l = lane; l < cn; l += 32) t += dz_s[j][l];
    t = warp_sum(t);
    if (lane == 0 && t != 0.0f) atomicAdd(db + wrow_s[j], t);
  }
  const bool e_ok = e < E, st_ok = e < out_cols;
  // The row loop is specialised on the number of slices of THIS image (block-uniform): images that use fewer slices
  // than the launch maximum S run code without the predicated-off FMAs of the missing ones (9 of 12 at c3).
  auto run = [&](auto sb_tag) {
    constexpr int SB = decltype(sb_tag)::value;   // 0: generic (S slots guarded by Sb); > 0: exactly SB slices
    constexpr int SN = SB > 0 ? SB : S;
    float2 wj[SN], dwj[SN], colsum = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < SN; ++j) {
      dwj[j] = make_float2(0.f, 0.f);
      wj[j] = ((SB > 0 || j < Sb) && e_ok) ? *reinterpret_cast<const float2*>(W + (long long)wrow_s[j] * ldw + e)
                               : make_float2(0.f, 0.f);
    }
    constexpr int UN = 8;  // rows in flight per warp
    // row pointers advance by a fixed stride: no 64-bit multiply per access
    const __nv_bfloat16* hp = hs + (r0 + rp) * ldh + e;
    __nv_bfloat16* zp = dZ + (r0 + rp) * lddz + e;
    const long long hstep = (long long)TB_RP * ldh, zstep = (long long)TB_RP * lddz;
    for (int l0 = rp; l0 < cn; l0 += TB_RP * UN, hp += UN * hstep, zp += UN * zstep) {
      uint32_t hraw[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        hraw[u] = 0u;
        if (l0 + TB_RP * u < cn && e_ok) hraw[u] = *reinterpret_cast<const uint32_t*>(hp + u * hstep);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int l = l0 + TB_RP * u;
        if (l < cn) {
          const float2 hv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hraw[u]));
          float ox = 0.f, oy = 0.f;
#pragma unroll
          for (int j = 0; j < SN; ++j) {
            if (SB > 0 || j < Sb) {  // (generic form: block-uniform guard)
              const float dzl = dz_s[j][l];
              if (PK) {
                const uint64_t d2 = pack2(dzl, dzl);
                const uint64_t o2 = fma2(d2, pack2(wj[j].x, wj[j].y), pack2(ox, oy));
                const uint64_t w2 = fma2(d2, pack2(hv.x, hv.y), pack2(dwj[j].x, dwj[j].y));
                unpack2(o2, ox, oy);
                unpack2(w2, dwj[j].x, dwj[j].y);
              } else {
                ox = fmaf(dzl, wj[j].x, ox);
                oy = fmaf(dzl, wj[j].y, oy);
                dwj[j].x = fmaf(dzl, hv.x, dwj[j].x);
                dwj[j].y = fmaf(dzl, hv.y, dwj[j].y);
              }
            }
          }
          if (DROP) {
            const float hx = hv.x * keep, hy = hv.y * keep;
            ox *= (hv.x != 0.0f ? inv_keep : 0.0f) * hx * (1.0f - hx);
            oy *= (hv.y != 0.0f ? inv_keep : 0.0f) * hy * (1.0f - hy);
          } else {
            ox *= hv.x * (1.0f - hv.x);
            oy *= hv.y * (1.0f - hv.y);
          }
          colsum.x += ox;
          colsum.y += oy;
          if (st_ok) {  // columns E .. out_cols are the zero K-padding of the next GEMM (h = 0 there)
            if (accumulate && e_ok) {
              const float2 pv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(zp + u * zstep));
              ox += pv.x;
              oy += pv.y;
            }
            *reinterpret_cast<__nv_bfloat162*>(zp + u * zstep) = __floats2bfloat162_rn(ox, oy);
          }
        }
      }
    }
    // block reductions over the row phases: dbelow (column sums of dZ) and the dW rows, one atomic per column
    *reinterpret_cast<float2*>(&red_s[rp][e]) = colsum;
    __syncthreads();
    for (int x = threadIdx.x; x < E; x += THREADS) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < TB_RP; ++r) t += red_s[r][x];
      if (dbelow != nullptr && t != 0.0f) atomicAdd(dbelow + x, t);
    }
#pragma unroll
    for (int j = 0; j < SN; ++j) {
      if (SB == 0 && j >= Sb) break;
      __syncthreads();
      *reinterpret_cast<float2*>(&red_s[rp][e]) = dwj[j];
      __syncthreads();
      for (int x = threadIdx.x; x < E; x += THREADS) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < TB_RP; ++r) t += red_s[r][x];
        if (t != 0.0f) atomicAdd(dW + (long long)wrow_s[j] * ldw + x, t);
      }
    }
  };
  if (Sb == S) run(std::integral_constant<int, S>());
  else if (S > 1 && Sb == S - 1) run(std::integral_constant<int, (S > 1 ? S - 1 : 1)>());
  else if (S > 2 && Sb == S - 2) run(std::integral_constant<int, (S > 2 ? S - 2 : 1)>());
  else if (S > 4 && Sb == S - 3) run(std::integral_constant<int, (S > 4 ? S - 3 : 1)>());
  else run(std::integral_constant<int, 0>());

}

// ---------------------------------------------------------------------------------------------------------
// Demand-driven relation table (tensor-core mode): only the relation columns ("slots") the image's own program uses
// are evaluated, LL[b][slot j][l] = logsigmoid(H2[row0[b] + l, :] . W[wrow_j, :] + bias[wrow_j]), self pairs = diag.
// Same streaming structure as the table-layer backward: block = (128-row chunk, image), warp per row, lanes over
// column pairs, one warp reduction per (row, slot); results are staged in shared memory and written coalesced.
template <int S, int NC>
__global__ void __launch_bounds__(256) rel_slots_fwd_kernel(
    const __nv_bfloat16* __restrict__ hs, long long ldh, int E, const float* __restrict__ W, long long ldw,
    const float* __restrict__ bias, const int32_t* __restrict__ slot_wrow, const int32_t* __restrict__ img_slot,
    int first, const int64_t* __restrict__ slot_blk, const int32_t* __restrict__ stride,
    const int32_t* __restrict__ row0, const int32_t* __restrict__ img_rows, const int32_t* __restrict__ img_n,
    float diag, float* __restrict__ ll, float* __restrict__ pp) {
  __shared__ float zs[S][TB_ROWS];
  const int b = blockIdx.y;
  const int rows = img_rows[b];
  const int c = blockIdx.x * TB_ROWS;
  if (c >= rows) return;
  const int j0 = img_slot[b] + first;
  const int Sb = min(img_slot[b + 1] - j0, S);
  if (Sb <= 0) return;
  const int cn = min(TB_ROWS, rows - c);
  const long long r0 = (long long)row0[b] + c;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2 wj[S][NC];
#pragma unroll
  for (int j = 0; j < S; ++j) {
    const int wr = j < Sb ? slot_wrow[j0 + j] : 0;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int e = 2 * lane + 64 * k;
      wj[j][k] = (j < Sb && e < E) ? *reinterpret_cast<const float2*>(W + (long long)wr * ldw + e)
                                   : make_float2(0.f, 0.f);
    }
  }
  // two rows per warp iteration (10 row loads in flight per lane), pointer-increment addressing
  const __nv_bfloat16* hp = hs + (r0 + warp) * ldh + 2 * lane;
  const long long step = 8 * ldh;
  for (int l = warp; l < cn; l += 16, hp += 2 * step) {
    const bool two = l + 8 < cn;
    uint32_t raw0[NC], raw1[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const bool ok = 2 * lane + 64 * k < E;
      raw0[k] = ok ? *reinterpret_cast<const uint32_t*>(hp + 64 * k) : 0u;
      raw1[k] = (ok && two) ? *reinterpret_cast<const uint32_t*>(hp + step + 64 * k) : 0u;
    }
    float2 h0[NC], h1[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      h0[k] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw0[k]));
      h1[k] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw1[k]));
    }
#pragma unroll
    for (int j = 0; j < S; ++j) {
      if (j < Sb) {  // block-uniform
        float z0 = 0.f, z1 = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          z0 = fmaf(h0[k].x, wj[j][k].x, fmaf(h0[k].y, wj[j][k].y, z0));
          z1 = fmaf(h1[k].x, wj[j][k].x, fmaf(h1[k].y, wj[j][k].y, z1));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          z0 += __shfl_xor_sync(0xffffffffu, z0, o);
          z1 += __shfl_xor_sync(0xffffffffu, z1, o);
        }
        if (lane == 0) {
          zs[j][l] = z0;
          if (two) zs[j][l + 8] = z1;
        }
      }
    }
  }
  __syncthreads();
  const int n = img_n[b];
  const long long base = slot_blk[b] + (long long)first * stride[b] + c;
  for (int idx = threadIdx.x; idx < Sb * cn; idx += 256) {
    const int j = idx / cn, l = idx - j * cn;
    const int lg = c + l;
    const float x = zs[j][l] + __ldg(bias + slot_wrow[j0 + j]);
    const float e = __expf(-fabsf(x));
    const float v = fminf(x, 0.0f) - __logf(1.0f + e);
    const bool is_diag = (lg / n) == (lg % n);
    ll[base + (long long)j * stride[b] + l] = is_diag ? diag : v;
    if (pp != nullptr) pp[base + (long long)j * stride[b] + l] = is_diag ? 0.0f : __fdividef(x >= 0.0f ? 1.0f : e, 1.0f + e);
  }
}

}  // namespace dfol

using namespace dfol;

extern "C" int dfol_cast_jobs(const void* jobs, int job_num, int64_t max_elements, void* stream) {
  DFOL_REQUIRE(jobs && job_num > 0 && max_elements > 0, "dfol_cast_jobs: bad arguments");
  long long bx = (max_elements + 255) / 256;
  if (bx > 1024) bx = 1024;
  dim3 grid((unsigned)bx, job_num);
  cast_jobs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const CastJob*>(jobs));
  return finish_launch("dfol_cast_jobs");
}

extern "C" int dfol_cast_job_size(void) { return (int)sizeof(CastJob); }

extern "C" int dfol_obj_finish(const float* features, int64_t ldf, int feature_dim, float* obj, int64_t ldo, int F,
                               void* obj16, int64_t ld16, int64_t rows, void* stream) {
  DFOL_REQUIRE(features && obj && obj16 && ld16 >= F + 4 && ldo >= F + 4, "dfol_obj_finish: bad arguments");
  if (rows == 0) return 0;
  long long bx = (rows * ld16 + 255) / 256;
  if (bx > 148 * 16) bx = 148 * 16;
  obj_finish_kernel<<<(unsigned)bx, 256, 0, (cudaStream_t)stream>>>(features, ldf, feature_dim, obj, ldo, F,
                                                                    reinterpret_cast<__nv_bfloat16*>(obj16), ld16, rows);
  return finish_launch("dfol_obj_finish");
}

extern "C" int dfol_pair_hidden_fwd_tc(const float* uv, int64_t lduv, const float* obj_pos, int64_t ldpos,
                                       const float* wg, int64_t ldw, const float* bias, void* h_out, int64_t ldh, int H,
                                       void* geo_out, const int32_t* pair_row, const int32_t* obj_row,
                                       const int32_t* img_n, int image_num, int max_n, void* stream) {
  DFOL_REQUIRE(uv && obj_pos && wg && bias && h_out && pair_row && obj_row && img_n,
               "dfol_pair_hidden_fwd_tc: null pointer");
  if (image_num == 0) return 0;
  DFOL_REQUIRE(H == 128 || H == 256 || H == 384 || H == 512, "dfol_pair_hidden_fwd_tc: H must be 128/256/384/512");
  DFOL_REQUIRE((lduv % 4) == 0 && (ldh % 4) == 0 && ldh >= H && (ldpos % 4) == 0 &&
                   (reinterpret_cast<uintptr_t>(uv) % 16) == 0 && (reinterpret_cast<uintptr_t>(h_out) % 16) == 0 &&
                   (reinterpret_cast<uintptr_t>(obj_pos) % 16) == 0 && (reinterpret_cast<uintptr_t>(bias) % 16) == 0,
               "dfol_pair_hidden_fwd_tc: strides must be multiples of 4 and buffers 16-byte aligned");
  DFOL_REQUIRE(max_n >= 1, "dfol_pair_hidden_fwd_tc: empty batch");
  // subjects per block: every block first stages its 32-object V tile (32 KB) and its weights, so blocks as large as
  // the batch allows (measured: 8 subjects per block 0.150 ms, all 48 subjects 0.120 ms at B = 256, N = 48) while the
  // grid still fills the GPU (3 resident blocks on each of 148 SMs)
  static const int ts_env = [] { const char* e = getenv("DFOL_PF_SUBJECTS"); return e ? atoi(e) : 0; }();
  int ts;
  if (ts_env > 0) {
    ts = ((ts_env + 7) / 8) * 8;
  } else {
    const long long tiles = (long long)((max_n + PF_TO - 1) / PF_TO) * image_num;
    ts = max_n <= 64 ? ((max_n + 7) / 8) * 8 : (((max_n + 1) / 2 + 7) / 8) * 8;
    while (ts > 8 && tiles * ((max_n + ts - 1) / ts) < 3 * 148) ts -= 8;
  }
  dim3 grid((max_n + PF_TO - 1) / PF_TO, image_num, (max_n + ts - 1) / ts);
  const size_t smem = (size_t)PF_TO * (H / 4) * sizeof(float4);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(h_out);
  float4* geo = reinterpret_cast<float4*>(geo_out);
  static const int pk_env = [] { const char* e = getenv("DFOL_PK_FWD"); return e ? atoi(e) : -1; }();
  const int pk = pk_env >= 0 ? pk_env : DFOL_PK_FWD_DEFAULT;
#define DFOL_PF_LAUNCH_PK(G, PK)                                                                                 \
  {                                                                                                              \
    auto kern = pair_hidden_fwd_tc_kernel<G, PK>;                                                                \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                          \
    kern<<<grid, 256, smem, st>>>(uv, lduv, obj_pos, ldpos, wg, ldw, bias, out, ldh, geo, pair_row, obj_row,    \
                                  img_n, ts);                                                                    \
  }
#define DFOL_PF_LAUNCH(G)                                                                                        \
  {                                                                                                              \
    if (pk == 0) DFOL_PF_LAUNCH_PK(G, 0) else if (pk == 1) DFOL_PF_LAUNCH_PK(G, 1) else DFOL_PF_LAUNCH_PK(G, 2)  \
  }
  switch (H / 128) {
    case 1: DFOL_PF_LAUNCH(1) break;
    case 2: DFOL_PF_LAUNCH(2) break;
    case 3: DFOL_PF_LAUNCH(3) break;
    default: DFOL_PF_LAUNCH(4) break;
  }
#undef DFOL_PF_LAUNCH_PK
#undef DFOL_PF_LAUNCH
  return finish_launch("dfol_pair_hidden_fwd_tc");
}

extern "C" int dfol_pair_hidden_bwd_tc(const void* dz, int64_t lddz, const void* geo, void* du_out, void* dv_out,
                                       int64_t ldo, float* dwg, int64_t ldw, float* dbias, int H,
                                       const int32_t* pair_row, const int32_t* obj_row, const int32_t* img_n,
                                       int image_num, int max_n, void* stream) {
  DFOL_REQUIRE(dz && geo && du_out && dv_out && dwg && dbias && pair_row && obj_row && img_n,
               "dfol_pair_hidden_bwd_tc: null pointer");
  if (image_num == 0) return 0;
  DFOL_REQUIRE((H % 64) == 0 && (lddz % 2) == 0 && (ldo % 2) == 0 && max_n >= 1 && max_n <= 128,
               "dfol_pair_hidden_bwd_tc: H %% 64 == 0, even strides, 1 <= max_n <= 128");
  // DFOL_PB_MODE: 0 (default) = by image size; 1 = cp.async ring; 64 = register-pipelined kernel
  static const int pk_env = [] { const char* e = getenv("DFOL_PK_PHB"); return e ? atoi(e) : -1; }();
  static const int mode_env = [] { const char* e = getenv("DFOL_PB_MODE"); return e ? atoi(e) : 0; }();
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* dzp = reinterpret_cast<const __nv_bfloat16*>(dz);
  const float4* gp = reinterpret_cast<const float4*>(geo);
  __nv_bfloat16* dup = reinterpret_cast<__nv_bfloat16*>(du_out);
  __nv_bfloat16* dvp = reinterpret_cast<__nv_bfloat16*>(dv_out);
  // (measured, B = 256: N = 48 0.119 ms ring / 0.126 registers; N = 64 0.205 / 0.244; N = 100 0.486 / 0.467 -- the
  //  256-thread ring kernels need 122 registers, two blocks per SM)
  if ((mode_env == 0 ? max_n <= 64 : mode_env == 1) && (lddz % 8) == 0 && (reinterpret_cast<uintptr_t>(dz) % 16) == 0) {
    dim3 grid(H / 64, image_num);
#define DFOL_PBA_LAUNCH_PK(T, NI, STAGES, PK)                                                                     \
  {                                                                                                               \
    auto kern = pair_hidden_bwd_async_kernel<T, NI, STAGES, PK>;                                                  \
    const int smem = STAGES * ((T / 32) * NI) * 144;                                                              \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                                \
    kern<<<grid, T, smem, st>>>(dzp, lddz, gp, dup, dvp, ldo, dwg, ldw, dbias, pair_row, obj_row, img_n);         \
  }
#define DFOL_PBA_LAUNCH(T, NI, STAGES)                                                                            \
  {                                                                                                               \
    if (pk == 0) DFOL_PBA_LAUNCH_PK(T, NI, STAGES, 0)                                                             \
    else if (pk == 1) DFOL_PBA_LAUNCH_PK(T, NI, STAGES, 1)                                                        \
    else DFOL_PBA_LAUNCH_PK(T, NI, STAGES, 2)                                                                     \
  }
    const int pk = pk_env >= 0 ? pk_env : DFOL_PK_PHB_RING_DEFAULT;
    if (max_n <= 32) DFOL_PBA_LAUNCH(128, 8, 4)
    else if (max_n <= 48) DFOL_PBA_LAUNCH(128, 12, 4)
    else if (max_n <= 64) DFOL_PBA_LAUNCH(256, 8, 4)
    else if (max_n <= 104) DFOL_PBA_LAUNCH(256, 13, 4)
    else DFOL_PBA_LAUNCH(256, 16, 4)
#undef DFOL_PBA_LAUNCH
#undef DFOL_PBA_LAUNCH_PK
    return finish_launch("dfol_pair_hidden_bwd_tc");
  }
  dim3 grid(H / 64, image_num);
#define DFOL_PB_LAUNCH(T, NI, MINB)                                                                               \
  pair_hidden_bwd_tc_kernel<T, NI, 64, MINB><<<grid, T, 0, st>>>(dzp, lddz, gp, dup, dvp, ldo, dwg, ldw, dbias,    \
                                                                 pair_row, obj_row, img_n)
  // one warp per row: 4 row groups (128 threads) for small images.  Scalar FMAs: packed pairs measured slower here
  // (N = 100: 0.467 ms scalar, 0.532 packed)
  if (max_n <= 32) DFOL_PB_LAUNCH(128, 8, 5);
  else if (max_n <= 48) DFOL_PB_LAUNCH(128, 12, 5);
  else if (max_n <= 64) DFOL_PB_LAUNCH(256, 8, 2);
  else if (max_n <= 104) DFOL_PB_LAUNCH(256, 13, 2);
  else DFOL_PB_LAUNCH(256, 16, 2);
#undef DFOL_PB_LAUNCH
  return finish_launch("dfol_pair_hidden_bwd_tc");
}

extern "C" int dfol_table_layer_bwd_tc(const float* g, const int32_t* slice_goff, const int32_t* slice_col,
                                       const int32_t* slice_wrow, const int32_t* img_slice, int image_num,
                                       int max_rows, int max_slices, const float* ll, const int64_t* blk,
                                       const int32_t* stride, const int32_t* row0, const int32_t* img_rows,
                                       const float* W, int64_t ldw, const void* h_saved, int64_t ldh, int E, void* dZ,
                                       int64_t lddz, int out_cols, float* dW, float* db, float* dbelow, float keep,
                                       void* stream) {
  DFOL_REQUIRE(g && slice_goff && slice_col && slice_wrow && img_slice && ll && blk && stride && row0 && img_rows &&
                   W && h_saved && dZ && dW && db,
               "dfol_table_layer_bwd_tc: null pointer");
  DFOL_REQUIRE(keep > 0.0f && keep <= 1.0f, "dfol_table_layer_bwd_tc: keep = 1 - dropout p must be in (0, 1]");
  if (image_num == 0 || max_rows == 0) return 0;
  DFOL_REQUIRE(out_cols >= E && out_cols <= lddz && out_cols <= 320 && (E % 2) == 0 && (out_cols % 2) == 0 &&
                   (ldw % 2) == 0 && (ldh % 2) == 0 && (lddz % 2) == 0,
               "dfol_table_layer_bwd_tc: E <= out_cols <= min(lddz, 320), even sizes and strides");
  dim3 grid((max_rows + TB_ROWS - 1) / TB_ROWS, image_num);
  DFOL_REQUIRE(grid.y <= 65535, "dfol_table_layer_bwd_tc: too many images");
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(h_saved);
  __nv_bfloat16* dzp = reinterpret_cast<__nv_bfloat16*>(dZ);
  static const int pk_env = [] { const char* e = getenv("DFOL_PK_TBL"); return e ? atoi(e) : -1; }();
  const int pk = pk_env >= 0 ? pk_env : DFOL_PK_TBL_DEFAULT;
  // slices are consumed in groups of at most 12 per pass; later passes accumulate into dZ
  int first = 0;
  do {
    const int left = max_slices - first;
    const int acc = first > 0 ? 1 : 0;
#define DFOL_TB_ARGS g, slice_goff, slice_col, slice_wrow, img_slice, first, acc, ll, blk, stride, row0, img_rows, W, ldw, \
                     hp, ldh, E, dzp, lddz, out_cols, dW, db, dbelow, keep
#define DFOL_TB_LAUNCH(S)                                                                                \
  do {                                                                                                   \
    if (keep < 1.0f) table_layer_bwd_tc_kernel<S, 5, true, false><<<grid, 320, 0, st>>>(DFOL_TB_ARGS);   \
    else if (pk) table_layer_bwd_tc_kernel<S, 5, false, true><<<grid, 320, 0, st>>>(DFOL_TB_ARGS);       \
    else table_layer_bwd_tc_kernel<S, 5, false, false><<<grid, 320, 0, st>>>(DFOL_TB_ARGS);              \
  } while (0)
    if (left <= 1) { DFOL_TB_LAUNCH(1); first += 1; }
    else if (left == 2) { DFOL_TB_LAUNCH(2); first += 2; }
    else if (left <= 4) { DFOL_TB_LAUNCH(4); first += 4; }
    else if (left <= 8) { DFOL_TB_LAUNCH(8); first += 8; }
    else { DFOL_TB_LAUNCH(12); first += 12; }
#undef DFOL_TB_LAUNCH
#undef DFOL_TB_ARGS
  } while (first < max_slices);
  return finish_launch("dfol_table_layer_bwd_tc");
}

extern "C" int dfol_rel_slots_fwd(const void* h_saved, int64_t ldh, int E, const float* W, int64_t ldw,
                                  const float* bias, const int32_t* slot_wrow, const int32_t* img_slot, int max_slots,
                                  const int64_t* slot_blk, const int32_t* stride, const int32_t* row0,
                                  const int32_t* img_rows, const int32_t* img_n, int image_num, int max_rows,
                                  float diag_value, float* ll, float* p_out, void* stream) {
  DFOL_REQUIRE(h_saved && W && bias && slot_wrow && img_slot && slot_blk && stride && row0 && img_rows && img_n && ll,
               "dfol_rel_slots_fwd: null pointer");
  if (image_num == 0 || max_rows == 0 || max_slots == 0) return 0;
  DFOL_REQUIRE(E <= 320 && (E % 2) == 0 && (ldw % 2) == 0 && (ldh % 2) == 0,
               "dfol_rel_slots_fwd: E <= 320, even sizes and strides");
  dim3 grid((max_rows + TB_ROWS - 1) / TB_ROWS, image_num);
  DFOL_REQUIRE(grid.y <= 65535, "dfol_rel_slots_fwd: too many images");
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(h_saved);
  for (int first = 0; first < max_slots;) {
    const int left = max_slots - first;
#define DFOL_RS_LAUNCH(S)                                                                                          \
  rel_slots_fwd_kernel<S, 5><<<grid, 256, 0, st>>>(hp, ldh, E, W, ldw, bias, slot_wrow, img_slot, first, slot_blk,  \
                                                   stride, row0, img_rows, img_n, diag_value, ll, p_out)
    if (left <= 1) { DFOL_RS_LAUNCH(1); first += 1; }
    else if (left == 2) { DFOL_RS_LAUNCH(2); first += 2; }
    else if (left <= 4) { DFOL_RS_LAUNCH(4); first += 4; }
    else { DFOL_RS_LAUNCH(8); first += 8; }
#undef DFOL_RS_LAUNCH
  }
  return finish_launch("dfol_rel_slots_fwd");
}
