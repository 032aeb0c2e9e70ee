// Pair hidden layer on the tensor cores (sm_100a): H1[(s,o), :] = elu(U[s] + V[o] + Wg . geo(s,o) + b) as a GROUPED GEMM
//     H1_tile = elu(A_tile . B_b^T),   A[(s,o), :] = [e_s | e_o | geo(s,o) | 1 | 0 ...]   (one-hot rows, K = 2 n_b + 5)
//                                      B_b[h, :]   = [U_b[:, h] | V_b[:, h] | Wg[h, :] | bias[h] | 0 ...]
// (reference: the first Linear of the relation network applied to [obj_s | obj_o | geo],
// batch_gqa_boxfeatures_pipeline.py:260-281 + gqa_interpreter_experiments.py:167; U|V = obj . [W_s|W_o]^T is computed once
// per object by the caller.)  The SIMT kernel spends 4 FMAs per output element on the geometry term and runs at 1.8 TB/s;
// here the additions and the geometry term are ONE 128 x H x K MMA chain per 128-row tile and the epilogue is the ELU and
// the bf16 store only.
//
//   all threads   : zero the A tile(s) in shared memory (128-byte swizzle, K-major: the layout TMA would have produced);
//   threads 0-127 : row r of the tile = pair (s,o): geometry from the two boxes, then the <= 7 non-zeros of the row;
//   warp 0 lane 0 : TMA loads of the per-image operand B_b (gathered to bf16 by pair_hidden_operand_kernel);
//   warp 1 lane 0 : tcgen05.mma 128 x H x 16 into an H-column TMEM accumulator;
//   warps 2-5     : epilogue, lane = row: tcgen05.ld, ELU, bf16, 32-byte stores along the row.
#include "tc_common.cuh"

namespace dfol {

constexpr int PM_BM = 128;
constexpr int PM_BK = 64;
constexpr int PM_MAX_KB = 4;  // K = 2 n + 5 <= 256
constexpr int PM_THREADS = 192;

struct PmParams {
  int H, num_kb;
  const float* pos; long long ldpos;
  __nv_bfloat16* h_out; long long ldh;
  float4* geo_out;
  const int32_t* pair_row; const int32_t* obj_row; const int32_t* img_n;
};

__device__ __forceinline__ uint32_t pm_swz(int row, int col) {  // byte offset of bf16 (row, col) in a [128][64] SW128 box
  return (uint32_t)(row * 128 + (((col >> 3) ^ (row & 7)) << 4) + ((col & 7) << 1));
}

__global__ void __launch_bounds__(PM_THREADS) pair_hidden_mma_kernel(const __grid_constant__ CUtensorMap tmap_b,
                                                                     PmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[PM_MAX_KB];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int b = blockIdx.y;
  const int n = p.img_n[b];
  const int rows = n * n;
  const int c = blockIdx.x * PM_BM;
  if (c >= rows) return;  // (whole CTA: before any barrier / TMEM allocation)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, num_kb = p.num_kb;
  const uint32_t a_bytes = PM_BM * PM_BK * 2, b_bytes = (uint32_t)H * PM_BK * 2;
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_tiles = tiles;                                   // [num_kb][128 x 64] bf16
  uint8_t* b_tiles = tiles + (size_t)num_kb * a_bytes;        // [num_kb][H x 64] bf16
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)H) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_kb; ++s) mbar_init(&full_bar[s], 1);
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // zero the one-hot operand
  {
    uint4* az = reinterpret_cast<uint4*>(a_tiles);
    const int total = num_kb * (int)(a_bytes / 16);
    for (int i = threadIdx.x; i < total; i += PM_THREADS) az[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (threadIdx.x == 0) {  // operand B_b of this image: all K blocks in flight
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_expect_tx(&full_bar[kb], b_bytes);
      tma_load_2d(&tmap_b, &full_bar[kb], b_tiles + (size_t)kb * b_bytes, kb * PM_BK, b * H);
    }
  }
  if (threadIdx.x < PM_BM) {  // row r of the tile: pair (s, o)
    const int r = threadIdx.x;
    const int l = c + r;
    if (l < rows) {
      const int s = l / n, o = l - s * n;
      const long long t0 = p.obj_row[b];
      const float4 ps = __ldg(reinterpret_cast<const float4*>(p.pos + (t0 + s) * p.ldpos));
      const float4 po = __ldg(reinterpret_cast<const float4*>(p.pos + (t0 + o) * p.ldpos));
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s != o) {  // (self pairs keep a zero geometry, like the SIMT kernel; their table entries are the diagonal)
        const float dx = ps.x + ps.z / 2.0f - po.x - po.z / 2.0f;
        const float dy = ps.y + ps.w / 2.0f - po.y - po.w / 2.0f;
        const float dist = sqrtf(dx * dx + dy * dy);
        g.x = dist;
        g.y = asinf(dy / fmaxf(dist, 1e-10f));
        const float sx = po.x - ps.x, sy = po.y - ps.y;
        g.z = (sx > 0.0f) ? 1.0f : (sx < 0.0f ? -1.0f : 0.0f);
        g.w = (sy > 0.0f) ? 1.0f : (sy < 0.0f ? -1.0f : 0.0f);
      }
      if (p.geo_out != nullptr) p.geo_out[(long long)p.pair_row[b] + l] = g;
      const __nv_bfloat16 one = __float2bfloat16(1.0f);
      auto put = [&](int col, __nv_bfloat16 v) {
        *reinterpret_cast<__nv_bfloat16*>(a_tiles + (size_t)(col >> 6) * a_bytes + pm_swz(r, col & 63)) = v;
      };
      put(s, one);
      put(n + o, one);
      put(2 * n, __float2bfloat16(g.x));
      put(2 * n + 1, __float2bfloat16(g.y));
      put(2 * n + 2, __float2bfloat16(g.z));
      put(2 * n + 3, __float2bfloat16(g.w));
      put(2 * n + 4, one);
    }
  }
  // generic-proxy writes of the A tile must be visible to the tensor-core (async) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(H >> 3) << 17) |
                             ((uint32_t)(PM_BM >> 4) << 24);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[kb], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t da = make_smem_desc(smem_u32(a_tiles + (size_t)kb * a_bytes));
        const uint64_t db = make_smem_desc(smem_u32(b_tiles + (size_t)kb * b_bytes));
#pragma unroll
        for (int k = 0; k < PM_BK / 16; ++k) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&tmem_full_bar);
    }
  } else if (warp >= 2) {
    const int quad = warp & 3;
    const int l = c + quad * 32 + lane;
    const bool row_ok = l < rows;
    __nv_bfloat16* dst = p.h_out + ((long long)p.pair_row[b] + l) * p.ldh;
    mbar_wait(&tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int c0 = 0; c0 < H; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(trow + (uint32_t)c0, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (!row_ok) continue;
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float x0 = __uint_as_float(r[2 * j]), x1 = __uint_as_float(r[2 * j + 1]);
        x0 = x0 > 0.0f ? x0 : __expf(x0) - 1.0f;
        x1 = x1 > 0.0f ? x1 : __expf(x1) - 1.0f;
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
        pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      uint4* d4 = reinterpret_cast<uint4*>(dst + c0);
      d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    if (row_ok)
      for (int h = H; h < (int)p.ldh; ++h) dst[h] = __float2bfloat16(0.0f);  // zero K padding of the next GEMM
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
}

// B_b[h, k] (bf16, K-major, [images * H][Kp]): U_b[k][h] | V_b[k - n][h] | Wg[h][k - 2n] | bias[h] | 0.
// grid (Kp / 32, images), block (32, 8): 32 x 32 tiles transposed through shared memory (coalesced on both sides).
__global__ void __launch_bounds__(256) pair_hidden_operand_kernel(const float* __restrict__ uv, long long lduv,
                                                                  const float* __restrict__ wg, long long ldw,
                                                                  const float* __restrict__ bias, int H,
                                                                  const int32_t* __restrict__ obj_row,
                                                                  const int32_t* __restrict__ img_n,
                                                                  __nv_bfloat16* __restrict__ Bm, int Kp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.y, k0 = blockIdx.x * 32;
  const int n = img_n[b];
  const long long t0 = obj_row[b];
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int h0 = 0; h0 < H; h0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + ty + 8 * i, h = h0 + tx;
      float v = 0.0f;
      if (h < H) {
        if (k < n) v = uv[(t0 + k) * lduv + h];
        else if (k < 2 * n) v = uv[(t0 + k - n) * lduv + H + h];
        else if (k < 2 * n + 4) v = wg[(long long)h * ldw + (k - 2 * n)];
        else if (k == 2 * n + 4) v = bias[h];
      }
      tile[ty + 8 * i][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int h = h0 + ty + 8 * i, k = k0 + tx;
      if (h < H && k < Kp) Bm[((long long)b * H + h) * Kp + k] = __float2bfloat16(tile[tx][ty + 8 * i]);
    }
    __syncthreads();
  }
}

}  // namespace dfol

using namespace dfol;

extern "C" int dfol_pair_hidden_fwd_mma(const float* uv, int64_t lduv, const float* obj_pos, int64_t ldpos,
                                        const float* wg, int64_t ldw, const float* bias, void* h_out, int64_t ldh, int H,
                                        void* geo_out, const int32_t* pair_row, const int32_t* obj_row,
                                        const int32_t* img_n, int image_num, int max_n, void* operand_workspace,
                                        void* stream) {
  const char* who = "dfol_pair_hidden_fwd_mma";
  DFOL_REQUIRE(uv && obj_pos && wg && bias && h_out && pair_row && obj_row && img_n && operand_workspace,
               "%s: null pointer", who);
  if (image_num == 0) return 0;
  DFOL_REQUIRE(H >= 16 && H <= 256 && (H % 16) == 0, "%s: H must be a multiple of 16, at most 256", who);
  DFOL_REQUIRE(max_n >= 1 && 2 * max_n + 5 <= PM_BK * PM_MAX_KB, "%s: at most %d objects per image", who,
               (PM_BK * PM_MAX_KB - 5) / 2);
  DFOL_REQUIRE((ldpos % 4) == 0 && (reinterpret_cast<uintptr_t>(obj_pos) % 16) == 0 && (ldh % 8) == 0 && ldh >= H &&
                   (reinterpret_cast<uintptr_t>(h_out) % 16) == 0 &&
                   (reinterpret_cast<uintptr_t>(operand_workspace) % 16) == 0,
               "%s: alignment (positions float4, output rows 16 bytes)", who);
  DFOL_REQUIRE(image_num <= 65535, "%s: too many images", who);
  cudaStream_t st = (cudaStream_t)stream;
  const int Kp = (2 * max_n + 5 + PM_BK - 1) / PM_BK * PM_BK;
  __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(operand_workspace);
  pair_hidden_operand_kernel<<<dim3(Kp / 32, image_num), dim3(32, 8), 0, st>>>(uv, lduv, wg, ldw, bias, H, obj_row,
                                                                              img_n, Bm, Kp);
  alignas(64) CUtensorMap mb;
  int rc = encode_map_bf16(&mb, Bm, (int64_t)image_num * H, Kp, Kp, H);
  if (rc != 0) return rc;
  PmParams p;
  p.H = H; p.num_kb = Kp / PM_BK; p.pos = obj_pos; p.ldpos = ldpos;
  p.h_out = reinterpret_cast<__nv_bfloat16*>(h_out); p.ldh = ldh; p.geo_out = reinterpret_cast<float4*>(geo_out);
  p.pair_row = pair_row; p.obj_row = obj_row; p.img_n = img_n;
  const size_t smem = (size_t)p.num_kb * (PM_BM + H) * PM_BK * 2 + 1024;
  {
    cudaError_t e = cudaFuncSetAttribute(pair_hidden_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return (int)e; }
  }
  dim3 grid((max_n * max_n + PM_BM - 1) / PM_BM, image_num);
  pair_hidden_mma_kernel<<<grid, PM_THREADS, smem, st>>>(mb, p);
  return finish_launch(who);
}
