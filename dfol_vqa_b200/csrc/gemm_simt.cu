// fp32 SIMT GEMM with the fused epilogues of the visual oracle (parity path; see include/dfol_b200.h).
//
// 128x64 output tile per 256-thread block, BK = 16, each thread an 8x4 micro tile (two float4 of rows, one
// float4 of columns), register-prefetched global loads, generic operand strides so the same kernel serves
// forward (A.B^T), data-gradient (A.B) and weight-gradient (A^T.B, split-K + atomics) contractions.
#include "dfol_common.cuh"

namespace dfol {

constexpr int BM = 128, BN = 64, BK = 16, GEMM_THREADS = 256;
constexpr int AS_LD = BM + 4, BS_LD = BN + 4;

struct GemmParams {
  const float* A; long long sam, sak;
  const float* B; long long sbk, sbn;
  float* C; long long ldc;
  const float* bias;
  int M, N, K;
  int act, accumulate, split_k;
  const float* mul_src; long long ld_mul; int mul_mode;
  int store;
  const int32_t* row_img; const int32_t* img_row; const int64_t* img_blk; const int32_t* img_stride;
  const int32_t* img_n; float diag_value;
};

// ROWFAST: consecutive threads own consecutive row groups (coalesced transposed table stores);
// otherwise consecutive threads own consecutive column groups (coalesced row-major stores).
template <bool ROWFAST, bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_f32_kernel(GemmParams p) {
  __shared__ __align__(16) float As[2][BK][AS_LD];
  __shared__ __align__(16) float Bs[2][BK][BS_LD];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tr = ROWFAST ? (tid & 15) : (tid >> 4);
  const int tc = ROWFAST ? (tid >> 4) : (tid & 15);

  // K range of this split
  const int k_per = ((p.K + p.split_k - 1) / p.split_k + BK - 1) / BK * BK;
  const int k_begin = blockIdx.z * k_per;
  const int k_end = min(p.K, k_begin + k_per);
  if (k_begin >= k_end && p.split_k > 1) return;

  // global -> register staging: A tile 128x16 = 8 per thread, B tile 64x16 = 4 per thread
  float ra[8], rb[4];
  auto load_tiles = [&](int kb) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m, k;
      if (A_KCONTIG) { k = tid & 15; m = (tid >> 4) + 16 * i; } else { m = tid & 127; k = (tid >> 7) + 2 * i; }
      int gm = m0 + m, gk = kb + k;
      ra[i] = (gm < p.M && gk < k_end) ? __ldg(p.A + (long long)gm * p.sam + (long long)gk * p.sak) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n, k;
      if (B_KCONTIG) { k = tid & 15; n = (tid >> 4) + 16 * i; } else { n = tid & 63; k = (tid >> 6) + 4 * i; }
      int gn = n0 + n, gk = kb + k;
      rb[i] = (gn < p.N && gk < k_end) ? __ldg(p.B + (long long)gk * p.sbk + (long long)gn * p.sbn) : 0.0f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m, k;
      if (A_KCONTIG) { k = tid & 15; m = (tid >> 4) + 16 * i; } else { m = tid & 127; k = (tid >> 7) + 2 * i; }
      As[buf][k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n, k;
      if (B_KCONTIG) { k = tid & 15; n = (tid >> 4) + 16 * i; } else { n = tid & 63; k = (tid >> 6) + 4 * i; }
      Bs[buf][k][n] = rb[i];
    }
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  int buf = 0;
  load_tiles(k_begin);
  store_tiles(0);
  __syncthreads();
  for (int kb = k_begin; kb < k_end; kb += BK) {
    const bool more = kb + BK < k_end;
    if (more) load_tiles(kb + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][4 * tr]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + 4 * tr]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][4 * tc]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? 4 * tr + i : 64 + 4 * tr + (i - 4));
    if (m >= p.M) continue;
    long long tbase = 0;
    int tstride = 0;
    bool is_diag = false;
    if (p.store == 1) {
      const int b = p.row_img[m];
      const int l = m - p.img_row[b];
      tbase = p.img_blk[b] + l;
      tstride = p.img_stride[b];
      if (p.img_n != nullptr) {
        const int n_obj = p.img_n[b];
        is_diag = (l / n_obj) == (l % n_obj);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + 4 * tc + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.split_k > 1) {
        atomicAdd(p.C + (long long)m * p.ldc + n, v);
        continue;
      }
      if (p.bias != nullptr) v += p.bias[n];
      v = act_apply(v, p.act);
      if (p.mul_mode != DFOL_MUL_NONE) {
        const float h = p.mul_src[(long long)m * p.ld_mul + n];
        v *= (p.mul_mode == DFOL_MUL_SIGMOID_GRAD) ? h * (1.0f - h) : (h > 0.0f ? 1.0f : h + 1.0f);
      }
      if (p.store == 1) {
        p.C[tbase + (long long)n * tstride] = is_diag ? p.diag_value : v;
      } else {
        float* dst = p.C + (long long)m * p.ldc + n;
        *dst = p.accumulate ? (*dst + v) : v;
      }
    }
  }
}

template <bool ROWFAST>
static void launch_gemm(const GemmParams& p, dim3 grid, cudaStream_t s) {
  const bool ak = (p.sak == 1), bk = (p.sbk == 1);
  if (ak && bk) gemm_f32_kernel<ROWFAST, true, true><<<grid, GEMM_THREADS, 0, s>>>(p);
  else if (ak && !bk) gemm_f32_kernel<ROWFAST, true, false><<<grid, GEMM_THREADS, 0, s>>>(p);
  else if (!ak && bk) gemm_f32_kernel<ROWFAST, false, true><<<grid, GEMM_THREADS, 0, s>>>(p);
  else gemm_f32_kernel<ROWFAST, false, false><<<grid, GEMM_THREADS, 0, s>>>(p);
}

}  // namespace dfol

extern "C" int dfol_gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
                             float* C, int64_t ldc, const float* bias, int M, int N, int K, int act, int accumulate,
                             int split_k, const float* mul_src, int64_t ld_mul, int mul_mode, int store,
                             const int32_t* row_img, const int32_t* img_row, const int64_t* img_blk,
                             const int32_t* img_stride, const int32_t* img_n, float diag_value, void* stream) {
  using namespace dfol;
  DFOL_REQUIRE(A && B && C, "dfol_gemm_f32: null operand");
  DFOL_REQUIRE(M >= 0 && N >= 0 && K >= 0, "dfol_gemm_f32: negative size");
  if (M == 0 || N == 0) return 0;
  if (split_k < 1) split_k = 1;
  DFOL_REQUIRE(split_k == 1 || (bias == nullptr && act == DFOL_ACT_NONE && mul_mode == DFOL_MUL_NONE && store == 0),
               "dfol_gemm_f32: split-K excludes bias/activation/multiplier/table store");
  DFOL_REQUIRE(store == 0 || (row_img && img_row && img_blk && img_stride), "dfol_gemm_f32: table store needs maps");
  DFOL_REQUIRE(mul_mode == DFOL_MUL_NONE || mul_src != nullptr, "dfol_gemm_f32: multiplier source missing");
  GemmParams p{A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, act, accumulate, split_k, mul_src, ld_mul, mul_mode,
               store, row_img, img_row, img_blk, img_stride, img_n, diag_value};
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, split_k);
  DFOL_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "dfol_gemm_f32: grid too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (store == 1) launch_gemm<true>(p, grid, s);
  else launch_gemm<false>(p, grid, s);
  return finish_launch("dfol_gemm_f32");
}
