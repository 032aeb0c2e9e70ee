// Weight-gradient contraction on tcgen05: C[M,N] += A^T . B with A [K, M] and B [K, N] bf16 row-major, i.e. the
// reduction runs over ROWS (K = pair rows / object rows, 10^4 .. 10^6) and both operands are "MN-major" for the
// tensor core: dW = dZ^T . H of a Linear layer without transposing the activations.
//
//   TMA boxes of 64 k-rows x 64 elements (128 B, 128-byte swizzle); an M tile of 128 is two boxes, an N tile of BN is
//   BN/64 boxes.  The UMMA shared-memory descriptors describe the canonical MN-major layout
//   ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) [elements]: 128-byte rows are k, groups of 8 k-rows are SBO = 1024 B apart,
//   64-element blocks along M/N are LBO = 8192 B apart (one box); a_major = b_major = MN in the instruction
//   descriptor; one UMMA consumes 16 k-rows = 2 groups (+2048 B per step).
//   Split-K over blockIdx.z; the epilogue adds the fp32 partial tile into C with red.global.add.f32.
#include <cstdlib>
#include "tc_common.cuh"

namespace dfol {

constexpr int WG_BM = 128, WG_BK = 64, WG_THREADS = 192, WG_MAX_STAGES = 4;

struct WgParams {
  float* C; long long ldc;
  // optional row segments of the output: rows [i*seg_rows, (i+1)*seg_rows) go to Cseg[i] (seg_rows = 0: single C)
  float* Cseg[4]; long long ldcseg[4]; int seg_rows;
  int M, N, K;
  int BN, stages;
  int kb_per_split;  // 64-row blocks per split
};

__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address (16-byte units)
  d |= (uint64_t)(8192 >> 4) << 16;            // leading byte offset: next 64-element block along M/N
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset: next group of 8 k-rows
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(WG_THREADS) gemm_bf16_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                        const __grid_constant__ CUtensorMap tmap_b,
                                                                        WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * p.BN, m0 = blockIdx.y * WG_BM;
  const int total_kb = (p.K + WG_BK - 1) / WG_BK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int num_kb = min(p.kb_per_split, total_kb - kb0);
  const int nbox = p.BN / 64;
  const uint32_t a_bytes = 2 * 8192, b_bytes = (uint32_t)nbox * 8192;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)p.BN) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (num_kb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
        int s = 0;
        uint32_t phase = 0;   // ring slot and parity as running counters (no division on the single-thread path)
        for (int i = 0; i < num_kb; ++i) {
          mbar_wait(&empty_bar[s], phase ^ 1);
          mbar_expect_tx(&full_bar[s], stage_bytes);
          uint8_t* sa = tiles + (size_t)s * stage_bytes;
          const int krow = (kb0 + i) * WG_BK;
          tma_load_2d(&tmap_a, &full_bar[s], sa, m0, krow);
          tma_load_2d(&tmap_a, &full_bar[s], sa + 8192, m0 + 64, krow);
          for (int j = 0; j < nbox; ++j) tma_load_2d(&tmap_b, &full_bar[s], sa + a_bytes + j * 8192, n0 + 64 * j, krow);
          if (++s == p.stages) { s = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        // D = F32, A = B = BF16, both MN-major (bits 15, 16), N = BN, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(WG_BM >> 4) << 24);
        int s = 0;
        uint32_t phase = 0;
        for (int i = 0; i < num_kb; ++i) {
          mbar_wait(&full_bar[s], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(tiles + (size_t)s * stage_bytes);
          const uint64_t da = make_smem_desc_mn(sa), db = make_smem_desc_mn(sa + a_bytes);
#pragma unroll
          for (int k = 0; k < WG_BK / 16; ++k) {
            // 16 k-rows = two 8-row groups = 2048 bytes = +128 in the 16-byte start-address field
            umma_bf16(tmem_base, da + 128 * k, db + 128 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (++s == p.stages) { s = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full_bar);
      }
    } else {
      const int quad = warp & 3;
      const int m = m0 + quad * 32 + lane;
      mbar_wait(&tmem_full_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
      for (int c0 = 0; c0 < p.BN && n0 + c0 < p.N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(trow + (uint32_t)c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < p.M) {
          float* dst;
          if (p.seg_rows > 0) {
            // (no dynamic indexing of the parameter arrays: that would spill them to local memory)
            const int sg = m / p.seg_rows;
            float* cb = sg == 0 ? p.Cseg[0] : (sg == 1 ? p.Cseg[1] : p.Cseg[2]);
            const long long cl = sg == 0 ? p.ldcseg[0] : (sg == 1 ? p.ldcseg[1] : p.ldcseg[2]);
            dst = cb + (long long)(m - sg * p.seg_rows) * cl + n0 + c0;
          } else {
            dst = p.C + (long long)m * p.ldc + n0 + c0;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c0 + j < p.N) atomicAdd(dst + j, __uint_as_float(r[j]));
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

}  // namespace dfol

using namespace dfol;

static int launch_wgrad(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int M, int N,
                        int64_t K, float* const* Cseg, const int64_t* ldcseg, int seg_rows, void* stream) {
  DFOL_REQUIRE(A && B && (C || seg_rows > 0), "dfol_gemm_bf16_tc_wgrad: null pointer");
  DFOL_REQUIRE(M > 0 && N > 0 && K > 0 && K < (1ll << 31), "dfol_gemm_bf16_tc_wgrad: bad sizes");
  DFOL_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0 && lda >= M && ldb >= N,
               "dfol_gemm_bf16_tc_wgrad: lda >= M, ldb >= N, both multiples of 8 elements");
  DFOL_REQUIRE((reinterpret_cast<uintptr_t>(A) % 16) == 0 && (reinterpret_cast<uintptr_t>(B) % 16) == 0,
               "dfol_gemm_bf16_tc_wgrad: operands must be 16-byte aligned");
  WgParams p;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = (int)K;
  p.seg_rows = seg_rows;
  for (int i = 0; i < 4; ++i) {
    p.Cseg[i] = (seg_rows > 0 && i * seg_rows < M) ? Cseg[i] : nullptr;
    p.ldcseg[i] = (seg_rows > 0 && i * seg_rows < M) ? ldcseg[i] : 0;
  }
  const int n64 = (N + 63) / 64;
  const int n_tiles = (n64 + 3) / 4;
  p.BN = ((n64 + n_tiles - 1) / n_tiles) * 64;
  const int m_tiles = (M + WG_BM - 1) / WG_BM;
  const int total_kb = (int)((K + WG_BK - 1) / WG_BK);
  // about one CTA per SM: split K so that tiles * splits ~ 148
  int splits = 148 / (n_tiles * m_tiles);
  if (splits < 1) splits = 1;
  // every split ends with tile-size red.global.add traffic: keep at least min_kb K blocks of work per split
  // (12288 object rows, 300 x 256 output: 0.045 ms with 48 splits of 4 blocks, 0.034 ms with 12 splits of 16)
  static const int min_kb = [] { const char* e = getenv("DFOL_WG_MIN_KB"); return e && atoi(e) > 0 ? atoi(e) : 16; }();
  if (splits > total_kb / min_kb) splits = total_kb / min_kb;
  if (splits < 1) splits = 1;
  p.kb_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
  const int stage_bytes = 2 * 8192 + (p.BN / 64) * 8192;
  int stages = WG_MAX_STAGES;
  if (stages > p.kb_per_split) stages = p.kb_per_split;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) { set_error("dfol_gemm_bf16_tc_wgrad: %s", cudaGetErrorString(e)); return (int)e; }
  }
  // tensor maps over [K rows, M or N columns]; boxes of 64 rows x 64 columns
  alignas(64) CUtensorMap ma, mb;
  int rc = encode_map_bf16(&ma, A, K, M, lda, WG_BK);
  if (rc != 0) return rc;
  rc = encode_map_bf16(&mb, B, K, N, ldb, WG_BK);
  if (rc != 0) return rc;
  dim3 grid(n_tiles, m_tiles, splits);
  gemm_bf16_tc_wgrad_kernel<<<grid, WG_THREADS, smem, (cudaStream_t)stream>>>(ma, mb, p);
  return finish_launch("dfol_gemm_bf16_tc_wgrad");
}

extern "C" int dfol_gemm_bf16_tc_wgrad(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                                       int M, int N, int64_t K, void* stream) {
  return launch_wgrad(A, lda, B, ldb, C, ldc, M, N, K, nullptr, nullptr, 0, stream);
}

extern "C" int dfol_gemm_bf16_tc_wgrad_seg(const void* A, int64_t lda, const void* B, int64_t ldb, float* C0, int64_t ldc0,
                                           float* C1, int64_t ldc1, float* C2, int64_t ldc2, int seg_rows, int M, int N,
                                           int64_t K, void* stream) {
  DFOL_REQUIRE(seg_rows > 0 && (seg_rows % 128) == 0 && M <= 3 * seg_rows && C0 && (M <= seg_rows || C1) &&
                   (M <= 2 * seg_rows || C2),
               "dfol_gemm_bf16_tc_wgrad_seg: up to three segments of seg_rows (multiple of 128) rows each");
  float* cs[4] = {C0, C1, C2, nullptr};
  const int64_t ls[4] = {ldc0, ldc1, ldc2, 0};
  return launch_wgrad(A, lda, B, ldb, nullptr, 0, M, N, K, cs, ls, seg_rows, stream);
}
