// bf16 tensor-core GEMM for the visual oracle on sm_100a: C = epilogue(A . B^T), A [M,K], B [N,K] bf16 K-major.
//
//   warp 0 (one lane)  : TMA producer  -- cp.async.bulk.tensor 2D tiles (128B swizzle) into a 2-4 stage smem ring
//   warp 1 (one lane)  : MMA issuer    -- tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16 per instruction, fp32
//                                         accumulator in TMEM; tcgen05.commit releases smem stages / signals epilogue
//   warps 2-5          : epilogue      -- tcgen05.ld (32 lanes x 16 columns), bias + activation, then either a
//                                         row-major bf16/fp32 store or the per-image transposed table store
//                                         (lanes = consecutive rows -> 128 B coalesced per table column)
// One 128 x BN output tile per CTA, BN <= 256 chosen by the host so that two CTAs fit on an SM (<= 256 TMEM
// columns and ~110 KB smem each): while one CTA drains its accumulator through the epilogue the other one issues
// MMAs.  Every mbarrier wait is bounded (trap instead of hanging the GPU).
#include <cstdlib>
#include "tc_common.cuh"

namespace dfol {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;  // 64 bf16 = 128 bytes = one swizzle atom row
constexpr int TC_UMMA_K = 16;
constexpr int TC_THREADS = 192;
constexpr int TC_MAX_STAGES = 4;

struct TcParams {
  void* C; long long ldc;
  const float* bias;
  int M, N, K;
  int n_store;   // row-major stores cover columns [0, n_store): columns >= N are written as zero (K padding)
  int BN, stages;
  int act, out_bf16, store;
  const int32_t* row_img; const int32_t* img_row; const int64_t* img_blk; const int32_t* img_stride;
  const int32_t* img_n; float diag_value;
  // optional epilogue multiplier (backward): v *= act'(h) with h = mul_src[m * ld_mul + n] (bf16), DFOL_MUL_*
  const __nv_bfloat16* mul_src; long long ld_mul; int mul_mode;
  float keep;  // < 1: mul_src holds post-dropout activations (0 / h / keep): act' at h = src * keep, factor (src != 0) / keep
  int exact;   // accurate expf / log1pf activations (fp32 parity mode on split-bf16 operands) instead of the MUFU forms
};

template <int ACT>
__device__ __forceinline__ float act_fast(float x) {
  if (ACT == DFOL_ACT_ELU) return x > 0.0f ? x : __expf(x) - 1.0f;
  if (ACT == DFOL_ACT_SIGMOID) return __fdividef(1.0f, 1.0f + __expf(-x));
  if (ACT == DFOL_ACT_LOGSIGMOID) return fminf(x, 0.0f) - __logf(1.0f + __expf(-fabsf(x)));
  return x;
}

constexpr int ST_F32 = 0, ST_BF16 = 1, ST_TABLE = 2;

// Tiles are enumerated N-fastest (blockIdx.x = N tile) so that the N tiles sharing one 128-row A tile run
// back-to-back and A is read from HBM once.
template <int ACT, int STORE>
__global__ void __launch_bounds__(TC_THREADS) gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                  const __grid_constant__ CUtensorMap tmap_b,
                                                                  TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float bias_s[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * p.BN;
  const int num_kb = p.K / TC_BK;
  const uint32_t a_bytes = TC_BM * TC_BK * 2, b_bytes = (uint32_t)p.BN * TC_BK * 2;
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
  if (warp >= 2) {  // stage the bias of this N tile (zero beyond N) while the pipeline fills
    for (int i = threadIdx.x - 64; i < p.BN; i += TC_THREADS - 64)
      bias_s[i] = (p.bias != nullptr && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.0f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
      int s = 0;
      uint32_t phase = 0;   // ring slot and parity as running counters (no division on the single-thread path)
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], phase ^ 1);
        mbar_expect_tx(&full_bar[s], stage_bytes);
        uint8_t* sa = tiles + (size_t)s * stage_bytes;
        tma_load_2d(&tmap_a, &full_bar[s], sa, kb * TC_BK, m0);
        tma_load_2d(&tmap_b, &full_bar[s], sa + a_bytes, kb * TC_BK, n0);
        if (++s == p.stages) { s = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = BF16, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      int s = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(tiles + (size_t)s * stage_bytes);
        const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + a_bytes);
#pragma unroll
        for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the 16-byte start-address field
          umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
        if (++s == p.stages) { s = 0; phase ^= 1u; }
      }
      umma_commit(&tmem_full_bar);   // accumulator complete
    }
  } else {
    // ---------------- epilogue: warps 2..5 own TMEM lane quadrants (warp % 4) ----------------
    const int quad = warp & 3;
    const int m = m0 + quad * 32 + lane;
    const bool row_ok = m < p.M;
    float* tdst = nullptr;
    long long tstride = 0;
    bool is_diag = false;
    if (STORE == ST_TABLE && row_ok) {
      const int b = p.row_img[m];
      const int l = m - p.img_row[b];
      tstride = p.img_stride[b];
      tdst = reinterpret_cast<float*>(p.C) + p.img_blk[b] + l + (long long)n0 * tstride;
      if (p.img_n != nullptr) {
        const int n_obj = p.img_n[b];
        is_diag = (l / n_obj) == (l % n_obj);
      }
    }
    const float diag = p.diag_value;
    const int n_valid = p.N - n0;        // columns of this tile that carry data
    const int n_cover = p.n_store - n0;  // columns of this tile that are stored (zero beyond n_valid)
    mbar_wait(&tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int c0 = 0; c0 < p.BN && c0 < n_cover; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(trow + (uint32_t)c0, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (!row_ok) continue;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x = __uint_as_float(r[j]) + bias_s[c0 + j];
        v[j] = p.exact ? act_apply(x, ACT) : act_fast<ACT>(x);
      }
      if (STORE != ST_TABLE && p.mul_mode != DFOL_MUL_NONE && c0 + 16 <= n_valid) {
        const uint4* hp = reinterpret_cast<const uint4*>(p.mul_src + (long long)m * p.ld_mul + n0 + c0);
        const uint4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[j]));
          if (p.keep < 1.0f) {
            const float ik = 1.0f / p.keep, hx = h.x * p.keep, hy = h.y * p.keep;
            const float gx = (p.mul_mode == DFOL_MUL_SIGMOID_GRAD) ? hx * (1.0f - hx) : (hx > 0.0f ? 1.0f : hx + 1.0f);
            const float gy = (p.mul_mode == DFOL_MUL_SIGMOID_GRAD) ? hy * (1.0f - hy) : (hy > 0.0f ? 1.0f : hy + 1.0f);
            v[2 * j] *= (h.x != 0.0f ? ik : 0.0f) * gx;
            v[2 * j + 1] *= (h.y != 0.0f ? ik : 0.0f) * gy;
          } else if (p.mul_mode == DFOL_MUL_SIGMOID_GRAD) {
            v[2 * j] *= h.x * (1.0f - h.x);
            v[2 * j + 1] *= h.y * (1.0f - h.y);
          } else {
            v[2 * j] *= (h.x > 0.0f ? 1.0f : h.x + 1.0f);
            v[2 * j + 1] *= (h.y > 0.0f ? 1.0f : h.y + 1.0f);
          }
        }
      }
      if (STORE == ST_TABLE) {
        float* d = tdst + (long long)c0 * tstride;
        if (c0 + 16 <= n_valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) d[j * tstride] = is_diag ? diag : v[j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < n_valid) d[j * tstride] = is_diag ? diag : v[j];
        }
      } else {
        if (c0 + 16 > n_valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j >= n_valid) v[j] = 0.0f;  // K padding of the next layer
        }
        if (STORE == ST_BF16) {
          __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + (long long)m * p.ldc + n0 + c0;
          if (c0 + 16 <= n_cover) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              pk[j] = *reinterpret_cast<uint32_t*>(&h2);
            }
            uint4* dst = reinterpret_cast<uint4*>(crow);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < n_cover) crow[j] = __float2bfloat16(v[j]);
          }
        } else {
          float* crow = reinterpret_cast<float*>(p.C) + (long long)m * p.ldc + n0 + c0;
          if (c0 + 16 <= n_cover && (p.ldc & 3) == 0) {
            float4* dst = reinterpret_cast<float4*>(crow);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < n_cover) crow[j] = v[j];
          }
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

typedef void (*TcKernel)(const CUtensorMap, const CUtensorMap, TcParams);

template <int STORE>
static TcKernel pick_act(int act) {
  switch (act) {
    case DFOL_ACT_ELU: return gemm_bf16_tc_kernel<DFOL_ACT_ELU, STORE>;
    case DFOL_ACT_SIGMOID: return gemm_bf16_tc_kernel<DFOL_ACT_SIGMOID, STORE>;
    case DFOL_ACT_LOGSIGMOID: return gemm_bf16_tc_kernel<DFOL_ACT_LOGSIGMOID, STORE>;
    default: return gemm_bf16_tc_kernel<DFOL_ACT_NONE, STORE>;
  }
}
static TcKernel pick_kernel(int act, int store_kind) {
  if (store_kind == ST_TABLE) return pick_act<ST_TABLE>(act);
  if (store_kind == ST_BF16) return pick_act<ST_BF16>(act);
  return pick_act<ST_F32>(act);
}

}  // namespace dfol

using namespace dfol;

static int launch_tc(const char* who, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                     const float* bias, int M, int N, int K, int act, int out_bf16, int store, const int32_t* row_img,
                     const int32_t* img_row, const int64_t* img_blk, const int32_t* img_stride, const int32_t* img_n,
                     float diag_value, const void* mul_src, int64_t ld_mul, int mul_mode, int store_cols,
                     void* stream, float keep = 1.0f, int exact = 0) {
  DFOL_REQUIRE(keep > 0.0f && keep <= 1.0f, "%s: keep = 1 - dropout p must be in (0, 1]", who);
  DFOL_REQUIRE(A && B && C, "%s: null pointer", who);
  DFOL_REQUIRE(M > 0 && N > 0 && K > 0 && (K % TC_BK) == 0, "%s: K must be a positive multiple of 64", who);
  DFOL_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0 && lda >= K && ldb >= K,
               "%s: lda/ldb must be >= K and multiples of 8 elements", who);
  DFOL_REQUIRE((reinterpret_cast<uintptr_t>(A) % 16) == 0 && (reinterpret_cast<uintptr_t>(B) % 16) == 0,
               "%s: operands must be 16-byte aligned", who);
  DFOL_REQUIRE(store == 0 || (row_img && img_row && img_blk && img_stride), "%s: table maps missing", who);
  DFOL_REQUIRE(store == 0 || !out_bf16, "%s: tables are fp32", who);
  DFOL_REQUIRE(mul_mode == DFOL_MUL_NONE ||
                   (mul_src && store == 0 && (N % 16) == 0 && (ld_mul % 8) == 0 &&
                    (reinterpret_cast<uintptr_t>(mul_src) % 16) == 0),
               "%s: multiplier needs a 16-byte aligned bf16 source, ld %% 8 == 0 and N %% 16 == 0", who);
  // columns to cover: row-major outputs are zero-filled up to ldc (the next layer's K padding)
  int n_store = N;
  if (store == 0) {
    DFOL_REQUIRE(ldc >= N, "%s: ldc < N", who);
    DFOL_REQUIRE(!out_bf16 || (ldc % 8) == 0, "%s: bf16 output needs ldc %% 8 == 0", who);
    n_store = store_cols > 0 ? store_cols : (int)ldc;
    DFOL_REQUIRE(n_store >= N && n_store <= ldc, "%s: N <= store_cols <= ldc", who);
  }
  const int cover = (n_store + 15) / 16 * 16;
  int n_tiles = (cover + 255) / 256;
  int BN = ((cover + n_tiles - 1) / n_tiles + 15) / 16 * 16;
  {
    // wave quantisation: M / 128 row tiles x n_tiles column tiles on 2 x 148 CTA slots.  One more column tile (narrower
    // BN) can turn "one full wave + a 30 % one" into two nearly full waves of cheaper tiles: cost ~ waves x (BN + 64)
    // (the 64 stands for the A tile and the fixed part of a tile).  DFOL_TC_WAVE=0 keeps the widest tiles.
    static const int wave = [] { const char* e = getenv("DFOL_TC_WAVE"); return e ? atoi(e) : 1; }();
    const long long m_tiles = (M + TC_BM - 1) / TC_BM;
    const long long slots = 2 * 148;
    if (wave && store == 0 && m_tiles * n_tiles > slots) {
      const int n2 = n_tiles + 1;
      const int BN2 = ((cover + n2 - 1) / n2 + 15) / 16 * 16;
      const long long c1 = ((m_tiles * n_tiles + slots - 1) / slots) * (BN + 64);
      const long long c2 = ((m_tiles * n2 + slots - 1) / slots) * (BN2 + 64);
      if (BN2 >= 64 && c2 < c1) { n_tiles = n2; BN = BN2; }
    }
  }
  TcParams p;
  p.C = C; p.ldc = ldc; p.bias = bias; p.M = M; p.N = N; p.K = K; p.n_store = n_store; p.BN = BN;
  p.act = act; p.out_bf16 = out_bf16; p.store = store;
  p.row_img = row_img; p.img_row = img_row; p.img_blk = img_blk; p.img_stride = img_stride; p.img_n = img_n;
  p.diag_value = diag_value;
  p.mul_src = reinterpret_cast<const __nv_bfloat16*>(mul_src); p.ld_mul = ld_mul; p.mul_mode = mul_mode;
  p.keep = keep;
  p.exact = exact;
  const int stage_bytes = (TC_BM + BN) * TC_BK * 2;
  // two CTAs per SM (DFOL_TC_SMEM_KB: ring budget per CTA for experiments; 200 = one CTA per SM with a deep ring)
  static const int smem_kb = [] { const char* e = getenv("DFOL_TC_SMEM_KB"); return e && atoi(e) >= 48 ? atoi(e) : 110; }();
  int stages = (smem_kb * 1024) / stage_bytes;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < 2) stages = 2;
  if (stages > K / TC_BK) stages = K / TC_BK;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  TcKernel kernel = pick_kernel(act, store == 1 ? ST_TABLE : (out_bf16 ? ST_BF16 : ST_F32));
  {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return (int)e; }
  }
  alignas(64) CUtensorMap ma, mb;
  int rc = encode_map_bf16(&ma, A, M, K, lda, TC_BM);
  if (rc != 0) return rc;
  rc = encode_map_bf16(&mb, B, N, K, ldb, BN);
  if (rc != 0) return rc;
  dim3 grid(n_tiles, (M + TC_BM - 1) / TC_BM);
  DFOL_REQUIRE(grid.y <= 65535, "%s: M too large for one launch (%d row tiles)", who, (int)grid.y);
  kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(ma, mb, p);
  return finish_launch(who);
}

extern "C" int dfol_gemm_bf16_tc(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                                 const float* bias, int M, int N, int K, int act, int out_bf16, int store,
                                 const int32_t* row_img, const int32_t* img_row, const int64_t* img_blk,
                                 const int32_t* img_stride, const int32_t* img_n, float diag_value, void* stream) {
  return launch_tc("dfol_gemm_bf16_tc", A, lda, B, ldb, C, ldc, bias, M, N, K, act, out_bf16, store, row_img, img_row,
                   img_blk, img_stride, img_n, diag_value, nullptr, 0, DFOL_MUL_NONE, 0, stream);
}

extern "C" int dfol_gemm_bf16_tc_dgrad(const void* dZ, int64_t lddz, const void* Wt, int64_t ldwt, void* dX,
                                       int64_t lddx, int store_cols, int M, int N, int K, const void* h_saved,
                                       int64_t ldh, int mul_mode, float keep, void* stream) {
  return launch_tc("dfol_gemm_bf16_tc_dgrad", dZ, lddz, Wt, ldwt, dX, lddx, nullptr, M, N, K, DFOL_ACT_NONE, 1, 0,
                   nullptr, nullptr, nullptr, nullptr, nullptr, 0.0f, h_saved, ldh, mul_mode, store_cols, stream, keep);
}

// ---------------------------------------------------------------------------------------------------------
// fp32 parity mode on the tensor cores: an fp32 operand x is split into three bf16 parts x = h + m + l
// (h = bf16(x), m = bf16(x - h), l = bf16(x - h - m): 24 mantissa bits) and the product a . b is evaluated as the six
// terms hl + lh + mm + hm + mh + hh (the dropped ml + lm + ll are below 2^-24 relative) by ONE bf16 GEMM over operands
// concatenated along K: A' = [Ah | Al | Am | Ah | Am | Ah], B' = [Bl | Bh | Bm | Bm | Bh | Bh], fp32 accumulation in
// TMEM.  The same tcgen05 kernels as the bf16 mode, six times the MMA work, fp32-level results.
namespace dfol {
// Block order = ascending magnitude of the product terms: (hl, lh, mm) ~ 2^-16, (hm, mh) ~ 2^-8, hh last.  The tensor
// core's fp32 accumulation truncates at the magnitude of the running sum: adding the small terms FIRST keeps their
// low bits (measured: hh first leaves a relative bias of ~1.4e-5 at K = 2048, small terms first ~3e-6).
//   pattern 0 (A side): h l m h m h      pattern 1 (B side): l h m m h h
template <int STACKED>
__global__ void __launch_bounds__(256) split3_bf16_kernel(const float* __restrict__ src, long long lds, long long rows,
                                                          int cols, __nv_bfloat16* __restrict__ dst, long long ldd,
                                                          int Kp, int pattern) {
  // one thread = 8 consecutive columns of one row: 16-byte stores into each of the six blocks
  const int width = STACKED ? (int)ldd : Kp;          // columns covered per block (multiple of 8)
  const int groups = width / 8;
  const long long total = rows * (long long)groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / groups;
    const int c0 = (int)(idx - r * groups) * 8;
    float x[8];
    const float* sp = src + r * lds + c0;
    if (c0 + 8 <= cols && ((reinterpret_cast<uintptr_t>(sp) & 15) == 0)) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(sp)), b = __ldg(reinterpret_cast<const float4*>(sp) + 1);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (c0 + j < cols) ? __ldg(sp + j) : 0.0f;
    }
    uint32_t hp[4], mp[4], lp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
      const float2 hf = __bfloat1622float2(h);
      const float r0 = x[2 * j] - hf.x, r1 = x[2 * j + 1] - hf.y;
      const __nv_bfloat162 m = __floats2bfloat162_rn(r0, r1);
      const float2 mf = __bfloat1622float2(m);
      const __nv_bfloat162 l = __floats2bfloat162_rn(r0 - mf.x, r1 - mf.y);
      hp[j] = *reinterpret_cast<const uint32_t*>(&h);
      mp[j] = *reinterpret_cast<const uint32_t*>(&m);
      lp[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    const uint4 H = make_uint4(hp[0], hp[1], hp[2], hp[3]), Mv = make_uint4(mp[0], mp[1], mp[2], mp[3]),
                L = make_uint4(lp[0], lp[1], lp[2], lp[3]);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      // A side: h l m h m h ; B side: l h m m h h
      const uint4 v = pattern == 0 ? (k == 1 ? L : ((k == 2 || k == 4) ? Mv : H))
                                   : (k == 0 ? L : ((k == 2 || k == 3) ? Mv : H));
      __nv_bfloat16* dp = STACKED ? dst + (k * rows + r) * ldd + c0 : dst + r * ldd + (long long)k * Kp + c0;
      *reinterpret_cast<uint4*>(dp) = v;
    }
  }
}
}  // namespace dfol

extern "C" int dfol_split3_bf16(const float* src, int64_t lds, int64_t rows, int cols, void* dst, int64_t ldd, int Kp,
                                int pattern, int stacked, void* stream) {
  DFOL_REQUIRE(src && dst && rows > 0 && cols > 0, "dfol_split3_bf16: bad arguments");
  DFOL_REQUIRE(stacked ? (ldd >= cols && (ldd % 8) == 0) : (Kp >= cols && (Kp % 8) == 0 && ldd >= 6ll * Kp && (ldd % 8) == 0),
               "dfol_split3_bf16: destination too narrow or not a multiple of 8 columns");
  DFOL_REQUIRE((reinterpret_cast<uintptr_t>(dst) % 16) == 0, "dfol_split3_bf16: destination must be 16-byte aligned");
  const long long total = rows * (long long)((stacked ? ldd : Kp) / 8);
  const long long want = (total + 255) / 256;
  const int blocks = (int)(want < 148 * 32 ? want : 148 * 32);
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
  if (stacked) split3_bf16_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, lds, rows, cols, d, ldd, Kp, pattern);
  else split3_bf16_kernel<0><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, lds, rows, cols, d, ldd, Kp, pattern);
  return finish_launch("dfol_split3_bf16");
}

extern "C" int dfol_gemm_bf16_tc_exact(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                                       int store_cols, const float* bias, int M, int N, int K, int act, int store,
                                       const int32_t* row_img, const int32_t* img_row, const int64_t* img_blk,
                                       const int32_t* img_stride, const int32_t* img_n, float diag_value, void* stream) {
  return launch_tc("dfol_gemm_bf16_tc_exact", A, lda, B, ldb, C, ldc, bias, M, N, K, act, 0, store, row_img, img_row,
                   img_blk, img_stride, img_n, diag_value, nullptr, 0, DFOL_MUL_NONE, store_cols, stream, 1.0f, 1);
}
