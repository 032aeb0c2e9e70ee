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
#include <cuda.h>
#include <cuda_bf16.h>

#include "dfol_common.cuh"

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
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (long long spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spin > (1ll << 26)) __trap();  // never hang the device: a lost arrival becomes a launch error
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128-byte swizzle, rows of 128 B packed densely: stride between 8-row groups = 1024 B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);   // start address, 16-byte units
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

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
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % p.stages;
        const uint32_t phase = (kb / p.stages) & 1;
        mbar_wait(&empty_bar[s], phase ^ 1);
        mbar_expect_tx(&full_bar[s], stage_bytes);
        uint8_t* sa = tiles + (size_t)s * stage_bytes;
        tma_load_2d(&tmap_a, &full_bar[s], sa, kb * TC_BK, m0);
        tma_load_2d(&tmap_b, &full_bar[s], sa + a_bytes, kb * TC_BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = BF16, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % p.stages;
        const uint32_t phase = (kb / p.stages) & 1;
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
      for (int j = 0; j < 16; ++j) v[j] = act_fast<ACT>(__uint_as_float(r[j]) + bias_s[c0 + j]);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

}  // namespace dfol

using namespace dfol;

// 2D bf16 tensor map, 128-byte swizzle, box = box_rows x 64 elements (out-of-bounds elements read as zero).
static int encode_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld_elems, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  DFOL_REQUIRE(fn != nullptr, "dfol_gemm_bf16_tc: cuTensorMapEncodeTiled unavailable (driver)");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("dfol_gemm_bf16_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -2;
  }
  return 0;
}

extern "C" int dfol_gemm_bf16_tc(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                                 const float* bias, int M, int N, int K, int act, int out_bf16, int store,
                                 const int32_t* row_img, const int32_t* img_row, const int64_t* img_blk,
                                 const int32_t* img_stride, const int32_t* img_n, float diag_value, void* stream) {
  DFOL_REQUIRE(A && B && C, "dfol_gemm_bf16_tc: null pointer");
  DFOL_REQUIRE(M > 0 && N > 0 && K > 0 && (K % TC_BK) == 0, "dfol_gemm_bf16_tc: K must be a positive multiple of 64");
  DFOL_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0 && lda >= K && ldb >= K,
               "dfol_gemm_bf16_tc: lda/ldb must be >= K and multiples of 8 elements");
  DFOL_REQUIRE((reinterpret_cast<uintptr_t>(A) % 16) == 0 && (reinterpret_cast<uintptr_t>(B) % 16) == 0,
               "dfol_gemm_bf16_tc: operands must be 16-byte aligned");
  DFOL_REQUIRE(store == 0 || (row_img && img_row && img_blk && img_stride), "dfol_gemm_bf16_tc: table maps missing");
  DFOL_REQUIRE(store == 0 || !out_bf16, "dfol_gemm_bf16_tc: tables are fp32");
  // columns to cover: row-major outputs are zero-filled up to ldc (the next layer's K padding)
  int n_store = N;
  if (store == 0) {
    DFOL_REQUIRE(ldc >= N, "dfol_gemm_bf16_tc: ldc < N");
    DFOL_REQUIRE(!out_bf16 || (ldc % 8) == 0, "dfol_gemm_bf16_tc: bf16 output needs ldc %% 8 == 0");
    n_store = (int)ldc;
  }
  const int cover = (n_store + 15) / 16 * 16;
  const int n_tiles = (cover + 255) / 256;
  const int BN = ((cover + n_tiles - 1) / n_tiles + 15) / 16 * 16;
  TcParams p;
  p.C = C; p.ldc = ldc; p.bias = bias; p.M = M; p.N = N; p.K = K; p.n_store = n_store; p.BN = BN;
  p.act = act; p.out_bf16 = out_bf16; p.store = store;
  p.row_img = row_img; p.img_row = img_row; p.img_blk = img_blk; p.img_stride = img_stride; p.img_n = img_n;
  p.diag_value = diag_value;
  const int stage_bytes = (TC_BM + BN) * TC_BK * 2;
  int stages = (110 * 1024) / stage_bytes;  // two CTAs per SM
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < 2) stages = 2;
  if (stages > K / TC_BK) stages = K / TC_BK;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  TcKernel kernel = pick_kernel(act, store == 1 ? ST_TABLE : (out_bf16 ? ST_BF16 : ST_F32));
  {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { set_error("dfol_gemm_bf16_tc: %s", cudaGetErrorString(e)); return (int)e; }
  }
  alignas(64) CUtensorMap ma, mb;
  int rc = encode_map(&ma, A, M, K, lda, TC_BM);
  if (rc != 0) return rc;
  rc = encode_map(&mb, B, N, K, ldb, BN);
  if (rc != 0) return rc;
  dim3 grid(n_tiles, (M + TC_BM - 1) / TC_BM);
  DFOL_REQUIRE(grid.y <= 65535, "dfol_gemm_bf16_tc: M too large for one launch (%d row tiles)", (int)grid.y);
  kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(ma, mb, p);
  return finish_launch("dfol_gemm_bf16_tc");
}
