// Persistent, weights-resident bf16 tensor-core GEMM for the pair-level layers of the relation network (sm_100a):
//   C[M, N] = epilogue(A[M, K] . B[N, K]^T)   with M ~ 10^5..10^6 pair rows and a SMALL weight matrix B
//   (N <= 320, K <= 320, N*K*2 <= 160 KB).
//
// One CTA per SM, resident for the whole launch:
//   warp 0 (one lane) : TMA producer -- loads B ONCE into shared memory (all K blocks, 128B swizzle), then streams the
//                       128-row A tiles of this CTA's M tiles through a ring of 64-column K blocks
//   warp 1 (one lane) : tcgen05.mma issuer -- 128 x (N/halves) x 16 per instruction, fp32 accumulators in TMEM; when
//                       2*N <= 512 columns the accumulator is double buffered so the MMAs of tile i+1 overlap the
//                       epilogue of tile i
//   warps 2..17       : epilogue -- four warps per TMEM lane quadrant, each owning a quarter of the N columns: tcgen05.ld
//                       bias + activation (forward) or activation-derivative multiplier (dgrad), bf16 row-major store
//                       (zero K-padding for the next layer) and, for the forward layer, the demand-driven relation
//                       columns: z_j = sigmoid(row) . W_emb[wrow_j] accumulated in registers while the row passes
//                       through, combined across the four column quarters in shared memory, written as
//                       logsigmoid(z_j + b_j) straight into the compact relation table (the P x E activation is never
//                       re-read; at inference it is not even written).
// Every mbarrier wait is bounded (trap instead of hanging the GPU).
#include <stdlib.h>

#include "tc_common.cuh"

namespace dfol {

constexpr int RS_BM = 128;
constexpr int RS_BK = 64;
constexpr int RS_THREADS = 64 + 512;  // TMA warp, MMA warp, 16 epilogue warps
constexpr int RS_MAX_STAGES = 4;
constexpr int RS_MAX_KB = 5;     // K <= 320
constexpr int RS_SLOTS = 2;      // relation slots evaluated per epilogue pass
constexpr int RS_MAX_CH = 5;     // 16-column chunks per epilogue warp (320 / 4 / 16)

struct RsParams {
  void* C; long long ldc;        // bf16 row-major output (nullptr: not stored)
  const float* bias;
  int M, N, K;
  int n_store;                   // columns [N, n_store) are written as zero
  int BN;                        // accumulator columns (multiple of 64, <= 320)
  int halves;                    // MMA instructions per k step (BN > 256 -> 2)
  int stages, acc_bufs;
  int act;
  const __nv_bfloat16* mul_src; long long ld_mul; int mul_mode;
  // demand-driven relation slots (forward only; slot_wrow == nullptr: off)
  const float* W_emb; long long ldw; const float* b_emb;
  const int32_t* slot_wrow; const int32_t* img_slot; const int64_t* slot_blk; const int32_t* rel_stride;
  const int32_t* row_img; const int32_t* img_row; const int32_t* img_n;
  float* ll; float diag_value; int max_slots;
#ifdef DFOL_RS_EXPERIMENTS
  int debug;  // timing experiments only (tools/time_resident.py): 1 = no epilogue, 2 = no MMAs, 4 = no stores
#endif
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int ACT>
__device__ __forceinline__ float rs_act(float x) {
  if (ACT == DFOL_ACT_ELU) return x > 0.0f ? x : __expf(x) - 1.0f;
  if (ACT == DFOL_ACT_SIGMOID) {  // 0.5 tanh(x/2) + 0.5: one MUFU op instead of exp + reciprocal
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
  }
  return x;
}

template <int ACT, bool SLOTS>
__global__ void __launch_bounds__(RS_THREADS, 1) gemm_bf16_tc_resident_kernel(
    const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, RsParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t b_full;
  __shared__ __align__(8) uint64_t a_full[RS_MAX_STAGES];
  __shared__ __align__(8) uint64_t a_empty[RS_MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float bias_s[320];
  __shared__ float zpart[4][RS_BM][RS_SLOTS];  // slot partial sums of the four column quarters

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.K / RS_BK;
  const int num_tiles = (p.M + RS_BM - 1) / RS_BM;
  const int BNh = p.BN / p.halves;
  const uint32_t a_bytes = RS_BM * RS_BK * 2;
  const uint32_t b_kb_bytes = (uint32_t)p.BN * RS_BK * 2;  // one K block of B: BN rows x 128 B
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_tiles = base;
  uint8_t* a_tiles = base + (size_t)num_kb * b_kb_bytes;

  if (threadIdx.x == 0) {
    mbar_init(&b_full, 1);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  for (int i = threadIdx.x; i < 320; i += RS_THREADS)
    bias_s[i] = (p.bias != nullptr && i < p.N) ? __ldg(p.bias + i) : 0.0f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
      // B: every K block, `halves` boxes of BNh rows each
      mbar_expect_tx(&b_full, (uint32_t)num_kb * b_kb_bytes);
      for (int kb = 0; kb < num_kb; ++kb)
        for (int h = 0; h < p.halves; ++h)
          tma_load_2d(&tmap_b, &b_full, b_tiles + (size_t)kb * b_kb_bytes + (size_t)h * BNh * 128, kb * RS_BK,
                      h * BNh);
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t phase = (it / p.stages) & 1;
          mbar_wait(&a_empty[s], phase ^ 1);
          mbar_expect_tx(&a_full[s], a_bytes);
          tma_load_2d(&tmap_a, &a_full[s], a_tiles + (size_t)s * a_bytes, kb * RS_BK, tile * RS_BM);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BNh >> 3) << 17) |
                             ((uint32_t)(RS_BM >> 4) << 24);
      mbar_wait(&b_full, 0);
      int it = 0, local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int buf = local % p.acc_bufs;
        const uint32_t use = (uint32_t)(local / p.acc_bufs);
        mbar_wait(&acc_empty[buf], (use & 1) ^ 1);  // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + (uint32_t)(buf * p.BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t phase = (it / p.stages) & 1;
          mbar_wait(&a_full[s], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_smem_desc(smem_u32(a_tiles + (size_t)s * a_bytes));
          const uint32_t sb = smem_u32(b_tiles + (size_t)kb * b_kb_bytes);
#pragma unroll
          for (int k = 0; k < RS_BK / 16; ++k) {
            for (int h = 0; h < p.halves; ++h) {
              const uint64_t db = make_smem_desc(sb + (uint32_t)(h * BNh * 128));
#ifdef DFOL_RS_EXPERIMENTS
              if (p.debug & 2) continue;
#endif
              umma_bf16(acc + (uint32_t)(h * BNh), da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&a_empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ---------------- epilogue: warps 2..17; TMEM lane quadrant = warp % 4, column quarter = (warp - 2) / 4 ----------
    const int quad = warp & 3;
    const int cq = (warp - 2) >> 2;
    const int r_in_tile = quad * 32 + lane;
    const int cw = p.BN / 4;                      // columns per warp (multiple of 16)
    const int cbeg = cq * cw;
    const int nch = cw / 16;                      // 16-column chunks per warp (<= 5)
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int buf = local % p.acc_bufs;
      const uint32_t use = (uint32_t)(local / p.acc_bufs);
      const int m = tile * RS_BM + r_in_tile;
      const bool row_ok = m < p.M;
      // slot metadata of this row
      int Sb = 0, j0 = 0, l_img = 0, n_obj = 1;
      long long ll_base = 0, ll_stride = 0;
      if (SLOTS && row_ok) {
        const int b = __ldg(p.row_img + m);
        j0 = __ldg(p.img_slot + b);
        Sb = __ldg(p.img_slot + b + 1) - j0;
        l_img = m - __ldg(p.img_row + b);
        n_obj = __ldg(p.img_n + b);
        ll_stride = __ldg(p.rel_stride + b);
        ll_base = __ldg(p.slot_blk + b) + l_img;
      }
      // dgrad: the activation-derivative operand of this row does not depend on the accumulator: fetch it now
      uint4 hpre[RS_MAX_CH][2];
      if (!SLOTS && p.mul_mode != DFOL_MUL_NONE && row_ok) {
#pragma unroll
        for (int c = 0; c < RS_MAX_CH; ++c) {
          if (c < nch && cbeg + 16 * c + 16 <= p.N) {
            const uint4* hp = reinterpret_cast<const uint4*>(p.mul_src + (long long)m * p.ld_mul + cbeg + 16 * c);
            hpre[c][0] = __ldg(hp);
            hpre[c][1] = __ldg(hp + 1);
          }
        }
      }
      const int passes = SLOTS ? max(1, (p.max_slots + RS_SLOTS - 1) / RS_SLOTS) : 1;
      bool passes_skip = false;
      mbar_wait(&acc_full[buf], use & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * p.BN + cbeg);
#ifdef DFOL_RS_EXPERIMENTS
      if (p.debug & 1) passes_skip = true;
#endif
      for (int pass = 0; pass < passes && !passes_skip; ++pass) {
        // a later pass is skipped by warps whose rows have no slots left (warp-uniform decision)
        const bool warp_needs = pass == 0 || __any_sync(0xffffffffu, Sb > pass * RS_SLOTS);
        float z[RS_SLOTS];
        const float* wr[RS_SLOTS];
#pragma unroll
        for (int j = 0; j < RS_SLOTS; ++j) {
          const int jj = pass * RS_SLOTS + j;
          z[j] = 0.f;
          wr[j] = (SLOTS && jj < Sb) ? p.W_emb + (long long)__ldg(p.slot_wrow + j0 + jj) * p.ldw : nullptr;
        }
        if (warp_needs) {
          uint32_t r[2][16];
          tmem_ld16(trow, r[0]);
#pragma unroll
          for (int c = 0; c < RS_MAX_CH; ++c) {
            if (c < nch) {
              const int c0 = cbeg + 16 * c;
              // embedding rows of this chunk (L1-resident broadcast loads) are requested before the TMEM wait
              float4 w4[RS_SLOTS][4];
              const bool full = c0 + 16 <= p.N;
              if (SLOTS && full) {
#pragma unroll
                for (int j = 0; j < RS_SLOTS; ++j)
                  if (wr[j] != nullptr) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) w4[j][q] = __ldg(reinterpret_cast<const float4*>(wr[j] + c0) + q);
                  }
              }
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              if (c + 1 < nch) tmem_ld16(trow + (uint32_t)(16 * (c + 1)), r[(c + 1) & 1]);
              const bool computes = row_ok && (pass == 0 || c0 < p.N);
              if (computes) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = rs_act<ACT>(__uint_as_float(r[c & 1][j]) + bias_s[c0 + j]);
                if (!SLOTS && p.mul_mode != DFOL_MUL_NONE && full) {
                  const uint32_t hw[8] = {hpre[c][0].x, hpre[c][0].y, hpre[c][0].z, hpre[c][0].w,
                                          hpre[c][1].x, hpre[c][1].y, hpre[c][1].z, hpre[c][1].w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[j]));
                    if (p.mul_mode == DFOL_MUL_SIGMOID_GRAD) {
                      v[2 * j] *= h.x * (1.0f - h.x);
                      v[2 * j + 1] *= h.y * (1.0f - h.y);
                    } else {
                      v[2 * j] *= (h.x > 0.0f ? 1.0f : h.x + 1.0f);
                      v[2 * j + 1] *= (h.y > 0.0f ? 1.0f : h.y + 1.0f);
                    }
                  }
                }
                if (!full) {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (c0 + j >= p.N) v[j] = 0.0f;  // K padding of the next layer / outside the embedding width
                }
                if (SLOTS) {
#pragma unroll
                  for (int j = 0; j < RS_SLOTS; ++j) {
                    if (wr[j] != nullptr) {
                      if (full) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                          z[j] = fmaf(v[4 * q], w4[j][q].x, z[j]);
                          z[j] = fmaf(v[4 * q + 1], w4[j][q].y, z[j]);
                          z[j] = fmaf(v[4 * q + 2], w4[j][q].z, z[j]);
                          z[j] = fmaf(v[4 * q + 3], w4[j][q].w, z[j]);
                        }
                      } else {
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                          if (c0 + q < p.N) z[j] = fmaf(v[q], __ldg(wr[j] + c0 + q), z[j]);
                      }
                    }
                  }
                }
                bool do_store = pass == 0 && p.C != nullptr && c0 < p.n_store;
#ifdef DFOL_RS_EXPERIMENTS
                if (p.debug & 4) { if (v[0] == 123.456f) p.ll[0] = v[3]; do_store = false; }
#endif
                if (do_store) {
                  __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + (long long)m * p.ldc + c0;
                  if (c0 + 16 <= p.n_store) {
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
                      if (c0 + j < p.n_store) crow[j] = __float2bfloat16(v[j]);
                  }
                }
              }
            }
          }
        }
        if (SLOTS) {
          // combine the four column quarters of every row in shared memory; warp cq finalises slot cq % RS_SLOTS
#pragma unroll
          for (int j = 0; j < RS_SLOTS; ++j) zpart[cq][r_in_tile][j] = z[j];
          named_bar_sync(1 + quad, 128);
          if (cq < RS_SLOTS && row_ok) {
            const int jj = pass * RS_SLOTS + cq;
            if (jj < Sb) {
              const float x = zpart[0][r_in_tile][cq] + zpart[1][r_in_tile][cq] + zpart[2][r_in_tile][cq] +
                              zpart[3][r_in_tile][cq] + __ldg(p.b_emb + __ldg(p.slot_wrow + j0 + jj));
              const float v = fminf(x, 0.0f) - __logf(1.0f + __expf(-fabsf(x)));
              const bool is_diag = (l_img / n_obj) == (l_img % n_obj);
              p.ll[ll_base + (long long)jj * ll_stride] = is_diag ? p.diag_value : v;
            }
          }
          named_bar_sync(1 + quad, 128);  // zpart may be overwritten by the next pass / tile
        }
      }
      // this warp is done with the accumulator buffer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

typedef void (*RsKernel)(const CUtensorMap, const CUtensorMap, RsParams);

}  // namespace dfol

using namespace dfol;

static int launch_resident(const char* who, const void* A, int64_t lda, const void* B, int64_t ldb, RsParams p,
                           void* stream) {
  DFOL_REQUIRE(A && B, "%s: null pointer", who);
  DFOL_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0 && (p.K % RS_BK) == 0 && p.K <= RS_BK * RS_MAX_KB,
               "%s: K must be a multiple of 64, at most %d", who, RS_BK * RS_MAX_KB);
  DFOL_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0 && lda >= p.K && ldb >= p.K,
               "%s: lda/ldb must be >= K and multiples of 8 elements", who);
  DFOL_REQUIRE((reinterpret_cast<uintptr_t>(A) % 16) == 0 && (reinterpret_cast<uintptr_t>(B) % 16) == 0,
               "%s: operands must be 16-byte aligned", who);
  DFOL_REQUIRE(p.C == nullptr || (p.ldc >= p.n_store && (p.ldc % 8) == 0 && p.n_store >= p.N),
               "%s: N <= store_cols <= ldc, ldc %% 8 == 0", who);
  DFOL_REQUIRE(p.mul_mode == DFOL_MUL_NONE || (p.mul_src && (p.N % 16) == 0 && (p.ld_mul % 8) == 0 &&
                                               (reinterpret_cast<uintptr_t>(p.mul_src) % 16) == 0),
               "%s: multiplier needs a 16-byte aligned bf16 source, ld %% 8 == 0 and N %% 16 == 0", who);
  const int cover = ((p.C != nullptr ? p.n_store : p.N) + 63) / 64 * 64;
  DFOL_REQUIRE(cover <= 320, "%s: at most 320 output columns", who);
  p.BN = cover;
  p.halves = cover > 256 ? 2 : 1;
  DFOL_REQUIRE(((cover / p.halves) % 16) == 0 && ((cover / 4) % 16) == 0, "%s: bad column tiling", who);
  p.acc_bufs = (2 * cover <= 512) ? 2 : 1;
  const int num_kb = p.K / RS_BK;
  const size_t b_bytes = (size_t)num_kb * cover * RS_BK * 2;
  const size_t a_stage = (size_t)RS_BM * RS_BK * 2;
  const size_t budget = 220 * 1024;  // + ~4 KB static (barriers, bias, slot partials) <= 227 KB
  DFOL_REQUIRE(b_bytes + 2 * a_stage + 1024 <= budget, "%s: weight matrix does not fit in shared memory", who);
  int stages = (int)((budget - 1024 - b_bytes) / a_stage);
  if (stages > RS_MAX_STAGES) stages = RS_MAX_STAGES;
  p.stages = stages;
  const size_t smem = b_bytes + (size_t)stages * a_stage + 1024;
  RsKernel kernel;
  if (p.slot_wrow != nullptr) kernel = gemm_bf16_tc_resident_kernel<DFOL_ACT_SIGMOID, true>;
  else if (p.act == DFOL_ACT_SIGMOID) kernel = gemm_bf16_tc_resident_kernel<DFOL_ACT_SIGMOID, false>;
  else if (p.act == DFOL_ACT_ELU) kernel = gemm_bf16_tc_resident_kernel<DFOL_ACT_ELU, false>;
  else kernel = gemm_bf16_tc_resident_kernel<DFOL_ACT_NONE, false>;
  {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return (int)e; }
  }
  alignas(64) CUtensorMap ma, mb;
  int rc = encode_map_bf16(&ma, A, p.M, p.K, lda, RS_BM);
  if (rc != 0) return rc;
  rc = encode_map_bf16(&mb, B, p.N, p.K, ldb, cover / p.halves);
  if (rc != 0) return rc;
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
#ifdef DFOL_RS_EXPERIMENTS
  {
    const char* dbg = getenv("DFOL_RS_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
#endif
  const int tiles = (p.M + RS_BM - 1) / RS_BM;
  const int grid = tiles < sms ? tiles : sms;
  kernel<<<grid, RS_THREADS, smem, (cudaStream_t)stream>>>(ma, mb, p);
  return finish_launch(who);
}

extern "C" int dfol_pair_layer_fwd_tc(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                                      int store_cols, const float* bias, int M, int N, int K, int act,
                                      const float* W_emb, int64_t ldw, const float* b_emb, const int32_t* slot_wrow,
                                      const int32_t* img_slot, int max_slots, const int64_t* slot_blk,
                                      const int32_t* rel_stride, const int32_t* row_img, const int32_t* img_row,
                                      const int32_t* img_n, float diag_value, float* ll, void* stream) {
  RsParams p = {};
  p.C = C; p.ldc = ldc; p.bias = bias; p.M = M; p.N = N; p.K = K;
  p.n_store = C != nullptr ? (store_cols > 0 ? store_cols : (int)ldc) : N;
  p.act = act; p.mul_src = nullptr; p.ld_mul = 0; p.mul_mode = DFOL_MUL_NONE;
  if (slot_wrow != nullptr) {
    DFOL_REQUIRE(act == DFOL_ACT_SIGMOID, "dfol_pair_layer_fwd_tc: the slot epilogue follows a sigmoid layer");
    DFOL_REQUIRE(W_emb && b_emb && img_slot && slot_blk && rel_stride && row_img && img_row && img_n && ll,
                 "dfol_pair_layer_fwd_tc: slot tables missing");
    DFOL_REQUIRE((ldw % 4) == 0 && (reinterpret_cast<uintptr_t>(W_emb) % 16) == 0,
                 "dfol_pair_layer_fwd_tc: embedding rows must be 16-byte aligned");
  } else {
    DFOL_REQUIRE(C != nullptr, "dfol_pair_layer_fwd_tc: nothing to compute");
  }
  p.W_emb = W_emb; p.ldw = ldw; p.b_emb = b_emb; p.slot_wrow = slot_wrow; p.img_slot = img_slot;
  p.slot_blk = slot_blk; p.rel_stride = rel_stride; p.row_img = row_img; p.img_row = img_row; p.img_n = img_n;
  p.ll = ll; p.diag_value = diag_value; p.max_slots = max_slots;
  return launch_resident("dfol_pair_layer_fwd_tc", A, lda, B, ldb, p, stream);
}

extern "C" int dfol_pair_layer_dgrad_tc(const void* dZ, int64_t lddz, const void* Wt, int64_t ldwt, void* dX,
                                        int64_t lddx, int store_cols, int M, int N, int K, const void* h_saved,
                                        int64_t ldh, int mul_mode, void* stream) {
  RsParams p = {};
  DFOL_REQUIRE(dX != nullptr, "dfol_pair_layer_dgrad_tc: null output");
  p.C = dX; p.ldc = lddx; p.bias = nullptr; p.M = M; p.N = N; p.K = K;
  p.n_store = store_cols > 0 ? store_cols : (int)lddx;
  p.act = DFOL_ACT_NONE;
  p.mul_src = reinterpret_cast<const __nv_bfloat16*>(h_saved); p.ld_mul = ldh; p.mul_mode = mul_mode;
  p.slot_wrow = nullptr;
  return launch_resident("dfol_pair_layer_dgrad_tc", dZ, lddz, Wt, ldwt, p, stream);
}
