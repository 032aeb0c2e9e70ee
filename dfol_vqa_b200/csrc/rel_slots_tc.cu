// Demand-driven relation table on the tensor cores (sm_100a): the same result as rel_slots_fwd_kernel
//   LL[b][slot j][l] = logsigmoid(H2[row0[b] + l, :] . W[wrow_j, :] + bias[wrow_j]),   self pairs = diag
// evaluated as a GROUPED GEMM: image b multiplies its n_b^2 activation rows by ITS OWN <= 16 embedding rows.
//
// The SIMT kernel pays one dot product + one warp reduction per (row, slot): with 9 relation columns per image
// (relation-chain programs, N = 100) it is issue-bound at 1.3 TB/s and needs two passes over the 1.6 GB activation.
// Here a CTA owns one 128-row tile of one image: all K blocks of the tile are fetched by TMA at once (80 KB in flight per
// CTA, two CTAs per SM), one elected thread issues 128 x 16 x 16 tcgen05.mma into a 32-column TMEM accumulator, and four
// epilogue warps (lane = row) add the bias, take the log-sigmoid and write the slot columns -- consecutive lanes write
// consecutive table entries.  The activation is read once, whatever the number of slots (<= 16 per pass).
//
// B operand: dfol_rel_slot_weights gathers, per image, its slot rows of the fp32 embedding matrix into a bf16
// [16 * images][K] matrix (zero rows beyond the image's slot count), so that B tiles are plain TMA boxes.
#include "tc_common.cuh"

namespace dfol {

constexpr int RS_BM = 128;
constexpr int RS_BK = 64;
constexpr int RS_BN = 16;      // slots per image and pass
constexpr int RS_MAX_KB = 5;   // K <= 320
constexpr int RS_THREADS = 192;

struct RsParams {
  int K, first;
  const float* bias;
  const int32_t* slot_wrow; const int32_t* img_slot; const int64_t* slot_blk; const int32_t* stride;
  const int32_t* row0; const int32_t* img_rows; const int32_t* img_n;
  float diag; float* ll;
  float* pp;  // optional probability table e^{ll} (0 on self pairs), same layout
};

__global__ void __launch_bounds__(RS_THREADS) rel_slots_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                  const __grid_constant__ CUtensorMap tmap_b,
                                                                  RsParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[RS_MAX_KB];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int b = blockIdx.y;
  const int rows = p.img_rows[b];
  const int c = blockIdx.x * RS_BM;
  const int j0 = p.img_slot[b] + p.first;
  const int Sb = min(p.img_slot[b + 1] - j0, RS_BN);
  if (c >= rows || Sb <= 0) return;  // (whole CTA: before any barrier / TMEM allocation)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = p.row0[b] + c;
  const int num_kb = p.K / RS_BK;
  constexpr uint32_t a_bytes = RS_BM * RS_BK * 2, b_bytes = RS_BN * RS_BK * 2, stage_bytes = a_bytes + b_bytes;
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  if (threadIdx.x == 0) {
    for (int s = 0; s < num_kb; ++s) mbar_init(&full_bar[s], 1);
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(32u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {  // every K block has its own stage: all loads of the tile are in flight at once
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_expect_tx(&full_bar[kb], stage_bytes);
        uint8_t* sa = tiles + (size_t)kb * stage_bytes;
        tma_load_2d(&tmap_a, &full_bar[kb], sa, kb * RS_BK, m0);
        tma_load_2d(&tmap_b, &full_bar[kb], sa + a_bytes, kb * RS_BK, b * RS_BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = BF16, both K-major, N = 16, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(RS_BN >> 3) << 17) |
                             ((uint32_t)(RS_BM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[kb], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(tiles + (size_t)kb * stage_bytes);
        const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + a_bytes);
#pragma unroll
        for (int k = 0; k < RS_BK / 16; ++k) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&tmem_full_bar);
    }
  } else {
    // epilogue: warps 2..5 own TMEM lane quadrants (warp % 4); lane = row of the tile
    const int quad = warp & 3;
    const int l = c + quad * 32 + lane;  // row inside the image
    const bool row_ok = l < rows;
    const int n_obj = p.img_n[b];
    const bool is_diag = row_ok && (l / n_obj) == (l % n_obj);
    float bj[RS_BN];
#pragma unroll
    for (int j = 0; j < RS_BN; ++j) bj[j] = (j < Sb) ? __ldg(p.bias + __ldg(p.slot_wrow + j0 + j)) : 0.0f;
    const long long st = p.stride[b];
    const long long doff = p.slot_blk[b] + (long long)p.first * st + l;
    float* dst = p.ll + doff;
    mbar_wait(&tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16), r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row_ok) {
#pragma unroll
      for (int j = 0; j < RS_BN; ++j) {
        if (j < Sb) {
          const float x = __uint_as_float(r[j]) + bj[j];
          const float e = __expf(-fabsf(x));
          dst[(long long)j * st] = is_diag ? p.diag : fminf(x, 0.0f) - __logf(1.0f + e);
          if (p.pp != nullptr) p.pp[doff + (long long)j * st] = is_diag ? 0.0f : __fdividef(x >= 0.0f ? 1.0f : e, 1.0f + e);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32u));
}

// Wb[(16 b + j), :] = bf16(W[slot_wrow[img_slot[b] + first + j], :]) for j < slots of image b (zero rows beyond, zero K
// padding): the per-image B operands of the grouped GEMM.
__global__ void __launch_bounds__(256) rel_slot_weights_kernel(const float* __restrict__ W, long long ldw, int E,
                                                               const int32_t* __restrict__ slot_wrow,
                                                               const int32_t* __restrict__ img_slot, int first,
                                                               __nv_bfloat16* __restrict__ Wb, int K) {
  const int b = blockIdx.x;
  const int j0 = img_slot[b] + first;
  const int Sb = min(img_slot[b + 1] - j0, RS_BN);
  for (int idx = threadIdx.x; idx < RS_BN * K; idx += blockDim.x) {
    const int j = idx / K, e = idx - j * K;
    float v = 0.0f;
    if (j < Sb && e < E) v = W[(long long)slot_wrow[j0 + j] * ldw + e];
    Wb[((long long)b * RS_BN + j) * K + e] = __float2bfloat16(v);
  }
}

}  // namespace dfol

using namespace dfol;

extern "C" int dfol_rel_slots_fwd_tc(const void* h_saved, int64_t ldh, int64_t total_rows, int E, int K, const float* W,
                                     int64_t ldw, const float* bias, const int32_t* slot_wrow, const int32_t* img_slot,
                                     int max_slots, const int64_t* slot_blk, const int32_t* stride, const int32_t* row0,
                                     const int32_t* img_rows, const int32_t* img_n, int image_num, int max_rows,
                                     float diag_value, void* wb_workspace, float* ll, float* p_out, void* stream) {
  const char* who = "dfol_rel_slots_fwd_tc";
  DFOL_REQUIRE(h_saved && W && bias && slot_wrow && img_slot && slot_blk && stride && row0 && img_rows && img_n && ll &&
                   wb_workspace,
               "%s: null pointer", who);
  if (image_num == 0 || max_rows == 0 || max_slots == 0) return 0;
  DFOL_REQUIRE(K > 0 && (K % RS_BK) == 0 && K <= RS_BK * RS_MAX_KB && E <= K && ldh >= K && (ldh % 8) == 0,
               "%s: K must be a multiple of 64, at most %d, E <= K <= ldh, ldh %% 8 == 0", who, RS_BK * RS_MAX_KB);
  DFOL_REQUIRE((reinterpret_cast<uintptr_t>(h_saved) % 16) == 0 && (reinterpret_cast<uintptr_t>(wb_workspace) % 16) == 0,
               "%s: operands must be 16-byte aligned", who);
  DFOL_REQUIRE(image_num <= 65535, "%s: too many images", who);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(wb_workspace);
  const size_t smem = (size_t)(K / RS_BK) * (RS_BM + RS_BN) * RS_BK * 2 + 1024;
  {
    cudaError_t e = cudaFuncSetAttribute(rel_slots_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return (int)e; }
  }
  alignas(64) CUtensorMap ma, mb;
  int rc = encode_map_bf16(&ma, h_saved, total_rows, K, ldh, RS_BM);
  if (rc != 0) return rc;
  rc = encode_map_bf16(&mb, wb, (int64_t)image_num * RS_BN, K, K, RS_BN);
  if (rc != 0) return rc;
  for (int first = 0; first < max_slots; first += RS_BN) {
    rel_slot_weights_kernel<<<image_num, 256, 0, st>>>(W, ldw, E, slot_wrow, img_slot, first, wb, K);
    RsParams p;
    p.K = K; p.first = first; p.bias = bias; p.slot_wrow = slot_wrow; p.img_slot = img_slot; p.slot_blk = slot_blk;
    p.stride = stride; p.row0 = row0; p.img_rows = img_rows; p.img_n = img_n; p.diag = diag_value; p.ll = ll; p.pp = p_out;
    dim3 grid((max_rows + RS_BM - 1) / RS_BM, image_num);
    rel_slots_tc_kernel<<<grid, RS_THREADS, smem, st>>>(ma, mb, p);
  }
  return finish_launch(who);
}
