// Cluster-split, weights-resident bf16 tensor-core GEMM for the pair-level layers (sm_100a):
//   C[M, N] = epilogue(A[M, K] . B[N, K]^T),  M ~ 10^5..10^6 pair rows, small weight matrix (N <= 384, K <= 320).
//
// The CTAs of a thread-block cluster (two in practice; the code handles 1-4) sit on different SMs and split N: CTA r
// keeps rows [r*BNh, (r+1)*BNh) of B resident in shared memory (<= 96 KB instead of the whole matrix), which leaves
// room for the A ring and for TMA-staged epilogue tiles.  Per CTA:
//   warp 0 (one lane) : TMA producer.  Every 128 x 64 A block is fetched ONCE per cluster: its two 64-row boxes are
//                       dealt round-robin to the CTAs, each loaded with .multicast::cluster into all rings; the
//                       epilogue operand tile (saved activation of the dgrad) is a plain TMA load of this CTA's
//                       columns.
//   warp 1 (one lane) : tcgen05.mma issuer, 128 x BNh x 16, fp32 accumulators double buffered in TMEM;
//                       tcgen05.commit.multicast releases a ring stage in ALL CTAs of the cluster.
//   warps 2..9        : epilogue (8 warps: 4 TMEM lane quadrants x 2 column parts; DFOL_CL_EPI_WARPS): tcgen05.ld, bias + activation or activation-derivative multiplier (operand read from
//                       the 128B-swizzled shared tile, conflict free), bf16 result written to a swizzled shared tile
//                       -- no per-thread global access.
//   warp 10 (one lane): store warp: waits until the eight epilogue warps have published the tile, issues the
//                       cp.async.bulk.tensor store (full 128-byte lines) and hands the buffer back once the store has
//                       read it.  (Clock stamps, DFOL_CL_TRACE=1: with the store issued by an epilogue thread, the ~1500
//                       cycles a bulk store needs to read 32-48 KB and a 256-thread barrier sat inside every tile's
//                       epilogue, which is the role that paces these kernels.)
// Fused weight gradient (WG, dgrad only): the dgrad already holds both operands of dW = dZ^T . H of its layer in shared
// memory -- every 128 x 64 block of dZ passes through the A ring and this CTA's columns of the saved activation H sit
// in the epilogue operand tile.  Per ring stage the MMA warp issues, after the four K-major dgrad MMAs, eight MN-major
// MMAs (reduction over the tile's 128 rows) dWt[this CTA's H columns, the stage's 64 dZ columns] += H^T . dZ into a
// second TMEM accumulator that lives for the whole kernel (K columns next to a single-buffered dgrad accumulator, which
// the epilogue copies to registers and releases at once); one red.global.add pass per CTA at the end.  The separate
// wgrad launch, which re-read dZ and H from HBM, disappears.
// Every mbarrier wait is bounded (trap instead of hanging the GPU).
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"

namespace dfol {

constexpr int CL_BM = 128;
constexpr int CL_BK = 64;
#ifndef DFOL_CL_EPI_WARPS
#define DFOL_CL_EPI_WARPS 8   // (16 measured level: forward 0.173 against 0.179 ms at c1 size, dgrad 0.310 against 0.298)
#endif
constexpr int CL_EW = DFOL_CL_EPI_WARPS;    // epilogue warps: 4 TMEM lane quadrants x CL_EW / 4 column parts
constexpr int CL_CP = CL_EW / 4;            // column parts of a CTA's BNh columns
constexpr int CL_THREADS = 64 + 32 * CL_EW + 32;   // producer, MMA issuer, epilogue warps, store warp
constexpr int CL_MAX_STAGES = 8;
constexpr int CL_MAX_KB = 5;
constexpr int CL_ABOX = 2;        // row boxes per A block, dealt round-robin to the CTAs of the cluster
constexpr int CL_EC = 3;          // in-place operand/result tiles of the dgrad (operand prefetched two tiles ahead)
constexpr int CL_MAX_CH = 192 / CL_CP / 16;   // 16-column chunks per epilogue warp
constexpr int CL_MAX_BOX = 3;     // 64-column boxes per staging tile (BNh <= 192)
constexpr int CL_WG_CH = 128 / CL_CP / 16;   // the same with the fused weight gradient (BNh = 128: held in registers)

struct ClParams {
  const float* bias;
  int M, N, K;
  int BNh;              // columns per CTA (multiple of 64)
  int stages;
  int act, mul_mode;    // mul_mode != NONE: epilogue operand tile (bf16, same shape as C) is loaded through tmap_e
  int cluster_size;     // CTAs per cluster (the launch attribute)
  int n_store;          // stored width of C (64-column boxes entirely beyond it are skipped)
  float keep;           // < 1: the multiplier operand holds post-dropout activations (0 / h / keep)
  int prefetch;         // tiles of L2 prefetch distance (0 = off)
  float* dW;            // WG: C_w[K, N] += A^T . E (fp32, red.global.add), row stride lddw
  long long lddw;
  int K_real;           // WG: rows of dW (columns of A that carry a weight row)
  int ec;               // in-place operand/result tiles of the dgrad in use (2..CL_EC)
  int mc;               // 1: every A block is fetched once per cluster (multicast); 0: every CTA loads its own copy
  long long* trace;     // DFOL_CL_TRACE=1: clock64 stamps of block 0's roles, first 16 tiles ([role: load, mma, epilogue, store][tile][16])
  int debug;            // DFOL_CL_DEBUG ablation bits (timing experiments only; results are wrong when set)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  // non-.aligned forms: the single-lane producer / MMA roles leave their warps diverged
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void cl_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cl_named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void cl_prefetch_l2(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}

// MN-major operand, 128-byte swizzle: rows of 128 B are the reduction index (groups of 8 rows 1024 B apart), 64-element
// blocks along M/N are `lbo` bytes apart (gemm_tcgen05_wgrad.cu has the layout in full)
__device__ __forceinline__ uint64_t cl_smem_desc_mn(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// byte offset of the 16-byte piece `piece` (0..7) of row `row` inside a [rows][64] bf16 box with 128-byte swizzle
__device__ __forceinline__ uint32_t swz128(int row, int piece) {
  return (uint32_t)(row * 128 + ((piece ^ (row & 7)) << 4));
}

#define CL_STAMP(role, k)                                                                        \
  do {                                                                                           \
    if (p.trace != nullptr && blockIdx.x == 0 && local < 16) p.trace[((role) * 16 + local) * 16 + (k)] = clock64(); \
  } while (0)

template <int ACT>
__device__ __forceinline__ float cl_act(float x) {
  if (ACT == DFOL_ACT_ELU) return x > 0.0f ? x : __expf(x) - 1.0f;
  if (ACT == DFOL_ACT_SIGMOID) {  // 0.5 tanh(x/2) + 0.5: one MUFU op instead of exp + reciprocal
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
  }
  return x;
}

template <int ACT, bool HAS_E, bool WG>
__global__ void __launch_bounds__(CL_THREADS, 1)
    gemm_bf16_tc_cluster_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                                const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_e,
                                ClParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t b_full;
  __shared__ __align__(8) uint64_t a_full[CL_MAX_STAGES];
  __shared__ __align__(8) uint64_t a_empty[CL_MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ __align__(8) uint64_t e_full[CL_EC];
  __shared__ __align__(8) uint64_t e_empty[CL_EC];
  __shared__ __align__(8) uint64_t w_full;
  __shared__ __align__(8) uint64_t c_ready[CL_EC][CL_MAX_BOX];   // per 64-column box of a staging tile
  __shared__ __align__(8) uint64_t c_free[CL_MAX_BOX];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank(), CS = (uint32_t)p.cluster_size;
  const uint16_t cmask = (uint16_t)((1u << CS) - 1u);
  const int cluster_id = blockIdx.x / (int)CS, num_clusters = gridDim.x / (int)CS;
  const int num_kb = p.K / CL_BK;
  const int num_tiles = (p.M + CL_BM - 1) / CL_BM;
  const int BNh = p.BNh;
  const int nbox = BNh / 64;                                // 64-column boxes of the C / E tiles of this CTA
  const int n0 = (int)rank * BNh;                           // first column of this CTA
  const uint32_t a_bytes = CL_BM * CL_BK * 2;               // 16 KB per stage
  const uint32_t b_kb_bytes = (uint32_t)BNh * CL_BK * 2;
  const uint32_t box_bytes = CL_BM * 128;                   // one 128 x 64 bf16 box
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_tiles = base;
  uint8_t* a_tiles = b_tiles + (size_t)num_kb * b_kb_bytes;
  // staging tiles: one result tile (forward), or two tiles that first receive the epilogue operand by TMA and are
  // then overwritten IN PLACE with the result (dgrad: double buffered so the next operand load overlaps)
  uint8_t* c_tile = a_tiles + (size_t)p.stages * a_bytes;
  const uint32_t ec_bytes = (uint32_t)nbox * box_bytes;
  // bias of this CTA's columns (forward only: the region exists only without an epilogue operand)
  float* bias_s = reinterpret_cast<float*>(c_tile + (size_t)(HAS_E ? p.ec : 1) * ec_bytes);

  if (threadIdx.x == 0) {
    mbar_init(&b_full, 1);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], p.mc ? CS : 1u); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], CL_EW); }
    for (int i = 0; i < CL_EC; ++i) { mbar_init(&e_full[i], 1); mbar_init(&e_empty[i], 1); }
    mbar_init(&w_full, 1);
    for (int b = 0; b < CL_MAX_BOX; ++b) {
      // epilogue warps that write into box b: four per column part that overlaps it
      uint32_t cnt = 0;
      for (int h = 0; h < CL_CP; ++h)
        if (h * (BNh / CL_CP) < 64 * (b + 1) && (h + 1) * (BNh / CL_CP) > 64 * b) cnt += 4;
      for (int i = 0; i < CL_EC; ++i) mbar_init(&c_ready[i][b], cnt > 0 ? cnt : 1u);
      mbar_init(&c_free[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (!HAS_E)
    for (int i = threadIdx.x; i < 192; i += CL_THREADS)
      bias_s[i] = (p.bias != nullptr && i < BNh && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.0f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any multicast load / remote arrival
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
      mbar_expect_tx(&b_full, (uint32_t)num_kb * b_kb_bytes);
      for (int kb = 0; kb < num_kb; ++kb)
        tma_load_2d(&tmap_b, &b_full, b_tiles + (size_t)kb * b_kb_bytes, kb * CL_BK, n0);
      int local = 0, s = 0, eb = 0;
      uint32_t phase = 0, eph = 0;      // ring / operand-tile slots and parities as running counters (no divisions)
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
        // The A ring holds at most one tile per SM: too few bytes in flight for the HBM latency.  The tiles this
        // cluster will need next are pulled into L2 by bulk prefetches (no shared memory involved), so that the ring's
        // own loads are L2 hits.
        for (int ahead = (local == 0 ? 1 : p.prefetch); p.prefetch > 0 && ahead <= p.prefetch; ++ahead) {
          const int pt = tile + ahead * num_clusters;
          if (pt >= num_tiles) break;
          for (int kb = 0; kb < num_kb; ++kb)
            for (uint32_t j = rank; j < CL_ABOX; j += CS)
              cl_prefetch_l2(&tmap_a, kb * CL_BK, pt * CL_BM + (int)j * (CL_BM / CL_ABOX));
          if (HAS_E)
            for (int j = 0; j < nbox; ++j) cl_prefetch_l2(&tmap_e, n0 + 64 * j, pt * CL_BM);
        }
        CL_STAMP(0, 0);
        if (HAS_E) {
          mbar_wait(&e_empty[eb], eph ^ 1u);  // the store that used this buffer has read it
          CL_STAMP(0, 1);
          mbar_expect_tx(&e_full[eb], ec_bytes);
          for (int j = 0; j < nbox; ++j)
            tma_load_2d(&tmap_e, &e_full[eb], c_tile + (size_t)eb * ec_bytes + (size_t)j * box_bytes, n0 + 64 * j,
                        tile * CL_BM);
          if (++eb == p.ec) { eb = 0; eph ^= 1u; }
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&a_empty[s], phase ^ 1);  // released by the MMAs of ALL CTAs of the cluster
          CL_STAMP(0, 2 + kb);
          mbar_expect_tx(&a_full[s], a_bytes);
          // the row boxes of the block are dealt round-robin to the CTAs; each is multicast to every ring
          if (p.mc) {
            for (uint32_t j = rank; j < CL_ABOX; j += CS)
              tma_load_2d_mc(&tmap_a, &a_full[s], a_tiles + (size_t)s * a_bytes + j * (a_bytes / CL_ABOX), kb * CL_BK,
                             tile * CL_BM + (int)j * (CL_BM / CL_ABOX), cmask);
          } else {
            for (uint32_t j = 0; j < CL_ABOX; ++j)
              tma_load_2d(&tmap_a, &a_full[s], a_tiles + (size_t)s * a_bytes + j * (a_bytes / CL_ABOX), kb * CL_BK,
                          tile * CL_BM + (int)j * (CL_BM / CL_ABOX));
          }
          if (++s == p.stages) { s = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BNh >> 3) << 17) |
                             ((uint32_t)(CL_BM >> 4) << 24);
      mbar_wait(&b_full, 0);
      // WG: M = this CTA's BNh columns of E, N = the 64 columns of dZ in the ring stage, both operands MN-major
      const uint32_t idesc_w = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                               ((uint32_t)(BNh >> 4) << 24);
      int local = 0, s = 0, eb = 0;
      uint32_t phase = 0, eph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
        // WG: one dgrad accumulator (columns 0..BNh), the weight-gradient accumulator behind it
        const int buf = WG ? 0 : (local & 1);
        const uint32_t use = WG ? (uint32_t)local : (uint32_t)(local >> 1);
        CL_STAMP(1, 0);
        if (WG) {
          mbar_wait(&e_full[eb], eph);  // the H tile is an MMA operand too
        } else {
          mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
        }
        CL_STAMP(1, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + (uint32_t)(buf * 256);
        const uint32_t e_addr = smem_u32(c_tile + (size_t)eb * ec_bytes);
        if (WG && p.stages == num_kb) {
          // The ring holds exactly one tile (stage kb = k-block kb, adjacent stages 16 KB apart like the boxes of an
          // MN-major operand): the weight-gradient MMAs take TWO k-blocks at once (N = 128), which halves the fetches of
          // their M operand -- operand reads from shared memory are what bounds these MMAs (measured per MMA, from the
          // issue stamps of DFOL_CL_TRACE: 128 x 64 x 16 MN-major ~120 cycles, 128 x 128 x 16 MN-major ~140, the K-major
          // 128 x 128 x 16 of the dgrad ~93 -- 60-90 B/clk of operand bytes -- against 32 / 64 / 64 by the tensor pipe;
          // interleaving the independent accumulator chains changes nothing).  They go first: they do not touch the dgrad accumulator, which the epilogue of the
          // previous tile may still be copying to registers.
          for (int kb = 0; kb < num_kb; kb += 2) {
            const int nk = min(2, num_kb - kb);
            for (int q = 0; q < nk; ++q) {
              mbar_wait(&a_full[kb + q], phase);
              CL_STAMP(1, 2 + kb + q);
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_addr = smem_u32(a_tiles + (size_t)kb * a_bytes);
            const uint64_t we = cl_smem_desc_mn(e_addr, box_bytes), wa = cl_smem_desc_mn(a_addr, a_bytes);
            const uint32_t accw = tmem_base + (uint32_t)(BNh + CL_BK * kb);
            const uint32_t idw = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                 ((uint32_t)((64 * nk) >> 3) << 17) | ((uint32_t)(BNh >> 4) << 24);
            if (!(p.debug & 1))
#pragma unroll
            for (int k = 0; k < CL_BM / 16; ++k)
              umma_bf16(accw, we + 128 * k, wa + 128 * k, idw, (local > 0 || k > 0) ? 1u : 0u);
            if (kb == 0) {
              mbar_wait(&acc_empty[0], (use & 1) ^ 1);
              CL_STAMP(1, 8);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            for (int q = 0; q < nk; ++q) {
              const uint64_t da = make_smem_desc(a_addr + (uint32_t)q * a_bytes);
              const uint64_t db = make_smem_desc(smem_u32(b_tiles + (size_t)(kb + q) * b_kb_bytes));
              if (!(p.debug & 2))
#pragma unroll
              for (int k = 0; k < CL_BK / 16; ++k)
                umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (kb + q > 0 || k > 0) ? 1u : 0u);
              if (p.mc) umma_commit_mc(&a_empty[kb + q], cmask);
              else umma_commit(&a_empty[kb + q]);
            }
          }
          phase ^= 1u;
        } else
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&a_full[s], phase);
          CL_STAMP(1, 2 + kb);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(a_tiles + (size_t)s * a_bytes);
          if (WG) {
            // weight-gradient MMAs first: they do not touch the dgrad accumulator, which the epilogue of the previous
            // tile may still be copying to registers
            const uint64_t we = cl_smem_desc_mn(e_addr, box_bytes), wa = cl_smem_desc_mn(a_addr, box_bytes);
            const uint32_t accw = tmem_base + (uint32_t)(BNh + CL_BK * kb);
            if (!(p.debug & 1))
#pragma unroll
            for (int k = 0; k < CL_BM / 16; ++k)   // 16 rows = two 8-row groups = 2048 B = +128 in the address field
              umma_bf16(accw, we + 128 * k, wa + 128 * k, idesc_w, (local > 0 || k > 0) ? 1u : 0u);
            if (kb == 0) {
              mbar_wait(&acc_empty[0], (use & 1) ^ 1);
              CL_STAMP(1, 8);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
          }
          const uint64_t da = make_smem_desc(a_addr);
          const uint64_t db = make_smem_desc(smem_u32(b_tiles + (size_t)kb * b_kb_bytes));
          if (!(p.debug & 2))
#pragma unroll
          for (int k = 0; k < CL_BK / 16; ++k)
            umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          if (p.mc) umma_commit_mc(&a_empty[s], cmask);  // the stage is free everywhere once every CTA's MMAs have read it
          else umma_commit(&a_empty[s]);
          if (++s == p.stages) { s = 0; phase ^= 1u; }
        }
        umma_commit(&acc_full[buf]);
        CL_STAMP(1, 9);
        if (++eb == p.ec) { eb = 0; eph ^= 1u; }
      }
      if (WG) umma_commit(&w_full);
    }
  } else if (warp < 2 + CL_EW) {
    // ---------------- epilogue: warps 2..; TMEM lane quadrant = warp % 4, column part = (warp - 2) / 4 ----------------
    const int quad = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int cw = BNh / CL_CP;
    const int cbeg = ch * cw;
    const int nchunks = cw / 16;
    const bool stamper = (warp == 2 && lane == 0);
    // derivative factor of the saved activation: 0 elu', 1 sigmoid', 2 either one behind a dropout mask
    const int emode = (p.keep < 1.0f) ? 2 : (p.mul_mode == DFOL_MUL_SIGMOID_GRAD ? 1 : 0);
    const int my_boxes = (min(max(p.n_store - n0, 0), BNh) + 63) / 64;   // boxes the store warp stores (and frees)
    int local = 0, eb = 0;
    uint32_t eph = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
      const int buf = WG ? 0 : (local & 1);
      const uint32_t use = WG ? (uint32_t)local : (uint32_t)(local >> 1);
      if (stamper) CL_STAMP(2, 0);
      mbar_wait(&acc_full[buf], use & 1);
      if (stamper) CL_STAMP(2, 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint8_t* stage = c_tile + (HAS_E ? (size_t)eb * ec_bytes : 0);
      if (HAS_E) mbar_wait(&e_full[eb], eph);
      if (stamper) CL_STAMP(2, 2);
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 256 + cbeg);
      constexpr int NR = WG ? CL_WG_CH : 2;
      uint32_t r[NR][16];
      if (WG) {
        // single accumulator: copy this thread's columns to registers and hand the accumulator straight back, so the
        // next tile's MMAs run under this tile's epilogue arithmetic
#pragma unroll
        for (int c = 0; c < CL_WG_CH; ++c)
          if (c < nchunks) tmem_ld16(trow + (uint32_t)(16 * c), r[c % NR]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) cl_mbar_arrive(&acc_empty[0]);
        if (stamper) CL_STAMP(2, 3);
      } else {
        tmem_ld16(trow, r[0]);
      }
#pragma unroll
      for (int c = 0; c < CL_MAX_CH; ++c) {
        if (c >= nchunks) break;
        const int c0 = cbeg + 16 * c;  // column inside this CTA's half
        if (!WG) {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c + 1 < nchunks) tmem_ld16(trow + (uint32_t)(16 * (c + 1)), r[(c + 1) % NR]);
        }
        const int box = c0 >> 6, piece = (c0 & 63) >> 3;  // 16-byte piece index of the chunk's first 8 columns
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = cl_act<ACT>(__uint_as_float(r[c % NR][j]) + (HAS_E ? 0.0f : bias_s[c0 + j]));
        if (HAS_E) {
          const uint8_t* eb8 = stage + (size_t)box * box_bytes;
          const uint4 h0 = *reinterpret_cast<const uint4*>(eb8 + swz128(row, piece));
          const uint4 h1 = *reinterpret_cast<const uint4*>(eb8 + swz128(row, piece + 1));
          const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
          if (emode == 0) {          // elu'(h) = 1 (h > 0), h + 1 (h <= 0)  ==  min(h, 0) + 1
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[j]));
              v[2 * j] *= fminf(h.x, 0.0f) + 1.0f;
              v[2 * j + 1] *= fminf(h.y, 0.0f) + 1.0f;
            }
          } else if (emode == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[j]));
              v[2 * j] *= h.x * (1.0f - h.x);
              v[2 * j + 1] *= h.y * (1.0f - h.y);
            }
          } else {                   // act' at h = operand * keep, mask factor (operand != 0) / keep
            const float ik = 1.0f / p.keep;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[j]));
              const float hx = h.x * p.keep, hy = h.y * p.keep;
              const float gx = (p.mul_mode == DFOL_MUL_SIGMOID_GRAD) ? hx * (1.0f - hx) : (hx > 0.0f ? 1.0f : hx + 1.0f);
              const float gy = (p.mul_mode == DFOL_MUL_SIGMOID_GRAD) ? hy * (1.0f - hy) : (hy > 0.0f ? 1.0f : hy + 1.0f);
              v[2 * j] *= (h.x != 0.0f ? ik : 0.0f) * gx;
              v[2 * j + 1] *= (h.y != 0.0f ? ik : 0.0f) * gy;
            }
          }
        }
        if (n0 + c0 + 16 > p.N) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c0 + j >= p.N) v[j] = 0.0f;  // K padding of the next layer
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
          pk[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        // forward: ONE staging tile -- the store warp must have read the previous tile's box out of it before the
        // first write into the box; the chunk's arithmetic is already done by then
        if (!HAS_E && local >= 1 && box < my_boxes && (c == 0 || ((c0 - 16) >> 6) != box))
          mbar_wait(&c_free[box], (uint32_t)((local - 1) & 1));
        uint8_t* cb = stage + (size_t)box * box_bytes;
        *reinterpret_cast<uint4*>(cb + swz128(row, piece)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(cb + swz128(row, piece + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        if (c + 1 == nchunks || ((c0 + 16) >> 6) != box) {
          // this warp's part of the box is complete: publish it to the async proxy and tell the store warp, which
          // stores the box while the remaining chunks are still being computed
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) cl_mbar_arrive(&c_ready[HAS_E ? eb : 0][box]);
        }
      }
      if (!WG) {
        // accumulator buffer is free again
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) cl_mbar_arrive(&acc_empty[buf]);
      }
      if (stamper) CL_STAMP(2, 4);
      if (++eb == p.ec) { eb = 0; eph ^= 1u; }
    }
    if (WG && local > 0) {
      // weight gradient of this CTA: lanes = its columns of E (columns n0.. of dW), TMEM columns BNh.. = rows of dW;
      // consecutive lanes add to consecutive addresses
      mbar_wait(&w_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int kw = p.K / CL_CP;                   // dW rows of this column part (multiple of 16: K % 64 == 0)
      const uint32_t wrow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)BNh;
      const int col = n0 + row;
      for (int k0 = ch * kw; k0 < (ch + 1) * kw && k0 < p.K_real; k0 += 16) {
        uint32_t w[16];
        tmem_ld16(wrow + (uint32_t)k0, w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (col < p.N) {
          float* dst = p.dW + (long long)k0 * p.lddw + col;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (k0 + j < p.K_real) atomicAdd(dst + (long long)j * p.lddw, __uint_as_float(w[j]));
        }
      }
    }
  } else {
    // ---------------- store warp (warp 10, one lane): staging tile -> global by TMA, off the epilogue's critical path;
    // the tile's buffer goes back (to the operand loads of the dgrad / to the next tile's epilogue) as soon as the bulk
    // store has READ it
    if (lane == 0) {
      int local = 0, eb = 0;
      uint32_t eph = 0;
      const int my_cols = min(max(p.n_store - n0, 0), BNh);
      const int my_boxes = (my_cols + 63) / 64;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
        const int slot = HAS_E ? eb : 0;
        const uint32_t par = HAS_E ? eph : (uint32_t)(local & 1);
        const uint8_t* stage = c_tile + (size_t)slot * ec_bytes;
        for (int j = 0; j < my_boxes; ++j) {   // box by box, as the epilogue warps complete them
          mbar_wait(&c_ready[slot][j], par);
          if (j == 0) CL_STAMP(3, 0);
          tma_store_2d(&tmap_c, stage + (size_t)j * box_bytes, n0 + 64 * j, tile * CL_BM);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        CL_STAMP(3, 2);
        if (HAS_E) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          cl_mbar_arrive(&e_empty[eb]);
        } else {
          // (bulk groups complete in order: <= k groups pending means the first my_boxes - k boxes have been read)
          if (my_boxes >= 3) { asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); cl_mbar_arrive(&c_free[my_boxes - 3]); }
          if (my_boxes >= 2) { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); cl_mbar_arrive(&c_free[my_boxes - 2]); }
          if (my_boxes >= 1) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); cl_mbar_arrive(&c_free[my_boxes - 1]); }
        }
        CL_STAMP(3, 1);
        if (++eb == p.ec) { eb = 0; eph ^= 1u; }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // the peer may still multicast into this CTA's ring / arrive on its barriers until it is done
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

typedef void (*ClKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, ClParams);

}  // namespace dfol

using namespace dfol;

static int launch_cluster(const char* who, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                          int store_cols, const float* bias, int M, int N, int K, int act, const void* E, int64_t lde,
                          int mul_mode, void* stream, float keep = 1.0f, float* dW = nullptr, int64_t lddw = 0,
                          int k_real = 0) {
  DFOL_REQUIRE(A && B && C, "%s: null pointer", who);
  DFOL_REQUIRE(keep > 0.0f && keep <= 1.0f, "%s: keep = 1 - dropout p must be in (0, 1]", who);
  DFOL_REQUIRE(M > 0 && N > 0 && K > 0 && (K % CL_BK) == 0 && K <= CL_BK * CL_MAX_KB,
               "%s: K must be a multiple of 64, at most %d", who, CL_BK * CL_MAX_KB);
  DFOL_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0 && (ldc % 8) == 0 && lda >= K && ldb >= K,
               "%s: strides must be multiples of 8 elements, lda/ldb >= K", who);
  DFOL_REQUIRE((reinterpret_cast<uintptr_t>(A) % 16) == 0 && (reinterpret_cast<uintptr_t>(B) % 16) == 0 &&
                   (reinterpret_cast<uintptr_t>(C) % 16) == 0,
               "%s: operands must be 16-byte aligned", who);
  const int n_store = store_cols > 0 ? store_cols : (int)ldc;
  DFOL_REQUIRE(n_store >= N && n_store <= ldc && (n_store % 8) == 0, "%s: N <= store_cols <= ldc, multiple of 8", who);
  const bool has_e = mul_mode != DFOL_MUL_NONE;
  DFOL_REQUIRE(!has_e || (E && (lde % 8) == 0 && (reinterpret_cast<uintptr_t>(E) % 16) == 0 && lde >= N),
               "%s: multiplier operand must be 16-byte aligned with ld %% 8 == 0", who);
  ClParams p;
  p.bias = bias; p.M = M; p.N = N; p.K = K; p.act = act; p.mul_mode = mul_mode; p.keep = keep;
  static const int pf_env = [] { const char* e = getenv("DFOL_CL_PREFETCH"); return e ? atoi(e) : 1; }();
  p.prefetch = pf_env;   // (measured at c3: dgrad 1.08 ms without, 0.97 ms at distance 1, 0.98 at 2, 1.31 at 4)
  // columns per CTA and cluster size.  Two CTAs (half of the stored width each, whole 64-column boxes) measured best:
  // clusters of 3-4 CTAs leave room for a 7-8 stage A ring but run 2x slower (every ring stage is released by a
  // commit from every CTA of the cluster, and the lockstep of 3-4 SMs costs more than the deeper ring gains).
  p.BNh = ((n_store + 1) / 2 + 63) / 64 * 64;
  DFOL_REQUIRE(p.BNh <= 192, "%s: at most 384 output columns", who);
  const int CS = (n_store + p.BNh - 1) / p.BNh;
  p.n_store = n_store;
  p.cluster_size = CS;
  p.dW = dW; p.lddw = lddw; p.K_real = k_real;
  static const int dbg_env = [] { const char* e = getenv("DFOL_CL_DEBUG"); return e ? atoi(e) : 0; }();
  p.debug = dbg_env;
  static const int trace_env = [] { const char* e = getenv("DFOL_CL_TRACE"); return e ? atoi(e) : 0; }();
  p.trace = nullptr;
  if (trace_env) {
    cudaMalloc(&p.trace, 4 * 16 * 16 * sizeof(long long));
    cudaMemset(p.trace, 0, 4 * 16 * 16 * sizeof(long long));
  }
  static const int mc_env = [] { const char* e = getenv("DFOL_CL_MC"); return e ? atoi(e) : 1; }();
  p.mc = mc_env;
  if (dW != nullptr) {
    DFOL_REQUIRE(has_e && p.BNh == 128 && p.BNh + K <= 512 && k_real > 0 && k_real <= K && lddw >= N,
                 "%s: the fused weight gradient needs 65..256 output columns, K <= 384 and a multiplier operand", who);
  }
  const int num_kb = K / CL_BK;
  const size_t b_bytes = (size_t)num_kb * p.BNh * CL_BK * 2;
  const size_t box = (size_t)CL_BM * 128;
  static const int ec_env = [] { const char* e = getenv("DFOL_CL_EC"); return e ? atoi(e) : 0; }();
  // two operand tiles (each goes back to the loads as soon as the store warp's bulk store has read it) leave room for a
  // ring of one whole tile: c4 dgrad 0.314 ms with three tiles / three stages, 0.296 ms with two / five
  p.ec = (ec_env >= 2 && ec_env <= CL_EC) ? ec_env : 2;
  const size_t tiles_bytes = (size_t)(p.BNh / 64) * box * (has_e ? p.ec : 1);  // dgrad: in-place operand/result tiles
  const size_t a_stage = (size_t)CL_BM * CL_BK * 2;
  DFOL_REQUIRE(!has_e || bias == nullptr, "%s: the multiplier epilogue has no bias", who);
  const size_t bias_bytes = has_e ? 0 : 192 * sizeof(float);
  const size_t budget = 232448 - 1536 - bias_bytes;  // 227 KB - static shared memory (barriers: < 1.5 KB)
  DFOL_REQUIRE(b_bytes + tiles_bytes + 2 * a_stage + 1024 <= budget, "%s: does not fit in shared memory", who);
  int stages = (int)((budget - 1024 - b_bytes - tiles_bytes) / a_stage);
  if (stages > CL_MAX_STAGES) stages = CL_MAX_STAGES;
  p.stages = stages;
  const size_t smem = b_bytes + tiles_bytes + (size_t)stages * a_stage + 1024 + bias_bytes;
  ClKernel kernel;
  if (has_e && dW != nullptr) kernel = gemm_bf16_tc_cluster_kernel<DFOL_ACT_NONE, true, true>;
  else if (has_e) kernel = gemm_bf16_tc_cluster_kernel<DFOL_ACT_NONE, true, false>;
  else if (act == DFOL_ACT_SIGMOID) kernel = gemm_bf16_tc_cluster_kernel<DFOL_ACT_SIGMOID, false, false>;
  else if (act == DFOL_ACT_ELU) kernel = gemm_bf16_tc_cluster_kernel<DFOL_ACT_ELU, false, false>;
  else kernel = gemm_bf16_tc_cluster_kernel<DFOL_ACT_NONE, false, false>;
  DFOL_REQUIRE(!has_e || act == DFOL_ACT_NONE, "%s: the multiplier epilogue has no activation", who);
  {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return (int)e; }
  }
  alignas(64) CUtensorMap ma, mb, mc, me;
  int rc = encode_map_bf16(&ma, A, M, K, lda, 64);
  if (rc != 0) return rc;
  rc = encode_map_bf16(&mb, B, N, K, ldb, p.BNh);
  if (rc != 0) return rc;
  rc = encode_map_bf16(&mc, C, M, n_store, ldc, CL_BM);
  if (rc != 0) return rc;
  rc = encode_map_bf16(&me, has_e ? E : C, M, has_e ? N : n_store, has_e ? lde : ldc, CL_BM);
  if (rc != 0) return rc;
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles = (M + CL_BM - 1) / CL_BM;
  int clusters = sms / CS;
  if (clusters > tiles) clusters = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS * clusters, 1, 1);
  cfg.blockDim = dim3(CL_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  {
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, ma, mb, mc, me, p);
    if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return (int)e; }
  }
  if (p.trace != nullptr) {   // timing experiment: per-tile stamps of block 0 (cycles relative to the first stamp)
    static long long host[4 * 16 * 16];
    cudaDeviceSynchronize();
    cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    long long t0 = 0;
    for (int i = 0; i < 4 * 16 * 16; ++i) if (host[i] && (!t0 || host[i] < t0)) t0 = host[i];
    static const char* names[4] = {"load", "mma ", "epi ", "stor"};
    fprintf(stderr, "[%s trace] M=%d N=%d K=%d stages=%d ec=%d wg=%d\n", who, M, N, K, p.stages, p.ec, dW != nullptr);
    for (int t = 0; t < 12; ++t)
      for (int r = 0; r < 4; ++r) {
        fprintf(stderr, "  tile %2d %s:", t, names[r]);
        for (int k = 0; k < 10; ++k) fprintf(stderr, " %7lld", host[(r * 16 + t) * 16 + k] ? host[(r * 16 + t) * 16 + k] - t0 : -1);
        fprintf(stderr, "\n");
      }
  }
  return finish_launch(who);
}

extern "C" int dfol_pair_layer_fwd_cluster(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                                           int store_cols, const float* bias, int M, int N, int K, int act,
                                           void* stream) {
  return launch_cluster("dfol_pair_layer_fwd_cluster", A, lda, B, ldb, C, ldc, store_cols, bias, M, N, K, act, nullptr,
                        0, DFOL_MUL_NONE, stream);
}

extern "C" int dfol_pair_layer_dgrad_cluster(const void* dZ, int64_t lddz, const void* Wt, int64_t ldwt, void* dX,
                                             int64_t lddx, int store_cols, int M, int N, int K, const void* h_saved,
                                             int64_t ldh, int mul_mode, float keep, void* stream) {
  return launch_cluster("dfol_pair_layer_dgrad_cluster", dZ, lddz, Wt, ldwt, dX, lddx, store_cols, nullptr, M, N, K,
                        DFOL_ACT_NONE, h_saved, ldh, mul_mode, stream, keep);
}

/* dgrad + weight gradient of the same layer in one pass over dZ and h_saved:
 *   dX = (dZ . Wt^T) * act'(h_saved)   and   dW[k, n] += sum_rows dZ[row, k] * h_saved[row, n]   (k < k_real) */
extern "C" int dfol_pair_layer_dgrad_wgrad_cluster(const void* dZ, int64_t lddz, const void* Wt, int64_t ldwt, void* dX,
                                                   int64_t lddx, int store_cols, int M, int N, int K,
                                                   const void* h_saved, int64_t ldh, int mul_mode, float keep,
                                                   float* dW, int64_t lddw, int k_real, void* stream) {
  DFOL_REQUIRE(dW != nullptr, "dfol_pair_layer_dgrad_wgrad_cluster: null weight gradient");
  return launch_cluster("dfol_pair_layer_dgrad_wgrad_cluster", dZ, lddz, Wt, ldwt, dX, lddx, store_cols, nullptr, M, N,
                        K, DFOL_ACT_NONE, h_saved, ldh, mul_mode, stream, keep, dW, lddw, k_real);
}
