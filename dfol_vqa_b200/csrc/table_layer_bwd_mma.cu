// Backward of the relation table layer LL = logsigmoid(H2 . W^T + b) on the tensor cores (sm_100a), from the compact
// gradient slices the backward interpreter emits (same contract as table_layer_bwd_tc_kernel, tc_support.cu):
//   dz_j[l]    = g_j[l] * (1 - exp(LL_j[l]))                               (logsigmoid', slice j of the image)
//   dZ2[l, e]  = (sum_j dz_j[l] W[wrow_j, e]) * h (1 - h),  h = H2[l, e]    (sigmoid' of the layer below)
//   dW[wrow_j, e] += sum_l dz_j[l] h;   db[wrow_j] += sum_l dz_j[l];   dbelow[e] += sum_l dZ2[l, e]
// (reference: autograd of classifier_oracle.py:149-154 restricted to the relation columns a program reads).
//
// The SIMT kernel spends 4 FMAs per (row, slice, column pair): with 9-12 slices per image (relation chains, N = 100) it
// is compute-bound at 1.4-1.7 TB/s.  Here the three contractions are tcgen05 MMAs on ONE 128-row tile of H2 that TMA
// brought to shared memory, and the SIMT work is the sigmoid' multiply only:
//   MMA-A  acc[128 x C]   = DZ (128 x S) . Wslots (S x C)       K = S (16 or 32 slices, bf16), fp32 in TMEM
//   MMA-B  dWt[C x S]    += H2^T (C x 128) . DZ (128 x S)       per image, accumulated in TMEM over the image's tiles
//   epilogue: dZ2 = acc * h (1 - h) written IN PLACE over the H2 tile (bf16, 128-byte swizzle), stored by TMA
//   colsum[C] += column sums of the dZ2 tile, by the epilogue warps from shared memory (registers over all tiles of the
//                CTA; round 2 had them as a third MMA against a tile of ones: 24 MN-major MMAs, ~4600 cycles per tile)
// H2 is read once and dZ2 written once (the algorithmic traffic); the operands of MMA-B / MMA-C are the SAME shared
// tiles seen through MN-major descriptors (rows = K), the slice gradients DZ are one small tile ([S][128] bf16) that
// serves MMA-A as an MN-major A operand and MMA-B as a K-major B operand.
//
// Persistent CTAs, each owning a contiguous run of (image, tile) pairs; warp roles:
//   warp 0 (lane 0)  TMA producer: H2 tile (C/64 boxes of 128 x 64) + the image's slice rows of W (bf16, gathered by
//                    table_slice_weights_kernel) into a 2-stage ring
//   warp 1 (lane 0)  tcgen05.mma issuer (+ TMEM allocation)
//   warps 2-5        DZ builders: thread = row, gathers g and LL of the tile's S slices, writes the DZ tile, keeps the
//                    db partial sums in registers until the image ends
//   warps 6-13       epilogue (TMEM lane quadrant = warp % 4, column half = (warp - 6) / 4) + column sums
//   warp 14 (lane 0) store warp: bulk store of the finished tile, hands the buffer back when the store has read it
// Every mbarrier wait is bounded (trap instead of hanging the GPU).
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"

namespace dfol {

constexpr int TM_BM = 128;
constexpr int TM_THREADS = 64 + 128 + 256 + 32;   // producer, MMA issuer, 4 DZ warps, 8 epilogue warps, store warp
constexpr uint32_t TM_BOX = 128 * 128;   // bytes of one 128-row x 64-column bf16 box
constexpr int TM_MAX_NB = 5;             // columns <= 320
constexpr uint32_t TM_ACC_COL = 0, TM_DW_COL = 320, TM_CS_COL = 416;   // TMEM columns (512 allocated)

struct TmParams {
  const float* g; const int32_t* slice_goff; const int32_t* slice_col; const int32_t* slice_wrow;
  const int32_t* img_slice;
  const float* ll; const int64_t* blk; const int32_t* stride; const int32_t* row0; const int32_t* img_rows;
  const int32_t* tile_start;   // [images + 1]: prefix sums of ceil(img_rows / 128)
  int images, total_tiles, E, NB;
  int debug;   // ablation switches (DFOL_TBL_DEBUG; results are wrong when set): 1 no gathers, 2 no epilogue math, 4 no MMA-B/C, 8 no store
  __nv_bfloat16* dZ; long long lddz;
  float* dW; long long ldw; float* db; float* dbelow;
  int trace_first, trace_block;   // DFOL_TBL_TRACE=1+first tile of the window, DFOL_TBL_TRACE_BLOCK
  long long* trace;   // DFOL_TBL_TRACE=1: clock64 stamps of block 0's roles, first 16 active tiles ([role][tile][16])
};

__device__ __forceinline__ uint64_t tm_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);      // start address (16-byte units)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16; // leading byte offset: next 64-element block along M / N
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset: next group of 8 k-rows
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tm_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tm_named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tm_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tm_prefetch_l2(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ uint32_t tm_swz(int row, int piece) {  // 16-byte piece of a 128-byte row, 128B swizzle
  return (uint32_t)(row * 128 + ((piece ^ (row & 7)) << 4));
}

#define TM_STAMP(role, k)                                                                              \
  do {                                                                                                 \
    if (p.trace != nullptr && blockIdx.x == p.trace_block && i >= p.trace_first && i < p.trace_first + 16)                       \
      p.trace[((role) * 16 + i - p.trace_first) * 16 + (k)] = clock64();                                 \
  } while (0)

// walk over the (image, tile) pairs of this CTA
struct TmTile {
  int tile, end, b, c, cn, j0, Sb;
  bool last_of_image;
  __device__ __forceinline__ void locate(const TmParams& p) {
    int lo = 0, hi = p.images - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (p.tile_start[mid] <= tile) lo = mid; else hi = mid - 1;
    }
    b = lo;
  }
  __device__ __forceinline__ void load(const TmParams& p, int SP) {
    const int t0 = p.tile_start[b], t1 = p.tile_start[b + 1];
    c = (tile - t0) * TM_BM;
    cn = min(TM_BM, p.img_rows[b] - c);
    j0 = p.img_slice[b];
    Sb = min(p.img_slice[b + 1] - j0, SP);
    last_of_image = (tile + 1 == t1) || (tile + 1 == end);
  }
  __device__ __forceinline__ bool valid() const { return tile < end; }
  __device__ __forceinline__ void next(const TmParams& p) {
    ++tile;
    if (tile < end) while (tile >= p.tile_start[b + 1]) ++b;
  }
};

template <int SP>
__global__ void __launch_bounds__(TM_THREADS, 1)
    table_layer_bwd_mma_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_w,
                               const __grid_constant__ CUtensorMap tmap_z, TmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t h_full[2], dz_full[2], dz_free[2], buf_free[2];
  __shared__ __align__(8) uint64_t mma_done, acc_empty, st_ready, all_done;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NB = p.NB;
  constexpr uint32_t WBOX = SP * 128;                 // bytes of one [SP x 64] bf16 box
  const uint32_t stage_bytes = (uint32_t)NB * (TM_BOX + WBOX);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* dz_tiles = base + 2 * (size_t)stage_bytes;          // [2][2 boxes of SP x 64]
  uint8_t* ones_tile = dz_tiles + 2 * 2 * WBOX;                // [2 boxes of 16 x 64], all 1.0
  auto h_tile = [&](int buf) { return base + (size_t)buf * stage_bytes; };
  auto w_tile = [&](int buf) { return base + (size_t)buf * stage_bytes + (size_t)NB * TM_BOX; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&h_full[i], 1); mbar_init(&dz_full[i], 128); mbar_init(&dz_free[i], 1);
      mbar_init(&buf_free[i], 1u + (uint32_t)p.NB);   // the tile's store has read it + the NB column-sum warps
    }
    mbar_init(&mma_done, 1); mbar_init(&acc_empty, 8); mbar_init(&st_ready, 1); mbar_init(&all_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  for (int i = threadIdx.x; i < 4096 / 4; i += TM_THREADS) reinterpret_cast<uint32_t*>(ones_tile)[i] = 0x3F803F80u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  TmTile t;
  t.tile = (int)(((long long)p.total_tiles * blockIdx.x) / gridDim.x);
  t.end = (int)(((long long)p.total_tiles * (blockIdx.x + 1)) / gridDim.x);
  if (t.valid()) t.locate(p);
  // the three 128-row blocks of the transposed accumulators: block m holds columns 64 * mbox[m] + lane; when the last
  // block would run past the tile it is shifted left and only its upper lanes are new
  const int mblocks = (NB + 1) / 2;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_h)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
      // The two shared stages hold one tile in flight while one is being worked on: not enough bytes in flight to cover
      // the HBM latency.  The NEXT tile is therefore pulled into L2 by a bulk prefetch (no shared memory needed) while
      // this one is loaded, so that its own load is an L2 hit.  Distance ONE tile: at three tiles ahead the prefetched
      // lines were evicted by the kernel's own store stream before use (ncu: 3.33 GB of DRAM reads against 1.64 GB of
      // tile data at c3 -- the L2 turns over in ~20 us at 6 TB/s).
      TmTile u = t;
      auto prefetch_next = [&]() {
        while (u.valid()) {
          u.load(p, SP);
          const bool active = u.Sb > 0;
          if (active) {
            const int row = p.row0[u.b] + u.c;
            for (int k = 0; k < NB; ++k) tm_prefetch_l2(&tmap_h, 64 * k, row);
          }
          u.next(p);
          if (active) return;
        }
      };
      prefetch_next();   // (tile 0: about to be loaded anyway; moves the cursor to tile 1)
      int i = 0;
      for (; t.valid(); t.next(p)) {
        t.load(p, SP);
        if (t.Sb <= 0) continue;
        prefetch_next();
        const int buf = i & 1;
        const uint32_t n = (uint32_t)(i >> 1);
        TM_STAMP(0, 0);
        mbar_wait(&buf_free[buf], (n & 1) ^ 1);
        TM_STAMP(0, 1);
        mbar_expect_tx(&h_full[buf], stage_bytes);
        const int row = p.row0[t.b] + t.c;
        for (int k = 0; k < NB; ++k) tma_load_2d(&tmap_h, &h_full[buf], h_tile(buf) + (size_t)k * TM_BOX, 64 * k, row);
        for (int k = 0; k < NB; ++k) tma_load_2d(&tmap_w, &h_full[buf], w_tile(buf) + (size_t)k * WBOX, 64 * k, SP * t.b);
        ++i;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t id_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TM_BM >> 4) << 24);
      const int n_lo = min(NB, 3) * 64, n_hi = NB * 64 - n_lo;
      const uint32_t idA_lo = id_base | (1u << 15) | (1u << 16) | ((uint32_t)(n_lo >> 3) << 17);
      const uint32_t idA_hi = id_base | (1u << 15) | (1u << 16) | ((uint32_t)(n_hi >> 3) << 17);
      const uint32_t idB = id_base | (1u << 15) | ((uint32_t)(SP >> 3) << 17);   // A MN-major, B K-major, N = SP
      const uint32_t idC = id_base | (1u << 15) | ((uint32_t)(16 >> 3) << 17);
      int i = 0, prev_b = -1;
      for (; t.valid(); t.next(p)) {
        t.load(p, SP);
        if (t.Sb <= 0) continue;
        const int buf = i & 1;
        const uint32_t n = (uint32_t)(i >> 1);
        TM_STAMP(1, 0);
        mbar_wait(&h_full[buf], n & 1);
        TM_STAMP(1, 1);
        mbar_wait(&dz_full[buf], n & 1);
        TM_STAMP(1, 2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t hs = smem_u32(h_tile(buf)), ws = smem_u32(w_tile(buf));
        const uint32_t ds = smem_u32(dz_tiles + (size_t)buf * 2 * WBOX);
        const bool fresh = (t.b != prev_b);
        // MMA-B: dWt[block m] += H2^T . DZ   (A: H2 boxes MN-major, M = column, K = row; B: DZ K-major, N = slice).  Within
        // an image it only ADDS to its own accumulators, so it is issued BEFORE waiting for the epilogue of the previous
        // tile (off the epilogue -> MMA-A -> epilogue path); the first tile of an image overwrites them and has to wait
        // until the previous image's rows have been flushed (the flush precedes the acc_empty arrival).
        auto issue_dw = [&]() {
          for (int k = 0; k < ((p.debug & 4) ? 1 : TM_BM / 16); ++k) {
            const uint64_t db = make_smem_desc(ds + (uint32_t)(k >> 2) * WBOX + 32u * (k & 3));
            for (int m = 0; m < mblocks; ++m) {
              const int mb = min(2 * m, NB - 2);
              umma_bf16(tmem_base + TM_DW_COL + (uint32_t)(m * SP), tm_desc_mn(hs + (uint32_t)mb * TM_BOX + 2048u * k, TM_BOX),
                        db, idB, (fresh && k == 0) ? 0u : 1u);
            }
          }
        };
        if (!fresh) issue_dw();
        TM_STAMP(1, 3);
        mbar_wait(&acc_empty, (uint32_t)(i & 1) ^ 1u);   // the epilogue of tile i-1 has drained the accumulators
        TM_STAMP(1, 4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // MMA-A: acc = DZ . Wslots   (A: DZ MN-major, M = row, K = slice; B: Wslots MN-major, K = slice, N = column)
#pragma unroll
        for (int k = 0; k < SP / 16; ++k) {
          const uint64_t da = tm_desc_mn(ds + 2048u * k, WBOX);
          umma_bf16(tmem_base + TM_ACC_COL, da, tm_desc_mn(ws + 2048u * k, WBOX), idA_lo, k > 0 ? 1u : 0u);
          if (n_hi > 0)
            umma_bf16(tmem_base + TM_ACC_COL + (uint32_t)n_lo, da, tm_desc_mn(ws + 3 * WBOX + 2048u * k, WBOX), idA_hi,
                      k > 0 ? 1u : 0u);
        }
        if (fresh) issue_dw();
        umma_commit(&mma_done);
        umma_commit(&dz_free[buf]);
        TM_STAMP(1, 5);
        prev_b = t.b;
        ++i;
      }
      umma_commit(&all_done);
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ DZ builders: thread = row of the tile
    const int l = threadIdx.x - 64;
    float dbacc[SP];
#pragma unroll
    for (int j = 0; j < SP; ++j) dbacc[j] = 0.0f;
    int i = 0;
    for (; t.valid(); t.next(p)) {
      t.load(p, SP);
      if (t.Sb <= 0) continue;
      const int buf = i & 1;
      const uint32_t n = (uint32_t)(i >> 1);
      if (l == 0) TM_STAMP(2, 0);
      mbar_wait(&dz_free[buf], (n & 1) ^ 1);
      if (l == 0) TM_STAMP(2, 1);
      uint8_t* dz = dz_tiles + (size_t)buf * 2 * WBOX + (size_t)(l >> 6) * WBOX;
      const int lc = l & 63;
      const bool ok = l < t.cn;
      const int lr = ok ? l : 0;                         // (clamped: every load below is in range and unconditional)
      const float* gp = p.g + t.c + lr;
      const float* lp = p.ll + p.blk[t.b] + t.c + lr;
      const long long st = p.stride[t.b];
      // all loads of the tile are issued before any is used: S independent gathers in flight per thread instead of a
      // chain of two DRAM latencies per slice (slices beyond the image's count re-read its last one, masked below)
      float gv[SP], lv[SP];
#pragma unroll
      for (int j = 0; j < SP; ++j) {
        const int jj = t.j0 + min(j, t.Sb - 1);
        gv[j] = (p.debug & 1) ? 0.5f : __ldg(gp + __ldg(p.slice_goff + jj));
        lv[j] = (p.debug & 1) ? -1.0f : __ldg(lp + (long long)__ldg(p.slice_col + jj) * st);
      }
#pragma unroll
      for (int j = 0; j < SP; ++j) {
        const float v = (j < t.Sb && ok) ? gv[j] * (1.0f - __expf(lv[j])) : 0.0f;
        dbacc[j] += v;
        *reinterpret_cast<__nv_bfloat16*>(dz + j * 128 + (((lc >> 3) ^ (j & 7)) << 4) + ((lc & 7) << 1)) =
            __float2bfloat16(v);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tm_arrive(&dz_full[buf]);
      if (l == 0) TM_STAMP(2, 2);
      if (t.last_of_image) {   // db[wrow_j] += sum over the image's rows handled by this CTA
#pragma unroll
        for (int j = 0; j < SP; ++j) {
          const float s = warp_sum(dbacc[j]);
          if (lane == 0 && j < t.Sb && s != 0.0f) atomicAdd(p.db + __ldg(p.slice_wrow + t.j0 + j), s);
          dbacc[j] = 0.0f;
        }
      }
      ++i;
    }
  } else if (warp < 14) {
    // ------------------------------------------------------------------ epilogue (256 threads)
    const int quad = warp & 3;
    const int ch = (warp - 6) >> 2;
    const int row = quad * 32 + lane;
    const int te = threadIdx.x - 192;           // 0 .. 255
    const bool issuer = (te == 0);
    const int chunks = NB * 4;                  // 16-column chunks of the tile
    const int c_begin = ch == 0 ? 0 : (chunks + 1) / 2, c_end = ch == 0 ? (chunks + 1) / 2 : chunks;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    // column sums of dZ2 (bias gradient of the layer below): warp `box` of the epilogue group sums box `box` of the
    // finished tile from shared memory -- lane = (row quarter, 16-byte piece), eight consecutive lanes read one whole
    // 128-byte row (conflict free) -- and keeps 8 column sums in registers over all tiles of the CTA.  (As 24 MN-major
    // tcgen05 MMAs against a tile of ones this cost ~4600 cycles of tensor-pipe time per tile, and the tile's buffer
    // could not go back to the loads before they had run: clock stamps, DFOL_TBL_TRACE.)
    float cs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] = 0.0f;
    const bool cs_thread = te < NB * 32;
    const int cs_box = te >> 5, cs_rq = (te >> 3) & 3, cs_piece = te & 7;
    int i = 0;
    for (; t.valid(); t.next(p)) {
      t.load(p, SP);
      const long long grow = (long long)p.row0[t.b] + t.c;
      if (t.Sb <= 0) {
        // image without (more) slices: its rows of dZ2 are zero
        const int pieces = NB * 8;
        for (int idx = te; idx < t.cn * pieces; idx += 256) {
          const int r = idx / pieces, pc = idx - r * pieces;
          *reinterpret_cast<uint4*>(p.dZ + (grow + r) * p.lddz + pc * 8) = make_uint4(0u, 0u, 0u, 0u);
        }
        continue;
      }
      const int buf = i & 1;
      const uint32_t n = (uint32_t)(i >> 1);
      if (issuer) TM_STAMP(3, 0);
      mbar_wait(&mma_done, (uint32_t)(i & 1));
      if (issuer) TM_STAMP(3, 1);
      mbar_wait(&h_full[buf], n & 1);   // (already complete: acquires the TMA writes for this thread's generic reads)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (issuer) TM_STAMP(3, 2);
      uint8_t* hb = h_tile(buf);
      uint32_t r[2][16];
      tmem_ld16(trow + TM_ACC_COL + (uint32_t)(16 * c_begin), r[0]);
      const int nch = c_end - c_begin;
#pragma unroll
      for (int cc = 0; cc < 2 * TM_MAX_NB; ++cc) {   // (compile-time register indices: r[] must not go to local memory)
        if (cc >= nch || (p.debug & 2)) break;
        const int c = c_begin + cc;
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (cc + 1 < nch) tmem_ld16(trow + TM_ACC_COL + (uint32_t)(16 * (c + 1)), r[(cc + 1) & 1]);
        const int c0 = 16 * c, box = c0 >> 6, piece = (c0 & 63) >> 3;
        uint8_t* bb = hb + (size_t)box * TM_BOX;
        uint4* q0 = reinterpret_cast<uint4*>(bb + tm_swz(row, piece));
        uint4* q1 = reinterpret_cast<uint4*>(bb + tm_swz(row, piece + 1));
        const uint4 h0 = *q0, h1 = *q1;
        const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[j]));
          const float vx = __uint_as_float(r[cc & 1][2 * j]) * (h.x * (1.0f - h.x));
          const float vy = __uint_as_float(r[cc & 1][2 * j + 1]) * (h.y * (1.0f - h.y));
          const __nv_bfloat162 o = __floats2bfloat162_rn(vx, vy);
          pk[j] = *reinterpret_cast<const uint32_t*>(&o);
        }
        *q0 = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *q1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      if (issuer) TM_STAMP(3, 3);
      if (t.last_of_image) {
        // flush the image's dW rows: block m of the transposed accumulator holds columns 64 * mb + row
        for (int m = ch; m < mblocks; m += 2) {
          const int mb = min(2 * m, NB - 2);
          const int e = 64 * mb + row;
          const bool mine = e >= 128 * m && e < p.E;
#pragma unroll
          for (int q = 0; q < SP / 16; ++q) {
            uint32_t w[16];
            tmem_ld16(trow + TM_DW_COL + (uint32_t)(m * SP + 16 * q), w);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (mine) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int jj = 16 * q + j;
                if (jj < t.Sb) {
                  const float v = __uint_as_float(w[j]);
                  if (v != 0.0f) atomicAdd(p.dW + (long long)__ldg(p.slice_wrow + t.j0 + jj) * p.ldw + e, v);
                }
              }
            }
          }
        }
      }
      // accumulators drained (and flushed): the next tile's MMAs may overwrite them
      if (issuer) TM_STAMP(3, 4);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tm_arrive(&acc_empty);
      // dZ2 tile complete in shared memory -> visible to the async proxy (TMA store) and to the column-sum warps
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tm_named_bar(1, 256);
      if (issuer) TM_STAMP(3, 5);
      if (t.cn == TM_BM && issuer) tm_arrive(&st_ready);   // the store warp stores the tile and frees its buffer
      if (cs_thread) {
        const uint8_t* cb = hb + (size_t)cs_box * TM_BOX;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
          const int r = cs_rq * 32 + k;   // (rows beyond the image's last one hold zeros: their DZ rows are zero)
          const uint4 v = *reinterpret_cast<const uint4*>(cb + tm_swz(r, cs_piece));
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
            cs[2 * j] += f.x;
            cs[2 * j + 1] += f.y;
          }
        }
        __syncwarp();
        if (lane == 0) tm_arrive(&buf_free[buf]);
      }
      if (issuer) TM_STAMP(3, 6);
      if (t.cn != TM_BM) {
        // partial tile at the end of an image: a box store would run into the next image's rows
        const int pieces = NB * 8;
        for (int idx = te; idx < t.cn * pieces; idx += 256) {
          const int rr = idx / pieces, pc = idx - rr * pieces;
          const uint4 v = *reinterpret_cast<const uint4*>(hb + (size_t)(pc >> 3) * TM_BOX + tm_swz(rr, pc & 7));
          *reinterpret_cast<uint4*>(p.dZ + (grow + rr) * p.lddz + pc * 8) = v;
        }
        tm_named_bar(1, 256);
        if (issuer) tm_arrive(&buf_free[buf]);
      }
      ++i;
    }
    if (i > 0) {
      mbar_wait(&all_done, 0);   // (every MMA of the CTA has completed before the accumulators are released)
      if (cs_thread && p.dbelow != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int e = 64 * cs_box + 8 * cs_piece + j;
          if (e < p.E && cs[j] != 0.0f) atomicAdd(p.dbelow + e, cs[j]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ store warp (one lane): dZ2 tile -> global by TMA,
    // off the epilogue's path; the buffer goes back to the loads as soon as the bulk store has read it
    if (lane == 0) {
      int i = 0;
      uint32_t fulls = 0;
      for (; t.valid(); t.next(p)) {
        t.load(p, SP);
        if (t.Sb <= 0) continue;
        const int buf = i & 1;
        if (t.cn == TM_BM) {
          mbar_wait(&st_ready, fulls & 1u);
          ++fulls;
          const long long grow = (long long)p.row0[t.b] + t.c;
          uint8_t* hb = h_tile(buf);
          for (int k = 0; k < ((p.debug & 8) ? 0 : NB); ++k) tm_store_2d(&tmap_z, hb + (size_t)k * TM_BOX, 64 * k, (int)grow);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          tm_arrive(&buf_free[buf]);
        }
        ++i;
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// Wb[(SP b + j), :] = bf16(W[slice_wrow[img_slice[b] + j], :]) for j < slices of image b (zero rows beyond, zero K
// padding): the per-image B operand of MMA-A.
__global__ void __launch_bounds__(256) table_slice_weights_kernel(const float* __restrict__ W, long long ldw, int E,
                                                                  const int32_t* __restrict__ slice_wrow,
                                                                  const int32_t* __restrict__ img_slice, int SP,
                                                                  __nv_bfloat16* __restrict__ Wb, int K) {
  const int b = blockIdx.x;
  const int j0 = img_slice[b];
  const int Sb = min(img_slice[b + 1] - j0, SP);
  for (int idx = threadIdx.x; idx < SP * K; idx += blockDim.x) {
    const int j = idx / K, e = idx - j * K;
    float v = 0.0f;
    if (j < Sb && e < E) v = W[(long long)slice_wrow[j0 + j] * ldw + e];
    Wb[((long long)b * SP + j) * K + e] = __float2bfloat16(v);
  }
}

}  // namespace dfol

using namespace dfol;

extern "C" int dfol_table_layer_bwd_mma(const float* g, const int32_t* slice_goff, const int32_t* slice_col,
                                        const int32_t* slice_wrow, const int32_t* img_slice, int image_num,
                                        int max_slices, const float* ll, const int64_t* blk, const int32_t* stride,
                                        const int32_t* row0, const int32_t* img_rows, const int32_t* tile_start,
                                        int total_tiles, int64_t total_rows, const float* W, int64_t ldw,
                                        const void* h_saved, int64_t ldh, int E, void* dZ, int64_t lddz, int cols,
                                        float* dW, float* db, float* dbelow, void* wb_workspace, void* stream) {
  const char* who = "dfol_table_layer_bwd_mma";
  DFOL_REQUIRE(g && slice_goff && slice_col && slice_wrow && img_slice && ll && blk && stride && row0 && img_rows &&
                   tile_start && W && h_saved && dZ && dW && db && wb_workspace,
               "%s: null pointer", who);
  if (image_num == 0 || total_tiles == 0 || total_rows == 0) return 0;
  DFOL_REQUIRE(max_slices >= 1 && max_slices <= 32, "%s: 1..32 slices per image", who);
  DFOL_REQUIRE(cols >= 128 && cols <= 64 * TM_MAX_NB && (cols % 64) == 0 && E <= cols && ldh >= cols && lddz >= cols &&
                   (ldh % 8) == 0 && (lddz % 8) == 0,
               "%s: 128 <= cols <= %d, multiple of 64, E <= cols <= ldh, lddz (multiples of 8)", who, 64 * TM_MAX_NB);
  DFOL_REQUIRE((reinterpret_cast<uintptr_t>(h_saved) % 16) == 0 && (reinterpret_cast<uintptr_t>(dZ) % 16) == 0 &&
                   (reinterpret_cast<uintptr_t>(wb_workspace) % 16) == 0,
               "%s: operands must be 16-byte aligned", who);
  cudaStream_t st = (cudaStream_t)stream;
  const int SP = max_slices <= 16 ? 16 : 32;
  const int NB = cols / 64;
  __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(wb_workspace);
  table_slice_weights_kernel<<<image_num, 256, 0, st>>>(W, ldw, E, slice_wrow, img_slice, SP, wb, cols);
  alignas(64) CUtensorMap mh, mw, mz;
  int rc = encode_map_bf16(&mh, h_saved, total_rows, cols, ldh, TM_BM);
  if (rc != 0) return rc;
  rc = encode_map_bf16(&mw, wb, (int64_t)image_num * SP, cols, cols, SP);
  if (rc != 0) return rc;
  rc = encode_map_bf16(&mz, dZ, total_rows, cols, lddz, TM_BM);
  if (rc != 0) return rc;
  TmParams p;
  p.g = g; p.slice_goff = slice_goff; p.slice_col = slice_col; p.slice_wrow = slice_wrow; p.img_slice = img_slice;
  p.ll = ll; p.blk = blk; p.stride = stride; p.row0 = row0; p.img_rows = img_rows; p.tile_start = tile_start;
  p.images = image_num; p.total_tiles = total_tiles; p.E = E; p.NB = NB;
  static const int dbg = [] { const char* e = getenv("DFOL_TBL_DEBUG"); return e ? atoi(e) : 0; }();
  p.debug = dbg;
  static const int trace_env = [] { const char* e = getenv("DFOL_TBL_TRACE"); return e ? atoi(e) : 0; }();
  p.trace = nullptr;
  p.trace_first = trace_env > 0 ? trace_env - 1 : 0;
  static const int trace_blk = [] { const char* e = getenv("DFOL_TBL_TRACE_BLOCK"); return e ? atoi(e) : 0; }();
  p.trace_block = trace_blk;
  if (trace_env) {
    cudaMalloc(&p.trace, 4 * 16 * 16 * sizeof(long long));
    cudaMemset(p.trace, 0, 4 * 16 * 16 * sizeof(long long));
  }
  p.dZ = reinterpret_cast<__nv_bfloat16*>(dZ); p.lddz = lddz; p.dW = dW; p.ldw = ldw; p.db = db; p.dbelow = dbelow;
  const size_t wbox = (size_t)SP * 128;
  const size_t smem = 2 * (size_t)NB * (TM_BOX + wbox) + 4 * wbox + 4096 + 1024;
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = total_tiles < sms ? total_tiles : sms;
  cudaError_t e;
  if (SP == 16) {
    e = cudaFuncSetAttribute(table_layer_bwd_mma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return (int)e; }
    table_layer_bwd_mma_kernel<16><<<grid, TM_THREADS, smem, st>>>(mh, mw, mz, p);
  } else {
    e = cudaFuncSetAttribute(table_layer_bwd_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return (int)e; }
    table_layer_bwd_mma_kernel<32><<<grid, TM_THREADS, smem, st>>>(mh, mw, mz, p);
  }
  if (p.trace != nullptr) {   // timing experiment: per-tile stamps of block 0 (cycles relative to the first stamp)
    static long long host[4 * 16 * 16];
    cudaDeviceSynchronize();
    cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    long long t0 = 0;
    for (int i = 0; i < 4 * 16 * 16; ++i) if (host[i] && (!t0 || host[i] < t0)) t0 = host[i];
    static const char* names[4] = {"load", "mma ", "dz  ", "epi "};
    fprintf(stderr, "[%s trace] tiles=%d SP=%d NB=%d block=%d first=%d\n", who, total_tiles, SP, NB, p.trace_block, p.trace_first);
    for (int t = 0; t < 12; ++t)
      for (int r = 0; r < 4; ++r) {
        fprintf(stderr, "  tile %2d %s:", t, names[r]);
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %7lld", host[(r * 16 + t) * 16 + k] ? host[(r * 16 + t) * 16 + k] - t0 : -1);
        fprintf(stderr, "\n");
      }
  }
  return finish_launch(who);
}
