// Pair-level forward chain of the relation head in ONE kernel (sm_100a):
//   H1[(s,o), :] = elu(U[s] + V[o] + Wg . geo(s,o) + b1)          (first Linear of the relation network through the U/V
//                                                                   decomposition, batch_gqa_boxfeatures_pipeline.py:257-279,
//                                                                   gqa_interpreter_experiments.py:167)
//   H2           = sigmoid(H1 . W2^T + b2)                         (second Linear, classifier_oracle.py:154)
// The separate kernels wrote H1 (pair_hidden_fwd_tc) and read it back as the A operand of the layer-2 GEMM
// (gemm_bf16_tc_cluster_kernel): 2 x P x 512 bytes of HBM traffic for a tensor that is produced from ~1 KB of per-object
// data per row.  Here the A operand is PRODUCED IN SHARED MEMORY: four producer warps per CTA evaluate the hidden layer
// for the tile's rows straight into the 128-byte-swizzled K-major stage buffers the tcgen05 MMAs read, and (training
// only) also store it to HBM for the backward pass -- the forward pass never reads H1 from memory.
//
// Same cluster-of-two structure as gemm_bf16_tc_cluster_kernel: the two CTAs of a cluster split the N columns of W2
// (resident in shared memory) and need the SAME 128-row A tile.  Each CTA produces 64 of the 128 rows and hands its half
// to the peer with a bulk shared::cta -> shared::cluster copy that completes on the peer's "stage full" mbarrier, so
// every hidden-layer element is computed once per cluster.
//   warp 0 (one lane) : TMA load of this CTA's rows of W2 (once)
//   warp 1 (one lane) : tcgen05.mma issuer, fp32 accumulators double buffered in TMEM, commit.multicast frees a stage
//                       in both CTAs
//   warps 2-5         : producers (lane = 8 hidden units of the row, warp w = rows w, w+4, ...; the V rows of 8 rows in
//                       flight per thread)
//   warps 6-13        : epilogue (bias, sigmoid, bf16, swizzled staging tile, TMA store) -- as in the cluster GEMM
// The arithmetic of a hidden-layer element is the SIMT kernel's, operation for operation (packed fp32 pairs, one-MUFU
// ELU), and the MMA chain is the cluster GEMM's: H1, geo and H2 are bit-identical to the unfused path.
#include <cstdlib>
#include "tc_common.cuh"

namespace dfol {

constexpr int PC_BM = 128;
constexpr int PC_BK = 64;
constexpr int PC_KB = 4;                       // K = 256 hidden units = 4 stages of 64
constexpr int PC_PROD = 128;                   // producer threads (4 warps)
constexpr int PC_THREADS = 64 + PC_PROD + 256;
constexpr int PC_MAX_CH = 6;

struct PcParams {
  const float* uv; long long lduv;             // [T, 2 * 256] fp32: U | V
  const float* pos; long long ldpos;           // [T, >= 4] normalised box (x, y, w, h)
  const float* wg; long long ldw;              // geometry weights wg[h * ldw + c], c < 4
  const float* bias1;                          // [256]
  const float* bias2;                          // [N]
  const int32_t* pair_img; const int32_t* pair_row; const int32_t* obj_row; const int32_t* img_n;
  __nv_bfloat16* h1_out; long long ldh1;       // [M, 256] or NULL (inference)
  float4* geo_out;                             // [M] or NULL
  int M, N, BNh, n_store;
  int debug;   // ablation switches (DFOL_PC_DEBUG; wrong results): 1 no row tables, 2 no hidden-layer math, 4 no H1 store, 8 no peer copy
};

__device__ __forceinline__ uint32_t pc_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void pc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void pc_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pc_named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint32_t pc_mapa(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// bulk copy of this CTA's shared memory into the peer's, completing (as transaction bytes) on the PEER's mbarrier
__device__ __forceinline__ void pc_bulk_to_peer(uint32_t dst_cluster_addr, const void* src, uint32_t bytes,
                                                uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster_addr), "r"(smem_u32(src)), "r"(bytes), "r"(bar_cluster_addr)
               : "memory");
}
__device__ __forceinline__ void pc_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void pc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t pc_swz(int row, int piece) {
  return (uint32_t)(row * 128 + ((piece ^ (row & 7)) << 4));
}
// packed fp32 pair arithmetic (both halves IEEE round-to-nearest: bit-identical to the scalar operations)
__device__ __forceinline__ uint64_t pc_pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void pc_unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t pc_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t pc_add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float pc_elu(float a) {   // (elu_fast of tc_support.cu)
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * 1.4426950408889634f));
  return a > 0.0f ? a : e - 1.0f;
}
__device__ __forceinline__ float pc_sigmoid(float x) {   // (cl_act<SIGMOID> of the cluster GEMM)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float4 pc_geometry(const float4 ps, const float4 po) {   // (pair_geometry4 of tc_support.cu)
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

__global__ void __launch_bounds__(PC_THREADS, 1)
    pair_chain_fwd_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_c,
                          PcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t b_full;
  __shared__ __align__(8) uint64_t a_full[PC_KB];
  __shared__ __align__(8) uint64_t a_empty[PC_KB];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float bias_s[192];
  __shared__ __align__(16) float4 geo_s[64];
  __shared__ int2 rowinfo_s[64];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = pc_ctarank();
  constexpr uint32_t CS = 2;
  const uint16_t cmask = 3;
  const int cluster_id = blockIdx.x / (int)CS, num_clusters = gridDim.x / (int)CS;
  const int num_tiles = (p.M + PC_BM - 1) / PC_BM;
  const int BNh = p.BNh;
  const int nbox = BNh / 64;
  const int n0 = (int)rank * BNh;
  constexpr uint32_t a_bytes = PC_BM * PC_BK * 2;           // 16 KB per stage
  const uint32_t b_kb_bytes = (uint32_t)BNh * PC_BK * 2;
  constexpr uint32_t box_bytes = PC_BM * 128;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_tiles = base;
  uint8_t* a_tiles = b_tiles + (size_t)PC_KB * b_kb_bytes;
  uint8_t* c_tile = a_tiles + (size_t)PC_KB * a_bytes;

  if (threadIdx.x == 0) {
    mbar_init(&b_full, 1);
    for (int s = 0; s < PC_KB; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], CS); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  for (int i = threadIdx.x; i < 192; i += PC_THREADS)
    bias_s[i] = (p.bias2 != nullptr && i < BNh && n0 + i < p.N) ? __ldg(p.bias2 + n0 + i) : 0.0f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  pc_cluster_sync();  // both CTAs' barriers are initialised before any remote copy / arrival
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
      mbar_expect_tx(&b_full, (uint32_t)PC_KB * b_kb_bytes);
      for (int kb = 0; kb < PC_KB; ++kb)
        tma_load_2d(&tmap_b, &b_full, b_tiles + (size_t)kb * b_kb_bytes, kb * PC_BK, n0);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BNh >> 3) << 17) |
                             ((uint32_t)(PC_BM >> 4) << 24);
      mbar_wait(&b_full, 0);
      int local = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
        const int buf = local & 1;
        const uint32_t use = (uint32_t)(local >> 1);
        mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + (uint32_t)(buf * 256);
        for (int kb = 0; kb < PC_KB; ++kb) {
          mbar_wait(&a_full[kb], (uint32_t)(local & 1));   // both halves of the stage: own rows + the peer's copy
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_smem_desc(smem_u32(a_tiles + (size_t)kb * a_bytes));
          const uint64_t db = make_smem_desc(smem_u32(b_tiles + (size_t)kb * b_kb_bytes));
#pragma unroll
          for (int k = 0; k < PC_BK / 16; ++k)
            umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          pc_commit_mc(&a_empty[kb], cmask);   // free in BOTH CTAs once every CTA's MMAs have read it
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp < 2 + PC_PROD / 32) {
    // ------------------------------------------------------------------ producers (128 threads)
    // lane = 8 hidden units = one 16-byte piece of the row (the 32 lanes of a warp cover the 256 units); warp pw owns the
    // rows pw, pw + 4, ... of this CTA's 64 rows (16 rows).  A thread keeps the geometry weights of its 8 units in
    // registers and has the V rows of 8 of its rows in flight at a time: the loads of the second half are issued before
    // the first half is evaluated, so the L2 latency of U / V is paid once per tile, not once per row.
    const int pt = threadIdx.x - 64;             // 0 .. 127
    const int pw = pt >> 5;
    const int h0 = 8 * lane;
    const int kb_mine = lane >> 3, piece = lane & 7;
    uint64_t bz[4], wk[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      bz[q] = pc_pack2(__ldg(p.bias1 + h0 + 2 * q), __ldg(p.bias1 + h0 + 2 * q + 1));
      const float* w0 = p.wg + (long long)(h0 + 2 * q) * p.ldw;
#pragma unroll
      for (int c = 0; c < 4; ++c) wk[c][q] = pc_pack2(__ldg(w0 + c), __ldg(w0 + p.ldw + c));
    }
    const uint32_t peer = rank ^ 1u;
    int local = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
      const int row0 = tile * PC_BM + 64 * (int)rank;   // first of this CTA's 64 rows
      // the four stages must have been read by the MMAs of both CTAs (which also means the peer has received the
      // copies this CTA sent out of them)
      for (int kb = 0; kb < PC_KB; ++kb) mbar_wait(&a_empty[kb], (uint32_t)(local & 1) ^ 1u);
      if (pt < 64) {
        const int row = row0 + pt;
        int2 info = make_int2(-1, -1);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < p.M && (p.debug & 1)) {
          info = make_int2(row & 1023, (row >> 3) & 1023);
        } else if (row < p.M) {
          const int b = __ldg(p.pair_img + row);
          const int n = __ldg(p.img_n + b);
          const int l = row - __ldg(p.pair_row + b);
          const int s = l / n, o = l - s * n;
          const int t0 = __ldg(p.obj_row + b);
          info = make_int2(t0 + s, t0 + o);
          if (s != o) {
            const float4 ps = __ldg(reinterpret_cast<const float4*>(p.pos + (long long)(t0 + s) * p.ldpos));
            const float4 po = __ldg(reinterpret_cast<const float4*>(p.pos + (long long)(t0 + o) * p.ldpos));
            g = pc_geometry(ps, po);
          }
          if (p.geo_out != nullptr) p.geo_out[row] = g;
        }
        rowinfo_s[pt] = info;
        geo_s[pt] = g;
      }
      pc_named_bar(2, PC_PROD);
      uint8_t* stage = a_tiles + (size_t)kb_mine * a_bytes;
      const float* vbase = p.uv + PC_KB * PC_BK + h0;
      int prev_ts = -2;
      uint64_t u[4] = {0ull, 0ull, 0ull, 0ull};
      auto row_out = [&](int rl, const int2 info, const float4 va, const float4 vb) {
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (info.x >= 0 && !(p.debug & 2)) {
          if (info.x != prev_ts) {
            const float4* up = reinterpret_cast<const float4*>(p.uv + (long long)info.x * p.lduv + h0);
            const float4 a = __ldg(up), b = __ldg(up + 1);
            u[0] = pc_add2(pc_pack2(a.x, a.y), bz[0]); u[1] = pc_add2(pc_pack2(a.z, a.w), bz[1]);
            u[2] = pc_add2(pc_pack2(b.x, b.y), bz[2]); u[3] = pc_add2(pc_pack2(b.z, b.w), bz[3]);
            prev_ts = info.x;
          }
          const float4 g = geo_s[rl];
          const uint64_t gp[4] = {pc_pack2(g.x, g.x), pc_pack2(g.y, g.y), pc_pack2(g.z, g.z), pc_pack2(g.w, g.w)};
          uint64_t acc[4] = {pc_add2(u[0], pc_pack2(va.x, va.y)), pc_add2(u[1], pc_pack2(va.z, va.w)),
                             pc_add2(u[2], pc_pack2(vb.x, vb.y)), pc_add2(u[3], pc_pack2(vb.z, vb.w))};
          uint32_t pk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[q] = pc_fma2(wk[c][q], gp[c], acc[q]);
            float x0, x1;
            pc_unpack2(acc[q], x0, x1);
            const __nv_bfloat162 hv = __floats2bfloat162_rn(pc_elu(x0), pc_elu(x1));
            pk[q] = *reinterpret_cast<const uint32_t*>(&hv);
          }
          out = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          if (p.h1_out != nullptr && !(p.debug & 4))
            *reinterpret_cast<uint4*>(p.h1_out + (long long)(row0 + rl) * p.ldh1 + h0) = out;
        }
        *reinterpret_cast<uint4*>(stage + pc_swz(64 * (int)rank + rl, piece)) = out;
      };
      // 16 rows per thread in 4 batches of 4: the V rows of batch b + 1 are requested before batch b is evaluated
      float4 va0[4], vb0[4], va1[4], vb1[4];
      int2 inf0[4], inf1[4];
#define PC_LOAD(INF, VA, VB, BT)                                                                         \
  _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                                        \
    INF[k] = rowinfo_s[pw + 4 * (4 * (BT) + k)];                                                         \
    VA[k] = VB[k] = make_float4(0.f, 0.f, 0.f, 0.f);                                                     \
    if (INF[k].x >= 0) {                                                                                 \
      const float4* vp = reinterpret_cast<const float4*>(vbase + (long long)INF[k].y * p.lduv);          \
      VA[k] = __ldg(vp); VB[k] = __ldg(vp + 1);                                                          \
    }                                                                                                    \
  }
#define PC_EVAL(INF, VA, VB, BT)                                                                         \
  _Pragma("unroll") for (int k = 0; k < 4; ++k) row_out(pw + 4 * (4 * (BT) + k), INF[k], VA[k], VB[k]);
      PC_LOAD(inf0, va0, vb0, 0)
      PC_LOAD(inf1, va1, vb1, 1)
      PC_EVAL(inf0, va0, vb0, 0)
      PC_LOAD(inf0, va0, vb0, 2)
      PC_EVAL(inf1, va1, vb1, 1)
      PC_LOAD(inf1, va1, vb1, 3)
      PC_EVAL(inf0, va0, vb0, 2)
      PC_EVAL(inf1, va1, vb1, 3)
#undef PC_LOAD
#undef PC_EVAL
      // this CTA's half of the four stages is complete: make it visible to the async proxy (MMA, bulk copy), then one
      // thread publishes it locally and ships it to the peer
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      pc_named_bar(2, PC_PROD);
      if (pt == 0) {
        for (int kb = 0; kb < PC_KB; ++kb) {
          uint8_t* half = a_tiles + (size_t)kb * a_bytes + (size_t)rank * (a_bytes / 2);
          if (p.debug & 8) { pc_arrive(&a_full[kb]); continue; }
          mbar_expect_tx(&a_full[kb], a_bytes / 2);   // (arrival + the bytes the PEER sends into this stage)
          pc_bulk_to_peer(pc_mapa(smem_u32(half), peer), half, a_bytes / 2, pc_mapa(smem_u32(&a_full[kb]), peer));
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: the last 8 warps
    constexpr int EW0 = 2 + PC_PROD / 32;        // first epilogue warp (a multiple of 2: quadrants via warp & 3)
    const int quad = warp & 3;
    const int ch = (warp - EW0) >> 2;
    const int row = quad * 32 + lane;
    const int cw = BNh / 2;
    const int cbeg = ch * cw;
    const int nchunks = cw / 16;
    const bool issuer = (warp == EW0 && lane == 0);
    int local = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
      const int buf = local & 1;
      const uint32_t use = (uint32_t)(local >> 1);
      mbar_wait(&acc_full[buf], use & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // the previous tile's TMA store must have finished READING the staging tile before it is overwritten
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      pc_named_bar(1, 256);
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 256 + cbeg);
      uint32_t r[2][16];
      tmem_ld16(trow, r[0]);
#pragma unroll
      for (int c = 0; c < PC_MAX_CH; ++c) {
        if (c >= nchunks) break;
        const int c0 = cbeg + 16 * c;
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c + 1 < nchunks) tmem_ld16(trow + (uint32_t)(16 * (c + 1)), r[(c + 1) & 1]);
        const int box = c0 >> 6, piece = (c0 & 63) >> 3;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = pc_sigmoid(__uint_as_float(r[c & 1][j]) + bias_s[c0 + j]);
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
        uint8_t* cb = c_tile + (size_t)box * box_bytes;
        *reinterpret_cast<uint4*>(cb + pc_swz(row, piece)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(cb + pc_swz(row, piece + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) pc_arrive(&acc_empty[buf]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      pc_named_bar(1, 256);
      if (issuer) {
        const int my_cols = min(max(p.n_store - n0, 0), BNh);
        const int my_boxes = (my_cols + 63) / 64;
        for (int j = 0; j < my_boxes; ++j) pc_store_2d(&tmap_c, c_tile + (size_t)j * box_bytes, n0 + 64 * j, tile * PC_BM);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  pc_cluster_sync();  // the peer may still copy into this CTA's ring / arrive on its barriers until it is done
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

}  // namespace dfol

using namespace dfol;

extern "C" int dfol_pair_chain_fwd(const float* uv, int64_t lduv, const float* obj_pos, int64_t ldpos, const float* wg,
                                   int64_t ldw, const float* bias1, const void* W2, int64_t ldw2, const float* bias2,
                                   void* H2, int64_t ldh2, int store_cols, void* h1_out, int64_t ldh1, void* geo_out,
                                   const int32_t* pair_img, const int32_t* pair_row, const int32_t* obj_row,
                                   const int32_t* img_n, int64_t M, int N, int H, void* stream) {
  const char* who = "dfol_pair_chain_fwd";
  DFOL_REQUIRE(uv && obj_pos && wg && bias1 && W2 && H2 && pair_img && pair_row && obj_row && img_n, "%s: null pointer",
               who);
  if (M == 0) return 0;
  DFOL_REQUIRE(H == PC_KB * PC_BK, "%s: the hidden width must be %d", who, PC_KB * PC_BK);
  DFOL_REQUIRE(M > 0 && M < (1ll << 31) && N > 0, "%s: bad sizes", who);
  DFOL_REQUIRE((lduv % 4) == 0 && (ldpos % 4) == 0 && (reinterpret_cast<uintptr_t>(uv) % 16) == 0 &&
                   (reinterpret_cast<uintptr_t>(obj_pos) % 16) == 0,
               "%s: U|V and the positions are read as float4 (16-byte aligned rows)", who);
  DFOL_REQUIRE((ldw2 % 8) == 0 && ldw2 >= H && (ldh2 % 8) == 0 && (reinterpret_cast<uintptr_t>(W2) % 16) == 0 &&
                   (reinterpret_cast<uintptr_t>(H2) % 16) == 0,
               "%s: bf16 operands must be 16-byte aligned with strides %% 8 == 0", who);
  DFOL_REQUIRE(h1_out == nullptr || ((ldh1 % 8) == 0 && ldh1 >= H && (reinterpret_cast<uintptr_t>(h1_out) % 16) == 0),
               "%s: H1 rows must be 16-byte aligned", who);
  const int n_store = store_cols > 0 ? store_cols : (int)ldh2;
  DFOL_REQUIRE(n_store >= N && n_store <= ldh2 && (n_store % 8) == 0, "%s: N <= store_cols <= ldh2, multiple of 8", who);
  PcParams p;
  p.uv = uv; p.lduv = lduv; p.pos = obj_pos; p.ldpos = ldpos; p.wg = wg; p.ldw = ldw; p.bias1 = bias1; p.bias2 = bias2;
  p.pair_img = pair_img; p.pair_row = pair_row; p.obj_row = obj_row; p.img_n = img_n;
  p.h1_out = reinterpret_cast<__nv_bfloat16*>(h1_out); p.ldh1 = ldh1; p.geo_out = reinterpret_cast<float4*>(geo_out);
  p.M = (int)M; p.N = N; p.n_store = n_store;
  static const int dbg = [] { const char* e = getenv("DFOL_PC_DEBUG"); return e ? atoi(e) : 0; }();
  p.debug = dbg;
  p.BNh = ((n_store + 1) / 2 + 63) / 64 * 64;
  DFOL_REQUIRE(p.BNh <= 192 && n_store > p.BNh, "%s: 192 < stored columns <= 384 (two CTAs split them)", who);
  const size_t b_bytes = (size_t)PC_KB * p.BNh * PC_BK * 2;
  const size_t smem = b_bytes + (size_t)PC_KB * PC_BM * PC_BK * 2 + (size_t)(p.BNh / 64) * PC_BM * 128 + 1024;
  DFOL_REQUIRE(smem <= 225 * 1024, "%s: does not fit in shared memory", who);
  {
    cudaError_t e = cudaFuncSetAttribute(pair_chain_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return (int)e; }
  }
  alignas(64) CUtensorMap mb, mc;
  int rc = encode_map_bf16(&mb, W2, N, H, ldw2, p.BNh);
  if (rc != 0) return rc;
  rc = encode_map_bf16(&mc, H2, M, n_store, ldh2, PC_BM);
  if (rc != 0) return rc;
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles = (int)((M + PC_BM - 1) / PC_BM);
  int clusters = sms / 2;
  if (clusters > tiles) clusters = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters, 1, 1);
  cfg.blockDim = dim3(PC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  {
    cudaError_t e = cudaLaunchKernelEx(&cfg, pair_chain_fwd_kernel, mb, mc, p);
    if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return (int)e; }
  }
  return finish_launch(who);
}
