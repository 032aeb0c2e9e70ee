// Microbenchmark: aggregate bandwidth of 2-D tensor-map TMA loads (bf16, 128-byte swizzle, 64-column boxes) streaming a
// row-major [rows x cols] matrix tile by tile (128-row tiles, all column boxes of a tile in flight), as the pair-layer
// GEMMs do.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma2d_bw tma2d_bw.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void stream2d(const __grid_constant__ CUtensorMap map, int tiles, int boxes, int box_rows, int depth,
                         float* sink) {
  extern __shared__ __align__(1024) char ring[];
  __shared__ __align__(8) uint64_t full[16];
  const uint32_t box_bytes = (uint32_t)box_rows * 128u;
  const uint32_t tile_bytes = box_bytes * boxes * (128 / box_rows);
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int n = 0;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) ++n;
  auto issue = [&](int k) {
    const int tile = blockIdx.x + k * gridDim.x;
    const int b = k % depth;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[b])), "r"(tile_bytes)
                 : "memory");
    char* dst = ring + (size_t)b * tile_bytes;
    for (int j = 0; j < boxes; ++j)
      for (int r = 0; r < 128 / box_rows; ++r) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&full[b])), "r"(j * 64),
            "r"(tile * 128 + r * box_rows)
            : "memory");
        dst += box_bytes;
      }
  };
  if (threadIdx.x == 0)
    for (int k = 0; k < depth && k < n; ++k) issue(k);
  float acc = 0.f;
  for (int k = 0; k < n; ++k) {
    const int b = k % depth;
    const uint32_t parity = (k / depth) & 1;
    uint32_t done = 0;
    while (!done)
      asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}"
                   : "=r"(done) : "r"(smem_u32(&full[b])), "r"(parity) : "memory");
    acc += reinterpret_cast<const float*>(ring + (size_t)b * tile_bytes)[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0 && k + depth < n) issue(k + depth);
  }
  if (acc == 123.f) sink[0] = acc;
}

int main() {
  const long long rows = 589824 * 2, cols = 256;
  void* src; float* sink;
  cudaMalloc(&src, rows * cols * 2); cudaMalloc(&sink, 4);
  cudaMemset(src, 1, rows * cols * 2);
  cudaFuncSetAttribute(stream2d, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* sym; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  Enc enc = (Enc)sym;
  for (int box_rows : {128, 64})
    for (int boxes : {4, 2})
      for (int depth : {1, 2, 3}) {
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
        cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
        cuuint32_t es[2] = {1, 1};
        enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const size_t tile_bytes = (size_t)128 * 128 * boxes;
        const size_t smem = tile_bytes * depth;
        if (smem > 200 * 1024) continue;
        const int tiles = (int)(rows / 128);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        stream2d<<<148, 128, smem>>>(map, tiles, boxes, box_rows, depth, sink);
        cudaEventRecord(e0);
        stream2d<<<148, 128, smem>>>(map, tiles, boxes, box_rows, depth, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("box %3d rows x 64 cols, %d column boxes/tile (%3zu KB/tile), depth %d (in flight %3zu KB/SM): %7.1f GB/s (%s)\n",
               box_rows, boxes, tile_bytes / 1024, depth, smem / 1024, (double)tiles * tile_bytes / ms / 1e6,
               cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
