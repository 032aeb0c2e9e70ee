// Microbenchmark: aggregate global->shared bandwidth of cp.async.bulk (1-D TMA copies) as a function of copy size,
// ring depth and CTAs per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_bw bulk_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void stream_kernel(const char* __restrict__ src, size_t bytes_per_cta, int copy_bytes, int depth,
                              float* sink) {
  extern __shared__ __align__(128) char ring[];
  __shared__ __align__(8) uint64_t full[16];
  const char* base = src + (size_t)blockIdx.x * bytes_per_cta;
  const int n = (int)(bytes_per_cta / copy_bytes);
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int k) {
    const int b = k % depth;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[b])), "r"(copy_bytes)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(ring + (size_t)b * copy_bytes)),
                 "l"(base + (size_t)k * copy_bytes), "r"(copy_bytes), "r"(smem_u32(&full[b]))
                 : "memory");
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
    acc += reinterpret_cast<const float*>(ring + (size_t)b * copy_bytes)[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0 && k + depth < n) issue(k + depth);
  }
  if (acc == 123.f) sink[0] = acc;
}

int main() {
  const size_t total = (size_t)3 << 30;
  char* src; float* sink;
  cudaMalloc(&src, total); cudaMalloc(&sink, 4);
  cudaMemset(src, 1, total);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int sizes[] = {4096, 16384, 40960, 65536};
  for (int per_sm = 1; per_sm <= 2; ++per_sm)
    for (int s : sizes)
      for (int depth : {2, 4, 8}) {
        const size_t smem = (size_t)s * depth;
        if (smem * per_sm > 200 * 1024) continue;
        const int ctas = 148 * per_sm;
        size_t per_cta = (total / ctas) / s * s;
        if (per_cta > ((size_t)16 << 20)) per_cta = ((size_t)16 << 20) / s * s;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        stream_kernel<<<ctas, 128, smem>>>(src, per_cta, s, depth, sink);
        cudaEventRecord(e0);
        stream_kernel<<<ctas, 128, smem>>>(src, per_cta, s, depth, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("ctas/SM %d copy %6d B depth %d in-flight/SM %4zu KB : %7.1f GB/s (%s)\n", per_sm, s, depth,
               smem * per_sm / 1024, (double)per_cta * ctas / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
