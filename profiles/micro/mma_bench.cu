// Microbenchmark (diagnostics): cost of back-to-back tcgen05.mma kind::f16 (bf16, M=128, K=16, SS mode) as a function of N
// and of the number of independent accumulators.  One CTA, one issuing thread, operands resident in smem (SWIZZLE_128B).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void __launch_bounds__(128, 1) bench(int N, int nacc, int n_mma, int b_mn, unsigned long long* out, int nwarps) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 49152 / 4; i += 128) ((uint32_t*)sm)[i] = 0x3c003c00;  // finite bf16 data
  if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar[i]))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(su32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < nwarps) {
    const int w = threadIdx.x >> 5;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t ad = desc(su32(sm), 16, 1024);
    const uint64_t bd = b_mn ? desc(su32(sm + 16384), 2048, 1024) : desc(su32(sm + 16384), 16, 1024);
    for (int rep = 0; rep < 3; rep++) {
      unsigned long long t0 = gt();
      for (int i = 0; i < n_mma; i++) {
        const uint32_t d = tm + ((i % nacc) + w * nacc) * N;
        const uint32_t acc = i >= nacc;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(ad + (uint64_t)((i & 3) * 2)), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
      unsigned long long t1 = gt();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(&bar[w])) : "memory");
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0,1,0,P;\n}" : "=r"(ok) : "r"(su32(&bar[w])), "r"(rep & 1) : "memory");
      unsigned long long t2 = gt();
      out[2 * w] = t1 - t0; out[2 * w + 1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}
int main() {
  unsigned long long* out; CK(cudaMalloc(&out, 64));
  CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const int n_mma = 64;
  // one issuing warp, N = 256 (the chain kernels' MMA): 64 MMAs into ONE accumulator vs alternating between TWO accumulators
  for (int nacc : {1, 2}) {
    bench<<<1, 128, 65536>>>(256, 1, n_mma, nacc, out, 1);
    CK(cudaDeviceSynchronize());
    unsigned long long h[8]; CK(cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost));
    printf("N=256 issuing warps=1 accumulators=%d: %d MMAs issued in %7.1f ns, complete after %7.1f ns -> %5.1f ns per MMA (math floor 65 ns)\n", nacc, n_mma,
           (double)h[0], (double)h[1], (double)h[1] / n_mma);
  }
  for (int N : {64, 128})
    for (int nw : {1, 2, 4}) {
      if (N * nw > 512) continue;
      bench<<<1, 128, 65536>>>(N, 1, n_mma, 1, out, nw);
      CK(cudaDeviceSynchronize());
      unsigned long long h[8]; CK(cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost));
      double mx = 0; for (int w = 0; w < nw; w++) mx = h[2 * w + 1] > mx ? h[2 * w + 1] : mx;
      printf("N=%3d issuing warps=%d: %d MMAs per warp, all complete after %7.1f ns -> %5.1f ns per MMA aggregate (%.0f cycles; math floor %d)\n", N, nw, n_mma, mx,
             mx / (n_mma * nw), mx / (n_mma * nw) * 1.965, N / 2);
    }
  return 0;
}
