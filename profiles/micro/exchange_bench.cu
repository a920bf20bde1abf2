// Microbenchmark (diagnostics, not product): how fast can the 8 CTAs of a cluster all-gather 8 x 16 KB blocks?
//   mode 0: every CTA TMA-loads all 8 blocks from global/L2 (unicast)                     [what euler_cluster v1 does]
//   mode 1: every CTA TMA-loads ITS block with .multicast::cluster to all 8 CTAs
//   mode 2: every CTA bulk-copies its block smem -> smem of all 8 peers (cp.async.bulk.shared::cluster.shared::cta)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exchange_bench exchange_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cudaTypedefs.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)
constexpr int NC = 8, BLK = 16384, ROUNDS = 200;

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\nselp.u32 %0,1,0,P;\n}" : "=r"(ok) : "r"(su32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release;\nbarrier.cluster.wait.acquire;" ::: "memory"); }

__global__ void __cluster_dims__(NC, 1, 1) __launch_bounds__(128, 1)
bench(const __grid_constant__ CUtensorMap map, int mode, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = sm;                  // [8][16 KB]
  uint8_t* sSrc = sm + NC * BLK;     // 16 KB staging (mode 2)
  uint64_t* full = (uint64_t*)(sSrc + BLK);   // [8]
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) { for (int i = 0; i < NC; i++) mbar_init(&full[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  csync();
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) {
    t0 = gt();
    for (int r = 0; r < ROUNDS; r++) {
      for (int i = 0; i < NC; i++) mbar_expect(&full[i], BLK);
      if (mode == 0) {
        for (int i = 0; i < NC; i++)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(su32(sA + i * BLK)), "l"((uint64_t)&map), "r"(su32(&full[i])), "r"(i * 64), "r"((int)(blockIdx.x / NC) * 128) : "memory");
      } else if (mode == 1) {
        uint16_t mask = 0xFF;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                     ::"r"(su32(sA + rank * BLK)), "l"((uint64_t)&map), "r"(su32(&full[rank])), "r"((int)rank * 64), "r"((int)(blockIdx.x / NC) * 128), "h"(mask) : "memory");
      } else {
        for (uint32_t p = 0; p < NC; p++) {
          uint32_t dst, bar;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(su32(sA + rank * BLK)), "r"(p));
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar) : "r"(su32(&full[rank])), "r"(p));
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "r"(su32(sSrc)), "r"(BLK), "r"(bar) : "memory");
        }
      }
      for (int i = 0; i < NC; i++) mbar_wait(&full[i], r & 1);
      // round barrier so that nobody overwrites a peer's buffers early (cheap remote-free emulation)
      asm volatile("barrier.cluster.arrive.release;\nbarrier.cluster.wait.acquire;" ::: "memory");
    }
    out[blockIdx.x] = gt() - t0;
  } else {
    for (int r = 0; r < ROUNDS; r++) asm volatile("barrier.cluster.arrive.release;\nbarrier.cluster.wait.acquire;" ::: "memory");
  }
  csync();
}

int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  auto enc = (PFN_cuTensorMapEncodeTiled_v12000)p;
  const int clusters = 2;
  uint16_t* g; CK(cudaMalloc(&g, (size_t)clusters * 128 * 512 * 2)); CK(cudaMemset(g, 0, (size_t)clusters * 128 * 512 * 2));
  CUtensorMap map; cuuint64_t dims[2] = {512, (cuuint64_t)clusters * 128}; cuuint64_t st[1] = {1024}; cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  unsigned long long* out; CK(cudaMalloc(&out, 64 * 8));
  const int smem = NC * BLK + BLK + 1024 + 256;
  CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int mode = 0; mode < 3; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      bench<<<NC * clusters, 128, smem>>>(map, mode, out);
      CK(cudaDeviceSynchronize());
    }
    unsigned long long h[64]; CK(cudaMemcpy(h, out, NC * clusters * 8, cudaMemcpyDeviceToHost));
    double mx = 0; for (int i = 0; i < NC * clusters; i++) mx = h[i] > mx ? h[i] : mx;
    printf("mode %d: %.3f us per all-gather round of 8 x 16 KB per CTA (incl. one cluster barrier)  -> %.1f GB/s received per SM\n", mode,
           mx / ROUNDS / 1e3, 128.0 * 1024 / (mx / ROUNDS));
  }
  // barrier-only baseline
  return 0;
}
