// Microbenchmark (diagnostics): what does the weight ring of the large-batch chain kernel (chain2_tc.cu) cost per 16 KB stage as a pure TMA
// stream -- a producer thread refilling an n-deep ring of 16 KB slots, a consumer thread releasing each slot `hold_ns` after it landed
// (hold = 0: pure TMA throughput; hold = 180: the two tcgen05.mma of a stage) -- for the access patterns of the kernel and for alternatives:
//   mode 0  forward today : 3-D box {64 n, 32 k, 4 chunks} of the row-major [in][512] bf16 weights, SWIZZLE_128B   (128 rows of 128 B)
//   mode 1  backward today: 2-D box {32 k, 256 n} of the same matrix read as W^T, SWIZZLE_64B                      (256 rows of 64 B)
//   mode 2  pre-tiled     : the 16 KB stage image contiguous in global memory, one 1-D cp.async.bulk
//   mode 3  backward, 64-wide: 2-D box {64 k, 128 n} SWIZZLE_128B (half the requests of mode 1 for the same bytes)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_ring_bench tma_ring_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.u32 %0,1,0,P;\n}" : "=r"(ok) : "r"(su32(bar)), "r"(parity) : "memory");
}
constexpr int STAGE = 16384;
__global__ void __launch_bounds__(64, 1) ring(const __grid_constant__ CUtensorMap m3, const __grid_constant__ CUtensorMap mT, const __grid_constant__ CUtensorMap mT64,
                                              const uint8_t* tiled, int mode, int nstage, int n_total, int hold_ns, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], empty[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; i++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&empty[i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int n = 0; n < n_total; n++) {
      wait(&empty[stage], phase ^ 1);
      const int layer = (n / 32) & 3, h = (n / 16) & 1, ks = n & 15;
      uint8_t* dst = sm + stage * STAGE;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&full[stage])), "r"(STAGE) : "memory");
      if (mode == 0) {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(su32(dst)), "l"((uint64_t)&m3), "r"(su32(&full[stage])), "r"(0), "r"(layer * 512 + ks * 32), "r"(h * 4) : "memory");
      } else if (mode == 1) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(su32(dst)), "l"((uint64_t)&mT), "r"(su32(&full[stage])), "r"(ks * 32), "r"(layer * 512 + h * 256) : "memory");
      } else if (mode == 3) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(su32(dst)), "l"((uint64_t)&mT64), "r"(su32(&full[stage])), "r"((ks & 7) * 64), "r"(layer * 512 + h * 256 + (ks >> 3) * 128) : "memory");
      } else {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(su32(dst)), "l"(tiled + (size_t)(n & 127) * STAGE), "r"(STAGE), "r"(su32(&full[stage])) : "memory");
      }
      if (++stage == nstage) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    unsigned long long t0 = 0;
    for (int n = 0; n < n_total; n++) {
      wait(&full[stage], phase);
      if (n == 8) t0 = gt();
      if (hold_ns > 0) { const unsigned long long t = gt(); while (gt() - t < (unsigned long long)hold_ns) { } }
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(&empty[stage])) : "memory");
      if (++stage == nstage) { stage = 0; phase ^= 1; }
    }
    out[blockIdx.x] = gt() - t0;
  }
}
typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncFn enc = (EncFn)fn;
  const int rows = 2048;
  uint16_t* W; CK(cudaMalloc(&W, (size_t)rows * 512 * 2)); CK(cudaMemset(W, 0, (size_t)rows * 512 * 2));
  uint8_t* tiled; CK(cudaMalloc(&tiled, 128 * STAGE)); CK(cudaMemset(tiled, 0, 128 * STAGE));
  unsigned long long* out; CK(cudaMalloc(&out, 256 * 8));
  CUtensorMap m3, mT, mT64;
  {
    cuuint64_t dims[3] = {64, (cuuint64_t)rows, 8}; cuuint64_t strides[2] = {1024, 128}; cuuint32_t box[3] = {64, 32, 4}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, W, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("3-D map refused (%d)\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {512, (cuuint64_t)rows}; cuuint64_t strides[1] = {1024}; cuuint32_t box[2] = {32, 256}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mT, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("2-D map refused (%d)\n", (int)r); return 1; }
    cuuint32_t box64[2] = {64, 128};
    r = enc(&mT64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, dims, strides, box64, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("2-D 64-wide map refused (%d)\n", (int)r); return 1; }
  }
  CK(cudaFuncSetAttribute(ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * STAGE + 1024));
  const int n_total = 32 * 40 + 8;
  const char* names[4] = {"fwd 3-D box, 128 x 128 B ", "bwd 2-D box, 256 x 64 B  ", "pre-tiled, 1-D bulk 16 KB", "bwd 2-D box, 128 x 128 B "};
  for (int hold : {0, 180})
    for (int ncta : {1, 128})
      for (int mode = 0; mode < 4; mode++) {
        printf("hold %3d ns, %3d CTAs, %s:", hold, ncta, names[mode]);
        for (int nstage : {2, 3, 4, 5, 8}) {
          double worst = 0;
          for (int rep = 0; rep < 3; rep++) {
            ring<<<ncta, 64, 8 * STAGE + 1024>>>(m3, mT, mT64, tiled, mode, nstage, n_total, hold, out);
            CK(cudaDeviceSynchronize());
            unsigned long long h[128]; CK(cudaMemcpy(h, out, ncta * 8, cudaMemcpyDeviceToHost));
            double mx = 0; for (int i = 0; i < ncta; i++) mx = h[i] > mx ? (double)h[i] : mx;
            worst = mx;
          }
          printf("  %d-deep %6.1f ns/stage", nstage, worst / (n_total - 8));
        }
        printf("\n");
      }
  return 0;
}
