// dp_comm.cu -- the gradient exchange of the data-parallel step as kernels of this library over NVLink peer memory / NVLS
// multicast (no NCCL on the step).  The reference is single-device: this replaces nothing in it, it is what makes
// "larger-batch runs partitioned across the 8 B200s" (north_star) one CUDA graph per step.
//
//   dp_reduce_kernel<MC>: one launch per gradient bucket, as soon as that network's gradients are final on this rank.
//     CTA b of rank r   1. tells CTA b of every peer "my gradients are final" (st.release.sys into the peer's flag pad) and waits
//                          for the same word from every peer (ld.acquire.sys on its own pad);
//                       2. reduces ITS 1/world share of the bucket:  MC: multimem.ld_reduce.add.v4.f32 through the NVSwitch
//                          (one load returns the sum over all ranks' arenas), multimem.st.v4.f32 of the sum into every arena;
//                          !MC: loads from the `world` peer mappings summed in rank order, stores to every peer mapping;
//                       3. fence.sys, second flag round: when the kernel ends on a rank, every rank's share has landed in its
//                          arena (the optimizer pass that follows in the stream reads plain local memory).
//     Each element is summed exactly once (by its owner) and the same bits go to all ranks: replicas stay bit-identical.
//   Flags are monotonically increasing epochs kept in device memory, so a captured CUDA graph replays correctly.
#include "dp_comm.cuh"
#include "step.cuh"

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// A peer that never arrives (crashed rank, mismatched step sequence) must surface as an error, not as a hung GPU.
static __device__ __noinline__ void dp_timeout(int bucket, int cta, int peer, uint32_t want, uint32_t have) {
  printf("fql_b200: data-parallel flag timeout (bucket %d cta %d waiting for rank %d: epoch %u, have %u)\n", bucket, cta, peer, want, have);
  __trap();
}
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t want, int bucket, int cta, int peer) {
  uint32_t v = ld_acquire_sys(p);
  if ((int32_t)(v - want) >= 0) return;
  const unsigned long long t0 = gtimer();
  while ((int32_t)((v = ld_acquire_sys(p)) - want) < 0) {
    if (gtimer() - t0 > 60ull * 1000000000ull) dp_timeout(bucket, cta, peer, want, v);
  }
}

struct DpArgs {
  int rank, world, bucket, S;
  float* base[FQL_DP_MAX_RANKS];
  float* mc;
  int64_t arena, off, n4;       // bucket = floats [off, off + 4 * n4) of every seed's arena
  int64_t o_flags, o_epochs, o_raw;
  const float* raw_local;       // optional [S][FQL_NUM_RAW]
};

template <bool MC>
__global__ void __launch_bounds__(512) dp_reduce_kernel(const DpArgs a) {
  const int b = blockIdx.x, tid = threadIdx.x;
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(a.base[a.rank] + a.o_flags) + ((int64_t)a.bucket * DP_MAX_CTAS + b) * FQL_DP_MAX_RANKS;
  uint32_t* my_epoch = reinterpret_cast<uint32_t*>(a.base[a.rank] + a.o_epochs) + a.bucket * DP_MAX_CTAS + b;
  __shared__ uint32_t s_epoch;
  if (tid == 0) s_epoch = *my_epoch;
  __syncthreads();
  const uint32_t e1 = s_epoch + 1, e2 = s_epoch + 2;
  // ---- 1. every rank's gradients of this bucket are final (its kernel runs behind them in its stream)
  if (tid < a.world) {
    uint32_t* peer = reinterpret_cast<uint32_t*>(a.base[tid] + a.o_flags) + ((int64_t)a.bucket * DP_MAX_CTAS + b) * FQL_DP_MAX_RANKS + a.rank;
    st_release_sys(peer, e1);
    wait_flag(my_flags + tid, e1, a.bucket, b, tid);
  }
  __syncthreads();
  // ---- 2. my share of every seed's bucket
  const int64_t chunk = (a.n4 + a.world - 1) / a.world;
  const int64_t lo = a.rank * chunk, hi = (lo + chunk < a.n4) ? lo + chunk : a.n4;
  const int64_t per_seed = hi > lo ? hi - lo : 0;
  for (int64_t i = (int64_t)b * blockDim.x + tid; i < per_seed * a.S; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / per_seed, e = a.off + (int64_t)s * a.arena + (lo + (i - s * per_seed)) * 4;
    float4 v;
    if (MC) {
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "l"(a.mc + e)
                   : "memory");
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.mc + e), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                   : "memory");
    } else {
      float4 p[FQL_DP_MAX_RANKS];
#pragma unroll
      for (int r = 0; r < FQL_DP_MAX_RANKS; r++)
        if (r < a.world) p[r] = __ldcg(reinterpret_cast<const float4*>(a.base[r] + e));   // all peers' loads in flight together
      v = p[0];
#pragma unroll
      for (int r = 1; r < FQL_DP_MAX_RANKS; r++)
        if (r < a.world) { v.x += p[r].x; v.y += p[r].y; v.z += p[r].z; v.w += p[r].w; }
#pragma unroll
      for (int r = 0; r < FQL_DP_MAX_RANKS; r++)
        if (r < a.world) __stcg(reinterpret_cast<float4*>(a.base[r] + e), v);
    }
  }
  // the metric accumulators of this rank -> slot [rank] of every rank's gather buffer (fql_finalize_info_seed reduces them)
  if (a.raw_local && b == 0) {
    const int n = a.S * FQL_NUM_RAW;
    for (int i = tid; i < n * a.world; i += blockDim.x) {
      const int r = i / n, k = i - r * n;
      __stcg(a.base[r] + a.o_raw + (int64_t)a.rank * n + k, a.raw_local[k]);
    }
  }
  // ---- 3. everything I wrote is visible system-wide before any peer is told so
  __threadfence_system();
  __syncthreads();
  if (tid < a.world) {
    uint32_t* peer = reinterpret_cast<uint32_t*>(a.base[tid] + a.o_flags) + ((int64_t)a.bucket * DP_MAX_CTAS + b) * FQL_DP_MAX_RANKS + a.rank;
    st_release_sys(peer, e2);
    wait_flag(my_flags + tid, e2, a.bucket, b, tid);
  }
  __syncthreads();
  if (tid == 0) *my_epoch = e2;
}

}  // namespace

DpLamArgs dp_lam_args(const DpState& dp) {
  DpLamArgs l;
  memset(&l, 0, sizeof(l));
  if (!dp.active) return l;
  l.rank = dp.comm.rank;
  l.world = dp.comm.world;
  for (int r = 0; r < dp.comm.world; r++) l.slots[r] = reinterpret_cast<unsigned long long*>(reinterpret_cast<float*>(dp.comm.base[r]) + dp.lay.lam);
  l.epoch = reinterpret_cast<uint32_t*>(reinterpret_cast<float*>(dp.comm.base[dp.comm.rank]) + dp.lay.epochs) + DP_BUCKETS * DP_MAX_CTAS;
  return l;
}

int dp_reduce_bucket(const DpState& dp, int bucket, int64_t off, int64_t n, const float* raw_local, cudaStream_t st) {
  FQL_REQUIRE(dp.active, "dp_reduce_bucket: no communicator attached");
  FQL_REQUIRE(bucket >= 0 && bucket < DP_BUCKETS && off % 4 == 0 && n % 4 == 0, "dp_reduce_bucket: bad bucket %d [%lld, +%lld)", bucket,
              (long long)off, (long long)n);
  DpArgs a;
  memset(&a, 0, sizeof(a));
  a.rank = dp.comm.rank; a.world = dp.comm.world; a.bucket = bucket; a.S = dp.S;
  for (int r = 0; r < dp.comm.world; r++) a.base[r] = reinterpret_cast<float*>(dp.comm.base[r]);
  a.mc = reinterpret_cast<float*>(dp.comm.base_mc);
  a.arena = dp.arena; a.off = off; a.n4 = n / 4;
  a.o_flags = dp.lay.flags; a.o_epochs = dp.lay.epochs; a.o_raw = dp.lay.raw_all;
  a.raw_local = raw_local;
  // one CTA per 8192 float4 of this rank's share (>= 4 float4 in flight per thread), every rank launches the same grid
  const int64_t share = (a.n4 + a.world - 1) / a.world * dp.S;
  int ctas = (int)((share + 8191) / 8192);
  ctas = ctas < 1 ? 1 : (ctas > DP_MAX_CTAS ? DP_MAX_CTAS : ctas);
  if (a.mc) dp_reduce_kernel<true><<<ctas, 512, 0, st>>>(a);
  else dp_reduce_kernel<false><<<ctas, 512, 0, st>>>(a);
  FQL_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t fql_dp_symmetric_bytes(const FqlDims* d, int32_t world) {
  Layout L;
  if (fql_build_layout(d, &L)) return 0;
  if (world < 1 || world > FQL_DP_MAX_RANKS) {
    fql_set_error("fql_dp_symmetric_bytes: world %d outside [1, %d]", world, FQL_DP_MAX_RANKS);
    return 0;
  }
  return (size_t)dp_layout(d->num_seeds, L.arena).total * sizeof(float);
}
