// dp_comm.cu -- the gradient exchange of the data-parallel step as kernels of this library over NVLink peer memory / NVLS
// multicast (no NCCL on the step).  The reference is single-device: this replaces nothing in it, it is what makes
// "larger-batch runs partitioned across the 8 B200s" (north_star) one CUDA graph per step.
//
//   dp_reduce_kernel<MC>: one launch per gradient bucket, as soon as that network's gradients are final on this rank.
//     rank r            1. CTA 0 tells every peer "my gradients are final" (st.release.sys into the peer's flag pad); every CTA
//                          waits for the same word from every peer (ld.acquire.sys on its own pad);
//                       2. reduces ITS 1/world share of the bucket, the CTAs striding over it with several independent 16-byte
//                          reductions in flight per thread (a round trip through the switch is ~3 us: one at a time ran at
//                          57 GB/s):  MC: multimem.ld_reduce.add.v4.f32 through the NVSwitch (one load returns the sum over all
//                          ranks' arenas), multimem.st.v4.f32 of the sum into every arena;  !MC: loads from the `world` peer
//                          mappings summed in rank order, stores to every peer mapping;
//                       3. fence.sys, arrival counter: the last CTA runs the second flag round -- when the kernel ends on a
//                          rank, every rank's share has landed in its arena (the optimizer pass that follows in the stream reads
//                          plain local memory).
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
    __nanosleep(32);
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

// multimem: one load returns the sum over all ranks' arenas (reduced in the NVSwitch), one store lands in every arena
__device__ __forceinline__ float4 mc_ld_reduce(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// DP_UNROLL independent reductions in flight per thread (a round trip through the switch is ~3 us)
template <bool MC, int DP_UNROLL>
__global__ void __launch_bounds__(512) dp_reduce_kernel(const DpArgs a) {
  const int b = blockIdx.x, tid = threadIdx.x;
  // flag words of this bucket: [rank] written by the peers (monotonic epochs); local: the epoch of the last completed exchange and
  // the arrival counter of this launch's CTAs
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(a.base[a.rank] + a.o_flags) + (int64_t)a.bucket * DP_MAX_CTAS * FQL_DP_MAX_RANKS;
  uint32_t* my_epoch = reinterpret_cast<uint32_t*>(a.base[a.rank] + a.o_epochs) + a.bucket * DP_MAX_CTAS;
  uint32_t* my_count = my_epoch + 1;
  __shared__ uint32_t s_epoch;
  __shared__ int s_last;
  if (tid == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(my_epoch);
  __syncthreads();
  const uint32_t e1 = s_epoch + 1, e2 = s_epoch + 2;
  // ---- 1. every rank's gradients of this bucket are final (its kernel runs behind them in its stream): CTA 0 tells the peers,
  //         every CTA waits for all of them
  if (tid < a.world) {
    if (b == 0) {
      uint32_t* peer = reinterpret_cast<uint32_t*>(a.base[tid] + a.o_flags) + (int64_t)a.bucket * DP_MAX_CTAS * FQL_DP_MAX_RANKS + a.rank;
      st_release_sys(peer, e1);
    }
    wait_flag(my_flags + tid, e1, a.bucket, b, tid);
  }
  __syncthreads();
  // ---- 2. my share of every seed's bucket, DP_UNROLL independent 16-byte reductions in flight per thread
  const int64_t chunk = (a.n4 + a.world - 1) / a.world;
  const int64_t lo = a.rank * chunk, hi = (lo + chunk < a.n4) ? lo + chunk : a.n4;
  const int64_t per_seed = hi > lo ? hi - lo : 0;
  const int64_t total = per_seed * a.S, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)b * blockDim.x + tid; i0 < total; i0 += DP_UNROLL * stride) {
    int64_t e[DP_UNROLL];
    float4 v[DP_UNROLL];
#pragma unroll
    for (int u = 0; u < DP_UNROLL; u++) {
      const int64_t i = i0 + u * stride;
      e[u] = -1;
      if (i < total) {
        const int64_t s = i / per_seed;
        e[u] = a.off + s * a.arena + (lo + (i - s * per_seed)) * 4;
      }
    }
    if (MC) {
#pragma unroll
      for (int u = 0; u < DP_UNROLL; u++)
        if (e[u] >= 0) v[u] = mc_ld_reduce(a.mc + e[u]);
#pragma unroll
      for (int u = 0; u < DP_UNROLL; u++)
        if (e[u] >= 0) mc_st(a.mc + e[u], v[u]);
    } else {
#pragma unroll
      for (int u = 0; u < DP_UNROLL; u++) {
        if (e[u] < 0) continue;
        float4 p[FQL_DP_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < FQL_DP_MAX_RANKS; r++)
          if (r < a.world) p[r] = __ldcg(reinterpret_cast<const float4*>(a.base[r] + e[u]));   // all peers' loads in flight together
        float4 t = p[0];
#pragma unroll
        for (int r = 1; r < FQL_DP_MAX_RANKS; r++)
          if (r < a.world) { t.x += p[r].x; t.y += p[r].y; t.z += p[r].z; t.w += p[r].w; }
        v[u] = t;
      }
#pragma unroll
      for (int u = 0; u < DP_UNROLL; u++) {
        if (e[u] < 0) continue;
#pragma unroll
        for (int r = 0; r < FQL_DP_MAX_RANKS; r++)
          if (r < a.world) __stcg(reinterpret_cast<float4*>(a.base[r] + e[u]), v[u]);
      }
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
  // ---- 3. everything this rank wrote is visible system-wide before any peer is told so: the last CTA to arrive tells the peers
  //         and waits for theirs -- when the kernel ends on a rank, every rank's share has landed in its arena
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    const uint32_t t = atomicAdd(my_count, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (tid < a.world) {
    uint32_t* peer = reinterpret_cast<uint32_t*>(a.base[tid] + a.o_flags) + (int64_t)a.bucket * DP_MAX_CTAS * FQL_DP_MAX_RANKS + a.rank;
    st_release_sys(peer, e2);
    wait_flag(my_flags + tid, e2, a.bucket, b, tid);
  }
  __syncthreads();
  if (tid == 0) {
    *my_count = 0;
    *my_epoch = e2;
  }
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
  // one CTA per 1024 float4 of this rank's share (two reductions in flight per thread), at most one CTA per SM pair
  const int64_t share = (a.n4 + a.world - 1) / a.world * dp.S;
  static const int max_ctas = getenv("FQL_DP_CTAS") ? atoi(getenv("FQL_DP_CTAS")) : DP_REDUCE_CTAS;
  static const int unroll = getenv("FQL_DP_UNROLL") ? atoi(getenv("FQL_DP_UNROLL")) : 4;
  static const int threads = getenv("FQL_DP_THREADS") ? atoi(getenv("FQL_DP_THREADS")) : 512;
  int ctas = (int)((share + 1023) / 1024);
  ctas = ctas < 1 ? 1 : (ctas > max_ctas ? max_ctas : ctas);
  if (a.mc) {
    if (unroll >= 8) dp_reduce_kernel<true, 8><<<ctas, threads, 0, st>>>(a);
    else if (unroll >= 4) dp_reduce_kernel<true, 4><<<ctas, threads, 0, st>>>(a);
    else dp_reduce_kernel<true, 1><<<ctas, threads, 0, st>>>(a);
  } else {
    if (unroll >= 4) dp_reduce_kernel<false, 4><<<ctas, threads, 0, st>>>(a);
    else dp_reduce_kernel<false, 1><<<ctas, threads, 0, st>>>(a);
  }
  FQL_CHECK_LAUNCH();
  return 0;
}

// Stand-alone exchange of grads[s][off, off + n) of every seed (callers that drive fql_step_grads / fql_step_apply themselves, and
// profiles/micro/dp_reduce_bench.py).  Enqueue-only; every rank must call it with the same arguments.
int dp_allreduce_range(const DpState& dp, int bucket, long long off, long long n, cudaStream_t st) {
  return dp_reduce_bucket(dp, bucket, off, n, nullptr, st);
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
