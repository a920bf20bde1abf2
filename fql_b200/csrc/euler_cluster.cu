// euler_cluster.cu -- compute_flow_actions (agents/fql.py:155-171) as ONE persistent thread-block-cluster kernel.
//
// The Euler integration is the longest dependent chain of an update: flow_steps x 5 Dense layers on the same rows.  At
// batch 256 that is only two 128-row tiles, so the per-row-tile fused kernel (mlp_tc.cu) leaves 146 SMs idle and a
// kernel-per-layer schedule pays a launch + pipeline-fill per layer.  Here a cluster of 8 CTAs owns one 128-row tile:
//
//   CTA j computes output columns [64j, 64j+64) of every hidden layer with the FULL K: D_j = A[128 x K] * W_l[K x 64j..]
//     A: all 8 K-blocks ([128][64] bf16, SWIZZLE_128B) resident in smem; K-block i is the slice CTA i produced for the
//        previous layer, all-gathered through L2 with TMA multicast: epilogue -> st.global (bf16 row-major scratch) -> ONE
//        cp.async.bulk.tensor ... .multicast::cluster per CTA that lands its slice in all 8 CTAs' smem and completes their
//        per-block mbarriers directly (MMAs start on the first block that lands).  "smem may be overwritten" is a multicast
//        tcgen05.commit from every CTA's MMA warp.  (Measured alternatives, profiles/micro: DSMEM bulk copies 15 GB/s/SM,
//        unicast gather + remote arrives 7.6 us/layer.)
//     B: the CTA's [K][64] weight slice (MN-major, straight from the Flax [in,out] bf16 shadow), prefetched by TMA while the
//        previous layer's epilogue and exchange run (weights do not depend on activations)
//     D: 128 x 64 fp32 in TMEM, epilogue one thread per row: + bias, GELU(tanh), bf16
//   the narrow last Dense (N = action_dim <= 32, zero-padded to 64) is computed redundantly by every CTA, so each CTA applies
//   a += v / flow_steps to ITS copy of the first-layer operand tile (resident in smem for all steps) with no exchange.
//
// No kernel boundary, no global barrier: 4 L2 exchanges + 5 MMAs chains per Euler step.
#include "step.cuh"
#include "tc_prims.cuh"

#include <cudaTypedefs.h>

using namespace tc;

namespace {

constexpr int NC = 8;                 // CTAs per cluster = column slices
constexpr int TILE_M = 128;
constexpr int KB = 64;                // K-block
constexpr int A_BLK = TILE_M * 128;   // 16 KB: [128][64] bf16
constexpr int B_BLK = KB * 128;       //  8 KB: [64 k][64 n] bf16
constexpr int MAX_A = 32;
constexpr int NMMA = 4;                // MMA-issuer warps (k-step w of every K block -> warp w, own TMEM accumulator)
constexpr int NEPI = 4;                // epilogue warps (4: one thread per row, both 32-column halves; 8 was measured slower:
                                       // the 416-thread CTA caps registers at 128 and the epilogue spills)
constexpr int NTHREADS = 32 * (1 + NMMA + NEPI);
constexpr int EPI0 = 32 * (1 + NMMA);  // first epilogue thread

struct EulerArgs {
  int H, K0, K0pad, A, F, n_steps, NL;
  int S, M, tiles;
  int x_rows_s;                 // rows of one seed in the X0b tensor
  int x_row0;                   // first row (inside a seed) of the Euler rows
  int w_row[FQL_MAXL];          // first shadow row (units of H elements) of layer l < NL-1
  int w_rows_s;
  int wl_row, wl_rows_s;        // padded last layer (units of 64 elements)
  const float* params;
  long long arena;
  long long off_b[FQL_MAXL];
  const float* a0;              // [S][M][A] initial actions (noise)
  float* target;                // [S][M][A] clip(final)
  __nv_bfloat16* hx;            // exchange scratch [2][S*tiles*128][H]
  long long hx_buf_elems;
  unsigned long long* dbg;  // optional [CTA][16] globaltimer stamps of iteration DBG_IT (diagnostics)
};

constexpr int DBG_IT = 7;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ float gelu_fast(float x) {
  const float u = FQL_GELU_C * (x + FQL_GELU_A * x * x * x);
  return 0.5f * x * (1.0f + tanh_approx(u));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("fql_b200: cluster mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) euler_cluster_kernel(const __grid_constant__ CUtensorMap mapX,
                                                               const __grid_constant__ CUtensorMap mapW,
                                                               const __grid_constant__ CUtensorMap mapWL,
                                                               const __grid_constant__ CUtensorMap mapHx, const EulerArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nkb = a.H / KB, nkb_x = a.K0pad / KB;
  uint8_t* sA = smem;                        // [nkb][16 KB]
  uint8_t* sB = sA + nkb * A_BLK;            // [nkb][8 KB]  weight slice of the current layer
  uint8_t* sX = sB + nkb * B_BLK;            // [nkb_x][16 KB] first-layer operand tile, resident for all steps
  float* sBias = reinterpret_cast<float*>(sX + nkb_x * A_BLK);  // [2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 2 * 64);
  uint64_t* full_a = bars;          // [8]
  uint64_t* full_b = bars + 8;
  uint64_t* acc_full = bars + 9;
  uint64_t* free_a = bars + 10;     // count NC: every CTA's MMAs of the current layer are done -> sA may be overwritten
  uint64_t* a_ready = bars + 11;    // count 4: the Euler update of the resident X tile is done (local)
  uint64_t* x_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t j = cluster_ctarank();                  // column slice
  const int cl = blockIdx.x / NC;                        // cluster id = (seed, tile)
  const int tile = cl % a.tiles, s = cl / a.tiles;
  const int NL = a.NL;
  const int total = a.n_steps * NL;
  const int hx_row = (s * a.tiles + tile) * TILE_M;      // this tile's rows in the exchange scratch
  unsigned long long* dbg = a.dbg ? a.dbg + blockIdx.x * 16 : nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapWL);
    tma_prefetch_desc(&mapHx);
    for (int i = 0; i < 8; i++) mbar_init(&full_a[i], 1);
    mbar_init(full_b, 1);
    mbar_init(acc_full, NMMA);
    mbar_init(free_a, NC * NMMA);
    mbar_init(a_ready, NEPI);
    mbar_init(x_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 64 * NMMA);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();  // every CTA's barriers are initialised before any remote arrive
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_expect_tx(x_full, nkb_x * A_BLK);
      const int xrow = s * a.x_rows_s + a.x_row0 + tile * TILE_M;
      for (int kb = 0; kb < nkb_x; kb++) tma_load_2d(sX + kb * A_BLK, &mapX, x_full, kb * KB, xrow);
      int n_arm = 0;
      for (int it = 0; it < total; it++) {
        const int l = it % NL;
        const bool last = (l == NL - 1);
        const int K = (l == 0) ? a.K0 : a.H;
        const int kblocks = (K + KB - 1) / KB;
        // weights of this layer: sB is free once the previous layer's MMAs have completed
        if (it > 0) mbar_wait(acc_full, (it - 1) & 1);
        mbar_expect_tx(full_b, kblocks * B_BLK);
        for (int kb = 0; kb < kblocks; kb++) {
          if (!last) tma_load_2d(sB + kb * B_BLK, &mapW, full_b, (int)j * 64, a.w_row[l] + s * a.w_rows_s + kb * KB);
          else tma_load_2d(sB + kb * B_BLK, &mapWL, full_b, 0, a.wl_row + s * a.wl_rows_s + kb * KB);
        }
        if (l >= 1) {
          // arm the per-block barriers; the data arrives as TMA multicasts issued by the 8 CTAs' epilogues
          for (int kb = 0; kb < nkb; kb++) mbar_expect_tx(&full_a[kb], A_BLK);
          if (dbg && it == DBG_IT) {  // diagnostics: true arrival time of every block (this thread is otherwise idle here)
            uint32_t seen = 0;
            while (seen != 0xFF)
              for (int kb = 0; kb < 8; kb++)
                if (!(seen >> kb & 1) && mbar_try_wait(&full_a[kb], n_arm & 1)) { dbg[8 + kb] = gtime(); seen |= 1u << kb; }
          }
          n_arm++;
        }
      }
    }
  } else if (warp <= NMMA) {
    // ================= MMA issuers =================
    // One tcgen05.mma costs its issuing warp ~80 ns for any N (profiles/micro/mma_bench.cu) but the cost is per warp, so four
    // warps each issue k-step w of every 64-wide K block into their own 64 TMEM columns (summed by the epilogue).
    if (lane == 0) {
      const int mw = warp - 1;
      const uint32_t idesc = make_idesc_bf16(TILE_M, 64, false, true);
      const uint64_t a_t0 = make_smem_desc(0, 16, 1024), b_t0 = make_smem_desc(0, B_BLK, 1024);
      const uint64_t a_t = a_t0 + (uint64_t)(mw * 2), b_t = b_t0 + (uint64_t)(mw * 128);
      const uint32_t sa0 = smem_u32(sA) >> 4, sx0 = smem_u32(sX) >> 4, sb0 = smem_u32(sB) >> 4;
      const uint32_t tacc = tmem_base + mw * 64;
      int n_a = 0;  // uses of the full_a barriers
      int n_step = 0;
      for (int it = 0; it < total; it++) {
        const int l = it % NL;
        const int K = (l == 0) ? a.K0 : a.H;
        const int kblocks = (K + KB - 1) / KB;
        if (l == 0) {
          if (it == 0) mbar_wait(x_full, 0);
          else { mbar_wait(a_ready, (n_step - 1) & 1); }
          n_step++;
        }
        mbar_wait(full_b, it & 1);
        tc_fence_after();
        const uint32_t a0d = (l == 0 ? sx0 : sa0);
        if (l >= 1) {
          // hidden / last layers (K = H = 8 blocks): warp w takes K-blocks 2w and 2w+1 entirely -> two barrier waits per layer
          for (int i = 0; i < 2; i++) {
            const int kb = mw * 2 + i;
            mbar_wait(&full_a[kb], n_a & 1);
            tc_fence_after();
            if (dbg && mw == 0 && it == DBG_IT && i == 0) dbg[5] = gtime();
            if (dbg && mw == NMMA - 1 && it == DBG_IT && i == 1) dbg[6] = gtime();
            umma_bf16_x4(tacc, a_t0 + (uint64_t)(a0d + kb * (A_BLK >> 4)), b_t0 + (uint64_t)(sb0 + kb * (B_BLK >> 4)), 2, 128, idesc, i != 0);
          }
        } else {
          // first layer (K0 <= 128): k-step w of every block -> warp w
          for (int kb = 0; kb < kblocks; kb++)
            if (mw < (a.K0 - kb * KB + 15) / 16)
              umma_bf16(tacc, a_t + (uint64_t)(a0d + kb * (A_BLK >> 4)), b_t + (uint64_t)(sb0 + kb * (B_BLK >> 4)), idesc, kb != 0);
        }
        if (l >= 1) n_a++;
        umma_commit(acc_full);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(free_a)),
                     "h"((uint16_t)0xFF)
                     : "memory");
      }
    }
  } else {
    // ================= epilogue: one thread per (row, 32-column half) =================
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int half0 = (warp - (1 + NMMA)) >> 2;      // first 32-column half this warp handles (stride NEPI/4)
    const int row = q * 32 + lane;
    const int grow = tile * TILE_M + row;
    const bool valid = grow < a.M;
    const uint32_t t_lane0 = tmem_base + ((uint32_t)(q * 32) << 16);
    const int et = threadIdx.x - EPI0;
    float act[MAX_A];
#pragma unroll
    for (int c = 0; c < MAX_A; c++) act[c] = 0.f;
    if (valid && half0 == 0) {
#pragma unroll
      for (int c = 0; c < MAX_A; c++)
        if (c < a.A) act[c] = a.a0[((int64_t)s * a.M + grow) * a.A + c];
    }
    int n_exch = 0;
    for (int it = 0; it < total; it++) {
      const int l = it % NL, step = it / NL;
      const bool last = (l == NL - 1);
      float* sb = sBias + (it & 1) * 64;
      {
        const int N = last ? a.A : a.H;
        const int col = last ? et : (int)j * 64 + et;
        if (et < 64) sb[et] = (col < N) ? a.params[(int64_t)s * a.arena + a.off_b[l] + col] : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NEPI) : "memory");
      }
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      if (dbg && et == 0 && it == DBG_IT - 1) dbg[0] = gtime();
      if (dbg && et == 0 && it == DBG_IT) dbg[7] = gtime();
      // partial accumulators of the NMMA issuer warps, loads issued in pairs (a first layer with K0 < 64 has ceil(K0/16) of them)
      const int live = (l == 0 && a.K0 < KB) ? (a.K0 + 15) / 16 : NMMA;
      const int buf = n_exch & 1;
      if (!last) n_exch++;
#pragma unroll 1
      for (int half = half0; half < 2; half += NEPI / 4) {
      const uint32_t t_lane = t_lane0 + half * 32;
      uint32_t r0[32];
      if (!last || half == 0) {
        uint32_t t1[32];
        tmem_ld32(t_lane, r0);
        if (live > 1) tmem_ld32(t_lane + 64, t1);
        tmem_wait_ld();
        if (live > 1) {
#pragma unroll
          for (int i = 0; i < 32; i++) r0[i] = __float_as_uint(__uint_as_float(r0[i]) + __uint_as_float(t1[i]));
        }
        if (live > 2) {
          uint32_t t2[32];
          tmem_ld32(t_lane + 128, t1);
          if (live > 3) tmem_ld32(t_lane + 192, t2);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; i++) {
            float v = __uint_as_float(r0[i]) + __uint_as_float(t1[i]);
            if (live > 3) v += __uint_as_float(t2[i]);
            r0[i] = __float_as_uint(v);
          }
        }
      }
      if (!last) {
        uint4* dst = reinterpret_cast<uint4*>(a.hx + (int64_t)buf * a.hx_buf_elems + (int64_t)(hx_row + row) * a.H + j * 64 + half * 32);
#pragma unroll
        for (int c = 0; c < 4; c++) {
          float h[8];
#pragma unroll
          for (int i = 0; i < 8; i++) h[i] = gelu_fast(__uint_as_float(r0[c * 8 + i]) + sb[half * 32 + c * 8 + i]);
          dst[c] = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
        }
      } else if (half == 0) {
        // Euler step on this CTA's resident copy of the first-layer operand (every CTA computes the same last layer)
        const float inv = (float)a.n_steps;
#pragma unroll
        for (int c = 0; c < MAX_A; c++)
          if (c < a.A) {
            act[c] += (__uint_as_float(r0[c]) + sb[c]) / inv;
            const int col = a.F + c;
            *reinterpret_cast<__nv_bfloat16*>(sX + (col >> 6) * A_BLK + sw128_off(row, (col & 63) >> 3) + (col & 7) * 2) =
                __float2bfloat16(act[c]);
          }
        {
          const int col = a.F + a.A;
          *reinterpret_cast<__nv_bfloat16*>(sX + (col >> 6) * A_BLK + sw128_off(row, (col & 63) >> 3) + (col & 7) * 2) =
              __float2bfloat16((float)((double)(step + 1) / (double)a.n_steps));
        }
        if (step == a.n_steps - 1 && valid && j == 0) {
#pragma unroll
          for (int c = 0; c < MAX_A; c++)
            if (c < a.A) a.target[((int64_t)s * a.M + grow) * a.A + c] = fminf(fmaxf(act[c], -1.0f), 1.0f);
        }
      }
      }  // halves
      if (!last) {
        if (dbg && et == 0 && it == DBG_IT - 1) dbg[1] = gtime();
        // publish: my slice is in the scratch -> (CTA barrier) -> one thread multicasts it into all 8 CTAs' sA[j]
        tc_fence_before();
        asm volatile("bar.sync 2, %0;" ::"n"(32 * NEPI) : "memory");
        if (dbg && et == 0 && it == DBG_IT - 1) dbg[2] = gtime();
        if (et == 0) {
          asm volatile("fence.proxy.async.global;" ::: "memory");  // this CTA's st.global (ordered by the bar.sync) -> its own TMA read
          mbar_wait_cluster(free_a, it & 1);  // every CTA finished reading sA for this layer
          if (dbg && it == DBG_IT - 1) dbg[3] = gtime();
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
              ::"r"(smem_u32(sA + j * A_BLK)), "l"(reinterpret_cast<uint64_t>(&mapHx)), "r"(smem_u32(&full_a[j])), "r"((int)j * KB),
              "r"(buf * (a.S * a.tiles * TILE_M) + hx_row), "h"((uint16_t)0xFF)
              : "memory");
        }
      } else {
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA may exit while a peer can still arrive on its barriers
  if (warp == 1) tmem_dealloc(tmem_base, 64 * NMMA);
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_map_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows) {
  auto enc = get_encode();
  FQL_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FQL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) inner=%llu rows=%llu", (int)r, (unsigned long long)inner,
              (unsigned long long)rows);
  return 0;
}

}  // namespace

size_t tc_euler_scratch_elems(const FqlDims* d, int M) {
  const int tiles = (M + TILE_M - 1) / TILE_M;
  return (size_t)2 * d->num_seeds * tiles * TILE_M * d->hidden;
}

int tc_euler_cluster(const TcEulerSpec& f, cudaStream_t st) {
  const FqlDims* d = f.d;
  const Layout& L = *f.L;
  FQL_TRY(tc_supported(d));
  FQL_REQUIRE(d->hidden == 512, "euler_cluster_kernel is built for hidden = 512 (8 column slices of 64)");
  const NetView& nv = L.net[FQL_NET_ACTOR_BC_FLOW];
  EulerArgs a;
  memset(&a, 0, sizeof(a));
  a.H = d->hidden; a.K0 = nv.in_dim; a.K0pad = (int)round_up64(nv.in_dim, 64); a.A = d->action_dim; a.F = d->obs_dim;
  a.n_steps = d->flow_steps; a.NL = nv.n_layers;
  a.S = d->num_seeds; a.M = f.M; a.tiles = (f.M + TILE_M - 1) / TILE_M;
  a.x_rows_s = f.Mcap0; a.x_row0 = f.r0_in;
  const int64_t seed_elems = tc_shadow_seed_elems(d, L);
  a.w_rows_s = (int)(seed_elems / d->hidden);
  a.wl_rows_s = (int)(seed_elems / 64);
  for (int l = 0; l < nv.n_layers; l++) {
    a.w_row[l] = (int)(nv.off_w[l] / d->hidden);
    a.off_b[l] = nv.off_b[l];
  }
  int64_t wl = L.arena;
  for (int t = 0; t < FQL_NET_ACTOR_BC_FLOW; t++) wl += (int64_t)L.net[t].ens * d->hidden * 64;
  a.wl_row = (int)(wl / 64);
  a.params = f.params; a.arena = L.arena; a.a0 = f.a0; a.target = f.target;
  a.hx = reinterpret_cast<__nv_bfloat16*>(f.scratch);
  a.hx_buf_elems = (long long)a.S * a.tiles * TILE_M * a.H;
  a.dbg = reinterpret_cast<unsigned long long*>(f.dbg);
  FQL_REQUIRE(f.scratch != nullptr && f.a0 && f.target, "tc_euler_cluster: NULL argument");
  const int smem = (a.H / KB) * (A_BLK + B_BLK) + (a.K0pad / KB) * A_BLK + 2 * 64 * 4 + 256 + 1024;
  FQL_REQUIRE(smem <= 232448, "euler_cluster_kernel: shared memory %d > 227 KB", smem);
  CUtensorMap mapX, mapW, mapWL, mapHx;
  FQL_TRY(make_map_2d(&mapX, f.X0b, a.K0pad, (uint64_t)a.S * f.Mcap0, 64, TILE_M));
  FQL_TRY(make_map_2d(&mapW, f.shadow, d->hidden, (uint64_t)a.S * a.w_rows_s, 64, KB));
  FQL_TRY(make_map_2d(&mapWL, f.shadow, 64, (uint64_t)a.S * a.wl_rows_s, 64, KB));
  FQL_TRY(make_map_2d(&mapHx, f.scratch, d->hidden, (uint64_t)2 * a.S * a.tiles * TILE_M, 64, TILE_M));
  static bool attr_set = false;
  if (!attr_set) {
    FQL_CHECK_CUDA(cudaFuncSetAttribute(euler_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(NC * a.tiles * a.S);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;  // (claiming all 227 KB to keep other kernels off these SMs was measured: no effect)
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FQL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, euler_cluster_kernel, mapX, mapW, mapWL, mapHx, a));
  FQL_CHECK_LAUNCH();
  return 0;
}
