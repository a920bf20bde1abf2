// optim.cu -- fused multi-tensor optimizer pass over the flat arenas (north_star #5):
//   gradient statistics   utils/flax_utils.py:139-149  (per-leaf max / min / L2 -> max, min, L1-of-L2)
//   optax.adam apply      utils/flax_utils.py:120-130, agents/fql.py:237 (b1 .9, b2 .999, eps 1e-8, eps_root 0)
//   Polyak target update  agents/fql.py:113-120 -- from the PRE-step critic and PRE-step target
// One pass: each CTA owns one FQL_LEAF_PAD-float block (never straddles a leaf), reads g,p,m,v once, writes p,m,v
// (+ target block for critic leaves).  Target-critic leaves have exactly-zero gradients so Adam is the
// identity on them; they are only written by the Polyak half.  HBM-bound: 28 B/param (+8 B/critic param).
#include "step.cuh"

#include <cuda_bf16.h>

namespace {

__global__ void __launch_bounds__(256) adam_polyak_stats_kernel(Layout L, FqlHparams hp, float* __restrict__ params,
                                                                float* __restrict__ mu, float* __restrict__ nu,
                                                                const float* __restrict__ grads,
                                                                const int32_t* __restrict__ count,
                                                                float* __restrict__ partials, __nv_bfloat16* __restrict__ shadow,
                                                                int64_t shadow_seed, int blk0, int nblk) {
  const int blk = blk0 + blockIdx.x, s = blockIdx.y;
  const int64_t off = (int64_t)blk * FQL_LEAF_PAD + threadIdx.x * 4;
  const int64_t base = (int64_t)s * L.arena;
  float* part = partials + ((int64_t)s * nblk + blk) * 4;
  const NetView& tgt = L.net[FQL_NET_TARGET_CRITIC];
  const NetView& cri = L.net[FQL_NET_CRITIC];
  if (off >= tgt.begin && off < tgt.end) {  // block-uniform
    if (threadIdx.x == 0) { part[0] = 0.f; part[1] = 0.f; part[2] = 0.f; }
    return;
  }
  // optax bias_correction: 1 - decay**count in float32.  A correctly rounded float pow (via double) once per CTA: one ulp
  // of pow(0.999f, t) is 7.5e-6 of (1 - 0.999^8), so a sloppy powf would show up in the parameters.
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const double t = (double)(count[0] + 1);
    s_bc[0] = 1.0f - (float)pow((double)hp.beta1, t);
    s_bc[1] = 1.0f - (float)pow((double)hp.beta2, t);
  }
  __syncthreads();
  const float bc1 = s_bc[0], bc2 = s_bc[1];
  float4 g = *reinterpret_cast<const float4*>(grads + base + off);
  float4 p = *reinterpret_cast<const float4*>(params + base + off);
  float4 m = *reinterpret_cast<const float4*>(mu + base + off);
  float4 v = *reinterpret_cast<const float4*>(nu + base + off);
  float gr[4] = {g.x, g.y, g.z, g.w}, pr[4] = {p.x, p.y, p.z, p.w}, mr[4] = {m.x, m.y, m.z, m.w}, vr[4] = {v.x, v.y, v.z, v.w};
  float pn[4];
  float mx = -INFINITY, mn = INFINITY, sq = 0.f;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float gi = gr[i];
    mx = fmaxf(mx, gi);
    mn = fminf(mn, gi);
    sq += gi * gi;
    mr[i] = hp.beta1 * mr[i] + hp.one_minus_beta1 * gi;
    vr[i] = hp.beta2 * vr[i] + hp.one_minus_beta2 * gi * gi;
    const float mhat = mr[i] / bc1;
    const float vhat = vr[i] / bc2;
    pn[i] = pr[i] + (-hp.lr * (mhat / (sqrtf(vhat) + hp.eps)));
  }
  *reinterpret_cast<float4*>(params + base + off) = make_float4(pn[0], pn[1], pn[2], pn[3]);
  if (shadow) {  // bf16 tensor-core operand copy of the fresh parameters (same [in,out] layout)
    __nv_bfloat162 lo = __floats2bfloat162_rn(pn[0], pn[1]), hi = __floats2bfloat162_rn(pn[2], pn[3]);
    *reinterpret_cast<uint2*>(shadow + (int64_t)s * shadow_seed + off) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
  *reinterpret_cast<float4*>(mu + base + off) = make_float4(mr[0], mr[1], mr[2], mr[3]);
  *reinterpret_cast<float4*>(nu + base + off) = make_float4(vr[0], vr[1], vr[2], vr[3]);
  if (off >= cri.begin && off < cri.end) {  // Polyak with the pre-step critic values still in registers
    const int64_t toff = base + off - cri.begin + tgt.begin;
    float4 tp = *reinterpret_cast<const float4*>(params + toff);
    const float omt = hp.one_minus_tau;
    tp.x = pr[0] * hp.tau + tp.x * omt;
    tp.y = pr[1] * hp.tau + tp.y * omt;
    tp.z = pr[2] * hp.tau + tp.z * omt;
    tp.w = pr[3] * hp.tau + tp.w * omt;
    *reinterpret_cast<float4*>(params + toff) = tp;
    if (shadow) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(tp.x, tp.y), hi = __floats2bfloat162_rn(tp.z, tp.w);
      *reinterpret_cast<uint2*>(shadow + (int64_t)s * shadow_seed + (toff - base)) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
  // block reduce of the statistics
  __shared__ float smx[8], smn[8], ssq[8];
  mx = warp_max(mx);
  mn = warp_min(mn);
  sq = warp_sum(sq);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { smx[w] = mx; smn[w] = mn; ssq[w] = sq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = smx[0], b = smn[0], c = ssq[0];
#pragma unroll
    for (int i = 1; i < 8; i++) { a = fmaxf(a, smx[i]); b = fminf(b, smn[i]); c += ssq[i]; }
    part[0] = a; part[1] = b; part[2] = c;
  }
}

// one CTA per seed, one WARP per leaf (lanes stride over the leaf's block partials): per-leaf L2 norms, then the three
// scalars of flax_utils.py:147-149
__global__ void __launch_bounds__(1024) grad_stats_final_kernel(Layout L, const float* __restrict__ partials,
                                                                float* __restrict__ gstats, int32_t* count_inc) {
  const int s = blockIdx.x;
  const int nblk = L.leaf_blk[L.n_leaves];
  const float* part = partials + (int64_t)s * nblk * 4;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __shared__ float smx[32], smn[32], snorm[32];
  float wmx = 0.f, wmn = 0.f, wnorm = 0.f;  // the target critic's zero gradients are leaves too
  for (int leaf = w; leaf < L.n_leaves; leaf += 32) {
    if (L.leaf_net[leaf] == FQL_NET_TARGET_CRITIC) continue;
    float mx = -INFINITY, mn = INFINITY, sq = 0.f;
    for (int b = L.leaf_blk[leaf] + lane; b < L.leaf_blk[leaf + 1]; b += 32) {
      mx = fmaxf(mx, part[b * 4 + 0]);
      mn = fminf(mn, part[b * 4 + 1]);
      sq += part[b * 4 + 2];
    }
    mx = warp_max(mx);
    mn = warp_min(mn);
    sq = warp_sum(sq);
    wmx = fmaxf(wmx, mx);
    wmn = fminf(wmn, mn);
    wnorm += sqrtf(sq);
  }
  if (lane == 0) { smx[w] = wmx; smn[w] = wmn; snorm[w] = wnorm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = smx[0], b = smn[0], c = snorm[0];
    for (int i = 1; i < 32; i++) { a = fmaxf(a, smx[i]); b = fminf(b, smn[i]); c += snorm[i]; }
    gstats[s * 4 + 0] = a;
    gstats[s * 4 + 1] = b;
    gstats[s * 4 + 2] = c;
    if (s == 0 && count_inc) count_inc[0] += 1;  // optax count / TrainState.step advance (flax_utils.py:126)
  }
}

__global__ void zero_kernel(float4* p, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace

int launch_adam_polyak_stats(const Layout& L, const FqlHparams& hp, int S, float* params, float* mu, float* nu,
                             const float* grads, const int32_t* count, float* partials, void* shadow, int64_t shadow_seed, cudaStream_t st,
                             int blk0, int blk1) {
  const int nblk = L.leaf_blk[L.n_leaves];
  if (blk1 < 0) blk1 = nblk;
  if (blk1 <= blk0) return 0;
  dim3 grid(blk1 - blk0, S);
  adam_polyak_stats_kernel<<<grid, 256, 0, st>>>(L, hp, params, mu, nu, grads, count, partials, reinterpret_cast<__nv_bfloat16*>(shadow), shadow_seed, blk0, nblk);
  FQL_CHECK_LAUNCH();
  return 0;
}

int launch_grad_stats_final(const Layout& L, int S, const float* partials, float* gstats, int32_t* count_inc, cudaStream_t st) {
  grad_stats_final_kernel<<<S, 1024, 0, st>>>(L, partials, gstats, count_inc);
  FQL_CHECK_LAUNCH();
  return 0;
}

int launch_zero(float* p, int64_t n, cudaStream_t st) {
  if (n == 0) return 0;
  const int64_t n4 = n / 4;  // arenas and workspace slices are multiples of 4 floats
  zero_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4*>(p), n4);
  FQL_CHECK_LAUNCH();
  return 0;
}
