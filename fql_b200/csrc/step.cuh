// step.cuh -- shapes, workspace map and launcher prototypes of one FQLAgent.update (agents/fql.py:122-133).
#pragma once
#include "common.cuh"

enum {  // raw accumulators per seed (all-reduced across data-parallel ranks: [0..8] SUM, [9..10] MAX)
  RAW_CRITIC_SQ = 0, RAW_Q_SUM = 1, RAW_BC_SQ = 2, RAW_DISTILL_SQ = 3, RAW_QPI_SUM = 4, RAW_QPI_ABS = 5, RAW_MSE = 6,
  RAW_Q_MAX = 9, RAW_Q_NEGMIN = 10,
};

struct StepShape {
  int S, B, GB, F, A, H, NH;  // seeds, local batch, global batch, obs, action, hidden width, hidden layers
  int q_agg_min, normalize_q_loss, flow_steps;
};

struct PassBuf {               // activations of one grouped MLP pass: [G][Mcap][H] per hidden layer
  float* Z[FQL_MAXL];          // pre-activation (kept for backward / LN)
  float* Hh[FQL_MAXL];         // post activation (+LN): next layer's input
  float* mu[FQL_MAXL];         // LN row statistics [G][Mcap]
  float* rstd[FQL_MAXL];
  float* out;                  // [G][Mcap][out_dim]
  int G, Mcap;
};

struct WsPtrs {
  float *XO, *XF, *XC, *vel;   // first-layer inputs [S][3B][F+A], [S][2B][F+A+1], [3][S][B][F+A]; vel [S][B][A]
  float *O_out, *F_out, *C_out;
  float *dq, *dqs, *dpred, *dapi, *target, *dX0;
  float *raw_local;            // [S][FQL_NUM_RAW] when the caller passes none
  float *gstats;               // [S][4]: grad max, min, L1-of-L2 norm
  float *partials;             // [S][blocks][4] per 1024-float arena block: max, min, sumsq
  PassBuf pO, pF, pC;
  float *dC[2], *dCp[2], *dF[2], *dO[2];  // backward ping-pong [S][E][B][H]
};

StepShape make_shape(const FqlDims* d);
size_t carve_workspace(const FqlDims* d, const Layout& L, void* base, WsPtrs* w);  // returns bytes needed

int launch_prep(const StepShape& sh, const FqlBatch& b, const WsPtrs& w, cudaStream_t st);
int launch_post_onestep(const StepShape& sh, const FqlBatch& b, const WsPtrs& w, float* raw, cudaStream_t st);
int launch_critic_post(const StepShape& sh, const FqlHparams& hp, const FqlBatch& b, const WsPtrs& w, float* raw, cudaStream_t st);
int launch_bc_post(const StepShape& sh, const WsPtrs& w, float* raw, cudaStream_t st);
int launch_euler_update(const StepShape& sh, const WsPtrs& w, int step, cudaStream_t st);
int launch_actor_grad(const StepShape& sh, const FqlHparams& hp, const WsPtrs& w, float* raw, cudaStream_t st);
int launch_finalize_info(const StepShape& sh, const FqlHparams& hp, const float* raw, const float* gstats, float* info,
                         int with_grad_stats, cudaStream_t st);
int launch_clip(const float* in, float* out, int64_t n, cudaStream_t st);
int launch_concat(const float* x0, int k0, const float* x1, int k1, float c2, int k2, float* out, int64_t rows, cudaStream_t st);
int launch_euler_inplace(float* X, const float* v, int F, int A, int64_t rows, int step, int nsteps, float* out, cudaStream_t st);

// optim.cu
int launch_adam_polyak_stats(const Layout& L, const FqlHparams& hp, int S, float* params, float* mu, float* nu,
                             const float* grads, const int32_t* count, float* partials, void* shadow, cudaStream_t st);
int launch_grad_stats_final(const Layout& L, int S, const float* partials, float* gstats, int32_t* count_inc, cudaStream_t st);
int launch_zero(float* p, int64_t n, cudaStream_t st);
