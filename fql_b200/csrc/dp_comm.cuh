// dp_comm.cuh -- data-parallel exchange over NVLink peer memory: layout of the symmetric buffer and the launchers.
#pragma once
#include "common.cuh"

constexpr int DP_MAX_CTAS = 64;     // flag / epoch words reserved per bucket in the symmetric buffer
constexpr int DP_REDUCE_CTAS = 96;  // CTAs of one bucket-reduce launch
constexpr int DP_BUCKETS = 4;       // flag sets: bc-flow | critic | one-step actor (+ metric gather) | whole arena (fp32 schedule)

struct DpLayout {                   // float offsets inside one rank's symmetric buffer
  int64_t grads, raw_all, lam, flags, epochs, total;
};
inline DpLayout dp_layout(int S, int64_t arena) {
  DpLayout l;
  l.grads = 0;
  l.raw_all = round_up64((int64_t)S * arena, 1024);
  l.lam = l.raw_all + round_up64((int64_t)FQL_DP_MAX_RANKS * S * FQL_NUM_RAW, 64);
  l.flags = l.lam + round_up64((int64_t)2 * FQL_DP_MAX_RANKS * S, 64);                 // 64-bit {epoch, value} words
  l.epochs = l.flags + (int64_t)DP_BUCKETS * DP_MAX_CTAS * FQL_DP_MAX_RANKS;          // uint32 [bucket][cta][rank]
  l.total = l.epochs + round_up64((int64_t)DP_BUCKETS * DP_MAX_CTAS + S, 64);         // local: uint32 [bucket][cta], then [S] (lam)
  return l;
}

struct DpState {
  int active = 0;
  FqlDpComm comm;
  DpLayout lay;
  int S = 0;
  int64_t arena = 0;
};

// global mean|q| of config['normalize_q_loss'] (agents/fql.py:74-76) inside critic_post_kernel: one {epoch, value} word per rank
struct DpLamArgs {
  int rank, world;                          // world <= 1: inactive
  unsigned long long* slots[FQL_DP_MAX_RANKS];  // rank p's [FQL_DP_MAX_RANKS][S] words
  uint32_t* epoch;                          // local [S]
};
DpLamArgs dp_lam_args(const DpState& dp);

// Reduce grads[s][off, off + n) of every seed over all ranks, in place on every rank.  raw_local != NULL: the same launch also
// gathers the [S][FQL_NUM_RAW] metric accumulators into every rank's raw_all[rank].
int dp_reduce_bucket(const DpState& dp, int bucket, int64_t off, int64_t n, const float* raw_local, cudaStream_t st);
inline const float* dp_raw_all(const DpState& dp) { return reinterpret_cast<const float*>(dp.comm.base[dp.comm.rank]) + dp.lay.raw_all; }
int dp_allreduce_range(const DpState& dp, int bucket, long long off, long long n, cudaStream_t st);
