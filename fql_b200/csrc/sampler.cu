// sampler.cu -- device side of Dataset.sample (utils/datasets.py:68-112): fancy-index gather, frame stacking and the
// edge-padded random crop, on a dataset that lives in HBM.  Indices (and crop offsets) are drawn on the host from the
// global numpy MT19937 exactly as the reference does, so results are bit-exact; this is pure byte movement (HBM-bound).
#include "common.cuh"

namespace {

// dst[i,:] = src[idx[i],:]   (utils/datasets.py:96  arr[idxs])
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, const int64_t* __restrict__ idxs,
                                   int64_t n_idx, int64_t row_elems) {
  for (int64_t i = blockIdx.x; i < n_idx; i += gridDim.x) {
    const T* s = src + idxs[i] * row_elems;
    T* d = dst + i * row_elems;
    for (int64_t c = threadIdx.x; c < row_elems; c += blockDim.x) d[c] = s[c];
  }
}

// Frame stack (datasets.py:73-87) + batched_random_crop (datasets.py:17-33, 102-112) in one pass over the output.
// obs block j (j=0 oldest)        = O[max(idx-(fs-1-j), init)]
// next_obs block j<fs-1           = O[max(idx-(fs-2-j), init)] ; block fs-1 = NO[idx]
// crop: out[y,x] = img[clamp(y+cy-pad), clamp(x+cx-pad)]  (edge pad + dynamic_slice at (cy,cx,0))
__global__ void gather_frames_kernel(const uint8_t* __restrict__ O, const uint8_t* __restrict__ NO, uint8_t* __restrict__ obs_out,
                                     uint8_t* __restrict__ next_out, const int64_t* __restrict__ idxs,
                                     const int64_t* __restrict__ init_idxs, const int64_t* __restrict__ crop, int H, int W, int C,
                                     int fs, int pad) {
  const int64_t b = blockIdx.x / H;
  const int y = (int)(blockIdx.x % H);
  const int64_t idx = idxs[b];
  const int64_t init = init_idxs ? init_idxs[b] : 0;
  int sy = y, dx = 0;
  if (crop) {
    sy = min(max(y + (int)crop[b * 2 + 0] - pad, 0), H - 1);
    dx = (int)crop[b * 2 + 1] - pad;
  }
  const int CO = fs * C;
  const int64_t frame_bytes = (int64_t)H * W * C;
  for (int t = threadIdx.x; t < W * CO; t += blockDim.x) {
    const int x = t / CO, co = t % CO;
    const int j = co / C, c = co % C;
    const int sx = min(max(x + dx, 0), W - 1);
    const int64_t pix = ((int64_t)sy * W + sx) * C + c;
    const int64_t fo = max(idx - (int64_t)(fs - 1 - j), init);
    const int64_t out_off = (((int64_t)b * H + y) * W + x) * CO + co;
    obs_out[out_off] = O[fo * frame_bytes + pix];
    if (next_out) {
      uint8_t v;
      if (j == fs - 1) v = NO[idx * frame_bytes + pix];
      else v = O[max(idx - (int64_t)(fs - 2 - j), init) * frame_bytes + pix];
      next_out[out_off] = v;
    }
  }
}

// ---- Philox-4x32-10 ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; i++) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// tensor ids 0..4 = z_next, x0, t, z, z_metric.  normal via Box-Muller, uniform = 24 random bits in [0,1).
// The Philox counter is the element's index in the GLOBAL [S][global_batch][A] tensor (quad = index / 4, lane = index % 4): a
// data-parallel rank that holds rows [row0, row0 + B) of every seed draws exactly the slice a single device would have drawn for
// those rows (R ranks == 1 device), and no two ranks share a noise row.
__global__ void fill_noise_kernel(uint64_t seed, uint64_t step, float* z_next, float* x0, float* t, float* z, float* z_metric,
                                  int S, int B, int A, int64_t GB, int64_t row0) {
  const int tensor = blockIdx.y;
  float* out = tensor == 0 ? z_next : tensor == 1 ? x0 : tensor == 2 ? t : tensor == 3 ? z : z_metric;
  const int w = tensor == 2 ? 1 : A;
  const int64_t n = (int64_t)S * B * w;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // local element
  if (out == nullptr || i >= n) return;
  const int64_t per_seed = (int64_t)B * w;
  const int64_t s = i / per_seed, r = i - s * per_seed;
  const int64_t ge = (s * GB + row0) * w + r;                        // global element
  const int64_t q = ge >> 2;
  const int lane = (int)(ge & 3);
  uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32) ^ ((uint32_t)tensor << 28), (uint32_t)step, (uint32_t)(step >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  float v;
  if (tensor == 2) {
    v = (float)(c[lane] >> 8) * (1.0f / 16777216.0f);
  } else {
    const int p = lane & 2;
    const float u1 = ((float)(c[p] >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
    const float u2 = (float)(c[p + 1] >> 8) * (1.0f / 16777216.0f);
    const float rr = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    v = (lane & 1) ? rr * sn : rr * cs;
  }
  out[i] = v;
}

}  // namespace

extern "C" int fql_gather_rows(const void* src, void* dst, const int64_t* idxs, int64_t n_idx, int64_t row_bytes, void* stream) {
  FQL_REQUIRE(src && dst && idxs && n_idx >= 0 && row_bytes > 0, "fql_gather_rows: bad arguments");
  if (n_idx == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)(n_idx < 65535 * 16 ? n_idx : 65535 * 16);
  const uintptr_t al = (uintptr_t)src | (uintptr_t)dst | (uintptr_t)row_bytes;
  if ((al & 15) == 0) {
    const int64_t re = row_bytes / 16;
    gather_rows_kernel<uint4><<<grid, re >= 256 ? 256 : (re > 32 ? 128 : 32), 0, st>>>((const uint4*)src, (uint4*)dst, idxs, n_idx, re);
  } else if ((al & 3) == 0) {
    const int64_t re = row_bytes / 4;
    gather_rows_kernel<uint32_t><<<grid, re >= 256 ? 256 : (re > 32 ? 128 : 32), 0, st>>>((const uint32_t*)src, (uint32_t*)dst, idxs, n_idx, re);
  } else {
    gather_rows_kernel<uint8_t><<<grid, 128, 0, st>>>((const uint8_t*)src, (uint8_t*)dst, idxs, n_idx, row_bytes);
  }
  FQL_CHECK_LAUNCH();
  return 0;
}

extern "C" int fql_gather_frames(const uint8_t* obs_src, const uint8_t* next_src, uint8_t* obs_out, uint8_t* next_out,
                                 const int64_t* idxs, const int64_t* init_idxs, const int64_t* crop_from, int64_t n_idx,
                                 int32_t img_h, int32_t img_w, int32_t img_c, int32_t frame_stack, int32_t pad, void* stream) {
  FQL_REQUIRE(obs_src && obs_out && idxs, "fql_gather_frames: NULL argument");
  FQL_REQUIRE(img_h > 0 && img_w > 0 && img_c > 0 && frame_stack >= 1 && pad >= 0, "fql_gather_frames: bad geometry");
  FQL_REQUIRE(!next_out || next_src, "fql_gather_frames: next_out without next_src");
  FQL_REQUIRE(frame_stack == 1 || init_idxs, "fql_gather_frames: frame stacking needs init_idxs");
  if (n_idx == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int per_row = img_w * frame_stack * img_c;
  const int threads = per_row >= 512 ? 512 : (per_row >= 256 ? 256 : 128);
  gather_frames_kernel<<<(unsigned)(n_idx * img_h), threads, 0, st>>>(obs_src, next_src, obs_out, next_out, idxs, init_idxs,
                                                                    crop_from, img_h, img_w, img_c, frame_stack, pad);
  FQL_CHECK_LAUNCH();
  return 0;
}

extern "C" int fql_fill_noise_rows(const FqlDims* d, uint64_t seed, uint64_t step, int64_t row_offset, float* z_next, float* x0, float* t,
                                   float* z, float* z_metric, void* stream) {
  FQL_TRY(fql_validate_dims(d));
  FQL_REQUIRE(row_offset >= 0 && row_offset + d->batch <= d->global_batch, "fql_fill_noise_rows: rows [%lld, %lld) outside the global batch %d",
              (long long)row_offset, (long long)row_offset + d->batch, d->global_batch);
  const int64_t n_act = (int64_t)d->num_seeds * d->batch * d->action_dim;
  dim3 grid((unsigned)((n_act + 255) / 256), 5);
  fill_noise_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(seed, step, z_next, x0, t, z, z_metric, d->num_seeds, d->batch,
                                                                               d->action_dim, d->global_batch, row_offset);
  FQL_CHECK_LAUNCH();
  return 0;
}

extern "C" int fql_fill_noise(const FqlDims* d, uint64_t seed, uint64_t step, float* z_next, float* x0, float* t, float* z,
                              float* z_metric, void* stream) {
  FQL_REQUIRE(d && d->global_batch == d->batch, "fql_fill_noise: data-parallel ranks (global_batch != batch) must call fql_fill_noise_rows "
                                                "with their row offset, or every rank draws the same noise");
  return fql_fill_noise_rows(d, seed, step, 0, z_next, x0, t, z, z_metric, stream);
}
