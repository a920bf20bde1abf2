// encoder.cu -- ImpalaEncoder('impala_small') forward / backward (utils/encoders.py:10-108) in fp32, NHWC, on CUDA cores.
//   x = u8/255 -> 3 x [conv3x3 SAME -> max_pool 3x3/2 SAME(-inf) -> (relu -> conv -> relu -> conv) + skip] -> relu -> flatten
//     -> Dense(2048->512) -> gelu(tanh)                                     (encoders.py:83-100, networks.py:34-61 activate_final)
// Round-1 status: correct (parity-tested against oracle/encoder_oracle.py) but not yet tuned: direct convolutions with the
// weights in shared memory; the tensor-core implicit-GEMM version is a next-round item.
#include "step.cuh"

namespace {

constexpr int kStacks[3] = {16, 32, 32};

// One thread per output pixel, all COUT channels in registers, weights [9][CIN][COUT] in smem (broadcast reads).
//   in_u8 != NULL : input is uint8, scaled by 1/255 (encoders.py:84)         relu_in : apply relu to the input on load
//   flip          : use W'[ky][kx][co][ci] = W[2-ky][2-kx][ci][co]  (input gradient of the same convolution; then CIN/COUT swap)
//   skip          : out = conv + bias + skip                       mask_src : out *= (mask_src > 0)       accumulate : out += ...
template <int COUT>
__global__ void __launch_bounds__(128) conv3x3_kernel(const float* __restrict__ in_f, const uint8_t* __restrict__ in_u8,
                                                      const float* __restrict__ Wg, const float* __restrict__ bias,
                                                      const float* __restrict__ skip, const float* __restrict__ mask_src,
                                                      float* __restrict__ out, int B, int H, int W, int CIN, int relu_in, int flip,
                                                      int accumulate) {
  // each thread computes PX horizontally adjacent output pixels x all COUT channels: every weight fetched from smem feeds PX FMAs
  constexpr int PX = 4;
  extern __shared__ float ws[];  // [9][CIN][COUT]
  for (int i = threadIdx.x; i < 9 * CIN * COUT; i += blockDim.x) {
    const int co = i % COUT, ci = (i / COUT) % CIN, tap = i / (COUT * CIN);
    // forward: W[tap][ci][co] (HWIO).  flipped: the stored kernel is [3][3][COUT(as in)][CIN(as out)] -> W[8-tap][co][ci]
    ws[i] = flip ? Wg[((8 - tap) * COUT + co) * CIN + ci] : Wg[i];
  }
  __syncthreads();
  const int WQ = (W + PX - 1) / PX;
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (int64_t)B * H * WQ) return;
  const int x0 = (int)(q % WQ) * PX, y = (int)((q / WQ) % H);
  const int64_t b = q / ((int64_t)WQ * H);
  float acc[PX][COUT];
#pragma unroll
  for (int p = 0; p < PX; p++)
#pragma unroll
    for (int co = 0; co < COUT; co++) acc[p][co] = bias ? bias[co] : 0.f;
  for (int ky = 0; ky < 3; ky++) {
    const int iy = y + ky - 1;
    if (iy < 0 || iy >= H) continue;
    const int64_t rowoff = (b * H + iy) * W;
    for (int ci = 0; ci < CIN; ci++) {
      // the PX+2 input values of this row / channel that the PX outputs touch
      float v[PX + 2];
#pragma unroll
      for (int j = 0; j < PX + 2; j++) {
        const int ix = x0 + j - 1;
        float t = 0.f;
        if (ix >= 0 && ix < W) {
          const int64_t o = (rowoff + ix) * CIN + ci;
          t = in_u8 ? (float)in_u8[o] / 255.0f : in_f[o];
          if (relu_in) t = fmaxf(t, 0.f);
        }
        v[j] = t;
      }
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        const float4* w4 = reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * CIN + ci) * COUT);
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; c4++) {
          const float4 w = w4[c4];
#pragma unroll
          for (int p = 0; p < PX; p++) {
            const float a = v[p + kx];
            acc[p][c4 * 4 + 0] = fmaf(a, w.x, acc[p][c4 * 4 + 0]);
            acc[p][c4 * 4 + 1] = fmaf(a, w.y, acc[p][c4 * 4 + 1]);
            acc[p][c4 * 4 + 2] = fmaf(a, w.z, acc[p][c4 * 4 + 2]);
            acc[p][c4 * 4 + 3] = fmaf(a, w.w, acc[p][c4 * 4 + 3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < PX; p++) {
    const int x = x0 + p;
    if (x >= W) break;
    const int64_t pix = (b * H + y) * W + x;
    float* o = out + pix * COUT;
#pragma unroll
    for (int co = 0; co < COUT; co++) {
      float t = acc[p][co];
      if (skip) t += skip[pix * COUT + co];
      if (mask_src) t = mask_src[pix * COUT + co] > 0.f ? t : 0.f;
      o[co] = accumulate ? o[co] + t : t;
    }
  }
}

// Input gradient towards a runtime channel count (the first conv of a stack: COUT_fwd -> CIN_fwd in {img_c,16,32}): generic version
__global__ void __launch_bounds__(128) conv3x3_dgrad_generic_kernel(const float* __restrict__ dout, const float* __restrict__ Wg,
                                                                    float* __restrict__ din, int B, int H, int W, int CIN, int COUT) {
  extern __shared__ float ws[];  // W[tap][ci][co]
  for (int i = threadIdx.x; i < 9 * CIN * COUT; i += blockDim.x) ws[i] = Wg[i];
  __syncthreads();
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (int64_t)B * H * W) return;
  const int x = (int)(pix % W), y = (int)((pix / W) % H);
  const int64_t b = pix / ((int64_t)W * H);
  for (int ci = 0; ci < CIN; ci++) {
    float acc = 0.f;
    for (int ky = 0; ky < 3; ky++) {
      const int oy = y - ky + 1;
      if (oy < 0 || oy >= H) continue;
      for (int kx = 0; kx < 3; kx++) {
        const int ox = x - kx + 1;
        if (ox < 0 || ox >= W) continue;
        const float* d = dout + ((b * H + oy) * W + ox) * COUT;
        const float* w = ws + ((ky * 3 + kx) * CIN + ci) * COUT;
        for (int co = 0; co < COUT; co++) acc = fmaf(d[co], w[co], acc);
      }
    }
    din[pix * CIN + ci] = acc;
  }
}

// Weight + bias gradient: one CTA per image, partial sums [B][9*CIN*COUT + COUT]; rows of the image staged in smem.
__global__ void __launch_bounds__(256) conv3x3_wgrad_kernel(const float* __restrict__ in_f, const uint8_t* __restrict__ in_u8,
                                                            const float* __restrict__ dout, float* __restrict__ partial, int H, int W,
                                                            int CIN, int COUT, int relu_in, int rows_per_cta) {
  extern __shared__ float sm[];
  float* s_in = sm;                              // [3][W+2][CIN]
  float* s_do = sm + 3 * (W + 2) * CIN;          // [W][COUT]
  const int b = blockIdx.x, y_begin = blockIdx.y * rows_per_cta, y_end = min(H, y_begin + rows_per_cta);
  const int nout = 9 * CIN * COUT;
  constexpr int MAXO = 36;                       // ceil(9*32*32 / 256)
  // a row's 64-term dot product runs in fp32; rows are accumulated in double: these are sums of up to B*H*W signed terms with
  // heavy cancellation (the critic's TD errors change sign), and a plain fp32 running sum misses the 1e-5 parity bar
  double acc[MAXO];
#pragma unroll
  for (int k = 0; k < MAXO; k++) acc[k] = 0.0;
  double bacc = 0.0;
  for (int y = y_begin; y < y_end; y++) {
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * (W + 2) * CIN; i += blockDim.x) {
      const int ci = i % CIN, xx = (i / CIN) % (W + 2), r = i / (CIN * (W + 2));
      const int iy = y + r - 1, ix = xx - 1;
      float v = 0.f;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        const int64_t off = (((int64_t)b * H + iy) * W + ix) * CIN + ci;
        v = in_u8 ? (float)in_u8[off] / 255.0f : in_f[off];
        if (relu_in) v = fmaxf(v, 0.f);
      }
      s_in[i] = v;
    }
    for (int i = threadIdx.x; i < W * COUT; i += blockDim.x) s_do[i] = dout[(((int64_t)b * H + y) * W) * COUT + i];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MAXO; k++) {
      const int o = threadIdx.x + k * 256;
      if (o < nout) {
        const int co = o % COUT, ci = (o / COUT) % CIN, tap = o / (COUT * CIN);
        const int ky = tap / 3, kx = tap % 3;
        const float* ip = s_in + (ky * (W + 2) + kx) * CIN + ci;
        float a = 0.f;
        for (int x = 0; x < W; x++) a = fmaf(ip[x * CIN], s_do[x * COUT + co], a);
        acc[k] += (double)a;
      }
    }
    if (threadIdx.x < COUT) {
      float r = 0.f;
      for (int x = 0; x < W; x++) r += s_do[x * COUT + threadIdx.x];
      bacc += (double)r;
    }
  }
  float* p = partial + ((int64_t)b * gridDim.y + blockIdx.y) * (nout + COUT);
#pragma unroll
  for (int k = 0; k < MAXO; k++) {
    const int o = threadIdx.x + k * 256;
    if (o < nout) p[o] = (float)acc[k];
  }
  if (threadIdx.x < COUT) p[nout + threadIdx.x] = (float)bacc;
}

// out[i] = sum_b partial[b][i]   (deterministic); writes the kernel gradient and the bias gradient
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int B, int n_w, int n_b, float* __restrict__ gw, float* __restrict__ gb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_w + n_b) return;
  double s = 0.0;
  for (int b = 0; b < B; b++) s += (double)partial[(int64_t)b * (n_w + n_b) + i];
  if (i < n_w) gw[i] = (float)s;
  else gb[i - n_w] = (float)s;
}

// max_pool 3x3 stride 2 'SAME' with -inf padding: total pad = (Ho-1)*2+3-H, low pad = total/2 (0 for even H, 1 for odd H);
// window o covers inputs [2o - lo, 2o - lo + 2]; the first maximum in row-major order wins.
__global__ void maxpool_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, uint8_t* __restrict__ arg, int B, int H, int W, int C) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * Ho * Wo * C) return;
  const int c = (int)(i % C), xo = (int)((i / C) % Wo), yo = (int)((i / ((int64_t)C * Wo)) % Ho);
  const int64_t b = i / ((int64_t)C * Wo * Ho);
  const int lo_h = ((Ho - 1) * 2 + 3 - H) / 2, lo_w = ((Wo - 1) * 2 + 3 - W) / 2;
  float best = -INFINITY;
  int bk = -1;
  for (int ky = 0; ky < 3; ky++) {
    const int iy = 2 * yo + ky - lo_h;
    if (iy < 0 || iy >= H) continue;
    for (int kx = 0; kx < 3; kx++) {
      const int ix = 2 * xo + kx - lo_w;
      if (ix < 0 || ix >= W) continue;
      const float v = in[((b * H + iy) * W + ix) * C + c];
      if (v > best || bk < 0) { best = v; bk = ky * 3 + kx; }
    }
  }
  out[i] = best;
  arg[i] = (uint8_t)bk;
}

// gather form of the pooling gradient: an input pixel sums the windows whose argmax points at it
__global__ void maxpool_bwd_kernel(const float* __restrict__ dout, const uint8_t* __restrict__ arg, float* __restrict__ din, int B, int H,
                                   int W, int C) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H * W * C) return;
  const int c = (int)(i % C), x = (int)((i / C) % W), y = (int)((i / ((int64_t)C * W)) % H);
  const int64_t b = i / ((int64_t)C * W * H);
  const int lo_h = ((Ho - 1) * 2 + 3 - H) / 2, lo_w = ((Wo - 1) * 2 + 3 - W) / 2;
  float s = 0.f;
  for (int yo = (y + lo_h) / 2 - 1; yo <= (y + lo_h) / 2; yo++) {       // windows with 2*yo - lo <= y <= 2*yo - lo + 2
    const int ky = y + lo_h - 2 * yo;
    if (yo < 0 || yo >= Ho || ky < 0 || ky > 2) continue;
    for (int xo = (x + lo_w) / 2 - 1; xo <= (x + lo_w) / 2; xo++) {
      const int kx = x + lo_w - 2 * xo;
      if (xo < 0 || xo >= Wo || kx < 0 || kx > 2) continue;
      const int64_t o = ((b * Ho + yo) * Wo + xo) * C + c;
      if (arg[o] == ky * 3 + kx) s += dout[o];
    }
  }
  din[i] = s;
}

__global__ void relu_copy_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fmaxf(in[i], 0.f);
}
// dz = dout * gelu'(z)
__global__ void gelu_grad_mul_kernel(const float* __restrict__ dout, const float* __restrict__ z, float* __restrict__ dz, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dz[i] = dout[i] * gelu_tanh_grad_f(z[i]);
}
// dx = dflat * (x_last > 0)
__global__ void relu_mask_kernel(const float* __restrict__ d, const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[i] > 0.f ? d[i] : 0.f;
}

// out[m][f] = sum_e dX0[e][m][f], f < F: the feature part of a first-layer input gradient (critic input is broadcast to both heads)
__global__ void extract_feat_grad_kernel(const float* __restrict__ dX0, float* __restrict__ out, int E, int64_t M, int K0, int F) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * F) return;
  const int64_t m = i / F;
  const int f = (int)(i % F);
  float s = 0.f;
  for (int e = 0; e < E; e++) s += dX0[((int64_t)e * M + m) * K0 + f];
  out[i] = s;
}

template <typename K, typename... Args>
int launch1d(K kern, int64_t n, cudaStream_t st, Args... args) {
  if (n <= 0) return 0;
  kern<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(args...);
  FQL_CHECK_LAUNCH();
  return 0;
}

int conv_fwd(const float* in_f, const uint8_t* in_u8, const float* Wg, const float* bias, const float* skip, const float* mask,
             float* out, int B, int H, int W, int CIN, int COUT, int relu_in, int flip, int accumulate, cudaStream_t st) {
  const int64_t nq = (int64_t)B * H * ((W + 3) / 4);
  const size_t smem = (size_t)9 * CIN * COUT * sizeof(float);
  const unsigned grid = (unsigned)((nq + 127) / 128);
  if (COUT == 16) conv3x3_kernel<16><<<grid, 128, smem, st>>>(in_f, in_u8, Wg, bias, skip, mask, out, B, H, W, CIN, relu_in, flip, accumulate);
  else if (COUT == 32) conv3x3_kernel<32><<<grid, 128, smem, st>>>(in_f, in_u8, Wg, bias, skip, mask, out, B, H, W, CIN, relu_in, flip, accumulate);
  else FQL_REQUIRE(false, "conv3x3: unsupported channel count %d", COUT);
  FQL_CHECK_LAUNCH();
  return 0;
}

int conv_wgrad(const float* in_f, const uint8_t* in_u8, const float* dout, float* partial, float* gw, float* gb, int B, int H, int W, int CIN,
               int COUT, int relu_in, cudaStream_t st) {
  FQL_REQUIRE(9 * CIN * COUT <= 36 * 256, "conv3x3 wgrad: too many weights");
  const size_t smem = ((size_t)3 * (W + 2) * CIN + (size_t)W * COUT) * sizeof(float);
  // one CTA per (image, chunk of rows): enough CTAs to fill the GPU at small batch, partials reduced deterministically afterwards
  const int chunks = (B >= 512 || H < 16) ? 1 : (H >= 32 ? 4 : 2);
  const int rows = (H + chunks - 1) / chunks;
  conv3x3_wgrad_kernel<<<dim3(B, chunks), 256, smem, st>>>(in_f, in_u8, dout, partial, H, W, CIN, COUT, relu_in, rows);
  FQL_CHECK_LAUNCH();
  const int n_w = 9 * CIN * COUT;
  reduce_partials_kernel<<<(n_w + COUT + 255) / 256, 256, 0, st>>>(partial, B * chunks, n_w, COUT, gw, gb);
  FQL_CHECK_LAUNCH();
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// buffers of one encoder pass
// ---------------------------------------------------------------------------------------------------------------
size_t enc_carve(const FqlDims* d, int64_t B, void* base, EncBuf* e, bool for_backward) {
  if (d->precision != FQL_PRECISION_FP32) return enc_tc_carve(d, B, base, e, for_backward);
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](int64_t nfloats) {
    off = (off + 255) & ~(size_t)255;
    float* r = base ? reinterpret_cast<float*>(p + off) : nullptr;
    off += (size_t)nfloats * 4;
    return r;
  };
  int H = d->reserved[0], W = d->reserved[1];
  int64_t biggest = 0;
  for (int i = 0; i < 3; i++) {
    const int f = kStacks[i];
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    biggest = biggest > B * H * W * f ? biggest : B * H * W * f;
    e->pl[i] = take(B * Ho * Wo * f);
    e->c1[i] = take(B * Ho * Wo * f);
    e->x[i] = take(B * Ho * Wo * f);
    e->arg[i] = reinterpret_cast<uint8_t*>(take((B * Ho * Wo * f + 3) / 4));
    H = Ho; W = Wo;
  }
  e->flat_dim = H * W * kStacks[2];
  e->flat = take(B * e->flat_dim);
  e->z = take(B * d->obs_dim);
  e->scratch = take(biggest);   // pre-pool conv output / its gradient
  if (for_backward) {
    e->dz = take(B * d->obs_dim);
    e->dflat = take(B * e->flat_dim);
    int H2 = d->reserved[0], W2 = d->reserved[1];
    int64_t big2 = 0;
    for (int i = 0; i < 3; i++) {
      const int Ho = (H2 + 1) / 2, Wo = (W2 + 1) / 2;
      const int64_t n = B * Ho * Wo * kStacks[i];
      big2 = big2 > n ? big2 : n;
      H2 = Ho; W2 = Wo;
    }
    e->da = take(big2);
    e->db = take(big2);
    e->dc = take(big2);
    int wmax = 9 * 32 * 32 + 32;
    e->partial = take(4 * B * (int64_t)wmax);
  }
  return off + 256;
}

// features [B, 512] = encoder(obs_u8 [B,H,W,C]); `enc` = offsets of this encoder's leaves, params = arena of the seed
int enc_forward(const FqlDims* d, const EncView& v, const float* params, const uint8_t* obs, int64_t B, const EncBuf& e, float* feat,
                cudaStream_t st) {
  if (d->precision != FQL_PRECISION_FP32) return enc_tc_forward(d, v, params, obs, B, e, feat, st);
  int H = d->reserved[0], W = d->reserved[1], C = d->reserved[2];
  const float* xin_f = nullptr;
  const uint8_t* xin_u8 = obs;
  for (int i = 0; i < 3; i++) {
    const int f = kStacks[i];
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    FQL_TRY(conv_fwd(xin_f, xin_u8, params + v.off_cw[i][0], params + v.off_cb[i][0], nullptr, nullptr, e.scratch, (int)B, H, W, C, f, 0, 0, 0, st));
    FQL_TRY(launch1d(maxpool_fwd_kernel, B * Ho * Wo * f, st, e.scratch, e.pl[i], e.arg[i], (int)B, H, W, f));
    FQL_TRY(conv_fwd(e.pl[i], nullptr, params + v.off_cw[i][1], params + v.off_cb[i][1], nullptr, nullptr, e.c1[i], (int)B, Ho, Wo, f, f, 1, 0, 0, st));
    FQL_TRY(conv_fwd(e.c1[i], nullptr, params + v.off_cw[i][2], params + v.off_cb[i][2], e.pl[i], nullptr, e.x[i], (int)B, Ho, Wo, f, f, 1, 0, 0, st));
    xin_f = e.x[i]; xin_u8 = nullptr;
    H = Ho; W = Wo; C = f;
  }
  FQL_TRY(launch1d(relu_copy_kernel, B * e.flat_dim, st, e.x[2], e.flat, B * e.flat_dim));
  GemmArgs a;  // z = flat @ Wd + bd ; feat = gelu(z)
  memset(&a, 0, sizeof(a));
  a.P = 1; a.S = 1; a.E = 1; a.M = (int)B; a.N = d->obs_dim; a.K = e.flat_dim;
  a.A.base[0] = e.flat; a.lda = e.flat_dim;
  a.B.base[0] = params + v.off_dw; a.ldb = d->obs_dim;
  a.bias.base[0] = params + v.off_db;
  a.out_pre.base[0] = e.z; a.ld_pre = d->obs_dim;
  a.out.base[0] = feat; a.ldo = d->obs_dim;
  a.act_gelu = 1;
  FQL_TRY(launch_gemm(a, st));
  return 0;
}

// parameter gradients of one encoder given d(loss)/d(features) [B,512]; grads = gradient arena of the seed
int enc_backward(const FqlDims* d, const EncView& v, const float* params, float* grads, const uint8_t* obs, int64_t B, const EncBuf& e,
                 const float* dfeat, cudaStream_t st) {
  if (d->precision != FQL_PRECISION_FP32) return enc_tc_backward(d, v, params, grads, obs, B, e, dfeat, st);
  const int F = d->obs_dim;
  FQL_TRY(launch1d(gelu_grad_mul_kernel, B * F, st, dfeat, e.z, e.dz, B * F));
  GemmArgs a;  // dWd = flat^T dz
  memset(&a, 0, sizeof(a));
  a.P = 1; a.S = 1; a.E = 1; a.M = e.flat_dim; a.N = F; a.K = (int)B; a.trans_a = 1;
  a.A.base[0] = e.flat; a.lda = e.flat_dim;
  a.B.base[0] = e.dz; a.ldb = F;
  a.out.base[0] = grads + v.off_dw; a.ldo = F;
  FQL_TRY(launch_gemm(a, st));
  ColSumArgs c;
  memset(&c, 0, sizeof(c));
  c.P = 1; c.S = 1; c.E = 1; c.M = (int)B; c.N = F; c.ld = F;
  c.X.base[0] = e.dz;
  c.out.base[0] = grads + v.off_db;
  FQL_TRY(launch_colsum(c, st));
  memset(&a, 0, sizeof(a));  // dflat = dz Wd^T
  a.P = 1; a.S = 1; a.E = 1; a.M = (int)B; a.N = e.flat_dim; a.K = F; a.trans_b = 1;
  a.A.base[0] = e.dz; a.lda = F;
  a.B.base[0] = params + v.off_dw; a.ldb = F;
  a.out.base[0] = e.dflat; a.ldo = e.flat_dim;
  FQL_TRY(launch_gemm(a, st));
  // geometry of every stack
  int Hs[4], Ws[4], Cs[4];
  Hs[0] = d->reserved[0]; Ws[0] = d->reserved[1]; Cs[0] = d->reserved[2];
  for (int i = 0; i < 3; i++) { Hs[i + 1] = (Hs[i] + 1) / 2; Ws[i + 1] = (Ws[i] + 1) / 2; Cs[i + 1] = kStacks[i]; }
  float* dx = e.da;  // gradient w.r.t. the stack output x[i]
  float* other = e.dc;
  FQL_TRY(launch1d(relu_mask_kernel, B * e.flat_dim, st, e.dflat, e.x[2], dx, B * e.flat_dim));
  for (int i = 2; i >= 0; i--) {
    const int f = kStacks[i], Ho = Hs[i + 1], Wo = Ws[i + 1], H = Hs[i], W = Ws[i], C = Cs[i];
    const int64_t n = B * Ho * Wo * f;
    // conv2: input relu(c1), output gradient dx
    FQL_TRY(conv_wgrad(e.c1[i], nullptr, dx, e.partial, grads + v.off_cw[i][2], grads + v.off_cb[i][2], (int)B, Ho, Wo, f, f, 1, st));
    // dc1 = dgrad(dx, W2) * (c1 > 0)
    FQL_TRY(conv_fwd(dx, nullptr, params + v.off_cw[i][2], nullptr, nullptr, e.c1[i], e.db, (int)B, Ho, Wo, f, f, 0, 1, 0, st));
    // conv1: input relu(pl), output gradient dc1
    FQL_TRY(conv_wgrad(e.pl[i], nullptr, e.db, e.partial, grads + v.off_cw[i][1], grads + v.off_cb[i][1], (int)B, Ho, Wo, f, f, 1, st));
    // dpl = dx (skip) + dgrad(dc1, W1) * (pl > 0)   -> accumulate into dx in place
    FQL_TRY(conv_fwd(e.db, nullptr, params + v.off_cw[i][1], nullptr, nullptr, e.pl[i], dx, (int)B, Ho, Wo, f, f, 0, 1, 1, st));
    // dc0 = pool_bwd(dpl)
    FQL_TRY(launch1d(maxpool_bwd_kernel, B * H * W * f, st, dx, e.arg[i], e.scratch, (int)B, H, W, f));
    // conv0: input x[i-1] (or the pixels), output gradient dc0
    FQL_TRY(conv_wgrad(i ? e.x[i - 1] : nullptr, i ? nullptr : obs, e.scratch, e.partial, grads + v.off_cw[i][0], grads + v.off_cb[i][0], (int)B, H, W, C,
                       f, 0, st));
    if (i > 0) {  // gradient w.r.t. the previous stack's output (no relu between stacks): C = 16 or 32 output channels
      FQL_TRY(conv_fwd(e.scratch, nullptr, params + v.off_cw[i][0], nullptr, nullptr, nullptr, other, (int)B, H, W, f, C, 0, 1, 0, st));
      float* t = dx; dx = other; other = t;
    }
    (void)n;
  }
  return 0;
}

int launch_extract_feat_grad(const float* dX0, float* out, int E, int64_t M, int K0, int F, cudaStream_t st) {
  return launch1d(extract_feat_grad_kernel, M * F, st, dX0, out, E, M, K0, F);
}
