// train.cu -- the training step that feeds the denoiser (sm_100a, FP32).
//
//   corrupt_kernel      : dataset.py:120-238 (flips, Bernoulli(1/2) choice of multiplicative /
//                         additive truncated-normal noise, round) + the sub-sampling corruption
//                         of README.md:49-55.  Counter-based Philox4x32-10; every float op is an
//                         explicit round-to-nearest mul/add/div/sqrt so that an independent IEEE-754
//                         float32 restatement (the numpy checker used by tests/) reproduces the
//                         stream bit for bit.
//   loss kernels        : loss.py:40-131,190-247 (hinged MAE, RMSE, metrics), per-sample sums
//   train step          : train_loop.py:263-312 -- forward with BN batch statistics
//                         (backbone_blocks.py:167-246, Keras BatchNormalization(center=False)),
//                         loss, analytic backward (head, BN-train, conv dgrad/wgrad, ReLU mask),
//                         L1/L2 weight regularisation (loss.py:181-187), flat gradient vector in
//                         Keras trainable_variables order
//   adam kernels        : optimizer.py:145-224 (Adam + global_clipnorm), train_loop.py:421-434
//
// Feature maps are NHWC float32 over [n, h, w] (no pow2 canvas: the reference trains on the
// hydra model directly, train_loop.py:276-277).
#include <math.h>

#include "kernels.cuh"

namespace bfcnn {

// =====================================================================================
// Philox4x32-10 and the bit-reproducible normal sampler
// =====================================================================================
struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}

// tf.random.uniform float: 23 mantissa bits -> [0,1)
__device__ __forceinline__ float u01(uint32_t r) { return __fmul_rn((float)(r >> 9), 1.1920928955078125e-07f); }

// ln(u) for u in [1e-7, 1): u = m * 2^e, m in (sqrt(1/2), sqrt(2)]; ln m = 2 atanh((m-1)/(m+1))
__device__ __forceinline__ float ln_det(float u) {
  const uint32_t b = __float_as_uint(u);
  int e = (int)(b >> 23) - 127;
  float m = __uint_as_float((b & 0x007FFFFFu) | 0x3F800000u);
  if (m > 1.41421354f) { m = __fmul_rn(m, 0.5f); e += 1; }
  const float s = __fdiv_rn(__fadd_rn(m, -1.0f), __fadd_rn(m, 1.0f));
  const float s2 = __fmul_rn(s, s);
  float p = 0.222222224f;                            // 2/9
  p = __fadd_rn(__fmul_rn(p, s2), 0.285714298f);     // 2/7
  p = __fadd_rn(__fmul_rn(p, s2), 0.400000006f);     // 2/5
  p = __fadd_rn(__fmul_rn(p, s2), 0.666666687f);     // 2/3
  p = __fadd_rn(__fmul_rn(p, s2), 2.0f);
  return __fadd_rn(__fmul_rn((float)e, 0.693147182f), __fmul_rn(s, p));
}

// sin, cos of 2*pi*v for v in [0,1): quadrant reduction + Taylor polynomials on [-pi/4, pi/4]
__device__ __forceinline__ void sincos_turns_det(float v, float& sn, float& cs) {
  const int q = (int)__fadd_rn(__fmul_rn(v, 4.0f), 0.5f);
  const float f = __fadd_rn(v, -__fmul_rn((float)q, 0.25f));
  const float a = __fmul_rn(f, 6.28318548f);
  const float a2 = __fmul_rn(a, a);
  float ps = 2.75573188e-06f;                          // 1/9!
  ps = __fadd_rn(__fmul_rn(ps, a2), -1.98412701e-04f); // -1/7!
  ps = __fadd_rn(__fmul_rn(ps, a2), 8.33333377e-03f);  // 1/5!
  ps = __fadd_rn(__fmul_rn(ps, a2), -1.66666672e-01f); // -1/3!
  ps = __fadd_rn(__fmul_rn(ps, a2), 1.0f);
  const float s0 = __fmul_rn(a, ps);
  float pc = -2.75573200e-07f;                         // -1/10!
  pc = __fadd_rn(__fmul_rn(pc, a2), 2.48015876e-05f);  // 1/8!
  pc = __fadd_rn(__fmul_rn(pc, a2), -1.38888892e-03f); // -1/6!
  pc = __fadd_rn(__fmul_rn(pc, a2), 4.16666679e-02f);  // 1/4!
  pc = __fadd_rn(__fmul_rn(pc, a2), -0.5f);
  pc = __fadd_rn(__fmul_rn(pc, a2), 1.0f);
  switch (q & 3) {
    case 0: sn = s0; cs = pc; break;
    case 1: sn = pc; cs = -s0; break;
    case 2: sn = -s0; cs = -pc; break;
    default: sn = -pc; cs = s0; break;
  }
}

// one Box-Muller pair (TF BoxMullerFloat: u1 clamped to 1e-7, z0 = sin*r, z1 = cos*r)
__device__ __forceinline__ void box_muller_det(uint32_t r0, uint32_t r1, float& z0, float& z1) {
  float u1 = u01(r0);
  if (u1 < 1.0e-7f) u1 = 1.0e-7f;
  const float rad = __fsqrt_rn(__fmul_rn(-2.0f, ln_det(u1)));
  float sn, cs;
  sincos_turns_det(u01(r1), sn, cs);
  z0 = __fmul_rn(sn, rad);
  z1 = __fmul_rn(cs, rad);
}

// tf.random.truncated_normal: standard normals with |z| >= 2 rejected and re-drawn.
// Value slot `slot` of pixel `pix` of sample (g_lo, g_hi): attempt a uses counter
// (pix, slot | a << 8, g_lo, g_hi); the first of its four normals inside (-2, 2) wins.
__device__ __forceinline__ float truncated_normal_det(uint32_t pix, uint32_t slot, uint32_t g_lo, uint32_t g_hi,
                                                      uint32_t k0, uint32_t k1) {
  for (uint32_t a = 0;; ++a) {
    const U4 r = philox4x32_10(pix, slot | (a << 8), g_lo, g_hi, k0, k1);
    float z0, z1;
    box_muller_det(r.x, r.y, z0, z1);
    if (fabsf(z0) < 2.0f) return z0;
    if (fabsf(z1) < 2.0f) return z1;
    box_muller_det(r.z, r.w, z0, z1);
    if (fabsf(z0) < 2.0f) return z0;
    if (fabsf(z1) < 2.0f) return z1;
  }
}

__global__ void __launch_bounds__(256)
corrupt_kernel(const uint8_t* __restrict__ clean_u8, float* __restrict__ clean_f32, float* __restrict__ noisy_f32,
               int n, int h, int w, uint64_t seed, uint64_t sample_offset, bfcnn_noise_cfg cfg) {
  const int s = blockIdx.y;
  const uint64_t g = sample_offset + (uint64_t)s;
  const uint32_t g_lo = (uint32_t)g, g_hi = (uint32_t)(g >> 32);
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  // The decisions of dataset.py:141-142,170-187 (flips, noise on/off, sigmas) are drawn once per CALL of prepare_data_fn in
  // the reference, and a call sees the no_crops_per_image crops of ONE image (dataset.py:276-297: load + random_crops ->
  // prepare -> unbatch -> shuffle -> batch).  draw_group = k > 1: runs of k consecutive global sample indices (the crops of
  // one image) share the draws of the run's first index; 0 / 1: every sample draws for itself (k = 1 crop per image, the
  // reference default); < 0: the whole call shares the draws of sample_offset.  The per-pixel noise is always per sample.
  uint64_t gd = g;
  if (cfg.draw_group > 1) gd = g - g % (uint64_t)cfg.draw_group;
  else if (cfg.draw_group < 0) gd = sample_offset;
  const uint32_t d_lo = (uint32_t)gd, d_hi = (uint32_t)(gd >> 32);
  const U4 ra = philox4x32_10(0u, 0xFFFFFFFFu, d_lo, d_hi, k0, k1);
  const U4 rb = philox4x32_10(1u, 0xFFFFFFFFu, d_lo, d_hi, k0, k1);
  const bool add_on = cfg.additive_max > 0.f, mul_on = cfg.multiplicative_max > 0.f;
  const bool flip_lr = cfg.random_left_right && (u01(ra.x) > 0.5f);
  const bool flip_ud = cfg.random_up_down && (u01(ra.y) > 0.5f);
  const bool use_add = add_on && (u01(ra.z) > 0.5f);
  const bool use_mul = mul_on && (u01(ra.w) > 0.5f);
  const float sigma_add = __fadd_rn(cfg.additive_min, __fmul_rn(__fadd_rn(cfg.additive_max, -cfg.additive_min), u01(rb.x)));
  const float sigma_mul = __fadd_rn(cfg.multiplicative_min,
                                    __fmul_rn(__fadd_rn(cfg.multiplicative_max, -cfg.multiplicative_min), u01(rb.y)));
  const bool subsample = cfg.subsample && (u01(rb.z) > 0.5f);

  const uint8_t* src = clean_u8 + (size_t)s * h * w * 3;
  const size_t base = (size_t)s * h * w * 3;
  const int npx = h * w;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npx; pix += gridDim.x * blockDim.x) {
    const int y = pix / w, x = pix - y * w;
    const int cy = flip_ud ? (h - 1 - y) : y, cx = flip_lr ? (w - 1 - x) : x;
    const uint8_t* pc = src + ((size_t)cy * w + cx) * 3;
    const uint8_t* pn = pc;
    if (subsample) {  // stride-2 decimation + nearest x2 up-sampling of the (flipped) image
      const int y2 = y & ~1, x2 = x & ~1;
      const int sy = flip_ud ? (h - 1 - y2) : y2, sx = flip_lr ? (w - 1 - x2) : x2;
      pn = src + ((size_t)sy * w + sx) * 3;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = (float)pn[c];
      if (use_mul) {  // x * TN(1, sigma_mul)          dataset.py:190-206
        const float z = truncated_normal_det((uint32_t)pix, (uint32_t)c, g_lo, g_hi, k0, k1);
        v = __fmul_rn(v, __fadd_rn(1.0f, __fmul_rn(sigma_mul, z)));
      }
      if (use_add) {  // x + TN(0, sigma_add)          dataset.py:209-225
        const float z = truncated_normal_det((uint32_t)pix, (uint32_t)(3 + c), g_lo, g_hi, k0, k1);
        v = __fadd_rn(v, __fmul_rn(sigma_add, z));
      }
      if (cfg.round_values) v = rintf(v);  // dataset.py:228
      noisy_f32[base + (size_t)pix * 3 + c] = v;
      if (clean_f32) clean_f32[base + (size_t)pix * 3 + c] = (float)pc[c];  // dataset.py:233-235
    }
  }
}

int run_corrupt(bfcnn_handle* h, const uint8_t* clean_u8, float* clean_f32, float* noisy_f32, int n, int height,
                int width, uint64_t seed, uint64_t sample_offset, const bfcnn_noise_cfg* cfg, cudaStream_t st) {
  BF_REQUIRE((long long)height * width < (1ll << 31), "image too large for the corruption kernel");
  BF_REQUIRE(n <= 65535, "batch too large for the corruption kernel");
  BF_REQUIRE(cfg->additive_max <= 0.f || cfg->additive_min <= cfg->additive_max, "additive noise range is inverted");
  BF_REQUIRE(cfg->multiplicative_max <= 0.f || cfg->multiplicative_min <= cfg->multiplicative_max,
             "multiplicative noise range is inverted");
  const int npx = height * width;
  dim3 grid((unsigned)std::min((npx + 255) / 256, 4 * h->sm_count), (unsigned)n);
  corrupt_kernel<<<grid, 256, 0, st>>>(clean_u8, clean_f32, noisy_f32, n, height, width, seed, sample_offset, *cfg);
  h->launches++;
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

// =====================================================================================
// multi-scale ground truth (utilities.py:625-685 multiscales_generator_fn; train_loop.py:239-247): one level =
// tf.nn.avg_pool2d(ksize 2x2, strides 2, VALID) -> clip [0,255] -> tf.round.  HBM-bound: 48 B read, 12 B written per
// output pixel; one thread per output pixel, the two input rows are read as 6 contiguous floats each.
// =====================================================================================
__global__ void __launch_bounds__(256)
downscale2x_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int h, int w, int clip_values, int round_values) {
  const int ho = h >> 1, wo = w >> 1;
  const long long total = (long long)n * ho * wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % wo), y = (int)((i / wo) % ho), b = (int)(i / ((long long)wo * ho));
    const float* r0 = in + (((size_t)b * h + 2 * y) * w + 2 * x) * 3;
    const float* r1 = r0 + (size_t)w * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = __fmul_rn(__fadd_rn(__fadd_rn(r0[c], r0[3 + c]), __fadd_rn(r1[c], r1[3 + c])), 0.25f);
      if (clip_values) v = fminf(fmaxf(v, 0.f), 255.f);
      if (round_values) v = rintf(v);
      out[(size_t)i * 3 + c] = v;
    }
  }
}

int run_downscale2x(bfcnn_handle* h, const float* in, float* out, int n, int height, int width, int clip_values, int round_values,
                    cudaStream_t st) {
  const long long total = (long long)n * (height / 2) * (width / 2);
  if (total == 0) return BFCNN_OK;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 8);
  downscale2x_kernel<<<blocks, 256, 0, st>>>(in, out, n, height, width, clip_values, round_values);
  h->launches++;
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

// =====================================================================================
// loss (loss.py)
// =====================================================================================
// keras.activations.relu(x, threshold=t, max_value=m): x*[x > t] (t != 0) or relu(x), then clip [0, m]
__device__ __forceinline__ float keras_relu(float x, float t, float m) {
  x = (t != 0.f) ? ((x > t) ? x : 0.f) : fmaxf(x, 0.f);
  return fminf(fmaxf(x, 0.f), m);
}
// d/dx of the above (mask is not differentiated; clip_by_value passes the gradient on [0, m])
__device__ __forceinline__ float keras_relu_grad(float x, float t, float m) {
  const bool pass = (t != 0.f) ? (x > t) : (x > 0.f);
  return (pass && x <= m) ? 1.f : 0.f;
}

struct LossTerms { float abs_e, hinged, sq, sqh; };
__device__ __forceinline__ LossTerms loss_terms(float e, float hinge, float cutoff) {
  LossTerms t;
  const float ae = fabsf(e);
  t.abs_e = keras_relu(ae, 0.f, 255.f);                 // mae_actual   loss.py:194-198
  t.hinged = keras_relu(ae, hinge, cutoff);             // loss.py:209-214
  const float d0 = keras_relu(e, 0.f, 255.f);           // mse_actual   loss.py:200-204 (relu of the SIGNED error)
  t.sq = d0 * d0;
  const float d1 = keras_relu(e, hinge, cutoff * cutoff);  // loss.py:232-237
  t.sqh = d1 * d1;
  return t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// block-reduce K floats per thread and add them (as doubles) to dst[0..K)
template <int K>
__device__ __forceinline__ void block_accumulate(float (&v)[K], double* dst, float* s_red /*[K]*/) {
  if (threadIdx.x < K) s_red[threadIdx.x] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float a = warp_sum(v[k]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_red[k], a);
  }
  __syncthreads();
  if (threadIdx.x < K) atomicAdd(&dst[threadIdx.x], (double)s_red[threadIdx.x]);
}

// sums[s][4] += {sum|e|, sum hinged, sum relu(e)^2, sum hinged-relu(e)^2} of sample s
__global__ void __launch_bounds__(256)
loss_reduce_kernel(const float* __restrict__ gt, const float* __restrict__ pred, double* __restrict__ sums,
                   int per_sample, float hinge, float cutoff) {
  __shared__ float s_red[4];
  const int s = blockIdx.y;
  const float* g = gt + (size_t)s * per_sample;
  const float* p = pred + (size_t)s * per_sample;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += gridDim.x * blockDim.x) {
    const LossTerms t = loss_terms(g[i] - p[i], hinge, cutoff);
    acc[0] += t.abs_e; acc[1] += t.hinged; acc[2] += t.sq; acc[3] += t.sqh;
  }
  block_accumulate<4>(acc, sums + 4 * s, s_red);
}

// scalars[0..4] = total, mae, rmse, hinged-mae, ssim loss ; coef[0] = c_mae ; coef[1+s] = c_mse[s]
__global__ void loss_finalize_kernel(const double* __restrict__ sums, const double* __restrict__ ssim_sums, double ssim_cnt, int n,
                                     int per_sample, bfcnn_loss_cfg cfg, float* __restrict__ scalars, float* __restrict__ coef) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double cnt = (double)per_sample;
  double mae = 0, hm = 0, rm = 0, rmh = 0;
  for (int s = 0; s < n; ++s) {
    mae += sums[4 * s + 0] / cnt;
    hm += sums[4 * s + 1] / cnt;
    rm += sqrt(sums[4 * s + 2] / cnt + 1e-3);      // DEFAULT_EPSILON constants.py:7
    const double r = sqrt(sums[4 * s + 3] / cnt + 1e-3);
    rmh += r;
    if (coef) coef[1 + s] = (cfg.mse_multiplier > 0.f) ? (float)((double)cfg.mse_multiplier / ((double)n * cnt * r)) : 0.f;
  }
  mae /= n; hm /= n; rm /= n; rmh /= n;
  double total = 0, ssim_loss = 0;
  if (cfg.mae_multiplier > 0.f) total += hm * (double)cfg.mae_multiplier;
  if (cfg.ssim_multiplier > 0.f && ssim_sums) {     // 1 - mean over images of (mean over windows and channels)
    double m = 0;
    for (int s = 0; s < n; ++s) m += ssim_sums[s] / ssim_cnt;
    ssim_loss = 1.0 - m / n;
    total += ssim_loss * (double)cfg.ssim_multiplier;
  }
  if (cfg.mse_multiplier > 0.f) total += rmh * (double)cfg.mse_multiplier;
  scalars[0] = (float)total; scalars[1] = (float)mae; scalars[2] = (float)rm; scalars[3] = (float)hm; scalars[4] = (float)ssim_loss;
  if (coef) coef[0] = (cfg.mae_multiplier > 0.f) ? (float)((double)cfg.mae_multiplier / ((double)n * cnt)) : 0.f;
}

static int loss_grid_x(const bfcnn_handle* h, int per_sample, int n) {
  const int want = (per_sample + 255) / 256;
  return std::max(1, std::min(want, (8 * h->sm_count + n - 1) / n));
}

// (run_loss is defined after the SSIM kernels)
// =====================================================================================
// SSIM term (row N3): tf.image.ssim(gt, pred, filter_size=7, max_val=255) as called by loss.py:217-225
// (TF 2.13 image_ops_impl._ssim_helper: 7x7 Gaussian (sigma 1.5, softmax-normalised), VALID windows, per channel
//  luminance * contrast-structure, mean over windows and channels, then loss = 1 - mean over the batch)
// =====================================================================================
constexpr int SS_F = 7;                       // filter size
constexpr int SS_TW = 32, SS_TH = 8;          // windows (forward) / pixels (backward) per CTA
constexpr float SS_C1 = (0.01f * 255.f) * (0.01f * 255.f), SS_C2 = (0.03f * 255.f) * (0.03f * 255.f);

// The 49 taps travel as a kernel argument (the constant bank of the launch): a __constant__ symbol is per DEVICE, and a
// process-wide "uploaded" flag left handles on a second GPU of the same process with an all-zero filter.
struct GaussTaps { float g[SS_F * SS_F]; };
static GaussTaps ssim_taps() {
  GaussTaps t;
  double g1[SS_F], sum = 0;
  for (int i = 0; i < SS_F; ++i) { const double c = i - (SS_F - 1) / 2.0; g1[i] = exp(-0.5 * c * c / (1.5 * 1.5)); sum += g1[i]; }
  for (int i = 0; i < SS_F; ++i)
    for (int j = 0; j < SS_F; ++j) t.g[i * SS_F + j] = (float)((g1[i] / sum) * (g1[j] / sum));   // softmax of the sum of exponents
  return t;
}

// per window and channel: S = l*cs and its partial derivatives w.r.t. (mean_y, E[xy], E[x^2+y^2]) -> maps [n][hv][wv][3][3];
// per-image sum of S -> ssum[n]
__global__ void __launch_bounds__(SS_TW * SS_TH)
ssim_forward_kernel(const float* __restrict__ gt, const float* __restrict__ pred, float* __restrict__ maps,
                    double* __restrict__ ssum, int h, int w, int hv, int wv, const GaussTaps taps) {
  __shared__ float s_x[(SS_TH + SS_F - 1) * (SS_TW + SS_F - 1) * 3];
  __shared__ float s_y[(SS_TH + SS_F - 1) * (SS_TW + SS_F - 1) * 3];
  __shared__ float s_red[1];
  constexpr int PW = SS_TW + SS_F - 1, PH = SS_TH + SS_F - 1;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * SS_TW, y0 = blockIdx.y * SS_TH, b = blockIdx.z;
  const float* gb = gt + (size_t)b * h * w * 3;
  const float* pb = pred + (size_t)b * h * w * 3;
  for (int i = tid; i < PH * PW * 3; i += SS_TW * SS_TH) {
    const int c = i % 3, p = i / 3;
    const int lx = p % PW, ly = p / PW;
    const int gx = x0 + lx, gy = y0 + ly;
    float xv = 0.f, yv = 0.f;
    if (gx < w && gy < h) { xv = gb[((size_t)gy * w + gx) * 3 + c]; yv = pb[((size_t)gy * w + gx) * 3 + c]; }
    s_x[i] = xv; s_y[i] = yv;
  }
  __syncthreads();
  const int lx = tid % SS_TW, ly = tid / SS_TW;
  const int wx = x0 + lx, wy = y0 + ly;
  float acc = 0.f;
  if (wx < wv && wy < hv) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float mx = 0.f, my = 0.f, exy = 0.f, e2 = 0.f;
      for (int dy = 0; dy < SS_F; ++dy)
#pragma unroll
        for (int dx = 0; dx < SS_F; ++dx) {
          const float g = taps.g[dy * SS_F + dx];
          const int o = ((ly + dy) * PW + lx + dx) * 3 + c;
          const float xv = s_x[o], yv = s_y[o];
          mx = fmaf(g, xv, mx); my = fmaf(g, yv, my);
          exy = fmaf(g, xv * yv, exy); e2 = fmaf(g, xv * xv + yv * yv, e2);
        }
      const float num0 = 2.f * mx * my, den0 = mx * mx + my * my;
      const float lden = den0 + SS_C1, l = (num0 + SS_C1) / lden;
      const float Nn = 2.f * exy - num0 + SS_C2, D = e2 - den0 + SS_C2;
      const float cs = Nn / D;
      acc += l * cs;
      const float dl_dmy = (2.f * mx * lden - (num0 + SS_C1) * 2.f * my) / (lden * lden);
      const float dcs_dmy = (-2.f * mx * D + 2.f * my * Nn) / (D * D);
      float* m = maps + ((((size_t)b * hv + wy) * wv + wx) * 3 + c) * 3;
      m[0] = cs * dl_dmy + l * dcs_dmy;     // dS/d mean_y
      m[1] = l * 2.f / D;                   // dS/d E[xy]
      m[2] = -l * Nn / (D * D);             // dS/d E[x^2 + y^2]
    }
  }
  if (tid == 0) s_red[0] = 0.f;
  __syncthreads();
  const float wsum = warp_sum(acc);
  if ((tid & 31) == 0) atomicAdd(&s_red[0], wsum);
  __syncthreads();
  if (tid == 0) atomicAdd(&ssum[b], (double)s_red[0]);
}

// dpred[p][c] = coef * sum_{windows containing p} g * (dS/dmy + dS/dExy * x_p + dS/dE2 * 2 y_p)
__global__ void __launch_bounds__(SS_TW * SS_TH)
ssim_backward_kernel(const float* __restrict__ gt, const float* __restrict__ pred, const float* __restrict__ maps,
                     float* __restrict__ dpred, int h, int w, int hv, int wv, float coef, const GaussTaps taps) {
  __shared__ float s_m[(SS_TH + SS_F - 1) * (SS_TW + SS_F - 1) * 9];
  constexpr int PW = SS_TW + SS_F - 1, PH = SS_TH + SS_F - 1;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * SS_TW, y0 = blockIdx.y * SS_TH, b = blockIdx.z;
  // windows (wy, wx) with wy in [y - 6, y], wx in [x - 6, x]: tile origin (y0 - 6, x0 - 6)
  for (int i = tid; i < PH * PW * 9; i += SS_TW * SS_TH) {
    const int k = i % 9, p = i / 9;
    const int lx = p % PW, ly = p / PW;
    const int wx = x0 + lx - (SS_F - 1), wy = y0 + ly - (SS_F - 1);
    float v = 0.f;
    if (wx >= 0 && wx < wv && wy >= 0 && wy < hv) v = maps[(((size_t)b * hv + wy) * wv + wx) * 9 + k];
    s_m[i] = v;
  }
  __syncthreads();
  const int lx = tid % SS_TW, ly = tid / SS_TW;
  const int gx = x0 + lx, gy = y0 + ly;
  if (gx >= w || gy >= h) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float sa = 0.f, sb = 0.f, sc = 0.f;
    for (int dy = 0; dy < SS_F; ++dy)
#pragma unroll
      for (int dx = 0; dx < SS_F; ++dx) {
        // pixel p sits at offset (dy, dx) inside window (gy - dy, gx - dx)
        const float g = taps.g[dy * SS_F + dx];
        const float* m = s_m + (((ly + (SS_F - 1) - dy) * PW + lx + (SS_F - 1) - dx) * 3 + c) * 3;
        sa = fmaf(g, m[0], sa); sb = fmaf(g, m[1], sb); sc = fmaf(g, m[2], sc);
      }
    const size_t o = (((size_t)b * h + gy) * w + gx) * 3 + c;
    dpred[o] = coef * (sa + sb * gt[o] + sc * 2.f * pred[o]);
  }
}

// SSIM forward: per-window derivative maps + per-image sums (ssim_sums[n] must be zeroed by the caller)
static int run_ssim(bfcnn_handle* h, const float* gt, const float* pred, int n, int height, int width, float* maps,
                    double* ssim_sums, cudaStream_t st) {
  BF_REQUIRE(height >= SS_F && width >= SS_F, "SSIM needs height and width >= 7 (tf.image.ssim, filter_size=7)");
  const int hv = height - SS_F + 1, wv = width - SS_F + 1;
  dim3 grid((wv + SS_TW - 1) / SS_TW, (hv + SS_TH - 1) / SS_TH, n);
  BF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "image too large for the SSIM grid");
  ssim_forward_kernel<<<grid, SS_TW * SS_TH, 0, st>>>(gt, pred, maps, ssim_sums, height, width, hv, wv, ssim_taps());
  h->launches++;
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

int run_loss(bfcnn_handle* h, const float* gt, const float* pred, int n, int height, int width,
             const bfcnn_loss_cfg* cfg, float* out5, cudaStream_t st) {
  BF_REQUIRE(n <= 65535, "batch too large");
  BF_REQUIRE((long long)height * width * 3 < (1ll << 31), "image too large");
  const int per_sample = height * width * 3;
  const bool use_ssim = cfg->ssim_multiplier > 0.f;
  BF_CHECK(h->ws_stats.reserve((size_t)(5 * n) * sizeof(double) + 64));
  double* sums = h->ws_stats.as<double>();
  double* ssim_sums = sums + 4 * n;
  float* scal = reinterpret_cast<float*>(sums + 5 * n);
  BF_CUDA(cudaMemsetAsync(sums, 0, (size_t)5 * n * sizeof(double), st));
  dim3 grid((unsigned)loss_grid_x(h, per_sample, n), (unsigned)n);
  loss_reduce_kernel<<<grid, 256, 0, st>>>(gt, pred, sums, per_sample, cfg->hinge, cfg->cutoff);
  double ssim_cnt = 1;
  if (use_ssim) {
    const size_t nwin = (size_t)n * (height - SS_F + 1 > 0 ? height - SS_F + 1 : 0) * (width - SS_F + 1 > 0 ? width - SS_F + 1 : 0);
    BF_CHECK(h->ws_grads.reserve(std::max<size_t>(nwin, 1) * 9 * sizeof(float)));
    BF_CHECK(run_ssim(h, gt, pred, n, height, width, h->ws_grads.as<float>(), ssim_sums, st));
    ssim_cnt = 3.0 * (height - SS_F + 1) * (width - SS_F + 1);
  }
  loss_finalize_kernel<<<1, 32, 0, st>>>(sums, use_ssim ? ssim_sums : nullptr, ssim_cnt, n, per_sample, *cfg, scal, nullptr);
  h->launches += 2;
  BF_CUDA(cudaGetLastError());
  BF_CUDA(cudaMemcpyAsync(out5, scal, 5 * sizeof(float), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  return BFCNN_OK;
}

// =====================================================================================
// training forward helpers
// =====================================================================================
// per-step derived weights: collapsed head [16][4] and the dgrad kernels
// dgrad[l][8 - tap][co][ci] = W_l[tap][ci][co]  (a correlation of dOut with the flipped, transposed kernel)
__global__ void __launch_bounds__(256)
train_prep_kernel(const float* __restrict__ vars, const long long* __restrict__ conv_off /*[2N]*/, int nconv,
                  long long h0_off, long long h1_off, int F, float* __restrict__ dgrad_w, float* __restrict__ head_c) {
  const int l = blockIdx.x;
  if (l < nconv) {
    const float* w = vars + conv_off[l];
    float* d = dgrad_w + (size_t)l * 9 * C * C;
    for (int i = threadIdx.x; i < 9 * C * C; i += blockDim.x) {
      const int co = i % C, ci = (i / C) % C, tap = i / (C * C);
      d[((8 - tap) * C + co) * C + ci] = w[i];
    }
  } else {
    for (int i = threadIdx.x; i < C * 4; i += blockDim.x) {
      const int ci = i >> 2, o = i & 3;
      float a = 0.f;
      if (o < 3)
        for (int f = 0; f < F; ++f) a = fmaf(vars[h0_off + ci * F + f], vars[h1_off + f * 3 + o], a);
      head_c[i] = a;
    }
  }
}

// batch statistics -> (mean, biased var, invstd, scale = gamma*invstd); optional moving update
// (Keras fused BN: moving <- moving*m + batch*(1-m), moving_var takes the unbiased variance)
__global__ void bn_finalize_kernel(const double* __restrict__ stats /*[2][16]*/, double cnt, float eps, float momentum,
                                   float* __restrict__ vars, long long gamma_off, long long mean_off, long long var_off,
                                   float* __restrict__ bn /*[4][16]*/, float* __restrict__ coef /*[3][16]*/, int update_moving) {
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");   // the next conv may run its setup now (conv_t5.cu, launch_pdl)
  const int c = threadIdx.x;
  if (c >= C) return;
  const double m = stats[c] / cnt;
  double v = stats[C + c] / cnt - m * m;
  if (v < 0) v = 0;
  const double inv = 1.0 / sqrt(v + (double)eps);
  bn[c] = (float)m; bn[C + c] = (float)v; bn[2 * C + c] = (float)inv;
  bn[3 * C + c] = (float)((double)vars[gamma_off + c] * inv);
  // x_next = x + (u - mean) * scale as ca * x + cb * u + cc: the fused prologue of the next conv_a (conv_t5.cu)
  coef[c] = 1.0f; coef[C + c] = bn[3 * C + c]; coef[2 * C + c] = (float)(-m * (double)bn[3 * C + c]);
  if (update_moving) {
    const double ub = v * (cnt / fmax(cnt - 1.0, 1.0));
    vars[mean_off + c] = (float)((double)vars[mean_off + c] * (double)momentum + m * (1.0 - (double)momentum));
    vars[var_off + c] = (float)((double)vars[var_off + c] * (double)momentum + ub * (1.0 - (double)momentum));
  }
}

// x_next = x + (u - mean) * scale          (BN(center=False) then Add, backbone_blocks.py:191-242)
__global__ void __launch_bounds__(256)
bn_residual_kernel(const float4* __restrict__ x, const float4* __restrict__ u, float4* __restrict__ out,
                   const float* __restrict__ bn, long long n4) {
  __shared__ float s_bn[4 * C];
  if (threadIdx.x < 4 * C) s_bn[threadIdx.x] = bn[threadIdx.x];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 3) * 4;
    const float4 xv = x[i], uv = u[i];
    float4 r;
    r.x = xv.x + (uv.x - s_bn[c + 0]) * s_bn[3 * C + c + 0];
    r.y = xv.y + (uv.y - s_bn[c + 1]) * s_bn[3 * C + c + 1];
    r.z = xv.z + (uv.z - s_bn[c + 2]) * s_bn[3 * C + c + 2];
    r.w = xv.w + (uv.w - s_bn[c + 3]) * s_bn[3 * C + c + 3];
    out[i] = r;
  }
}

// -------------------------------------------------------------------------------------
// head: y = Wc^T x ; t = tanh(2y)*0.51 ; pred = (clip(t,+-0.5)+0.5)*255   (model.py:297-342)
// -------------------------------------------------------------------------------------
__device__ __forceinline__ void head_forward_px(const float (&x)[C], const float* __restrict__ s_wc /*[16][4]*/,
                                                float (&th)[3], float (&pred)[3]) {
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    float y = 0.f;
#pragma unroll
    for (int ci = 0; ci < C; ++ci) y = fmaf(x[ci], s_wc[ci * 4 + o], y);
    th[o] = tanhf(2.0f * y);
    const float t = th[o] * 0.51f;
    pred[o] = (fminf(fmaxf(t, -0.5f), 0.5f) + 0.5f) * 255.0f;
  }
}

__device__ __forceinline__ void load_px16(const float* __restrict__ p, float (&x)[C]) {
  float4 v[4];
  ldg256(p, v[0], v[1]);
  ldg256(p + 8, v[2], v[3]);
#pragma unroll
  for (int q = 0; q < 4; ++q) { x[4 * q] = v[q].x; x[4 * q + 1] = v[q].y; x[4 * q + 2] = v[q].z; x[4 * q + 3] = v[q].w; }
}

// forward head + per-sample loss sums in one pass over the last feature map
__global__ void __launch_bounds__(256)
head_loss_kernel(const float* __restrict__ feat, const float* __restrict__ gt, const float* __restrict__ wc,
                 double* __restrict__ sums, float* __restrict__ pred_out /* [n,h,w,3] or nullptr (SSIM needs the tensor) */,
                 int px_per_sample, float hinge, float cutoff) {
  __shared__ float s_wc[C * 4];
  __shared__ float s_red[4];
  if (threadIdx.x < C * 4) s_wc[threadIdx.x] = wc[threadIdx.x];
  __syncthreads();
  const int s = blockIdx.y;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < px_per_sample; i += gridDim.x * blockDim.x) {
    const size_t p = (size_t)s * px_per_sample + i;
    float x[C], th[3], pr[3];
    load_px16(feat + p * C, x);
    head_forward_px(x, s_wc, th, pr);
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      const LossTerms t = loss_terms(gt[p * 3 + o] - pr[o], hinge, cutoff);
      acc[0] += t.abs_e; acc[1] += t.hinged; acc[2] += t.sq; acc[3] += t.sqh;
      if (pred_out) pred_out[p * 3 + o] = pr[o];
    }
  }
  block_accumulate<4>(acc, sums + 4 * s, s_red);
}

// backward of loss + head: dX = Wc * dy ; G[ci][o] += x[ci]*dy[o]
// Four lanes per pixel, lane q of the quad owns channels 4q..4q+3: one coalesced 16-byte load and store per lane, 12 G
// accumulators (the one-thread-per-pixel version held 48 and a whole pixel, 153 registers: one CTA per SM and 126 us
// against an HBM floor of 45 us).  Lane o < 3 of a quad evaluates output channel o (tanh, loss derivative); the three dy
// travel through the quad by shuffle.
__global__ void __launch_bounds__(256)
head_backward_kernel(const float* __restrict__ feat, const float* __restrict__ gt, const float* __restrict__ wc,
                     const float* __restrict__ coef, const float* __restrict__ dpred_extra /* SSIM term or nullptr */,
                     float* __restrict__ dfeat, double* __restrict__ G /*[16][3]*/, int px_per_sample, float hinge, float cutoff) {
  __shared__ float s_wc[C * 4];
  __shared__ float s_red[C * 3];
  if (threadIdx.x < C * 4) s_wc[threadIdx.x] = wc[threadIdx.x];
  if (threadIdx.x < C * 3) s_red[threadIdx.x] = 0.f;
  __syncthreads();
  const int s = blockIdx.y;
  const int q = threadIdx.x & 3, lane = threadIdx.x & 31;
  const float c_mae = coef[0], c_mse = coef[1 + s];
  float w4[4][3];   // this lane's rows of Wc
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int o = 0; o < 3; ++o) w4[k][o] = s_wc[(4 * q + k) * 4 + o];
  float g[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int o = 0; o < 3; ++o) g[k][o] = 0.f;
  const int quads = (gridDim.x * blockDim.x) >> 2;
  // every lane of a warp runs the same number of iterations (shuffles inside): the bound is per quad-slot of the warp
  for (int i0 = (blockIdx.x * blockDim.x + (threadIdx.x & ~31)) >> 2; i0 < px_per_sample; i0 += quads) {
    const int i = i0 + (lane >> 2);
    const bool ok = i < px_per_sample;
    const size_t p = (size_t)s * px_per_sample + (ok ? i : 0);
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) x = *reinterpret_cast<const float4*>(feat + p * C + 4 * q);
    float y[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      float a = fmaf(x.x, w4[0][o], fmaf(x.y, w4[1][o], fmaf(x.z, w4[2][o], x.w * w4[3][o])));
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      y[o] = a;
    }
    float dyq = 0.f;
    if (ok && q < 3) {
      const float yo = q == 0 ? y[0] : (q == 1 ? y[1] : y[2]);
      const float th = tanhf(2.0f * yo);
      const float t = th * 0.51f;
      const float pr = (fminf(fmaxf(t, -0.5f), 0.5f) + 0.5f) * 255.0f;
      const float e = gt[p * 3 + q] - pr;
      const float ae = fabsf(e);
      const float sgn = (e > 0.f) ? 1.f : ((e < 0.f) ? -1.f : 0.f);
      float de = c_mae * sgn * keras_relu_grad(ae, hinge, cutoff);
      de += c_mse * keras_relu(e, hinge, cutoff * cutoff) * keras_relu_grad(e, hinge, cutoff * cutoff);
      const float dpred = -de + (dpred_extra ? dpred_extra[p * 3 + q] : 0.f);
      const float pass = (t >= -0.5f && t <= 0.5f) ? 1.f : 0.f;   // clip_by_value gradient
      dyq = dpred * 255.0f * pass * 0.51f * 2.0f * (1.0f - th * th);
    }
    float dy[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) dy[o] = __shfl_sync(0xffffffffu, dyq, (lane & ~3) + o);
    if (ok) {
      float4 dx;
      dx.x = w4[0][0] * dy[0] + w4[0][1] * dy[1] + w4[0][2] * dy[2];
      dx.y = w4[1][0] * dy[0] + w4[1][1] * dy[1] + w4[1][2] * dy[2];
      dx.z = w4[2][0] * dy[0] + w4[2][1] * dy[1] + w4[2][2] * dy[2];
      dx.w = w4[3][0] * dy[0] + w4[3][1] * dy[1] + w4[3][2] * dy[2];
      *reinterpret_cast<float4*>(dfeat + p * C + 4 * q) = dx;
      const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int o = 0; o < 3; ++o) g[k][o] = fmaf(xv[k], dy[o], g[k][o]);
    }
  }
  // lanes with equal q hold the same G entries: sum them inside the warp, then the CTA, then double atomics
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      float v = g[k][o];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 4) atomicAdd(&s_red[(4 * q + k) * 3 + o], v);
    }
  __syncthreads();
  if (threadIdx.x < C * 3) atomicAdd(&G[threadIdx.x], (double)s_red[threadIdx.x]);
}

// dH0 = G * H1^T ; dH1 = H0^T * G  (+ L2 regulariser gradient 2*0.01*lambda*w, model.py:275)
__global__ void head_grad_finalize_kernel(const double* __restrict__ G, const float* __restrict__ vars, long long h0_off,
                                          long long h1_off, int F, float reg2, float* __restrict__ g_h0,
                                          float* __restrict__ g_h1) {
  for (int i = threadIdx.x; i < C * F; i += blockDim.x) {
    const int ci = i / F, f = i % F;
    double a = 0;
    for (int o = 0; o < 3; ++o) a += G[ci * 3 + o] * (double)vars[h1_off + f * 3 + o];
    g_h0[i] = (float)(a + (double)reg2 * (double)vars[h0_off + i]);
  }
  for (int i = threadIdx.x; i < F * 3; i += blockDim.x) {
    const int f = i / 3, o = i % 3;
    double a = 0;
    for (int ci = 0; ci < C; ++ci) a += (double)vars[h0_off + ci * F + f] * G[ci * 3 + o];
    g_h1[i] = (float)(a + (double)reg2 * (double)vars[h1_off + i]);
  }
}

// -------------------------------------------------------------------------------------
// BN (training) backward: du = scale * (dy - mean(dy) - xhat * mean(dy*xhat)) ; dgamma = sum(dy*xhat)
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float4* __restrict__ dy, const float4* __restrict__ u, const float* __restrict__ bn,
                     double* __restrict__ sums /*[2][16]*/, long long n4) {
  __shared__ float s_bn[4 * C];
  __shared__ float s_red[2 * C];
  if (threadIdx.x < 4 * C) s_bn[threadIdx.x] = bn[threadIdx.x];
  __syncthreads();
  // a thread always sees the same channel quad: stride is a multiple of 4 float4
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = (int)(i0 & 3) * 4;
  for (long long i = i0; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 d = dy[i], uv = u[i];
    a[0] += d.x; a[1] += d.y; a[2] += d.z; a[3] += d.w;
    a[4] = fmaf(d.x, (uv.x - s_bn[c + 0]) * s_bn[2 * C + c + 0], a[4]);
    a[5] = fmaf(d.y, (uv.y - s_bn[c + 1]) * s_bn[2 * C + c + 1], a[5]);
    a[6] = fmaf(d.z, (uv.z - s_bn[c + 2]) * s_bn[2 * C + c + 2], a[6]);
    a[7] = fmaf(d.w, (uv.w - s_bn[c + 3]) * s_bn[2 * C + c + 3], a[7]);
  }
  if (threadIdx.x < 2 * C) s_red[threadIdx.x] = 0.f;
  __syncthreads();
  // lanes with equal (lane & 3) share the channel quad
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float v = a[k];
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if ((threadIdx.x & 31) < 4) atomicAdd(&s_red[(k >> 2) * C + c + (k & 3)], v);
  }
  __syncthreads();
  if (threadIdx.x < 2 * C) atomicAdd(&sums[threadIdx.x], (double)s_red[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float4* __restrict__ dy, const float4* __restrict__ u, const float* __restrict__ bn,
                    const double* __restrict__ sums, double cnt, float4* __restrict__ du, float* __restrict__ g_gamma,
                    long long n4) {
  __shared__ float s_bn[4 * C];
  __shared__ float s_m[2 * C];
  if (threadIdx.x < 4 * C) s_bn[threadIdx.x] = bn[threadIdx.x];
  if (threadIdx.x < 2 * C) s_m[threadIdx.x] = (float)(sums[threadIdx.x] / cnt);
  if (blockIdx.x == 0 && threadIdx.x < C) g_gamma[threadIdx.x] = (float)sums[C + threadIdx.x];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 3) * 4;
    const float4 d = dy[i], uv = u[i];
    float4 r;
    r.x = s_bn[3 * C + c + 0] * (d.x - s_m[c + 0] - (uv.x - s_bn[c + 0]) * s_bn[2 * C + c + 0] * s_m[C + c + 0]);
    r.y = s_bn[3 * C + c + 1] * (d.y - s_m[c + 1] - (uv.y - s_bn[c + 1]) * s_bn[2 * C + c + 1] * s_m[C + c + 1]);
    r.z = s_bn[3 * C + c + 2] * (d.z - s_m[c + 2] - (uv.z - s_bn[c + 2]) * s_bn[2 * C + c + 2] * s_m[C + c + 2]);
    r.w = s_bn[3 * C + c + 3] * (d.w - s_m[c + 3] - (uv.w - s_bn[c + 3]) * s_bn[2 * C + c + 3] * s_m[C + c + 3]);
    du[i] = r;
  }
}

// The same as per-channel coefficients of the fused conv prologue (conv_t5.cu): du = ca * dy + cb * u + cc, and dgamma
__global__ void bn_bwd_coef_kernel(const float* __restrict__ bn, const double* __restrict__ sums, double cnt,
                                   float* __restrict__ coef /*[3][16]*/, float* __restrict__ g_gamma) {
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");   // the next conv may run its setup now (conv_t5.cu, launch_pdl)
  const int c = threadIdx.x;
  if (c >= C) return;
  const double m1 = sums[c] / cnt, m2 = sums[C + c] / cnt;
  const double mean = bn[c], inv = bn[2 * C + c], scale = bn[3 * C + c];
  coef[c] = (float)scale;
  coef[C + c] = (float)(-scale * inv * m2);
  coef[2 * C + c] = (float)(scale * (mean * inv * m2 - m1));
  g_gamma[c] = (float)sums[C + c];
}

// -------------------------------------------------------------------------------------
// conv wgrad: dW[tap][ci][co] = sum_p A[p + tap][ci] * G[p][co]      (zero padding)
// CTA tile 64 x 8 pixels; 4 thread groups of 64, group g walks rows 2g, 2g+1 of the tile;
// thread (ci, q) of a group owns dW[0..8][ci][4q..4q+3] (36 accumulators) and slides a
// 3x3 window of A[.][ci] along x.  Persistent grid; per-CTA partial sums go to
// partial[cta][2304] and are summed in fixed order by wgrad_reduce_kernel (deterministic).
// -------------------------------------------------------------------------------------
constexpr int WG_W = 64, WG_H = 8;
constexpr int WG_A_FLOATS = (WG_H + 2) * (WG_W + 2) * C;
constexpr int WG_G_FLOATS = WG_H * WG_W * C;
constexpr size_t WG_SMEM = (size_t)(WG_A_FLOATS + WG_G_FLOATS) * sizeof(float);

__global__ void __launch_bounds__(256, 2)
wgrad3x3_kernel(const float* __restrict__ A, const float* __restrict__ G, float* __restrict__ partial,
                int n, int h, int w, int tiles_x, int tiles_y) {
  extern __shared__ __align__(16) float smem[];
  float* s_a = smem;                 // [WG_H+2][WG_W+2][16]
  float* s_g = smem + WG_A_FLOATS;   // [WG_H][WG_W][16]
  const int tid = threadIdx.x;
  const int grp = tid >> 6, t64 = tid & 63, ci = t64 >> 2, q = t64 & 3;
  float acc[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[t][k] = 0.f;

  const int ntiles = tiles_x * tiles_y * n;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int x0 = tx * WG_W, y0 = ty * WG_H;
    const float* a_b = A + (size_t)b * h * w * C;
    const float* g_b = G + (size_t)b * h * w * C;
    __syncthreads();
    for (int i = tid; i < (WG_H + 2) * (WG_W + 2) * 4; i += 256) {
      const int c4 = i & 3, p = i >> 2;
      const int lx = p % (WG_W + 2), ly = p / (WG_W + 2);
      const int gx = x0 + lx - 1, gy = y0 + ly - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gx >= 0 && gx < w && gy >= 0 && gy < h) v = *reinterpret_cast<const float4*>(a_b + ((size_t)gy * w + gx) * C + 4 * c4);
      reinterpret_cast<float4*>(s_a)[i] = v;
    }
    for (int i = tid; i < WG_H * WG_W * 4; i += 256) {
      const int c4 = i & 3, p = i >> 2;
      const int lx = p % WG_W, ly = p / WG_W;
      const int gx = x0 + lx, gy = y0 + ly;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gx < w && gy < h) v = *reinterpret_cast<const float4*>(g_b + ((size_t)gy * w + gx) * C + 4 * c4);
      reinterpret_cast<float4*>(s_g)[i] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int rr = 0; rr < 2; ++rr) {
      const int r = 2 * grp + rr;
      const float* a0 = s_a + ((r + 0) * (WG_W + 2)) * C + ci;
      const float* a1 = s_a + ((r + 1) * (WG_W + 2)) * C + ci;
      const float* a2 = s_a + ((r + 2) * (WG_W + 2)) * C + ci;
      const float* gp = s_g + (r * WG_W) * C + 4 * q;
      float w00 = a0[0], w01 = a0[C], w10 = a1[0], w11 = a1[C], w20 = a2[0], w21 = a2[C];
#pragma unroll 4
      for (int x = 0; x < WG_W; ++x) {
        const float w02 = a0[(x + 2) * C], w12 = a1[(x + 2) * C], w22 = a2[(x + 2) * C];
        const float4 gv = *reinterpret_cast<const float4*>(gp + x * C);
        const float wv[9] = {w00, w01, w02, w10, w11, w12, w20, w21, w22};
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          acc[t][0] = fmaf(wv[t], gv.x, acc[t][0]);
          acc[t][1] = fmaf(wv[t], gv.y, acc[t][1]);
          acc[t][2] = fmaf(wv[t], gv.z, acc[t][2]);
          acc[t][3] = fmaf(wv[t], gv.w, acc[t][3]);
        }
        w00 = w01; w01 = w02; w10 = w11; w11 = w12; w20 = w21; w21 = w22;
      }
    }
  }
  // reduce the 4 groups in fixed order through shared memory
  __syncthreads();
  float* s_red = smem;  // [4][2304]
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) s_red[grp * 2304 + (t * C + ci) * C + 4 * q + k] = acc[t][k];
  __syncthreads();
  float* dst = partial + (size_t)blockIdx.x * 2304;
  for (int i = tid; i < 2304; i += 256) dst[i] = (s_red[i] + s_red[2304 + i]) + (s_red[2 * 2304 + i] + s_red[3 * 2304 + i]);
}

// base conv wgrad: dWb[tap][c3][co] = sum_p xn[p + tap][c3] * G[p][co], xn = clip(x,0,255)/255 - 0.5 inside the
// image, 0 outside (zero padding of the NORMALISED tensor).  Thread = one (tap, c3), all 16 cout, every ngroups-th pixel.
constexpr int WB_W = 32, WB_H = 8;  // k0 <= 7: 2352 outputs <= the 4096 floats of the gradient tile (group sums reuse it)
__global__ void __launch_bounds__(256)
wgrad_base_kernel(const float* __restrict__ img, const float* __restrict__ G, float* __restrict__ partial,
                  int n, int h, int w, int k0, int tiles_x, int tiles_y) {
  extern __shared__ __align__(16) float smem[];
  const int r0 = (k0 - 1) >> 1;
  const int aw = WB_W + 2 * r0, ah = WB_H + 2 * r0;
  float* s_a = smem;                    // [ah][aw][3] normalised
  float* s_g = smem + ((ah * aw * 3 + 3) & ~3);  // [WB_H][WB_W][16]
  const int tid = threadIdx.x;
  const int nout = k0 * k0 * 3 * C;
  // thread -> ((tap, c3) = u, pixel group): ncombo = k0*k0*3 <= 147, ngroups = 256 / ncombo (k0 = 3: 27 combos x 9 groups)
  const int ncombo = k0 * k0 * 3, ngroups = 256 / ncombo;
  const bool active = tid < ngroups * ncombo;
  const int u = tid % ncombo, grp = tid / ncombo;
  const int c3 = u % 3, tap = u / 3, dy = tap / k0, dx = tap % k0;
  float acc[C];
#pragma unroll
  for (int k = 0; k < C; ++k) acc[k] = 0.f;
  const int ntiles = tiles_x * tiles_y * n;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int x0 = tx * WB_W, y0 = ty * WB_H;
    const float* i_b = img + (size_t)b * h * w * 3;
    const float* g_b = G + (size_t)b * h * w * C;
    __syncthreads();
    for (int i = tid; i < ah * aw * 3; i += 256) {
      const int c = i % 3, p = i / 3;
      const int lx = p % aw, ly = p / aw;
      const int gx = x0 + lx - r0, gy = y0 + ly - r0;
      float v = 0.f;
      if (gx >= 0 && gx < w && gy >= 0 && gy < h)
        v = __fsub_rn(__fdiv_rn(fminf(fmaxf(i_b[((size_t)gy * w + gx) * 3 + c], 0.f), 255.f), 255.f), 0.5f);
      s_a[i] = v;
    }
    for (int i = tid; i < WB_H * WB_W * 4; i += 256) {
      const int c4 = i & 3, p = i >> 2;
      const int lx = p % WB_W, ly = p / WB_W;
      const int gx = x0 + lx, gy = y0 + ly;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gx < w && gy < h) v = *reinterpret_cast<const float4*>(g_b + ((size_t)gy * w + gx) * C + 4 * c4);
      reinterpret_cast<float4*>(s_g)[i] = v;
    }
    __syncthreads();
    if (active) {
      // register tiling: one (tap, c3) and all 16 cout per thread, every `ngroups`-th pixel of the tile: 1 scalar + 4
      // broadcast vector loads per 16 FMAs (one output per thread needed 2 loads per FMA and was LDS-bound)
      for (int px = grp; px < WB_H * WB_W; px += ngroups) {
        const int y = px / WB_W, x = px - y * WB_W;
        const float av = s_a[((y + dy) * aw + x + dx) * 3 + c3];
        const float4* gp = reinterpret_cast<const float4*>(s_g + px * C);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 gv = gp[q];
          acc[4 * q] = fmaf(av, gv.x, acc[4 * q]); acc[4 * q + 1] = fmaf(av, gv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(av, gv.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(av, gv.w, acc[4 * q + 3]);
        }
      }
    }
  }
  // the pixel groups of the CTA are summed in fixed order (deterministic); s_g (256 x 16 floats) holds [ngroups][nout]
  __syncthreads();
  if (active) {
#pragma unroll
    for (int co = 0; co < C; ++co) s_g[grp * nout + u * C + co] = acc[co];
  }
  __syncthreads();
  float* dst = partial + (size_t)blockIdx.x * nout;
  for (int o = tid; o < nout; o += 256) {
    float a = 0.f;
    for (int gq = 0; gq < ngroups; ++gq) a += s_g[gq * nout + o];
    dst[o] = a;
  }
}

// out[i] = sum_cta partial[cta][i] (fixed order) + reg1 * sign(w[i])      (L1(0.01)*lambda, loss.py:181-187)
// 32 outputs x 8 part-lanes per CTA: warp j sums parts j, j+8, ... (coalesced 128-byte rows) in order, the 8 warp sums are
// combined in fixed order through shared memory, so the result is deterministic (one thread per output over ~300 parts
// was latency-bound: 25 us per launch, 26 launches per step).
constexpr int WR_OUT = 32, WR_LANES = 8;
__global__ void __launch_bounds__(WR_OUT * WR_LANES)
wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, int nout, const float* __restrict__ wts, float reg1,
                    float* __restrict__ out) {
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");   // the next conv may run its setup now (conv_t5.cu, launch_pdl)
  __shared__ double s_sum[WR_LANES][WR_OUT];
  const int j = threadIdx.x >> 5, li = threadIdx.x & 31;
  const int i = blockIdx.x * WR_OUT + li;
  double a = 0;
  if (i < nout)
    for (int p = j; p < nparts; p += WR_LANES) a += (double)partial[(size_t)p * nout + i];
  s_sum[j][li] = a;
  __syncthreads();
  if (j != 0 || i >= nout) return;
#pragma unroll
  for (int k = 1; k < WR_LANES; ++k) a += s_sum[k][li];
  const float wv = wts[i];
  const float sg = (wv > 0.f) ? 1.f : ((wv < 0.f) ? -1.f : 0.f);
  out[i] = (float)(a + (double)reg1 * (double)sg);
}

// dgamma with no regulariser: plain copy handled in bn_bwd_apply_kernel.

// regularisation value: 0.01*sum|w_backbone kernels| + 0.01*sum w_head^2   (regularizers "l1"/"l2")
__global__ void __launch_bounds__(256)
reg_loss_kernel(const float* __restrict__ vars, const long long* __restrict__ seg_off, const int* __restrict__ seg_len,
                const int* __restrict__ seg_kind, int nseg, double* __restrict__ out) {
  __shared__ double s_red[256];
  double a = 0;
  for (int s = 0; s < nseg; ++s) {
    const float* w = vars + seg_off[s];
    const int kind = seg_kind[s];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < seg_len[s]; i += gridDim.x * blockDim.x) {
      const double v = (double)w[i];
      a += (kind == 1) ? fabs(v) : v * v;
    }
  }
  s_red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) s_red[threadIdx.x] += s_red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, s_red[0] * 0.01);
}

// losses5 = total, denoiser total, mae, regularisation, ssim loss
__global__ void step_scalars_kernel(const float* __restrict__ loss_scal, const double* __restrict__ reg, float lambda,
                                    float* __restrict__ out5) {
  if (threadIdx.x != 0) return;
  out5[0] = (float)((double)loss_scal[0] + reg[0] * (double)lambda);
  out5[1] = loss_scal[0];
  out5[2] = loss_scal[1];
  out5[3] = (float)reg[0];
  out5[4] = loss_scal[4];
}

// =====================================================================================
// the training step
// =====================================================================================
struct TrainWs {
  // offsets into ws_stats (bytes)
  size_t bn_stats, bn_params, bn_coef, bwd_sums, loss_sums, ssim_sums, loss_scal, loss_coef, G, reg, out4, end;
};

static TrainWs plan_stats(int N, int n) {
  TrainWs w;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t r = o; o += (bytes + 255) & ~size_t(255); return r; };
  w.bn_stats = take((size_t)std::max(N, 1) * 2 * C * sizeof(double));
  w.bwd_sums = take((size_t)std::max(N, 1) * 2 * C * sizeof(double));
  w.loss_sums = take((size_t)4 * n * sizeof(double));
  w.ssim_sums = take((size_t)n * sizeof(double));
  w.G = take((size_t)C * 3 * sizeof(double));
  w.reg = take(sizeof(double));
  w.bn_params = take((size_t)std::max(N, 1) * 4 * C * sizeof(float));
  w.bn_coef = take((size_t)std::max(N, 1) * 6 * C * sizeof(float));   // per block: forward [3][16], backward [3][16]
  w.loss_scal = take(8 * sizeof(float));
  w.loss_coef = take((size_t)(1 + n) * sizeof(float));
  w.out4 = take(8 * sizeof(float));
  w.end = o;
  return w;
}

int run_train_step(bfcnn_handle* h, const float* clean, const float* noisy, int n, int height, int width,
                   const bfcnn_loss_cfg* cfg, float* flat_grads, float* losses4, int update_moving, cudaStream_t st) {
  const VarLayout& L = h->lay;
  const int N = L.N, k0 = L.k0, F = L.F;
  BF_REQUIRE(n <= 65535, "batch too large");
  BF_REQUIRE((long long)height * width * 3 < (1ll << 31), "image too large");
  const Extent e{n, height, width, height, width};
  const size_t npx = (size_t)n * height * width;
  const size_t map_floats = npx * C;
  const long long n4 = (long long)(map_floats / 4);
  const double cnt = (double)npx;
  static bool attr_set_dev[64] = {};   // function attributes are per device: one flag per device ordinal
  bool& attr_set = attr_set_dev[h->device & 63];
  if (!attr_set) {
    BF_CUDA(cudaFuncSetAttribute((const void*)wgrad3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
    BF_CUDA(cudaFuncSetAttribute((const void*)wgrad_base_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_set = true;
  }
  h->packed_valid = false;  // moving statistics (and, after adam, the weights) change on the device

  // ---- workspaces
  const int n_saved = 3 * N + 1;                       // X_0..X_N, T_i, U_i
  // behind them: the ReLU masks of the T_i as 16-bit maps (tcgen05 engine: the conv_b dgrad reads 2 instead of 64 B/pixel)
  const size_t mask_elems = (npx + 1) & ~(size_t)1;
  BF_CHECK(h->ws_train.reserve((size_t)n_saved * map_floats * sizeof(float) + (size_t)N * mask_elems * sizeof(uint16_t)));
  const bool use_ssim = cfg->ssim_multiplier > 0.f;
  // SSIM scratch behind the 4 gradient maps: pred [npx*3], dpred [npx*3], derivative maps [<= npx*9]
  BF_CHECK(h->ws_grads.reserve(((size_t)4 * map_floats + (use_ssim ? npx * 15 : 0)) * sizeof(float)));
  const int wg_grid = 2 * h->sm_count;
  const size_t nbase = (size_t)k0 * k0 * 3 * C;
  const size_t part_floats = (size_t)wg_grid * std::max<size_t>(2304, nbase);
  BF_CHECK(h->ws_feat[2].reserve((part_floats + (size_t)2 * std::max(N, 1) * 9 * C * C + C * 4) * sizeof(float)));
  const TrainWs W = plan_stats(N, n);
  BF_CHECK(h->ws_stats.reserve(W.end));
  uint8_t* sb = h->ws_stats.as<uint8_t>();
  double* bn_stats = reinterpret_cast<double*>(sb + W.bn_stats);
  double* bwd_sums = reinterpret_cast<double*>(sb + W.bwd_sums);
  double* loss_sums = reinterpret_cast<double*>(sb + W.loss_sums);
  double* ssim_sums = reinterpret_cast<double*>(sb + W.ssim_sums);
  double* Gd = reinterpret_cast<double*>(sb + W.G);
  double* regd = reinterpret_cast<double*>(sb + W.reg);
  float* bn_params = reinterpret_cast<float*>(sb + W.bn_params);
  float* bn_coef = reinterpret_cast<float*>(sb + W.bn_coef);
  float* loss_scal = reinterpret_cast<float*>(sb + W.loss_scal);
  float* loss_coef = reinterpret_cast<float*>(sb + W.loss_coef);
  float* out4_d = reinterpret_cast<float*>(sb + W.out4);
  BF_CUDA(cudaMemsetAsync(sb, 0, W.bn_params, st));   // all double accumulators

  float* saved = h->ws_train.as<float>();
  auto Xm = [&](int i) { return saved + (size_t)i * map_floats; };               // X_0..X_N
  auto Tm = [&](int i) { return saved + (size_t)(N + 1 + i) * map_floats; };     // T_0..T_{N-1}
  auto Um = [&](int i) { return saved + (size_t)(2 * N + 1 + i) * map_floats; }; // U_0..U_{N-1}
  auto Mm = [&](int i) { return reinterpret_cast<uint16_t*>(saved + (size_t)n_saved * map_floats) + (size_t)i * mask_elems; };
  float* gbuf = h->ws_grads.as<float>();
  float* dXa = gbuf; float* dXb = gbuf + map_floats; float* dU = gbuf + 2 * map_floats; float* dT = gbuf + 3 * map_floats;
  float* ss_pred = gbuf + 4 * map_floats; float* ss_dpred = ss_pred + npx * 3; float* ss_maps = ss_dpred + npx * 3;
  float* partial = h->ws_feat[2].as<float>();
  float* dgrad_w = partial + part_floats;                       // [2N][9][16][16]
  float* head_c = dgrad_w + (size_t)2 * std::max(N, 1) * 9 * C * C;  // [16][4]
  float* vars = h->d_vars.as<float>();

  // ---- tables: conv offsets (for the prep kernel) and regulariser segments; they depend on the architecture only and are
  // uploaded once per handle (a per-step upload needed a stream synchronisation: the host vectors die with their scope)
  if (!h->d_train_tables.p) {
    std::vector<long long> off(2 * N + 8, 0);
    std::vector<int> len(2 * N + 8, 0), kind(2 * N + 8, 0);
    int ns = 0;
    for (int i = 0; i < N; ++i) { off[ns] = (long long)L.wa[i]; len[ns] = 9 * C * C; kind[ns++] = 1;
                                  off[ns] = (long long)L.wb[i]; len[ns] = 9 * C * C; kind[ns++] = 1; }
    off[ns] = (long long)L.base; len[ns] = (int)nbase; kind[ns++] = 1;
    off[ns] = (long long)L.h0; len[ns] = C * F; kind[ns++] = 2;
    off[ns] = (long long)L.h1; len[ns] = F * 3; kind[ns++] = 2;
    const size_t nt = (size_t)(2 * N + 8);
    BF_CHECK(h->d_train_tables.reserve(nt * (sizeof(long long) + 2 * sizeof(int))));
    uint8_t* tb = h->d_train_tables.as<uint8_t>();
    BF_CUDA(cudaMemcpy(tb, off.data(), nt * sizeof(long long), cudaMemcpyHostToDevice));
    BF_CUDA(cudaMemcpy(tb + nt * sizeof(long long), len.data(), nt * sizeof(int), cudaMemcpyHostToDevice));
    BF_CUDA(cudaMemcpy(tb + nt * (sizeof(long long) + sizeof(int)), kind.data(), nt * sizeof(int), cudaMemcpyHostToDevice));
  }
  const long long* tab_off = h->d_train_tables.as<long long>();
  const int* tab_len = reinterpret_cast<const int*>(tab_off + (2 * N + 8));
  const int* tab_kind = tab_len + (2 * N + 8);
  const int nseg = 2 * N + 3;
  train_prep_kernel<<<2 * N + 1, 256, 0, st>>>(vars, tab_off, 2 * N, (long long)L.h0, (long long)L.h1, F, dgrad_w, head_c);
  reg_loss_kernel<<<8, 256, 0, st>>>(vars, tab_off, tab_len, tab_kind, nseg, regd);
  h->launches += 2;

  const int ew_blocks = (int)std::min<long long>((n4 + 255) / 256, (long long)h->sm_count * 8);
  const int ew_blocks4 = std::max(4, ew_blocks & ~3);  // multiple of 4 blocks*256 keeps (i & 3) fixed per thread

  // conv engine: tensor cores with the fp16 hi/lo split (FP32-grade, conv_x3.cu) unless BFCNN_TRAIN_CONV=fp32
  // conv engine: tensor cores with the fp16 hi/lo split (conv_x3.cu; default) or FP32 FFMA (bfcnn_set_train_engine)
  const int x3_mode = h->train_engine >= 1 ? 3 : 0;
  const bool t5 = h->train_engine == 2;   // the same arithmetic on tcgen05 (conv_t5.cu)
  // back-propagated gradients are ~255/(n*h*w*3) in magnitude: a power-of-two pre-scale brings them to O(1) for the split
  const float gscale = exp2f(floorf(log2f(fmaxf((float)(npx * 3) / 256.f, 1.f))));
  auto conv = [&](const float* in, float* out, const float* wts, const float* res, double* stats, ConvEpi epi) {
    const bool backward = (epi == CONV_MASK || epi == CONV_RESIDUAL);
    const bool use_x3 = (x3_mode & (backward ? 2 : 1)) != 0;
    if (use_x3 && t5) return launch_conv3x3_t5(h, in, out, wts, res, stats, epi, e, backward ? gscale * 64.0f : 64.0f, st);
    return use_x3 ? launch_conv3x3_x3(h, in, out, wts, res, stats, epi, e, backward ? gscale * 64.0f : 64.0f, st)
                  : launch_conv3x3_f32(h, in, out, wts, nullptr, res, stats, epi, e, st);
  };
  // ---- forward (training mode)
  // tcgen05 engine: BN + Add of block i-1 is the prologue of block i's conv_a, and the BN backward of block i the prologue of
  // its conv_b dgrad (conv_t5.cu, PRO = 1); the element-wise kernels remain for the last block and for the other engines
  const bool fuse = t5;
  if (k0 == 3 && x3_mode) BF_CHECK(launch_base_conv3_x3(h, noisy, Xm(0), vars + L.base, e, st));   // tensor cores (conv_x3.cu)
  else BF_CHECK(launch_base_conv(h, noisy, false, Xm(0), vars + L.base, e, st));
  for (int i = 0; i < N; ++i) {
    if (fuse && i > 0)
      BF_CHECK(launch_conv3x3_t5(h, Xm(i - 1), Tm(i), vars + L.wa[i], nullptr, nullptr, CONV_RELU, e, 64.0f, st, Um(i - 1), Xm(i),
                                 bn_coef + (size_t)(i - 1) * 6 * C, Mm(i)));
    else if (fuse)
      BF_CHECK(launch_conv3x3_t5(h, Xm(i), Tm(i), vars + L.wa[i], nullptr, nullptr, CONV_RELU, e, 64.0f, st, nullptr, nullptr, nullptr, Mm(i)));
    else
      BF_CHECK(conv(Xm(i), Tm(i), vars + L.wa[i], nullptr, nullptr, CONV_RELU));
    BF_CHECK(conv(Tm(i), Um(i), vars + L.wb[i], nullptr, bn_stats + (size_t)i * 2 * C, CONV_STATS));
    bn_finalize_kernel<<<1, 32, 0, st>>>(bn_stats + (size_t)i * 2 * C, cnt, h->arch.bn_epsilon, h->arch.bn_momentum, vars,
                                         (long long)L.gamma[i], (long long)L.mean[i], (long long)L.var[i],
                                         bn_params + (size_t)i * 4 * C, bn_coef + (size_t)i * 6 * C, update_moving);
    h->launches++;
    if (!fuse || i == N - 1) {
      bn_residual_kernel<<<ew_blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(Xm(i)), reinterpret_cast<const float4*>(Um(i)),
                                                    reinterpret_cast<float4*>(Xm(i + 1)), bn_params + (size_t)i * 4 * C, n4);
      h->launches++;
    }
  }
  // ---- loss
  const int px_per_sample = height * width;
  dim3 lgrid((unsigned)loss_grid_x(h, px_per_sample, n), (unsigned)n);
  head_loss_kernel<<<lgrid, 256, 0, st>>>(Xm(N), clean, head_c, loss_sums, use_ssim ? ss_pred : nullptr, px_per_sample, cfg->hinge,
                                          cfg->cutoff);
  double ssim_cnt = 1;
  if (use_ssim) {
    BF_CHECK(run_ssim(h, clean, ss_pred, n, height, width, ss_maps, ssim_sums, st));
    const int hv = height - SS_F + 1, wv = width - SS_F + 1;
    ssim_cnt = 3.0 * hv * wv;
    // d(total)/d(pred) of  m * (1 - mean_b mean_{c,w} S)
    const float coef = -cfg->ssim_multiplier / (float)((double)n * ssim_cnt);
    dim3 bgrid((width + SS_TW - 1) / SS_TW, (height + SS_TH - 1) / SS_TH, n);
    ssim_backward_kernel<<<bgrid, SS_TW * SS_TH, 0, st>>>(clean, ss_pred, ss_maps, ss_dpred, height, width, hv, wv, coef, ssim_taps());
    h->launches++;
  }
  loss_finalize_kernel<<<1, 32, 0, st>>>(loss_sums, use_ssim ? ssim_sums : nullptr, ssim_cnt, n, px_per_sample * 3, *cfg, loss_scal,
                                         loss_coef);
  step_scalars_kernel<<<1, 32, 0, st>>>(loss_scal, regd, cfg->regularization, out4_d);
  // ---- backward: head
  const dim3 bgrid4((unsigned)std::max(1, std::min((px_per_sample * 4 + 255) / 256, (16 * h->sm_count + n - 1) / n)), (unsigned)n);
  head_backward_kernel<<<bgrid4, 256, 0, st>>>(Xm(N), clean, head_c, loss_coef, use_ssim ? ss_dpred : nullptr, dXa, Gd, px_per_sample,
                                              cfg->hinge, cfg->cutoff);
  const float reg1 = cfg->regularization * 0.01f;          // d/dw lambda*0.01*|w|
  const float reg2 = cfg->regularization * 0.01f * 2.0f;   // d/dw lambda*0.01*w^2
  head_grad_finalize_kernel<<<1, 256, 0, st>>>(Gd, vars, (long long)L.h0, (long long)L.h1, F, reg2, flat_grads + L.t_h0,
                                               flat_grads + L.t_h1);
  h->launches += 5;
  // ---- backward: residual blocks
  const int tiles_x = (width + WG_W - 1) / WG_W, tiles_y = (height + WG_H - 1) / WG_H;
  const int wg_blocks = std::min(wg_grid, tiles_x * tiles_y * n);
  // dW = act (x) grad + L1 sub-gradient, engine as for the convs
  auto wgrad = [&](const float* act, const float* grad, const float* wts, float* out) -> int {
    int parts = wg_blocks;
    if (x3_mode) {
      BF_CHECK(launch_wgrad3x3_x3(h, act, grad, partial, wg_grid, e, gscale * 64.0f, &parts, st));
    } else {
      wgrad3x3_kernel<<<wg_blocks, 256, WG_SMEM, st>>>(act, grad, partial, n, height, width, tiles_x, tiles_y);
      h->launches++;
    }
    wgrad_reduce_kernel<<<2304 / WR_OUT, WR_OUT * WR_LANES, 0, st>>>(partial, parts, 2304, wts, reg1, out);
    h->launches++;
    return BFCNN_OK;
  };
  float* dX = dXa; float* dXn = dXb;
  for (int i = N - 1; i >= 0; --i) {
    const float* bnp = bn_params + (size_t)i * 4 * C;
    double* bs = bwd_sums + (size_t)i * 2 * C;
    bn_bwd_reduce_kernel<<<ew_blocks4, 256, 0, st>>>(reinterpret_cast<const float4*>(dX), reinterpret_cast<const float4*>(Um(i)), bnp, bs, n4);
    // conv_b: dU = BN backward of dX ; dT = dgrad(dU) masked by ReLU ; dWb = T (x) dU
    if (fuse) {
      float* cf = bn_coef + (size_t)i * 6 * C + 3 * C;
      bn_bwd_coef_kernel<<<1, 32, 0, st>>>(bnp, bs, cnt, cf, flat_grads + L.t_gamma[i]);
      BF_CHECK(launch_conv3x3_t5(h, dX, dT, dgrad_w + (size_t)(2 * i + 1) * 9 * C * C, Tm(i), nullptr, CONV_MASK, e, gscale * 64.0f, st,
                                 Um(i), dU, cf, Mm(i)));
    } else {
      bn_bwd_apply_kernel<<<ew_blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(dX), reinterpret_cast<const float4*>(Um(i)), bnp, bs, cnt,
                                                     reinterpret_cast<float4*>(dU), flat_grads + L.t_gamma[i], n4);
      BF_CHECK(conv(dU, dT, dgrad_w + (size_t)(2 * i + 1) * 9 * C * C, Tm(i), nullptr, CONV_MASK));
    }
    BF_CHECK(wgrad(Tm(i), dU, vars + L.wb[i], flat_grads + L.t_wb[i]));
    // conv_a: dX_i = dgrad(dT) + dX_{i+1} ; dWa = X_i (x) dT
    BF_CHECK(conv(dT, dXn, dgrad_w + (size_t)(2 * i) * 9 * C * C, dX, nullptr, CONV_RESIDUAL));
    BF_CHECK(wgrad(Xm(i), dT, vars + L.wa[i], flat_grads + L.t_wa[i]));
    h->launches += 2;
    std::swap(dX, dXn);
  }
  // ---- backward: base conv (weights only; no gradient into the image)
  {
    int blocks = 0;
    if (k0 == 3 && x3_mode) {
      // tensor cores, the arithmetic of the 3x3 wgrads (conv_x3.cu)
      BF_CHECK(launch_wgrad_base3_x3(h, noisy, dX, partial, (int)(part_floats / nbase), e, gscale * 64.0f, &blocks, st));
    } else {
      const int r0 = (k0 - 1) / 2;
      const int btx = (width + WB_W - 1) / WB_W, bty = (height + WB_H - 1) / WB_H;
      // latency-bound tile loads: up to 6 CTAs per SM, as many as the partial-sum buffer holds rows for
      blocks = (int)std::min<size_t>(std::min<size_t>((size_t)btx * bty * n, (size_t)6 * h->sm_count), part_floats / nbase);
      const size_t smem = (size_t)((((WB_H + 2 * r0) * (WB_W + 2 * r0) * 3 + 3) & ~3) + WB_H * WB_W * C) * sizeof(float);
      wgrad_base_kernel<<<blocks, 256, smem, st>>>(noisy, dX, partial, n, height, width, k0, btx, bty);
      h->launches++;
    }
    wgrad_reduce_kernel<<<(unsigned)((nbase + WR_OUT - 1) / WR_OUT), WR_OUT * WR_LANES, 0, st>>>(partial, blocks, (int)nbase, vars + L.base, reg1,
                                                                         flat_grads + L.t_base);
    h->launches++;
  }
  BF_CUDA(cudaGetLastError());
  h->tr_n = n; h->tr_h = height; h->tr_w = width;
  h->tr_out5_off = W.out4;
  // losses4 == nullptr: the step stays asynchronous (no device-to-host copy, no synchronisation); the scalars stay on the
  // device until bfcnn_train_losses() fetches them, so steps chain without a host round trip
  if (losses4) {
    BF_CUDA(cudaMemcpyAsync(losses4, out4_d, 5 * sizeof(float), cudaMemcpyDeviceToHost, st));
    BF_CUDA(cudaStreamSynchronize(st));
  }
  return BFCNN_OK;
}

int run_train_losses(bfcnn_handle* h, float* losses5, cudaStream_t st) {
  BF_REQUIRE(h->tr_n > 0 && h->ws_stats.p != nullptr, "no training step has run on this handle");
  BF_CUDA(cudaMemcpyAsync(losses5, h->ws_stats.as<uint8_t>() + h->tr_out5_off, 5 * sizeof(float), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  return BFCNN_OK;
}

// which: 0 = X_i (input of block i; i = N is the stack output), 1 = T_i = ReLU(conv_a), 2 = U_i = conv_b output before BN
int run_saved_activation(bfcnn_handle* h, int which, int index, float* out, cudaStream_t st) {
  const int N = h->lay.N;
  BF_REQUIRE(h->tr_n > 0 && h->ws_train.p != nullptr, "no training step has run on this handle");
  BF_REQUIRE(which >= 0 && which <= 2, "which must be 0 (X), 1 (T) or 2 (U)");
  BF_REQUIRE(index >= 0 && index < (which == 0 ? N + 1 : N), "activation index out of range");
  const size_t map_floats = (size_t)h->tr_n * h->tr_h * h->tr_w * C;
  const size_t slot = which == 0 ? (size_t)index : (which == 1 ? (size_t)(N + 1 + index) : (size_t)(2 * N + 1 + index));
  BF_CUDA(cudaMemcpyAsync(out, h->ws_train.as<float>() + slot * map_floats, map_floats * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return BFCNN_OK;
}

// =====================================================================================
// Adam with global-norm clipping over the flat trainable vector (optimizer.py:145-224)
// =====================================================================================
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, int n, float scale, double* __restrict__ out) {
  __shared__ double s_red[256];
  double a = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double v = (double)g[i] * (double)scale;
    a += v * v;
  }
  s_red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) s_red[threadIdx.x] += s_red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, s_red[0]);
}

// keras Adam: m += (g-m)(1-b1); v += (g^2-v)(1-b2); w -= m*alpha/(sqrt(v)+eps),
// alpha = lr*sqrt(1-b2^t)/(1-b1^t); g pre-scaled by grad_scale and by clip/max(norm, clip)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ vars, const long long* __restrict__ map /*[n] trainable -> variable index*/,
            const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int n, float grad_scale,
            const double* __restrict__ sumsq, bfcnn_adam_cfg cfg, float alpha) {
  float clip = 1.f;
  if (cfg.global_clipnorm > 0.f) {
    const float norm = (float)sqrt(*sumsq);
    clip = cfg.global_clipnorm / fmaxf(norm, cfg.global_clipnorm);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale * clip;
    const float mi = m[i] + (gi - m[i]) * (1.f - cfg.beta_1);
    const float vi = v[i] + (gi * gi - v[i]) * (1.f - cfg.beta_2);
    m[i] = mi; v[i] = vi;
    const long long j = map[i];
    vars[j] = vars[j] - mi * alpha / (sqrtf(vi) + cfg.epsilon);
  }
}

int run_adam_step(bfcnn_handle* h, const float* flat_grads, float grad_scale, const bfcnn_adam_cfg* cfg, int64_t step,
                  cudaStream_t st) {
  const VarLayout& L = h->lay;
  const int nt = (int)L.t_total;
  const bool fresh = (h->adam_m.p == nullptr);
  BF_CHECK(h->adam_m.reserve((size_t)nt * sizeof(float)));
  BF_CHECK(h->adam_v.reserve((size_t)nt * sizeof(float) + 16 + sizeof(double) + (size_t)nt * sizeof(long long)));
  float* v = h->adam_v.as<float>();
  double* sumsq = reinterpret_cast<double*>(h->adam_v.as<uint8_t>() + (((size_t)nt * sizeof(float) + 7) & ~size_t(7)));
  long long* map = reinterpret_cast<long long*>(sumsq + 1);
  if (fresh) {
    BF_CUDA(cudaMemsetAsync(h->adam_m.p, 0, (size_t)nt * sizeof(float), st));
    BF_CUDA(cudaMemsetAsync(v, 0, (size_t)nt * sizeof(float), st));
    std::vector<long long> hm((size_t)nt);
    auto fill = [&](size_t t_off, size_t v_off, size_t len) { for (size_t k = 0; k < len; ++k) hm[t_off + k] = (long long)(v_off + k); };
    fill(L.t_base, L.base, (size_t)L.k0 * L.k0 * 3 * C);
    for (int i = 0; i < L.N; ++i) { fill(L.t_wa[i], L.wa[i], 9 * C * C); fill(L.t_wb[i], L.wb[i], 9 * C * C); fill(L.t_gamma[i], L.gamma[i], C); }
    fill(L.t_h0, L.h0, (size_t)C * L.F);
    fill(L.t_h1, L.h1, (size_t)L.F * 3);
    BF_CUDA(cudaMemcpyAsync(map, hm.data(), hm.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
    BF_CUDA(cudaStreamSynchronize(st));
  }
  BF_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(double), st));
  if (cfg->global_clipnorm > 0.f) {
    sumsq_kernel<<<32, 256, 0, st>>>(flat_grads, nt, grad_scale, sumsq);
    h->launches++;
  }
  const double b1t = pow((double)cfg->beta_1, (double)step), b2t = pow((double)cfg->beta_2, (double)step);
  const float alpha = (float)((double)cfg->learning_rate * sqrt(1.0 - b2t) / (1.0 - b1t));
  adam_kernel<<<(nt + 255) / 256, 256, 0, st>>>(h->d_vars.as<float>(), map, flat_grads, h->adam_m.as<float>(), v, nt, grad_scale,
                                                sumsq, *cfg, alpha);
  h->launches++;
  h->packed_valid = false;
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

}  // namespace bfcnn
