// host_pack.cu -- variable layout, BN folding, head collapse and packing into kernel layouts.
//
// Folding: Keras BatchNormalization(center=False) still subtracts the moving mean,
// so conv_b + BN(inference) == conv with w' = w*gamma/sqrt(var+eps) plus the per-channel
// constant b' = -mean*gamma/sqrt(var+eps)   (backbone_resnet.py:129-135, constants.py:9).
// The two linear 1x1 head convs (model.py:297-340) multiply into one [16,3] matrix.
// All folding is done in double and rounded once.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace bfcnn {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int DevBuf::reserve(size_t n) {
  if (n <= bytes) return BFCNN_OK;
  release();
  const size_t want = (n + 255) & ~size_t(255);
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    p = nullptr;
    set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? BFCNN_ERR_OUT_OF_MEMORY : BFCNN_ERR_CUDA;
  }
  bytes = want;
  return BFCNN_OK;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}

void VarLayout::build(const bfcnn_arch& a) {
  N = a.no_layers; k0 = a.base_kernel; F = a.head_filters;
  wa.assign(N, 0); wb.assign(N, 0); gamma.assign(N, 0); mean.assign(N, 0); var.assign(N, 0);
  t_wa.assign(N, 0); t_wb.assign(N, 0); t_gamma.assign(N, 0);
  size_t o = 0, t = 0;
  base = o; t_base = t; o += (size_t)k0 * k0 * 3 * C; t += (size_t)k0 * k0 * 3 * C;
  for (int i = 0; i < N; ++i) {
    wa[i] = o; t_wa[i] = t; o += 9 * C * C; t += 9 * C * C;
    wb[i] = o; t_wb[i] = t; o += 9 * C * C; t += 9 * C * C;
    gamma[i] = o; t_gamma[i] = t; o += C; t += C;
    mean[i] = o; o += C;
    var[i] = o; o += C;
  }
  h0 = o; t_h0 = t; o += (size_t)C * F; t += (size_t)C * F;
  h1 = o; t_h1 = t; o += (size_t)F * 3; t += (size_t)F * 3;
  total = o; t_total = t;
}

int pack_weights(bfcnn_handle* h) {
  const VarLayout& L = h->lay;
  const float* v = h->h_vars.data();
  const int N = L.N, k0 = L.k0, F = L.F;
  const double eps = (double)h->arch.bn_epsilon;

  std::vector<float> conv((size_t)2 * N * 9 * C * C), bias((size_t)2 * N * C, 0.f), head(C * 4, 0.f);
  for (int i = 0; i < N; ++i) {
    memcpy(&conv[(size_t)(2 * i) * 9 * C * C], v + L.wa[i], sizeof(float) * 9 * C * C);
    double s[C];
    for (int c = 0; c < C; ++c) {
      s[c] = (double)v[L.gamma[i] + c] / sqrt((double)v[L.var[i] + c] + eps);
      bias[(size_t)(2 * i + 1) * C + c] = (float)(-(double)v[L.mean[i] + c] * s[c]);
    }
    float* dst = &conv[(size_t)(2 * i + 1) * 9 * C * C];
    for (int k = 0; k < 9 * C; ++k)
      for (int c = 0; c < C; ++c) dst[k * C + c] = (float)((double)v[L.wb[i] + k * C + c] * s[c]);
  }
  for (int ci = 0; ci < C; ++ci)
    for (int o = 0; o < 3; ++o) {
      double a = 0.0;
      for (int f = 0; f < F; ++f) a += (double)v[L.h0 + ci * F + f] * (double)v[L.h1 + f * 3 + o];
      head[ci * 4 + o] = (float)a;
    }

  // tcgen05 B operands (fused_stream.cu): per conv, per dx, N = 48 rows n = j*16 + cout with j <-> dy = 1 - j
  // (input row q feeds output rows q-1, q, q+1), K = 16 cin, SWIZZLE_NONE K-major core matrices:
  // byte offset(n, k) = (k/8)*768 + (n/8)*128 + (n%8)*16 + (k%8)*2
  std::vector<__half> umma((size_t)2 * N * 3 * 48 * 16);
  for (int l = 0; l < 2 * N; ++l)
    for (int dxi = 0; dxi < 3; ++dxi)
      for (int n = 0; n < 48; ++n)
        for (int k = 0; k < 16; ++k) {
          const int j = n / 16, co = n % 16, dy = 1 - j;
          const int tap = (dy + 1) * 3 + dxi;
          const size_t off = (size_t)(l * 3 + dxi) * 768 + (k / 8) * 384 + (n / 8) * 64 + (n % 8) * 8 + (k % 8);
          umma[off] = __float2half_rn(conv[((size_t)l * 9 + tap) * C * C + k * C + co]);
        }

  // hi/lo variant for the F16X3 arithmetic on tcgen05: per conv [hi][lo], each as above
  std::vector<__half> umma3((size_t)2 * N * 2 * 3 * 48 * 16);
  for (int l = 0; l < 2 * N; ++l)
    for (int dxi = 0; dxi < 3; ++dxi)
      for (int n = 0; n < 48; ++n)
        for (int k = 0; k < 16; ++k) {
          const int j = n / 16, co = n % 16, dy = 1 - j;
          const int tap = (dy + 1) * 3 + dxi;
          const size_t off = (size_t)dxi * 768 + (k / 8) * 384 + (n / 8) * 64 + (n % 8) * 8 + (k % 8);
          const float v0 = conv[((size_t)l * 9 + tap) * C * C + k * C + co];
          const __half hi = __float2half_rn(v0);
          umma3[(size_t)(l * 2 + 0) * 2304 + off] = hi;
          umma3[(size_t)(l * 2 + 1) * 2304 + off] = __float2half_rn(v0 - __half2float(hi));
        }

  // Last pass of the F16 stack: the collapsed 1x1 head [16][3] is folded INTO the last conv_b (its 16 output channels
  // become the 3 head outputs, columns 3..15 zero) and the residual reaches the head through one extra MMA per row
  // (A = the block's input row, B = the head matrix), so the accumulator of the last layer holds the pre-tanh head value
  // and the epilogue neither loads the residual nor multiplies 16 x 3 per pixel.  [dx 3][N 48][K 16] + [N 16][K 16].
  std::vector<__half> last((size_t)3 * 48 * 16 + 16 * 16, __float2half_rn(0.f));
  if (N > 0) {
    const int l = 2 * N - 1;
    for (int dxi = 0; dxi < 3; ++dxi)
      for (int n = 0; n < 48; ++n)
        for (int k = 0; k < 16; ++k) {
          const int j = n / 16, o = n % 16, dy = 1 - j;
          const int tap = (dy + 1) * 3 + dxi;
          double a = 0.0;
          if (o < 3)
            for (int co = 0; co < C; ++co) a += (double)conv[((size_t)l * 9 + tap) * C * C + k * C + co] * (double)head[co * 4 + o];
          last[(size_t)dxi * 768 + (k / 8) * 384 + (n / 8) * 64 + (n % 8) * 8 + (k % 8)] = __float2half_rn((float)a);
        }
    for (int n = 0; n < 3; ++n)
      for (int k = 0; k < 16; ++k) last[(size_t)2304 + (k / 8) * 128 + (n / 8) * 64 + (n % 8) * 8 + (k % 8)] = __float2half_rn(head[k * 4 + n]);
  }

  BF_CUDA(cudaSetDevice(h->device));
  const size_t nbase = (size_t)k0 * k0 * 3 * C;
  BF_CHECK(h->d_vars.reserve(L.total * sizeof(float)));
  BF_CHECK(h->d_base_f32.reserve(nbase * sizeof(float)));
  BF_CHECK(h->d_conv_f32.reserve(std::max<size_t>(conv.size(), 1) * sizeof(float)));
  BF_CHECK(h->d_bias_f32.reserve(std::max<size_t>(bias.size(), 1) * sizeof(float)));
  BF_CHECK(h->d_head_f32.reserve(head.size() * sizeof(float)));
  BF_CHECK(h->d_conv_umma.reserve(std::max<size_t>(umma.size(), 1) * sizeof(__half)));
  BF_CHECK(h->d_conv_umma_x3.reserve(std::max<size_t>(umma3.size(), 1) * sizeof(__half)));
  BF_CHECK(h->d_last_umma.reserve(last.size() * sizeof(__half)));
  BF_CUDA(cudaMemcpy(h->d_vars.p, v, L.total * sizeof(float), cudaMemcpyHostToDevice));
  BF_CUDA(cudaMemcpy(h->d_base_f32.p, v + L.base, nbase * sizeof(float), cudaMemcpyHostToDevice));
  if (N > 0) {
    BF_CUDA(cudaMemcpy(h->d_conv_f32.p, conv.data(), conv.size() * sizeof(float), cudaMemcpyHostToDevice));
    BF_CUDA(cudaMemcpy(h->d_bias_f32.p, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
    BF_CUDA(cudaMemcpy(h->d_conv_umma.p, umma.data(), umma.size() * sizeof(__half), cudaMemcpyHostToDevice));
    BF_CUDA(cudaMemcpy(h->d_conv_umma_x3.p, umma3.data(), umma3.size() * sizeof(__half), cudaMemcpyHostToDevice));
    BF_CUDA(cudaMemcpy(h->d_last_umma.p, last.data(), last.size() * sizeof(__half), cudaMemcpyHostToDevice));
  }
  BF_CUDA(cudaMemcpy(h->d_head_f32.p, head.data(), head.size() * sizeof(float), cudaMemcpyHostToDevice));
  h->packed_valid = true;
  return BFCNN_OK;
}

}  // namespace bfcnn
