// conv_f32.cu -- FP32 (FFMA) layer kernels: the reference-grade path and the training forward.
//
//   base_conv_kernel   : normalise (utilities.py:449-461) + base conv k0 x k0, 3->16
//                        (backbone_resnet.py:258-262)
//   conv3x3_c16_kernel : one 3x3 16->16 conv with fused epilogue (bias, ReLU, residual add,
//                        per-channel batch statistics)  (backbone_blocks.py:167-246)
//   head_kernel        : 1x1 16->3 (collapsed) + tanh(2y)*0.51 + denormalise + round
//                        (model.py:297-342, utilities.py:435-443, module_denoiser.py:71-73)
//
// Feature maps are NHWC float32 over the work extent [He, We] (common.cuh::Extent).
#include "kernels.cuh"

namespace bfcnn {

// ------------------------------------------------------------------------------------
// base conv
// ------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256)
base_conv_kernel(const TIn* __restrict__ img, float* __restrict__ out, const float* __restrict__ w,
                 int n, int h, int wd, int he, int we, int k0) {
  extern __shared__ float sw[];  // [k0*k0*3][16]
  const int nw = k0 * k0 * 3 * C;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int r = (k0 - 1) >> 1;
  const long long total = (long long)n * he * we;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % we);
    const int y = (int)((idx / we) % he);
    const int b = (int)(idx / ((long long)we * he));
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int dy = 0; dy < k0; ++dy) {
      const int yy = y + dy - r;
      if (yy < 0 || yy >= he) continue;  // zero padding of the NORMALISED tensor
      for (int dx = 0; dx < k0; ++dx) {
        const int xx = x + dx - r;
        if (xx < 0 || xx >= we) continue;
        float v[3];
        if (yy < h && xx < wd) {
          const TIn* p = img + (((long long)b * h + yy) * wd + xx) * 3;
          v[0] = (float)p[0]; v[1] = (float)p[1]; v[2] = (float)p[2];
        } else {
          v[0] = v[1] = v[2] = 0.f;  // raw zeros of the pow2 canvas (utilities.py:749)
        }
        const float* wt = sw + (dy * k0 + dx) * 3 * C;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          // layer_normalize: clip to [0,255], /255, -0.5
          const float xn = __fsub_rn(__fdiv_rn(fminf(fmaxf(v[ci], 0.f), 255.f), 255.f), 0.5f);
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] = fmaf(xn, wt[ci * C + c], acc[c]);
        }
      }
    }
    float4* o = reinterpret_cast<float4*>(out + idx * C);
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
  }
}

int launch_base_conv(bfcnn_handle* h, const void* img, bool img_is_u8, float* out, const float* w,
                     const Extent& e, cudaStream_t st) {
  const int k0 = h->arch.base_kernel;
  const size_t smem = (size_t)k0 * k0 * 3 * C * sizeof(float);
  const long long total = (long long)e.n * e.he * e.we;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16);
  if (img_is_u8)
    base_conv_kernel<uint8_t><<<blocks, 256, smem, st>>>((const uint8_t*)img, out, w, e.n, e.h, e.w, e.he, e.we, k0);
  else
    base_conv_kernel<float><<<blocks, 256, smem, st>>>((const float*)img, out, w, e.n, e.h, e.w, e.he, e.we, k0);
  h->launches++;
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

// ------------------------------------------------------------------------------------
// 3x3 16->16 conv, FP32
// ------------------------------------------------------------------------------------
// CTA tile: 64 (x) by 16 (y) outputs, 256 threads, each thread 4 consecutive x by 16 cout.
// Input tile in shared memory is channel-planar: plane c holds rows -1..16, columns -1..64
// at [row+1][col+4]; plane base = c*PLANE + 8*(c>>2) floats so that the NHWC->planar
// transposing stores of one warp (8 pixels x 16 channels) hit 32 distinct banks.
constexpr int CT_W = 64, CT_H = 16;
constexpr int CT_PITCH = 72;                       // floats per smem row (16 B aligned)
constexpr int CT_PLANE = (CT_H + 2) * CT_PITCH;    // 1296, multiple of 8
constexpr int CT_IN_FLOATS = C * CT_PLANE + 8 * 4;
constexpr int CT_W_FLOATS = 9 * C * C;
constexpr size_t CT_SMEM = (size_t)(CT_IN_FLOATS + CT_W_FLOATS + 2 * C) * sizeof(float);

__device__ __forceinline__ int ct_plane_base(int c) { return c * CT_PLANE + 8 * (c >> 2); }

template <bool RELU, bool RESIDUAL, bool STATS, bool MASK>
__global__ void __launch_bounds__(256, 2)
conv3x3_c16_kernel(const float* __restrict__ in, float* __restrict__ out,
                   const float* __restrict__ w,      // [9][16 cin][16 cout]
                   const float* __restrict__ bias,   // [16] or nullptr
                   const float* __restrict__ res,    // NHWC residual (RESIDUAL) / ReLU mask source (MASK)
                   double* __restrict__ stats,       // [2][16] sum, sumsq (STATS)
                   int he, int we) {
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;
  float* s_w = smem + CT_IN_FLOATS;
  float* s_stat = s_w + CT_W_FLOATS;  // [2][16]

  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H, b = blockIdx.z;
  const float* in_b = in + (long long)b * he * we * C;

  for (int i = tid; i < CT_W_FLOATS / 4; i += 256)
    reinterpret_cast<float4*>(s_w)[i] = reinterpret_cast<const float4*>(w)[i];
  if (STATS && tid < 2 * C) s_stat[tid] = 0.f;

  // ---- load the (CT_H+2) x (CT_W+2) input tile, zero outside the extent ("same" padding)
  // one float4 (4 channels of one pixel) per thread per iteration
  constexpr int LW = CT_W + 2;
  for (int i = tid; i < (CT_H + 2) * LW * 4; i += 256) {
    const int q = i & 3;
    const int p = i >> 2;
    const int lx = p % LW, ly = p / LW;
    const int gx = x0 + lx - 1, gy = y0 + ly - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gx >= 0 && gx < we && gy >= 0 && gy < he)
      v = *reinterpret_cast<const float4*>(in_b + ((long long)gy * we + gx) * C + 4 * q);
    const int o = ly * CT_PITCH + lx + 3;
    s_in[ct_plane_base(4 * q + 0) + o] = v.x;
    s_in[ct_plane_base(4 * q + 1) + o] = v.y;
    s_in[ct_plane_base(4 * q + 2) + o] = v.z;
    s_in[ct_plane_base(4 * q + 3) + o] = v.w;
  }
  __syncthreads();

  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][C];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < C; ++c) acc[p][c] = 0.f;

#pragma unroll 1
  for (int ci = 0; ci < C; ++ci) {
    const float* pl = s_in + ct_plane_base(ci) + ty * CT_PITCH + 4 * tx;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const float* row = pl + dy * CT_PITCH;
      float v[6];
      v[0] = row[3];
      const float4 m = *reinterpret_cast<const float4*>(row + 4);
      v[1] = m.x; v[2] = m.y; v[3] = m.z; v[4] = m.w;
      v[5] = row[8];
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float4* wp = reinterpret_cast<const float4*>(s_w + ((dy * 3 + dx) * C + ci) * C);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = wp[q];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            acc[p][4 * q + 0] = fmaf(v[p + dx], wv.x, acc[p][4 * q + 0]);
            acc[p][4 * q + 1] = fmaf(v[p + dx], wv.y, acc[p][4 * q + 1]);
            acc[p][4 * q + 2] = fmaf(v[p + dx], wv.z, acc[p][4 * q + 2]);
            acc[p][4 * q + 3] = fmaf(v[p + dx], wv.w, acc[p][4 * q + 3]);
          }
        }
      }
    }
  }

  // ---- epilogue
  const int gy = y0 + ty;
  float ssum[C], ssq[C];
  if (STATS) {
#pragma unroll
    for (int c = 0; c < C; ++c) ssum[c] = ssq[c] = 0.f;
  }
  if (gy < he) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int gx = x0 + 4 * tx + p;
      if (gx >= we) continue;
      const long long o = (((long long)b * he + gy) * we + gx) * C;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 r = make_float4(acc[p][4 * q], acc[p][4 * q + 1], acc[p][4 * q + 2], acc[p][4 * q + 3]);
        if (bias != nullptr) {
          const float4 bv = *reinterpret_cast<const float4*>(bias + 4 * q);
          r.x += bv.x; r.y += bv.y; r.z += bv.z; r.w += bv.w;
        }
        if (STATS) {
          ssum[4 * q] += r.x; ssum[4 * q + 1] += r.y; ssum[4 * q + 2] += r.z; ssum[4 * q + 3] += r.w;
          ssq[4 * q] += r.x * r.x; ssq[4 * q + 1] += r.y * r.y; ssq[4 * q + 2] += r.z * r.z; ssq[4 * q + 3] += r.w * r.w;
        }
        if (RELU) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
        if (RESIDUAL) {
          const float4 rv = *reinterpret_cast<const float4*>(res + o + 4 * q);
          r.x += rv.x; r.y += rv.y; r.z += rv.z; r.w += rv.w;
        }
        if (MASK) {  // ReLU backward: pass the gradient where the saved activation is > 0
          const float4 mv = *reinterpret_cast<const float4*>(res + o + 4 * q);
          r.x = mv.x > 0.f ? r.x : 0.f; r.y = mv.y > 0.f ? r.y : 0.f;
          r.z = mv.z > 0.f ? r.z : 0.f; r.w = mv.w > 0.f ? r.w : 0.f;
        }
        *reinterpret_cast<float4*>(out + o + 4 * q) = r;
      }
    }
  }
  if (STATS) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float a = ssum[c], q2 = ssq[c];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, s);
        q2 += __shfl_xor_sync(0xffffffffu, q2, s);
      }
      if ((tid & 31) == 0) { atomicAdd(&s_stat[c], a); atomicAdd(&s_stat[C + c], q2); }
    }
    __syncthreads();
    if (tid < 2 * C) atomicAdd(&stats[tid], (double)s_stat[tid]);
  }
}

int launch_conv3x3_f32(bfcnn_handle* h, const float* in, float* out, const float* w, const float* bias,
                       const float* res, double* stats, ConvEpi epi, const Extent& e, cudaStream_t st) {
  static bool attr_set_dev[64] = {};   // function attributes are per device: one flag per device ordinal
  bool& attr_set = attr_set_dev[h->device & 63];
  auto set_attr = [](const void* f) {
    return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_SMEM);
  };
  if (!attr_set) {
    BF_CUDA(set_attr((const void*)conv3x3_c16_kernel<true, false, false, false>));
    BF_CUDA(set_attr((const void*)conv3x3_c16_kernel<false, true, false, false>));
    BF_CUDA(set_attr((const void*)conv3x3_c16_kernel<false, false, false, false>));
    BF_CUDA(set_attr((const void*)conv3x3_c16_kernel<false, false, true, false>));
    BF_CUDA(set_attr((const void*)conv3x3_c16_kernel<false, false, false, true>));
    attr_set = true;
  }
  dim3 grid((e.we + CT_W - 1) / CT_W, (e.he + CT_H - 1) / CT_H, e.n);
  BF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "image too large for the FP32 conv grid");
  switch (epi) {
    case CONV_STATS:
      conv3x3_c16_kernel<false, false, true, false><<<grid, 256, CT_SMEM, st>>>(in, out, w, bias, nullptr, stats, e.he, e.we);
      break;
    case CONV_RELU:
      conv3x3_c16_kernel<true, false, false, false><<<grid, 256, CT_SMEM, st>>>(in, out, w, bias, nullptr, nullptr, e.he, e.we);
      break;
    case CONV_RESIDUAL:
      conv3x3_c16_kernel<false, true, false, false><<<grid, 256, CT_SMEM, st>>>(in, out, w, bias, res, nullptr, e.he, e.we);
      break;
    case CONV_MASK:
      conv3x3_c16_kernel<false, false, false, true><<<grid, 256, CT_SMEM, st>>>(in, out, w, bias, res, nullptr, e.he, e.we);
      break;
    case CONV_PLAIN:
      conv3x3_c16_kernel<false, false, false, false><<<grid, 256, CT_SMEM, st>>>(in, out, w, bias, nullptr, nullptr, e.he, e.we);
      break;
    default:
      set_error("unsupported conv epilogue");
      return BFCNN_ERR_INTERNAL;
  }
  h->launches++;
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

// ------------------------------------------------------------------------------------
// head: collapsed 1x1 16->3, tanh(2y)*0.51, denormalise, [round-half-even -> uint8]
// ------------------------------------------------------------------------------------
template <bool OUT_U8>
__global__ void __launch_bounds__(256)
head_kernel(const float* __restrict__ feat, void* __restrict__ out, const float* __restrict__ wh,  // [16][4]
            int n, int h, int wd, int he, int we) {
  __shared__ float sw[C * 4];
  if (threadIdx.x < C * 4) sw[threadIdx.x] = wh[threadIdx.x];
  __syncthreads();
  const long long total = (long long)n * h * wd;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % wd);
    const int y = (int)((idx / wd) % h);
    const int b = (int)(idx / ((long long)wd * h));
    const float4* f = reinterpret_cast<const float4*>(feat + (((long long)b * he + y) * we + x) * C);
    float y3[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = f[q];
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int o = 0; o < 3; ++o) y3[o] = fmaf(vv[i], sw[(4 * q + i) * 4 + o], y3[o]);
    }
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      const float r = head_activation(y3[o]);
      if (OUT_U8)
        reinterpret_cast<uint8_t*>(out)[idx * 3 + o] = (uint8_t)__float2int_rn(r);
      else
        reinterpret_cast<float*>(out)[idx * 3 + o] = r;
    }
  }
}

int launch_head(bfcnn_handle* h, const float* feat, void* out, bool out_u8, const float* wh,
                const Extent& e, cudaStream_t st) {
  const long long total = (long long)e.n * e.h * e.w;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16);
  if (out_u8)
    head_kernel<true><<<blocks, 256, 0, st>>>(feat, out, wh, e.n, e.h, e.w, e.he, e.we);
  else
    head_kernel<false><<<blocks, 256, 0, st>>>(feat, out, wh, e.n, e.h, e.w, e.he, e.we);
  h->launches++;
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

}  // namespace bfcnn
