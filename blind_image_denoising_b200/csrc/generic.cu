// generic.cu -- the reference-grade FP32 path for every OTHER `type: "resnet"` configuration the reference's builder accepts
// (SURVEY 8f N4): 1 to 3 convs per block with any odd kernel size, any filter counts, grouped and depthwise middle convs
// (backbone_resnet.py:149-178, the in-tree config resnet_color_1x6_bn_32x128x32_1x3x1_128x128_depthwise_l1_relu.json),
// initial / final BatchNormalization (:266-276), ChannelwiseMultiplier / Multiplier scalings (custom_layers.py:1028-1162,
// backbone_blocks.py:216-221).  One layer per launch, float32 NHWC, BN and multipliers folded into a per-channel
// (scale, bias) of the conv that precedes them by the host (generic.py).  The tcgen05 stacks stay specialised on the
// 16-channel two-conv family the north star names; this file makes the rest of the family load and run with the same
// semantics (pow2 canvas included, materialised literally as module_denoiser.py:56 does), not fast.
#include "kernels.cuh"

namespace bfcnn {

// uint8 [n,h,w,3] -> normalised float canvas [n,hc,wc,3]: clip(x,0,255)/255 - 0.5, raw zeros (= -0.5) bottom / right
// (utilities.py:449-461 after utilities.py:736-751)
__global__ void __launch_bounds__(256)
generic_prepare_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int n, int h, int w, int hc, int wc) {
  const long long total = (long long)n * hc * wc * 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    const long long p = i / 3;
    const int x = (int)(p % wc), y = (int)((p / wc) % hc), b = (int)(p / ((long long)wc * hc));
    float v = 0.f;
    if (y < h && x < w) v = (float)img[(((long long)b * h + y) * w + x) * 3 + c];
    out[i] = __fsub_rn(__fdiv_rn(v, 255.f), 0.5f);
  }
}

// One conv layer, "same" zero padding, stride 1, no bias (bias-free family):
//   depth_multiplier == 0: Conv2D(groups)           kernel [k,k,cin/groups,cout]   (Keras HWIO)
//   depth_multiplier  > 0: DepthwiseConv2D           kernel [k,k,cin,depth_multiplier], cout = cin * depth_multiplier,
//                          output channel ci * depth_multiplier + m  (tf.nn.depthwise_conv2d)
// then y = acc * scale[co] + bias[co] (folded BatchNormalization / multipliers; nullptr = identity), ReLU, + residual.
// One thread per (pixel, output channel): consecutive threads read consecutive weights and write consecutive outputs; the
// input pixel is a broadcast within the channel group.
__global__ void __launch_bounds__(256)
generic_conv_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ wts,
                    const float* __restrict__ scale, const float* __restrict__ bias, const float* __restrict__ res,
                    int n, int h, int w, int cin, int cout, int k, int groups, int dm, int relu) {
  const long long total = (long long)n * h * w * cout;
  const int r = (k - 1) >> 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    const long long p = i / cout;
    const int x = (int)(p % w), y = (int)((p / w) % h), b = (int)(p / ((long long)w * h));
    const float* in_b = in + (long long)b * h * w * cin;
    float acc = 0.f;
    if (dm > 0) {
      const int ci = co / dm, m = co - ci * dm;
      for (int dy = 0; dy < k; ++dy) {
        const int yy = y + dy - r;
        if (yy < 0 || yy >= h) continue;
        for (int dx = 0; dx < k; ++dx) {
          const int xx = x + dx - r;
          if (xx < 0 || xx >= w) continue;
          acc = fmaf(in_b[((long long)yy * w + xx) * cin + ci], wts[((long long)(dy * k + dx) * cin + ci) * dm + m], acc);
        }
      }
    } else {
      const int cin_g = cin / groups, cout_g = cout / groups, g = co / cout_g;
      for (int dy = 0; dy < k; ++dy) {
        const int yy = y + dy - r;
        if (yy < 0 || yy >= h) continue;
        for (int dx = 0; dx < k; ++dx) {
          const int xx = x + dx - r;
          if (xx < 0 || xx >= w) continue;
          const float* ip = in_b + ((long long)yy * w + xx) * cin + g * cin_g;
          const float* wp = wts + (long long)(dy * k + dx) * cin_g * cout + co;
          for (int c = 0; c < cin_g; ++c) acc = fmaf(ip[c], wp[(long long)c * cout], acc);
        }
      }
    }
    float v = acc;
    if (scale) v *= scale[co];
    if (bias) v += bias[co];
    if (relu) v = fmaxf(v, 0.f);
    if (res) v += res[i];
    out[i] = v;
  }
}

// head output y [n,hc,wc,3] (after the two 1x1 convs) -> tanh(2y)*0.51 -> (clip(+-0.5)+0.5)*255 -> crop [n,h,w,3] ->
// float32 or round-half-even uint8 (model.py:342, utilities.py:435-443, module_denoiser.py:68-73)
__global__ void __launch_bounds__(256)
generic_finish_kernel(const float* __restrict__ y, void* __restrict__ out, int n, int h, int w, int hc, int wc, int out_u8) {
  const long long total = (long long)n * h * w * 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    const long long p = i / 3;
    const int x = (int)(p % w), yy = (int)((p / w) % h), b = (int)(p / ((long long)w * h));
    const float v = head_activation(y[(((long long)b * hc + yy) * wc + x) * 3 + c]);
    if (out_u8) reinterpret_cast<uint8_t*>(out)[i] = (uint8_t)__float2int_rn(v);
    else reinterpret_cast<float*>(out)[i] = v;
  }
}

static int generic_grid(long long total) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)sms * 16));
}

}  // namespace bfcnn

using namespace bfcnn;

extern "C" {

int bfcnn_generic_prepare(int device, const uint8_t* img, float* canvas, int n, int height, int width, int canvas_h,
                          int canvas_w, void* stream) {
  BF_REQUIRE(n >= 0 && height >= 0 && width >= 0 && canvas_h >= height && canvas_w >= width, "bad image / canvas dimensions");
  const long long total = (long long)n * canvas_h * canvas_w * 3;
  if (total == 0) return BFCNN_OK;
  BF_REQUIRE(img != nullptr && canvas != nullptr, "NULL pointer");
  BF_CUDA(cudaSetDevice(device));
  generic_prepare_kernel<<<generic_grid(total), 256, 0, (cudaStream_t)stream>>>(img, canvas, n, height, width, canvas_h, canvas_w);
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

int bfcnn_generic_conv2d(int device, const float* in, float* out, const float* weights, const float* scale, const float* bias,
                         const float* residual, int n, int height, int width, int cin, int cout, int kernel, int groups,
                         int depth_multiplier, int relu, void* stream) {
  BF_REQUIRE(n >= 0 && height >= 0 && width >= 0 && cin >= 1 && cout >= 1, "bad tensor dimensions");
  BF_REQUIRE(kernel >= 1 && kernel <= 15 && (kernel & 1) == 1, "kernel size must be odd and <= 15");
  if (depth_multiplier > 0) BF_REQUIRE(cout == cin * depth_multiplier, "depthwise: cout must equal cin * depth_multiplier");
  else BF_REQUIRE(groups >= 1 && cin % groups == 0 && cout % groups == 0, "groups must divide cin and cout");
  const long long total = (long long)n * height * width * cout;
  if (total == 0) return BFCNN_OK;
  BF_REQUIRE(in != nullptr && out != nullptr && weights != nullptr && in != out, "NULL or aliased pointer");
  BF_CUDA(cudaSetDevice(device));
  generic_conv_kernel<<<generic_grid(total), 256, 0, (cudaStream_t)stream>>>(in, out, weights, scale, bias, residual, n, height, width,
                                                                            cin, cout, kernel, groups, depth_multiplier, relu);
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

int bfcnn_generic_finish(int device, const float* y, void* out, int n, int height, int width, int canvas_h, int canvas_w,
                         int out_u8, void* stream) {
  BF_REQUIRE(n >= 0 && height >= 0 && width >= 0 && canvas_h >= height && canvas_w >= width, "bad image / canvas dimensions");
  const long long total = (long long)n * height * width * 3;
  if (total == 0) return BFCNN_OK;
  BF_REQUIRE(y != nullptr && out != nullptr, "NULL pointer");
  BF_CUDA(cudaSetDevice(device));
  generic_finish_kernel<<<generic_grid(total), 256, 0, (cudaStream_t)stream>>>(y, out, n, height, width, canvas_h, canvas_w, out_u8);
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

}  // extern "C"
