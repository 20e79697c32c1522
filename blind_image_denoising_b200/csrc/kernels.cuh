// kernels.cuh -- launchers shared between the translation units of libbfcnn_b200.so
#pragma once
#include <algorithm>
#include <cuda.h>   // CUtensorMap (types only)
#include <cuda_fp16.h>
#include "common.cuh"

namespace bfcnn {

// model.py:342 tanh(2y)*0.51, then utilities.py:435-443 (clip(+-0.5)+0.5)*255
__device__ __forceinline__ float head_activation(float y) {
  float t = tanhf(2.0f * y) * 0.51f;
  t = fminf(fmaxf(t, -0.5f), 0.5f);
  return (t + 0.5f) * 255.0f;
}

// Programmatic dependent launch.  Host: the kernel's CTAs may take their SMs as the previous kernel's CTAs exit (or, when
// that kernel released its dependents early, while it still runs) and execute whatever precedes pdl_wait_for_previous():
// setup that reads nothing the previous kernels of the stream wrote.  A kernel launched the ordinary way in between is a
// full fence (train_prep / adam, the writers of the weights the conv setups read, are launched that way).
template <typename Kernel, typename... Args>
inline cudaError_t launch_pdl(Kernel kernel, int grid, int threads, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
// Device: tell the scheduler that the NEXT kernel's CTAs may be placed (they wait at this same point), then wait until the
// previous kernel has completed and its writes are visible.
__device__ __forceinline__ void pdl_wait_for_previous() {
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a thread moves a whole 32-byte sector, so an fp32 NHWC16 pixel is
// two full-sector accesses instead of four half-sector ones at a 64-byte stride (those were the critical path of the
// training conv's epilogue: 115 -> 72 us, and of the head kernels)
__device__ __forceinline__ void ldg256(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float4& a, const float4& b) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y),
               "f"(b.z), "f"(b.w) : "memory");
}

// ---- conv_f32.cu
int launch_base_conv(bfcnn_handle* h, const void* img, bool img_is_u8, float* out, const float* w,
                     const Extent& e, cudaStream_t st);
enum ConvEpi { CONV_PLAIN = 0, CONV_RELU = 1, CONV_RESIDUAL = 2, CONV_STATS = 3, CONV_MASK = 4 };
int launch_conv3x3_f32(bfcnn_handle* h, const float* in, float* out, const float* w, const float* bias,
                       const float* res, double* stats, ConvEpi epi, const Extent& e, cudaStream_t st);
int launch_head(bfcnn_handle* h, const float* feat, void* out, bool out_u8, const float* wh,
                const Extent& e, cudaStream_t st);

// ---- conv_x3.cu: the same layer contract on tensor cores, fp16 hi/lo split (FP32-grade); training fwd + dgrad
int launch_conv3x3_x3(bfcnn_handle* h, const float* in, float* out, const float* w, const float* res, double* stats,
                      ConvEpi epi, const Extent& e, float in_scale, cudaStream_t st);

// ---- conv_t5.cu: the same layer contract on tcgen05 (row-streaming, F16X3 arithmetic); default training conv engine
// in2 / out2 / coef: fused prologue (the conv input is ca*in + cb*in2 + cc per channel, written to out2), ReLU / mask epilogues
// relu_mask: a [n,h,w] uint16 map, bit c = (channel c of the ReLU output > 0): written by CONV_RELU, read by CONV_MASK instead
// of the 64 B/pixel activation `res`
int launch_conv3x3_t5(bfcnn_handle* h, const float* in, float* out, const float* w, const float* res, double* stats,
                      ConvEpi epi, const Extent& e, float in_scale, cudaStream_t st, const float* in2 = nullptr,
                      float* out2 = nullptr, const float* coef = nullptr, uint16_t* relu_mask = nullptr);

int launch_wgrad3x3_x3(bfcnn_handle* h, const float* act, const float* grad, float* partial, int max_parts, const Extent& e,
                       float g_scale, int* parts_out, cudaStream_t st);
// training forward: normalise + base conv k0 = 3, fp32 image [n,h,w,3] (0..255) -> fp32 NHWC16, on the same arithmetic
int launch_base_conv3_x3(bfcnn_handle* h, const float* img, float* out, const float* w, const Extent& e, cudaStream_t st);
// base-conv wgrad (k0 = 3) on the same arithmetic: img = the fp32 image [n,h,w,3] (0..255), partial [grid][432]
int launch_wgrad_base3_x3(bfcnn_handle* h, const float* img, const float* grad, float* partial, int max_parts, const Extent& e,
                          float g_scale, int* parts_out, cudaStream_t st);

// ---- base_conv.cu: normalise + base conv into the fp16 NHWC16 map of the streaming stacks, and its TMA descriptor
// fp16 NHWC16 feature map as a 5-D TMA tensor {ch8, half, x, y, n}, box = box_x pixels x box_y rows of one channel half
int make_feature_tmap(CUtensorMap* out, const __half* base, const Extent& e, int box_x, int box_y, long long row_px = 0);
// feat_lo != nullptr (F16X3 stack): also the lo part of the fp16 hi/lo split
// img_stride / row_stride (pixels; 0 = the plain [n][he][we] map): where pixel (b, y, x) goes, b * img_stride + y * row_stride + x
int launch_base_conv_f16(bfcnn_handle* h, const uint8_t* d_in, __half* feat, const Extent& e, cudaStream_t st,
                         __half* feat_lo = nullptr, long long img_stride = 0, long long row_stride = 0);
// ---- base_conv_t5.cu: the k0 = 3 base conv on tcgen05 into the virtual-row map [he][n (we + 1)][16] (feat_lo: F16X3)
int launch_base_conv3_t5(bfcnn_handle* h, const uint8_t* d_in, __half* feat, __half* feat_lo, const Extent& e, cudaStream_t st);
// ---- fused_stream.cu: the fused conv-BN-ReLU stack on tcgen05 as a row-streaming pipeline, F16 arithmetic
int run_fused_stack_stream(bfcnn_handle* h, const uint8_t* d_in, void* d_out, bool out_u8, const Extent& e,
                           cudaStream_t st);
// ---- fused_stream_x3.cu: the same pipeline in the F16X3 arithmetic (fp16 hi/lo operand parts, FP32-grade)
int run_fused_stack_stream_x3(bfcnn_handle* h, const uint8_t* d_in, void* d_out, bool out_u8, const Extent& e,
                              cudaStream_t st);

// ---- train.cu
int run_corrupt(bfcnn_handle* h, const uint8_t* clean_u8, float* clean_f32, float* noisy_f32, int n,
                int height, int width, uint64_t seed, uint64_t sample_offset,
                const bfcnn_noise_cfg* cfg, cudaStream_t st);
int run_loss(bfcnn_handle* h, const float* gt, const float* pred, int n, int height, int width,
             const bfcnn_loss_cfg* cfg, float* out4, cudaStream_t st);
int run_train_step(bfcnn_handle* h, const float* clean, const float* noisy, int n, int height,
                   int width, const bfcnn_loss_cfg* cfg, float* flat_grads, float* losses4,
                   int update_moving, cudaStream_t st);
int run_downscale2x(bfcnn_handle* h, const float* in, float* out, int n, int height, int width, int clip_values, int round_values,
                    cudaStream_t st);
int run_train_losses(bfcnn_handle* h, float* losses5, cudaStream_t st);
int run_saved_activation(bfcnn_handle* h, int which, int index, float* out, cudaStream_t st);
int run_adam_step(bfcnn_handle* h, const float* flat_grads, float grad_scale, const bfcnn_adam_cfg* cfg,
                  int64_t step, cudaStream_t st);

}  // namespace bfcnn
