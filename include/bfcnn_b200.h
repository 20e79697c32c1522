/*
 * bfcnn_b200.h -- C ABI of libbfcnn_b200.so: the B200 (sm_100a) hot path of the bfcnn
 * bias-free ResNet denoiser (resnet_color_1xN_bn_16x3x3) and of the training step
 * that feeds it.
 *
 * The reference (NikolasMarkou/blind_image_denoising, bfcnn 3.2.0) is pure
 * Python/TensorFlow and has NO FFI layer; its "operator interface" for
 * this path is a handful of Python callables.  Each entry point below names the
 * reference callable (file:line under /root/reference) whose arithmetic it replaces.
 * The Python package `blind_image_denoising_b200` (and its `bfcnn` alias) binds these
 * with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (BFCNN_OK) or a negative bfcnn_status; the message of
 *     the last failure on the calling thread is available from bfcnn_last_error();
 *   - nothing throws across the ABI and nothing ever falls back to the CPU: without a
 *     usable CUDA device every compute call fails with BFCNN_ERR_CUDA;
 *   - the caller owns every in/out buffer; the handle owns weights and workspaces;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream);
 *     calls on one handle are stream-ordered and not re-entrant; one handle per device;
 *   - images are NHWC uint8 / float32, C = 3; weights are ONE flat float32 vector in
 *     Keras `hydra.variables` order (see bfcnn_num_weights).
 */
#ifndef BFCNN_B200_H_
#define BFCNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFCNN_ABI_VERSION 3

typedef enum bfcnn_status {
  BFCNN_OK = 0,
  BFCNN_ERR_INVALID_ARGUMENT = -1,
  BFCNN_ERR_CUDA = -2,
  BFCNN_ERR_OUT_OF_MEMORY = -3,
  BFCNN_ERR_UNSUPPORTED = -4,
  BFCNN_ERR_INTERNAL = -5
} bfcnn_status;

/* Arithmetic the conv stack runs in (bfcnn_denoise_*). */
typedef enum bfcnn_precision {
  BFCNN_PREC_FP32 = 0,   /* FP32 FFMA, layer by layer: the reference-grade path            */
  BFCNN_PREC_F16 = 1,    /* fused stack on tcgen05 (UMMA, TMEM accumulators): fp16 operands, fp32 accumulate */
  BFCNN_PREC_F16X3 = 2   /* fused stack on tcgen05, fp16 hi/lo split of activations and weights (3 MMAs per product): fp32-grade */
} bfcnn_precision;

/* Flags of bfcnn_denoise_*. */
#define BFCNN_FLAG_IN_DEVICE   1u  /* `in` is a device pointer (else host)                        */
#define BFCNN_FLAG_OUT_DEVICE  2u  /* `out` is a device pointer (else host)                       */
#define BFCNN_FLAG_NO_PAD_POW2 4u  /* do NOT emulate pad_to_power_of_2 (utilities.py:736-751)     */
#define BFCNN_FLAG_IN_F32      8u  /* `in` holds float32 pixels (0..255 scale; the hydra model's own input,
                                      model.py:100-102) instead of uint8; BFCNN_PREC_FP32 only   */

/* Hyper-parameters of bfcnn/backbone_resnet.py:19-49 + bfcnn/model.py:267-275 for the
 * 16-channel, two-3x3-convs-per-block family. */
typedef struct bfcnn_arch {
  int32_t no_layers;     /* N residual blocks                      */
  int32_t base_kernel;   /* k0 of the base conv (odd, <= 7)        */
  int32_t filters;       /* 16                                     */
  int32_t head_filters;  /* F of the 1x1 head (model.py:268)       */
  int32_t in_channels;   /* 3                                      */
  int32_t out_channels;  /* 3                                      */
  float bn_epsilon;      /* constants.py:9  (1e-3)                 */
  float bn_momentum;     /* constants.py:11 (0.995)                */
} bfcnn_arch;

/* dataset.py:94-99,170-187 noise ranges + the sub-sampling corruption of README.md:49-55. */
typedef struct bfcnn_noise_cfg {
  float additive_min, additive_max;             /* sigma_add ~ U(min,max); max<=0 disables  */
  float multiplicative_min, multiplicative_max; /* sigma_mul ~ U(min,max); max<=0 disables  */
  int32_t random_left_right;                    /* dataset.py:141-148                       */
  int32_t random_up_down;                       /* dataset.py:150-158                       */
  int32_t subsample;                            /* 1: w.p. 1/2 decimate x2 + nearest up x2  */
  int32_t round_values;                         /* dataset.py:228 (the reference always rounds; 0 is for tests) */
  int32_t draw_group;                           /* who shares the call-level draws of dataset.py:141-187 (flips, on/off
                                                   switches, sigmas): 0/1 = every sample its own (no_crops_per_image = 1,
                                                   the reference default); k > 1 = runs of k consecutive global sample
                                                   indices, i.e. the k crops of one image (dataset.py:276-297); < 0 = the
                                                   whole call.  Per-pixel noise values are always per sample.       */
} bfcnn_noise_cfg;

/* loss.py:152-187 configuration. */
typedef struct bfcnn_loss_cfg {
  float hinge;            /* loss.py:165 */
  float cutoff;           /* loss.py:166 */
  float mae_multiplier;   /* loss.py:169 */
  float mse_multiplier;   /* loss.py:177 */
  float regularization;   /* loss.py:181 */
  float ssim_multiplier;  /* loss.py:171-173: (1 - mean tf.image.ssim(gt, pred, filter_size=7, max_val=255)) * m; row N3 */
} bfcnn_loss_cfg;

/* fused Adam of optimizer.py:145-224 (row N1). */
typedef struct bfcnn_adam_cfg {
  float learning_rate, beta_1, beta_2, epsilon;
  float global_clipnorm; /* <= 0 disables (optimizer.py:169) */
} bfcnn_adam_cfg;

typedef struct bfcnn_handle bfcnn_handle;

/* ---- library ---------------------------------------------------------------- */
int bfcnn_abi_version(void);
const char* bfcnn_last_error(void);
/* number of visible CUDA devices, or a negative status (never "0 and carry on"). */
int bfcnn_device_count(void);
/* floats in the flat Keras-order variable vector / in the trainable subset. */
int64_t bfcnn_num_weights(const bfcnn_arch* arch);
int64_t bfcnn_num_trainable(const bfcnn_arch* arch);

/* ---- model ------------------------------------------------------------------
 * replaces: bfcnn.load_model / tf.saved_model.load (bfcnn/__init__.py:81-97) and
 * model_builder (bfcnn/model.py:58-162) for this family.  `weights` (host) holds
 * bfcnn_num_weights(arch) floats: base kernel [k0,k0,3,16] HWIO, then per block
 * W_a, W_b [3,3,16,16], gamma, moving_mean, moving_var [16], then head [1,1,16,F],
 * [1,1,F,3].  BN is folded into W_b + a per-channel constant at load. */
int bfcnn_create(const bfcnn_arch* arch, const float* weights, size_t n_floats, int device,
                 bfcnn_handle** out);
void bfcnn_destroy(bfcnn_handle* h);
/* Workspaces (feature maps, staging buffers, saved activations) grow to the largest call seen and are kept for reuse; this
 * returns them to the driver (weights, packed operands and optimizer state stay).  The next call allocates again. */
int bfcnn_release_workspaces(bfcnn_handle* h);
int bfcnn_set_weights(bfcnn_handle* h, const float* weights, size_t n_floats);
int bfcnn_get_weights(bfcnn_handle* h, float* weights, size_t n_floats);

/* ---- inference ----------------------------------------------------------------
 * replaces: DenoiserModule.__call__ (bfcnn/module_denoiser.py:39-75):
 * uint8 NHWC -> float -> [pow2 canvas] -> normalise -> resnet -> head -> tanh(2y)*0.51
 * -> denormalise -> [crop] -> round-half-even -> uint8 NHWC. */
int bfcnn_denoise_u8(bfcnn_handle* h, const uint8_t* in, uint8_t* out, int n, int height, int width,
                     int precision, uint32_t flags, void* stream);
/* same, but returns the pre-round float32 prediction (0..255 scale) for parity checks. */
int bfcnn_denoise_f32(bfcnn_handle* h, const uint8_t* in, float* out, int n, int height, int width,
                      int precision, uint32_t flags, void* stream);
/* kernels launched by this handle since creation (bench.py's gpu_launches). */
int64_t bfcnn_launch_count(const bfcnn_handle* h);
/* record/elapse device time spent in the conv-stack kernels of the LAST denoise call
 * (events on the caller's stream; ms).  Used for the roofline of the dominant kernel. */
int bfcnn_last_stack_ms(bfcnn_handle* h, float* ms);

/* Per-launch device times of the conv-stack kernels of the LAST denoise call made with timing switched on (CUDA events
 * around every launch on the caller's stream; bench.py's roofline of the dominant kernel is measured with these inside
 * its own run).  kinds[i]: 0 = base conv, 1 = pass of the fused stack, 2 = last pass (head + encode fused in), -1 = the
 * idle time on the stream between the launch before and the launch after this entry. */
int bfcnn_set_kernel_timing(bfcnn_handle* h, int on);
int bfcnn_kernel_times(bfcnn_handle* h, float* ms, int* kinds, int capacity, int* count);

/* ---- training step -------------------------------------------------------------
 * replaces: dataset_builder.prepare_data_fn (bfcnn/dataset.py:120-238).
 * clean_u8 [n,h,w,3] -> clean_f32, noisy_f32 [n,h,w,3] (device pointers).  Sample s
 * uses Philox4x32-10 stream (seed, sample_offset + s): bit-identical to the oracle. */
int bfcnn_corrupt(bfcnn_handle* h, const uint8_t* clean_u8, float* clean_f32, float* noisy_f32,
                  int n, int height, int width, uint64_t seed, uint64_t sample_offset,
                  const bfcnn_noise_cfg* cfg, void* stream);

/* replaces: one level of multiscales_generator_fn (bfcnn/utilities.py:625-685, called at bfcnn/train_loop.py:239-247,274):
 * tf.nn.avg_pool2d(2x2, strides 2, VALID) -> clip [0,255] (clip_values) -> tf.round (round_values).  in: device float32
 * [n,h,w,3]; out: device float32 [n,h/2,w/2,3]. */
int bfcnn_downscale2x(bfcnn_handle* h, const float* in, float* out, int n, int height, int width, int clip_values,
                      int round_values, void* stream);

/* replaces: loss_function_builder(...)["denoiser"] (bfcnn/loss.py:190-247).
 * gt, pred: device float32 [n,h,w,3]; out5 (host): total, mae, rmse, hinged-mae, ssim loss (1 - mean SSIM; 0 when
 * ssim_multiplier == 0).  SSIM needs height, width >= 7. */
int bfcnn_loss(bfcnn_handle* h, const float* gt, const float* pred, int n, int height, int width,
               const bfcnn_loss_cfg* cfg, float* out5, void* stream);

/* replaces: train_step_single_gpu (bfcnn/train_loop.py:263-312): forward with BN batch
 * statistics, hinged-MAE (+RMSE) loss, L1/L2 weight regularisation, backward.
 * clean, noisy: device float32 [n,h,w,3] (0..255).  flat_grads: device float32
 * [bfcnn_num_trainable] in Keras trainable_variables order (the buffer a data-parallel
 * caller all-reduces).  losses5 (host, or NULL = asynchronous, see bfcnn_train_losses): total, denoiser total, mae,
 * regularisation, ssim loss.
 * update_moving != 0 applies the BN moving-statistics update (momentum 0.995). */
int bfcnn_train_step(bfcnn_handle* h, const float* clean, const float* noisy, int n, int height,
                     int width, const bfcnn_loss_cfg* cfg, float* flat_grads, float* losses5,
                     int update_moving, void* stream);

/* bfcnn_train_step with losses5 == NULL is ASYNCHRONOUS: nothing is copied back and the host does not wait, so that
 * corrupt -> step -> all-reduce -> Adam chain on the stream without a host round trip (the reference reads its loss
 * tensors only when it logs them, train_loop.py:439-559).  bfcnn_train_losses fetches the five scalars of the LAST step
 * (same order as losses5) and synchronises the stream. */
int bfcnn_train_losses(bfcnn_handle* h, float* losses5, void* stream);

/* Test hook: copy one activation map saved by the LAST bfcnn_train_step to `out` (device float32 [n,h,w,16]).
 * which: 0 = X_i, the input of block i (index N = output of the stack; backbone_blocks.py:240-242), 1 = T_i =
 * ReLU(conv_a(X_i)) (:174-178), 2 = U_i = conv_b(T_i) before BatchNormalization (:191-196).  The gradient-parity tests
 * read the ReLU masks the kernels actually used from T_i. */
int bfcnn_saved_activation(bfcnn_handle* h, int which, int index, float* out, void* stream);

/* Engine of the 3x3 convs inside bfcnn_train_step: 2 (default) = tcgen05 with the fp16 hi/lo split (3 MMAs per product,
 * conv error ~4e-6 on O(1) data, i.e. FP32-grade; row-streaming kernel, conv_t5.cu), 1 = the same arithmetic on
 * mma.sync (conv_x3.cu), 0 = FP32 FFMA (conv error ~1e-6).  They differ only in rounding; on tiny batches a single ReLU
 * whose pre-activation is ~1e-6 can switch and move one pixel's worth of gradient (tests/test_training_gpu.py states the
 * gates per engine). */
int bfcnn_set_train_engine(bfcnn_handle* h, int engine);

/* One 3x3 16->16 "same" convolution layer (keras Conv2D of utilities.py:195-196, no bias), device float32 NHWC16
 * in/out, weights [3,3,16,16] HWIO on the device.  engine 0 = FP32 FFMA, 1 = mma.sync with the fp16 hi/lo split, 2 = tcgen05
 * with the fp16 hi/lo split (the engines of the training step); exposed so that the layer kernels can be tested in isolation. */
int bfcnn_conv3x3(bfcnn_handle* h, const float* in, const float* weights, float* out, int n, int height, int width,
                  int engine, int relu, void* stream);

/* The one exchange step of data-parallel training (the reference is single-device and accumulates micro-batches
 * instead, train_loop.py:404-437): in-place ncclAllReduce(sum, float32) of the flat gradient over `nccl_comm` (an
 * ncclComm_t of the caller, passed as void*) on `stream`.  Averaging is left to bfcnn_adam_step's grad_scale = 1/world.
 * NCCL is dlopen'ed (libnccl.so.2) on first use; BFCNN_ERR_UNSUPPORTED if it is not installed. */
int bfcnn_allreduce_grads(bfcnn_handle* h, float* flat_grads, void* nccl_comm, void* stream);

/* replaces: optimizer.apply_gradients with keras Adam + global_clipnorm
 * (bfcnn/optimizer.py:145-224, bfcnn/train_loop.py:314-321, 421-434).  flat_grads is
 * the (all-reduced, averaged) device gradient; step counts from 1. */
int bfcnn_adam_step(bfcnn_handle* h, const float* flat_grads, float grad_scale,
                    const bfcnn_adam_cfg* cfg, int64_t step, void* stream);

/* ---- every other `type: "resnet"` configuration -------------------------------------------------
 * Reference-grade FP32 layer kernels for the block layouts the tcgen05 stacks do not cover: 1-3 convs per block, any
 * odd kernel size / filter counts, grouped and depthwise convs (bfcnn/backbone_resnet.py:149-178; the in-tree config
 * bfcnn/configs/resnet_color_1x6_bn_32x128x32_1x3x1_128x128_depthwise_l1_relu.json), initial / final BatchNormalization
 * (:266-276) and the ChannelwiseMultiplier / Multiplier scalings (bfcnn/custom_layers.py:1028-1162), which the host folds
 * into a per-channel (scale, bias) of the preceding conv.  All pointers are device pointers, float32 NHWC; no handle.
 *
 * prepare : uint8 [n,h,w,3] -> clip/255-0.5 on the zero-padded pow2 canvas [n,hc,wc,3] (module_denoiser.py:53-56,
 *           utilities.py:449-461,736-751)
 * conv2d  : Conv2D(groups) with kernel [k,k,cin/groups,cout] (depth_multiplier = 0) or DepthwiseConv2D with kernel
 *           [k,k,cin,depth_multiplier]; "same" zero padding, stride 1, no bias; y = acc*scale + bias (either may be NULL),
 *           optional ReLU, optional residual add (utilities.py:195-215, backbone_blocks.py:240-242)
 * finish  : head output [n,hc,wc,3] -> tanh(2y)*0.51 -> denormalise -> crop -> float32 or round-half-even uint8
 *           (model.py:342, utilities.py:435-443, module_denoiser.py:68-73) */
int bfcnn_generic_prepare(int device, const uint8_t* img, float* canvas, int n, int height, int width, int canvas_h,
                          int canvas_w, void* stream);
int bfcnn_generic_conv2d(int device, const float* in, float* out, const float* weights, const float* scale, const float* bias,
                         const float* residual, int n, int height, int width, int cin, int cout, int kernel, int groups,
                         int depth_multiplier, int relu, void* stream);
int bfcnn_generic_finish(int device, const float* y, void* out, int n, int height, int width, int canvas_h, int canvas_w,
                         int out_u8, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BFCNN_B200_H_ */
