// api.cu -- the extern "C" surface of libbfcnn_b200.so (include/bfcnn_b200.h)
#include <dlfcn.h>
#include <math.h>
#include <string.h>

#include "kernels.cuh"

using namespace bfcnn;

namespace {

int validate_arch(const bfcnn_arch* a) {
  BF_REQUIRE(a != nullptr, "arch is NULL");
  BF_REQUIRE(a->filters == C, "filters must be 16");
  BF_REQUIRE(a->in_channels == 3 && a->out_channels == 3, "in/out channels must be 3");
  BF_REQUIRE(a->no_layers >= 0 && a->no_layers <= 64, "no_layers must be in [0,64]");
  BF_REQUIRE(a->base_kernel == 1 || a->base_kernel == 3 || a->base_kernel == 5 || a->base_kernel == 7,
             "base_kernel must be 1, 3, 5 or 7");
  BF_REQUIRE(a->head_filters >= 1 && a->head_filters <= 64, "head_filters must be in [1,64]");
  BF_REQUIRE(a->bn_epsilon > 0.f, "bn_epsilon must be > 0");
  return BFCNN_OK;
}

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// utilities.py:736-751 emulation without the padded work
Extent make_extent(const bfcnn_handle* h, int n, int height, int width, uint32_t flags) {
  Extent e;
  e.n = n; e.h = height; e.w = width;
  if (flags & BFCNN_FLAG_NO_PAD_POW2) {
    e.he = height; e.we = width;
  } else {
    const int R = (h->arch.base_kernel - 1) / 2 + 2 * h->arch.no_layers;
    e.he = std::min(next_pow2(height), height + R);
    e.we = std::min(next_pow2(width), width + R);
  }
  return e;
}

int check_images(int n, int height, int width) {
  BF_REQUIRE(n >= 0 && height >= 0 && width >= 0, "negative image dimension");
  BF_REQUIRE((long long)n * height * width * 3 < (1ll << 40), "image batch too large");
  return BFCNN_OK;
}

// the conv stack of one batch (device pointers)
int run_stack(bfcnn_handle* h, const uint8_t* d_in, bool in_u8, void* d_out, bool out_u8, const Extent& e, int precision, cudaStream_t st) {
  if (precision == BFCNN_PREC_FP32) {
    const size_t feat = (size_t)e.n * e.he * e.we * C * sizeof(float);
    BF_CHECK(h->ws_feat[0].reserve(feat));
    BF_CHECK(h->ws_feat[1].reserve(feat));
    float* X = h->ws_feat[0].as<float>();
    float* T = h->ws_feat[1].as<float>();
    h->feat_tag[0] = h->feat_tag[1] = 0ull;   // the buffers no longer hold a streaming stack's layout
    BF_CHECK(launch_base_conv(h, d_in, in_u8, X, h->d_base_f32.as<float>(), e, st));
    for (int i = 0; i < h->arch.no_layers; ++i) {
      const float* wa = h->d_conv_f32.as<float>() + (size_t)(2 * i) * 9 * C * C;
      const float* wb = h->d_conv_f32.as<float>() + (size_t)(2 * i + 1) * 9 * C * C;
      const float* bb = h->d_bias_f32.as<float>() + (size_t)(2 * i + 1) * C;
      BF_CHECK(launch_conv3x3_f32(h, X, T, wa, nullptr, nullptr, nullptr, CONV_RELU, e, st));
      BF_CHECK(launch_conv3x3_f32(h, T, X, wb, bb, X, nullptr, CONV_RESIDUAL, e, st));
    }
    return launch_head(h, X, d_out, out_u8, h->d_head_f32.as<float>(), e, st);
  }
  if (precision == BFCNN_PREC_F16) return run_fused_stack_stream(h, d_in, d_out, out_u8, e, st);
  return run_fused_stack_stream_x3(h, d_in, d_out, out_u8, e, st);
}

int denoise_impl(bfcnn_handle* h, const uint8_t* in, void* out, bool out_u8, int n, int height, int width,
                 int precision, uint32_t flags, void* stream) {
  BF_REQUIRE(h != nullptr, "handle is NULL");
  BF_CHECK(check_images(n, height, width));
  BF_REQUIRE(precision == BFCNN_PREC_FP32 || precision == BFCNN_PREC_F16 || precision == BFCNN_PREC_F16X3, "unknown precision");
  const size_t npx = (size_t)n * height * width;
  if (npx == 0) return BFCNN_OK;  // empty batch / empty image: nothing to do
  // float32 input (the hydra model's own signature, model.py:100-102: the normaliser clips to [0,255]) runs on the
  // reference-grade FP32 path only: the tensor-core stacks rely on uint8 pixels being exact in fp16
  const bool in_u8 = !(flags & BFCNN_FLAG_IN_F32);
  BF_REQUIRE(in_u8 || precision == BFCNN_PREC_FP32, "float32 input needs precision BFCNN_PREC_FP32");
  const size_t isz = in_u8 ? 1 : sizeof(float);
  BF_REQUIRE(in != nullptr && out != nullptr, "in/out is NULL");
  BF_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t px_img = (size_t)height * width;
  const size_t in_bytes = npx * 3 * isz, osz = out_u8 ? 1 : sizeof(float), out_bytes = npx * 3 * osz;
  if (!h->packed_valid) {
    // a training / optimiser step changed the variables on the device: fold and pack them again
    BF_CUDA(cudaDeviceSynchronize());
    BF_CUDA(cudaMemcpy(h->h_vars.data(), h->d_vars.p, h->lay.total * sizeof(float), cudaMemcpyDeviceToHost));
    BF_CHECK(pack_weights(h));
  }
  const bool in_host = !(flags & BFCNN_FLAG_IN_DEVICE), out_host = !(flags & BFCNN_FLAG_OUT_DEVICE);
  if (in_host && out_host && st == nullptr) {
    // host in, host out, no caller stream: run on a stream the handle owns instead of the legacy default stream, so
    // that two handles driven from two host threads overlap (the copies of one batch with the conv stack of the other;
    // blind_image_denoising_b200.PipelinedDenoiser).  The call still returns only after its D2H copy has completed.
    if (!h->s_compute) BF_CUDA(cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking));
    st = h->s_compute;
  }
  if (in_host) BF_CHECK(h->ws_in.reserve(in_bytes));
  if (out_host) BF_CHECK(h->ws_out.reserve(out_bytes));
  const uint8_t* d_in = in_host ? h->ws_in.as<uint8_t>() : in;
  uint8_t* d_out = out_host ? h->ws_out.as<uint8_t>() : reinterpret_cast<uint8_t*>(out);

  // Host buffers: split the batch into chunks of >= ~4 MP and pipeline H2D(i+1) | conv stack(i) | D2H(i-1) on three
  // streams (pinned host memory makes the copies truly asynchronous).  Device buffers: one chunk on the caller's stream.
  int per_chunk = n;
  if ((in_host || out_host) && n > 1) per_chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, (4u << 20) / std::max<size_t>(px_img, 1)));
  const int chunks = (n + per_chunk - 1) / per_chunk;
  if (chunks > 1) {
    if (!h->s_h2d) BF_CUDA(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    if (!h->s_d2h) BF_CUDA(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    while ((int)h->ev_pool.size() < 2 * chunks + 1) {
      cudaEvent_t e;
      BF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->ev_pool.push_back(e);
    }
    // the copy streams start after whatever the caller queued on `st`
    BF_CUDA(cudaEventRecord(h->ev_pool[2 * chunks], st));
    BF_CUDA(cudaStreamWaitEvent(h->s_h2d, h->ev_pool[2 * chunks], 0));
  }
  BF_CUDA(cudaEventRecord(h->ev0, st));
  for (int c = 0; c < chunks; ++c) {
    const int i0 = c * per_chunk, nc = std::min(per_chunk, n - i0);
    const size_t off_px = (size_t)i0 * px_img * 3, off_in = off_px * isz, off_out = off_px * osz;
    const size_t cin = (size_t)nc * px_img * 3 * isz, cout = (size_t)nc * px_img * 3 * osz;
    if (in_host) {
      cudaStream_t sc = chunks > 1 ? h->s_h2d : st;
      BF_CUDA(cudaMemcpyAsync(h->ws_in.as<uint8_t>() + off_in, reinterpret_cast<const uint8_t*>(in) + off_in, cin, cudaMemcpyHostToDevice, sc));
      if (chunks > 1) {
        BF_CUDA(cudaEventRecord(h->ev_pool[2 * c], sc));
        BF_CUDA(cudaStreamWaitEvent(st, h->ev_pool[2 * c], 0));
      }
    }
    const Extent e = make_extent(h, nc, height, width, flags);
    BF_CHECK(run_stack(h, d_in + off_in, in_u8, d_out + off_out, out_u8, e, precision, st));
    if (out_host) {
      cudaStream_t sc = chunks > 1 ? h->s_d2h : st;
      if (chunks > 1) {
        BF_CUDA(cudaEventRecord(h->ev_pool[2 * c + 1], st));
        BF_CUDA(cudaStreamWaitEvent(sc, h->ev_pool[2 * c + 1], 0));
      }
      BF_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(out) + off_out, h->ws_out.as<uint8_t>() + off_out, cout, cudaMemcpyDeviceToHost, sc));
    }
  }
  BF_CUDA(cudaEventRecord(h->ev1, st));
  h->ev_valid = true;
  if (out_host) {
    // Wait for the result with a BLOCKING event: cudaStreamSynchronize spins, and with one spinning worker thread per
    // model instance and rank (PipelinedDenoiser: two per GPU) eight ranks keep sixteen host cores busy doing nothing
    // while the threads that have to issue the next copies wait for a core.
    if (!h->ev_done) BF_CUDA(cudaEventCreateWithFlags(&h->ev_done, cudaEventBlockingSync | cudaEventDisableTiming));
    if (chunks > 1) {
      BF_CUDA(cudaEventRecord(h->ev_done, h->s_d2h));
      BF_CUDA(cudaEventSynchronize(h->ev_done));
    }
    BF_CUDA(cudaEventRecord(h->ev_done, st));
    BF_CUDA(cudaEventSynchronize(h->ev_done));
  }
  return BFCNN_OK;
}

}  // namespace

extern "C" {

int bfcnn_abi_version(void) { return BFCNN_ABI_VERSION; }

const char* bfcnn_last_error(void) { return bfcnn::get_error(); }

int bfcnn_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return BFCNN_ERR_CUDA;
  }
  return n;
}

int64_t bfcnn_num_weights(const bfcnn_arch* arch) {
  if (validate_arch(arch) != BFCNN_OK) return BFCNN_ERR_INVALID_ARGUMENT;
  VarLayout L;
  L.build(*arch);
  return (int64_t)L.total;
}

int64_t bfcnn_num_trainable(const bfcnn_arch* arch) {
  if (validate_arch(arch) != BFCNN_OK) return BFCNN_ERR_INVALID_ARGUMENT;
  VarLayout L;
  L.build(*arch);
  return (int64_t)L.t_total;
}

int bfcnn_create(const bfcnn_arch* arch, const float* weights, size_t n_floats, int device, bfcnn_handle** out) {
  BF_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  BF_CHECK(validate_arch(arch));
  BF_REQUIRE(weights != nullptr, "weights is NULL");
  int ndev = 0;
  BF_CUDA(cudaGetDeviceCount(&ndev));
  if (ndev <= 0) {
    set_error("no CUDA device visible: libbfcnn_b200 has no CPU path");
    return BFCNN_ERR_CUDA;
  }
  BF_REQUIRE(device >= 0 && device < ndev, "device index out of range");
  BF_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  BF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return BFCNN_ERR_UNSUPPORTED;
  }
  bfcnn_handle* h = new (std::nothrow) bfcnn_handle();
  if (!h) { set_error("out of host memory"); return BFCNN_ERR_OUT_OF_MEMORY; }
  h->arch = *arch;
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->lay.build(*arch);
  if (n_floats != h->lay.total) {
    set_error("invalid argument: expected %zu weight floats, got %zu", h->lay.total, n_floats);
    delete h;
    return BFCNN_ERR_INVALID_ARGUMENT;
  }
  h->h_vars.assign(weights, weights + n_floats);
  int s = pack_weights(h);
  if (s == BFCNN_OK && cudaEventCreate(&h->ev0) != cudaSuccess) s = BFCNN_ERR_CUDA;
  if (s == BFCNN_OK && cudaEventCreate(&h->ev1) != cudaSuccess) s = BFCNN_ERR_CUDA;
  if (s != BFCNN_OK) { bfcnn_destroy(h); return s; }
  *out = h;
  return BFCNN_OK;
}

void bfcnn_destroy(bfcnn_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  h->d_vars.release(); h->d_base_f32.release(); h->d_conv_f32.release(); h->d_bias_f32.release();
  h->d_head_f32.release(); h->d_last_umma.release(); h->d_conv_umma.release(); h->d_conv_umma_x3.release();
  h->ws_in.release(); h->ws_out.release();
  for (auto& b : h->ws_feat) b.release();
  h->ws_train.release(); h->ws_stats.release(); h->ws_grads.release();
  h->adam_m.release(); h->adam_v.release();
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ktime_ev) cudaEventDestroy(e);
  if (h->s_compute) cudaStreamDestroy(h->s_compute);
  if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
  if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
  if (h->ev_done) cudaEventDestroy(h->ev_done);
  h->d_train_tables.release();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
}

int bfcnn_release_workspaces(bfcnn_handle* h) {
  BF_REQUIRE(h != nullptr, "handle is NULL");
  BF_CUDA(cudaSetDevice(h->device));
  BF_CUDA(cudaDeviceSynchronize());
  h->ws_in.release(); h->ws_out.release();
  for (auto& b : h->ws_feat) b.release();
  h->feat_tag[0] = h->feat_tag[1] = 0ull;
  h->ws_train.release(); h->ws_stats.release(); h->ws_grads.release();
  h->tr_n = h->tr_h = h->tr_w = 0;
  return BFCNN_OK;
}

int bfcnn_set_weights(bfcnn_handle* h, const float* weights, size_t n_floats) {
  BF_REQUIRE(h != nullptr && weights != nullptr, "NULL argument");
  if (n_floats != h->lay.total) {
    set_error("invalid argument: expected %zu weight floats, got %zu", h->lay.total, n_floats);
    return BFCNN_ERR_INVALID_ARGUMENT;
  }
  BF_CUDA(cudaSetDevice(h->device));
  BF_CUDA(cudaDeviceSynchronize());
  h->h_vars.assign(weights, weights + n_floats);
  return pack_weights(h);
}

int bfcnn_get_weights(bfcnn_handle* h, float* weights, size_t n_floats) {
  BF_REQUIRE(h != nullptr && weights != nullptr, "NULL argument");
  if (n_floats != h->lay.total) {
    set_error("invalid argument: expected %zu weight floats, got %zu", h->lay.total, n_floats);
    return BFCNN_ERR_INVALID_ARGUMENT;
  }
  BF_CUDA(cudaSetDevice(h->device));
  // the device copy is authoritative (training updates it in place)
  BF_CUDA(cudaDeviceSynchronize());
  BF_CUDA(cudaMemcpy(h->h_vars.data(), h->d_vars.p, n_floats * sizeof(float), cudaMemcpyDeviceToHost));
  memcpy(weights, h->h_vars.data(), n_floats * sizeof(float));
  return BFCNN_OK;
}

int bfcnn_denoise_u8(bfcnn_handle* h, const uint8_t* in, uint8_t* out, int n, int height, int width,
                     int precision, uint32_t flags, void* stream) {
  return denoise_impl(h, in, out, true, n, height, width, precision, flags, stream);
}

int bfcnn_denoise_f32(bfcnn_handle* h, const uint8_t* in, float* out, int n, int height, int width,
                      int precision, uint32_t flags, void* stream) {
  return denoise_impl(h, in, out, false, n, height, width, precision, flags, stream);
}

int64_t bfcnn_launch_count(const bfcnn_handle* h) { return h ? h->launches : 0; }

int bfcnn_last_stack_ms(bfcnn_handle* h, float* ms) {
  BF_REQUIRE(h != nullptr && ms != nullptr, "NULL argument");
  BF_REQUIRE(h->ev_valid, "no denoise call has been timed yet");
  BF_CUDA(cudaSetDevice(h->device));
  BF_CUDA(cudaEventSynchronize(h->ev1));
  BF_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return BFCNN_OK;
}

int bfcnn_set_kernel_timing(bfcnn_handle* h, int on) {
  BF_REQUIRE(h != nullptr, "handle is NULL");
  h->ktime_on = on != 0;
  h->ktime_n = 0;
  return BFCNN_OK;
}

int bfcnn_kernel_times(bfcnn_handle* h, float* ms, int* kinds, int capacity, int* count) {
  BF_REQUIRE(h != nullptr && ms != nullptr && kinds != nullptr && count != nullptr, "NULL argument");
  BF_CUDA(cudaSetDevice(h->device));
  int n = 0;
  for (int i = 0; i < h->ktime_n && n < capacity; ++i) {
    BF_CUDA(cudaEventSynchronize(h->ktime_ev[2 * i + 1]));
    if (i > 0 && n + 1 < capacity) {   // idle time on the stream between the previous launch and this one
      BF_CUDA(cudaEventElapsedTime(&ms[n], h->ktime_ev[2 * i - 1], h->ktime_ev[2 * i]));
      kinds[n++] = -1;
    }
    BF_CUDA(cudaEventElapsedTime(&ms[n], h->ktime_ev[2 * i], h->ktime_ev[2 * i + 1]));
    kinds[n++] = h->ktime_kind[i];
  }
  *count = n;
  return BFCNN_OK;
}

int bfcnn_corrupt(bfcnn_handle* h, const uint8_t* clean_u8, float* clean_f32, float* noisy_f32, int n,
                  int height, int width, uint64_t seed, uint64_t sample_offset, const bfcnn_noise_cfg* cfg,
                  void* stream) {
  BF_REQUIRE(h != nullptr && cfg != nullptr, "NULL argument");
  BF_CHECK(check_images(n, height, width));
  if ((size_t)n * height * width == 0) return BFCNN_OK;
  BF_REQUIRE(clean_u8 != nullptr && noisy_f32 != nullptr, "NULL image pointer");
  BF_CUDA(cudaSetDevice(h->device));
  return run_corrupt(h, clean_u8, clean_f32, noisy_f32, n, height, width, seed, sample_offset, cfg,
                     (cudaStream_t)stream);
}

int bfcnn_loss(bfcnn_handle* h, const float* gt, const float* pred, int n, int height, int width,
               const bfcnn_loss_cfg* cfg, float* out4, void* stream) {
  BF_REQUIRE(h != nullptr && cfg != nullptr && out4 != nullptr, "NULL argument");
  BF_CHECK(check_images(n, height, width));
  BF_REQUIRE((size_t)n * height * width > 0, "loss of an empty batch is undefined");
  BF_REQUIRE(gt != nullptr && pred != nullptr, "NULL image pointer");
  BF_REQUIRE(cfg->hinge >= 0.f && cfg->cutoff > 0.f, "hinge must be >= 0 and cutoff > 0");
  BF_CUDA(cudaSetDevice(h->device));
  return run_loss(h, gt, pred, n, height, width, cfg, out4, (cudaStream_t)stream);
}

int bfcnn_train_step(bfcnn_handle* h, const float* clean, const float* noisy, int n, int height, int width,
                     const bfcnn_loss_cfg* cfg, float* flat_grads, float* losses4, int update_moving,
                     void* stream) {
  BF_REQUIRE(h != nullptr && cfg != nullptr, "NULL argument");   // losses4 may be NULL: asynchronous step
  BF_CHECK(check_images(n, height, width));
  BF_REQUIRE((size_t)n * height * width > 0, "train step on an empty batch is undefined");
  BF_REQUIRE(clean != nullptr && noisy != nullptr && flat_grads != nullptr, "NULL pointer");
  BF_REQUIRE(cfg->hinge >= 0.f && cfg->cutoff > 0.f, "hinge must be >= 0 and cutoff > 0");
  BF_CUDA(cudaSetDevice(h->device));
  return run_train_step(h, clean, noisy, n, height, width, cfg, flat_grads, losses4, update_moving,
                        (cudaStream_t)stream);
}

int bfcnn_downscale2x(bfcnn_handle* h, const float* in, float* out, int n, int height, int width, int clip_values,
                      int round_values, void* stream) {
  BF_REQUIRE(h != nullptr, "handle is NULL");
  BF_CHECK(check_images(n, height, width));
  if ((size_t)n * (height / 2) * (width / 2) == 0) return BFCNN_OK;
  BF_REQUIRE(in != nullptr && out != nullptr, "NULL image pointer");
  BF_CUDA(cudaSetDevice(h->device));
  return run_downscale2x(h, in, out, n, height, width, clip_values, round_values, (cudaStream_t)stream);
}

int bfcnn_train_losses(bfcnn_handle* h, float* losses5, void* stream) {
  BF_REQUIRE(h != nullptr && losses5 != nullptr, "NULL argument");
  BF_CUDA(cudaSetDevice(h->device));
  return run_train_losses(h, losses5, (cudaStream_t)stream);
}

int bfcnn_saved_activation(bfcnn_handle* h, int which, int index, float* out, void* stream) {
  BF_REQUIRE(h != nullptr && out != nullptr, "NULL argument");
  BF_CUDA(cudaSetDevice(h->device));
  return run_saved_activation(h, which, index, out, (cudaStream_t)stream);
}

int bfcnn_set_train_engine(bfcnn_handle* h, int engine) {
  BF_REQUIRE(h != nullptr, "handle is NULL");
  BF_REQUIRE(engine == 0 || engine == 1 || engine == 2,
             "engine must be 0 (FP32 FFMA), 1 (mma.sync, fp16 hi/lo split) or 2 (tcgen05, fp16 hi/lo split)");
  h->train_engine = engine;
  return BFCNN_OK;
}

int bfcnn_conv3x3(bfcnn_handle* h, const float* in, const float* weights, float* out, int n, int height, int width,
                  int engine, int relu, void* stream) {
  BF_REQUIRE(h != nullptr && in != nullptr && weights != nullptr && out != nullptr, "NULL argument");
  BF_CHECK(check_images(n, height, width));
  if ((size_t)n * height * width == 0) return BFCNN_OK;
  BF_CUDA(cudaSetDevice(h->device));
  const Extent e{n, height, width, height, width};
  cudaStream_t st = (cudaStream_t)stream;
  const ConvEpi epi = relu ? CONV_RELU : CONV_PLAIN;
  if (engine == 0) return launch_conv3x3_f32(h, in, out, weights, nullptr, nullptr, nullptr, epi, e, st);
  if (engine == 1) return launch_conv3x3_x3(h, in, out, weights, nullptr, nullptr, epi, e, 64.0f, st);
  if (engine == 2) return launch_conv3x3_t5(h, in, out, weights, nullptr, nullptr, epi, e, 64.0f, st);
  set_error("invalid argument: engine must be 0 (FP32 FFMA), 1 (mma.sync hi/lo split) or 2 (tcgen05 hi/lo split)");
  return BFCNN_ERR_INVALID_ARGUMENT;
}

// NCCL is bound at run time (dlopen): the library has no link-time dependency on it, inference-only users never load it,
// and inside a PyTorch process the already loaded libnccl.so.2 (the one torch.distributed uses) is the one that is found.
namespace {
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);
nccl_allreduce_fn g_nccl_allreduce = nullptr;
nccl_errstr_fn g_nccl_errstr = nullptr;
int load_nccl() {
  if (g_nccl_allreduce) return BFCNN_OK;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { set_error("libnccl.so.2 cannot be loaded: %s", dlerror()); return BFCNN_ERR_UNSUPPORTED; }
  g_nccl_errstr = (nccl_errstr_fn)dlsym(lib, "ncclGetErrorString");
  g_nccl_allreduce = (nccl_allreduce_fn)dlsym(lib, "ncclAllReduce");
  if (!g_nccl_allreduce) { set_error("libnccl has no ncclAllReduce"); return BFCNN_ERR_UNSUPPORTED; }
  return BFCNN_OK;
}
}  // namespace

int bfcnn_allreduce_grads(bfcnn_handle* h, float* flat_grads, void* nccl_comm, void* stream) {
  BF_REQUIRE(h != nullptr && flat_grads != nullptr && nccl_comm != nullptr, "NULL argument");
  BF_CUDA(cudaSetDevice(h->device));
  BF_CHECK(load_nccl());
  // one in-place sum over the flat gradient (<= 84 272 floats, latency-bound: one buffer, one collective, the compute
  // stream); the 1/world average is the grad_scale of bfcnn_adam_step
  const int r = g_nccl_allreduce(flat_grads, flat_grads, h->lay.t_total, /*ncclFloat32*/ 7, /*ncclSum*/ 0, nccl_comm, (cudaStream_t)stream);
  if (r != 0) {
    set_error("ncclAllReduce failed: %s", g_nccl_errstr ? g_nccl_errstr(r) : "unknown NCCL error");
    return BFCNN_ERR_CUDA;
  }
  return BFCNN_OK;
}

int bfcnn_adam_step(bfcnn_handle* h, const float* flat_grads, float grad_scale, const bfcnn_adam_cfg* cfg,
                    int64_t step, void* stream) {
  BF_REQUIRE(h != nullptr && cfg != nullptr && flat_grads != nullptr, "NULL argument");
  BF_REQUIRE(step >= 1, "step counts from 1");
  BF_CUDA(cudaSetDevice(h->device));
  return run_adam_step(h, flat_grads, grad_scale, cfg, step, (cudaStream_t)stream);
}

}  // extern "C"
