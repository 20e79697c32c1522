// common.cuh -- handle, error plumbing and shared geometry of libbfcnn_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/bfcnn_b200.h"

namespace bfcnn {

constexpr int C = 16;  // channels of every backbone conv

void set_error(const char* fmt, ...);
const char* get_error();

#define BF_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      bfcnn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                         \
      return (_e == cudaErrorMemoryAllocation) ? BFCNN_ERR_OUT_OF_MEMORY : BFCNN_ERR_CUDA; \
    }                                                                                     \
  } while (0)

#define BF_CHECK(expr)            \
  do {                            \
    int _s = (expr);              \
    if (_s != BFCNN_OK) return _s; \
  } while (0)

#define BF_REQUIRE(cond, msg)                      \
  do {                                             \
    if (!(cond)) {                                 \
      bfcnn::set_error("invalid argument: %s", msg); \
      return BFCNN_ERR_INVALID_ARGUMENT;           \
    }                                              \
  } while (0)

// A grow-only device buffer.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int reserve(size_t n);
  void release();
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Offsets (in floats) of the Keras-order flat variable vector.
struct VarLayout {
  int N, k0, F;
  size_t base;                   // [k0,k0,3,16]
  std::vector<size_t> wa, wb, gamma, mean, var;
  size_t h0, h1, total;
  // trainable subset (Keras trainable_variables order): base, (wa, wb, gamma)*N, h0, h1
  size_t t_base;
  std::vector<size_t> t_wa, t_wb, t_gamma;
  size_t t_h0, t_h1, t_total;
  void build(const bfcnn_arch& a);
};

// Work extent of one denoise call: the image lives on a pow2 canvas with raw
// zeros bottom/right; only rows < He = min(Hc, H+R) and cols < We = min(Wc, W+R) can
// influence the cropped output, so every feature map is computed on [0,He) x [0,We).
struct Extent {
  int n, h, w;    // image
  int he, we;     // work extent
};

}  // namespace bfcnn

struct bfcnn_handle {
  bfcnn_arch arch;
  int device = 0;
  bfcnn::VarLayout lay;
  std::vector<float> h_vars;  // Keras-order host copy

  // raw variables on the device (training reads/updates these)
  bfcnn::DevBuf d_vars;
  // inference-time folded / packed weights
  bfcnn::DevBuf d_base_f32;   // [k0*k0*3][16]
  bfcnn::DevBuf d_conv_f32;   // [2N][9][16 cin][16 cout], conv_b folded with BN scale
  bfcnn::DevBuf d_bias_f32;   // [2N][16], zero for conv_a, BN constant for conv_b
  bfcnn::DevBuf d_head_f32;   // [16][4] collapsed head (4th column zero)
  bfcnn::DevBuf d_conv_umma_x3; // the same with a lo part: [2N][hi/lo][dx 3][N 48][K 16] fp16
  bfcnn::DevBuf d_last_umma;  // last conv_b with the collapsed head folded in [dx 3][N 48][K 16] + the head matrix [N 16][K 16] (fp16)
  bfcnn::DevBuf d_conv_umma;  // tcgen05 B operands: [2N][dx 3][N 48 = (dy, cout)][K 16] fp16, K-major core matrices
  bool packed_valid = false;

  // workspaces
  bfcnn::DevBuf ws_in, ws_out;          // staging for host<->device images
  bfcnn::DevBuf ws_feat[3];             // feature maps
  // what ws_feat[0 / 1] currently hold: the streaming stacks zero the separator columns of their virtual-row layout once
  // per (buffer, geometry) and skip it while the tag still matches (a 2-D memset of ~9000 32-byte rows cost ~0.35 ms per
  // buffer and call on 4K frames); every other user of the buffers resets the tag to 0
  unsigned long long feat_tag[2] = {0ull, 0ull};
  bfcnn::DevBuf ws_train;               // saved activations for backward
  bfcnn::DevBuf ws_stats;               // BN batch statistics, reductions
  bfcnn::DevBuf ws_grads;               // scratch gradients
  bfcnn::DevBuf adam_m, adam_v;         // Adam moments over the trainable vector

  bfcnn::DevBuf d_train_tables;         // conv offsets / regulariser segments of the training step (uploaded once)
  int tr_n = 0, tr_h = 0, tr_w = 0;     // shape of the last training step (its saved activations sit in ws_train)
  size_t tr_out5_off = 0;               // byte offset of the last step's five loss scalars inside ws_stats
  int train_engine = 2;   // convs of the training step: 2 = tcgen05, fp16 hi/lo split (conv_t5.cu), 1 = the same on mma.sync (conv_x3.cu), 0 = FP32 FFMA
  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_done = nullptr;                   // blocking-sync event: host-buffer calls sleep on it instead of spinning
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;   // copy streams of the host-buffer pipeline (api.cu)
  cudaStream_t s_compute = nullptr;               // compute stream of host-to-host calls without a caller stream (api.cu)
  std::vector<cudaEvent_t> ev_pool;
  bool ev_valid = false;
  int sm_count = 148;
  // optional per-launch timing of the conv-stack kernels (bfcnn_set_kernel_timing): event pairs of the last call
  bool ktime_on = false;
  std::vector<cudaEvent_t> ktime_ev;
  std::vector<int> ktime_kind;      // 0 = base conv, 1 = pass, 2 = last pass (head fused in)
  int ktime_n = 0;
};

namespace bfcnn {
int pack_weights(bfcnn_handle* h);  // host fold + upload (host_pack.cu)
// identity of a streaming stack's feature-map layout in a workspace buffer (engine, allocation, geometry); never 0
inline unsigned long long feat_layout_tag(int engine, const void* ptr, const Extent& e) {
  unsigned long long t = 1469598103934665603ull;
  const unsigned long long parts[5] = {(unsigned long long)engine, (unsigned long long)(uintptr_t)ptr, (unsigned long long)e.n,
                                       (unsigned long long)e.he, (unsigned long long)e.we};
  for (unsigned long long v : parts) { t ^= v; t *= 1099511628211ull; }
  return t | 1ull;
}
// per-launch timing hooks (no-ops unless bfcnn_set_kernel_timing switched them on)
inline void ktime_begin(bfcnn_handle* h, cudaStream_t st, int kind) {
  if (!h->ktime_on) return;
  while ((int)h->ktime_ev.size() < 2 * (h->ktime_n + 1)) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    h->ktime_ev.push_back(e);
  }
  if ((int)h->ktime_kind.size() <= h->ktime_n) h->ktime_kind.resize(h->ktime_n + 1);
  h->ktime_kind[h->ktime_n] = kind;
  cudaEventRecord(h->ktime_ev[2 * h->ktime_n], st);
}
inline void ktime_end(bfcnn_handle* h, cudaStream_t st) {
  if (!h->ktime_on || (int)h->ktime_ev.size() < 2 * (h->ktime_n + 1)) return;
  cudaEventRecord(h->ktime_ev[2 * h->ktime_n + 1], st);
  h->ktime_n++;
}
}
