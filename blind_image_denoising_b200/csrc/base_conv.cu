// base_conv.cu -- normalise + base conv k0 x k0, 3 -> 16 (utilities.py:449-461, backbone_resnet.py:258-262) from the uint8
// image into the fp16 NHWC16 feature map the streaming stacks read (hi part, and the lo part of the fp16 hi/lo split for
// the F16X3 stack), and the TMA descriptor of that map.
//   base_conv3_t5_kernel  : k0 = 3, tcgen05 (base_conv_t5.cu; the default of the shipped models)
//   base_conv_f16_kernel  : any other odd k0 <= 7, FP32 FFMA (this file)
#include "kernels.cuh"
#include "umma_ptx.cuh"

namespace bfcnn {
namespace bconv {

using namespace tc5;

// ---------------------------------------------------------------------------- base conv -> fp16 NHWC16
// normalise (utilities.py:449-461) + base conv k0 x k0, 3 -> 16 (backbone_resnet.py:258-262) over the work extent.
// CTA tile 64 x 16 pixels; the uint8 halo tile goes to shared memory already normalised (exactly the reference's
// x/255 - 0.5 in fp32): 0 outside the work extent (zero padding of the NORMALISED tensor), -0.5 on the raw-zero pow2
// canvas (utilities.py:749).  Each thread owns 4 consecutive pixels x 16 cout.
constexpr int BC_W = 64, BC_H = 16;
__global__ void __launch_bounds__(256, 2)
base_conv_f16_kernel(const uint8_t* __restrict__ img, __half* __restrict__ out, __half* __restrict__ out_lo /* or nullptr */,
                     const float* __restrict__ w, int h, int wd, int he, int we, int k0,
                     long long img_stride /* pixels between images */, long long row_stride /* pixels between rows */) {
  extern __shared__ __align__(16) float bsm[];
  const int r0 = (k0 - 1) >> 1;
  const int tw = BC_W + 2 * r0, th = BC_H + 2 * r0;
  float* s_w = bsm;                              // [k0*k0*3][16]
  float* s_in = bsm + k0 * k0 * 3 * C;           // [th][tw][3]
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * BC_W, y0 = blockIdx.y * BC_H, b = blockIdx.z;
  for (int i = tid; i < k0 * k0 * 3 * C; i += 256) s_w[i] = w[i];
  const uint8_t* img_b = img + (long long)b * h * wd * 3;
  for (int i = tid; i < th * tw; i += 256) {
    const int ly = i / tw, lx = i - ly * tw;
    const int gy = y0 + ly - r0, gx = x0 + lx - r0;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (gy >= 0 && gy < he && gx >= 0 && gx < we) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      if (gy < h && gx < wd) {
        const uint8_t* sp = img_b + ((long long)gy * wd + gx) * 3;
        a0 = (float)sp[0]; a1 = (float)sp[1]; a2 = (float)sp[2];
      }
      v0 = __fsub_rn(__fdiv_rn(a0, 255.f), 0.5f);
      v1 = __fsub_rn(__fdiv_rn(a1, 255.f), 0.5f);
      v2 = __fsub_rn(__fdiv_rn(a2, 255.f), 0.5f);
    }
    s_in[i * 3 + 0] = v0; s_in[i * 3 + 1] = v1; s_in[i * 3 + 2] = v2;
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][C];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < C; ++c) acc[p][c] = 0.f;
  for (int dy = 0; dy < k0; ++dy)
    for (int dx = 0; dx < k0; ++dx) {
      const float* ip = s_in + ((ty + dy) * tw + 4 * tx + dx) * 3;
      const float* wp = s_w + (dy * k0 + dx) * 3 * C;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        float wv[C];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t4 = *reinterpret_cast<const float4*>(wp + ci * C + 4 * q);
          wv[4 * q] = t4.x; wv[4 * q + 1] = t4.y; wv[4 * q + 2] = t4.z; wv[4 * q + 3] = t4.w;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float xv = ip[p * 3 + ci];
#pragma unroll
          for (int c = 0; c < C; ++c) acc[p][c] = fmaf(xv, wv[c], acc[p][c]);
        }
      }
    }
  const int gy = y0 + ty;
  if (gy >= he) return;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int gx = x0 + 4 * tx + p;
    if (gx >= we) continue;
    uint4 lo, hi;
    lo.x = pack_h2(acc[p][0], acc[p][1]); lo.y = pack_h2(acc[p][2], acc[p][3]); lo.z = pack_h2(acc[p][4], acc[p][5]); lo.w = pack_h2(acc[p][6], acc[p][7]);
    hi.x = pack_h2(acc[p][8], acc[p][9]); hi.y = pack_h2(acc[p][10], acc[p][11]); hi.z = pack_h2(acc[p][12], acc[p][13]); hi.w = pack_h2(acc[p][14], acc[p][15]);
    const long long oo = ((long long)b * img_stride + (long long)gy * row_stride + gx) << 4;
    uint4* o = reinterpret_cast<uint4*>(out + oo);
    o[0] = lo;
    o[1] = hi;
    if (out_lo) {   // the fp16 residual of every channel (F16X3 arithmetic)
      uint4 l0, l1;
      float2 f;
      f = unpack_h2(lo.x); l0.x = pack_h2(acc[p][0] - f.x, acc[p][1] - f.y);
      f = unpack_h2(lo.y); l0.y = pack_h2(acc[p][2] - f.x, acc[p][3] - f.y);
      f = unpack_h2(lo.z); l0.z = pack_h2(acc[p][4] - f.x, acc[p][5] - f.y);
      f = unpack_h2(lo.w); l0.w = pack_h2(acc[p][6] - f.x, acc[p][7] - f.y);
      f = unpack_h2(hi.x); l1.x = pack_h2(acc[p][8] - f.x, acc[p][9] - f.y);
      f = unpack_h2(hi.y); l1.y = pack_h2(acc[p][10] - f.x, acc[p][11] - f.y);
      f = unpack_h2(hi.z); l1.z = pack_h2(acc[p][12] - f.x, acc[p][13] - f.y);
      f = unpack_h2(hi.w); l1.w = pack_h2(acc[p][14] - f.x, acc[p][15] - f.y);
      uint4* ol = reinterpret_cast<uint4*>(out_lo + oo);
      ol[0] = l0;
      ol[1] = l1;
    }
  }
}


}  // namespace bconv

// ------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda)
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_feature_tmap(CUtensorMap* out, const __half* base, const Extent& e, int box_x, int box_y, long long row_px) {
  if (row_px == 0) row_px = e.we;   // pixels between rows (>= e.we)
  static tmap_encode_fn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    BF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return BFCNN_ERR_CUDA;
    }
    enc = reinterpret_cast<tmap_encode_fn>(fn);
  }
  // fp16 NHWC16 viewed as {ch8, half, x, y, n}
  const cuuint64_t dims[5] = {8, 2, (cuuint64_t)e.we, (cuuint64_t)e.he, (cuuint64_t)e.n};
  const cuuint64_t strides[4] = {16, 32, (cuuint64_t)row_px * 32, (cuuint64_t)e.he * row_px * 32};
  const cuuint32_t box[5] = {8, 1, (cuuint32_t)box_x, (cuuint32_t)box_y, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (extent %d x %d x %d)", (int)r, e.n, e.he, e.we);
    return BFCNN_ERR_CUDA;
  }
  return BFCNN_OK;
}



int launch_base_conv_f16(bfcnn_handle* h, const uint8_t* d_in, __half* feat, const Extent& e, cudaStream_t st, __half* feat_lo,
                         long long img_stride, long long row_stride) {
  using namespace bconv;
  if (img_stride == 0) { img_stride = (long long)e.he * e.we; row_stride = e.we; }   // [n][he][we][16]
  const int k0 = h->arch.base_kernel, r0 = (k0 - 1) / 2;
  // k0 = 3 into the virtual-row map of the streaming stacks: the tcgen05 kernel (base_conv_t5.cu)
  if (k0 == 3 && img_stride == e.we + 1 && row_stride == (long long)e.n * (e.we + 1))
    return launch_base_conv3_t5(h, d_in, feat, feat_lo, e, st);
  {
    const size_t bsm = (size_t)(k0 * k0 * 3 * C + (BC_H + 2 * r0) * (BC_W + 2 * r0) * 3) * sizeof(float);
    dim3 grid((e.we + BC_W - 1) / BC_W, (e.he + BC_H - 1) / BC_H, e.n);
    BF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "image too large for the base conv grid");
    base_conv_f16_kernel<<<grid, 256, bsm, st>>>(d_in, feat, feat_lo, h->d_base_f32.as<float>(), e.h, e.w, e.he, e.we, k0, img_stride, row_stride);
  }
  h->launches++;
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

}  // namespace bfcnn
