// base_conv.cu -- normalise + base conv k0 x k0, 3 -> 16 (utilities.py:449-461, backbone_resnet.py:258-262) from the uint8
// image into the fp16 NHWC16 feature map the streaming stacks read (hi part, and the lo part of the fp16 hi/lo split for
// the F16X3 stack), and the TMA descriptor of that map.
//   base_conv3_mma_kernel : k0 = 3, implicit GEMM on mma.sync.m16n8k16 at FP32-grade accuracy (default)
//   base_conv_f16_kernel  : any odd k0 <= 7, FP32 FFMA
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


// ---------------------------------------------------------------------------- base conv 3x3 on mma.sync (k0 == 3)
// The FFMA kernel above takes as long as a whole residual pass on 4K frames (issue-bound, 16-byte stores at a 128-byte
// stride).  For k0 = 3 the same arithmetic runs as an implicit GEMM on mma.sync.m16n8k16 at FP32-grade accuracy:
//   * the uint8 tile goes to shared memory as 4 fp16 channels per pixel (r, g, b, m): the raw value / 256 (exact in fp16)
//     and m = 1 inside the work extent, 0 outside.  x/255 - 0.5 is folded into the weights:
//         sum_taps w (v/255 - 0.5 m) = sum_taps (256/255 w) (v/256) + (-0.5 sum_c w) m
//     so out-of-extent taps (v = 0, m = 0) contribute nothing (zero padding of the NORMALISED tensor) and raw-zero canvas
//     pixels (v = 0, m = 1) contribute -0.5 w (utilities.py:749), as in the FFMA kernel;
//   * K = (dy 3, dx 3, c 4) = 36, padded to 48: with a pixel stride of 4 halves the im2col row of (pixel, dy) is 12
//     contiguous halves of the tile, so an A fragment register is one aligned 32-bit shared load;
//   * the weights are split into fp16 hi + lo (two MMAs per product; the activations are exact), accumulation in fp32;
//   * each warp stages its 16-pixel x 16-channel result through shared memory and stores 512 contiguous bytes.
constexpr int BM_W = 64, BM_H = 16;                 // CTA tile (pixels); 8 warps x 2 rows x 4 m16 tiles
constexpr int BM_TW = BM_W + 2, BM_TH = BM_H + 3;   // halo tile (+1 row: the K padding reads row dy = 3, times zero weights)
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(256, 4)
base_conv3_mma_kernel(const uint8_t* __restrict__ img, __half* __restrict__ out, __half* __restrict__ out_lo /* or nullptr */,
                      const float* __restrict__ w, int h, int wd, int he, int we, int tiles_x, int tiles_y, int tiles,
                      long long img_stride, long long row_stride) {
  __shared__ __align__(16) __half s_in[BM_TH * BM_TW * 4];
  __shared__ __align__(16) __half s_out[8][16 * 16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  // B fragments (once per CTA; the CTA then walks tiles with a stride of gridDim.x): w'(k, n), k = dy*12 + dx*4 + c
  uint32_t bh[3][2][2], bl[3][2][2];
#pragma unroll
  for (int ks = 0; ks < 3; ++ks)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float wv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = ks * 16 + hh * 8 + 2 * t + e, n = nt * 8 + g;
          float v = 0.f;
          if (k < 36) {
            const int tap = k >> 2, c = k & 3;
            if (c < 3) v = w[(tap * 3 + c) * C + n] * (256.0f / 255.0f);
            else v = -0.5f * (w[(tap * 3 + 0) * C + n] + w[(tap * 3 + 1) * C + n] + w[(tap * 3 + 2) * C + n]);
          }
          wv[e] = v;
        }
        const __half h0 = __float2half_rn(wv[0]), h1 = __float2half_rn(wv[1]);
        const __half l0 = __float2half_rn(wv[0] - __half2float(h0)), l1 = __float2half_rn(wv[1] - __half2float(h1));
        bh[ks][nt][hh] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        bl[ks][nt][hh] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      }
  // A fragment offsets (in halves) of this thread's k pairs: k = ks*16 + hh*8 + 2t -> (dy = k / 12, k % 12)
  int koff[3][2];
#pragma unroll
  for (int ks = 0; ks < 3; ++ks)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int k = ks * 16 + hh * 8 + 2 * t;
      koff[ks][hh] = (k / 12) * (BM_TW * 4) + (k % 12);
    }
  __half* so = s_out[warp];
#pragma unroll 1
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
  const int txi = tile % tiles_x, tyi = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
  const int x0 = txi * BM_W, y0 = tyi * BM_H;
  // input tile: (v/256, m) as 4 halves per pixel
  const uint8_t* img_b = img + (long long)b * h * wd * 3;
  __syncthreads();   // the previous tile's fragment loads are done
  for (int i = tid; i < BM_TH * BM_TW; i += 256) {
    const int ly = i / BM_TW, lx = i - ly * BM_TW;
    const int gy = y0 + ly - 1, gx = x0 + lx - 1;
    uint32_t p01 = 0u, p23 = 0u;
    if (ly < BM_H + 2 && gy >= 0 && gy < he && gx >= 0 && gx < we) {
      int a0 = 0, a1 = 0, a2 = 0;
      if (gy < h && gx < wd) {
        const uint8_t* sp = img_b + ((long long)gy * wd + gx) * 3;
        a0 = sp[0]; a1 = sp[1]; a2 = sp[2];
      }
      const __half v0 = __float2half_rn((float)a0 * 0.00390625f), v1 = __float2half_rn((float)a1 * 0.00390625f);
      const __half v2 = __float2half_rn((float)a2 * 0.00390625f);
      p01 = (uint32_t)__half_as_ushort(v0) | ((uint32_t)__half_as_ushort(v1) << 16);
      p23 = (uint32_t)__half_as_ushort(v2) | (0x3C00u << 16);   // m = 1.0h
    }
    *reinterpret_cast<uint2*>(s_in + i * 4) = make_uint2(p01, p23);
  }
  __syncthreads();
#pragma unroll 1
  for (int mt = 0; mt < 8; ++mt) {
    const int ry = 2 * warp + (mt >> 2), px0 = (mt & 3) * 16;   // tile row, first pixel of the m16 tile
    const __half* base0 = s_in + (ry * BM_TW + px0 + g) * 4;     // pixel (ry, px0+g), tap (dy 0, dx 0)
    const __half* base1 = base0 + 8 * 4;                         // pixel px0 + g + 8
    float acc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
      uint32_t a[4];
      a[0] = *reinterpret_cast<const uint32_t*>(base0 + koff[ks][0]);
      a[1] = *reinterpret_cast<const uint32_t*>(base1 + koff[ks][0]);
      a[2] = *reinterpret_cast<const uint32_t*>(base0 + koff[ks][1]);
      a[3] = *reinterpret_cast<const uint32_t*>(base1 + koff[ks][1]);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        mma16816(acc[nt], a, bl[ks][nt][0], bl[ks][nt][1]);
        mma16816(acc[nt], a, bh[ks][nt][0], bh[ks][nt][1]);
      }
    }
    // stage [16 px][16 ch] fp16, then 32 lanes x 16 B = the 512 contiguous bytes of the 16 pixels; for the F16X3 stacks
    // a second round stores the lo part (the rounding error of the fp16 value) into the lo feature map
    const int gy = y0 + ry, gx = x0 + px0 + (lane >> 1);
    const long long o = (((long long)b * img_stride + (long long)gy * row_stride + gx) << 4) + (lane & 1) * 8;
    for (int part = 0; part < (out_lo ? 2 : 1); ++part) {
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const uint32_t h01 = pack_h2(acc[nt][0], acc[nt][1]), h23 = pack_h2(acc[nt][2], acc[nt][3]);
        if (part == 0) {
          *reinterpret_cast<uint32_t*>(so + g * 16 + nt * 8 + 2 * t) = h01;
          *reinterpret_cast<uint32_t*>(so + (g + 8) * 16 + nt * 8 + 2 * t) = h23;
        } else {
          const float2 f01 = unpack_h2(h01), f23 = unpack_h2(h23);
          *reinterpret_cast<uint32_t*>(so + g * 16 + nt * 8 + 2 * t) = pack_h2(acc[nt][0] - f01.x, acc[nt][1] - f01.y);
          *reinterpret_cast<uint32_t*>(so + (g + 8) * 16 + nt * 8 + 2 * t) = pack_h2(acc[nt][2] - f23.x, acc[nt][3] - f23.y);
        }
      }
      __syncwarp();
      if (gy < he && gx < we) {
        const uint4 v = *reinterpret_cast<const uint4*>(so + lane * 8);
        *reinterpret_cast<uint4*>((part == 0 ? out : out_lo) + o) = v;
      }
    }
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
  // k0 = 3 into the virtual-row map of the streaming stacks: the tcgen05 kernel (base_conv_t5.cu); BFCNN_BASE_MMA_SYNC=1
  // keeps the mma.sync kernel below for A/B runs
  static const bool mma_sync = getenv("BFCNN_BASE_MMA_SYNC") != nullptr && atoi(getenv("BFCNN_BASE_MMA_SYNC")) != 0;
  if (k0 == 3 && !mma_sync && img_stride == e.we + 1 && row_stride == (long long)e.n * (e.we + 1))
    return launch_base_conv3_t5(h, d_in, feat, feat_lo, e, st);
  if (k0 == 3) {
    const int tx = (e.we + BM_W - 1) / BM_W, ty = (e.he + BM_H - 1) / BM_H;
    const long long tiles = (long long)tx * ty * e.n;
    BF_REQUIRE(tiles < (1ll << 31), "too many base-conv tiles");
    const int g3 = (int)std::min<long long>(tiles, 4ll * h->sm_count);
    base_conv3_mma_kernel<<<g3, 256, 0, st>>>(d_in, feat, feat_lo, h->d_base_f32.as<float>(), e.h, e.w, e.he, e.we, tx, ty, (int)tiles,
                                              img_stride, row_stride);
  } else {
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
