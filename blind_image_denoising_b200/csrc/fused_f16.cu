// fused_f16.cu -- the fused conv-BN-ReLU residual stack on tensor cores (sm_100a).
//
// One CTA owns one spatial region of RW x rh pixels (16 channels, fp16, NHWC in shared
// memory) and runs `nblk` residual blocks on it back to back without leaving the SM:
//     [first pass: uint8 decode + normalise + base conv]             (module_denoiser.py:53,
//                                                                     utilities.py:449-461,
//                                                                     backbone_resnet.py:258-262)
//     nblk x { T = ReLU(conv_a(X)) ; X = X + conv_b'(T) + b' }        (backbone_blocks.py:167-246,
//                                                                     BN folded)
//     [last pass: collapsed 1x1 head + tanh(2y)*0.51 + denormalise + round + uint8 encode]
//                                                                    (model.py:297-342,
//                                                                     utilities.py:435-443,
//                                                                     module_denoiser.py:71-73)
// The region carries a halo of 2*nblk pixels (the receptive field of the fused blocks);
// after every conv the out-of-extent positions are forced to zero, which is exactly the
// per-layer "same" zero padding of the reference.  Between passes the 16-channel map is
// spilled to HBM as fp16 NHWC (two planes hi/lo in F16X3 mode).
//
// Each 3x3 16->16 conv is an implicit GEMM on mma.sync.m16n8k16 (HMMA, fp32 accumulate):
// M = 16 consecutive pixels of one row, N = 16 cout (two n-tiles), K = 16 cin per tap.
// A fragments come straight from the NHWC tile with ldmatrix (the tap shift is an address
// offset, no im2col); a warp marches down a 16-pixel-wide column strip and keeps the A
// fragments of the previous two rows in registers, so each input row is read from shared
// memory 3 times (dx) instead of 9.  B fragments (weights) live in registers for the layer.
//   F16   : fp16 operands, 1 MMA per (tap, n-tile)
//   F16X3 : activations and weights split hi+lo (fp16 each), 3 MMAs per (tap, n-tile):
//           hi*hi + lo*hi + hi*lo  -> fp32-grade results (precision table)
// Why mma.sync and not tcgen05 for this shape: DESIGN.md section 4.
#include "kernels.cuh"

namespace bfcnn {

constexpr int RW = 64;           // region width in pixels == shared-memory row pitch
constexpr int NTHREADS = 512;    // 16 warps: 4 column strips x 4 row bands
constexpr int SLACK_PX = 8;      // pixels of slack before and after every plane
constexpr int PX_BYTES = 32;     // 16 channels fp16
constexpr int MAX_SMEM = 232448; // 227 KB opt-in limit

enum Epi { EPI_RELU_TO_T = 0, EPI_RES_TO_X = 1, EPI_RES_TO_GLOBAL = 2, EPI_RES_HEAD = 3 };

struct FusedParams {
  const uint8_t* img;      // [n][h][w][3]
  const __half* fin;       // [P][n][he][we][16]
  __half* fout;            // [P][n][he][we][16]
  void* out;               // [n][h][w][3] uint8 or float
  const float* wbase;      // [k0*k0*3][16]
  const uint32_t* frags;   // [2N][2][9][2][64]
  const float* bias;       // [2N][16]
  const float* whead;      // [16][4]
  long long plane_stride;  // halves between the hi and lo planes of fin / fout
  int n, h, w, he, we;
  int k0, blk0, nblk;
  int first, last, out_u8;
  int rh, tw, th, tiles_x, tiles_y;
};

// ------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint2 b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
// byte offset of (pixel, 16-byte half) inside a plane: the two halves of a pixel are XOR
// swizzled by bit 2 of the pixel index so that any 8 consecutive pixels cover all 32 banks.
__device__ __forceinline__ int px_off(int pix, int half) {
  return pix * PX_BYTES + ((half ^ ((pix >> 2) & 1)) << 4);
}

struct LayerCtx {
  uint32_t src[2], dst[2];  // shared byte addresses of pixel 0 of the hi / lo planes
  const uint32_t* frag;     // [2 planes][9][2][64]
  const float* bias;        // [16] or nullptr
  int lvl;                  // rows/cols [lvl, size-lvl) are produced by this layer
  int halo;
  int oy, ox, b;
};

// ------------------------------------------------------------------------------------
// epilogue of one 16-pixel row segment: acc[nt][0..1] -> pixel cA, acc[nt][2..3] -> pixel cA+8
// ------------------------------------------------------------------------------------
template <int P, int EPI>
__device__ __forceinline__ void epilogue_row(const FusedParams& p, const LayerCtx& L, float (&acc)[2][4],
                                             int r, int x0, int lane, const float (&bv)[4],
                                             const float (&wh)[4][3]) {
  const int q = lane & 3;
  const int gy = L.oy + r;
  const bool row_in = (gy >= 0) && (gy < p.he);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = x0 + (lane >> 2) + 8 * j;
    const int gx = L.ox + c;
    const bool inside = row_in && (gx >= 0) && (gx < p.we);
    const int pix = r * RW + c;
    float v[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      v[nt][0] = acc[nt][2 * j + 0];
      v[nt][1] = acc[nt][2 * j + 1];
    }
    if (EPI == EPI_RELU_TO_T) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        v[nt][0] = inside ? fmaxf(v[nt][0], 0.f) : 0.f;
        v[nt][1] = inside ? fmaxf(v[nt][1], 0.f) : 0.f;
      }
    } else {
      // X + conv_b'(T) + b'   (Add([x, previous]), backbone_blocks.py:240-242)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const uint32_t a0 = L.dst[0] + px_off(pix, nt) + q * 4;
        float2 x = unpack_h2(lds32(a0));
        if (P == 2) {
          const float2 xl = unpack_h2(lds32(L.dst[1] + px_off(pix, nt) + q * 4));
          x.x += xl.x; x.y += xl.y;
        }
        v[nt][0] = inside ? (v[nt][0] + bv[2 * nt + 0]) + x.x : 0.f;
        v[nt][1] = inside ? (v[nt][1] + bv[2 * nt + 1]) + x.y : 0.f;
      }
    }
    if (EPI == EPI_RELU_TO_T || EPI == EPI_RES_TO_X) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int o = px_off(pix, nt) + q * 4;
        const uint32_t hi = pack_h2(v[nt][0], v[nt][1]);
        sts32(L.dst[0] + o, hi);
        if (P == 2) {
          const float2 hf = unpack_h2(hi);
          sts32(L.dst[1] + o, pack_h2(v[nt][0] - hf.x, v[nt][1] - hf.y));
        }
      }
    } else if (EPI == EPI_RES_TO_GLOBAL) {
      const bool in_tile = inside && (c >= L.halo) && (c < RW - L.halo);
      if (in_tile) {
        const long long o = ((((long long)L.b * p.he + gy) * p.we + gx) << 4) + 2 * q;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const uint32_t hi = pack_h2(v[nt][0], v[nt][1]);
          *reinterpret_cast<uint32_t*>(p.fout + o + 8 * nt) = hi;
          if (P == 2) {
            const float2 hf = unpack_h2(hi);
            *reinterpret_cast<uint32_t*>(p.fout + p.plane_stride + o + 8 * nt) =
                pack_h2(v[nt][0] - hf.x, v[nt][1] - hf.y);
          }
        }
      }
    } else {  // EPI_RES_HEAD
      float s[3];
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        s[o] = v[0][0] * wh[0][o];
        s[o] = fmaf(v[0][1], wh[1][o], s[o]);
        s[o] = fmaf(v[1][0], wh[2][o], s[o]);
        s[o] = fmaf(v[1][1], wh[3][o], s[o]);
        s[o] += __shfl_xor_sync(0xffffffffu, s[o], 1);
        s[o] += __shfl_xor_sync(0xffffffffu, s[o], 2);
      }
      const bool in_img = row_in && (gx >= 0) && (gy < p.h) && (gx < p.w) && (c >= L.halo) && (c < RW - L.halo);
      if (in_img && q < 3) {
        const float y = (q == 0) ? s[0] : ((q == 1) ? s[1] : s[2]);
        const float rv = head_activation(y);
        const long long o = (((long long)L.b * p.h + gy) * p.w + gx) * 3 + q;
        if (p.out_u8)
          reinterpret_cast<uint8_t*>(p.out)[o] = (uint8_t)__float2int_rn(rv);
        else
          reinterpret_cast<float*>(p.out)[o] = rv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// one conv layer over the region, F16: sliding window of A fragments down a column strip
// ------------------------------------------------------------------------------------
template <int EPI>
struct RowStepF16 {
  template <int PH>
  static __device__ __forceinline__ void run(const FusedParams& p, const LayerCtx& L, uint32_t (&A)[3][3][4],
                                             const uint2 (&B)[9][2], const int (&aoff)[3], int r, int x0,
                                             int lane, const float (&bv)[4], const float (&wh)[4][3]) {
    constexpr int SP = (PH + 2) % 3;
    const uint32_t rowp = L.src[0] + (r + 1) * (RW * PX_BYTES);
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) ldsm4(A[SP][dx], rowp + aoff[dx]);
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      constexpr int S0 = PH;  // slot of row r-1
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int slot = (S0 + dy) % 3;
        mma16816(acc[0], A[slot][dx], B[dy * 3 + dx][0]);
        mma16816(acc[1], A[slot][dx], B[dy * 3 + dx][1]);
      }
    }
    epilogue_row<1, EPI>(p, L, acc, r, x0, lane, bv, wh);
  }
};

template <int EPI>
__device__ __forceinline__ void conv_layer_f16(const FusedParams& p, const LayerCtx& L, int warp, int lane,
                                               const float (&wh)[4][3]) {
  const int strip = warp & 3, band = warp >> 2;
  const int x0 = strip * 16;
  const int nrows = p.rh - 2 * L.lvl;
  const int r_begin = L.lvl + (nrows * band) / 4;
  const int r_end = L.lvl + (nrows * (band + 1)) / 4;

  uint2 B[9][2];
  const uint2* fr = reinterpret_cast<const uint2*>(L.frag);
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    B[t][0] = __ldg(fr + (t * 2 + 0) * 32 + lane);
    B[t][1] = __ldg(fr + (t * 2 + 1) * 32 + lane);
  }
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (EPI != EPI_RELU_TO_T) {
    const int q = lane & 3;
    bv[0] = L.bias[2 * q]; bv[1] = L.bias[2 * q + 1]; bv[2] = L.bias[8 + 2 * q]; bv[3] = L.bias[8 + 2 * q + 1];
  }
  int aoff[3];
  {
    const int i = (lane & 7) + ((lane >> 3) & 1) * 8, hf = lane >> 4;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) aoff[dx] = px_off(x0 + dx - 1 + i, hf);
  }
  if (r_begin >= r_end) return;
  uint32_t A[3][3][4];
#pragma unroll
  for (int dx = 0; dx < 3; ++dx) {
    ldsm4(A[0][dx], L.src[0] + (r_begin - 1) * (RW * PX_BYTES) + aoff[dx]);
    ldsm4(A[1][dx], L.src[0] + (r_begin) * (RW * PX_BYTES) + aoff[dx]);
  }
  for (int r = r_begin; r < r_end; r += 3) {
    RowStepF16<EPI>::template run<0>(p, L, A, B, aoff, r, x0, lane, bv, wh);
    if (r + 1 < r_end) RowStepF16<EPI>::template run<1>(p, L, A, B, aoff, r + 1, x0, lane, bv, wh);
    if (r + 2 < r_end) RowStepF16<EPI>::template run<2>(p, L, A, B, aoff, r + 2, x0, lane, bv, wh);
  }
}

// ------------------------------------------------------------------------------------
// one conv layer, F16X3: hi*hi + lo*hi + hi*lo, A re-read per tap (B hi/lo fill the registers)
// ------------------------------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void conv_layer_x3(const FusedParams& p, const LayerCtx& L, int warp, int lane,
                                              const float (&wh)[4][3]) {
  const int strip = warp & 3, band = warp >> 2;
  const int x0 = strip * 16;
  const int nrows = p.rh - 2 * L.lvl;
  const int r_begin = L.lvl + (nrows * band) / 4;
  const int r_end = L.lvl + (nrows * (band + 1)) / 4;

  uint2 Bh[9][2], Bl[9][2];
  const uint2* fr = reinterpret_cast<const uint2*>(L.frag);
#pragma unroll
  for (int t = 0; t < 9; ++t) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      Bh[t][nt] = __ldg(fr + (t * 2 + nt) * 32 + lane);
      Bl[t][nt] = __ldg(fr + 9 * 2 * 32 + (t * 2 + nt) * 32 + lane);
    }
  }
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (EPI != EPI_RELU_TO_T) {
    const int q = lane & 3;
    bv[0] = L.bias[2 * q]; bv[1] = L.bias[2 * q + 1]; bv[2] = L.bias[8 + 2 * q]; bv[3] = L.bias[8 + 2 * q + 1];
  }
  int aoff[3];
  {
    const int i = (lane & 7) + ((lane >> 3) & 1) * 8, hf = lane >> 4;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) aoff[dx] = px_off(x0 + dx - 1 + i, hf);
  }
  for (int r = r_begin; r < r_end; ++r) {
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int rowb = (r + dy - 1) * (RW * PX_BYTES);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        uint32_t ah[4], al[4];
        ldsm4(ah, L.src[0] + rowb + aoff[dx]);
        ldsm4(al, L.src[1] + rowb + aoff[dx]);
        const int t = dy * 3 + dx;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          mma16816(acc[nt], al, Bh[t][nt]);
          mma16816(acc[nt], ah, Bl[t][nt]);
          mma16816(acc[nt], ah, Bh[t][nt]);
        }
      }
    }
    epilogue_row<2, EPI>(p, L, acc, r, x0, lane, bv, wh);
  }
}

template <int P, int EPI>
__device__ __forceinline__ void conv_layer(const FusedParams& p, const LayerCtx& L, int warp, int lane,
                                           const float (&wh)[4][3]) {
  if (P == 1) conv_layer_f16<EPI>(p, L, warp, lane, wh);
  else conv_layer_x3<EPI>(p, L, warp, lane, wh);
}

// ------------------------------------------------------------------------------------
// the pass kernel
// ------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(NTHREADS, 1)
fused_pass_kernel(const FusedParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int plane_bytes = (p.rh * RW + 2 * SLACK_PX) * PX_BYTES;
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
  uint32_t sX[2], sT[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    sX[k] = s0 + (k % P) * plane_bytes + SLACK_PX * PX_BYTES;
    sT[k] = s0 + (P + (k % P)) * plane_bytes + SLACK_PX * PX_BYTES;
  }
  uint8_t* s_extra = smem + 2 * P * plane_bytes;

  int t = blockIdx.x;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int b = t / p.tiles_y;
  const int halo = 2 * p.nblk;
  const int oy = ty * p.th - halo, ox = tx * p.tw - halo;

  // ---------------- stage the input region into X
  if (p.first) {
    const int k0 = p.k0, r0 = (k0 - 1) >> 1;
    const int sw = RW + 2 * r0, sh = p.rh + 2 * r0;
    float* s_wb = reinterpret_cast<float*>(s_extra);
    uint8_t* s_img = s_extra + ((k0 * k0 * 3 * C * 4 + 15) & ~15);
    for (int i = tid; i < k0 * k0 * 3 * C; i += NTHREADS) s_wb[i] = p.wbase[i];
    const uint8_t* img_b = p.img + (long long)b * p.h * p.w * 3;
    for (int i = tid; i < sh * sw; i += NTHREADS) {
      const int ly = i / sw, lx = i - ly * sw;
      const int gy = oy - r0 + ly, gx = ox - r0 + lx;
      uint8_t v0 = 0, v1 = 0, v2 = 0;  // raw zeros outside the image (pow2 canvas, utilities.py:749)
      if (gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) {
        const uint8_t* s = img_b + ((long long)gy * p.w + gx) * 3;
        v0 = s[0]; v1 = s[1]; v2 = s[2];
      }
      s_img[i * 3 + 0] = v0; s_img[i * 3 + 1] = v1; s_img[i * 3 + 2] = v2;
    }
    __syncthreads();
    // base conv (FP32 FFMA; 0.5 % of the FLOPs), one pixel x 16 cout per thread
    for (int pix = tid; pix < p.rh * RW; pix += NTHREADS) {
      const int r = pix / RW, c = pix % RW;
      const int gy = oy + r, gx = ox + c;
      float acc[C];
#pragma unroll
      for (int k = 0; k < C; ++k) acc[k] = 0.f;
      if (gy >= 0 && gy < p.he && gx >= 0 && gx < p.we) {
        for (int dy = 0; dy < k0; ++dy) {
          const int yy = gy + dy - r0;
          if (yy < 0 || yy >= p.he) continue;  // zero padding of the NORMALISED tensor
          for (int dx = 0; dx < k0; ++dx) {
            const int xx = gx + dx - r0;
            if (xx < 0 || xx >= p.we) continue;
            const uint8_t* s = s_img + ((r + dy) * sw + (c + dx)) * 3;
            const float* wt = s_wb + (dy * k0 + dx) * 3 * C;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
              const float xn = __fsub_rn(__fdiv_rn((float)s[ci], 255.f), 0.5f);
#pragma unroll
              for (int k = 0; k < C; ++k) acc[k] = fmaf(xn, wt[ci * C + k], acc[k]);
            }
          }
        }
      }
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint4 hi;
        hi.x = pack_h2(acc[8 * hf + 0], acc[8 * hf + 1]);
        hi.y = pack_h2(acc[8 * hf + 2], acc[8 * hf + 3]);
        hi.z = pack_h2(acc[8 * hf + 4], acc[8 * hf + 5]);
        hi.w = pack_h2(acc[8 * hf + 6], acc[8 * hf + 7]);
        sts128(sX[0] + px_off(pix, hf), hi);
        if (P == 2) {
          uint4 lo;
          float2 f;
          f = unpack_h2(hi.x); lo.x = pack_h2(acc[8 * hf + 0] - f.x, acc[8 * hf + 1] - f.y);
          f = unpack_h2(hi.y); lo.y = pack_h2(acc[8 * hf + 2] - f.x, acc[8 * hf + 3] - f.y);
          f = unpack_h2(hi.z); lo.z = pack_h2(acc[8 * hf + 4] - f.x, acc[8 * hf + 5] - f.y);
          f = unpack_h2(hi.w); lo.w = pack_h2(acc[8 * hf + 6] - f.x, acc[8 * hf + 7] - f.y);
          sts128(sX[1] + px_off(pix, hf), lo);
        }
      }
    }
  } else {
    for (int i = tid; i < p.rh * RW * 2; i += NTHREADS) {
      const int hf = i & 1, pix = i >> 1;
      const int r = pix / RW, c = pix % RW;
      const int gy = oy + r, gx = ox + c;
      const bool valid = (gy >= 0) && (gy < p.he) && (gx >= 0) && (gx < p.we);
      const long long o = valid ? (((((long long)b * p.he + gy) * p.we + gx) << 4) + 8 * hf) : 0;
#pragma unroll
      for (int k = 0; k < P; ++k)
        cp_async16_zfill(sX[k] + px_off(pix, hf), p.fin + k * p.plane_stride + o, valid);
    }
    cp_async_wait_all();
  }
  __syncthreads();

  // ---------------- residual blocks
  float wh[4][3];
  {
    const int q = lane & 3;
    const int chs[4] = {2 * q, 2 * q + 1, 8 + 2 * q, 8 + 2 * q + 1};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int o = 0; o < 3; ++o) wh[i][o] = p.last ? p.whead[chs[i] * 4 + o] : 0.f;
  }
  LayerCtx L;
  L.halo = halo; L.oy = oy; L.ox = ox; L.b = b;
  for (int blk = 0; blk < p.nblk; ++blk) {
    const int l = 2 * (p.blk0 + blk);
    // conv_a + ReLU : X -> T
    L.src[0] = sX[0]; L.src[1] = sX[1]; L.dst[0] = sT[0]; L.dst[1] = sT[1];
    L.frag = p.frags + (size_t)l * (2 * 9 * 2 * 64);
    L.bias = nullptr;
    L.lvl = 2 * blk + 1;
    conv_layer<P, EPI_RELU_TO_T>(p, L, warp, lane, wh);
    __syncthreads();
    // conv_b' + b' + skip : T (+X) -> X
    L.src[0] = sT[0]; L.src[1] = sT[1]; L.dst[0] = sX[0]; L.dst[1] = sX[1];
    L.frag = p.frags + (size_t)(l + 1) * (2 * 9 * 2 * 64);
    L.bias = p.bias + (l + 1) * C;
    L.lvl = 2 * blk + 2;
    if (blk + 1 < p.nblk) {
      conv_layer<P, EPI_RES_TO_X>(p, L, warp, lane, wh);
      __syncthreads();
    } else if (p.last) {
      conv_layer<P, EPI_RES_HEAD>(p, L, warp, lane, wh);
    } else {
      conv_layer<P, EPI_RES_TO_GLOBAL>(p, L, warp, lane, wh);
    }
  }
}

// ------------------------------------------------------------------------------------
// host: pass planning
// ------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

int run_fused_stack(bfcnn_handle* h, const uint8_t* d_in, void* d_out, bool out_u8, const Extent& e,
                    int precision, cudaStream_t st) {
  const int N = h->arch.no_layers, k0 = h->arch.base_kernel, r0 = (k0 - 1) / 2;
  const int P = (precision == BFCNN_PREC_F16X3) ? 2 : 1;
  if (N < 1) {
    set_error("the fused tensor-core stack needs no_layers >= 1 (use BFCNN_PREC_FP32)");
    return BFCNN_ERR_UNSUPPORTED;
  }
  static bool attr_set_dev[64] = {};   // function attributes are per device: one flag per device ordinal
  bool& attr_set = attr_set_dev[h->device & 63];
  if (!attr_set) {
    BF_CUDA(cudaFuncSetAttribute((const void*)fused_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    BF_CUDA(cudaFuncSetAttribute((const void*)fused_pass_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    attr_set = true;
  }
  int kb = env_int(P == 1 ? "BFCNN_KB_F16" : "BFCNN_KB_F16X3", P == 1 ? 2 : 1);
  kb = std::max(1, std::min(kb, N));
  const int passes = (N + kb - 1) / kb;

  const size_t feat_halves = (size_t)e.n * e.he * e.we * C;
  if (passes > 1) {
    BF_CHECK(h->ws_feat[0].reserve(feat_halves * P * sizeof(__half)));
    if (passes > 2) BF_CHECK(h->ws_feat[1].reserve(feat_halves * P * sizeof(__half)));
  }

  for (int ps = 0; ps < passes; ++ps) {
    FusedParams p;
    p.img = d_in; p.out = d_out;
    p.fin = (ps > 0) ? h->ws_feat[(ps - 1) & 1].as<__half>() : nullptr;
    p.fout = (ps + 1 < passes) ? h->ws_feat[ps & 1].as<__half>() : nullptr;
    p.wbase = h->d_base_f32.as<float>();
    p.frags = h->d_conv_frag.as<uint32_t>();
    p.bias = h->d_bias_f32.as<float>();
    p.whead = h->d_head_f32.as<float>();
    p.plane_stride = (long long)feat_halves;
    p.n = e.n; p.h = e.h; p.w = e.w; p.he = e.he; p.we = e.we;
    p.k0 = k0;
    p.blk0 = ps * kb;
    p.nblk = std::min(kb, N - p.blk0);
    p.first = (ps == 0); p.last = (ps + 1 == passes); p.out_u8 = out_u8 ? 1 : 0;
    const int halo = 2 * p.nblk;
    // shared-memory budget -> region rows
    size_t extra = 0;
    int rh_max;
    {
      const int px_bytes_all = 2 * P * PX_BYTES;
      const size_t fixed = (size_t)2 * P * 2 * SLACK_PX * PX_BYTES;
      if (p.first) {
        // extra = base weights + (rh+2r0)*(RW+2r0)*3 image bytes: solve for rh
        const size_t wb = ((size_t)k0 * k0 * 3 * C * 4 + 15) & ~size_t(15);
        const size_t per_row = (size_t)RW * px_bytes_all + (size_t)(RW + 2 * r0) * 3;
        rh_max = (int)((MAX_SMEM - fixed - wb - (size_t)2 * r0 * (RW + 2 * r0) * 3 - 64) / per_row);
      } else {
        rh_max = (int)((MAX_SMEM - fixed) / ((size_t)RW * px_bytes_all));
      }
    }
    const int rows_needed = p.last ? e.h : e.he;
    const int cols_needed = p.last ? e.w : e.we;
    const int th_max = rh_max - 2 * halo;
    p.tw = RW - 2 * halo;
    if (th_max < 1 || p.tw < 1) {
      set_error("fused pass does not fit shared memory (kb=%d)", kb);
      return BFCNN_ERR_INTERNAL;
    }
    p.tiles_y = (rows_needed + th_max - 1) / th_max;
    p.th = (rows_needed + p.tiles_y - 1) / p.tiles_y;   // balance the tile rows
    p.rh = p.th + 2 * halo;
    p.tiles_x = (cols_needed + p.tw - 1) / p.tw;
    if (p.first) {
      const size_t wb = ((size_t)k0 * k0 * 3 * C * 4 + 15) & ~size_t(15);
      extra = wb + (size_t)(p.rh + 2 * r0) * (RW + 2 * r0) * 3;
    }
    const size_t smem = (size_t)2 * P * (p.rh * RW + 2 * SLACK_PX) * PX_BYTES + extra;
    if (smem > MAX_SMEM) {
      set_error("internal: fused pass smem %zu > %d", smem, MAX_SMEM);
      return BFCNN_ERR_INTERNAL;
    }
    const long long grid = (long long)p.tiles_x * p.tiles_y * e.n;
    BF_REQUIRE(grid < (1ll << 31), "too many tiles");
    if (P == 1) fused_pass_kernel<1><<<(unsigned)grid, NTHREADS, smem, st>>>(p);
    else fused_pass_kernel<2><<<(unsigned)grid, NTHREADS, smem, st>>>(p);
    h->launches++;
    BF_CUDA(cudaGetLastError());
  }
  return BFCNN_OK;
}

}  // namespace bfcnn
