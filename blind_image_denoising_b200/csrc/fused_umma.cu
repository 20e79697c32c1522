// fused_umma.cu -- the fused conv-BN-ReLU residual stack on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// Same contract as fused_f16.cu (one CTA owns a spatial region with a halo of 2*nblk pixels, runs nblk residual
// blocks on it without leaving the SM, per-layer zero padding by masking), different engine:
//
//   * The region is 128 pixels wide: one region row == the M = 128 rows of one tcgen05.mma.
//   * Activations live in shared memory as two channel-half planes (8 channels = 16 B per pixel per plane).
//     That is exactly the SWIZZLE_NONE K-major canonical layout (core matrix = 8 pixels x 16 B, SBO = 128 B,
//     LBO = plane stride), and because it is affine in the pixel index a 3x3 tap shift is just
//     "descriptor start address += shift * 16 B" -- no im2col, no data movement (tools/umma_probe.cu, Q1).
//   * A 16-cout GEMM (N = 16) would be shared-memory bound on the A operand (4 KB per MMA for 8 math cycles).
//     So every MMA computes the contributions of one input row q to the THREE output rows q-1, q, q+1 at once:
//     B = [16 cin x (3 dy x 16 cout)] = N 48, and the accumulator of output row r sits in TMEM columns
//     [16(r+1), 16(r+1)+16), so the three dy partial sums are accumulated by the tensor core itself
//     (D columns 16q .. 16q+47 of MMA(q)).  dx is handled by three MMAs with the A start shifted by -1/0/+1 pixel.
//     3 MMAs (M128 N48 K16) per 128-pixel row per conv; the epilogue reads 16 columns per pixel.
//   * Accumulators never leave TMEM between the MMA and the epilogue; the epilogue (bias, ReLU / residual add,
//     border mask, fp16 pack) runs in 16 warps that each own one 32-lane TMEM quarter of a row, writes the next
//     layer's A operand straight back into the shared-memory planes and re-zeroes the drained TMEM block.
//   * One elected thread issues every MMA; per-row mbarriers couple it to the epilogue warps
//     (mma_done[q] via tcgen05.commit, epi_done[r] via mbarrier.arrive), so layer l+1 chases layer l down the
//     region a few rows behind and the tensor pipe never drains at a layer boundary.
//
// Reference arithmetic: module_denoiser.py:53-73, utilities.py:449-461 (normalise), backbone_resnet.py:258-262
// (base conv), backbone_blocks.py:167-246 (block), model.py:297-342 (head), utilities.py:435-443 (denormalise).
#include "kernels.cuh"

namespace bfcnn {
namespace umma {

constexpr int RW = 128;                 // region width == UMMA M
constexpr int SLACK_PX = 8;             // pixels of slack before/after every plane (tap shift -1/+1)
constexpr int NSETS = 4;                // epilogue warp sets (4 warps each, one per TMEM lane quarter)
constexpr int EPI_WARPS = 4 * NSETS;
constexpr int NTHREADS = 32 * (EPI_WARPS + 1);   // + the MMA issuer warp
constexpr int MAX_RH = 30;              // (RH + 2) accumulator blocks of 16 columns <= 512 TMEM columns
constexpr int W_LAYER_BYTES = 3 * 48 * 16 * 2;   // B operand of one conv: [dx 3][N 48][K 16] fp16
constexpr int MAX_SMEM = 232448;
constexpr int MAX_LAYERS = 8;           // conv layers fused per pass

enum Epi { EPI_RELU_TO_T = 0, EPI_RES_TO_X = 1, EPI_RES_TO_GLOBAL = 2, EPI_RES_HEAD = 3 };

struct Params {
  const uint8_t* img;      // [n][h][w][3]
  const __half* fin;       // [n][he][we][16]
  __half* fout;            // [n][he][we][16]
  void* out;               // [n][h][w][3] uint8 or float
  const float* wbase;      // [k0*k0*3][16]
  const uint8_t* wumma;    // [2N][W_LAYER_BYTES]
  const float* bias;       // [2N][16]
  const float* whead;      // [16][4]
  int n, h, w, he, we;
  int k0, blk0, nblk;
  int first, last, out_u8;
  int rh, tw, th, tiles_x, tiles_y;
};

// ---------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (sm_100); layout type 0 = SWIZZLE_NONE
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                    // D = F32
  d |= 0u << 7;                    // A = F16
  d |= 0u << 10;                   // B = F16   (both K-major: bits 15, 16 = 0)
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar), "r"(cnt) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(taddr), "r"(z)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }

// ---------------------------------------------------------------------------- shared-memory map
struct Smem {
  uint32_t bars;       // mma_done[32], epi_done[32] (8 B each)
  uint32_t tmem_slot;  // uint32 written by tcgen05.alloc
  uint32_t head;       // float [16][4]
  uint32_t bias;       // float [MAX_LAYERS][16]
  uint32_t wts;        // [nlayers][W_LAYER_BYTES]
  uint32_t X[2], T[2]; // byte address of pixel 0 of each channel-half plane
  uint32_t plane_bytes;
};
__host__ __device__ inline uint32_t plane_bytes_of(int rh) { return (uint32_t)(rh * RW + 2 * SLACK_PX) * 16u; }
constexpr uint32_t SM_BARS = 0, SM_TMEM = 512, SM_HEAD = 528, SM_BIAS = 784, SM_WTS = 784 + MAX_LAYERS * 64;  // 1296
__host__ __device__ inline uint32_t planes_offset(int nlayers) { return (SM_WTS + (uint32_t)nlayers * W_LAYER_BYTES + 127u) & ~127u; }

// ---------------------------------------------------------------------------- epilogue of one row quarter
template <int EPI>
__device__ __forceinline__ void epilogue_row(const Params& p, const Smem& S, const float* s_bias_l, const float* s_head,
                                             uint32_t taddr, int r, int c, int oy, int ox, int b, int halo, bool rezero) {
  uint32_t v[16];
  tmem_ld16(taddr, v);
  if (rezero) tmem_zero16(taddr);
  const int gy = oy + r, gx = ox + c;
  const bool inside = (gy >= 0) && (gy < p.he) && (gx >= 0) && (gx < p.we);
  const uint32_t pix_off = (uint32_t)(r * RW + c) * 16u;
  if (EPI == EPI_RELU_TO_T) {
    uint4 lo, hi;
    lo.x = pack_h2(fmaxf(__uint_as_float(v[0]), 0.f), fmaxf(__uint_as_float(v[1]), 0.f));
    lo.y = pack_h2(fmaxf(__uint_as_float(v[2]), 0.f), fmaxf(__uint_as_float(v[3]), 0.f));
    lo.z = pack_h2(fmaxf(__uint_as_float(v[4]), 0.f), fmaxf(__uint_as_float(v[5]), 0.f));
    lo.w = pack_h2(fmaxf(__uint_as_float(v[6]), 0.f), fmaxf(__uint_as_float(v[7]), 0.f));
    hi.x = pack_h2(fmaxf(__uint_as_float(v[8]), 0.f), fmaxf(__uint_as_float(v[9]), 0.f));
    hi.y = pack_h2(fmaxf(__uint_as_float(v[10]), 0.f), fmaxf(__uint_as_float(v[11]), 0.f));
    hi.z = pack_h2(fmaxf(__uint_as_float(v[12]), 0.f), fmaxf(__uint_as_float(v[13]), 0.f));
    hi.w = pack_h2(fmaxf(__uint_as_float(v[14]), 0.f), fmaxf(__uint_as_float(v[15]), 0.f));
    if (!inside) { lo = make_uint4(0u, 0u, 0u, 0u); hi = lo; }
    sts128(S.T[0] + pix_off, lo);
    sts128(S.T[1] + pix_off, hi);
    return;
  }
  // X + conv_b'(T) + b'   (Add([x, previous]), backbone_blocks.py:240-242; BN folded)
  float f[16];
  {
    const uint4 x0 = lds128(S.X[0] + pix_off), x1 = lds128(S.X[1] + pix_off);
    const uint32_t xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 xv = unpack_h2(xs[i]);
      f[2 * i] = (__uint_as_float(v[2 * i]) + s_bias_l[2 * i]) + xv.x;
      f[2 * i + 1] = (__uint_as_float(v[2 * i + 1]) + s_bias_l[2 * i + 1]) + xv.y;
    }
  }
  if (EPI == EPI_RES_TO_X || EPI == EPI_RES_TO_GLOBAL) {
    uint4 lo, hi;
    lo.x = pack_h2(f[0], f[1]); lo.y = pack_h2(f[2], f[3]); lo.z = pack_h2(f[4], f[5]); lo.w = pack_h2(f[6], f[7]);
    hi.x = pack_h2(f[8], f[9]); hi.y = pack_h2(f[10], f[11]); hi.z = pack_h2(f[12], f[13]); hi.w = pack_h2(f[14], f[15]);
    if (EPI == EPI_RES_TO_X) {
      if (!inside) { lo = make_uint4(0u, 0u, 0u, 0u); hi = lo; }
      sts128(S.X[0] + pix_off, lo);
      sts128(S.X[1] + pix_off, hi);
    } else {
      const bool in_tile = inside && (c >= halo) && (c < RW - halo);
      if (in_tile) {
        uint4* o = reinterpret_cast<uint4*>(p.fout + ((((long long)b * p.he + gy) * p.we + gx) << 4));
        o[0] = lo;
        o[1] = hi;
      }
    }
  } else {  // EPI_RES_HEAD: collapsed 1x1 head + tanh(2y)*0.51 + denormalise (+ round + uint8)
    const bool in_img = inside && (gy < p.h) && (gx < p.w) && (c >= halo) && (c < RW - halo);
    if (in_img) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        const float4 wv = *reinterpret_cast<const float4*>(s_head + ch * 4);
        s0 = fmaf(f[ch], wv.x, s0); s1 = fmaf(f[ch], wv.y, s1); s2 = fmaf(f[ch], wv.z, s2);
      }
      const float r0 = head_activation(s0), r1 = head_activation(s1), r2 = head_activation(s2);
      const long long o = (((long long)b * p.h + gy) * p.w + gx) * 3;
      if (p.out_u8) {
        uint8_t* d = reinterpret_cast<uint8_t*>(p.out) + o;
        d[0] = (uint8_t)__float2int_rn(r0); d[1] = (uint8_t)__float2int_rn(r1); d[2] = (uint8_t)__float2int_rn(r2);
      } else {
        float* d = reinterpret_cast<float*>(p.out) + o;
        d[0] = r0; d[1] = r1; d[2] = r2;
      }
    }
  }
}

// ---------------------------------------------------------------------------- the pass kernel
__global__ void __launch_bounds__(NTHREADS, 1)
umma_pass_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nl = 2 * p.nblk;
  const uint32_t s0 = smem_u32(smem);
  Smem S;
  S.bars = s0 + SM_BARS; S.tmem_slot = s0 + SM_TMEM; S.head = s0 + SM_HEAD; S.bias = s0 + SM_BIAS; S.wts = s0 + SM_WTS;
  S.plane_bytes = plane_bytes_of(p.rh);
  {
    const uint32_t pl = s0 + planes_offset(nl);
    S.X[0] = pl + 0 * S.plane_bytes + SLACK_PX * 16; S.X[1] = pl + 1 * S.plane_bytes + SLACK_PX * 16;
    S.T[0] = pl + 2 * S.plane_bytes + SLACK_PX * 16; S.T[1] = pl + 3 * S.plane_bytes + SLACK_PX * 16;
  }
  uint8_t* g_planes = smem + planes_offset(nl);
  float* s_head = reinterpret_cast<float*>(smem + SM_HEAD);
  float* s_bias = reinterpret_cast<float*>(smem + SM_BIAS);

  int t = blockIdx.x;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int b = t / p.tiles_y;
  const int halo = 2 * p.nblk;
  const int oy = ty * p.th - halo, ox = tx * p.tw - halo;

  // ---------------- one-time setup: barriers, TMEM, weights
  if (tid < 64) mbar_init(S.bars + tid * 8, tid < 32 ? 1u : 128u);   // [0,32) mma_done (commit), [32,64) epi_done (128 threads)
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(S.tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  for (int i = tid; i < nl * (W_LAYER_BYTES / 16); i += NTHREADS)
    reinterpret_cast<uint4*>(smem + SM_WTS)[i] =
        reinterpret_cast<const uint4*>(p.wumma + (size_t)(2 * p.blk0) * W_LAYER_BYTES)[i];
  for (int i = tid; i < nl * C; i += NTHREADS) s_bias[i] = p.bias[(size_t)(2 * p.blk0) * C + i];
  if (tid < C * 4) s_head[tid] = p.whead[tid];

  // ---------------- stage the input region into X
  if (p.first) {
    const int k0 = p.k0, r0 = (k0 - 1) >> 1;
    const int sw = RW + 2 * r0, sh = p.rh + 2 * r0;
    // the uint8 tile and the base weights alias the T planes (not written before the first epilogue)
    float* s_wb = reinterpret_cast<float*>(g_planes + 2 * S.plane_bytes);
    uint8_t* s_img = reinterpret_cast<uint8_t*>(s_wb) + ((k0 * k0 * 3 * C * 4 + 15) & ~15);
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
    // base conv (FP32 FFMA), one pixel x 16 cout per thread
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
      uint4 lo, hi;
      lo.x = pack_h2(acc[0], acc[1]); lo.y = pack_h2(acc[2], acc[3]); lo.z = pack_h2(acc[4], acc[5]); lo.w = pack_h2(acc[6], acc[7]);
      hi.x = pack_h2(acc[8], acc[9]); hi.y = pack_h2(acc[10], acc[11]); hi.z = pack_h2(acc[12], acc[13]); hi.w = pack_h2(acc[14], acc[15]);
      sts128(S.X[0] + (uint32_t)pix * 16u, lo);
      sts128(S.X[1] + (uint32_t)pix * 16u, hi);
    }
  } else {
    for (int i = tid; i < p.rh * RW * 2; i += NTHREADS) {
      const int hf = i & 1, pix = i >> 1;
      const int r = pix / RW, c = pix % RW;
      const int gy = oy + r, gx = ox + c;
      const bool valid = (gy >= 0) && (gy < p.he) && (gx >= 0) && (gx < p.we);
      const long long o = valid ? (((((long long)b * p.he + gy) * p.we + gx) << 4) + 8 * hf) : 0;
      cp_async16_zfill(S.X[hf] + (uint32_t)pix * 16u, p.fin + o, valid);
    }
    cp_async_wait_all();
  }
  // zero the plane slack (read by the -1/+1 tap shifts of the first / last row); in the first pass the T planes
  // were scratch for the uint8 tile, so this happens only after every thread finished the base conv
  if (p.first) __syncthreads();
  if (tid < 4 * 2 * SLACK_PX) {
    const int pl = tid / (2 * SLACK_PX), k = tid % (2 * SLACK_PX);
    const uint32_t off = (uint32_t)pl * S.plane_bytes + (k < SLACK_PX ? (uint32_t)k * 16u : S.plane_bytes - (uint32_t)(2 * SLACK_PX - k) * 16u);
    *reinterpret_cast<uint4*>(g_planes + off) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();   // generic-proxy writes of X -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM);

  if (warp < EPI_WARPS) {
    // ================= epilogue warps =================
    const int quarter = warp & 3, set = warp >> 2;
    const uint32_t tq = tmem + ((uint32_t)(quarter * 32) << 16);
    const int c = quarter * 32 + lane;
    // zero this warp's share of the accumulator blocks, then release the MMA issuer
    for (int blk = set; blk < p.rh + 2; blk += NSETS) tmem_zero16(tq + blk * 16);
    tmem_wait_st();
    tc_fence_before();
    asm volatile("bar.sync 1, %0;\n" ::"r"(NTHREADS) : "memory");
    for (int l = 0; l < nl; ++l) {
      const int r_lo = l + 1, r_hi = p.rh - l - 1;
      const bool is_b = (l & 1) != 0, last_layer = (l + 1 == nl);
      const float* s_bias_l = s_bias + l * C;
      int r = r_lo + ((set - (r_lo % NSETS)) + NSETS) % NSETS;
      for (; r < r_hi; r += NSETS) {
        mbar_wait(S.bars + (uint32_t)(r + 1) * 8, (uint32_t)(l & 1));
        tc_fence_after();
        const uint32_t taddr = tq + (uint32_t)(r + 1) * 16;
        if (!is_b) epilogue_row<EPI_RELU_TO_T>(p, S, s_bias_l, s_head, taddr, r, c, oy, ox, b, halo, true);
        else if (!last_layer) epilogue_row<EPI_RES_TO_X>(p, S, s_bias_l, s_head, taddr, r, c, oy, ox, b, halo, true);
        else if (p.last) epilogue_row<EPI_RES_HEAD>(p, S, s_bias_l, s_head, taddr, r, c, oy, ox, b, halo, false);
        else epilogue_row<EPI_RES_TO_GLOBAL>(p, S, s_bias_l, s_head, taddr, r, c, oy, ox, b, halo, false);
        if (!last_layer) {
          tmem_wait_st();
          fence_async_smem();
          tc_fence_before();
          mbar_arrive(S.bars + (uint32_t)(32 + r) * 8);
        }
      }
    }
  } else {
    // ================= MMA issuer warp =================
    asm volatile("bar.sync 1, %0;\n" ::"r"(NTHREADS) : "memory");   // accumulators are zero
    tc_fence_after();
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 48);
      for (int l = 0; l < nl; ++l) {
        const uint32_t src = (l & 1) ? S.T[0] : S.X[0];
        const uint32_t wl = S.wts + (uint32_t)l * W_LAYER_BYTES;
        const int q_lo = l, q_hi = p.rh - l;
        for (int q = q_lo; q < q_hi; ++q) {
          if (l > 0) {
            const uint32_t par = (uint32_t)((l - 1) & 1);
            if (q == q_lo) mbar_wait(S.bars + (uint32_t)(32 + q) * 8, par);
            if (q + 1 < q_hi) mbar_wait(S.bars + (uint32_t)(32 + q + 1) * 8, par);
            tc_fence_after();
          }
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const uint64_t ad = make_desc(src + (uint32_t)(q * RW + dx - 1) * 16u, S.plane_bytes, 128);
            const uint64_t bd = make_desc(wl + (uint32_t)dx * (48 * 16 * 2), 48 * 16, 128);
            mma_f16_ss(tmem + (uint32_t)q * 16, ad, bd, idesc, 1u);
          }
          umma_commit(S.bars + (uint32_t)q * 8);
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

}  // namespace umma

// ------------------------------------------------------------------------------------
// host: pass planning
// ------------------------------------------------------------------------------------
static int env_int_u(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

int run_fused_stack_umma(bfcnn_handle* h, const uint8_t* d_in, void* d_out, bool out_u8, const Extent& e, cudaStream_t st) {
  using namespace umma;
  const int N = h->arch.no_layers, k0 = h->arch.base_kernel, r0 = (k0 - 1) / 2;
  if (N < 1) {
    set_error("the fused tensor-core stack needs no_layers >= 1 (use BFCNN_PREC_FP32)");
    return BFCNN_ERR_UNSUPPORTED;
  }
  static bool attr_set = false;
  if (!attr_set) {
    BF_CUDA(cudaFuncSetAttribute((const void*)umma_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    attr_set = true;
  }
  int kb = env_int_u("BFCNN_KB_UMMA", 2);
  kb = std::max(1, std::min(std::min(kb, N), MAX_LAYERS / 2));
  const int passes = (N + kb - 1) / kb;
  const size_t feat_halves = (size_t)e.n * e.he * e.we * C;
  if (passes > 1) {
    BF_CHECK(h->ws_feat[0].reserve(feat_halves * sizeof(__half)));
    if (passes > 2) BF_CHECK(h->ws_feat[1].reserve(feat_halves * sizeof(__half)));
  }
  for (int ps = 0; ps < passes; ++ps) {
    Params p;
    p.img = d_in; p.out = d_out;
    p.fin = (ps > 0) ? h->ws_feat[(ps - 1) & 1].as<__half>() : nullptr;
    p.fout = (ps + 1 < passes) ? h->ws_feat[ps & 1].as<__half>() : nullptr;
    p.wbase = h->d_base_f32.as<float>();
    p.wumma = h->d_conv_umma.as<uint8_t>();
    p.bias = h->d_bias_f32.as<float>();
    p.whead = h->d_head_f32.as<float>();
    p.n = e.n; p.h = e.h; p.w = e.w; p.he = e.he; p.we = e.we;
    p.k0 = k0;
    p.blk0 = ps * kb;
    p.nblk = std::min(kb, N - p.blk0);
    p.first = (ps == 0); p.last = (ps + 1 == passes); p.out_u8 = out_u8 ? 1 : 0;
    const int nl = 2 * p.nblk, halo = 2 * p.nblk;
    // shared-memory budget -> region rows (<= MAX_RH by TMEM capacity)
    const size_t fixed = planes_offset(nl) + 64;
    int rh_max = (int)((MAX_SMEM - fixed) / ((size_t)4 * RW * 16)) - 1;   // 4 planes, 2 KB per row each (+ slack)
    while (rh_max > 0 && fixed + (size_t)4 * plane_bytes_of(rh_max) > (size_t)MAX_SMEM) --rh_max;
    rh_max = std::min(rh_max, MAX_RH);
    if (p.first) {
      // the uint8 tile + base weights alias the two T planes
      while (rh_max > 2 * halo + 1) {
        const size_t need = (((size_t)k0 * k0 * 3 * C * 4 + 15) & ~size_t(15)) + (size_t)(rh_max + 2 * r0) * (RW + 2 * r0) * 3;
        if (need <= (size_t)2 * plane_bytes_of(rh_max)) break;
        --rh_max;
      }
    }
    const int rows_needed = p.last ? e.h : e.he;
    const int cols_needed = p.last ? e.w : e.we;
    const int th_max = rh_max - 2 * halo;
    p.tw = RW - 2 * halo;
    if (th_max < 1 || p.tw < 1) {
      set_error("fused tcgen05 pass does not fit (kb=%d)", kb);
      return BFCNN_ERR_INTERNAL;
    }
    p.tiles_y = (rows_needed + th_max - 1) / th_max;
    p.th = (rows_needed + p.tiles_y - 1) / p.tiles_y;   // balance the tile rows
    p.rh = p.th + 2 * halo;
    p.tiles_x = (cols_needed + p.tw - 1) / p.tw;
    const size_t smem = planes_offset(nl) + (size_t)4 * plane_bytes_of(p.rh);
    if (smem > (size_t)MAX_SMEM || p.rh > MAX_RH) {
      set_error("internal: tcgen05 pass smem %zu rh %d", smem, p.rh);
      return BFCNN_ERR_INTERNAL;
    }
    const long long grid = (long long)p.tiles_x * p.tiles_y * e.n;
    BF_REQUIRE(grid < (1ll << 31), "too many tiles");
    umma_pass_kernel<<<(unsigned)grid, NTHREADS, smem, st>>>(p);
    h->launches++;
    BF_CUDA(cudaGetLastError());
  }
  return BFCNN_OK;
}

}  // namespace bfcnn
