// conv_x3.cu -- the training-step 3x3 16->16 convolutions on tensor cores at FP32-grade accuracy.
//
// Same contract as conv3x3_c16_kernel (conv_f32.cu): fp32 NHWC16 in, fp32 NHWC16 out, zero "same" padding, fused
// epilogue (ReLU / residual add / per-channel batch statistics / ReLU-mask for the backward pass).  The arithmetic is
// the F16X3 scheme: activations and weights are split into fp16 hi + lo parts on the fly and every tap
// issues hi*hi + lo*hi + hi*lo on mma.sync.m16n8k16 with fp32 accumulation -- error ~2^-21 relative, i.e. FP32-grade,
// which is what the gradient parity gate (cosine >= 0.9999, max error <= 1e-3 of the gradient scale) needs.
// Used for the forward convs (backbone_blocks.py:167-246 in training mode) and for the dgrad convs of
// train_loop.py:302-304 (a correlation of dOut with the flipped, transposed kernel).
#include "kernels.cuh"

namespace bfcnn {
namespace x3 {

constexpr int RW = 64;           // region width in pixels == shared-memory row pitch (62 output columns + halo)
constexpr int NT = 256;          // 8 warps: 4 column strips x 2 row bands
constexpr int SLACK_PX = 8;
constexpr int PX_BYTES = 32;     // 16 channels fp16
constexpr float W_SCALE = 256.f;  // weights ~0.1: the low part of the split would be an fp16 subnormal without it
constexpr int TH_MAX = 25;       // rows per tile: 2 CTAs per SM (107 KB each) overlap each other's load and math

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint2 b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
// the two 16-byte halves of a pixel are XOR-swizzled by bit 2 of the pixel index (conflict-free ldmatrix rows)
__device__ __forceinline__ int px_off(int pix, int half) { return pix * PX_BYTES + ((half ^ ((pix >> 2) & 1)) << 4); }

__device__ __forceinline__ void split_store(uint32_t hi_addr, uint32_t lo_addr, const float4& a, const float4& b) {
  uint4 hi, lo;
  hi.x = pack_h2(a.x, a.y); hi.y = pack_h2(a.z, a.w); hi.z = pack_h2(b.x, b.y); hi.w = pack_h2(b.z, b.w);
  float2 f;
  f = unpack_h2(hi.x); lo.x = pack_h2(a.x - f.x, a.y - f.y);
  f = unpack_h2(hi.y); lo.y = pack_h2(a.z - f.x, a.w - f.y);
  f = unpack_h2(hi.z); lo.z = pack_h2(b.x - f.x, b.y - f.y);
  f = unpack_h2(hi.w); lo.w = pack_h2(b.z - f.x, b.w - f.y);
  sts128(hi_addr, hi);
  sts128(lo_addr, lo);
}

template <int EPI>
__global__ void __launch_bounds__(NT, 2)
conv3x3_x3_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w /*[9][16 cin][16 cout]*/,
                  const float* __restrict__ res, double* __restrict__ stats, int h, int wd, int th, int tiles_x, int tiles_y,
                  float in_scale, float out_scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float s_stat[2 * C];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rh = th + 2;
  const int plane_bytes = (rh * RW + 2 * SLACK_PX) * PX_BYTES;
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t sH = s0 + SLACK_PX * PX_BYTES, sL = s0 + plane_bytes + SLACK_PX * PX_BYTES;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int oy = ty * th - 1, ox = tx * (RW - 2) - 1;
  if (EPI == CONV_STATS && tid < 2 * C) s_stat[tid] = 0.f;

  // ---- stage the region: fp32 -> fp16 hi + lo planes, zero outside the image ("same" padding)
  const float* in_b = in + (long long)b * h * wd * C;
  for (int i = tid; i < rh * RW * 2; i += NT) {
    const int hf = i & 1, pix = i >> 1;
    const int r = pix / RW, c = pix % RW;
    const int gy = oy + r, gx = ox + c;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bq = a;
    if (gy >= 0 && gy < h && gx >= 0 && gx < wd) {
      const float4* s = reinterpret_cast<const float4*>(in_b + ((long long)gy * wd + gx) * C + 8 * hf);
      a = s[0]; bq = s[1];
      // power-of-two pre-scale (exact): keeps the LOW part of the split out of the fp16 subnormal range (activations ~0.1
      // have low parts ~2e-5 < 6.1e-5; back-propagated gradients are ~1e-6 to begin with)
      a.x *= in_scale; a.y *= in_scale; a.z *= in_scale; a.w *= in_scale;
      bq.x *= in_scale; bq.y *= in_scale; bq.z *= in_scale; bq.w *= in_scale;
    }
    split_store(sH + px_off(pix, hf), sL + px_off(pix, hf), a, bq);
  }
  // ---- B fragments (weights) hi / lo: lane holds B[k][n], k = 2q, 2q+1, 2q+8, 2q+9, n = 8*nt + g
  uint2 Bh[9][2], Bl[9][2];
  {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float* wp = w + tap * C * C + nt * 8 + g;
        const float w0 = __ldg(wp + (2 * q) * C) * W_SCALE, w1 = __ldg(wp + (2 * q + 1) * C) * W_SCALE;
        const float w2 = __ldg(wp + (2 * q + 8) * C) * W_SCALE, w3 = __ldg(wp + (2 * q + 9) * C) * W_SCALE;
        const uint32_t h01 = pack_h2(w0, w1), h23 = pack_h2(w2, w3);
        const float2 f01 = unpack_h2(h01), f23 = unpack_h2(h23);
        Bh[tap][nt] = make_uint2(h01, h23);
        Bl[tap][nt] = make_uint2(pack_h2(w0 - f01.x, w1 - f01.y), pack_h2(w2 - f23.x, w3 - f23.y));
      }
  }
  __syncthreads();

  const int strip = warp & 3, band = warp >> 2;
  const int x0 = strip * 16;
  const int r_begin = 1 + (th * band) / 2, r_end = 1 + (th * (band + 1)) / 2;
  int aoff[3];
  {
    const int i = (lane & 7) + ((lane >> 3) & 1) * 8, hf = lane >> 4;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) aoff[dx] = px_off(x0 + dx - 1 + i, hf);
  }
  const int q = lane & 3;
  float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = r_begin; r < r_end; ++r) {
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int rowb = (r + dy - 1) * (RW * PX_BYTES);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        uint32_t ah[4], al[4];
        ldsm4(ah, sH + rowb + aoff[dx]);
        ldsm4(al, sL + rowb + aoff[dx]);
        const int tap = dy * 3 + dx;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          mma16816(acc[nt], al, Bh[tap][nt]);
          mma16816(acc[nt], ah, Bl[tap][nt]);
          mma16816(acc[nt], ah, Bh[tap][nt]);
        }
      }
    }
    // ---- epilogue: acc[nt][0..1] -> pixel x0 + lane/4, acc[nt][2..3] -> pixel x0 + lane/4 + 8; channels 8*nt + 2q, +1
    const int gy = oy + r;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = x0 + (lane >> 2) + 8 * j;
      const int gx = ox + c;
      if (c < 1 || c >= RW - 1 || gx >= wd || gy >= h) continue;
      const long long o = (((long long)b * h + gy) * wd + gx) * C;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        float2 v = make_float2(acc[nt][2 * j] * out_scale, acc[nt][2 * j + 1] * out_scale);
        const long long oc = o + nt * 8 + 2 * q;
        if (EPI == CONV_STATS) {
          ssum[2 * nt] += v.x; ssum[2 * nt + 1] += v.y;
          ssq[2 * nt] += v.x * v.x; ssq[2 * nt + 1] += v.y * v.y;
        }
        if (EPI == CONV_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
        if (EPI == CONV_RESIDUAL) {
          const float2 rv = *reinterpret_cast<const float2*>(res + oc);
          v.x += rv.x; v.y += rv.y;
        }
        if (EPI == CONV_MASK) {   // ReLU backward: pass the gradient where the saved activation is > 0
          const float2 mv = *reinterpret_cast<const float2*>(res + oc);
          v.x = mv.x > 0.f ? v.x : 0.f; v.y = mv.y > 0.f ? v.y : 0.f;
        }
        *reinterpret_cast<float2*>(out + oc) = v;
      }
    }
  }
  if (EPI == CONV_STATS) {
    // lanes that share q own the same 4 channels (2q, 2q+1, 8+2q, 8+2q+1)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a = ssum[k], s2 = ssq[k];
#pragma unroll
      for (int s = 4; s < 32; s <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, s); s2 += __shfl_xor_sync(0xffffffffu, s2, s); }
      if (lane < 4) {
        const int ch = (k >> 1) * 8 + 2 * q + (k & 1);
        atomicAdd(&s_stat[ch], a);
        atomicAdd(&s_stat[C + ch], s2);
      }
    }
    __syncthreads();
    if (tid < 2 * C) atomicAdd(&stats[tid], (double)s_stat[tid]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// wgrad on tensor cores: dW[tap][ci][co] = sum_p A[p + tap][ci] * G[p][co]  (zero padding), fp16 hi/lo split (3 MMAs).
// Per tap this is a [16 ci x 16 co] GEMM whose K dimension runs over pixels: A^T and G come straight from the NHWC16
// tiles with ldmatrix.trans (rows = pixels), 16 pixels of one image row per K step.  Every warp keeps the full
// [9][16][16] partial in registers (72 accumulators per lane) over all the tiles its CTA visits (persistent grid); the
// warps are summed in fixed order through shared memory and the per-CTA partials by wgrad_reduce_kernel: deterministic.
// ------------------------------------------------------------------------------------------------------------------
constexpr int WG_TH = 12;        // tile rows (two 108 KB CTAs per SM)
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// Staging of one wgrad tile by NTH threads: activations with the one-pixel halo and output gradients (zero on the halo
// columns and outside the image) from fp32 NHWC16 into the fp16 hi / lo planes.  Tiles walk the VIRTUAL row of the batch:
// the n images side by side, one zero column between neighbours (it is the zero padding both of them see), so a
// 256-pixel-wide image does not pay for a fifth, nearly empty 62-column tile.  All global loads of a batch of SB
// iterations are issued first, then the splits and shared-memory stores (split_store is inline asm with a memory
// clobber, so a load placed after it cannot move above it: one load per iteration meant 13 serial global-load
// latencies per tile).
template <int NTH, int SBA, int SBG>
__device__ __forceinline__ void wgrad_stage_tile(int lt, uint32_t aH, uint32_t aL, uint32_t gH, uint32_t gL, const float* __restrict__ A,
                                                 const float* __restrict__ G, int n, int h, int wd, int oy, int ox, float g_scale) {
  constexpr int RH = WG_TH + 2;
  // a thread stages the same column and channel half in every iteration (NTH / 2 is a multiple of RW)
  static_assert((NTH / 2) % RW == 0, "staging assumes a fixed column per thread");
  const int c_t = (lt >> 1) % RW, hf_t = lt & 1;
  const int vx = ox + c_t, vb = vx >= 0 ? vx / (wd + 1) : -1, gx_t = vx - vb * (wd + 1);
  const bool col_ok = vb >= 0 && vb < n && gx_t < wd;
  const long long col_off = col_ok ? (((long long)vb * h * wd + gx_t) * C + 8 * hf_t) : 0;
  const float* a_b = A + col_off;
  const float* g_b = G + col_off;
  for (int i0 = lt; i0 < RH * RW * 2; i0 += NTH * SBA) {
    float4 u[SBA], v[SBA];
#pragma unroll
    for (int k = 0; k < SBA; ++k) {
      const int i = i0 + k * NTH;
      const int gy = oy + (i >> 1) / RW;
      u[k] = make_float4(0.f, 0.f, 0.f, 0.f); v[k] = u[k];
      if (i < RH * RW * 2 && col_ok && gy >= 0 && gy < h) {
        const float4* sp = reinterpret_cast<const float4*>(a_b + (long long)gy * wd * C);
        u[k] = sp[0]; v[k] = sp[1];
      }
    }
#pragma unroll
    for (int k = 0; k < SBA; ++k) {
      const int i = i0 + k * NTH;
      if (i >= RH * RW * 2) break;
      const int hf = i & 1, pix = i >> 1;
      float4 a = u[k], bq = v[k];
      a.x *= 64.f; a.y *= 64.f; a.z *= 64.f; a.w *= 64.f; bq.x *= 64.f; bq.y *= 64.f; bq.z *= 64.f; bq.w *= 64.f;
      split_store(aH + px_off(pix, hf), aL + px_off(pix, hf), a, bq);
    }
  }
  for (int i0 = lt; i0 < WG_TH * RW * 2; i0 += NTH * SBG) {
    float4 u[SBG], v[SBG];
#pragma unroll
    for (int k = 0; k < SBG; ++k) {
      const int i = i0 + k * NTH;
      const int gy = oy + 1 + (i >> 1) / RW;
      u[k] = make_float4(0.f, 0.f, 0.f, 0.f); v[k] = u[k];
      if (i < WG_TH * RW * 2 && col_ok && c_t >= 1 && c_t < RW - 1 && gy < h) {
        const float4* sp = reinterpret_cast<const float4*>(g_b + (long long)gy * wd * C);
        u[k] = sp[0]; v[k] = sp[1];
      }
    }
#pragma unroll
    for (int k = 0; k < SBG; ++k) {
      const int i = i0 + k * NTH;
      if (i >= WG_TH * RW * 2) break;
      const int hf = i & 1, pix = i >> 1;
      float4 a = u[k], bq = v[k];
      a.x *= g_scale; a.y *= g_scale; a.z *= g_scale; a.w *= g_scale; bq.x *= g_scale; bq.y *= g_scale; bq.z *= g_scale; bq.w *= g_scale;
      split_store(gH + px_off(pix, hf), gL + px_off(pix, hf), a, bq);
    }
  }
}

// The MMA phase of one tile for one of 8 warps.  Work unit = half a tile row (two K steps of 16 pixels): 24 units, 3 per
// warp.  The two accumulators of a tap form two independent HMMA chains (each still receives lo*hi, hi*lo, hi*hi in order).
__device__ __forceinline__ void wgrad_mma_tile(float (&acc)[9][2][4], int warp, int lane, uint32_t aH, uint32_t aL, uint32_t gH, uint32_t gL) {
  // ldmatrix.trans row addresses: lanes 0-7 / 8-15 / 16-23 / 24-31 give the rows of the four 8x8 matrices
  //   (pixels 0-7, ch 0-7), (pixels 0-7, ch 8-15), (pixels 8-15, ch 0-7), (pixels 8-15, ch 8-15)
  const int lp = (lane & 7) + ((lane >> 4) & 1) * 8, lh = (lane >> 3) & 1;
  // for the B operand (G) the x4 order is (px 0-7, co 0-7), (px 8-15, co 0-7), (px 0-7, co 8-15), (px 8-15, co 8-15)
  const int gp = (lane & 7) + ((lane >> 3) & 1) * 8, gh = (lane >> 4) & 1;
  static_assert((RW / 16) % 2 == 0 && (WG_TH * 2) % 8 == 0, "units must divide evenly over the warps");
  for (int u = warp; u < WG_TH * 2; u += 8) {
    const int r = u >> 1;
#pragma unroll 1
    for (int ks = (u & 1) * (RW / 32); ks < ((u & 1) + 1) * (RW / 32); ++ks) {
      uint32_t bh[4], bl[4];
      const int gpix = r * RW + ks * 16 + gp;
      ldsm4t(bh, gH + px_off(gpix, gh));
      ldsm4t(bl, gL + px_off(gpix, gh));
      // the fragments of tap t + 1 are requested before the MMAs of tap t are issued (the asm statements keep their order, so
      // without this every group of six MMAs waited for its own ldmatrix: "short scoreboard" was the top HMMA stall)
      uint32_t ah[2][4], al[2][4];
      const int apix0 = r * RW + ks * 16 - 1 + lp;
      ldsm4t(ah[0], aH + px_off(apix0, lh));
      ldsm4t(al[0], aL + px_off(apix0, lh));
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int cur = t & 1, nxt = cur ^ 1;
        if (t + 1 < 9) {
          const int apix = (r + (t + 1) / 3) * RW + ks * 16 + (t + 1) % 3 - 1 + lp;
          ldsm4t(ah[nxt], aH + px_off(apix, lh));
          ldsm4t(al[nxt], aL + px_off(apix, lh));
        }
        // A fragment order of mma (a0: m 0-7 k 0-7, a1: m 8-15 k 0-7, a2: m 0-7 k 8-15, a3: m 8-15 k 8-15) == load order
        mma16816(acc[t][0], al[cur], make_uint2(bh[0], bh[1]));
        mma16816(acc[t][1], al[cur], make_uint2(bh[2], bh[3]));
        mma16816(acc[t][0], ah[cur], make_uint2(bl[0], bl[1]));
        mma16816(acc[t][1], ah[cur], make_uint2(bl[2], bl[3]));
        mma16816(acc[t][0], ah[cur], make_uint2(bh[0], bh[1]));
        mma16816(acc[t][1], ah[cur], make_uint2(bh[2], bh[3]));
      }
    }
  }
}

// the 8 MMA warps' accumulators in fixed order through shared memory, then one partial per CTA
__device__ __forceinline__ void wgrad_store_acc(const float (&acc)[9][2][4], int warp, int lane, uint8_t* smem) {
  float* s_red = reinterpret_cast<float*>(smem);   // [8][2304]
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      float* d = s_red + warp * 2304 + t * C * C + nt * 8 + 2 * q;
      d[g * C] = acc[t][nt][0]; d[g * C + 1] = acc[t][nt][1];
      d[(g + 8) * C] = acc[t][nt][2]; d[(g + 8) * C + 1] = acc[t][nt][3];
    }
}
template <int NTH>
__device__ __forceinline__ void wgrad_sum_cta(int tid, const uint8_t* smem, float* __restrict__ partial, float out_scale) {
  const float* s_red = reinterpret_cast<const float*>(smem);
  float* dst = partial + (size_t)blockIdx.x * 2304;
  for (int i = tid; i < 2304; i += NTH) {
    float a = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) a += s_red[w8 * 2304 + i];
    dst[i] = a * out_scale;
  }
}

// the tap shifts read one pixel before / after the activation planes: those products meet a zero gradient, but
// 0 * (NaN bit pattern left in shared memory) would still poison the sum, so the slack is zeroed once
__device__ __forceinline__ void wgrad_zero_slack(int t /* 0 .. 63 */, uint32_t buf, int a_plane) {
  const int pl = t / (4 * SLACK_PX), k = t % (4 * SLACK_PX);   // 16-byte chunks: 2 per pixel
  const uint32_t base = buf + pl * a_plane;
  const uint32_t off = k < 2 * SLACK_PX ? (uint32_t)k * 16u : (uint32_t)(a_plane - (4 * SLACK_PX - k) * 16);
  sts128(base + off, make_uint4(0u, 0u, 0u, 0u));
}

constexpr int WG_A_PLANE = ((WG_TH + 2) * RW + 2 * SLACK_PX) * PX_BYTES, WG_G_PLANE = (WG_TH * RW + 2 * SLACK_PX) * PX_BYTES;
constexpr int WG_BUF = 2 * WG_A_PLANE + 2 * WG_G_PLANE;   // hi + lo planes of A and of G: 108.5 KB

// ONE 16-warp CTA per SM, warps 0-7 multiply, warps 8-15 stage the next tile into the other of two shared-memory
// buffers, so global-load latency and the hi / lo split hide behind the MMAs of the previous tile instead of alternating
// with them (a first version, two 8-warp CTAs per SM that staged / multiplied in turn, left the tensor pipe 47 % active
// with half of the stall samples in the staging phase and at its two barriers -- profiles/r01i_train_ncu.md; this one:
// 56 %, profiles/r02_wgrad_ncu.md).  Hand-over through named barriers: FULL[b] (256 loader arrivals +
// 256 consumer waits), FREE[b] (the other way round, only when the CTA has another tile for that buffer).
constexpr int WS_NT = 512;
__device__ __forceinline__ void nbar_sync(int id, int cnt) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(cnt) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int cnt) { asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(cnt) : "memory"); }

__global__ void __launch_bounds__(WS_NT, 1)
wgrad3x3_x3_ws_kernel(const float* __restrict__ A, const float* __restrict__ G, float* __restrict__ partial, int n, int h, int wd,
                      int tiles_x, int tiles_y, float g_scale, float out_scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
  constexpr int BAR_FULL = 1, BAR_FREE = 3, BAR_MMA_WARPS = 5;
  if (tid < 2 * (2 * 2 * SLACK_PX * 2)) wgrad_zero_slack(tid & 63, s0 + (tid >> 6) * WG_BUF, WG_A_PLANE);
  __syncthreads();
  const int ntiles = tiles_x * tiles_y;
  if (warp < 8) {   // ---- consumers: the accumulators live in this branch only, the loaders keep their registers for loads
    float acc[9][2][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[t][nt][k] = 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t sb = s0 + b * WG_BUF;
      const uint32_t aH = sb + SLACK_PX * PX_BYTES, aL = aH + WG_A_PLANE;
      const uint32_t gH = sb + 2 * WG_A_PLANE + SLACK_PX * PX_BYTES, gL = gH + WG_G_PLANE;
      nbar_sync(BAR_FULL + b, WS_NT);
      wgrad_mma_tile(acc, warp, lane, aH, aL, gH, gL);
      if (tile + 2 * (int)gridDim.x < ntiles) nbar_arrive(BAR_FREE + b, WS_NT);   // the loaders may refill this buffer
    }
    nbar_sync(BAR_MMA_WARPS, WS_NT / 2);   // every consumer warp is done reading the tile buffers (all tiles are staged by then)
    wgrad_store_acc(acc, warp, lane, smem);
  } else {          // ---- loaders
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t sb = s0 + b * WG_BUF;
      const uint32_t aH = sb + SLACK_PX * PX_BYTES, aL = aH + WG_A_PLANE;
      const uint32_t gH = sb + 2 * WG_A_PLANE + SLACK_PX * PX_BYTES, gL = gH + WG_G_PLANE;
      if (it >= 2) nbar_sync(BAR_FREE + b, WS_NT);
      const int tx = tile % tiles_x, ty = tile / tiles_x;
      wgrad_stage_tile<WS_NT / 2, 7, 6>(tid - WS_NT / 2, aH, aL, gH, gL, A, G, n, h, wd, ty * WG_TH - 1, tx * (RW - 2) - 1, g_scale);
      __threadfence_block();
      nbar_arrive(BAR_FULL + b, WS_NT);
    }
  }
  __syncthreads();
  wgrad_sum_cta<WS_NT>(tid, smem, partial, out_scale);
}

// ------------------------------------------------------------------------------------------------------------------
// Base-conv wgrad (k0 = 3) on tensor cores: dWb[tap][c3][co] = sum_p xn[p + tap][c3] * G[p][co], xn = clip(x,0,255)/255 - 0.5
// inside the image and 0 outside (zero padding of the NORMALISED tensor; utilities.py:449-461, backbone_resnet.py:258-262).
// The same K-over-pixels scheme as the 3x3 wgrad above: the image tile sits in shared memory as one 16-byte chunk per
// pixel (r, g, b, 0 x 5; fp16 hi plane + lo plane), so ldmatrix.trans delivers [8 "channels" x 8 pixels] blocks and an
// m16 tile of the MMA is TWO taps (rows 0-7 / 8-15, three rows of each valid): 5 m-tiles x 2 n-tiles x 3 products = 30
// HMMA per 16 pixels.  Tile 64 x 8 pixels, warp = one tile row; per-CTA partials in fixed order (deterministic).
// (The FFMA kernel it replaces, train.cu::wgrad_base_kernel, took 199 us per step at 32 x 256 x 256: 8x its HBM floor.)
// ------------------------------------------------------------------------------------------------------------------
constexpr int WB3_W = 64, WB3_H = 8, WB3_AW = WB3_W + 2, WB3_AH = WB3_H + 2;
constexpr int WB3_A_PLANE = WB3_AH * WB3_AW * 16, WB3_G_PLANE = WB3_H * WB3_W * PX_BYTES;
constexpr int WB3_SMEM = 2 * WB3_A_PLANE + 2 * WB3_G_PLANE;   // 53.9 KB; the cross-warp sums (8 x 432 floats) reuse it
// One image tile with a one-pixel halo, normalised exactly as the forward pass does (clip(x,0,255)/255 - 0.5, 0 outside the
// image = the zero padding of the NORMALISED tensor), scaled by 64, as ONE 16-byte chunk per pixel (r, g, b, 0 x 5) in an
// fp16 hi plane and a lo plane: rows of ldmatrix (.trans for the wgrad: K = pixels; plain for the conv: M = pixels).
__device__ __forceinline__ void stage_image_tile(int tid, uint32_t aH, uint32_t aL, const float* __restrict__ i_b, int x0, int y0, int h, int wd) {
  for (int i = tid; i < WB3_AH * WB3_AW; i += NT) {
    const int ly = i / WB3_AW, lx = i - ly * WB3_AW;
    const int gx = x0 + lx - 1, gy = y0 + ly - 1;
    float v[3] = {0.f, 0.f, 0.f};
    if (gx >= 0 && gx < wd && gy >= 0 && gy < h) {
      const float* px = i_b + ((size_t)gy * wd + gx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = __fsub_rn(__fdiv_rn(fminf(fmaxf(px[c], 0.f), 255.f), 255.f), 0.5f) * 64.f;
    }
    uint4 hi, lo;
    hi.x = pack_h2(v[0], v[1]); hi.y = pack_h2(v[2], 0.f); hi.z = 0u; hi.w = 0u;
    float2 f = unpack_h2(hi.x);
    lo.x = pack_h2(v[0] - f.x, v[1] - f.y);
    f = unpack_h2(hi.y);
    lo.y = pack_h2(v[2] - f.x, 0.f); lo.z = 0u; lo.w = 0u;
    sts128(aH + i * 16, hi);
    sts128(aL + i * 16, lo);
  }
}
__global__ void __launch_bounds__(NT, 4)
wgrad_base3_x3_kernel(const float* __restrict__ img, const float* __restrict__ G, float* __restrict__ partial, int n, int h, int wd,
                      int tiles_x, int tiles_y, float g_scale, float out_scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t aH = s0, aL = s0 + WB3_A_PLANE, gH = s0 + 2 * WB3_A_PLANE, gL = gH + WB3_G_PLANE;
  float acc[5][2][4];
#pragma unroll
  for (int j = 0; j < 5; ++j)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[j][nt][k] = 0.f;
  // ldmatrix.trans row addresses (see wgrad_mma_tile): lane -> matrix (lane >> 3), row (lane & 7)
  const int mi = lane >> 3, ri = lane & 7;
  const int gp = (lane & 7) + ((lane >> 3) & 1) * 8, gh = (lane >> 4) & 1;
  const int ntiles = tiles_x * tiles_y * n;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int x0 = tx * WB3_W, y0 = ty * WB3_H;
    const float* i_b = img + (size_t)b * h * wd * 3;
    const float* g_b = G + (size_t)b * h * wd * C;
    __syncthreads();   // the previous tile's fragments are read
    stage_image_tile(tid, aH, aL, i_b, x0, y0, h, wd);
    for (int i = tid; i < WB3_H * WB3_W * 2; i += NT) {
      const int hf = i & 1, pix = i >> 1;
      const int ly = pix / WB3_W, lx = pix - ly * WB3_W;
      const int gx = x0 + lx, gy = y0 + ly;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bq = a;
      if (gx < wd && gy < h) {
        const float4* sp = reinterpret_cast<const float4*>(g_b + ((size_t)gy * wd + gx) * C + 8 * hf);
        a = sp[0]; bq = sp[1];
        a.x *= g_scale; a.y *= g_scale; a.z *= g_scale; a.w *= g_scale; bq.x *= g_scale; bq.y *= g_scale; bq.z *= g_scale; bq.w *= g_scale;
      }
      split_store(gH + px_off(pix, hf), gL + px_off(pix, hf), a, bq);
    }
    __syncthreads();
    const int r = warp;   // tile row of this warp
#pragma unroll 1
    for (int ks = 0; ks < WB3_W / 16; ++ks) {
      uint32_t bh[4], bl[4];
      const int gpix = r * WB3_W + ks * 16 + gp;
      ldsm4t(bh, gH + px_off(gpix, gh));
      ldsm4t(bl, gL + px_off(gpix, gh));
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        // matrices 0 / 2: tap 2j, pixels 0-7 / 8-15; matrices 1 / 3: tap 2j + 1 (j = 4: tap 8 again, its rows are dropped)
        const int t = min(2 * j + (mi & 1), 8), dy = t / 3, dx = t - 3 * dy;
        const uint32_t off = (uint32_t)(((r + dy) * WB3_AW + ks * 16 + ri + 8 * (mi >> 1) + dx) * 16);
        uint32_t ah[4], al[4];
        ldsm4t(ah, aH + off);
        ldsm4t(al, aL + off);
        mma16816(acc[j][0], al, make_uint2(bh[0], bh[1]));
        mma16816(acc[j][1], al, make_uint2(bh[2], bh[3]));
        mma16816(acc[j][0], ah, make_uint2(bl[0], bl[1]));
        mma16816(acc[j][1], ah, make_uint2(bl[2], bl[3]));
        mma16816(acc[j][0], ah, make_uint2(bh[0], bh[1]));
        mma16816(acc[j][1], ah, make_uint2(bh[2], bh[3]));
      }
    }
  }
  // the 8 warps' accumulators in fixed order through shared memory: [warp][tap 9][c3 3][co 16]
  __syncthreads();
  float* s_red = reinterpret_cast<float*>(smem);
  {
    const int g = lane >> 2, q = lane & 3;
    if (g < 3) {
#pragma unroll
      for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          float* d0 = s_red + warp * 432 + ((2 * j) * 3 + g) * C + nt * 8 + 2 * q;
          d0[0] = acc[j][nt][0]; d0[1] = acc[j][nt][1];
          if (j < 4) {
            float* d1 = s_red + warp * 432 + ((2 * j + 1) * 3 + g) * C + nt * 8 + 2 * q;
            d1[0] = acc[j][nt][2]; d1[1] = acc[j][nt][3];
          }
        }
    }
  }
  __syncthreads();
  float* dst = partial + (size_t)blockIdx.x * 432;
  for (int i = tid; i < 432; i += NT) {
    float a = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) a += s_red[w8 * 432 + i];
    dst[i] = a * out_scale;
  }
}

// Base conv (k0 = 3) of the training forward on the same tile: out[p][co] = sum_{tap, c3} xn[p + tap][c3] * W[tap][c3][co], fp32
// NHWC16 out.  M = 16 pixels of a tile row, K = two taps (8 padded channels each), N = 16: 5 k-tiles x 2 n-tiles x 3
// products per 16 pixels; the weight fragments (hi + lo, scaled by W_SCALE) live in registers for the whole kernel.
// (conv_f32.cu::base_conv_kernel<float>, the FFMA kernel it replaces in the step, normalised every pixel nine times:
// 134 us at 32 x 256 x 256 against an HBM floor of 25 us.)
__global__ void __launch_bounds__(NT, 3)
base_conv3_x3_kernel(const float* __restrict__ img, float* __restrict__ out, const float* __restrict__ w /*[9][3][16]*/, int n, int h, int wd,
                     int tiles_x, int tiles_y) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t aH = s0, aL = s0 + WB3_A_PLANE;
  const int g = lane >> 2, q = lane & 3;
  // B fragments: k-tile j = taps (2j, 2j + 1), b0 = k (2q, 2q + 1) of tap 2j, b1 = the same of tap 2j + 1; n = nt * 8 + g
  uint32_t wh[5][2][2], wl[5][2][2];
#pragma unroll
  for (int j = 0; j < 5; ++j)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        const int t = 2 * j + hb, co = nt * 8 + g;
        float v0 = 0.f, v1 = 0.f;
        if (t < 9) {
          if (2 * q < 3) v0 = w[(t * 3 + 2 * q) * C + co] * W_SCALE;
          if (2 * q + 1 < 3) v1 = w[(t * 3 + 2 * q + 1) * C + co] * W_SCALE;
        }
        const uint32_t hh = pack_h2(v0, v1);
        const float2 f = unpack_h2(hh);
        wh[j][nt][hb] = hh;
        wl[j][nt][hb] = pack_h2(v0 - f.x, v1 - f.y);
      }
  // ldmatrix (plain) row addresses: matrices (px 0-7, tap 2j), (px 8-15, tap 2j), (px 0-7, tap 2j + 1), (px 8-15, tap 2j + 1)
  const int mi = lane >> 3, ri = lane & 7;
  const float out_scale = 1.0f / (64.f * W_SCALE);
  const int ntiles = tiles_x * tiles_y * n;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int x0 = tx * WB3_W, y0 = ty * WB3_H;
    __syncthreads();
    stage_image_tile(tid, aH, aL, img + (size_t)b * h * wd * 3, x0, y0, h, wd);
    __syncthreads();
    const int r = warp, gy = y0 + r;
    if (gy >= h) continue;
    float* o_row = out + ((size_t)b * h + gy) * wd * C;
#pragma unroll 1
    for (int ks = 0; ks < WB3_W / 16; ++ks) {
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int t = min(2 * j + (mi >> 1), 8), dy = t / 3, dx = t - 3 * dy;   // j = 4: tap 9 does not exist, its weights are zero
        const uint32_t off = (uint32_t)(((r + dy) * WB3_AW + ks * 16 + ri + 8 * (mi & 1) + dx) * 16);
        uint32_t ah[4], al[4];
        ldsm4(ah, aH + off);
        ldsm4(al, aL + off);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          mma16816(acc[nt], al, make_uint2(wh[j][nt][0], wh[j][nt][1]));
          mma16816(acc[nt], ah, make_uint2(wl[j][nt][0], wl[j][nt][1]));
          mma16816(acc[nt], ah, make_uint2(wh[j][nt][0], wh[j][nt][1]));
        }
      }
      // c0, c1 = (pixel g, co 2q, 2q + 1), c2, c3 = (pixel g + 8, ...)
      const int gx0 = x0 + ks * 16;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        if (gx0 + g < wd) *reinterpret_cast<float2*>(o_row + (size_t)(gx0 + g) * C + nt * 8 + 2 * q) = make_float2(acc[nt][0] * out_scale, acc[nt][1] * out_scale);
        if (gx0 + g + 8 < wd) *reinterpret_cast<float2*>(o_row + (size_t)(gx0 + g + 8) * C + nt * 8 + 2 * q) = make_float2(acc[nt][2] * out_scale, acc[nt][3] * out_scale);
      }
    }
  }
}

}  // namespace x3

int launch_conv3x3_x3(bfcnn_handle* h, const float* in, float* out, const float* w, const float* res, double* stats,
                      ConvEpi epi, const Extent& e, float in_scale, cudaStream_t st) {
  using namespace x3;
  const int tiles_y = (e.he + TH_MAX - 1) / TH_MAX;
  const int th = (e.he + tiles_y - 1) / tiles_y;
  const int tiles_x = (e.we + (RW - 2) - 1) / (RW - 2);
  const size_t smem = (size_t)2 * ((th + 2) * RW + 2 * SLACK_PX) * PX_BYTES;
  static bool attr_set_dev[64] = {};   // function attributes are per device: one flag per device ordinal
  bool& attr_set = attr_set_dev[h->device & 63];
  if (!attr_set) {
    const int mx = 2 * ((TH_MAX + 2) * RW + 2 * SLACK_PX) * PX_BYTES;
    BF_CUDA(cudaFuncSetAttribute((const void*)conv3x3_x3_kernel<CONV_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    BF_CUDA(cudaFuncSetAttribute((const void*)conv3x3_x3_kernel<CONV_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    BF_CUDA(cudaFuncSetAttribute((const void*)conv3x3_x3_kernel<CONV_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    BF_CUDA(cudaFuncSetAttribute((const void*)conv3x3_x3_kernel<CONV_STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    BF_CUDA(cudaFuncSetAttribute((const void*)conv3x3_x3_kernel<CONV_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    attr_set = true;
  }
  const long long grid = (long long)tiles_x * tiles_y * e.n;
  BF_REQUIRE(grid < (1ll << 31), "too many tiles");
  BF_REQUIRE(in_scale > 0.f, "in_scale must be a positive power of two");
  const unsigned g = (unsigned)grid;
  switch (epi) {
    case CONV_PLAIN: conv3x3_x3_kernel<CONV_PLAIN><<<g, NT, smem, st>>>(in, out, w, res, stats, e.he, e.we, th, tiles_x, tiles_y, in_scale, 1.0f / (in_scale * W_SCALE)); break;
    case CONV_RELU: conv3x3_x3_kernel<CONV_RELU><<<g, NT, smem, st>>>(in, out, w, res, stats, e.he, e.we, th, tiles_x, tiles_y, in_scale, 1.0f / (in_scale * W_SCALE)); break;
    case CONV_RESIDUAL: conv3x3_x3_kernel<CONV_RESIDUAL><<<g, NT, smem, st>>>(in, out, w, res, stats, e.he, e.we, th, tiles_x, tiles_y, in_scale, 1.0f / (in_scale * W_SCALE)); break;
    case CONV_STATS: conv3x3_x3_kernel<CONV_STATS><<<g, NT, smem, st>>>(in, out, w, res, stats, e.he, e.we, th, tiles_x, tiles_y, in_scale, 1.0f / (in_scale * W_SCALE)); break;
    case CONV_MASK: conv3x3_x3_kernel<CONV_MASK><<<g, NT, smem, st>>>(in, out, w, res, stats, e.he, e.we, th, tiles_x, tiles_y, in_scale, 1.0f / (in_scale * W_SCALE)); break;
    default: set_error("unsupported conv epilogue"); return BFCNN_ERR_INTERNAL;
  }
  h->launches++;
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

// dW partials of one 3x3 conv: `partial` receives [returned grid][2304]; g_scale (a power of two) lifts the gradients
int launch_wgrad3x3_x3(bfcnn_handle* h, const float* act, const float* grad, float* partial, int max_parts, const Extent& e,
                       float g_scale, int* parts_out, cudaStream_t st) {
  using namespace x3;
  const size_t smem = (size_t)WG_BUF * 2;
  static bool attr_set_dev[64] = {};   // function attributes are per device: one flag per device ordinal
  bool& attr_set = attr_set_dev[h->device & 63];
  if (!attr_set) {
    BF_CUDA(cudaFuncSetAttribute((const void*)wgrad3x3_x3_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * WG_BUF));
    attr_set = true;
  }
  static_assert((size_t)8 * 2304 * 4 <= (size_t)WG_BUF, "cross-warp reduction buffer does not fit");
  static_assert(2 * WG_BUF <= 232448, "two tile buffers do not fit in shared memory");
  const long long vcols = (long long)e.n * (e.we + 1) - 1;   // virtual row: the images side by side, a zero column between them
  BF_REQUIRE(vcols < (1ll << 30), "batch too wide for the virtual row");
  const int tiles_x = (int)((vcols + (RW - 2) - 1) / (RW - 2)), tiles_y = (e.he + WG_TH - 1) / WG_TH;
  const long long ntiles = (long long)tiles_x * tiles_y;
  BF_REQUIRE(ntiles < (1ll << 30), "too many tiles");
  const int grid = (int)std::min<long long>(ntiles, std::min(max_parts, h->sm_count));
  wgrad3x3_x3_ws_kernel<<<grid, WS_NT, smem, st>>>(act, grad, partial, e.n, e.he, e.we, tiles_x, tiles_y, g_scale, 1.0f / (64.f * g_scale));
  h->launches++;
  BF_CUDA(cudaGetLastError());
  *parts_out = grid;
  return BFCNN_OK;
}

// dWb partials of the k0 = 3 base conv: `partial` receives [returned grid][432]
int launch_wgrad_base3_x3(bfcnn_handle* h, const float* img, const float* grad, float* partial, int max_parts, const Extent& e,
                          float g_scale, int* parts_out, cudaStream_t st) {
  using namespace x3;
  static bool attr_set_dev[64] = {};
  bool& attr_set = attr_set_dev[h->device & 63];
  if (!attr_set) {
    BF_CUDA(cudaFuncSetAttribute((const void*)wgrad_base3_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WB3_SMEM));
    attr_set = true;
  }
  static_assert(8 * 432 * 4 <= WB3_SMEM, "cross-warp reduction buffer does not fit");
  const int tiles_x = (e.we + WB3_W - 1) / WB3_W, tiles_y = (e.he + WB3_H - 1) / WB3_H;
  const long long ntiles = (long long)tiles_x * tiles_y * e.n;
  BF_REQUIRE(ntiles < (1ll << 30), "too many tiles");
  const int grid = (int)std::min<long long>(ntiles, std::min(max_parts, 4 * h->sm_count));
  wgrad_base3_x3_kernel<<<grid, NT, WB3_SMEM, st>>>(img, grad, partial, e.n, e.he, e.we, tiles_x, tiles_y, g_scale, 1.0f / (64.f * g_scale));
  h->launches++;
  BF_CUDA(cudaGetLastError());
  *parts_out = grid;
  return BFCNN_OK;
}

// training forward: normalise + base conv k0 = 3 from the fp32 image [n,h,w,3] into the fp32 NHWC16 map
int launch_base_conv3_x3(bfcnn_handle* h, const float* img, float* out, const float* w, const Extent& e, cudaStream_t st) {
  using namespace x3;
  static bool attr_set_dev[64] = {};
  bool& attr_set = attr_set_dev[h->device & 63];
  if (!attr_set) {
    BF_CUDA(cudaFuncSetAttribute((const void*)base_conv3_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * WB3_A_PLANE));
    attr_set = true;
  }
  BF_REQUIRE(e.he == e.h && e.we == e.w, "training maps have no canvas band");
  const int tiles_x = (e.we + WB3_W - 1) / WB3_W, tiles_y = (e.he + WB3_H - 1) / WB3_H;
  const long long ntiles = (long long)tiles_x * tiles_y * e.n;
  BF_REQUIRE(ntiles < (1ll << 30), "too many tiles");
  const int grid = (int)std::min<long long>(ntiles, 6ll * h->sm_count);
  base_conv3_x3_kernel<<<grid, NT, 2 * WB3_A_PLANE, st>>>(img, out, w, e.n, e.he, e.we, tiles_x, tiles_y);
  h->launches++;
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

}  // namespace bfcnn
