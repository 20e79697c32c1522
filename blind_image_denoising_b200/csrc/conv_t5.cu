// conv_t5.cu -- the training-step 3x3 16->16 convolution on tcgen05 at FP32-grade accuracy (F16X3 arithmetic).
//
// Same contract as conv3x3_x3_kernel (conv_x3.cu): fp32 NHWC16 in, fp32 NHWC16 out, zero "same" padding, fused epilogue
// (ReLU / residual add / per-channel batch statistics / ReLU mask for the backward pass), activations and weights split
// on the fly into fp16 hi + lo parts, every product issued as lo*hi + hi*lo + hi*hi with fp32 accumulation.  The engine
// is the row-streaming pipeline of fused_stream.cu reduced to ONE layer:
//
//   * a CTA owns a 128-lane column strip segment (126 output columns + one halo column per side) and streams down it in
//     steps of G = 4 rows;
//   * two groups of 8 converter warps take alternate steps: a thread owns one channel half of one pixel, issues the
//     256-bit loads of its 4 rows BEFORE it waits for the ring slot (two steps of global loads in flight), scales, splits
//     into hi / lo and stores the fp16 channel-half planes (SWIZZLE_NONE K-major UMMA layout, a dx tap is a descriptor
//     shift) into a ring of 4 groups;
//   * one elected thread issues, per input row, 9 MMAs (3 dx x {lo*hi, hi*lo, hi*hi}, M128 N48 K16): the N = 48 dy-scatter
//     accumulates the row into the TMEM blocks of output rows q-1, q, q+1 (block = row & 31; MMAs whose blocks straddle
//     the end of the 32-block ring are split, as in fused_stream.cu); its mbarrier waits are done by a helper warp;
//   * 8 epilogue warps (two row quarters each per step) drain the finished rows: scale, epilogue op, two 256-bit stores
//     per pixel (16-byte stores at a 64-byte stride were the critical path: 115 -> 72 us with full-sector accesses).
//
// Per 128-pixel row the shared-memory data pipe moves 9 x 5.5 KB of operands + 8 KB of plane stores (~450 cycles), HBM
// moves 128 x (64 + 64 [+ 64]) bytes (HBM floor 41-62 us at 2.1 MP); measured 72 us (conv_x3.cu: 117-181 us).
//
// Fused prologue (PRO = 1): the converters build the conv input from TWO maps, x = ca[ch] * in + cb[ch] * in2 + cc[ch], write
// it to out2 (the map the rest of the step needs: X_{i+1} in the forward pass, dU in the backward pass) and feed the MMAs
// from registers.  With the per-channel coefficients of train.cu::bn_coef_kernel this is
//   forward   X_{i+1} = X_i + (U_i - mean) * gamma / sigma                   (BN(center=False) + Add, backbone_blocks.py:191-242)
//   backward  dU_i = gamma / sigma * (dY - mean(dY) - xhat * mean(dY * xhat))  (BN batch-statistics backward)
// so bn_residual_kernel / bn_bwd_apply_kernel (a 192 B/pixel HBM round trip each) disappear from the step: the fused conv
// moves 64 B/pixel more than the plain one.  A fused step converts its 4 rows in two batches of 2 (both operands of 4 rows
// would not fit the 78-register budget of the 26-warp CTA).
//
// Reference: the forward convs of backbone_blocks.py:167-246 in training mode and the dgrad convs of
// train_loop.py:302-304 (a correlation of dOut with the flipped, transposed kernel, prepared by the caller).
#include "stream_common.cuh"

namespace bfcnn {
namespace t5 {

using namespace tc5;
using stream::RW; using stream::SLACK_PX; using stream::Split; using stream::Seg; using stream::seg_at; using stream::cta_rows;

constexpr int G = 4;                    // rows per step
constexpr int EPI_WARPS = 8;            // G rows x 4 TMEM lane quarters = 16 tasks per step, two per warp
constexpr int WARP_MMA = 8, WARP_HELP = 9, WARP_CVT = 10, CVT_WARPS = 8;   // converter warps per group: thread = (pixel, half)
constexpr int CVT_GROUPS = 2;           // converter groups take alternate steps; each issues its loads BEFORE it waits for
                                        // the ring slot, so two steps of global loads are in flight
constexpr int NTHREADS = 32 * (WARP_CVT + CVT_GROUPS * CVT_WARPS);
constexpr int KIN = 4;                  // input ring: groups of G rows
constexpr int PLANE_BYTES = (KIN * G * RW + 2 * SLACK_PX) * 16;   // one channel-half plane of the hi or lo part
constexpr int W_PART_BYTES = 3 * 48 * 16 * 2;                     // B operand of one part: [dx 3][N 48][K 16] fp16
constexpr float W_SCALE = 256.f;        // as conv_x3.cu: keeps the low part of the weights out of the fp16 subnormals
constexpr uint32_t BAR_MMA = 0, BAR_EPI = 2, BAR_FULL = 4, BAR_FREE = 4 + KIN, NBARS = 4 + 2 * KIN;
constexpr uint32_t SM_BARS = 0, SM_TMEM = 256, SM_STAT = 384, SM_WTS = 512, SM_PLANES = SM_WTS + 2 * W_PART_BYTES;   // 9728
constexpr int SMEM_BYTES = SM_PLANES + 4 * PLANE_BYTES;
constexpr int MIN_SHARE = 16;
constexpr int SEG_OVERHEAD = 2 + 2 * G;   // halo rows + pipeline fill / drain of a segment, in rows (cost-space split)

struct Params : Split {   // rows_needed = h, seg_overhead = SEG_OVERHEAD
  const float* in;
  float* out;
  const float* w;       // [9][16 cin][16 cout] fp32
  const float* res;     // residual / mask source (fp32 NHWC16) or nullptr
  const float* in2;     // PRO = 1: second input map
  float* out2;          // PRO = 1: the combined input is written here
  const float* coef;    // PRO = 1: [3][16] per-channel ca, cb, cc
  uint16_t* mask_out;   // CONV_RELU: bit c of pixel p = (output channel c > 0), or nullptr
  const uint16_t* mask_in;   // CONV_MASK: that map instead of `res` (2 B/pixel instead of 64), or nullptr
  double* stats;        // [32]: per-channel sum, sum of squares (CONV_STATS)
  int n, h, wd;
  float in_scale, out_scale;
};

// hi / lo split of 8 consecutive channels (one 16-byte chunk of each part)
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
  hi.x = pack_h2(a.x, a.y); hi.y = pack_h2(a.z, a.w); hi.z = pack_h2(b.x, b.y); hi.w = pack_h2(b.z, b.w);
  float2 f;
  f = unpack_h2(hi.x); lo.x = pack_h2(a.x - f.x, a.y - f.y);
  f = unpack_h2(hi.y); lo.y = pack_h2(a.z - f.x, a.w - f.y);
  f = unpack_h2(hi.z); lo.z = pack_h2(b.x - f.x, b.y - f.y);
  f = unpack_h2(hi.w); lo.w = pack_h2(b.z - f.x, b.w - f.y);
}

template <int EPI, int PRO>
__global__ void __launch_bounds__(NTHREADS, 1)
conv3x3_t5_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t s0 = smem_u32(smem);
  const uint32_t bars = s0 + SM_BARS;
  float* s_stat = reinterpret_cast<float*>(smem + SM_STAT);   // [32]
  // planes: hi half 0, hi half 1, lo half 0, lo half 1; pixel 0 of ring row 0 sits SLACK_PX pixels into each plane
  const uint32_t pl0 = s0 + SM_PLANES + SLACK_PX * 16;
  long long r0, r1;
  cta_rows(p, r0, r1);

  // ---------------- setup: barriers, TMEM, weights (fp32 -> scaled hi / lo in the UMMA B layout), zeroed planes
  if (tid < (int)NBARS) {
    const uint32_t cnt = tid < (int)BAR_EPI ? 1u : (tid < (int)BAR_FULL ? (uint32_t)EPI_WARPS : (tid < (int)BAR_FREE ? (uint32_t)CVT_WARPS : 1u));
    mbar_init(bars + tid * 8, cnt);
  }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(s0 + SM_TMEM) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid < 2 * C) s_stat[tid] = 0.f;
  {
    __half* wh = reinterpret_cast<__half*>(smem + SM_WTS);
    __half* wl = wh + W_PART_BYTES / 2;
    for (int i = tid; i < 3 * 48 * 16; i += NTHREADS) {
      const int dxi = i / 768, r = i - dxi * 768, nn = r >> 4, k = r & 15;   // (dx, n = j*16 + cout, k = cin)
      const int j = nn >> 4, co = nn & 15, tap = (2 - j) * 3 + dxi;            // block j <-> dy = 1 - j (host_pack.cu)
      const float v = p.w[(tap * C + k) * C + co] * W_SCALE;
      const __half hv = __float2half_rn(v);
      const int off = dxi * 768 + (k >> 3) * 384 + (nn >> 3) * 64 + (nn & 7) * 8 + (k & 7);
      wh[off] = hv;
      wl[off] = __float2half_rn(v - __half2float(hv));
    }
    for (uint32_t i = tid; i < (uint32_t)(4 * PLANE_BYTES) / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem + SM_PLANES)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM);
  // Programmatic dependent launch: the setup above (barriers, TMEM, weight split, zeroed planes: a few microseconds, 24
  // convs per step) read only the weights, which were final before the PREVIOUS kernel started; it overlaps that kernel when
  // it is one of the small ones that release their dependents at once (bn_finalize / bn_bwd_coef / wgrad_reduce).
  pdl_wait_for_previous();

  if (warp < EPI_WARPS) {
    // ================= epilogue warps =================
    const int quarter = warp & 3, rsel0 = warp >> 2;   // rows rsel0 and rsel0 + 2 of the group
    const int c = quarter * 32 + lane;
    const uint32_t tq = tmem + ((uint32_t)(quarter * 32) << 16);
    for (int blk = rsel0; blk < 32; blk += 2) tmem_zero16(tq + blk * 16);
    tmem_wait_st();
    tc_fence_before();
    asm volatile("bar.sync 1, %0;\n" ::"r"(32 * (EPI_WARPS + 1)) : "memory");
    float ssum[C], ssq[C];
    if (EPI == CONV_STATS) {
#pragma unroll
      for (int i = 0; i < C; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
    }
    uint32_t S = 0;
    for (long long a = r0; a < r1;) {
      const Seg sg = seg_at(p, a, r1);
      a += sg.yb - sg.ya;
      const int P = (sg.yb - sg.ya) + 2, Gm = (P + G - 1) / G, nsteps = Gm + 1;
      // column of the VIRTUAL row (all images side by side, one zero column between them): image vb, column gx
      const int y00 = sg.ya - 1, vx = sg.j * (RW - 2) - 1 + c;
      const int vb = vx >= 0 ? vx / (p.wd + 1) : -1, gx = vx - vb * (p.wd + 1);
      const bool col_out = (c >= 1) && (c < RW - 1) && (vb >= 0) && (vb < p.n) && (gx < p.wd);
      const long long px0 = ((long long)vb * p.h + y00) * p.wd + gx;   // pixel index of (vb, y00, gx)
      for (int sr = 0; sr < nsteps; ++sr, ++S) {
        const int w = sr - 1;
        // the residual / mask operand of BOTH rows of this step is fetched before the wait for the MMAs (its address does
        // not depend on them): the global-load latency hides behind the wait instead of sitting on the epilogue's path
        float4 rva[2][4];
        uint32_t mva[2] = {0u, 0u};
        const bool bitmask = (EPI == CONV_MASK) && p.mask_in != nullptr;   // warp-uniform
        if (EPI == CONV_RESIDUAL || EPI == CONV_MASK) {
#pragma unroll
          for (int k2 = 0; k2 < 2; ++k2) {
            const int rho = G * w + rsel0 + 2 * k2;
            if (w >= 0 && col_out && rho >= 1 && rho < P - 1) {
              if (bitmask) {
                mva[k2] = p.mask_in[px0 + (long long)rho * p.wd];
              } else {
                const long long o = (px0 + (long long)rho * p.wd) * C;
                ldg256(p.res + o, rva[k2][0], rva[k2][1]);
                ldg256(p.res + o + 8, rva[k2][2], rva[k2][3]);
              }
            }
          }
        }
        mbar_wait_sleep(bars + (BAR_MMA + (S & 1u)) * 8, (S >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2) {
          if (w < 0) break;
          const int rsel = rsel0 + 2 * k2;
          const int rho = G * w + rsel;
          const uint32_t taddr = tq + (uint32_t)(rho & 31) * 16u;
          uint32_t v[16];
          tmem_ld16_issue(taddr, v);
          const bool ok = col_out && rho >= 1 && rho < P - 1;
          const long long o = (px0 + (long long)rho * p.wd) * C;
          const float4 (&rv)[4] = rva[k2];
          tmem_ld_wait(v);
          tmem_zero16(taddr);
          if (ok) {
            float4 fo[4];
            uint32_t mbits = 0u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 f = make_float4(__uint_as_float(v[4 * q]) * p.out_scale, __uint_as_float(v[4 * q + 1]) * p.out_scale,
                                     __uint_as_float(v[4 * q + 2]) * p.out_scale, __uint_as_float(v[4 * q + 3]) * p.out_scale);
              if (EPI == CONV_STATS) {
                ssum[4 * q] += f.x; ssum[4 * q + 1] += f.y; ssum[4 * q + 2] += f.z; ssum[4 * q + 3] += f.w;
                ssq[4 * q] += f.x * f.x; ssq[4 * q + 1] += f.y * f.y; ssq[4 * q + 2] += f.z * f.z; ssq[4 * q + 3] += f.w * f.w;
              }
              if (EPI == CONV_RELU) {
                f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f);
                mbits |= ((f.x > 0.f ? 1u : 0u) | (f.y > 0.f ? 2u : 0u) | (f.z > 0.f ? 4u : 0u) | (f.w > 0.f ? 8u : 0u)) << (4 * q);
              }
              if (EPI == CONV_RESIDUAL) { f.x += rv[q].x; f.y += rv[q].y; f.z += rv[q].z; f.w += rv[q].w; }
              if (EPI == CONV_MASK) {   // ReLU backward: pass the gradient where the saved activation is > 0
                if (bitmask) {
                  const uint32_t mq = mva[k2] >> (4 * q);
                  f.x = (mq & 1u) ? f.x : 0.f; f.y = (mq & 2u) ? f.y : 0.f; f.z = (mq & 4u) ? f.z : 0.f; f.w = (mq & 8u) ? f.w : 0.f;
                } else {
                  f.x = rv[q].x > 0.f ? f.x : 0.f; f.y = rv[q].y > 0.f ? f.y : 0.f;
                  f.z = rv[q].z > 0.f ? f.z : 0.f; f.w = rv[q].w > 0.f ? f.w : 0.f;
                }
              }
              fo[q] = f;
            }
            stg256(p.out + o, fo[0], fo[1]);
            stg256(p.out + o + 8, fo[2], fo[3]);
            if (EPI == CONV_RELU && p.mask_out != nullptr) p.mask_out[px0 + (long long)rho * p.wd] = (uint16_t)mbits;
          }
          tmem_wait_st();
          tc_fence_before();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + (BAR_EPI + (S & 1u)) * 8);
      }
    }
    if (EPI == CONV_STATS) {
      // per-channel sums of this thread's pixels -> warp -> CTA (shared float atomics) -> global (double atomics)
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        float a = ssum[ch], s2 = ssq[ch];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, m); s2 += __shfl_xor_sync(0xffffffffu, s2, m); }
        if (lane == 0) { atomicAdd(&s_stat[ch], a); atomicAdd(&s_stat[C + ch], s2); }
      }
    }
  } else if (warp == WARP_MMA) {
    // ================= MMA issuer =================
    asm volatile("bar.sync 1, %0;\n" ::"r"(32 * (EPI_WARPS + 1)) : "memory");
    const uint32_t idesc0 = make_idesc_f16(128, 0);
    const uint32_t a_hi = desc_lo(pl0, PLANE_BYTES) - 1u;          // row 0, pixel -1 of the hi part; + RW per ring row
    const uint32_t a_lo = a_hi + (uint32_t)(2 * PLANE_BYTES / 16);
    const uint32_t b_hi = desc_lo(s0 + SM_WTS, 48 * 16), b_lo = b_hi + (uint32_t)(W_PART_BYTES / 16);
    constexpr uint32_t BDX = 48 * 16 * 2 / 16;
    uint32_t S = 0;
    for (long long a = r0; a < r1;) {
      const Seg sg = seg_at(p, a, r1);
      a += sg.yb - sg.ya;
      const int P = (sg.yb - sg.ya) + 2, Gm = (P + G - 1) / G, nsteps = Gm + 1;
      for (int sr = 0; sr < nsteps; ++sr, ++S) {
        asm volatile("bar.sync %0, 64;\n" ::"r"(2 + (S & 1u)) : "memory");   // the helper has seen this step's barriers
        tc_fence_after();
        if (elect_one_sync()) {
          if (sr < Gm) {
            const uint32_t slot_row = (uint32_t)((S % KIN) * G);   // ring rows of this step's group (S counts groups too: see below)
#pragma unroll 1
            for (int i = 0; i < G; ++i) {
              const int rho = G * sr + i;
              if (rho >= P) break;
              const uint32_t arow = (slot_row + (uint32_t)i) * RW;
              const int jlo = (rho == 0) ? 1 : 0, jhi = (rho == P - 1) ? 1 : 2;
              const int blk_lo = (rho - 1 + jlo) & 31, nb = jhi - jlo + 1;
              const int n1 = min(nb, 32 - blk_lo);
#pragma unroll 1
              for (int part = 0; part < 2; ++part) {   // blocks [blk_lo, blk_lo + n1), then the wrapped rest at column 0
                const int nblk = part == 0 ? n1 : nb - n1;
                if (nblk <= 0) break;
                const uint32_t d = tmem + (part == 0 ? (uint32_t)blk_lo * 16u : 0u);
                const uint32_t boff = (uint32_t)((jlo + (part == 0 ? 0 : n1)) * 16);
                const uint32_t id = idesc0 + ((uint32_t)(2 * nblk) << 17);
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                  mma_lo(d, a_lo + arow + dx, b_hi + boff + dx * BDX, id);
                  mma_lo_fill(d, a_hi + arow + dx, b_lo + boff + dx * BDX, id);      // hi*lo and hi*hi share their A operand:
                  mma_lo_lastuse(d, a_hi + arow + dx, b_hi + boff + dx * BDX, id);   // one shared-memory read (A collector)
                }
              }
            }
          }
          umma_commit(bars + (BAR_MMA + (S & 1u)) * 8);
          umma_commit(bars + (BAR_FREE + (S % KIN)) * 8);
                                                             // uses its slot, the epilogue-only one with nothing in it)
        }
        __syncwarp();
      }
    }
  } else if (warp == WARP_HELP) {
    // ================= barrier helper of the MMA issuer =================
    uint32_t S = 0;
    for (long long a = r0; a < r1;) {
      const Seg sg = seg_at(p, a, r1);
      a += sg.yb - sg.ya;
      const int P = (sg.yb - sg.ya) + 2, Gm = (P + G - 1) / G, nsteps = Gm + 1;
      for (int sr = 0; sr < nsteps; ++sr, ++S) {
        if (lane == 0 && S >= 2) mbar_wait_sleep(bars + (BAR_EPI + (S & 1u)) * 8, ((S - 2) >> 1) & 1u);
        if (lane == 1) mbar_wait_sleep(bars + (BAR_FULL + (S % KIN)) * 8, (S / KIN) & 1u);
        if (lane == 2 && sr == 0 && S >= 1) mbar_wait_sleep(bars + (BAR_EPI + ((S - 1) & 1u)) * 8, ((S - 1) >> 1) & 1u);
        __syncwarp();
        tc_fence_before();
        asm volatile("bar.arrive %0, 64;\n" ::"r"(2 + (S & 1u)) : "memory");
      }
    }
  } else {
    // ================= converter warps: fp32 rows -> scaled hi / lo fp16 planes =================
    // The ring slot of a step is S % KIN for EVERY step (the epilogue-only step at the end of a segment arrives with an
    // empty slot), so issuer, helper and converters index barriers and rows by the same global step counter and every
    // barrier completes exactly one phase per KIN steps.
    const int cw = warp - WARP_CVT, cgrp = cw / CVT_WARPS;
    const int ct = (cw % CVT_WARPS) * 32 + lane;
    const int c = ct >> 1, hf = ct & 1;   // pixel column of the strip, channel half
    uint32_t S = 0;
    for (long long a = r0; a < r1;) {
      const Seg sg = seg_at(p, a, r1);
      a += sg.yb - sg.ya;
      const int P = (sg.yb - sg.ya) + 2, Gm = (P + G - 1) / G, nsteps = Gm + 1;
      const int y00 = sg.ya - 1, vx = sg.j * (RW - 2) - 1 + c;
      const int vb = vx >= 0 ? vx / (p.wd + 1) : -1, gx = vx - vb * (p.wd + 1);
      const bool col_in = (vb >= 0) && (vb < p.n) && (gx < p.wd);   // separator columns and the outside are zero ("same" padding)
      const float* in_b = p.in + (long long)vb * p.h * p.wd * C + 8 * hf;
      float ca[8], cb[8], cc[8];   // PRO = 1: this thread's channel half of the per-channel coefficients
      if (PRO == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { ca[k] = p.coef[8 * hf + k]; cb[k] = p.coef[C + 8 * hf + k]; cc[k] = p.coef[2 * C + 8 * hf + k]; }
      }
      const bool col_side = col_in && (c >= 1) && (c < RW - 1);   // columns this strip owns (the halo columns belong to its neighbours)
      for (int sr = 0; sr < nsteps; ++sr, ++S) {
        if ((int)(S % CVT_GROUPS) != cgrp) continue;   // the other converter group's step
        const uint32_t slot = S % KIN;
        if (PRO == 1) {
          // two batches of G / 2 rows, both operands of a batch in flight together; the first batch's loads are issued
          // before the wait for the ring slot
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            float4 f[G / 2][2], g2[G / 2][2];
            bool ok[G / 2];
#pragma unroll
            for (int i = 0; i < G / 2; ++i) {
              const int rho = G * sr + (G / 2) * hb + i, gy = y00 + rho;
              ok[i] = sr < Gm && col_in && rho < P && gy >= 0 && gy < p.h;
              if (ok[i]) {
                const long long o = ((long long)gy * p.wd + gx) * C;
                ldg256(in_b + o, f[i][0], f[i][1]);
                ldg256(p.in2 + (in_b - p.in) + o, g2[i][0], g2[i][1]);
              }
            }
            if (hb == 0 && S >= (uint32_t)KIN) mbar_wait_sleep(bars + (BAR_FREE + slot) * 8, ((S / KIN) - 1u) & 1u);
#pragma unroll
            for (int i = 0; i < G / 2 && sr < Gm; ++i) {
              const int rho = G * sr + (G / 2) * hb + i;
              float4 u = make_float4(0.f, 0.f, 0.f, 0.f), v = u;   // outside the image: the zero padding, whatever the coefficients
              if (ok[i]) {
                u.x = fmaf(ca[0], f[i][0].x, fmaf(cb[0], g2[i][0].x, cc[0])); u.y = fmaf(ca[1], f[i][0].y, fmaf(cb[1], g2[i][0].y, cc[1]));
                u.z = fmaf(ca[2], f[i][0].z, fmaf(cb[2], g2[i][0].z, cc[2])); u.w = fmaf(ca[3], f[i][0].w, fmaf(cb[3], g2[i][0].w, cc[3]));
                v.x = fmaf(ca[4], f[i][1].x, fmaf(cb[4], g2[i][1].x, cc[4])); v.y = fmaf(ca[5], f[i][1].y, fmaf(cb[5], g2[i][1].y, cc[5]));
                v.z = fmaf(ca[6], f[i][1].z, fmaf(cb[6], g2[i][1].z, cc[6])); v.w = fmaf(ca[7], f[i][1].w, fmaf(cb[7], g2[i][1].w, cc[7]));
                // rows 0 and P - 1 of a segment are halo rows: another segment (or nobody, outside the image) owns them
                if (col_side && rho >= 1 && rho < P - 1) stg256(p.out2 + (in_b - p.in) + ((long long)(y00 + rho) * p.wd + gx) * C, u, v);
              }
              u.x *= p.in_scale; u.y *= p.in_scale; u.z *= p.in_scale; u.w *= p.in_scale;
              v.x *= p.in_scale; v.y *= p.in_scale; v.z *= p.in_scale; v.w *= p.in_scale;
              uint4 hi, lo;
              split8(u, v, hi, lo);
              const uint32_t dst = pl0 + (uint32_t)hf * PLANE_BYTES + ((slot * G + (uint32_t)((G / 2) * hb + i)) * RW + (uint32_t)c) * 16u;
              sts128(dst, hi);
              sts128(dst + 2 * PLANE_BYTES, lo);
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + (BAR_FULL + slot) * 8);
          continue;
        }
        // all loads of the step first (no shared-memory store in between: they stay in flight together, and while this
        // group waits for its ring slot)
        float4 f[G][2];
#pragma unroll
        for (int i = 0; i < G; ++i) {
          const int rho = G * sr + i, gy = y00 + rho;
          if (sr < Gm && col_in && rho < P && gy >= 0 && gy < p.h) {
            ldg256(in_b + ((long long)gy * p.wd + gx) * C, f[i][0], f[i][1]);
          } else {
            f[i][0] = make_float4(0.f, 0.f, 0.f, 0.f); f[i][1] = f[i][0];
          }
        }
        if (S >= (uint32_t)KIN) mbar_wait_sleep(bars + (BAR_FREE + slot) * 8, ((S / KIN) - 1u) & 1u);
#pragma unroll
        for (int i = 0; i < G && sr < Gm; ++i) {
          float4 u = f[i][0], v = f[i][1];
          u.x *= p.in_scale; u.y *= p.in_scale; u.z *= p.in_scale; u.w *= p.in_scale;
          v.x *= p.in_scale; v.y *= p.in_scale; v.z *= p.in_scale; v.w *= p.in_scale;
          uint4 hi, lo;
          split8(u, v, hi, lo);
          const uint32_t dst = pl0 + (uint32_t)hf * PLANE_BYTES + ((slot * G + (uint32_t)i) * RW + (uint32_t)c) * 16u;
          sts128(dst, hi);
          sts128(dst + 2 * PLANE_BYTES, lo);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + (BAR_FULL + slot) * 8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (EPI == CONV_STATS && tid < 2 * C) atomicAdd(&p.stats[tid], (double)s_stat[tid]);
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

}  // namespace t5

int launch_conv3x3_t5(bfcnn_handle* h, const float* in, float* out, const float* w, const float* res, double* stats,
                      ConvEpi epi, const Extent& e, float in_scale, cudaStream_t st, const float* in2, float* out2,
                      const float* coef, uint16_t* relu_mask) {
  using namespace t5;
  BF_REQUIRE(in_scale > 0.f, "in_scale must be a positive power of two");
  static bool attr_set_dev[64] = {};   // function attributes are per device: one flag per device ordinal
  bool& attr_set = attr_set_dev[h->device & 63];
  if (!attr_set) {
    BF_CUDA(cudaFuncSetAttribute((const void*)conv3x3_t5_kernel<CONV_PLAIN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    BF_CUDA(cudaFuncSetAttribute((const void*)conv3x3_t5_kernel<CONV_RELU, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    BF_CUDA(cudaFuncSetAttribute((const void*)conv3x3_t5_kernel<CONV_RESIDUAL, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    BF_CUDA(cudaFuncSetAttribute((const void*)conv3x3_t5_kernel<CONV_STATS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    BF_CUDA(cudaFuncSetAttribute((const void*)conv3x3_t5_kernel<CONV_MASK, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    BF_CUDA(cudaFuncSetAttribute((const void*)conv3x3_t5_kernel<CONV_RELU, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    BF_CUDA(cudaFuncSetAttribute((const void*)conv3x3_t5_kernel<CONV_MASK, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  Params p;
  p.in = in; p.out = out; p.w = w; p.res = res; p.stats = stats;
  p.in2 = in2; p.out2 = out2; p.coef = coef;
  p.mask_out = epi == CONV_RELU ? relu_mask : nullptr;
  p.mask_in = epi == CONV_MASK ? relu_mask : nullptr;
  BF_REQUIRE(epi != CONV_MASK || res != nullptr || relu_mask != nullptr, "mask epilogue needs the activation or its bit mask");
  const bool fused = in2 != nullptr;
  BF_REQUIRE(!fused || ((epi == CONV_RELU || epi == CONV_MASK) && out2 && coef), "fused conv prologue: ReLU / mask epilogues only");
  p.n = e.n; p.h = e.he; p.wd = e.we;
  // The images of the batch are laid side by side in one VIRTUAL row with a zero column between neighbours (which is the
  // zero padding of both): strips of 126 output columns run across image boundaries, so a 256-pixel-wide crop does not
  // waste a third of the lanes of its third strip.
  const long long vwidth = (long long)e.n * (e.we + 1) - 1;
  BF_REQUIRE(vwidth < (1ll << 30), "batch too wide for the virtual row");
  p.tiles_x = (int)((vwidth + (RW - 2) - 1) / (RW - 2));
  p.rows_needed = e.he; p.seg_overhead = SEG_OVERHEAD;
  const int grid = stream::plan_split(p, h->sm_count, MIN_SHARE);
  p.in_scale = in_scale;
  p.out_scale = 1.0f / (in_scale * W_SCALE);
  auto go = [&](auto kernel) { return launch_pdl(kernel, grid, NTHREADS, (size_t)SMEM_BYTES, st, p); };
  switch (epi) {
    case CONV_PLAIN: BF_CUDA(go(conv3x3_t5_kernel<CONV_PLAIN, 0>)); break;
    case CONV_RELU: BF_CUDA(fused ? go(conv3x3_t5_kernel<CONV_RELU, 1>) : go(conv3x3_t5_kernel<CONV_RELU, 0>)); break;
    case CONV_RESIDUAL: BF_CUDA(go(conv3x3_t5_kernel<CONV_RESIDUAL, 0>)); break;
    case CONV_STATS: BF_CUDA(go(conv3x3_t5_kernel<CONV_STATS, 0>)); break;
    case CONV_MASK: BF_CUDA(fused ? go(conv3x3_t5_kernel<CONV_MASK, 1>) : go(conv3x3_t5_kernel<CONV_MASK, 0>)); break;
    default: set_error("unsupported conv epilogue"); return BFCNN_ERR_INTERNAL;
  }
  h->launches++;
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

}  // namespace bfcnn
