// base_conv_t5.cu -- normalise + base conv 3x3, 3 -> 16 (utilities.py:449-461, backbone_resnet.py:258-262) on tcgen05, from the
// uint8 image straight into the fp16 NHWC16 feature map of the streaming stacks (hi part; hi + lo for the F16X3 stack).
//
// The row-streaming single-layer pipeline of conv_t5.cu with a different operand on each side:
//   * A operand: ONE 16-byte chunk per pixel, (r/256, g/256, b/256, m, 0, 0, 0, 0) in fp16 -- the raw value / 256 is exact
//     in fp16, m = 1 inside the work extent and 0 outside (zero padding of the NORMALISED tensor; raw-zero canvas pixels
//     have m = 1, utilities.py:749).  With LBO = 16 B the two K core matrices of an MMA are two NEIGHBOURING pixels, so
//     K = 16 spans the taps dx = -1, 0 and a second MMA, shifted by two pixels, the tap dx = +1 (its other half meets zero
//     weights): 2 MMAs per input row and weight part instead of 3, N = 48 dy-scatter as everywhere else.
//   * B operand: x/255 - 0.5 folded into the weights,  sum w (v/255 - 0.5 m) = sum (256/255 w)(v/256) + (-0.5 sum_c w) m,
//     split into fp16 hi + lo (the activations are exact, so hi + lo weights give FP32-grade results: 4 MMAs per row); the
//     F16 stack rounds this layer's output to fp16 anyway and takes the hi weights only (2 MMAs per row).
//   * converter warps: a thread owns one pixel column, reads the 3 bytes of its pixel in G = 4 rows, stores one chunk per
//     row; epilogue warps: TMEM -> fp16 (and the fp16 rounding error as the lo part) -> one 256-bit store per pixel and part.
// Per 128-pixel row: 4 x 5.5 KB of operands against 4 KB (8 KB) written to HBM: the kernel is bound by the HBM write of
// the feature map (32 or 64 B per pixel); the mma.sync kernel it replaces (base_conv.cu) ran at a third of that rate.
//
// Output layout: the virtual-row map of the streaming stacks, [he][vw = n (we + 1)][16] (fused_stream.cu): the images of
// the batch side by side, a zero column after each; strips of 126 output columns run across the image boundaries.
#include "stream_common.cuh"

namespace bfcnn {
namespace bt5 {

using namespace tc5;
using stream::RW; using stream::SLACK_PX; using stream::Split; using stream::Seg; using stream::seg_at; using stream::cta_rows;

constexpr int G = 4;                    // rows per step
constexpr int EPI_WARPS = 8;            // G rows x 4 TMEM lane quarters = 16 tasks per step, two per warp
constexpr int WARP_MMA = 8, WARP_HELP = 9, WARP_CVT = 10, CVT_WARPS = 4;   // converter warps per group: thread = pixel column
constexpr int CVT_GROUPS = 4;           // converter groups take every 4th step: four steps of global loads in flight (with two
                                        // the kernel waited for the byte loads: 25 % of the stall samples, ncu r02c)
constexpr int NTHREADS = 32 * (WARP_CVT + CVT_GROUPS * CVT_WARPS);
constexpr int KIN = 8;                  // input ring: groups of G rows
constexpr int PLANE_BYTES = (KIN * G * RW + 2 * SLACK_PX) * 16;   // one 16-byte chunk per pixel
constexpr int W_CHUNK_BYTES = 48 * 16 * 2;                        // B operand of one K chunk: [N 48][K 16] fp16
constexpr int W_PART_BYTES = 2 * W_CHUNK_BYTES;                   // chunks (dx -1, 0) and (dx +1, -)
constexpr float W_SCALE = 256.f;        // keeps the low part of the weights out of the fp16 subnormals
constexpr uint32_t BAR_MMA = 0, BAR_EPI = 2, BAR_FULL = 4, BAR_FREE = 4 + KIN, NBARS = 4 + 2 * KIN;
constexpr uint32_t SM_BARS = 0, SM_TMEM = 256, SM_WTS = 512, SM_PLANES = 7168;
static_assert(SM_WTS + 2 * W_PART_BYTES <= SM_PLANES, "weights overlap the plane");
constexpr int SMEM_BYTES = SM_PLANES + PLANE_BYTES;
constexpr int MIN_SHARE = 16;
constexpr int SEG_OVERHEAD = 2 + 2 * G;   // halo rows + pipeline fill / drain of a segment, in rows (cost-space split)

struct Params : Split {   // rows_needed = he, seg_overhead = SEG_OVERHEAD
  const uint8_t* img;   // [n][h][w][3]
  __half* out;          // [he][vw][16] hi part
  __half* out_lo;       // lo part or nullptr
  const float* w;       // [9][3][16 cout] fp32 (d_base_f32)
  int n, h, wd, he, we, vw;
};

template <bool WITH_LO>
__global__ void __launch_bounds__(NTHREADS, 1)
base_conv3_t5_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t s0 = smem_u32(smem);
  const uint32_t bars = s0 + SM_BARS;
  const uint32_t pl0 = s0 + SM_PLANES + SLACK_PX * 16;   // pixel 0 of ring row 0
  long long r0, r1;
  cta_rows(p, r0, r1);

  // ---------------- setup: barriers, TMEM, weights (normalisation folded in, scaled hi / lo, UMMA B layout), zeroed plane
  if (tid < (int)NBARS) {
    const uint32_t cnt = tid < (int)BAR_EPI ? 1u : (tid < (int)BAR_FULL ? (uint32_t)EPI_WARPS : (tid < (int)BAR_FREE ? (uint32_t)CVT_WARPS : 1u));
    mbar_init(bars + tid * 8, cnt);
  }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(s0 + SM_TMEM) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  {
    __half* wh = reinterpret_cast<__half*>(smem + SM_WTS);
    __half* wl = wh + W_PART_BYTES / 2;
    for (int i = tid; i < 2 * 48 * 16; i += NTHREADS) {
      const int ck = i / 768, r = i - ck * 768, nn = r >> 4, k = r & 15;   // (K chunk, n = j*16 + cout, k = (pixel t, channel))
      const int j = nn >> 4, co = nn & 15, t = k >> 3, ch = k & 7;
      const int dxi = 2 * ck + t;                                          // chunk 0: dx -1, 0; chunk 1: dx +1, (nothing)
      float v = 0.f;
      if (dxi < 3 && ch < 4) {
        const int tap = (2 - j) * 3 + dxi;                                 // block j <-> dy = 1 - j (host_pack.cu)
        const float* wt = p.w + (size_t)tap * 3 * C + co;
        v = ch < 3 ? wt[ch * C] * (256.0f / 255.0f) : -0.5f * (wt[0] + wt[C] + wt[2 * C]);
      }
      v *= W_SCALE;
      const __half hv = __float2half_rn(v);
      const int off = ck * 768 + (k >> 3) * 384 + (nn >> 3) * 64 + (nn & 7) * 8 + (k & 7);
      wh[off] = hv;
      wl[off] = __float2half_rn(v - __half2float(hv));
    }
    for (uint32_t i = tid; i < (uint32_t)PLANE_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem + SM_PLANES)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM);

  if (warp < EPI_WARPS) {
    // ================= epilogue warps =================
    const int quarter = warp & 3, rsel0 = warp >> 2;   // rows rsel0 and rsel0 + 2 of the group
    const int c = quarter * 32 + lane;
    const uint32_t tq = tmem + ((uint32_t)(quarter * 32) << 16);
    for (int blk = rsel0; blk < 32; blk += 2) tmem_zero16(tq + blk * 16);
    tmem_wait_st();
    tc_fence_before();
    asm volatile("bar.sync 1, %0;\n" ::"r"(32 * (EPI_WARPS + 1)) : "memory");
    constexpr float OUT_SCALE = 1.0f / W_SCALE;
    uint32_t S = 0;
    for (long long a = r0; a < r1;) {
      const Seg sg = seg_at(p, a, r1);
      a += sg.yb - sg.ya;
      const int P = (sg.yb - sg.ya) + 2, Gm = (P + G - 1) / G, nsteps = Gm + 1;
      const int y00 = sg.ya - 1, vx = sg.j * (RW - 2) - 1 + c;
      const int vb = vx >= 0 ? vx / (p.we + 1) : -1, gx = vx - vb * (p.we + 1);
      const bool col_out = (c >= 1) && (c < RW - 1) && (vb >= 0) && (vb < p.n) && (gx < p.we);
      const long long px0 = (long long)y00 * p.vw + vx;   // pixel index of (row y00, virtual column vx)
      for (int sr = 0; sr < nsteps; ++sr, ++S) {
        const int w = sr - 1;
        mbar_wait_sleep(bars + (BAR_MMA + (S & 1u)) * 8, (S >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2) {
          if (w < 0) break;
          const int rho = G * w + rsel0 + 2 * k2;
          const uint32_t taddr = tq + (uint32_t)(rho & 31) * 16u;
          uint32_t v[16];
          tmem_ld16(taddr, v);
          tmem_zero16(taddr);
          if (col_out && rho >= 1 && rho < P - 1) {
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * OUT_SCALE;
            uint4 ha, hb;
            ha.x = pack_h2(f[0], f[1]); ha.y = pack_h2(f[2], f[3]); ha.z = pack_h2(f[4], f[5]); ha.w = pack_h2(f[6], f[7]);
            hb.x = pack_h2(f[8], f[9]); hb.y = pack_h2(f[10], f[11]); hb.z = pack_h2(f[12], f[13]); hb.w = pack_h2(f[14], f[15]);
            const long long o = (px0 + (long long)rho * p.vw) << 4;
            stg256(p.out + o, ha, hb);
            if (WITH_LO) {   // the fp16 rounding error of every channel (F16X3 arithmetic)
              const uint32_t hs[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
              uint32_t ls[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 r = unpack_h2(hs[i]);
                ls[i] = pack_h2(f[2 * i] - r.x, f[2 * i + 1] - r.y);
              }
              stg256(p.out_lo + o, make_uint4(ls[0], ls[1], ls[2], ls[3]), make_uint4(ls[4], ls[5], ls[6], ls[7]));
            }
          }
          tmem_wait_st();
          tc_fence_before();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + (BAR_EPI + (S & 1u)) * 8);
      }
    }
  } else if (warp == WARP_MMA) {
    // ================= MMA issuer =================
    asm volatile("bar.sync 1, %0;\n" ::"r"(32 * (EPI_WARPS + 1)) : "memory");
    const uint32_t idesc0 = make_idesc_f16(128, 0);
    const uint32_t a_base = desc_lo(pl0, 16) - 1u;   // ring row 0, pixel -1; LBO = 16 B: the second K half is the NEXT pixel
    const uint32_t b_hi = desc_lo(s0 + SM_WTS, 48 * 16), b_lo = b_hi + (uint32_t)(W_PART_BYTES / 16);
    constexpr uint32_t BCK = W_CHUNK_BYTES / 16;
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
            const uint32_t slot_row = (uint32_t)((S % KIN) * G);
            const int rho0 = G * sr;
            const uint32_t blk0 = (uint32_t)(rho0 - 1) & 31u;
            if (rho0 >= 1 && rho0 + G < P && blk0 + G + 2 <= 32u) {
              // fast path (7 steps in 8): four interior rows whose G + 2 accumulator blocks do not wrap around the TMEM
              // ring -> straight-line issue; every instruction between two MMAs of the issuing thread is on the critical
              // path (with only 2-4 MMAs per row the per-row bookkeeping of the general path cost as much as the MMAs)
              const uint32_t id = idesc0 + (6u << 17);
              const uint32_t ar0 = a_base + slot_row * RW, d0 = tmem + blk0 * 16u;
#pragma unroll
              for (int i = 0; i < G; ++i) {
                const uint32_t ar = ar0 + (uint32_t)i * RW, d = d0 + (uint32_t)i * 16u;
                if (WITH_LO) {
                  mma_lo(d, ar, b_lo, id);
                  mma_lo(d, ar + 2u, b_lo + BCK, id);
                }
                mma_lo(d, ar, b_hi, id);
                mma_lo(d, ar + 2u, b_hi + BCK, id);
              }
            } else {
#pragma unroll 1
              for (int i = 0; i < G; ++i) {
                const int rho = rho0 + i;
                if (rho >= P) break;
                const uint32_t arow = a_base + (slot_row + (uint32_t)i) * RW;
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
                  if (WITH_LO) {
                    mma_lo(d, arow, b_lo + boff, id);
                    mma_lo(d, arow + 2u, b_lo + boff + BCK, id);
                  }
                  mma_lo(d, arow, b_hi + boff, id);
                  mma_lo(d, arow + 2u, b_hi + boff + BCK, id);
                }
              }
            }
          }
          umma_commit(bars + (BAR_MMA + (S & 1u)) * 8);
          umma_commit(bars + (BAR_FREE + (S % KIN)) * 8);
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
    // ================= converter warps: uint8 pixels -> (r, g, b, m) fp16 chunks =================
    // The ring slot of a step is S % KIN for EVERY step (the epilogue-only step at the end of a segment arrives with an
    // empty slot), so issuer, helper and converters index barriers and rows by the same global step counter.
    const int cw = warp - WARP_CVT, cgrp = cw / CVT_WARPS;
    const int c = (cw % CVT_WARPS) * 32 + lane;   // pixel column of the strip
    uint32_t S = 0;
    for (long long a = r0; a < r1;) {
      const Seg sg = seg_at(p, a, r1);
      a += sg.yb - sg.ya;
      const int P = (sg.yb - sg.ya) + 2, Gm = (P + G - 1) / G, nsteps = Gm + 1;
      const int y00 = sg.ya - 1, vx = sg.j * (RW - 2) - 1 + c;
      const int vb = vx >= 0 ? vx / (p.we + 1) : -1, gx = vx - vb * (p.we + 1);
      const bool col_ext = (vb >= 0) && (vb < p.n) && (gx < p.we);   // separator columns and the outside: all zero
      const bool col_img = col_ext && (gx < p.wd);
      const uint8_t* img_b = p.img + ((long long)vb * p.h * p.wd + gx) * 3;
      for (int sr = 0; sr < nsteps; ++sr, ++S) {
        if ((int)(S % CVT_GROUPS) != cgrp) continue;   // the other converter group's step
        const uint32_t slot = S % KIN;
        // all loads of the step first: they stay in flight together, and while this group waits for its ring slot
        uint32_t px[G];
#pragma unroll
        for (int i = 0; i < G; ++i) {
          const int rho = G * sr + i, gy = y00 + rho;
          px[i] = 0u;
          if (sr < Gm && col_img && rho < P && gy >= 0 && gy < p.h) {
            const uint8_t* sp = img_b + (long long)gy * p.wd * 3;
            px[i] = (uint32_t)sp[0] | ((uint32_t)sp[1] << 8) | ((uint32_t)sp[2] << 16);
          }
        }
        if (S >= (uint32_t)KIN) mbar_wait_sleep(bars + (BAR_FREE + slot) * 8, ((S / KIN) - 1u) & 1u);
#pragma unroll
        for (int i = 0; i < G && sr < Gm; ++i) {
          const int rho = G * sr + i, gy = y00 + rho;
          uint4 ch = make_uint4(0u, 0u, 0u, 0u);
          if (col_ext && rho < P && gy >= 0 && gy < p.he) {
            const __half v0 = __float2half_rn((float)(px[i] & 0xFFu) * 0.00390625f);
            const __half v1 = __float2half_rn((float)((px[i] >> 8) & 0xFFu) * 0.00390625f);
            const __half v2 = __float2half_rn((float)((px[i] >> 16) & 0xFFu) * 0.00390625f);
            ch.x = (uint32_t)__half_as_ushort(v0) | ((uint32_t)__half_as_ushort(v1) << 16);
            ch.y = (uint32_t)__half_as_ushort(v2) | (0x3C00u << 16);   // m = 1.0h
          }
          sts128(pl0 + ((slot * G + (uint32_t)i) * RW + (uint32_t)c) * 16u, ch);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + (BAR_FULL + slot) * 8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

}  // namespace bt5

// feat (and feat_lo) in the virtual-row layout [he][vw = n (we + 1)][16]
int launch_base_conv3_t5(bfcnn_handle* h, const uint8_t* d_in, __half* feat, __half* feat_lo, const Extent& e, cudaStream_t st) {
  using namespace bt5;
  static bool attr_set_dev[64] = {};   // function attributes are per device: one flag per device ordinal
  bool& attr_set = attr_set_dev[h->device & 63];
  if (!attr_set) {
    BF_CUDA(cudaFuncSetAttribute((const void*)base_conv3_t5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    BF_CUDA(cudaFuncSetAttribute((const void*)base_conv3_t5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  Params p;
  p.img = d_in; p.out = feat; p.out_lo = feat_lo; p.w = h->d_base_f32.as<float>();
  p.n = e.n; p.h = e.h; p.wd = e.w; p.he = e.he; p.we = e.we;
  const long long vw = (long long)e.n * (e.we + 1);
  BF_REQUIRE(vw < (1ll << 30), "batch too wide for the virtual row");
  p.vw = (int)vw;
  p.tiles_x = (int)((vw - 1 + (RW - 2) - 1) / (RW - 2));
  p.rows_needed = e.he; p.seg_overhead = SEG_OVERHEAD;
  const int grid = stream::plan_split(p, h->sm_count, MIN_SHARE);
  if (feat_lo) base_conv3_t5_kernel<true><<<(unsigned)grid, NTHREADS, SMEM_BYTES, st>>>(p);
  else base_conv3_t5_kernel<false><<<(unsigned)grid, NTHREADS, SMEM_BYTES, st>>>(p);
  h->launches++;
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

}  // namespace bfcnn
