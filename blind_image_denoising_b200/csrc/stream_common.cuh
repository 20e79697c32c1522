// stream_common.cuh -- what the row-streaming tcgen05 kernels share (fused_stream.cu: F16 stack, fused_stream_x3.cu: F16X3
// stack, conv_t5.cu: one training conv, base_conv_t5.cu: base conv): geometry constants, the cost-space split of the
// (strip, row) space over the persistent CTAs, the barrier-helper and TMA-producer warps of the inference stacks, the head
// activation, and the host side of a pass (split, programmatic dependent launch).
#pragma once
#include <algorithm>

#include "kernels.cuh"
#include "umma_ptx.cuh"

namespace bfcnn {
namespace stream {

using namespace tc5;

constexpr int RW = 128;                 // strip width == UMMA M
constexpr int SLACK_PX = 8;             // pixels of slack before/after every plane (tap shift -1/+1)
constexpr int ROW_BYTES = RW * 16;      // one row of one channel-half plane
constexpr int GROUP_BYTES = 2 * ROW_BYTES;
constexpr int MAX_SMEM = 232448;

// ---------------------------------------------------------------------------------------------------------------------
// Work split.  The (strip, row) space of a pass is linearised strip by strip; every strip costs rows_needed + seg_overhead
// COST units, the first seg_overhead of which stand for the halo rows and the pipeline fill / drain a CTA pays when it
// starts a new segment at a strip boundary.  Equal cost ranges instead of equal row ranges keep CTAs whose range spans
// two strips from running ~28 rows longer than the others (6 % at one 4K frame per pass).
// ---------------------------------------------------------------------------------------------------------------------
struct Split {
  int tiles_x, rows_needed;       // strips per (virtual) row, output rows per strip
  long long total_rows, share;    // linearised (strip, row) space; COST units each CTA owns
  int seg_overhead;               // cost of starting a segment at a strip start, in rows
};
__host__ __device__ __forceinline__ long long cost_to_row(const Split& p, long long c) {
  const long long per = (long long)p.rows_needed + p.seg_overhead;
  const long long s = c / per, off = c - s * per;
  const long long in = off - p.seg_overhead;
  return s * p.rows_needed + (in < 0 ? 0 : (in > p.rows_needed ? (long long)p.rows_needed : in));
}
struct Seg { int b, j, ya, yb; };
// the next segment of the linear row range [a, r1): rows [ya, yb) of strip j of image b
__device__ __forceinline__ Seg seg_at(const Split& p, long long a, long long r1) {
  Seg s;
  const long long strip = a / p.rows_needed;
  s.ya = (int)(a - strip * p.rows_needed);
  s.yb = (int)min((long long)p.rows_needed, (long long)s.ya + (r1 - a));
  s.b = (int)(strip / p.tiles_x);
  s.j = (int)(strip - (long long)s.b * p.tiles_x);
  return s;
}
// [r0, r1) of this CTA
__device__ __forceinline__ void cta_rows(const Split& p, long long& r0, long long& r1) {
  r0 = min(p.total_rows, cost_to_row(p, (long long)blockIdx.x * p.share));
  r1 = min(p.total_rows, cost_to_row(p, ((long long)blockIdx.x + 1) * p.share));
}
// host: fills total_rows / share from tiles_x, rows_needed, seg_overhead; returns the grid (<= sm_count persistent CTAs,
// fewer when a CTA would own less than min_share rows)
inline int plan_split(Split& p, int sm_count, int min_share) {
  p.total_rows = (long long)p.tiles_x * p.rows_needed;
  const long long total_cost = (long long)p.tiles_x * ((long long)p.rows_needed + p.seg_overhead);
  int grid = (int)std::min<long long>(sm_count, std::max<long long>(1, p.total_rows / min_share));
  p.share = (total_cost + grid - 1) / grid;
  return (int)((total_cost + p.share - 1) / p.share);
}

// ---------------------------------------------------------------------------------------------------------------------
// Inference stacks (19 warps: 16 epilogue, MMA issuer, barrier helper, TMA producer)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int EPI_WARPS = 16;           // 4 sets x 4 TMEM lane quarters
constexpr int WARP_MMA = 16;            // warp 16 issues the MMAs, warp 17 waits on its barriers, warp 18 is the TMA producer
constexpr int NTHREADS = 32 * 19;
constexpr int LAG = 3;                  // steps between consecutive layers
constexpr int KT = 3;                   // T rings: written by the epilogue of layer l at step w+1 (which has only seen layer l's
                                        // MMAs of that step), read by the MMAs of layer l+1 at step w+3: three groups
constexpr int W_LAYER_BYTES = 3 * 48 * 16 * 2;   // B operand of one conv: [dx 3][N 48][K 16] fp16
constexpr int MAX_NL = 4;
constexpr int MIN_SHARE = 8;            // rows per CTA below which fewer CTAs are launched
__host__ __device__ inline uint32_t plane_bytes_of(int rows) { return (uint32_t)(rows * RW + 2 * SLACK_PX) * 16u; }

enum Kind { KIND_A = 0, KIND_B_TO_X = 1, KIND_B_OUT = 2 };

// model.py:342 tanh(2y)*0.51, then utilities.py:435-443 (clip(+-0.5)+0.5)*255, with tanh(z) = 1 - 2/(exp(2z)+1) on the
// fast exp / divide units (absolute error ~1e-6 of the +-1 range, 1e-4 on the 0-255 scale): the head sits on the
// epilogue's critical path in the last pass
__device__ __forceinline__ float head_activation_fast(float y) {
  const float e = __expf(4.0f * y);
  float t = (1.0f - __fdividef(2.0f, e + 1.0f)) * 0.51f;
  t = fminf(fmaxf(t, -0.5f), 0.5f);
  return (t + 0.5f) * 255.0f;
}
__device__ __forceinline__ void store_rgb(void* out_row, int out_u8, float r, float g, float b) {
  if (out_u8) {
    uint8_t* d = reinterpret_cast<uint8_t*>(out_row);
    d[0] = (uint8_t)__float2int_rn(r); d[1] = (uint8_t)__float2int_rn(g); d[2] = (uint8_t)__float2int_rn(b);   // round half to even
  } else {
    float* d = reinterpret_cast<float*>(out_row);
    d[0] = r; d[1] = g; d[2] = b;
  }
}

struct EpiCtx {
  uint32_t tq;             // TMEM address of this warp's lane quarter, column 0
  uint32_t pix;            // byte offset of this thread's pixel inside a ring row
  int y00, he, h_img;      // row rho of the segment is image row y00 + rho
  int P, nl;
  bool col_ok, col_out;
  __half* fout_col;        // feature-map address of (b, y00, gx); row rho adds rho * row_halves
  uint8_t* out_col;        // output address of (b, y00, gx)
  long long row_halves, row_out;
  int gb0;                 // X0 ring group slot of the segment's group 0
};

// steps of a segment of `rows` output rows with nl layers in flight
__device__ __forceinline__ void seg_steps(int rows, int nl, int& P, int& Gm, int& nsteps) {
  P = rows + 2 * nl; Gm = (P + 1) >> 1; nsteps = Gm + LAG * (nl - 1) + 1;
}

// Barrier helper of the MMA issuer.  The issuing thread never touches shared memory: a completed mbarrier.try_wait on it
// costs ~360 cycles of tensor-pipe bubble (tools/umma_probe3.cu: the load queues behind the operand fetches, and the MMA
// queue is shallow).  Step S needs: epi_done(S-2) (lane 0: input rows written, accumulator blocks drained), x_full of
// layer 0's group (lane 1), and at a segment start epi_done(S-1) too (lane 2: every accumulator block drained before the
// ring restarts at row 0).  One barrier per lane, in parallel; the issuer is released through a named barrier
// (bar.arrive / bar.sync 2 + (S & 1): hardware barrier, no shared-memory traffic).
template <int K0, uint32_t BAR_EPI, uint32_t BAR_XFULL>
__device__ __forceinline__ void helper_warp_loop(const Split& sp, uint32_t bars, long long r0, long long r1, int nl, int lane) {
  uint32_t S = 0;
  long long gg = 0;
  for (long long a = r0; a < r1;) {
    const Seg sg = seg_at(sp, a, r1);
    a += sg.yb - sg.ya;
    int P, Gm, nsteps;
    seg_steps(sg.yb - sg.ya, nl, P, Gm, nsteps);
    for (int sr = 0; sr < nsteps; ++sr, ++S) {
      if (lane == 0 && S >= 2) mbar_wait_sleep(bars + (BAR_EPI + (S & 1u)) * 8, ((S - 2) >> 1) & 1u);
      if (lane == 1 && sr < Gm) {
        const long long k = gg + sr;
        mbar_wait_sleep(bars + (BAR_XFULL + (uint32_t)(k % K0)) * 8, (uint32_t)(k / K0) & 1u);
      }
      if (lane == 2 && sr == 0 && S >= 1) mbar_wait_sleep(bars + (BAR_EPI + ((S - 1) & 1u)) * 8, ((S - 1) >> 1) & 1u);
      __syncwarp();
      tc_fence_before();
      asm volatile("bar.arrive %0, 64;\n" ::"r"(2 + (S & 1u)) : "memory");
    }
    gg += Gm;
  }
}

// TMA producer (one elected thread): the input rows of layer 0 in groups of 2 rows, K0 groups deep; PARTS = 2 fetches the
// lo map ("image" 1 of the tensor map) behind the hi map.  Out-of-extent rows / columns are zero-filled by the hardware (=
// the "same" padding).  x0 = ring address of pixel 0, row slot 0, channel half 0 of the hi part; planes follow each other
// at x0_plane (hi half 0, hi half 1 [, lo half 0, lo half 1]).
template <int K0, int PARTS, uint32_t BAR_XFULL, uint32_t BAR_XFREE>
__device__ __forceinline__ void tma_producer_loop(const Split& sp, const CUtensorMap* tmap, uint32_t bars, uint32_t x0, uint32_t x0_plane,
                                                  long long r0, long long r1, int nl, int tw) {
  long long gg = 0;
  for (long long a = r0; a < r1;) {
    const Seg sg = seg_at(sp, a, r1);
    a += sg.yb - sg.ya;
    int P, Gm, nsteps;
    seg_steps(sg.yb - sg.ya, nl, P, Gm, nsteps);
    const int gx0 = sg.j * tw - nl, y00 = sg.ya - nl;   // gx0: column of the virtual row
    for (int g = 0; g < Gm; ++g, ++gg) {
      const uint32_t k = (uint32_t)(gg % K0), n = (uint32_t)(gg / K0);
      if (n >= 1) mbar_wait_sleep(bars + (BAR_XFREE + k) * 8, (n - 1) & 1u);
      const uint32_t bar = bars + (BAR_XFULL + k) * 8;
      mbar_arrive_expect_tx(bar, 2 * PARTS * GROUP_BYTES);
      const uint32_t dst = x0 + k * GROUP_BYTES;
#pragma unroll
      for (int part = 0; part < PARTS; ++part) {
        tma_load_q(dst + (2 * part) * x0_plane, tmap, 0, gx0, y00 + 2 * g, part, bar);
        tma_load_q(dst + (2 * part + 1) * x0_plane, tmap, 1, gx0, y00 + 2 * g, part, bar);
      }
    }
  }
}

// One-time CTA setup shared by the inference stacks: barrier init (counts: mma_done 1, epi_done one per epilogue warp,
// x_full 1, x_free one per warp of the 2 rows x 4 quarters that read the group -- 32 same-address arrivals per warp showed
// up as ~200 extra shared-memory wavefronts per step) and the TMEM allocation (all 512 columns) by warp 0.
template <uint32_t BAR_EPI, uint32_t BAR_XFULL, uint32_t BAR_XFREE, uint32_t NBARS>
__device__ __forceinline__ void init_barriers_and_tmem(uint32_t bars, uint32_t tmem_slot, int tid, int warp) {
  if (tid < (int)NBARS) {
    const uint32_t cnt = tid < (int)BAR_EPI ? 1u : (tid < (int)BAR_XFULL ? (uint32_t)EPI_WARPS : (tid < (int)BAR_XFREE ? 1u : 8u));
    mbar_init(bars + tid * 8, cnt);
  }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
}

// host: geometry of pass `blk0 .. blk0 + nblk` of an N-block model.  Rows / columns this pass has to produce: the layers that
// are still to come (rem) shrink the cone of influence by one pixel each, so beyond h + rem (w + rem) nothing can reach the
// cropped output any more (the band of the pow2 canvas is only as wide as the receptive field that is LEFT).
// What lies beyond keeps stale values: never read by a valid output.  Returns the grid.
inline int plan_pass(Split& sp, int& tw, const Extent& e, int N, int blk0, int nblk, int sm_count) {
  const int nl = 2 * nblk;
  tw = RW - 2 * nl;
  const int rem = 2 * (N - blk0 - nblk);
  sp.rows_needed = std::min(e.he, e.h + rem);
  // columns of the virtual row that need an output: up to the last needed column of the last image
  const long long cols_needed = (long long)(e.n - 1) * (e.we + 1) + std::min(e.we, e.w + rem);
  sp.tiles_x = (int)((cols_needed + tw - 1) / tw);
  sp.seg_overhead = 2 * nl + 2 * (LAG * (nl - 1) + 1);
  return plan_split(sp, sm_count, MIN_SHARE);
}

}  // namespace stream
}  // namespace bfcnn
