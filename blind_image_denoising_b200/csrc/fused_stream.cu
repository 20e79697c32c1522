// fused_stream.cu -- the fused conv-BN-ReLU residual stack on tcgen05 as a ROW-STREAMING pipeline (sm_100a).
//
// Same arithmetic and operand layouts as fused_umma.cu (M = 128 pixels of one image row per MMA, channel-half planes in
// the SWIZZLE_NONE K-major layout so that a dx tap is a descriptor shift, N = 48 dy-scatter into three accumulator
// blocks), but the work unit is no longer a 128 x rh region with a vertical halo that every layer recomputes.  A CTA
// owns a 128-pixel-wide column strip segment and streams DOWN it: all 2*nblk conv layers of the pass are in flight at
// once, layer l+1 trailing layer l by LAG = 3 row groups, every layer advancing one group of 2 rows per step.  Only the
// horizontal halo (2*nblk columns per side) is recomputed: 94 % of the MMAs are useful at nblk = 2 (the region kernel:
// 71 %), and a segment is as tall as the host cares to make it, so 148 CTAs get equal shares of the frame.
//
//   step s:  MMA issuer      layer l multiplies its input rows {2g, 2g+1}, g = s - 3l           (one elected thread)
//            epilogue warps  layer l drains its output rows    {2w, 2w+1}, w = s - 3l - 1       (16 warps)
//            barrier helper  waits on the mbarriers of step s for the issuer                    (one warp)
//            TMA producer    input rows of layer 0, K0 groups deep ring                          (one elected thread)
//
//   * Shared memory holds four row RINGS (fp16, two channel-half planes each): X0 (TMA input of block 0, also the
//     residual of block 0), T0 = ReLU(conv_a), X1 (block-0 output = block-1 input and residual), T1.
//   * The 32 TMEM accumulator blocks (16 columns each) are one ring: row r of layer l lives in block (r + 14 l) & 31.
//     At any step a layer has at most 7 rows touched-but-not-drained, the four windows sit 8 blocks apart and move
//     together.  tcgen05.mma faults when its D columns run past column 511 (tools/umma_probe5.cu), so an MMA whose
//     three blocks straddle the ring end is issued as N = 32 + N = 16 (2 input rows in 32); every output row still
//     receives its nine partial products in the same order, so results do not depend on where a row sits in the ring
//     -- strips, crops and whole frames stay bit-identical (tests/test_inference_gpu.py).
//   * mbarriers: mma_done[l][s & 1] (tcgen05.commit after layer l's MMAs of step s: the epilogue of layer l runs while
//     the later layers of the step are still being multiplied), epi_done[s & 1] (one arrival per epilogue warp at the end
//     of step s).  Step s may be issued once epi_done(s - 2) has completed, so the epilogue of step s - 1
//     overlaps the MMAs of step s; LAG = 3 is the smallest lag for which layer l+1's input group is already written by
//     then.  x_full[k] / x_free[k] (k < K0) couple the TMA producer to the issuer and to the block-0 residual readers.
//   * The issuing thread never touches shared memory (a completed mbarrier.try_wait on it costs ~360 cycles of
//     tensor-pipe bubble, tools/umma_probe3.cu): the helper warp does the waits and releases it through a named barrier.
//     Its descriptor arithmetic is incremental (from scratch it cost 40 % of the issue time).
//   * Rows outside a layer's valid cone (the first / last rows of a segment, the halo columns) are computed and
//     ignored; rows / columns outside the work extent are forced to zero by every epilogue (per-layer "same" padding).
//
// Reference arithmetic: module_denoiser.py:53-73, backbone_blocks.py:167-246 (block), model.py:297-342 (head),
// utilities.py:435-443 (denormalise).
#include "kernels.cuh"
#include "umma_ptx.cuh"

namespace bfcnn {
namespace ustream {

using namespace tc5;

constexpr int RW = 128;                 // strip width == UMMA M
constexpr int SLACK_PX = 8;             // pixels of slack before/after every plane (tap shift -1/+1)
constexpr int EPI_WARPS = 16;           // 4 sets x 4 TMEM lane quarters
constexpr int WARP_MMA = 16;            // warp 16 issues the MMAs, warp 17 waits on its barriers, warp 18 is the TMA producer
constexpr int NTHREADS = 32 * 19;
constexpr int LAG = 3;                  // steps between consecutive layers
constexpr int K0 = 11;                  // X0 ring: groups of 2 rows (TMA prefetch depth)
constexpr int KX = 8;                   // X1 ring groups: written at step w+4, last read (residual) at step w+10
constexpr int KT = 3;                   // T rings: written by the epilogue of layer l at step w+1 (which has only seen layer l's
                                        // MMAs of that step), read by the MMAs of layer l+1 at step w+3: three groups
constexpr int ROW_BYTES = RW * 16;      // one row of one channel-half plane
constexpr int GROUP_BYTES = 2 * ROW_BYTES;
constexpr int W_LAYER_BYTES = 3 * 48 * 16 * 2;   // B operand of one conv: [dx 3][N 48][K 16] fp16
constexpr int MAX_SMEM = 232448;
constexpr int MAX_NL = 4;
constexpr int MIN_SHARE = 8;            // rows per CTA below which fewer CTAs are launched

// barriers (8 B each)
// mma_done[l][s & 1] (per layer: the epilogue of layer l starts while the later layers of the step are still being
// multiplied), epi_done[s & 1], x_full[k], x_free[k]
constexpr uint32_t BAR_MMA = 0, BAR_EPI = 2 * 4, BAR_XFULL = BAR_EPI + 2, BAR_XFREE = BAR_XFULL + K0, NBARS = BAR_XFREE + K0;
constexpr uint32_t SM_BARS = 0, SM_TMEM = 512, SM_HEAD = 528, SM_BIAS = 784, SM_WTS = 1152;

__host__ __device__ inline uint32_t plane_bytes_of(int rows) { return (uint32_t)(rows * RW + 2 * SLACK_PX) * 16u; }
constexpr uint32_t HEAD_B_BYTES = 16 * 16 * 2;   // B operand of the head matrix [N 16][K 16] fp16 (last pass)
__host__ __device__ inline uint32_t head_b_offset(int nl) { return SM_WTS + (uint32_t)nl * W_LAYER_BYTES; }
__host__ __device__ inline uint32_t rings_offset(int nl) { return (head_b_offset(nl) + HEAD_B_BYTES + 127u) & ~127u; }
__host__ __device__ inline uint32_t smem_bytes(int nl) {
  uint32_t b = rings_offset(nl) + 2 * plane_bytes_of(2 * K0) + 2 * plane_bytes_of(2 * KT);
  if (nl > 2) b += 2 * plane_bytes_of(2 * KX) + 2 * plane_bytes_of(2 * KT);
  return b;
}

struct Params {
  // Feature maps between passes are laid out as ONE virtual row per image row: the n images side by side, each followed by
  // a zero column (the "same" padding of both neighbours): [he][vw = n (we + 1)][16].  Strips of tw columns run across the
  // image boundaries, so narrow images (64 x 256 x 256: three 128-lane strips per image otherwise) waste no lanes.
  const __half* fin;       // [he][vw][16]  input feature map of this pass
  __half* fout;            // [he][vw][16]
  int vw;                  // n * (we + 1)
  void* out;               // [n][h][w][3] uint8 or float
  const uint8_t* wumma;    // [2N][W_LAYER_BYTES]
  const uint8_t* wlast;    // last pass: the final conv_b with the collapsed head folded in [W_LAYER_BYTES] + the head matrix [HEAD_B_BYTES]
  const float* bias;       // [2N][16]
  const float* whead;      // [16][4]
  int n, h, w, he, we;
  int blk0, nblk;
  int out_u8;
  int tw, tiles_x, rows_needed;   // output columns per strip, strips per virtual row, output rows per strip
  long long total_rows, share;    // linearised (strip, row) space; COST units each CTA owns (see cost_to_row)
  int seg_overhead;               // cost of starting a segment at a strip start, in rows (halo rows + pipeline fill / drain)
  long long* trace;               // debug timeline (BFCNN_STREAM_TRACE=1) of CTA trace_block, steps [TRACE_S0, TRACE_S0 + 32)
  int trace_block;
};
constexpr uint32_t TRACE_S0 = 100;
#ifdef BFCNN_STREAM_TRACE_BUILD   // compile-time: the timeline costs the issuer ~5 % even when it is switched off at run time
#define STREAM_TRACE(slot) do { if (tr && S >= TRACE_S0 && S < TRACE_S0 + 32) p.trace[(S - TRACE_S0) * 8 + (slot)] = clock64(); } while (0)
#define STREAM_TRACE_PTR(cond) ((tr && (cond) && S >= TRACE_S0 && S < TRACE_S0 + 32) ? p.trace + (S - TRACE_S0) * 8 + 4 : nullptr)
#else
#define STREAM_TRACE(slot) do { } while (0)
#define STREAM_TRACE_PTR(cond) nullptr
#endif

// Work is split in COST space: every strip costs rows_needed + seg_overhead units, the first seg_overhead of which stand
// for the halo rows and the pipeline fill / drain a CTA pays when it starts a new segment at a strip boundary.  Equal
// cost ranges instead of equal row ranges keep CTAs whose range spans two strips from running ~28 rows longer than the
// others (6 % at one 4K frame per pass).
__device__ __forceinline__ long long cost_to_row(const Params& p, long long c) {
  const long long per = (long long)p.rows_needed + p.seg_overhead;
  const long long s = c / per, off = c - s * per;
  return s * p.rows_needed + max(0ll, min((long long)p.rows_needed, off - p.seg_overhead));
}
struct Seg { int b, j, ya, yb; };
// the next segment of the linear row range [a, r1): rows [ya, yb) of strip j of image b
__device__ __forceinline__ Seg seg_at(const Params& p, long long a, long long r1) {
  Seg s;
  const long long strip = a / p.rows_needed;
  s.ya = (int)(a - strip * p.rows_needed);
  s.yb = (int)min((long long)p.rows_needed, (long long)s.ya + (r1 - a));
  s.b = (int)(strip / p.tiles_x);
  s.j = (int)(strip - (long long)s.b * p.tiles_x);
  return s;
}

// rings: byte address of pixel 0, row slot 0, channel half 0; the other half is +plane
struct Rings {
  uint32_t x0, t0, x1, t1;
  uint32_t x0_plane, t_plane, x1_plane;
};

enum Kind { KIND_A = 0, KIND_B_TO_X = 1, KIND_B_OUT = 2 };

// Every shared-memory descriptor of this kernel has SBO = 128 B, version 1, SWIZZLE_NONE: the high word is one constant
// and the issuer's arithmetic (tap shifts, ring rows, weight blocks) touches the 14-bit start-address field of the low
// word only -- 32-bit adds instead of 64-bit ones on the issuing thread.
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ void mma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(idesc)
      : "memory");
}

// model.py:342 tanh(2y)*0.51, then utilities.py:435-443 (clip(+-0.5)+0.5)*255, with tanh(z) = 1 - 2/(exp(2z)+1) on the
// fast exp / divide units (absolute error ~1e-6 of the +-1 range, 1e-4 on the 0-255 scale): the head sits on the
// epilogue's critical path in the last pass
__device__ __forceinline__ float head_activation_fast(float y) {
  const float e = __expf(4.0f * y);
  float t = (1.0f - __fdividef(2.0f, e + 1.0f)) * 0.51f;
  t = fminf(fmaxf(t, -0.5f), 0.5f);
  return (t + 0.5f) * 255.0f;
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

// one (layer, row) task of one warp: 32 pixels of output row rho of layer l
template <int KIND, bool LAST_PASS>
__device__ __forceinline__ void epi_task(const Params& p, const Rings& R, const EpiCtx& E, uint32_t bars, const float (&bias)[16],
                                         const float* s_head, const float* s_bias_l, int l, int rho, long long* tp = nullptr) {
  const uint32_t taddr = E.tq + (uint32_t)((rho + 14 * l) & 31) * 16u;
  if (KIND == KIND_B_OUT && LAST_PASS) {
    // Last layer of the model: the collapsed head is folded into this conv's weights and the residual came in through an
    // MMA of its own (issuer), so accumulator columns 0..2 hold the pre-tanh head output of the pixel; add the head's
    // share of the BN constant (bias[0..2]), tanh(2y)*0.51 + denormalise (model.py:342, utilities.py:435-443), store.
    uint32_t y[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" : "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]) :: "memory");
    tmem_zero16(taddr);
    const bool inside_l = E.col_ok && ((unsigned)(E.y00 + rho) < (unsigned)E.he);
    if (E.col_out && inside_l && rho >= E.nl && rho < E.P - E.nl && (E.y00 + rho) < E.h_img) {
      const float r0o = head_activation_fast(__uint_as_float(y[0]) + bias[0]);
      const float r1o = head_activation_fast(__uint_as_float(y[1]) + bias[1]);
      const float r2o = head_activation_fast(__uint_as_float(y[2]) + bias[2]);
      if (p.out_u8) {
        uint8_t* d = E.out_col + (long long)rho * E.row_out;
        d[0] = (uint8_t)__float2int_rn(r0o); d[1] = (uint8_t)__float2int_rn(r1o); d[2] = (uint8_t)__float2int_rn(r2o);
      } else {
        float* d = reinterpret_cast<float*>(E.out_col) + (long long)rho * E.row_out;
        d[0] = r0o; d[1] = r1o; d[2] = r2o;
      }
    }
    if (l == 1) {   // one block in this pass: the X0 row was the residual operand of the issuer's extra MMA, now complete
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(bars + (BAR_XFREE + (uint32_t)((E.gb0 + (rho >> 1)) % K0)) * 8);
    }
    return;
  }
  uint32_t v[16];
  if (tp) tp[0] = clock64();
  tmem_ld16_issue(taddr, v);
  const bool inside = E.col_ok && ((unsigned)(E.y00 + rho) < (unsigned)E.he);
  if (KIND == KIND_A) {
    tmem_ld_wait(v);
    if (tp) tp[1] = clock64();
    tmem_zero16(taddr);
    if (tp) tp[2] = clock64();
    const uint32_t m = inside ? 0xFFFFFFFFu : 0u;
    uint4 lo, hi;
    lo.x = relu_h2(pack_h2(__uint_as_float(v[0]), __uint_as_float(v[1]))) & m; lo.y = relu_h2(pack_h2(__uint_as_float(v[2]), __uint_as_float(v[3]))) & m;
    lo.z = relu_h2(pack_h2(__uint_as_float(v[4]), __uint_as_float(v[5]))) & m; lo.w = relu_h2(pack_h2(__uint_as_float(v[6]), __uint_as_float(v[7]))) & m;
    hi.x = relu_h2(pack_h2(__uint_as_float(v[8]), __uint_as_float(v[9]))) & m; hi.y = relu_h2(pack_h2(__uint_as_float(v[10]), __uint_as_float(v[11]))) & m;
    hi.z = relu_h2(pack_h2(__uint_as_float(v[12]), __uint_as_float(v[13]))) & m; hi.w = relu_h2(pack_h2(__uint_as_float(v[14]), __uint_as_float(v[15]))) & m;
    const uint32_t dst = (l == 0 ? R.t0 : R.t1) + ((uint32_t)rho % (2 * KT)) * ROW_BYTES + E.pix;
    sts128(dst, lo);
    sts128(dst + R.t_plane, hi);
    if (tp) tp[3] = clock64();
  } else {
    // residual: X of this block (fp16) + the BN constant b' + the accumulator
    uint32_t xsrc, xplane;
    if (l == 1) {
      const int w = rho >> 1;
      xsrc = R.x0 + (uint32_t)(((E.gb0 + w) % K0) * 2 + (rho & 1)) * ROW_BYTES + E.pix;
      xplane = R.x0_plane;
    } else {
      xsrc = R.x1 + (uint32_t)(rho & (2 * KX - 1)) * ROW_BYTES + E.pix;
      xplane = R.x1_plane;
    }
    const uint4 xa = lds128(xsrc), xb = lds128(xsrc + xplane);
    tmem_ld_wait(v);
    tmem_zero16(taddr);
    float f[16];
    {
      const uint32_t xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 xv = unpack_h2(xs[i]);
        float b0, b1;
        if (LAST_PASS) {   // 96-register cap (19 warps): the last pass keeps the bias in shared memory (broadcast loads)
          const float2 bb = *reinterpret_cast<const float2*>(s_bias_l + 2 * i);
          b0 = bb.x; b1 = bb.y;
        } else {
          b0 = bias[2 * i]; b1 = bias[2 * i + 1];
        }
        f[2 * i] = __uint_as_float(v[2 * i]) + (xv.x + b0);
        f[2 * i + 1] = __uint_as_float(v[2 * i + 1]) + (xv.y + b1);
      }
    }
    {
      uint4 lo, hi;
      lo.x = pack_h2(f[0], f[1]); lo.y = pack_h2(f[2], f[3]); lo.z = pack_h2(f[4], f[5]); lo.w = pack_h2(f[6], f[7]);
      hi.x = pack_h2(f[8], f[9]); hi.y = pack_h2(f[10], f[11]); hi.z = pack_h2(f[12], f[13]); hi.w = pack_h2(f[14], f[15]);
      if (KIND == KIND_B_TO_X) {
        const uint32_t m = inside ? 0xFFFFFFFFu : 0u;
        lo.x &= m; lo.y &= m; lo.z &= m; lo.w &= m; hi.x &= m; hi.y &= m; hi.z &= m; hi.w &= m;
        const uint32_t dst = R.x1 + (uint32_t)(rho & (2 * KX - 1)) * ROW_BYTES + E.pix;
        sts128(dst, lo);
        sts128(dst + R.x1_plane, hi);
      } else if (E.col_out && inside && rho >= E.nl && rho < E.P - E.nl) {
        stg256(E.fout_col + (long long)rho * E.row_halves, lo, hi);   // the pixel's 32 bytes as one full-sector store
      }
    }
    if (l == 1) {
      // this warp's X0 pixels of the row are consumed: (generic read -> async-proxy TMA overwrite)
      fence_async_smem();
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(bars + (BAR_XFREE + (uint32_t)((E.gb0 + (rho >> 1)) % K0)) * 8);
    }
  }
}

template <bool LAST_PASS>
__global__ void __launch_bounds__(NTHREADS, 1)
stream_pass_kernel(const Params p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  const int nl = 2 * p.nblk, halo = nl;
  const uint32_t s0 = smem_u32(smem);
  const uint32_t bars = s0 + SM_BARS;
  float* s_head = reinterpret_cast<float*>(smem + SM_HEAD);
  float* s_bias = reinterpret_cast<float*>(smem + SM_BIAS);
  Rings R;
  {
    uint32_t o = s0 + rings_offset(nl);
    R.x0_plane = plane_bytes_of(2 * K0); R.t_plane = plane_bytes_of(2 * KT); R.x1_plane = plane_bytes_of(2 * KX);
    R.x0 = o + SLACK_PX * 16; o += 2 * R.x0_plane;
    R.t0 = o + SLACK_PX * 16; o += 2 * R.t_plane;
    R.x1 = o + SLACK_PX * 16; o += 2 * R.x1_plane;
    R.t1 = o + SLACK_PX * 16;
  }
  const long long r0 = min(p.total_rows, cost_to_row(p, (long long)blockIdx.x * p.share));
  const long long r1 = min(p.total_rows, cost_to_row(p, ((long long)blockIdx.x + 1) * p.share));
  const bool tr = (p.trace != nullptr) && ((int)blockIdx.x == p.trace_block);
  if (tr && tid == 0) p.trace[256] = clock64();

  // ---------------- one-time setup: barriers, TMEM, weights, zeroed rings
  if (tid < (int)NBARS) {
    // epi_done: one arrival per epilogue warp; x_free: one per warp of the 2 rows x 4 quarters that read the group
    // (32 same-address arrivals per warp showed up as ~200 extra shared-memory wavefronts per step)
    const uint32_t cnt = tid < (int)BAR_EPI ? 1u : (tid < (int)BAR_XFULL ? (uint32_t)EPI_WARPS : (tid < (int)BAR_XFREE ? 1u : 8u));
    mbar_init(bars + tid * 8, cnt);
  }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(s0 + SM_TMEM) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  // last pass: layer nl - 1 takes the head-folded weights, and the head matrix follows the conv weights
  for (int i = tid; i < (nl - (LAST_PASS ? 1 : 0)) * (W_LAYER_BYTES / 16); i += NTHREADS)
    reinterpret_cast<uint4*>(smem + SM_WTS)[i] = reinterpret_cast<const uint4*>(p.wumma + (size_t)(2 * p.blk0) * W_LAYER_BYTES)[i];
  if (LAST_PASS)
    for (int i = tid; i < (int)((W_LAYER_BYTES + HEAD_B_BYTES) / 16); i += NTHREADS)
      reinterpret_cast<uint4*>(smem + SM_WTS + (nl - 1) * W_LAYER_BYTES)[i] = reinterpret_cast<const uint4*>(p.wlast)[i];
  for (int i = tid; i < nl * C; i += NTHREADS) s_bias[i] = p.bias[(size_t)(2 * p.blk0) * C + i];
  if (LAST_PASS) {
    if (tid < C * 3) s_head[tid] = p.whead[(tid / 3) * 4 + (tid % 3)];   // packed [16][3]
  } else if (tid < C * 4) {
    s_head[tid] = p.whead[tid];
  }
  {
    // stale shared memory may hold NaN patterns; rows outside the valid cone are multiplied (and ignored), so start clean
    const uint32_t ro = rings_offset(nl), n16 = (smem_bytes(nl) - ro) / 16;
    for (uint32_t i = tid; i < n16; i += NTHREADS) reinterpret_cast<uint4*>(smem + ro)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();   // barrier inits + zeros -> async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM);
  // Programmatic dependent launch (host side: cudaLaunchAttributeProgrammaticStreamSerialization).  Everything above read
  // only constants of the model; from here on the feature maps of the previous kernel are read (TMA) and the buffer it read
  // from is overwritten, so wait for it to complete -- after telling the scheduler that the NEXT kernel's CTAs may be
  // placed as soon as ours exit (they will wait at this same point).
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  if (tr && tid == 0) p.trace[257] = clock64();

  if (warp < EPI_WARPS) {
    // ================= epilogue warps =================
    const int quarter = warp & 3, set = warp >> 2;
    const int c = quarter * 32 + lane;
    EpiCtx E;
    E.tq = tmem + ((uint32_t)(quarter * 32) << 16);
    E.pix = (uint32_t)c * 16u;
    E.he = p.he; E.h_img = p.h; E.nl = nl;
    E.row_halves = (long long)p.vw * 16; E.row_out = (long long)p.w * 3;
    // this warp's tasks of a step: t = set, set + 4 (< 2 nl): row parity t / nl of layer (t + parity) % nl
    int tl[2], tpar[2], ntask = 0;
    for (int t = set; t < 2 * nl && ntask < 2; t += 4) { tpar[ntask] = t / nl; tl[ntask] = (t + tpar[ntask]) % nl; ++ntask; }
    // the layer whose MMAs are issued first in a step (order 3, 1, 2, 0) comes first
    auto issue_rank = [](int l) { return l == 3 ? 0 : (l == 1 ? 1 : (l == 2 ? 2 : 3)); };
    if (ntask == 2 && issue_rank(tl[1]) < issue_rank(tl[0])) {
      const int a_ = tl[0], b_ = tpar[0];
      tl[0] = tl[1]; tpar[0] = tpar[1]; tl[1] = a_; tpar[1] = b_;
    }
    float bias[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bias[i] = 0.f;
    for (int k = 0; k < ntask && !LAST_PASS; ++k)
      if (tl[k] & 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) bias[i] = s_bias[tl[k] * C + i];   // at most one conv_b layer per warp (nl = 2, 4)
      }
    if (LAST_PASS) {   // the head's share of the last conv_b's BN constant (the other conv_b layers read s_bias per task)
      for (int ch = 0; ch < C; ++ch) {
        const float bv = p.bias[(size_t)(2 * p.blk0 + nl - 1) * C + ch];
        bias[0] = fmaf(bv, p.whead[ch * 4 + 0], bias[0]);
        bias[1] = fmaf(bv, p.whead[ch * 4 + 1], bias[1]);
        bias[2] = fmaf(bv, p.whead[ch * 4 + 2], bias[2]);
      }
    }
    for (int blk = set; blk < 32; blk += 4) tmem_zero16(E.tq + blk * 16);
    tmem_wait_st();
    tc_fence_before();
    asm volatile("bar.sync 1, %0;\n" ::"r"(32 * (EPI_WARPS + 1)) : "memory");   // accumulators are zero -> MMA issuer
    uint32_t S = 0;
    long long gg = 0;
    for (long long a = r0; a < r1;) {
      const Seg sg = seg_at(p, a, r1);
      a += sg.yb - sg.ya;
      const int P = (sg.yb - sg.ya) + 2 * nl, Gm = (P + 1) >> 1, nsteps = Gm + LAG * (nl - 1) + 1;
      {
        const int vx = sg.j * p.tw - halo + c;                     // column of the virtual row
        const int vb = vx >= 0 ? vx / (p.we + 1) : -1, gx = vx - vb * (p.we + 1);   // image, column inside the image
        E.y00 = sg.ya - nl; E.P = P;
        E.col_ok = (vb >= 0) && (vb < p.n) && (gx < p.we);           // separator columns and the outside stay zero
        E.col_out = E.col_ok && (c >= halo) && (c < RW - halo) && (!LAST_PASS || gx < p.w);
        E.fout_col = p.fout + (((long long)E.y00 * p.vw + vx) << 4);
        E.out_col = reinterpret_cast<uint8_t*>(p.out) + ((((long long)vb * p.h + E.y00) * p.w + gx) * 3) * (p.out_u8 ? 1 : 4);
        E.gb0 = (int)(gg % K0);
      }
      for (int sr = 0; sr < nsteps; ++sr, ++S) {
        for (int k = 0; k < ntask; ++k) {
          const int l = tl[k], w = sr - LAG * l - 1;
          mbar_wait_sleep(bars + (BAR_MMA + 2u * (uint32_t)l + (S & 1u)) * 8, (S >> 1) & 1u);
          tc_fence_after();
          if (w < 0 || w >= Gm) continue;
          const int rho = 2 * w + tpar[k];
          if ((l & 1) == 0) epi_task<KIND_A, LAST_PASS>(p, R, E, bars, bias, s_head, s_bias + l * C, l, rho,
                                                        STREAM_TRACE_PTR(warp == 0 && lane == 0));
          else if (l + 1 < nl) epi_task<KIND_B_TO_X, LAST_PASS>(p, R, E, bars, bias, s_head, s_bias + l * C, l, rho);
          else epi_task<KIND_B_OUT, LAST_PASS>(p, R, E, bars, bias, s_head, s_bias + l * C, l, rho);
        }
        fence_async_smem();   // T / X stores of this step -> async proxy (tensor core reads)
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + (BAR_EPI + (S & 1u)) * 8);
      }
      gg += Gm;
    }
  } else if (warp == WARP_MMA) {
    // ================= MMA issuer =================
    // The issuing thread never touches shared memory: a completed mbarrier.try_wait on it costs ~360 cycles of tensor-pipe
    // bubble (tools/umma_probe3.cu: the load queues behind the operand fetches, and the MMA queue is shallow).  The waits
    // of step S are done by the helper warp, which then releases the issuer through a named barrier (bar.arrive /
    // bar.sync 2 + (S & 1): hardware barrier, no shared-memory traffic).
    asm volatile("bar.sync 1, %0;\n" ::"r"(32 * (EPI_WARPS + 1)) : "memory");   // accumulators are zero
    const uint32_t idesc0 = make_idesc_f16(128, 0);   // + (blocks * 2) << 17: N = 16 per accumulator block
    const uint32_t adesc_x0 = desc_lo(R.x0, R.x0_plane), adesc_t0 = desc_lo(R.t0, R.t_plane);
    const uint32_t adesc_x1 = desc_lo(R.x1, R.x1_plane), adesc_t1 = desc_lo(R.t1, R.t_plane);
    const uint32_t bdesc0 = desc_lo(s0 + SM_WTS, 48 * 16);
    const uint32_t bdesc_head = desc_lo(s0 + head_b_offset(nl), 16 * 16);   // [N 16][K 16]: the K halves are 256 B apart
    uint32_t S = 0;
    long long gg = 0;
    for (long long a = r0; a < r1;) {
      const Seg sg = seg_at(p, a, r1);
      a += sg.yb - sg.ya;
      const int P = (sg.yb - sg.ya) + 2 * nl, Gm = (P + 1) >> 1, nsteps = Gm + LAG * (nl - 1) + 1;
      const int gb0 = (int)(gg % K0);
      // per-layer issue state at the layer's group 0: A descriptor of input row 0 (pixel -1), its ring row, TMEM block of row -1
      uint32_t st_ad[MAX_NL];
      int st_slot[MAX_NL], st_blk[MAX_NL];
#pragma unroll
      for (int l = 0; l < MAX_NL; ++l) {
        const uint32_t adl = l == 0 ? adesc_x0 : (l == 1 ? adesc_t0 : (l == 2 ? adesc_x1 : adesc_t1));
        st_slot[l] = (l == 0) ? 2 * gb0 : 0;
        st_ad[l] = adl + (uint32_t)(st_slot[l] * RW - 1);
        st_blk[l] = (14 * l - 1) & 31;
      }
      // last pass: the residual of the last block reaches the (folded) head through one MMA per output row, A = row rho of
      // the block's input ring at pixel 0 (X1 for two blocks in the pass, X0 for one), B = the head matrix, N = 16
      auto head_x_desc = [&](int rho) -> uint32_t {
        return nl == 4 ? adesc_x1 + (uint32_t)((rho & (2 * KX - 1)) * RW)
                       : adesc_x0 + (uint32_t)((((gb0 + (rho >> 1)) % K0) * 2 + (rho & 1)) * RW);
      };
      for (int sr = 0; sr < nsteps; ++sr, ++S) {
        if (lane == 0) STREAM_TRACE(0);
        asm volatile("bar.sync %0, 64;\n" ::"r"(2 + (S & 1u)) : "memory");   // the helper has seen this step's barriers
        tc_fence_after();
        if (elect_one_sync()) {
          STREAM_TRACE(1);
          // issue order inside a step: conv_b layers first (3, 1, 2, 0) -- the layers of a step are independent of each
          // other, and the conv_b epilogues (residual load, global / head stores) are the long ones
#pragma unroll
          for (int li = 0; li < MAX_NL; ++li) {
            const int l = (li == 0) ? 3 : ((li == 1) ? 1 : ((li == 2) ? 2 : 0));
            if (l >= nl) continue;
            const int g = sr - LAG * l;
            const uint32_t mbar_l = bars + (BAR_MMA + 2u * (uint32_t)l + (S & 1u)) * 8;
            if (g < 0 || g >= Gm) { umma_commit(mbar_l); continue; }
            const uint32_t bd = bdesc0 + (uint32_t)(l * (W_LAYER_BYTES / 16));
            // incremental state of the layer (the issuing thread must not fall behind the shallow MMA queue: descriptor
            // arithmetic from scratch cost ~200 cycles per group of six MMAs, 40 % of the issue time)
            const uint32_t ad = st_ad[l];
            const int blk0 = st_blk[l];   // accumulator block of output row 2g - 1
            const int rho0 = 2 * g;
            // advance the state AFTER the MMAs are queued (the thread would otherwise block on the full queue anyway)
            auto advance = [&]() {
              st_blk[l] = (blk0 + 2) & 31;
              const int rows = (l == 0) ? 2 * K0 : ((l == 2) ? 2 * KX : 2 * KT);
              st_slot[l] += 2;
              st_ad[l] = ad + 2 * RW;
              if (st_slot[l] >= rows) { st_slot[l] -= rows; st_ad[l] -= (uint32_t)(rows * RW); }
            };
            if (rho0 >= 1 && rho0 + 2 < P && blk0 <= 28) {
              // fast path (7 groups in 8): both rows are interior rows of the segment and their four accumulator blocks
              // do not wrap around the TMEM ring -> six N = 48 MMAs, straight-line
              const uint32_t d = tmem + (uint32_t)blk0 * 16u;
              const uint32_t id = idesc0 + (6u << 17);
              mma_lo(d, ad, bd, id);
              mma_lo(d, ad + 1, bd + (48 * 16 * 2 / 16), id);
              mma_lo(d, ad + 2, bd + 2 * (48 * 16 * 2 / 16), id);
              if (LAST_PASS && l == nl - 1) mma_lo(d + 16, head_x_desc(rho0), bdesc_head, idesc0 + (2u << 17));
              mma_lo(d + 16, ad + RW, bd, id);
              mma_lo(d + 16, ad + RW + 1, bd + (48 * 16 * 2 / 16), id);
              mma_lo(d + 16, ad + RW + 2, bd + 2 * (48 * 16 * 2 / 16), id);
              if (LAST_PASS && l == nl - 1) mma_lo(d + 32, head_x_desc(rho0 + 1), bdesc_head, idesc0 + (2u << 17));
              umma_commit(mbar_l);
              advance();
              continue;
            }
#pragma unroll
            for (int par = 0; par < 2; ++par) {
              const int rho = rho0 + par;
              if (rho >= P) break;
              const uint32_t adr = ad + (uint32_t)(par * RW);
              // accumulator blocks of output rows rho-1, rho, rho+1 (B blocks 0, 1, 2); the segment's first / last input
              // row has no row above / below
              const int jlo = (rho == 0) ? 1 : 0, jhi = (rho == P - 1) ? 1 : 2;
              const int blk_lo = (blk0 + par + jlo) & 31, nb = jhi - jlo + 1;
              const int n1 = min(nb, 32 - blk_lo);
              {
                const uint32_t d = tmem + (uint32_t)blk_lo * 16u;
                const uint32_t b = bd + (uint32_t)(jlo * 16);
                const uint32_t id = idesc0 + ((uint32_t)(2 * n1) << 17);
                mma_lo(d, adr, b, id);
                mma_lo(d, adr + 1, b + (48 * 16 * 2 / 16), id);
                mma_lo(d, adr + 2, b + 2 * (48 * 16 * 2 / 16), id);
              }
              if (n1 < nb) {   // the blocks wrap around the TMEM ring: second part at column 0
                const uint32_t b = bd + (uint32_t)((jlo + n1) * 16);
                const uint32_t id = idesc0 + ((uint32_t)(2 * (nb - n1)) << 17);
                mma_lo(tmem, adr, b, id);
                mma_lo(tmem, adr + 1, b + (48 * 16 * 2 / 16), id);
                mma_lo(tmem, adr + 2, b + 2 * (48 * 16 * 2 / 16), id);
              }
              // residual -> head of output row rho (same place in the row's accumulation order as on the fast path)
              if (LAST_PASS && l == nl - 1)
                mma_lo(tmem + (uint32_t)((blk0 + par + 1) & 31) * 16u, head_x_desc(rho), bdesc_head, idesc0 + (2u << 17));
            }
            umma_commit(mbar_l);
            advance();
          }
          STREAM_TRACE(2);
          STREAM_TRACE(3);
        }
        __syncwarp();
      }
      gg += Gm;
    }
    if (tr && lane == 0) { p.trace[258] = clock64(); p.trace[259] = S; }
  } else if (warp == WARP_MMA + 1) {
    // ================= barrier helper of the MMA issuer =================
    // step S needs: epi_done(S-2) (lane 0: input rows written, accumulator blocks drained), x_full of layer 0's group
    // (lane 1), and at a segment start epi_done(S-1) too (lane 2: every accumulator block drained before the ring
    // restarts at row 0).  One barrier per lane, in parallel.
    uint32_t S = 0;
    long long gg = 0;
    for (long long a = r0; a < r1;) {
      const Seg sg = seg_at(p, a, r1);
      a += sg.yb - sg.ya;
      const int P = (sg.yb - sg.ya) + 2 * nl, Gm = (P + 1) >> 1, nsteps = Gm + LAG * (nl - 1) + 1;
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
  } else {
    // ================= TMA producer =================
    if (elect_one_sync()) {
      long long gg = 0;
      for (long long a = r0; a < r1;) {
        const Seg sg = seg_at(p, a, r1);
        a += sg.yb - sg.ya;
        const int P = (sg.yb - sg.ya) + 2 * nl, Gm = (P + 1) >> 1;
        const int gx0 = sg.j * p.tw - halo, y00 = sg.ya - nl;
        for (int g = 0; g < Gm; ++g, ++gg) {
          const uint32_t k = (uint32_t)(gg % K0), n = (uint32_t)(gg / K0);
          if (n >= 1) mbar_wait_sleep(bars + (BAR_XFREE + k) * 8, (n - 1) & 1u);
          const uint32_t bar = bars + (BAR_XFULL + k) * 8;
          mbar_arrive_expect_tx(bar, 2 * GROUP_BYTES);
          const uint32_t dst = R.x0 + k * GROUP_BYTES;
          tma_load_q(dst, &tmap, 0, gx0, y00 + 2 * g, 0, bar);   // gx0: column of the virtual row
          tma_load_q(dst + R.x0_plane, &tmap, 1, gx0, y00 + 2 * g, 0, bar);
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (tr && tid == 0) p.trace[260] = clock64();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

}  // namespace ustream

int run_fused_stack_stream(bfcnn_handle* h, const uint8_t* d_in, void* d_out, bool out_u8, const Extent& e, cudaStream_t st) {
  using namespace ustream;
  const int N = h->arch.no_layers;
  BF_REQUIRE(N >= 1, "the fused tensor-core stack needs no_layers >= 1 (use BFCNN_PREC_FP32)");
  static bool attr_set_dev[64] = {};   // function attributes are per device: one flag per device ordinal
  bool& attr_set = attr_set_dev[h->device & 63];
  if (!attr_set) {
    BF_CUDA(cudaFuncSetAttribute((const void*)stream_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    BF_CUDA(cudaFuncSetAttribute((const void*)stream_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    attr_set = true;
  }
  const int kb = 2;
  const int passes = (N + kb - 1) / kb;
  const long long vw = (long long)e.n * (e.we + 1);   // virtual row: the images side by side, a zero column after each
  BF_REQUIRE(vw < (1ll << 30), "batch too wide for the virtual row");
  const size_t feat_halves = (size_t)e.he * vw * C;
  BF_CHECK(h->ws_feat[1].reserve(feat_halves * sizeof(__half)));
  if (passes > 1) BF_CHECK(h->ws_feat[0].reserve(feat_halves * sizeof(__half)));
  // the separator columns sit at a regular stride of (we + 1) pixels: zero them in both maps (nothing else writes
  // them); the last one lies outside the tensor map, where the TMA fills in zeros, so one image needs no memset.  Done once
  // per buffer and geometry (feat_tag): as a per-call 2-D memset it cost ~0.35 ms per buffer on 4 x 4K frames, 7 % of the step
  for (int k = (passes > 1 ? 0 : 1); k < 2 && e.n > 1; ++k) {
    const unsigned long long tag = feat_layout_tag(1, h->ws_feat[k].p, e);
    if (h->feat_tag[k] == tag) continue;
    BF_CUDA(cudaMemset2DAsync(h->ws_feat[k].as<__half>() + (size_t)e.we * C, (size_t)(e.we + 1) * C * sizeof(__half), 0, C * sizeof(__half),
                              (size_t)e.he * e.n, st));
    h->feat_tag[k] = tag;
  }
  // pass "-1": base conv into ws_feat[1] (pass ps reads ws_feat[(ps-1)&1], writes ws_feat[ps&1])
  h->ktime_n = 0;
  ktime_begin(h, st, 0);
  BF_CHECK(launch_base_conv_f16(h, d_in, h->ws_feat[1].as<__half>(), e, st, nullptr, e.we + 1, vw));
  ktime_end(h, st);
  Extent ev = e;   // what the TMA sees: one "image" of he rows and vw columns
  ev.n = 1; ev.we = (int)vw - 1;
  for (int ps = 0; ps < passes; ++ps) {
    Params p;
    const bool last = (ps + 1 == passes);
    p.out = d_out;
    p.fin = h->ws_feat[(ps + 1) & 1].as<__half>();
    p.fout = last ? nullptr : h->ws_feat[ps & 1].as<__half>();
    p.wumma = h->d_conv_umma.as<uint8_t>();
    p.wlast = h->d_last_umma.as<uint8_t>();
    p.bias = h->d_bias_f32.as<float>();
    p.whead = h->d_head_f32.as<float>();
    p.n = e.n; p.h = e.h; p.w = e.w; p.he = e.he; p.we = e.we; p.vw = (int)vw;
    p.blk0 = ps * kb;
    p.nblk = std::min(kb, N - p.blk0);
    p.out_u8 = out_u8 ? 1 : 0;
    const int nl = 2 * p.nblk;
    p.tw = RW - 2 * nl;
    // Rows / columns this pass has to produce: the layers that are still to come (rem) shrink the cone of influence by one
    // pixel each, so beyond h + rem (w + rem) nothing can reach the cropped output any more (the band of the
    // pow2 canvas is only as wide as the receptive field that is LEFT).  What lies beyond keeps stale values: never read
    // by a valid output.
    const int rem = 2 * (N - p.blk0 - p.nblk);
    p.rows_needed = std::min(e.he, e.h + rem);
    // columns of the virtual row that need an output: up to the last needed column of the last image
    const long long cols_needed = (long long)(e.n - 1) * (e.we + 1) + std::min(e.we, e.w + rem);
    p.tiles_x = (int)((cols_needed + p.tw - 1) / p.tw);
    p.total_rows = (long long)p.tiles_x * p.rows_needed;
    p.seg_overhead = 2 * nl + 2 * (LAG * (nl - 1) + 1);
    const long long total_cost = (long long)p.tiles_x * ((long long)p.rows_needed + p.seg_overhead);
    int grid = (int)std::min<long long>(h->sm_count, std::max<long long>(1, p.total_rows / MIN_SHARE));
    p.share = (total_cost + grid - 1) / grid;
    grid = (int)((total_cost + p.share - 1) / p.share);
    const size_t smem = smem_bytes(nl);
    BF_REQUIRE(smem <= (size_t)MAX_SMEM, "internal: streaming pass does not fit in shared memory");
    static const int trace_on = getenv("BFCNN_STREAM_TRACE") ? atoi(getenv("BFCNN_STREAM_TRACE")) : 0;
    p.trace = nullptr; p.trace_block = 0;
    if (trace_on && ps == std::min(1, passes - 1)) {
      BF_CHECK(h->ws_feat[2].reserve(264 * sizeof(long long)));
      BF_CUDA(cudaMemsetAsync(h->ws_feat[2].p, 0, 264 * sizeof(long long), st));
      p.trace = h->ws_feat[2].as<long long>(); p.trace_block = grid / 2;
    }
    CUtensorMap tmap;
    BF_CHECK(make_feature_tmap(&tmap, p.fin, ev, RW, 2, vw));
    ktime_begin(h, st, last ? 2 : 1);
    {
      // programmatic dependent launch: the CTAs of this pass take their SMs as the previous kernel's CTAs exit and run
      // their prologue (barriers, TMEM, weights, zeroed rings) while its tail is still working; griddepcontrol.wait in the
      // kernel holds every access to the feature maps until the previous kernel has completed
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      if (last) BF_CUDA(cudaLaunchKernelEx(&cfg, stream_pass_kernel<true>, p, tmap));
      else BF_CUDA(cudaLaunchKernelEx(&cfg, stream_pass_kernel<false>, p, tmap));
    }
    ktime_end(h, st);
    h->launches++;
    BF_CUDA(cudaGetLastError());
    if (p.trace) {
      long long tb[264];
      BF_CUDA(cudaMemcpyAsync(tb, p.trace, sizeof(tb), cudaMemcpyDeviceToHost, st));
      BF_CUDA(cudaStreamSynchronize(st));
      const long long t0 = tb[0];
      fprintf(stderr, "[stream trace] kernel: setup %lld, issue loop end %lld (%lld steps, %.0f cycles/step), total %lld\n", tb[257] - tb[256],
              tb[258] - tb[256], tb[259], (double)(tb[258] - tb[257]) / (double)std::max(1ll, tb[259]), tb[260] - tb[256]);
      fprintf(stderr, "[stream trace] pass %d grid %d share %lld: step | loop top, after bar.sync | MMAs issued, committed | epi w0 task A: start, ld done | zero issued, sts issued\n", ps, grid, p.share);
      for (int i = 0; i < 32; ++i) {
        fprintf(stderr, "  %3u |", TRACE_S0 + i);
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %7lld%s", tb[i * 8 + k] ? tb[i * 8 + k] - t0 : -1ll, (k & 1) ? " |" : "");
        fprintf(stderr, "\n");
      }
    }
  }
  return BFCNN_OK;
}

}  // namespace bfcnn
