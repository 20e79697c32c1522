// fused_stream.cu -- the fused conv-BN-ReLU residual stack on tcgen05 as a ROW-STREAMING pipeline (sm_100a).
//
// M = 128 pixels of one image row per MMA, activations as channel-half planes in the SWIZZLE_NONE K-major layout so that a dx
// tap is a descriptor shift, N = 48 dy-scatter into three accumulator blocks.  The work unit is not a
// 128 x rh region with a vertical halo that every layer recomputes (the first engine of round 1, removed).  A CTA
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
#include "stream_common.cuh"

namespace bfcnn {
namespace ustream {

using namespace stream;

constexpr int K0 = 11;                  // X0 ring: groups of 2 rows (TMA prefetch depth)
constexpr int KX = 8;                   // X1 ring groups: written at step w+4, last read (residual) at step w+10

// barriers (8 B each)
// mma_done[l][s & 1] (per layer: the epilogue of layer l starts while the later layers of the step are still being
// multiplied), epi_done[s & 1], x_full[k], x_free[k]
constexpr uint32_t BAR_MMA = 0, BAR_EPI = 2 * 4, BAR_XFULL = BAR_EPI + 2, BAR_XFREE = BAR_XFULL + K0, NBARS = BAR_XFREE + K0;
constexpr uint32_t SM_BARS = 0, SM_TMEM = 512, SM_HEAD = 528, SM_BIAS = 784, SM_WTS = 1152;

constexpr uint32_t HEAD_B_BYTES = 16 * 16 * 2;   // B operand of the head matrix [N 16][K 16] fp16 (last pass)
__host__ __device__ inline uint32_t head_b_offset(int nl) { return SM_WTS + (uint32_t)nl * W_LAYER_BYTES; }
__host__ __device__ inline uint32_t rings_offset(int nl) { return (head_b_offset(nl) + HEAD_B_BYTES + 127u) & ~127u; }
__host__ __device__ inline uint32_t smem_bytes(int nl) {
  uint32_t b = rings_offset(nl) + 2 * plane_bytes_of(2 * K0) + 2 * plane_bytes_of(2 * KT);
  if (nl > 2) b += 2 * plane_bytes_of(2 * KX) + 2 * plane_bytes_of(2 * KT);
  return b;
}

struct Params : Split {
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
  int tw;                  // output columns per strip
  long long* trace;        // debug timeline (BFCNN_STREAM_TRACE=1) of CTA trace_block, steps [TRACE_S0, TRACE_S0 + 32)
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

// rings: byte address of pixel 0, row slot 0, channel half 0; the other half is +plane
struct Rings {
  uint32_t x0, t0, x1, t1;
  uint32_t x0_plane, t_plane, x1_plane;
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
      store_rgb(E.out_col + (long long)rho * E.row_out * (p.out_u8 ? 1 : 4), p.out_u8, r0o, r1o, r2o);
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
  long long r0, r1;
  cta_rows(p, r0, r1);
  const bool tr = (p.trace != nullptr) && ((int)blockIdx.x == p.trace_block);
  if (tr && tid == 0) p.trace[256] = clock64();

  // ---------------- one-time setup: barriers, TMEM, weights, zeroed rings
  init_barriers_and_tmem<BAR_EPI, BAR_XFULL, BAR_XFREE, NBARS>(bars, s0 + SM_TMEM, tid, warp);
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
  pdl_wait_for_previous();   // programmatic dependent launch: everything above read only constants of the model
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
      int P, Gm, nsteps;
      seg_steps(sg.yb - sg.ya, nl, P, Gm, nsteps);
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
      int P, Gm, nsteps;
      seg_steps(sg.yb - sg.ya, nl, P, Gm, nsteps);
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
    helper_warp_loop<K0, BAR_EPI, BAR_XFULL>(p, bars, r0, r1, nl, lane);
  } else {
    if (elect_one_sync()) tma_producer_loop<K0, 1, BAR_XFULL, BAR_XFREE>(p, &tmap, bars, R.x0, R.x0_plane, r0, r1, nl, p.tw);
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
    const int grid = plan_pass(p, p.tw, e, N, p.blk0, p.nblk, h->sm_count);
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
    if (last) BF_CUDA(launch_pdl(stream_pass_kernel<true>, grid, NTHREADS, smem, st, p, tmap));
    else BF_CUDA(launch_pdl(stream_pass_kernel<false>, grid, NTHREADS, smem, st, p, tmap));
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
