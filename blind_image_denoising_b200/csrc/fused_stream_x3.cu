// fused_stream_x3.cu -- the row-streaming tcgen05 stack of fused_stream.cu in the F16X3 arithmetic (FP32-grade results).
//
// Same pipeline (column strips streamed top to bottom, layer l+1 trailing layer l by LAG = 3 row groups, rings in shared
// memory, the 32 TMEM blocks as one ring, barrier-helper warp, per-layer mma_done), with activations and weights split into
// fp16 hi + lo parts: every product is issued as lo*hi + hi*lo + hi*hi (9 MMAs per row and conv; the last two share their A
// operand through the tensor core's A collector), the rings carry both
// parts (four channel-half planes), the feature map between passes carries both parts (the virtual-row map of
// fused_stream.cu twice: the lo map behind the hi map, a second "image" to the tensor map).
// One residual block (two convs) per pass: the rings of two blocks with both parts do not fit in shared memory.
// Kept in its own translation unit so that the F16 kernel's code generation is untouched; the shared pieces are in
// stream_common.cuh.
//
// Reference arithmetic: module_denoiser.py:53-73, backbone_blocks.py:167-246 (block), model.py:297-342 (head),
// utilities.py:435-443 (denormalise).
#include "stream_common.cuh"

namespace bfcnn {
namespace ustream3 {

using namespace stream;

constexpr int K0 = 9;                   // X0 ring: groups of 2 rows (TMA prefetch depth)

// barriers (8 B each)
// mma_done[l][s & 1] (per layer: the epilogue of layer l starts while the later layers of the step are still being
// multiplied), epi_done[s & 1], x_full[k], x_free[k]
constexpr uint32_t BAR_MMA = 0, BAR_EPI = 2 * 4, BAR_XFULL = BAR_EPI + 2, BAR_XFREE = BAR_XFULL + K0, NBARS = BAR_XFREE + K0;
constexpr uint32_t SM_BARS = 0, SM_TMEM = 512, SM_HEAD = 528, SM_BIAS = 784, SM_WTS = 1152;

// weights: [layer][hi / lo part][W_LAYER_BYTES]; rings: X0 and T0, each hi half 0, hi half 1, lo half 0, lo half 1
__host__ __device__ inline uint32_t rings_offset(int nl) { return (SM_WTS + (uint32_t)nl * 2 * W_LAYER_BYTES + 127u) & ~127u; }
__host__ __device__ inline uint32_t smem_bytes(int nl) { return rings_offset(nl) + 4 * plane_bytes_of(2 * K0) + 4 * plane_bytes_of(2 * KT); }

struct Params : Split {
  // Feature maps between passes: ONE virtual row per image row, the n images side by side, each followed by a zero column
  // (the "same" padding of both neighbours), hi part then lo part: [2 parts][he][vw = n (we + 1)][16] (see fused_stream.cu).
  long long plane_halves;  // halves between the hi and the lo part of a feature map (he * vw * 16)
  const __half* fin;       // [2 parts][he][vw][16]  input feature map of this pass
  __half* fout;            // [2 parts][he][vw][16]
  int vw;                  // n * (we + 1)
  void* out;               // [n][h][w][3] uint8 or float
  const uint8_t* wumma;    // [2N][hi / lo][W_LAYER_BYTES]
  const float* bias;       // [2N][16]
  const float* whead;      // [16][4]
  int n, h, w, he, we;
  int blk0, nblk;
  int out_u8;
  int tw;                  // output columns per strip
};

// rings: byte address of pixel 0, row slot 0, channel half 0; the other half is +plane
struct Rings {
  uint32_t x0, t0;             // hi part; the lo part is + 2 planes
  uint32_t x0_plane, t_plane;
};

// one (layer, row) task of one warp: 32 pixels of output row rho of layer l
template <int KIND, bool LAST_PASS>
__device__ __forceinline__ void epi_task(const Params& p, const Rings& R, const EpiCtx& E, uint32_t bars, const float (&bias)[16],
                                         const float* s_head, const float* s_bias_l, int l, int rho) {
  const uint32_t taddr = E.tq + (uint32_t)((rho + 14 * l) & 31) * 16u;
  uint32_t v[16];
  tmem_ld16_issue(taddr, v);
  const bool inside = E.col_ok && ((unsigned)(E.y00 + rho) < (unsigned)E.he);
  if (KIND == KIND_A) {
    tmem_ld_wait(v);
    tmem_zero16(taddr);
    const uint32_t m = inside ? 0xFFFFFFFFu : 0u;
    // T = ReLU(acc), split into the hi part and the lo part (the rounding error of the hi part), per channel half
    uint32_t hp[8], lp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float r0 = fmaxf(__uint_as_float(v[2 * i]), 0.f), r1 = fmaxf(__uint_as_float(v[2 * i + 1]), 0.f);
      const uint32_t hh = pack_h2(r0, r1);
      const float2 hf = unpack_h2(hh);
      hp[i] = hh & m;
      lp[i] = pack_h2(r0 - hf.x, r1 - hf.y) & m;
    }
    const uint32_t dst = R.t0 + ((uint32_t)rho % (2 * KT)) * ROW_BYTES + E.pix;
    sts128(dst, make_uint4(hp[0], hp[1], hp[2], hp[3]));
    sts128(dst + R.t_plane, make_uint4(hp[4], hp[5], hp[6], hp[7]));
    sts128(dst + 2 * R.t_plane, make_uint4(lp[0], lp[1], lp[2], lp[3]));
    sts128(dst + 3 * R.t_plane, make_uint4(lp[4], lp[5], lp[6], lp[7]));
  } else {
    // residual: X of this block (fp16) + the BN constant b' + the accumulator
    const uint32_t xsrc = R.x0 + (uint32_t)(((E.gb0 + (rho >> 1)) % K0) * 2 + (rho & 1)) * ROW_BYTES + E.pix;
    const uint4 xa = lds128(xsrc), xb = lds128(xsrc + R.x0_plane);
    const uint4 xla = lds128(xsrc + 2 * R.x0_plane), xlb = lds128(xsrc + 3 * R.x0_plane);
    tmem_ld_wait(v);
    tmem_zero16(taddr);
    float xf[16];   // X = hi + lo
    {
      const uint32_t xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      const uint32_t xl[8] = {xla.x, xla.y, xla.z, xla.w, xlb.x, xlb.y, xlb.z, xlb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 a2 = unpack_h2(xs[i]), b2 = unpack_h2(xl[i]);
        xf[2 * i] = a2.x + b2.x; xf[2 * i + 1] = a2.y + b2.y;
      }
    }
    if (KIND == KIND_B_OUT && LAST_PASS) {
      // collapsed 1x1 head + tanh(2y)*0.51 + denormalise (+ round + uint8), accumulated channel pair by channel pair
      // (the 19-warp CTA caps the kernel at 96 registers; bias and head weights stay in shared memory)
      if (E.col_out && inside && rho >= E.nl && rho < E.P - E.nl && (E.y00 + rho) < E.h_img) {
        // s_head is packed [16 ch][3] here (12 broadcast LDS.128 per thread) and the BN constant b' of the last conv_b
        // enters as bias[0..2] = sum_ch b'[ch] w[ch][j], folded once per thread
        float s0 = bias[0], s1 = bias[1], s2 = bias[2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {   // 4 channels = 12 weights = 3 float4 per iteration
          const float4 wa = *reinterpret_cast<const float4*>(s_head + 12 * q), wb = *reinterpret_cast<const float4*>(s_head + 12 * q + 4);
          const float4 wc = *reinterpret_cast<const float4*>(s_head + 12 * q + 8);
          const float f0 = __uint_as_float(v[4 * q]) + xf[4 * q], f1 = __uint_as_float(v[4 * q + 1]) + xf[4 * q + 1];
          const float f2 = __uint_as_float(v[4 * q + 2]) + xf[4 * q + 2], f3 = __uint_as_float(v[4 * q + 3]) + xf[4 * q + 3];
          s0 = fmaf(f0, wa.x, s0); s1 = fmaf(f0, wa.y, s1); s2 = fmaf(f0, wa.z, s2);
          s0 = fmaf(f1, wa.w, s0); s1 = fmaf(f1, wb.x, s1); s2 = fmaf(f1, wb.y, s2);
          s0 = fmaf(f2, wb.z, s0); s1 = fmaf(f2, wb.w, s1); s2 = fmaf(f2, wc.x, s2);
          s0 = fmaf(f3, wc.y, s0); s1 = fmaf(f3, wc.z, s1); s2 = fmaf(f3, wc.w, s2);
        }
        const float r0o = head_activation_fast(s0), r1o = head_activation_fast(s1), r2o = head_activation_fast(s2);
        store_rgb(E.out_col + (long long)rho * E.row_out * (p.out_u8 ? 1 : 4), p.out_u8, r0o, r1o, r2o);
      }
      if (l == 1) {
        fence_async_smem();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(bars + (BAR_XFREE + (uint32_t)((E.gb0 + (rho >> 1)) % K0)) * 8);
      }
      return;
    }
    // X_new = acc + X + b', split into hi + lo and written to the next pass's feature map (lo images follow the hi images)
    if (E.col_out && inside && rho >= E.nl && rho < E.P - E.nl) {
      uint32_t hp[8], lp[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float b0, b1;
        if (LAST_PASS) {
          const float2 bb = *reinterpret_cast<const float2*>(s_bias_l + 2 * i);
          b0 = bb.x; b1 = bb.y;
        } else {
          b0 = bias[2 * i]; b1 = bias[2 * i + 1];
        }
        const float f0 = __uint_as_float(v[2 * i]) + (xf[2 * i] + b0), f1 = __uint_as_float(v[2 * i + 1]) + (xf[2 * i + 1] + b1);
        const uint32_t hh = pack_h2(f0, f1);
        const float2 hf = unpack_h2(hh);
        hp[i] = hh;
        lp[i] = pack_h2(f0 - hf.x, f1 - hf.y);
      }
      __half* o = E.fout_col + (long long)rho * E.row_halves;
      stg256(o, make_uint4(hp[0], hp[1], hp[2], hp[3]), make_uint4(hp[4], hp[5], hp[6], hp[7]));
      stg256(o + p.plane_halves, make_uint4(lp[0], lp[1], lp[2], lp[3]), make_uint4(lp[4], lp[5], lp[6], lp[7]));
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
    R.x0_plane = plane_bytes_of(2 * K0); R.t_plane = plane_bytes_of(2 * KT);
    R.x0 = o + SLACK_PX * 16; o += 4 * R.x0_plane;
    R.t0 = o + SLACK_PX * 16;
  }
  long long r0, r1;
  cta_rows(p, r0, r1);

  // ---------------- one-time setup: barriers, TMEM, weights, zeroed rings
  init_barriers_and_tmem<BAR_EPI, BAR_XFULL, BAR_XFREE, NBARS>(bars, s0 + SM_TMEM, tid, warp);
  for (int i = tid; i < nl * 2 * (W_LAYER_BYTES / 16); i += NTHREADS)   // [layer][hi / lo][W_LAYER_BYTES], as host_pack.cu lays them out
    reinterpret_cast<uint4*>(smem + SM_WTS)[i] = reinterpret_cast<const uint4*>(p.wumma + (size_t)(2 * p.blk0) * 2 * W_LAYER_BYTES)[i];
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
          if ((l & 1) == 0) epi_task<KIND_A, LAST_PASS>(p, R, E, bars, bias, s_head, s_bias + l * C, l, rho);
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
    // The issuing thread never touches shared memory: the waits of step S are done by the helper warp
    // (stream_common.cuh::helper_warp_loop), which releases the issuer through a named barrier.
    asm volatile("bar.sync 1, %0;\n" ::"r"(32 * (EPI_WARPS + 1)) : "memory");   // accumulators are zero
    const uint32_t idesc0 = make_idesc_f16(128, 0);   // + (blocks * 2) << 17: N = 16 per accumulator block
    const uint32_t adesc_x0 = desc_lo(R.x0, R.x0_plane), adesc_t0 = desc_lo(R.t0, R.t_plane);
    const uint32_t alo_x0 = (2 * R.x0_plane) >> 4, alo_t0 = (2 * R.t_plane) >> 4;   // hi -> lo part of a ring, 16-byte units
    constexpr uint32_t BDX = 48 * 16 * 2 / 16, BLO = W_LAYER_BYTES / 16;                // next dx block / the lo weights
    const uint32_t bdesc0 = desc_lo(s0 + SM_WTS, 48 * 16);
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
        const uint32_t adl = (l == 0) ? adesc_x0 : adesc_t0;
        st_slot[l] = (l == 0) ? 2 * gb0 : 0;
        st_ad[l] = adl + (uint32_t)(st_slot[l] * RW - 1);
        st_blk[l] = (14 * l - 1) & 31;
      }
      for (int sr = 0; sr < nsteps; ++sr, ++S) {
        asm volatile("bar.sync %0, 64;\n" ::"r"(2 + (S & 1u)) : "memory");   // the helper has seen this step's barriers
        tc_fence_after();
        if (elect_one_sync()) {
          // issue order inside a step: conv_b layers first (3, 1, 2, 0) -- the layers of a step are independent of each
          // other, and the conv_b epilogues (residual load, global / head stores) are the long ones
#pragma unroll
          for (int li = 0; li < MAX_NL; ++li) {
            const int l = (li == 0) ? 3 : ((li == 1) ? 1 : ((li == 2) ? 2 : 0));
            if (l >= nl) continue;
            const int g = sr - LAG * l;
            const uint32_t mbar_l = bars + (BAR_MMA + 2u * (uint32_t)l + (S & 1u)) * 8;
            if (g < 0 || g >= Gm) { umma_commit(mbar_l); continue; }
            const uint32_t bd = bdesc0 + (uint32_t)(l * 2 * (W_LAYER_BYTES / 16));
            const uint32_t alo = (l == 0) ? alo_x0 : alo_t0;
            // one input row: lo*hi + hi*lo + hi*hi for each dx (the F16X3 arithmetic, FP32-grade products)
            auto row9 = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t id) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                mma_lo(d, a + alo + dx, b + dx * BDX, id);
                // hi*lo and hi*hi share their A operand: fetched from shared memory once and kept in the tensor core's A
                // collector for the second (12.5 instead of 16.5 KB of operand reads per dx: 25.4 -> 23.8 ms per step of
                // four 4K frames in one process, bit-identical results)
                mma_lo_fill(d, a + dx, b + BLO + dx * BDX, id);
                mma_lo_lastuse(d, a + dx, b + dx * BDX, id);
              }
            };
            // incremental state of the layer (the issuing thread must not fall behind the shallow MMA queue: descriptor
            // arithmetic from scratch cost ~200 cycles per group of six MMAs, 40 % of the issue time)
            const uint32_t ad = st_ad[l];
            const int blk0 = st_blk[l];   // accumulator block of output row 2g - 1
            const int rho0 = 2 * g;
            // advance the state AFTER the MMAs are queued (the thread would otherwise block on the full queue anyway)
            auto advance = [&]() {
              st_blk[l] = (blk0 + 2) & 31;
              const int rows = (l == 0) ? 2 * K0 : 2 * KT;
              st_slot[l] += 2;
              st_ad[l] = ad + 2 * RW;
              if (st_slot[l] >= rows) { st_slot[l] -= rows; st_ad[l] -= (uint32_t)(rows * RW); }
            };
            if (rho0 >= 1 && rho0 + 2 < P && blk0 <= 28) {
              // fast path (7 groups in 8): both rows are interior rows of the segment and their four accumulator blocks
              // do not wrap around the TMEM ring -> six N = 48 MMAs, straight-line
              const uint32_t d = tmem + (uint32_t)blk0 * 16u;
              const uint32_t id = idesc0 + (6u << 17);
              row9(d, ad, bd, id);
              row9(d + 16, ad + RW, bd, id);
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
                row9(d, adr, b, id);
              }
              if (n1 < nb) {   // the blocks wrap around the TMEM ring: second part at column 0
                const uint32_t b = bd + (uint32_t)((jlo + n1) * 16);
                const uint32_t id = idesc0 + ((uint32_t)(2 * (nb - n1)) << 17);
                row9(tmem, adr, b, id);
              }
            }
            umma_commit(mbar_l);
            advance();
          }
        }
        __syncwarp();
      }
      gg += Gm;
    }
  } else if (warp == WARP_MMA + 1) {
    helper_warp_loop<K0, BAR_EPI, BAR_XFULL>(p, bars, r0, r1, nl, lane);
  } else {
    if (elect_one_sync()) tma_producer_loop<K0, 2, BAR_XFULL, BAR_XFREE>(p, &tmap, bars, R.x0, R.x0_plane, r0, r1, nl, p.tw);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

}  // namespace ustream3

int run_fused_stack_stream_x3(bfcnn_handle* h, const uint8_t* d_in, void* d_out, bool out_u8, const Extent& e, cudaStream_t st) {
  using namespace ustream3;
  const int N = h->arch.no_layers;
  BF_REQUIRE(N >= 1, "the fused tensor-core stack needs no_layers >= 1 (use BFCNN_PREC_FP32)");
  static bool attr_set_dev[64] = {};   // function attributes are per device: one flag per device ordinal
  bool& attr_set = attr_set_dev[h->device & 63];
  if (!attr_set) {
    BF_CUDA(cudaFuncSetAttribute((const void*)stream_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    BF_CUDA(cudaFuncSetAttribute((const void*)stream_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    attr_set = true;
  }
  const int passes = N;   // one residual block per pass
  const long long vw = (long long)e.n * (e.we + 1);   // virtual row: the images side by side, a zero column after each
  BF_REQUIRE(vw < (1ll << 30), "batch too wide for the virtual row");
  const size_t feat_halves = (size_t)e.he * vw * C;
  BF_CHECK(h->ws_feat[1].reserve(feat_halves * 2 * sizeof(__half)));   // hi part, then lo part
  if (passes > 1) BF_CHECK(h->ws_feat[0].reserve(feat_halves * 2 * sizeof(__half)));
  // the separator columns sit at a regular stride of (we + 1) pixels, in the hi and in the lo part: zero them in both maps
  // (nothing else writes them); the last one lies outside the tensor map, where the TMA fills in zeros, so one image needs
  // no memset.  Once per buffer and geometry (feat_tag), see fused_stream.cu.
  for (int k = (passes > 1 ? 0 : 1); k < 2 && e.n > 1; ++k) {
    const unsigned long long tag = feat_layout_tag(2, h->ws_feat[k].p, e);
    if (h->feat_tag[k] == tag) continue;
    for (int part = 0; part < 2; ++part)
      BF_CUDA(cudaMemset2DAsync(h->ws_feat[k].as<__half>() + part * feat_halves + (size_t)e.we * C, (size_t)(e.we + 1) * C * sizeof(__half), 0,
                                C * sizeof(__half), (size_t)e.he * e.n, st));
    h->feat_tag[k] = tag;
  }
  // pass "-1": base conv (hi + lo) into ws_feat[1] (pass ps reads ws_feat[(ps-1)&1], writes ws_feat[ps&1])
  h->ktime_n = 0;
  ktime_begin(h, st, 0);
  BF_CHECK(launch_base_conv_f16(h, d_in, h->ws_feat[1].as<__half>(), e, st, h->ws_feat[1].as<__half>() + feat_halves, e.we + 1, vw));
  ktime_end(h, st);
  Extent e2 = e;   // what the TMA sees: two "images" (hi part, lo part) of he rows and vw columns
  e2.n = 2; e2.we = (int)vw - 1;
  for (int ps = 0; ps < passes; ++ps) {
    Params p;
    const bool last = (ps + 1 == passes);
    p.out = d_out;
    p.fin = h->ws_feat[(ps + 1) & 1].as<__half>();
    p.fout = last ? nullptr : h->ws_feat[ps & 1].as<__half>();
    p.plane_halves = (long long)feat_halves;
    p.wumma = h->d_conv_umma_x3.as<uint8_t>();
    p.bias = h->d_bias_f32.as<float>();
    p.whead = h->d_head_f32.as<float>();
    p.n = e.n; p.h = e.h; p.w = e.w; p.he = e.he; p.we = e.we; p.vw = (int)vw;
    p.blk0 = ps;
    p.nblk = 1;
    p.out_u8 = out_u8 ? 1 : 0;
    const int nl = 2;
    const int grid = plan_pass(p, p.tw, e, N, p.blk0, p.nblk, h->sm_count);
    const size_t smem = smem_bytes(nl);
    BF_REQUIRE(smem <= (size_t)MAX_SMEM, "internal: streaming pass does not fit in shared memory");
    CUtensorMap tmap;
    BF_CHECK(make_feature_tmap(&tmap, p.fin, e2, RW, 2, vw));
    ktime_begin(h, st, last ? 2 : 1);
    if (last) BF_CUDA(launch_pdl(stream_pass_kernel<true>, grid, NTHREADS, smem, st, p, tmap));
    else BF_CUDA(launch_pdl(stream_pass_kernel<false>, grid, NTHREADS, smem, st, p, tmap));
    ktime_end(h, st);
    h->launches++;
    BF_CUDA(cudaGetLastError());
  }
  return BFCNN_OK;
}

}  // namespace bfcnn
