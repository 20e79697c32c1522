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
//   * One elected thread issues every MMA; mbarriers per group of 3 rows couple it to the epilogue warps
//     (mma_done[g] via tcgen05.commit, epi_done[g] via mbarrier.arrive), so layer l+1 chases layer l down the
//     region a few rows behind and the tensor pipe never drains at a layer boundary.
//   * The kernel is persistent (one CTA per SM, regions round-robin).  In the last layer of a region every
//     epilogue thread, once it has consumed its pixel of X, fetches the same pixel of the NEXT region into the
//     freed slot with cp.async; to the MMA issuer the next region's first layer is just "one more layer", so
//     the tensor pipe does not drain between regions either and HBM latency hides behind the last conv.
//
// Reference arithmetic: module_denoiser.py:53-73, utilities.py:449-461 (normalise), backbone_resnet.py:258-262
// (base conv), backbone_blocks.py:167-246 (block), model.py:297-342 (head), utilities.py:435-443 (denormalise).
#include "kernels.cuh"
#include "umma_ptx.cuh"

namespace bfcnn {
namespace umma {

constexpr int RW = 128;                 // region width == UMMA M
constexpr int SLACK_PX = 8;             // pixels of slack before/after every plane (tap shift -1/+1)
constexpr int NSETS = 4;                // epilogue warp sets (4 warps each, one per TMEM lane quarter); 13 warps -> 128 registers
constexpr int EPI_WARPS = 4 * NSETS;
constexpr int NTHREADS = 32 * (EPI_WARPS + 1);   // + the MMA issuer warp
constexpr int MAX_RH = 30;              // (RH + 2) accumulator blocks of 16 columns <= 512 TMEM columns
constexpr int W_LAYER_BYTES = 3 * 48 * 16 * 2;   // B operand of one conv: [dx 3][N 48][K 16] fp16
constexpr int MAX_SMEM = 232448;
constexpr int MAX_LAYERS = 8;           // conv layers fused per pass

enum Epi { EPI_RELU_TO_T = 0, EPI_RES_TO_X = 1, EPI_RES_TO_GLOBAL = 2, EPI_RES_HEAD = 3 };

constexpr int GROUP = 6;                // region rows per barrier group: the MMA issuer pays ~300 cycles per mbarrier wait
                                        // (the tensor pipe drains behind it, tools/umma_probe3.cu), so it waits per group, not per row
constexpr int MAX_GROUPS = (MAX_RH + GROUP - 1) / GROUP;

struct Params {
  const __half* fin;       // [n][he][we][16]  input feature map of this pass (base conv output for pass 0)
  __half* fout;            // [n][he][we][16]
  void* out;               // [n][h][w][3] uint8 or float
  const uint8_t* wumma;    // [2N][W_LAYER_BYTES]
  const float* bias;       // [2N][16]
  const float* whead;      // [16][4]
  int n, h, w, he, we;
  int blk0, nblk;
  int last, out_u8;
  int rh, tw, th, tiles_x, tiles_y, regions;
  long long* trace;        // debug timeline of CTA `trace_block` (BFCNN_UMMA_TRACE=1), else nullptr
  int trace_block;
};

using namespace tc5;

// ---------------------------------------------------------------------------- shared-memory map
struct Smem {
  uint32_t bars;       // mma_done[16], epi_done[16] (8 B each)
  uint32_t wts;        // [nlayers][W_LAYER_BYTES]
  uint32_t X[2], T[2]; // byte address of pixel 0 of each channel-half plane
  uint32_t plane_bytes;
};
__host__ __device__ inline uint32_t plane_bytes_of(int rh) { return (uint32_t)(rh * RW + 2 * SLACK_PX) * 16u; }
constexpr uint32_t SM_BARS = 0, SM_TMEM = 512, SM_HEAD = 528, SM_BIAS = 784, SM_WTS = 784 + MAX_LAYERS * 64;  // 1296
__host__ __device__ inline uint32_t planes_offset(int nlayers) { return (SM_WTS + (uint32_t)nlayers * W_LAYER_BYTES + 127u) & ~127u; }

struct Region { int b, oy, ox; };
__device__ __forceinline__ Region region_of(const Params& p, int it, int halo) {
  Region r;
  const int tx = it % p.tiles_x;
  it /= p.tiles_x;
  const int ty = it % p.tiles_y;
  r.b = it / p.tiles_y;
  r.oy = ty * p.th - halo;
  r.ox = tx * p.tw - halo;
  return r;
}

// Per-thread constants of one region: everything that does not depend on the row is computed once, so that the row
// loop below is ~60 instructions per conv_a row and ~25 per conv_b row (it was 170: the epilogue warps, not the tensor
// pipe, set the pace -- profiles/r01_umma_ncu.md).
struct EpiCtx {
  uint32_t tq;            // TMEM address of this warp's lane quarter, column 0
  uint32_t x0, x1, t0, t1;  // shared byte addresses of this thread's pixel column in region row 0
  int oy, he;             // row r is inside the extent iff (unsigned)(oy + r) < he
  bool col_ok;            // this thread's column is inside the extent
  bool col_out;           // ... and inside the tile (output) columns [halo, RW - halo) (and inside the image for the head)
  __half* fout_col;       // feature-map address of (b, oy, gx); row r adds r * row_halves
  uint8_t* out_col;       // output address of (b, oy, gx) (uint8 or float)
  long long row_halves;   // we * 16
  long long row_out;      // w * 3 elements
  int h_img;              // rows of the image (head: gy < h)
  // next region (prefetch of this warp's quarter-row by TMA)
  int nb, noy, nox_q;     // image index, first row, first column of the quarter in the next region
  uint32_t nx0, nx1;      // shared destinations (region row 0) of the quarter in the two channel-half planes
};

template <int EPI, bool PREFETCH>
__device__ __forceinline__ void epilogue_layer(const Params& p, const CUtensorMap* tmap, const Smem& S, const EpiCtx& E,
                                               const float* s_bias_next, const float* s_head, int set, int l, uint32_t parity) {
  const int r_lo = l + 1, r_hi = p.rh - l - 1;
  float bias[16];
  if (EPI == EPI_RELU_TO_T) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 bq = *reinterpret_cast<const float4*>(s_bias_next + 4 * q);
      bias[4 * q] = bq.x; bias[4 * q + 1] = bq.y; bias[4 * q + 2] = bq.z; bias[4 * q + 3] = bq.w;
    }
  }
  int waited = -1;   // highest mma_done index already observed in this layer
  for (int r = set; r < p.rh; r += NSETS) {
    const int grp = r / GROUP;
    const int need = min(r + 1, p.rh - 1) / GROUP;   // row r is complete once input row r+1 has been multiplied
    if (need > waited) {
      mbar_wait(S.bars + (uint32_t)need * 8, parity);
      tc_fence_after();
      waited = need;
    }
    const uint32_t taddr = E.tq + (uint32_t)(r + 1) * 16;
    const uint32_t po = (uint32_t)r * (RW * 16);
    if (r >= r_lo && r < r_hi) {
      const bool inside = E.col_ok && ((unsigned)(E.oy + r) < (unsigned)E.he);
      uint32_t v[16];
      tmem_ld16_issue(taddr, v);
      if (EPI == EPI_RELU_TO_T) {
        // T = ReLU(D); D <- X + b' (the residual the following conv_b accumulates onto)
        const uint4 xa = lds128(E.x0 + po), xb = lds128(E.x1 + po);
        tmem_ld_wait(v);
        {
          float f[16];
          const uint32_t xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 xv = unpack_h2(xs[i]);
            f[2 * i] = xv.x + bias[2 * i];
            f[2 * i + 1] = xv.y + bias[2 * i + 1];
          }
          tmem_st16(taddr, f);
        }
        const uint32_t m = inside ? 0xFFFFFFFFu : 0u;
        uint4 lo, hi;
        lo.x = relu_h2(pack_h2(__uint_as_float(v[0]), __uint_as_float(v[1]))) & m; lo.y = relu_h2(pack_h2(__uint_as_float(v[2]), __uint_as_float(v[3]))) & m;
        lo.z = relu_h2(pack_h2(__uint_as_float(v[4]), __uint_as_float(v[5]))) & m; lo.w = relu_h2(pack_h2(__uint_as_float(v[6]), __uint_as_float(v[7]))) & m;
        hi.x = relu_h2(pack_h2(__uint_as_float(v[8]), __uint_as_float(v[9]))) & m; hi.y = relu_h2(pack_h2(__uint_as_float(v[10]), __uint_as_float(v[11]))) & m;
        hi.z = relu_h2(pack_h2(__uint_as_float(v[12]), __uint_as_float(v[13]))) & m; hi.w = relu_h2(pack_h2(__uint_as_float(v[14]), __uint_as_float(v[15]))) & m;
        sts128(E.t0 + po, lo);
        sts128(E.t1 + po, hi);
      } else {
        tmem_ld_wait(v);
        tmem_zero16(taddr);
        if (EPI == EPI_RES_HEAD) {   // collapsed 1x1 head + tanh(2y)*0.51 + denormalise (+ round + uint8)
          if (E.col_out && inside && (E.oy + r) < E.h_img) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int ch = 0; ch < 16; ++ch) {
              const float4 wv = *reinterpret_cast<const float4*>(s_head + ch * 4);
              const float fv = __uint_as_float(v[ch]);
              s0 = fmaf(fv, wv.x, s0); s1 = fmaf(fv, wv.y, s1); s2 = fmaf(fv, wv.z, s2);
            }
            const float r0o = head_activation(s0), r1o = head_activation(s1), r2o = head_activation(s2);
            if (p.out_u8) {
              uint8_t* d = E.out_col + (long long)r * E.row_out;
              d[0] = (uint8_t)__float2int_rn(r0o); d[1] = (uint8_t)__float2int_rn(r1o); d[2] = (uint8_t)__float2int_rn(r2o);
            } else {
              float* d = reinterpret_cast<float*>(E.out_col) + (long long)r * E.row_out;
              d[0] = r0o; d[1] = r1o; d[2] = r2o;
            }
          }
        } else {
          uint4 lo, hi;
          lo.x = pack_h2(__uint_as_float(v[0]), __uint_as_float(v[1])); lo.y = pack_h2(__uint_as_float(v[2]), __uint_as_float(v[3]));
          lo.z = pack_h2(__uint_as_float(v[4]), __uint_as_float(v[5])); lo.w = pack_h2(__uint_as_float(v[6]), __uint_as_float(v[7]));
          hi.x = pack_h2(__uint_as_float(v[8]), __uint_as_float(v[9])); hi.y = pack_h2(__uint_as_float(v[10]), __uint_as_float(v[11]));
          hi.z = pack_h2(__uint_as_float(v[12]), __uint_as_float(v[13])); hi.w = pack_h2(__uint_as_float(v[14]), __uint_as_float(v[15]));
          if (EPI == EPI_RES_TO_X) {
            const uint32_t m = inside ? 0xFFFFFFFFu : 0u;
            lo.x &= m; lo.y &= m; lo.z &= m; lo.w &= m; hi.x &= m; hi.y &= m; hi.z &= m; hi.w &= m;
            sts128(E.x0 + po, lo);
            sts128(E.x1 + po, hi);
          } else if (E.col_out && inside) {
            uint4* o = reinterpret_cast<uint4*>(E.fout_col + (long long)r * E.row_halves);
            o[0] = lo;
            o[1] = hi;
          }
        }
      }
    } else if (EPI == EPI_RES_HEAD || EPI == EPI_RES_TO_GLOBAL) {
      tmem_zero16(taddr);   // rows that fell out of the valid range hold stale partial sums: clean for the next region
    }
    fence_async_smem();   // T/X stores of this row -> async proxy (tensor core); X reads of this row -> before the TMA overwrite
    if (PREFETCH) {
      // X row r is dead for this region: one lane fetches the next region's quarter-row (2 halves x 512 B) by TMA; the
      // x_ready barrier of the row group completes when the bytes have landed
      __syncwarp();
      if (elect_one_sync()) {
        const uint32_t bar = S.bars + (uint32_t)(32 + grp) * 8;
        mbar_arrive_expect_tx(bar, 1024u);
        tma_load_q(E.nx0 + po, tmap, 0, E.nox_q, E.noy + r, E.nb, bar);
        tma_load_q(E.nx1 + po, tmap, 1, E.nox_q, E.noy + r, E.nb, bar);
      }
    }
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(S.bars + (uint32_t)(16 + grp) * 8);
  }
}

// ---------------------------------------------------------------------------- the pass kernel (persistent)
template <bool LAST_PASS>
__global__ void __launch_bounds__(NTHREADS, 1)
umma_pass_kernel(const Params p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: no BSSY/BSYNC around role branches
  const int nl = 2 * p.nblk, halo = 2 * p.nblk;
  const int ng = (p.rh + GROUP - 1) / GROUP;
  const bool tr = (p.trace != nullptr) && ((int)blockIdx.x == p.trace_block);
  if (tr && tid == 0) p.trace[0] = clock64();
  const uint32_t s0 = smem_u32(smem);
  Smem S;
  S.bars = s0 + SM_BARS; S.wts = s0 + SM_WTS;
  S.plane_bytes = plane_bytes_of(p.rh);
  {
    const uint32_t pl = s0 + planes_offset(nl);
    S.X[0] = pl + 0 * S.plane_bytes + SLACK_PX * 16; S.X[1] = pl + 1 * S.plane_bytes + SLACK_PX * 16;
    S.T[0] = pl + 2 * S.plane_bytes + SLACK_PX * 16; S.T[1] = pl + 3 * S.plane_bytes + SLACK_PX * 16;
  }
  uint8_t* g_planes = smem + planes_offset(nl);
  float* s_head = reinterpret_cast<float*>(smem + SM_HEAD);
  float* s_bias = reinterpret_cast<float*>(smem + SM_BIAS);

  // ---------------- one-time setup: barriers, TMEM, weights, the first region
  // [0,16) mma_done[g] (tcgen05.commit); [16,32) epi_done[g]: every thread of the 128-pixel row arrives once per row of
  // the group; [32,48) x_ready[g] (next region's X rows): one arrive.expect_tx per warp per row + the TMA bytes;
  // [48] the first region's staging (one arrive.expect_tx per warp)
  if (tid < 49) {
    const int gi = tid & 15;
    const int rows = max(0, min(GROUP, p.rh - gi * GROUP));
    const uint32_t cnt = tid < 16 ? 1u : (tid < 32 ? (uint32_t)max(1, 128 * rows) : (tid < 48 ? (uint32_t)max(1, 4 * rows) : (uint32_t)(NTHREADS / 32)));
    mbar_init(S.bars + tid * 8, cnt);
  }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(s0 + SM_TMEM) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  for (int i = tid; i < nl * (W_LAYER_BYTES / 16); i += NTHREADS)
    reinterpret_cast<uint4*>(smem + SM_WTS)[i] =
        reinterpret_cast<const uint4*>(p.wumma + (size_t)(2 * p.blk0) * W_LAYER_BYTES)[i];
  for (int i = tid; i < nl * C; i += NTHREADS) s_bias[i] = p.bias[(size_t)(2 * p.blk0) * C + i];
  if (tid < C * 4) s_head[tid] = p.whead[tid];
  // zero the plane slack (read by the -1/+1 tap shifts of the first / last row)
  if (tid < 4 * 2 * SLACK_PX) {
    const int pl = tid / (2 * SLACK_PX), k = tid % (2 * SLACK_PX);
    const uint32_t off = (uint32_t)pl * S.plane_bytes + (k < SLACK_PX ? (uint32_t)k * 16u : S.plane_bytes - (uint32_t)(2 * SLACK_PX - k) * 16u);
    *reinterpret_cast<uint4*>(g_planes + off) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();   // barrier inits + slack zeros -> async proxy
  __syncthreads();
  {
    // the first region: quarter-row boxes by TMA, spread over the warps' elected lanes, one barrier
    const Region g0 = region_of(p, (int)blockIdx.x, halo);
    const int nbox = p.rh * 8;   // rows x 4 quarters x 2 halves
    const uint32_t bar = S.bars + 48 * 8;
    if (elect_one_sync()) {
      int mine = 0;
      for (int i = warp; i < nbox; i += NTHREADS / 32) ++mine;
      mbar_arrive_expect_tx(bar, (uint32_t)mine * 512u);
      for (int i = warp; i < nbox; i += NTHREADS / 32) {
        const int hf = i & 1, q = (i >> 1) & 3, r = i >> 3;
        tma_load_q(S.X[hf] + (uint32_t)(r * RW + q * 32) * 16u, &tmap, hf, g0.ox + q * 32, g0.oy + r, g0.b, bar);
      }
    }
    mbar_wait(bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM);
  if (tr && tid == 0) p.trace[1] = clock64();

  if (warp < EPI_WARPS) {
    // ================= epilogue warps =================
    const int quarter = warp & 3, set = warp >> 2;
    const int c = quarter * 32 + lane;
    EpiCtx E;
    E.tq = tmem + ((uint32_t)(quarter * 32) << 16);
    E.x0 = S.X[0] + (uint32_t)c * 16u; E.x1 = S.X[1] + (uint32_t)c * 16u;
    E.t0 = S.T[0] + (uint32_t)c * 16u; E.t1 = S.T[1] + (uint32_t)c * 16u;
    E.nx0 = S.X[0] + (uint32_t)(quarter * 32) * 16u; E.nx1 = S.X[1] + (uint32_t)(quarter * 32) * 16u;
    E.he = p.he; E.h_img = p.h;
    E.row_halves = (long long)p.we * 16; E.row_out = (long long)p.w * 3;
    // zero this warp's share of the accumulator blocks, then release the MMA issuer
    for (int blk = set; blk < p.rh + 2; blk += NSETS) tmem_zero16(E.tq + blk * 16);
    tmem_wait_st();
    tc_fence_before();
    asm volatile("bar.sync 1, %0;\n" ::"r"(NTHREADS) : "memory");
    uint32_t L = 0;   // global layer counter == mbarrier phase index
    for (int it = (int)blockIdx.x; it < p.regions; it += (int)gridDim.x) {
      const Region g = region_of(p, it, halo);
      const bool has_next = (it + (int)gridDim.x) < p.regions;
      {
        const int gx = g.ox + c;
        E.oy = g.oy;
        E.col_ok = (gx >= 0) && (gx < p.we);
        E.col_out = E.col_ok && (c >= halo) && (c < RW - halo) && (!LAST_PASS || gx < p.w);
        E.fout_col = p.fout + ((((long long)g.b * p.he + g.oy) * p.we + gx) << 4);
        E.out_col = reinterpret_cast<uint8_t*>(p.out) + ((((long long)g.b * p.h + g.oy) * p.w + gx) * 3) * (p.out_u8 ? 1 : 4);
        if (has_next) {
          const Region gn = region_of(p, it + (int)gridDim.x, halo);
          E.nb = gn.b; E.noy = gn.oy; E.nox_q = gn.ox + quarter * 32;
        }
      }
      for (int l = 0; l < nl; ++l, ++L) {
        const float* s_bias_next = s_bias + (l + 1) * C;   // conv_a pre-loads the bias of the conv_b that follows
        if ((l & 1) == 0) {
          if (has_next && l == nl - 2) epilogue_layer<EPI_RELU_TO_T, true>(p, &tmap, S, E, s_bias_next, s_head, set, l, L & 1u);
          else epilogue_layer<EPI_RELU_TO_T, false>(p, &tmap, S, E, s_bias_next, s_head, set, l, L & 1u);
        } else if (l + 1 < nl) {
          epilogue_layer<EPI_RES_TO_X, false>(p, &tmap, S, E, s_bias_next, s_head, set, l, L & 1u);
        } else {
          epilogue_layer<LAST_PASS ? EPI_RES_HEAD : EPI_RES_TO_GLOBAL, false>(p, &tmap, S, E, s_bias_next, s_head, set, l, L & 1u);
        }
        if (tr && lane == 0 && quarter == 0 && L < 8) p.trace[40 + set * 16 + L] = clock64();
      }
    }
  } else {
    // ================= MMA issuer warp =================
    asm volatile("bar.sync 1, %0;\n" ::"r"(NTHREADS) : "memory");   // accumulators are zero
    tc_fence_after();
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc_f16(128, 48);
      // descriptors with the start-address field at pixel 0 / layer 0; the per-MMA part is an add of 16-byte units
      const uint64_t adesc_x = make_desc(S.X[0], S.plane_bytes, 128), adesc_t = make_desc(S.T[0], S.plane_bytes, 128);
      const uint64_t bdesc0 = make_desc(S.wts, 48 * 16, 128);
      uint32_t L = 0;
      int nreg = 0;   // regions finished by this CTA
      for (int it = (int)blockIdx.x; it < p.regions; it += (int)gridDim.x, ++nreg) {
        for (int l = 0; l < nl; ++l, ++L) {
          const uint64_t ad0 = (l & 1) ? adesc_t : adesc_x;
          const uint64_t bd0 = bdesc0 + (uint64_t)(l * (W_LAYER_BYTES / 16));
          const int q_lo = l, q_hi = p.rh - l;
          if (tr && L < 8) p.trace[8 + 2 * L] = clock64();
          for (int grp = 0; grp < ng; ++grp) {
            if (L > 0) {
              const long long w0 = tr ? clock64() : 0;
              const uint32_t par = (L - 1) & 1u;
              if (grp == 0) mbar_wait(S.bars + (uint32_t)(16 + 0) * 8, par);
              if (grp + 1 < ng) mbar_wait(S.bars + (uint32_t)(16 + grp + 1) * 8, par);
              if (l == 0) {   // first layer of a later region: its X rows were fetched (cp.async) during the previous region
                const uint32_t xpar = (uint32_t)(nreg - 1) & 1u;
                if (grp == 0) mbar_wait(S.bars + (uint32_t)(32 + 0) * 8, xpar);
                if (grp + 1 < ng) mbar_wait(S.bars + (uint32_t)(32 + grp + 1) * 8, xpar);
                fence_async_smem();   // generic-proxy writes observed through the mbarrier -> async-proxy reads of the MMA
              }
              tc_fence_after();
              if (tr && L < 8) p.trace[24 + L] += clock64() - w0;
            }
            const int row_end = min(grp * GROUP + GROUP, q_hi);
            for (int q = max(grp * GROUP, q_lo); q < row_end; ++q) {
              const uint64_t ad = ad0 + (uint64_t)(q * RW - 1);
              const uint32_t d = tmem + (uint32_t)q * 16;
              mma_f16_ss(d, ad, bd0, idesc, 1u);
              mma_f16_ss(d, ad + 1, bd0 + (48 * 16 * 2 / 16), idesc, 1u);
              mma_f16_ss(d, ad + 2, bd0 + 2 * (48 * 16 * 2 / 16), idesc, 1u);
            }
            umma_commit(S.bars + (uint32_t)grp * 8);
          }
          if (tr && L < 8) p.trace[9 + 2 * L] = clock64();
        }
      }
    }
    __syncwarp();
  }
  if (tr && tid == 0) p.trace[2] = clock64();
  tc_fence_before();
  __syncthreads();
  if (tr && tid == 0) p.trace[3] = clock64();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------- base conv -> fp16 NHWC16
// normalise (utilities.py:449-461) + base conv k0 x k0, 3 -> 16 (backbone_resnet.py:258-262) over the work extent.
// CTA tile 64 x 16 pixels; the uint8 halo tile goes to shared memory already normalised (exactly the reference's
// x/255 - 0.5 in fp32): 0 outside the work extent (zero padding of the NORMALISED tensor), -0.5 on the raw-zero pow2
// canvas (utilities.py:749).  Each thread owns 4 consecutive pixels x 16 cout.
constexpr int BC_W = 64, BC_H = 16;
__global__ void __launch_bounds__(256, 2)
base_conv_f16_kernel(const uint8_t* __restrict__ img, __half* __restrict__ out, const float* __restrict__ w,
                     int h, int wd, int he, int we, int k0, long long img_stride, long long row_stride) {
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
    uint4* o = reinterpret_cast<uint4*>(out + (((long long)b * img_stride + (long long)gy * row_stride + gx) << 4));
    o[0] = lo;
    o[1] = hi;
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

}  // namespace umma

// ------------------------------------------------------------------------------------
// host: pass planning
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


static int env_int_u(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

int launch_base_conv_f16(bfcnn_handle* h, const uint8_t* d_in, __half* feat, const Extent& e, cudaStream_t st, __half* feat_lo,
                         long long img_stride, long long row_stride) {
  using namespace umma;
  if (img_stride == 0) { img_stride = (long long)e.he * e.we; row_stride = e.we; }   // [n][he][we][16]
  const int k0 = h->arch.base_kernel, r0 = (k0 - 1) / 2;
  const size_t bsm = (size_t)(k0 * k0 * 3 * C + (BC_H + 2 * r0) * (BC_W + 2 * r0) * 3) * sizeof(float);
  dim3 grid((e.we + BC_W - 1) / BC_W, (e.he + BC_H - 1) / BC_H, e.n);
  BF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "image too large for the base conv grid");
  static const int ffma_only = env_int_u("BFCNN_BASE_FFMA", 0);
  if (k0 == 3 && !ffma_only) {
    const int tx = (e.we + BM_W - 1) / BM_W, ty = (e.he + BM_H - 1) / BM_H;
    const long long tiles = (long long)tx * ty * e.n;
    BF_REQUIRE(tiles < (1ll << 31), "too many base-conv tiles");
    const int g3 = (int)std::min<long long>(tiles, 4ll * h->sm_count);
    base_conv3_mma_kernel<<<g3, 256, 0, st>>>(d_in, feat, feat_lo, h->d_base_f32.as<float>(), e.h, e.w, e.he, e.we, tx, ty, (int)tiles,
                                              img_stride, row_stride);
  } else {
    base_conv_f16_kernel<<<grid, 256, bsm, st>>>(d_in, feat, h->d_base_f32.as<float>(), e.h, e.w, e.he, e.we, k0, img_stride, row_stride);
  }
  h->launches++;
  BF_CUDA(cudaGetLastError());
  return BFCNN_OK;
}

int run_fused_stack_umma(bfcnn_handle* h, const uint8_t* d_in, void* d_out, bool out_u8, const Extent& e, cudaStream_t st) {
  using namespace umma;
  const int engine_regions = env_int_u("BFCNN_UMMA_REGIONS", 0);   // 1: the region kernel below (kept for A/B runs)
  if (!engine_regions && h->arch.no_layers >= 1) return run_fused_stack_stream(h, d_in, d_out, out_u8, e, st);
  const int N = h->arch.no_layers, k0 = h->arch.base_kernel;
  if (N < 1) {
    set_error("the fused tensor-core stack needs no_layers >= 1 (use BFCNN_PREC_FP32)");
    return BFCNN_ERR_UNSUPPORTED;
  }
  static bool attr_set_dev[64] = {};   // function attributes are per device: one flag per device ordinal
  bool& attr_set = attr_set_dev[h->device & 63];
  if (!attr_set) {
    BF_CUDA(cudaFuncSetAttribute((const void*)umma_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    BF_CUDA(cudaFuncSetAttribute((const void*)umma_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    attr_set = true;
  }
  int kb = env_int_u("BFCNN_KB_UMMA", 2);
  kb = std::max(1, std::min(std::min(kb, N), MAX_LAYERS / 2));
  const int passes = (N + kb - 1) / kb;
  const size_t feat_halves = (size_t)e.n * e.he * e.we * C;
  BF_CHECK(h->ws_feat[1].reserve(feat_halves * sizeof(__half)));
  if (passes > 1) BF_CHECK(h->ws_feat[0].reserve(feat_halves * sizeof(__half)));

  // pass "-1": base conv into ws_feat[1] (pass ps reads ws_feat[(ps-1)&1], writes ws_feat[ps&1])
  BF_CHECK(launch_base_conv_f16(h, d_in, h->ws_feat[1].as<__half>(), e, st));
  for (int ps = 0; ps < passes; ++ps) {
    Params p;
    p.out = d_out;
    p.fin = h->ws_feat[(ps + 1) & 1].as<__half>();
    p.fout = (ps + 1 < passes) ? h->ws_feat[ps & 1].as<__half>() : nullptr;
    p.wumma = h->d_conv_umma.as<uint8_t>();
    p.bias = h->d_bias_f32.as<float>();
    p.whead = h->d_head_f32.as<float>();
    p.n = e.n; p.h = e.h; p.w = e.w; p.he = e.he; p.we = e.we;
    p.blk0 = ps * kb;
    p.nblk = std::min(kb, N - p.blk0);
    p.last = (ps + 1 == passes); p.out_u8 = out_u8 ? 1 : 0;
    const int nl = 2 * p.nblk, halo = 2 * p.nblk;
    // shared-memory budget -> region rows (<= MAX_RH by TMEM capacity), a whole number of row groups when possible
    const size_t fixed = planes_offset(nl) + 64;
    int rh_max = MAX_RH;
    while (rh_max > 0 && fixed + (size_t)4 * plane_bytes_of(rh_max) > (size_t)MAX_SMEM) --rh_max;
    const int rows_needed = p.last ? e.h : e.he;
    const int cols_needed = p.last ? e.w : e.we;
    int th_max = rh_max - 2 * halo;
    p.tw = RW - 2 * halo;
    if (th_max < 1 || p.tw < 1) {
      set_error("fused tcgen05 pass does not fit (kb=%d)", kb);
      return BFCNN_ERR_INTERNAL;
    }
    if (rows_needed > th_max && (rh_max / GROUP) * GROUP - 2 * halo >= 1) th_max = (rh_max / GROUP) * GROUP - 2 * halo;
    p.tiles_y = (rows_needed + th_max - 1) / th_max;
    p.th = (rows_needed + p.tiles_y - 1) / p.tiles_y;   // balance the tile rows
    p.rh = p.th + 2 * halo;
    p.tiles_x = (cols_needed + p.tw - 1) / p.tw;
    const size_t smem = planes_offset(nl) + (size_t)4 * plane_bytes_of(p.rh);
    if (smem > (size_t)MAX_SMEM || p.rh > MAX_RH) {
      set_error("internal: tcgen05 pass smem %zu rh %d", smem, p.rh);
      return BFCNN_ERR_INTERNAL;
    }
    const long long regions = (long long)p.tiles_x * p.tiles_y * e.n;
    BF_REQUIRE(regions < (1ll << 31), "too many tiles");
    p.regions = (int)regions;
    const int grid = (int)std::min<long long>(regions, h->sm_count);
    static const int trace_on = env_int_u("BFCNN_UMMA_TRACE", 0);
    p.trace = nullptr; p.trace_block = 0;
    if (trace_on && ps == std::min(1, passes - 1)) {
      BF_CHECK(h->ws_feat[2].reserve(256 * sizeof(long long)));
      BF_CUDA(cudaMemsetAsync(h->ws_feat[2].p, 0, 256 * sizeof(long long), st));
      p.trace = h->ws_feat[2].as<long long>(); p.trace_block = grid / 2;
    }
    CUtensorMap tmap;
    BF_CHECK(make_feature_tmap(&tmap, p.fin, e, 32, 1));
    if (p.last) umma_pass_kernel<true><<<(unsigned)grid, NTHREADS, smem, st>>>(p, tmap);
    else umma_pass_kernel<false><<<(unsigned)grid, NTHREADS, smem, st>>>(p, tmap);
    if (p.trace) {
      long long t[256];
      BF_CUDA(cudaMemcpyAsync(t, p.trace, sizeof(t), cudaMemcpyDeviceToHost, st));
      BF_CUDA(cudaStreamSynchronize(st));
      fprintf(stderr, "[umma trace] pass %d rh %d nl %d regions %d grid %d: setup %lld total %lld (%.0f per region)\n", ps, p.rh, nl,
              p.regions, grid, t[1] - t[0], t[3] - t[0], (double)(t[3] - t[0]) / ((p.regions + grid - 1) / grid));
      for (int L = 0; L < 8; ++L) {
        fprintf(stderr, "  layer %d: mma issue [%lld .. %lld] waited %lld |", L, t[8 + 2 * L] - t[0], t[9 + 2 * L] - t[0], t[24 + L]);
        for (int s = 0; s < NSETS; ++s) fprintf(stderr, " epi%d end %lld", s, t[40 + s * 16 + L] - t[0]);
        fprintf(stderr, "\n");
      }
    }
    h->launches++;
    BF_CUDA(cudaGetLastError());
  }
  return BFCNN_OK;
}

}  // namespace bfcnn
