// umma_probe.cu -- microbenchmark + addressing check that fixed the design of the tcgen05 conv stack
// (DESIGN.md section 4).  Not part of the product library.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu && ./umma_probe
//
// Questions answered on a B200:
//   Q1  Does a SWIZZLE_NONE K-major A descriptor accept an arbitrary 16 B-aligned start address, so that a
//       3x3 tap shift is just "start += (dy*pitch+dx)*16 B" over a channel-half-planar pixel tile?
//   Q2  Does D += A*B accumulate on top of values put in TMEM with tcgen05.st (residual stream in TMEM)?
//   Q3  Cycles per tcgen05.mma (M=128, K=16, A and B from shared memory) for N = 16..256.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // layout_type = 0: SWIZZLE_NONE
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                    // D = F32
  d |= 0u << 7;                    // A = F16
  d |= 0u << 10;                   // B = F16
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
  return pred;
}
__device__ __forceinline__ void commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar), "r"(cnt));
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]));
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

constexpr int PITCH = 128;      // pixels per region row == one M tile
constexpr int ROWS = 26;
constexpr int NPIX = PITCH * (ROWS + 2) + 16;
constexpr int PLANE_BYTES = NPIX * 16;
constexpr int B_BYTES = 9 * 16 * 16 * 2;   // 9 taps [cout 16][cin 16] f16, stored as 3 (dx) x N=48 K-major blocks
constexpr int SMEM_BYTES = 2 * PLANE_BYTES + 2 * B_BYTES + 1024;

struct Params {
  const __half* act;   // [NPIX][16]
  const __half* wts;   // [9 taps: dy*3+dx][16 cout][16 cin]
  float* out;          // [ROWS][128][16] conv result (check mode)
  long long* cycles;   // [grid]
  int mode;            // 0 = correctness (N=48 dy-fused conv over rows), 1 = throughput
  int n_dim;           // throughput: N
  int mmas;            // throughput: MMAs per repetition
  int reps;
  int shift_px;        // throughput: pixel shift between successive MMAs' A start (0 => same address)
  int m_dim;           // throughput: M (64 or 128)
  int n2;              // throughput: if > 0 every other MMA uses N = n2 and a shifted A
  int preload;         // correctness: 1 => preload D with a constant via tcgen05.st and accumulate on top
  int group;           // throughput: consecutive MMAs sharing one D block (1 = every MMA its own block)
  int commit_every;    // throughput: tcgen05.commit to a scratch mbarrier every this many MMAs (0 = never)
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ __align__(8) uint64_t s_bar2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* pl0 = smem;
  uint8_t* pl1 = smem + PLANE_BYTES;
  uint8_t* sB = smem + 2 * PLANE_BYTES;   // [dx 3][kchunk 2][n 48][8 k] f16  (core matrix = 8 n x 16 B)

  // activations: plane0 = channels 0..7, plane1 = channels 8..15, 16 B per pixel each
  for (int i = tid; i < NPIX; i += 128) {
    const uint4* s = reinterpret_cast<const uint4*>(p.act + (size_t)i * 16);
    reinterpret_cast<uint4*>(pl0)[i] = s[0];
    reinterpret_cast<uint4*>(pl1)[i] = s[1];
  }
  // weights: B block for dx holds N = 48 rows n = j*16 + cout with j = 0,1,2 <-> dy = +1, 0, -1
  // (input row q contributes to out row q-dy: D column block j=0 is out row q-1 (dy=+1), j=1 row q, j=2 row q+1 (dy=-1))
  for (int i = tid; i < 3 * 2 * 48 * 8; i += 128) {
    const int k8 = i & 7, n = (i >> 3) % 48, kc = (i / (8 * 48)) & 1, dx = i / (8 * 48 * 2);
    const int j = n / 16, co = n % 16, dy = 1 - j;
    const int tap = (dy + 1) * 3 + dx;
    reinterpret_cast<__half*>(sB)[i] = p.wts[(tap * 16 + co) * 16 + kc * 8 + k8];
  }
  if (tid == 0) { mbar_init(smem_u32(&s_bar), 1); mbar_init(smem_u32(&s_bar2), 1); }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t a0 = smem_u32(pl0), b0 = smem_u32(sB), bar = smem_u32(&s_bar), bar2 = smem_u32(&s_bar2);
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  if (p.mode == 0) {
    // out rows r = 0..ROWS-1 live at tmem columns (r+1)*16 .. ; rows -1 and ROWS are scratch (cols 0 and (ROWS+1)*16)
    if (p.preload) {
      uint32_t v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(1000.0f + (float)c);
      for (int r = 0; r < ROWS + 2; ++r) tmem_st16(tmem + lane_base + r * 16, v);
    } else {
      uint32_t v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = 0u;
      for (int r = 0; r < ROWS + 2; ++r) tmem_st16(tmem + lane_base + r * 16, v);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (warp == 0 && elect_one_sync()) {
      const uint32_t idesc = make_idesc(128, 48);
      // input rows q = 0..ROWS-1 (smem row q+1; smem rows 0 and ROWS+1 are zero padding supplied by the host)
      for (int q = 0; q < ROWS; ++q) {
        for (int dx = 0; dx < 3; ++dx) {
          // A rows m = 0..127 <-> pixel (q, m + dx - 1): start shifted by one pixel = 16 B
          const uint32_t astart = a0 + (uint32_t)(((q + 1) * PITCH + 8 + dx - 1) * 16);
          const uint64_t ad = make_desc(astart, PLANE_BYTES, 128);
          const uint64_t bd = make_desc(b0 + dx * (2 * 48 * 16), 48 * 16, 128);
          mma(tmem + q * 16, ad, bd, idesc, 1u);   // columns q*16 .. q*16+47 = out rows q-1, q, q+1
        }
      }
      commit(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    for (int r = 0; r < ROWS; ++r) {
      uint32_t v[16];
      tmem_ld16(tmem + lane_base + (r + 1) * 16, v);
      float* o = p.out + ((size_t)r * 128 + tid) * 16;
#pragma unroll
      for (int c = 0; c < 16; ++c) o[c] = __uint_as_float(v[c]);
    }
  } else {
    if (warp == 0 && elect_one_sync()) {
      const uint32_t idesc = make_idesc(p.m_dim, p.n_dim);
      const uint32_t idesc2 = make_idesc(p.m_dim, p.n2 > 0 ? p.n2 : p.n_dim);
      uint32_t parity = 0;
      const uint64_t ad0 = make_desc(a0 + PITCH * 16, PLANE_BYTES, 128);
      const uint64_t bd = make_desc(b0, (uint32_t)p.n_dim * 16, 128);
      const uint32_t sh = (uint32_t)p.shift_px;   // in 16 B units == pixels
      for (int rep = -1; rep < p.reps; ++rep) {
        long long t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < p.mmas; i += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint32_t j = (uint32_t)(i + u);
            const uint32_t k = (p.group > 1 ? j / (uint32_t)p.group : j) & 15;
            if ((u & 1) && p.n2 > 0) mma(tmem + k * 16, ad0 + (uint64_t)(k * sh + 1), bd, idesc2, 1u);
            else mma(tmem + k * 16, ad0 + (uint64_t)(k * sh + (p.group > 1 ? j % (uint32_t)p.group : 0)), bd, idesc, 1u);
            if (p.commit_every > 0 && (j % (uint32_t)p.commit_every) == (uint32_t)p.commit_every - 1) commit(bar2);
          }
        }
        commit(bar);
        mbar_wait(bar, parity); parity ^= 1;
        long long t1 = clock64();
        if (rep == 0) p.cycles[blockIdx.x] = 0;
        if (rep >= 0) p.cycles[blockIdx.x] += t1 - t0;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));

  // integer-valued data: exact in f16 and in the fp32 accumulator
  std::vector<__half> act((size_t)NPIX * 16), wts(9 * 16 * 16);
  std::vector<float> actf(act.size()), wtsf(wts.size());
  srand(1);
  for (size_t i = 0; i < act.size(); ++i) {
    const int pix = (int)(i / 16);
    const int row = (pix - 8) / PITCH;   // smem row (0 = top zero pad)
    float v = (float)((rand() % 9) - 4);
    if (pix < 8 + PITCH || row >= ROWS + 1) v = 0.f;   // zero rows above/below
    actf[i] = v; act[i] = __float2half(v);
  }
  for (size_t i = 0; i < wts.size(); ++i) { float v = (float)((rand() % 7) - 3); wtsf[i] = v; wts[i] = __float2half(v); }
  __half *d_act, *d_wts; float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_act, act.size() * 2)); CK(cudaMalloc(&d_wts, wts.size() * 2));
  CK(cudaMalloc(&d_out, (size_t)ROWS * 128 * 16 * 4)); CK(cudaMalloc(&d_cyc, 148 * 8));
  CK(cudaMemcpy(d_act, act.data(), act.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_wts, wts.data(), wts.size() * 2, cudaMemcpyHostToDevice));

  // ---- Q1/Q2 correctness
  for (int preload = 0; preload < 2; ++preload) {
    Params p{d_act, d_wts, d_out, d_cyc, 0, 48, 0, 0, 0, 128, 0, preload, 1, 0};
    CK(cudaMemset(d_out, 0, (size_t)ROWS * 128 * 16 * 4));
    probe_kernel<<<1, 128, SMEM_BYTES>>>(p);
    CK(cudaDeviceSynchronize());
    std::vector<float> out((size_t)ROWS * 128 * 16);
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    // reference: out[r][x][co] = sum_{dy,dx,ci} W[dy+1][dx+1][co][ci] * in(r+dy, x+dx) on the linear pixel array
    double maxerr = 0; long bad = 0;
    for (int r = 0; r < ROWS; ++r)
      for (int x = 0; x < 128; ++x)
        for (int co = 0; co < 16; ++co) {
          double s = preload ? 1000.0 + co : 0.0;
          for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
              const int pix = 8 + (r + 1 + dy) * PITCH + x + dx;
              for (int ci = 0; ci < 16; ++ci)
                s += (double)wtsf[(((dy + 1) * 3 + dx + 1) * 16 + co) * 16 + ci] * actf[(size_t)pix * 16 + ci];
            }
          const double e = fabs(s - out[((size_t)r * 128 + x) * 16 + co]);
          if (e > maxerr) maxerr = e;
          if (e > 1e-3) { if (bad < 12) printf("  mismatch r %d x %d co %d: ref %.1f got %.1f\n", r, x, co, s, out[((size_t)r * 128 + x) * 16 + co]); ++bad; }
        }
    printf("Q%d correctness (preload=%d): max err %.6f, mismatches %ld of %d\n", preload ? 2 : 1, preload, maxerr, bad,
           ROWS * 128 * 16);
  }

  // ---- Q3 throughput
  auto run = [&](int grid, int M, int N, int n2, int shift, int group = 1, int commit_every = 0) {
    Params p{d_act, d_wts, d_out, d_cyc, 1, N, 384, 10, shift, M, n2, 0, group, commit_every};
    probe_kernel<<<grid, 128, SMEM_BYTES>>>(p);
    CK(cudaDeviceSynchronize());
    std::vector<long long> cyc(grid);
    CK(cudaMemcpy(cyc.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (auto c : cyc) avg += (double)c;
    avg /= grid;
    const double per = avg / (384.0 * 10.0);
    const double macs = n2 > 0 ? 0.5 * M * (N + n2) * 16 : (double)M * N * 16;
    printf("grid %3d M %3d N %3d n2 %3d shift %3d group %d commit %d: %.1f cyc/MMA -> %.0f MAC/clk/SM\n", grid, M, N, n2, shift, group,
           commit_every, per, macs / per);
  };
  for (int N : {16, 32, 48, 64, 96, 128, 144, 192, 256}) run(148, 128, N, 0, 1);
  for (int N : {16, 32, 48, 64, 96, 128, 256}) run(148, 64, N, 0, 1);
  run(148, 128, 96, 48, 1);
  run(148, 128, 144, 0, 128);
  run(1, 128, 96, 48, 1);
  // the conv stack's pattern: 3 dx-shifted MMAs per D block, rows 128 px apart, a commit per row
  run(148, 128, 48, 0, 128, 1, 0);
  run(148, 128, 48, 0, 128, 3, 0);
  run(148, 128, 48, 0, 128, 3, 3);
  run(148, 128, 48, 0, 128, 1, 3);
  run(148, 128, 48, 0, 128, 1, 1);
  run(148, 128, 48, 0, 1, 3, 3);
  printf("done\n");
  return 0;
}
