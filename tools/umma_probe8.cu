// umma_probe8.cu -- does a CTA pair (tcgen05.mma.cta_group::2, M = 256) lower the shared-memory operand cost of the
// N = 48 dy-scatter scheme?  Each CTA keeps its own 128 pixel rows of A and HALF of the B columns; the
// hardware sends both B halves to both tensor cores.  Measures cycles per MMA for cta_group 1 / 2 at N = 48 and N = 96
// (N = 96: hi and lo weight parts side by side, the F16X3 arithmetic with two accumulators) with the issue pattern of
// the streaming kernels (3 dx-shifted MMAs per row into one accumulator block), and checks the operand split
// numerically: A[m][k] = (k == 0) * (1 + cta), B[n][k] = (k == 0) * (n + 1)  ->  D[m][n] = (1 + cta)(n + 1) per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe8 umma_probe8.cu && ./umma_probe8
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
}
template <int CG>
__device__ __forceinline__ void commit(uint32_t mbar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar) : "memory");
  else
    asm volatile("{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\ttcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
  return pred;
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar), "r"(cnt)); }
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
constexpr int ROWS = 24;
constexpr int PLANE = (ROWS * 128 + 16) * 16;
constexpr int BBYTES = 3 * 96 * 32;
constexpr int SMEM = 2 * PLANE + BBYTES + 1024;

// out: [cta][128 lanes][N] floats of the check MMA; cycles: per leader CTA
template <int CG, int N>
__global__ void __launch_bounds__(128, 1) probe(long long* cycles, float* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar[2];
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  constexpr int NL = N / CG;   // B columns held by this CTA
  // A: two channel-half planes, pixel-major (16 B per pixel per plane); k = 0 is the first half of plane 0
  for (int i = tid; i < 2 * PLANE / 16; i += 128) {
    const bool plane0 = i < PLANE / 16;
    const __half v = __float2half(plane0 ? (float)(1 + rank) : 0.f);
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    q.x = (uint32_t)__half_as_ushort(v);   // k = 0 only
    reinterpret_cast<uint4*>(smem)[i] = q;
  }
  // B: per dx block [k half 2][n NL][8 halves]; value (n_global + 1) at k = 0
  for (int i = tid; i < BBYTES / 16; i += 128) reinterpret_cast<uint4*>(smem + 2 * PLANE)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int i = tid; i < 3 * NL; i += 128) {
    const int dx = i / NL, n = i % NL;
    __half* b = reinterpret_cast<__half*>(smem + 2 * PLANE + dx * (2 * NL * 16) + n * 16);
    b[0] = __float2half((float)(rank * NL + n + 1));
  }
  if (tid < 2) mbar_init(smem_u32(&s_bar[tid]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&s_tmem)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&s_tmem)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t a0 = smem_u32(smem) + 8 * 16, b0 = smem_u32(smem) + 2 * PLANE, bar = smem_u32(&s_bar[0]);
  const uint32_t idesc = make_idesc(128 * CG, N);
  uint32_t parity = 0;
  // ---- check: one MMA, accumulate off, into column 0
  if (rank == 0 && warp == 0 && elect_one_sync()) {
    mma<CG>(tmem, make_desc(a0, PLANE, 128), make_desc(b0, NL * 16, 128), idesc, 0u);
    commit<CG>(bar);
  }
  mbar_wait(bar, parity); parity ^= 1;
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  {
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      if (blockIdx.x < CG)
        for (int j = 0; j < 16; ++j) out[((size_t)blockIdx.x * 128 + tid) * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // ---- timing: the streaming kernels' issue pattern, 3 dx-shifted MMAs per row into one accumulator block
  if (rank == 0 && warp == 0 && elect_one_sync()) {
    long long total = 0;
    for (int rep = -1; rep < reps; ++rep) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int q = 0; q < ROWS; ++q) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
          mma<CG>(tmem + (q % 8) * (N / 3), make_desc(a0 + (q * 128 + dx) * 16, PLANE, 128), make_desc(b0 + dx * (2 * NL * 16), NL * 16, 128), idesc, 1u);
      }
      commit<CG>(bar);
      mbar_wait(bar, parity); parity ^= 1;
      if (rep >= 0) total += clock64() - t0;
    }
    cycles[blockIdx.x / CG] = total;
  } else if (CG == 2 && rank == 1 && warp == 0 && elect_one_sync()) {
    for (int rep = -1; rep < reps; ++rep) { mbar_wait(bar, parity); parity ^= 1; }   // the peer's operands stay alive
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
  }
}

template <int CG, int N>
void run(long long* d_cyc, float* d_out) {
  const int reps = 20;
  CK(cudaFuncSetAttribute(probe<CG, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  CK(cudaMemset(d_out, 0, 2 * 128 * 96 * sizeof(float)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = SMEM; cfg.stream = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, probe<CG, N>, d_cyc, d_out, reps));
  CK(cudaDeviceSynchronize());
  const int nlead = 148 / CG;
  std::vector<long long> c(nlead);
  CK(cudaMemcpy(c.data(), d_cyc, nlead * 8, cudaMemcpyDeviceToHost));
  double avg = 0; for (auto v : c) avg += (double)v; avg /= nlead;
  std::vector<float> o(2 * 128 * 96);
  CK(cudaMemcpy(o.data(), d_out, o.size() * sizeof(float), cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int r = 0; r < CG; ++r)
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        const float want = (float)((1 + r) * (n + 1));
        if (o[((size_t)r * 128 + m) * N + n] != want) { if (bad < 4) printf("  mismatch cta %d m %d n %d: %g vs %g\n", r, m, n, o[((size_t)r * 128 + m) * N + n], want); ++bad; }
      }
  const double per_mma = avg / reps / ROWS / 3;
  printf("cta_group %d  M %3d N %3d: %.1f cyc/MMA (%.1f per 128-pixel row-dx), operand split %s\n", CG, 128 * CG, N, per_mma, per_mma, bad ? "WRONG" : "ok");
}

int main() {
  long long* d_cyc; CK(cudaMalloc(&d_cyc, 148 * 8));
  float* d_out; CK(cudaMalloc(&d_out, 2 * 128 * 96 * sizeof(float)));
  run<1, 48>(d_cyc, d_out);
  run<1, 96>(d_cyc, d_out);
  run<2, 48>(d_cyc, d_out);
  run<2, 96>(d_cyc, d_out);
  printf("done\n");
  return 0;
}
