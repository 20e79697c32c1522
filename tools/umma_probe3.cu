// umma_probe3.cu -- issue-pattern microbenchmark for the tcgen05 conv stack.
// Pattern of fused_umma.cu: per region row q, GROUP dx-shifted MMAs (M128 N48 K16, A/B from shared memory) that
// accumulate into the same TMEM block, rows 128 pixels apart, a tcgen05.commit every COMMIT MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe3 umma_probe3.cu && ./umma_probe3
#include <cuda_runtime.h>
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
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc));
}
__device__ __forceinline__ void commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar) : "memory");
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
constexpr int ROWS = 24;
constexpr int PLANE = (ROWS * 128 + 16) * 16;
constexpr int SMEM = 2 * PLANE + 3 * 1536 + 1024;

// ORDER 0: row-major (q, dx) as in fused_umma.cu.  ORDER 1: dx-major over groups of 3 rows q, q+3, q+6 (disjoint D).
template <int N, int GROUP, int COMMIT, int ORDER, int WAITS = 0>
__global__ void __launch_bounds__(128, 1) probe(long long* cycles, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar[33];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < SMEM / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (tid < 33) mbar_init(smem_u32(&s_bar[tid]), 1);
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
  const uint32_t a0 = smem_u32(smem) + 8 * 16, b0 = smem_u32(smem) + 2 * PLANE, bar = smem_u32(&s_bar[32]);
  if (warp == 0 && elect_one_sync()) {
    const uint32_t idesc = make_idesc(128, N);
    uint32_t parity = 0;
    long long total = 0;
    for (int rep = -1; rep < reps; ++rep) {
      const long long t0 = clock64();
      int cnt = 0;
      if (ORDER == 0) {
#pragma unroll 1
        for (int q = 0; q < ROWS; ++q) {
          if (WAITS > 0 && (q % 3) == 0) {
#pragma unroll
            if (WAITS < 10) {
#pragma unroll
              for (int w = 0; w < (WAITS % 10); ++w) mbar_wait(smem_u32(&s_bar[31 - w]), 1);   // never-armed barrier: parity 1 reads "complete"
            }
            if (WAITS == 10) { volatile uint32_t* f = reinterpret_cast<volatile uint32_t*>(smem + SMEM - 64); while (*f == 0xdeadbeefu) {} }  // flag poll
            if (WAITS != 1 && WAITS != 10) asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          }
#pragma unroll
          for (int dx = 0; dx < GROUP; ++dx) {
            mma(tmem + q * 16, make_desc(a0 + (q * 128 + dx) * 16, PLANE, 128), make_desc(b0 + (dx % 3) * 1536, N * 16, 128), idesc);
            if (COMMIT > 0 && (++cnt % COMMIT) == 0) commit(smem_u32(&s_bar[q]));
          }
        }
      } else {
#pragma unroll 1
        for (int q0 = 0; q0 < ROWS; q0 += 9)
#pragma unroll 1
          for (int ph = 0; ph < 3; ++ph) {
#pragma unroll
            for (int dx = 0; dx < GROUP; ++dx)
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                const int q = q0 + ph + 3 * k;
                if (q < ROWS) mma(tmem + q * 16, make_desc(a0 + (q * 128 + dx) * 16, PLANE, 128), make_desc(b0 + (dx % 3) * 1536, N * 16, 128), idesc);
              }
            if (COMMIT > 0) commit(smem_u32(&s_bar[q0 + ph]));
          }
      }
      commit(bar);
      mbar_wait(bar, parity); parity ^= 1;
      if (rep >= 0) total += clock64() - t0;
    }
    cycles[blockIdx.x] = total;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

template <int N, int GROUP, int COMMIT, int ORDER, int WAITS = 0>
void run(long long* d_cyc, const char* what) {
  const int reps = 20;
  CK(cudaFuncSetAttribute(probe<N, GROUP, COMMIT, ORDER, WAITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  probe<N, GROUP, COMMIT, ORDER, WAITS><<<148, 128, SMEM>>>(d_cyc, reps);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(148);
  CK(cudaMemcpy(c.data(), d_cyc, 148 * 8, cudaMemcpyDeviceToHost));
  double avg = 0; for (auto v : c) avg += (double)v; avg /= 148;
  const double per_row = avg / reps / ROWS;
  printf("N %3d group %d commit %d order %d: %.1f cyc/row, %.1f cyc/MMA  (%s)\n", N, GROUP, COMMIT, ORDER, per_row, per_row / GROUP, what);
}

int main() {
  long long* d_cyc; CK(cudaMalloc(&d_cyc, 148 * 8));
  run<48, 1, 0, 0>(d_cyc, "1 MMA per row, no commits");
  run<48, 3, 0, 0>(d_cyc, "3 same-D MMAs per row, no commits");
  run<48, 3, 3, 0>(d_cyc, "3 same-D MMAs per row, commit per row (fused_umma.cu today)");
  run<48, 3, 1, 0>(d_cyc, "commit per MMA");
  run<48, 3, 0, 1>(d_cyc, "dx-major over 3 disjoint rows, no commits");
  run<48, 3, 3, 1>(d_cyc, "dx-major over 3 disjoint rows, commit per 9 MMAs");
  run<48, 3, 9, 0>(d_cyc, "row-major, commit per 3 rows");
  run<48, 3, 9, 0, 1>(d_cyc, "1 completed try_wait per 3 rows, NO tcgen05 fence");
  run<48, 3, 9, 0, 2>(d_cyc, "2 completed try_waits per 3 rows + fence::after_thread_sync");
  run<48, 3, 9, 0, 20>(d_cyc, "only fence::after_thread_sync per 3 rows");
  run<48, 3, 9, 0, 10>(d_cyc, "volatile smem flag poll per 3 rows, no fence");
  run<48, 3, 12, 0>(d_cyc, "row-major, commit per 4 rows");
  run<48, 3, 6, 0>(d_cyc, "row-major, commit per 2 rows");
  run<16, 9, 9, 0>(d_cyc, "N=16: 9 MMAs per row (plain implicit GEMM)");
  run<144, 1, 1, 0>(d_cyc, "N=144: 1 MMA per row");
  printf("done\n");
  return 0;
}
