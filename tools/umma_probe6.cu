// umma_probe6.cu -- TMEM read bandwidth (tcgen05.ld shapes) and its interference with concurrently running tcgen05.mma.
// One CTA per SM: warp 16 issues M128 N48 K16 MMAs (3 per row, as the conv stack), warps 0..NW-1 loop on tcgen05.ld.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe6 umma_probe6.cu && ./umma_probe6
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
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  } while (!done);
}
constexpr int ROWS = 24;
constexpr int PLANE = (ROWS * 128 + 16) * 16;
constexpr int SMEM = 2 * PLANE + 3 * 1536 + 1024;

// MODE 0: tcgen05.ld 32x32b.x16 ; 1: 32x32b.x32 ; 2: 16x256b.x4 (16 regs) ; 3: tcgen05.st 32x32b.x16 ; 4: ld.shared.v4 x2 (LSU control)
template <int MODE>
__device__ __forceinline__ uint32_t tm_op(uint32_t taddr, uint32_t saddr) {
  uint32_t v[32];
  if (MODE == 0) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    return v[0] ^ v[7] ^ v[15];
  } else if (MODE == 1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    return v[0] ^ v[17] ^ v[31];
  } else if (MODE == 2) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    return v[0] ^ v[7] ^ v[15];
  } else if (MODE == 3) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(taddr), "r"(z) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    return 0;
  } else {
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(saddr));
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "r"(saddr + 2048));
    return a ^ d ^ e ^ h;
  }
}

template <int MODE>
__global__ void __launch_bounds__(17 * 32, 1) probe(long long* out, int nw, int do_mma, int reps, int gap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ volatile int s_stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < SMEM / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&s_bar))); s_stop = 0; }
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
  if (warp == 16) {
    if (elect_one_sync()) {
      const uint32_t a0 = smem_u32(smem) + 8 * 16, b0 = smem_u32(smem) + 2 * PLANE, bar = smem_u32(&s_bar);
      const uint32_t idesc = make_idesc(128, 48);
      uint32_t parity = 0;
      const long long t0 = clock64();
      if (do_mma) {
        for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 1
          for (int q = 0; q < ROWS; ++q) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
              mma(tmem + q * 16, make_desc(a0 + (q * 128 + dx) * 16, PLANE, 128), make_desc(b0 + dx * 1536, 48 * 16, 128), idesc);
          }
          commit(bar);
          mbar_wait(bar, parity); parity ^= 1;
        }
      } else {
        while (clock64() - t0 < 400000) {}
      }
      out[blockIdx.x * 4 + 0] = clock64() - t0;
      s_stop = 1;
    }
    __syncwarp();
  } else if (warp < nw) {
    const uint32_t tq = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t sa = smem_u32(smem) + lane * 16;
    uint32_t acc = 0;
    long long n = 0;
    const long long t0 = clock64();
    while (!s_stop) {
#pragma unroll 4
      for (int i = 0; i < 8; ++i) acc ^= tm_op<MODE>(tq + (uint32_t)(((warp >> 2) * 8 + i) & (MODE == 1 ? 15 : 31)) * 16, sa + ((warp * 8 + i) % ROWS) * 2048);
      n += 8;
      if (gap) { const long long tg = clock64(); while (clock64() - tg < gap) {} }
    }
    const long long dt = clock64() - t0;
    if (lane == 0) { atomicAdd((unsigned long long*)&out[blockIdx.x * 4 + 1], (unsigned long long)n); if (warp == 0) out[blockIdx.x * 4 + 2] = dt; }
    if (acc == 0x12345678u) out[blockIdx.x * 4 + 3] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

template <int MODE>
void run(long long* d, int nw, int do_mma, int gap, const char* what) {
  const int reps = 20;
  CK(cudaMemset(d, 0, 148 * 4 * 8));
  CK(cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  probe<MODE><<<148, 17 * 32, SMEM>>>(d, nw, do_mma, reps, gap);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(148 * 4);
  CK(cudaMemcpy(c.data(), d, 148 * 4 * 8, cudaMemcpyDeviceToHost));
  double cyc = 0, ops = 0, dt = 0;
  for (int i = 0; i < 148; ++i) { cyc += c[i * 4]; ops += c[i * 4 + 1]; dt += c[i * 4 + 2]; }
  cyc /= 148; ops /= 148; dt /= 148;
  const double per_mma = do_mma ? cyc / (reps * ROWS * 3) : 0;
  printf("%-34s warps %2d gap %4d mma %d: %.1f cyc/MMA; %.0f ops in %.0f cyc = %.2f cyc/op(all warps), %.1f B/cyc @2KB\n", what, nw, gap, do_mma, per_mma, ops, dt,
         ops > 0 ? dt / ops : 0, ops > 0 ? ops * 2048 / dt : 0);
}

int main() {
  long long* d; CK(cudaMalloc(&d, 148 * 4 * 8));
  run<0>(d, 0, 1, 0, "MMA alone");
  for (int nw : {4, 8, 16}) run<0>(d, nw, 0, 0, "ld 32x32b.x16 alone");
  run<1>(d, 16, 0, 0, "ld 32x32b.x32 alone (4KB)");
  run<2>(d, 4, 0, 0, "ld 16x256b.x4 alone");
  for (int nw : {4, 16}) run<3>(d, nw, 0, 0, "st 32x32b.x16 alone");
  for (int nw : {4, 16}) run<4>(d, nw, 0, 0, "ld.shared.v4 x2 alone (1KB)");
  for (int nw : {4, 8, 16}) run<0>(d, nw, 1, 0, "MMA + ld 32x32b.x16 flat out");
  for (int gap : {200, 400, 800, 1600}) run<0>(d, 16, 1, gap, "MMA + ld 32x32b.x16 paced");
  for (int nw : {4, 16}) run<3>(d, nw, 1, 0, "MMA + st 32x32b.x16 flat out");
  for (int nw : {4, 16}) run<4>(d, nw, 1, 0, "MMA + ld.shared.v4 x2 flat out");
  for (int gap : {200, 400, 800}) run<4>(d, 16, 1, gap, "MMA + ld.shared.v4 x2 paced");
  printf("done\n");
  return 0;
}
