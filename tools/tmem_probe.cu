// tmem_probe.cu -- tcgen05.ld / tcgen05.st throughput per SM (is the epilogue of the
// tcgen05 conv stack TMEM-read-bound?).  Not part of the product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ uint32_t ld_x(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld_x<16>(uint32_t taddr) {
  uint32_t v[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  uint32_t a = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) a ^= v[i];
  return a;
}
template <>
__device__ __forceinline__ uint32_t ld_x<32>(uint32_t taddr) {
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  uint32_t a = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) a ^= v[i];
  return a;
}
__device__ __forceinline__ void st16(uint32_t taddr, uint32_t x) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(taddr), "r"(x));
}

// mode 0: x16 loads, 1: x32 loads, 2: x16 stores, 3: x16 load + x16 store (the epilogue's drain-and-zero)
__global__ void probe(int mode, int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
  // initialise every column this warp touches
  for (int c = 0; c < 512; c += 16) st16(tmem + c, 0u);
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  __syncthreads();
  uint32_t acc = 0;
  const long long t0 = clock64();
  const uint32_t cbase = (uint32_t)((warp >> 2) * 64) & 511u;
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (cbase + (uint32_t)(i * 32)) & 480u;
    if (mode == 0) acc ^= ld_x<16>(tmem + col);
    else if (mode == 1) acc ^= ld_x<32>(tmem + col);
    else if (mode == 2) st16(tmem + col, acc);
    else { acc ^= ld_x<16>(tmem + col); st16(tmem + col, 0u); }
  }
  if (mode >= 2) asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(s_tmem));
}

int main() {
  CK(cudaSetDevice(0));
  long long* d_cyc; uint32_t* d_sink;
  CK(cudaMalloc(&d_cyc, 148 * 8)); CK(cudaMalloc(&d_sink, 148 * 1024 * 4));
  const int iters = 4096;
  const char* names[4] = {"ld.x16", "ld.x32", "st.x16", "ld.x16+st.x16"};
  for (int mode = 0; mode < 4; ++mode)
    for (int threads : {128, 256, 512}) {
      probe<<<148, threads>>>(mode, iters, d_cyc, d_sink);
      CK(cudaDeviceSynchronize());
      std::vector<long long> c(148);
      CK(cudaMemcpy(c.data(), d_cyc, 148 * 8, cudaMemcpyDeviceToHost));
      double avg = 0; for (auto v : c) avg += (double)v; avg /= 148;
      const double bytes = (double)iters * (threads / 32) * 32 * ((mode == 1) ? 32 : 16) * 4 * ((mode == 3) ? 2 : 1);
      printf("%-14s warps %2d: %.1f cyc/iter/warp, %.0f B/clk/SM\n", names[mode], threads / 32, avg / iters, bytes / avg);
    }
  printf("done\n");
  return 0;
}
