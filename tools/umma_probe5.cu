// umma_probe5.cu -- does the D (accumulator) column range of tcgen05.mma wrap modulo the 512 TMEM columns?
// One CTA: zero all 512 columns, one M128 N48 K16 MMA (A = B = 1.0h -> every D element = 16) at column COL0, read back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe5 umma_probe5.cu && ./umma_probe5
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
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
constexpr int PLANE = (128 + 16) * 16;
constexpr int SMEM = 2 * PLANE + 1536 + 1024;

__global__ void __launch_bounds__(128, 1) probe(float* out, int col0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < SMEM / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&s_bar)));
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
  const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t z = 0u;
  for (int c = 0; c < 512; c += 16)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(tq + c), "r"(z) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0 && elect_one_sync()) {
    const uint32_t a0 = smem_u32(smem) + 8 * 16, b0 = smem_u32(smem) + 2 * PLANE;
    const uint32_t idesc = make_idesc(128, 48);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem + (uint32_t)col0),
                 "l"(make_desc(a0, PLANE, 128)), "l"(make_desc(b0, 48 * 16, 128)), "r"(idesc));
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&s_bar)) : "memory");
  }
  __syncwarp();
  mbar_wait(smem_u32(&s_bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int c = 0; c < 512; c += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tq + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int i = 0; i < 16; ++i) out[tid * 512 + c + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

int main() {
  float* d; CK(cudaMalloc(&d, 128 * 512 * 4));
  float* h = (float*)malloc(128 * 512 * 4);
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  const int cols[] = {64, 464, 480, 496};
  for (int col0 : cols) {
    probe<<<1, 128, SMEM>>>(d, col0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("col0 %d: CUDA error %s\n", col0, cudaGetErrorString(e)); return 0; }
    CK(cudaMemcpy(h, d, 128 * 512 * 4, cudaMemcpyDeviceToHost));
    printf("col0 %3d: nonzero 16-column blocks (lane 0 / lane 77):", col0);
    for (int lane : {0, 77}) {
      printf(" [");
      for (int c = 0; c < 512; c += 16) {
        int nz = 0; float val = 0;
        for (int i = 0; i < 16; ++i) if (h[lane * 512 + c + i] != 0.f) { ++nz; val = h[lane * 512 + c + i]; }
        if (nz) printf(" %d:%dx%.0f", c, nz, val);
      }
      printf(" ]");
    }
    printf("\n");
  }
  printf("done\n");
  return 0;
}
