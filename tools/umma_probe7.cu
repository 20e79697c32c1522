// umma_probe7.cu -- can the B operand (the weights, 1.5 KB per MMA, 27 % of the operand bytes of the streaming stack) stay
// in the tensor core's collector between the two rows of a step?  tcgen05.mma.ws (weight-stationary) keeps B in one of
// four collector buffers: row A fills b0..b2 (dx 0..2), row B uses them.  Questions:
//   Q1 numerics: D of the .ws sequence == D of the plain sequence (same operands, M = 128, dy-scatter N = 48 -> 64)
//   Q2 rate    : cycles per row of {plain N=48} vs {.ws N=64 all fill} vs {.ws N=64 fill + lastuse}
//   Q3         : does .ws accept N = 48 at M = 128?  (last: an illegal instruction ends the process)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe7 umma_probe7.cu && ./umma_probe7
// Not part of the product library.
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
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc));
}
#define WS_OP(NAME, QUAL)                                                                                                     \
  __device__ __forceinline__ void NAME(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {                                   \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::" QUAL        \
                 " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc));                                         \
  }
WS_OP(ws_fill0, "b0::fill") WS_OP(ws_fill1, "b1::fill") WS_OP(ws_fill2, "b2::fill")
WS_OP(ws_last0, "b0::lastuse") WS_OP(ws_last1, "b1::lastuse") WS_OP(ws_last2, "b2::lastuse")
WS_OP(ws_disc0, "b0::discard")
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(taddr), "r"(z) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

constexpr int ROWS = 24;                          // input rows per repetition (12 steps of 2 rows)
constexpr int NPIX = ROWS * 128 + 16;
constexpr int PLANE = NPIX * 16;
constexpr int BN = 64;                            // B tiles are stored with N = 64 rows (48 + 16 zero rows)
constexpr int BTILE = BN * 16 * 2;                // bytes of one B tile [N 64][K 16] fp16
constexpr int SMEM = 2 * PLANE + 3 * BTILE + 1024;

// mode 0: plain N=48; 1: .ws N=64, every MMA fills (no reuse); 2: .ws N=64, row A fills b0..b2, row B lastuse b0..b2;
// 3: as 2 with N=48; 4: plain N=64.
template <int MODE>
__global__ void __launch_bounds__(128, 1) probe(const __half* act, const __half* wts, float* dout, long long* cycles, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* pl0 = smem; uint8_t* pl1 = smem + PLANE;
  for (int i = tid; i < NPIX; i += 128) {
    const uint4* s = reinterpret_cast<const uint4*>(act + (size_t)i * 16);
    reinterpret_cast<uint4*>(pl0)[i] = s[0];
    reinterpret_cast<uint4*>(pl1)[i] = s[1];
  }
  // B tile (dx): element (n, k) at halves (k/8)*(BN*8) + n*8 + (k%8)   (LBO = BN*16 B, SBO = 128 B); rows 48..63 zero
  __half* sB = reinterpret_cast<__half*>(smem + 2 * PLANE);
  for (int i = tid; i < 3 * BN * 16; i += 128) {
    const int dx = i / (BN * 16), r = i % (BN * 16), n = r / 16, k = r % 16;
    sB[dx * BN * 16 + (k / 8) * (BN * 8) + n * 8 + (k % 8)] = n < 48 ? wts[(dx * 48 + n) * 16 + k] : __float2half(0.f);
  }
  if (tid == 0) mbar_init(smem_u32(&s_bar), 1);
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
  for (int b = 0; b < 32; ++b) tmem_zero16(tmem + ((uint32_t)(warp * 32) << 16) + b * 16);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t a0 = smem_u32(smem) + 8 * 16, b0 = smem_u32(smem) + 2 * PLANE, bar = smem_u32(&s_bar);
  constexpr int N = (MODE == 0 || MODE == 3) ? 48 : 64;
  if (warp == 0 && elect_one_sync()) {
    const uint32_t idesc = make_idesc(128, N);
    uint32_t parity = 0;
    long long total = 0;
    for (int rep = -1; rep < reps; ++rep) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int q = 0; q < ROWS; q += 2) {
        // input rows q, q+1 scatter into accumulator blocks q.. (dy): D column = q * 16 (wraps inside 512 columns: q < 24)
        const uint32_t dA = tmem + q * 16, dB = tmem + (q + 1) * 16;
        const uint64_t aA = make_desc(a0 + (q * 128) * 16, PLANE, 128), aB = make_desc(a0 + ((q + 1) * 128) * 16, PLANE, 128);
        const uint64_t bd0 = make_desc(b0, BN * 16, 128), bd1 = make_desc(b0 + BTILE, BN * 16, 128), bd2 = make_desc(b0 + 2 * BTILE, BN * 16, 128);
        if (MODE == 0 || MODE == 4) {
          mma(dA, aA, bd0, idesc); mma(dA, aA + 1, bd1, idesc); mma(dA, aA + 2, bd2, idesc);
          mma(dB, aB, bd0, idesc); mma(dB, aB + 1, bd1, idesc); mma(dB, aB + 2, bd2, idesc);
        } else if (MODE == 1) {
          ws_disc0(dA, aA, bd0, idesc); ws_disc0(dA, aA + 1, bd1, idesc); ws_disc0(dA, aA + 2, bd2, idesc);
          ws_disc0(dB, aB, bd0, idesc); ws_disc0(dB, aB + 1, bd1, idesc); ws_disc0(dB, aB + 2, bd2, idesc);
        } else {
          ws_fill0(dA, aA, bd0, idesc); ws_fill1(dA, aA + 1, bd1, idesc); ws_fill2(dA, aA + 2, bd2, idesc);
          ws_last0(dB, aB, bd0, idesc); ws_last1(dB, aB + 1, bd1, idesc); ws_last2(dB, aB + 2, bd2, idesc);
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
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (blockIdx.x == 0 && dout) {   // dump the 26 accumulator blocks: dout[block][lane 128][16]
    for (int b = 0; b < ROWS + 2; ++b) {
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + b * 16, v);
      for (int i = 0; i < 16; ++i) dout[((size_t)b * 128 + tid) * 16 + i] = __uint_as_float(v[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

template <int MODE>
std::vector<float> run(const __half* d_act, const __half* d_w, float* d_out, long long* d_cyc, const char* what) {
  const int reps = 20;
  CK(cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  // numerics: one repetition (rep -1 only) so that every block holds exactly one set of products
  probe<MODE><<<1, 128, SMEM>>>(d_act, d_w, d_out, d_cyc, 0);
  CK(cudaDeviceSynchronize());
  std::vector<float> out((size_t)(ROWS + 2) * 128 * 16);
  CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  probe<MODE><<<148, 128, SMEM>>>(d_act, d_w, nullptr, d_cyc, reps);
  CK(cudaDeviceSynchronize());
  std::vector<long long> c(148);
  CK(cudaMemcpy(c.data(), d_cyc, 148 * 8, cudaMemcpyDeviceToHost));
  double avg = 0; for (auto v : c) avg += (double)v; avg /= 148;
  printf("mode %d: %.1f cycles per input row (3 MMAs)  -- %s\n", MODE, avg / reps / ROWS, what);
  fflush(stdout);
  return out;
}

int main() {
  std::vector<__half> act((size_t)NPIX * 16), w(3 * 48 * 16);
  srand(1);
  for (auto& v : act) v = __float2half((float)(rand() % 17 - 8) / 8.f);
  for (auto& v : w) v = __float2half((float)(rand() % 9 - 4) / 16.f);
  __half *d_act, *d_w; float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_act, act.size() * 2)); CK(cudaMalloc(&d_w, w.size() * 2));
  CK(cudaMalloc(&d_out, (size_t)(ROWS + 2) * 128 * 16 * 4)); CK(cudaMalloc(&d_cyc, 148 * 8));
  CK(cudaMemcpy(d_act, act.data(), act.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_w, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
  auto ref = run<0>(d_act, d_w, d_out, d_cyc, "plain tcgen05.mma, N = 48 (the streaming stack today)");
  auto cmp = [&](const std::vector<float>& o, const char* name) {
    double md = 0; size_t bad = 0;
    for (size_t i = 0; i < (size_t)(ROWS + 2) * 128 * 16; ++i) { const double d = fabs((double)o[i] - ref[i]); if (d > md) md = d; if (d > 1e-3) ++bad; }
    printf("   %s vs plain N=48: max |diff| %.3g, %zu of %zu values differ\n", name, md, bad, (size_t)(ROWS + 2) * 128 * 16);
    fflush(stdout);
  };
  cmp(run<4>(d_act, d_w, d_out, d_cyc, "plain tcgen05.mma, N = 64 (zero-padded B)"), "plain N=64");
  cmp(run<1>(d_act, d_w, d_out, d_cyc, ".ws N = 64, collector b0::discard on every MMA (no reuse)"), ".ws discard");
  cmp(run<2>(d_act, d_w, d_out, d_cyc, ".ws N = 64, row A fills b0..b2, row B lastuse b0..b2"), ".ws fill/lastuse");
  cmp(run<3>(d_act, d_w, d_out, d_cyc, ".ws N = 48, row A fills, row B lastuse"), ".ws N=48");
  printf("done\n");
  return 0;
}
