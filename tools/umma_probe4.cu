// umma_probe4.cu -- can the dx taps be served from ONE shared-memory read of the activations?
//   tcgen05.cp.128x256b   : copy the 128-pixel x 16-channel fp16 A tile (SWIZZLE_NONE K-major planes) into TMEM
//   tcgen05.mma  [a_tmem] : use it as the A operand
//   tcgen05.shift.down    : move the TMEM rows (pixels) by one lane between the taps
// Questions: layout after cp (P1), MMA with A in TMEM == A in smem (P2), what shift does to the rows (P3),
// cycles per region row for cp + 3 x (mma N=48) + 2 x shift (P4).  Not part of the product library.
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
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void tc_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;\n" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void tc_shift_down(uint32_t taddr) { asm volatile("tcgen05.shift.cta_group::1.down [%0];\n" ::"r"(taddr) : "memory"); }
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

constexpr int ROWS = 24;
constexpr int NPIX = ROWS * 128 + 16;
constexpr int PLANE = NPIX * 16;
constexpr int SMEM = 2 * PLANE + 3 * 1536 + 512 + 1024;
constexpr int A_COL = 448;   // TMEM columns [448, 456) hold the A tile (8 columns = 16 fp16 per lane)

// mode 0: P1 dump of the A tile after cp; mode 1: after cp + 1 shift; mode 2: after cp + 2 shifts;
// mode 3: D = A(tmem) x I16 vs mode 4: D = A(smem) x I16; mode 5: throughput
__global__ void __launch_bounds__(128, 1) probe(const __half* act, uint32_t* dump, float* dout, long long* cycles, int mode, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ __align__(8) uint64_t s_bar2;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* pl0 = smem; uint8_t* pl1 = smem + PLANE;
  __half* sB = reinterpret_cast<__half*>(smem + 2 * PLANE);              // [3][N 48][K 16] core-matrix order (zeros except identity)
  __half* sI = reinterpret_cast<__half*>(smem + 2 * PLANE + 3 * 1536);   // identity [N 16][K 16] core-matrix order
  for (int i = tid; i < NPIX; i += 128) {
    const uint4* s = reinterpret_cast<const uint4*>(act + (size_t)i * 16);
    reinterpret_cast<uint4*>(pl0)[i] = s[0];
    reinterpret_cast<uint4*>(pl1)[i] = s[1];
  }
  for (int i = tid; i < 3 * 768; i += 128) sB[i] = __float2half(((i * 7) % 5 - 2) * 0.25f);
  for (int i = tid; i < 256; i += 128) {   // element (n, k): offset halves = (k/8)*128 + n*8 + (k%8)   (LBO = 256 B, SBO = 128 B)
    const int k8 = i & 7, n = (i >> 3) & 15, kc = i >> 7;
    sI[i] = __float2half((kc * 8 + k8) == n ? 1.f : 0.f);
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
  const uint32_t a0 = smem_u32(pl0) + 8 * 16, bar = smem_u32(&s_bar);
  const uint64_t adesc0 = make_desc(a0, PLANE, 128);
  const uint64_t idn = make_desc(smem_u32(sI), 256, 128);
  const uint64_t bdesc = make_desc(smem_u32(sB), 48 * 16, 128);
  if (mode <= 4) {
    if (warp == 0 && elect_one_sync()) {
      if (mode <= 3) {
        tc_cp_128x256b(tmem + A_COL, adesc0 + 128);            // region row 1 (pixels 128..255)
        if (mode == 1 || mode == 2) tc_shift_down(tmem + A_COL);
        if (mode == 2) tc_shift_down(tmem + A_COL);
        if (mode == 3) mma_ts(tmem + 0, tmem + A_COL, idn, make_idesc(128, 16), 0u);
      } else {
        mma_ss(tmem + 0, adesc0 + 128, idn, make_idesc(128, 16), 0u);
      }
      commit(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    if (mode <= 2) {
      uint32_t v[8];
      tmem_ld8(tmem + lane_base + A_COL, v);
      for (int i = 0; i < 8; ++i) dump[tid * 8 + i] = v[i];
    } else {
      uint32_t v[16];
      tmem_ld16(tmem + lane_base, v);
      for (int i = 0; i < 16; ++i) dout[tid * 16 + i] = __uint_as_float(v[i]);
    }
  } else {
    if (warp == 0 && elect_one_sync()) {
      const uint32_t idesc = make_idesc(128, 48);
      uint32_t parity = 0;
      long long total = 0;
      for (int rep = -1; rep < reps; ++rep) {
        const long long t0 = clock64();
        if (mode >= 11) {
#pragma unroll 1
          for (int q = 0; q < ROWS; ++q) {
            const uint32_t aslot = tmem + A_COL + (uint32_t)(q & 7) * 8;
            const uint32_t d = tmem + (uint32_t)q * 16;
            if (mode == 11 || mode == 13) tc_cp_128x256b(aslot, adesc0 + (uint64_t)(q * 128 + 1));
            if (mode != 13) {
              mma_ts(d, aslot, bdesc, idesc, 1u);
              mma_ts(d, aslot, bdesc + 96, idesc, 1u);
              mma_ts(d, aslot, bdesc + 192, idesc, 1u);
            }
            if (mode == 14) { tc_shift_down(aslot); tc_shift_down(aslot); }
            if ((q % 6) == 5) commit(bar + 0);
          }
        } else if (mode == 5) {
#pragma unroll 1
          for (int q = 0; q < ROWS; ++q) {
            const uint32_t aslot = tmem + A_COL + (uint32_t)(q & 7) * 8;
            const uint32_t d = tmem + (uint32_t)q * 16;
            tc_cp_128x256b(aslot, adesc0 + (uint64_t)(q * 128 + 1));
            mma_ts(d, aslot, bdesc, idesc, 1u);
            tc_shift_down(aslot);
            mma_ts(d, aslot, bdesc + 96, idesc, 1u);
            tc_shift_down(aslot);
            mma_ts(d, aslot, bdesc + 192, idesc, 1u);
            if ((q % 6) == 5) commit(bar + 0);
          }
        } else {
          const int IL = mode - 4;   // rows interleaved: 2, 3, 4, 6
#pragma unroll 1
          for (int q0 = 0; q0 < ROWS; q0 += IL) {
            for (int k = 0; k < IL; ++k) tc_cp_128x256b(tmem + A_COL + (uint32_t)((q0 + k) & 7) * 8, adesc0 + (uint64_t)((q0 + k) * 128 + 1));
            for (int k = 0; k < IL; ++k) mma_ts(tmem + (uint32_t)(q0 + k) * 16, tmem + A_COL + (uint32_t)((q0 + k) & 7) * 8, bdesc, idesc, 1u);
            for (int k = 0; k < IL; ++k) tc_shift_down(tmem + A_COL + (uint32_t)((q0 + k) & 7) * 8);
            for (int k = 0; k < IL; ++k) mma_ts(tmem + (uint32_t)(q0 + k) * 16, tmem + A_COL + (uint32_t)((q0 + k) & 7) * 8, bdesc + 96, idesc, 1u);
            for (int k = 0; k < IL; ++k) tc_shift_down(tmem + A_COL + (uint32_t)((q0 + k) & 7) * 8);
            for (int k = 0; k < IL; ++k) mma_ts(tmem + (uint32_t)(q0 + k) * 16, tmem + A_COL + (uint32_t)((q0 + k) & 7) * 8, bdesc + 192, idesc, 1u);
            if (((q0 + IL) % 6) == 0) commit(bar + 0);
          }
        }
        commit(smem_u32(&s_bar2));
        mbar_wait(smem_u32(&s_bar2), parity); parity ^= 1;
        const long long t1 = clock64();
        if (rep >= 0) total += t1 - t0;
      }
      cycles[blockIdx.x] = total;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

int main() {
  CK(cudaSetDevice(0));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  std::vector<__half> act((size_t)NPIX * 16);
  std::vector<float> actf(act.size());
  for (size_t i = 0; i < act.size(); ++i) {
    const int p = (int)(i / 16), c = (int)(i % 16);
    actf[i] = (float)(((p - 8) % 128) + c * 0.001f * 0 + ((p * 3 + c * 7) % 5) * 256);   // encodes the pixel index and channel
    actf[i] = (float)(((p - 8 + 1280) % 128) * 16 + c);                                  // value = 16*x + channel  (exact in fp16 up to 2047)
    act[i] = __float2half(actf[i]);
  }
  __half* d_act; uint32_t* d_dump; float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_act, act.size() * 2)); CK(cudaMalloc(&d_dump, 128 * 8 * 4)); CK(cudaMalloc(&d_out, 128 * 16 * 4)); CK(cudaMalloc(&d_cyc, 148 * 8));
  CK(cudaMemcpy(d_act, act.data(), act.size() * 2, cudaMemcpyHostToDevice));
  for (int mode = 0; mode <= 2; ++mode) {
    CK(cudaMemset(d_dump, 0xFF, 128 * 8 * 4));
    probe<<<1, 128, SMEM>>>(d_act, d_dump, d_out, d_cyc, mode, 0);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> dump(128 * 8);
    CK(cudaMemcpy(dump.data(), d_dump, dump.size() * 4, cudaMemcpyDeviceToHost));
    printf("mode %d (cp + %d shifts): lane -> 16 halves as (x, channel) pairs, lanes 0,1,2,31,32,33,127:\n", mode, mode);
    for (int lane : {0, 1, 2, 31, 32, 33, 63, 64, 127}) {
      printf("  lane %3d:", lane);
      for (int i = 0; i < 8; ++i) {
        const uint32_t v = dump[lane * 8 + i];
        const float f0 = __half2float(__ushort_as_half((unsigned short)(v & 0xFFFF))), f1 = __half2float(__ushort_as_half((unsigned short)(v >> 16)));
        printf(" (%d,%d)(%d,%d)", (int)f0 / 16, (int)f0 % 16, (int)f1 / 16, (int)f1 % 16);
      }
      printf("\n");
    }
  }
  for (int mode = 3; mode <= 4; ++mode) {
    probe<<<1, 128, SMEM>>>(d_act, d_dump, d_out, d_cyc, mode, 0);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * 16);
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) if (out[m * 16 + n] != (float)(m * 16 + n)) ++bad;
    printf("mode %d (D = A x I, A from %s): mismatches %ld of 2048; D[5][0..3] = %.0f %.0f %.0f %.0f\n", mode, mode == 3 ? "TMEM" : "smem", bad,
           out[80], out[81], out[82], out[83]);
  }
  for (int mode : {5, 11, 12, 13, 14}) {
    probe<<<148, 128, SMEM>>>(d_act, d_dump, d_out, d_cyc, mode, 20);
    CK(cudaDeviceSynchronize());
    std::vector<long long> c(148);
    CK(cudaMemcpy(c.data(), d_cyc, 148 * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto v : c) avg += (double)v; avg /= 148;
    const char* what = mode == 5 ? "cp + 3 mma + 2 shifts" : mode == 11 ? "cp + 3 mma (no shift)" : mode == 12 ? "3 mma A-in-TMEM only" : mode == 13 ? "cp only" : "3 mma + 2 shifts (no cp)";
    printf("P4 %-28s: %.1f cycles per row\n", what, avg / 20 / ROWS);
  }
  printf("done\n");
  return 0;
}
