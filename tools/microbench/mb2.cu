// Instruction-throughput microbenchmarks (warp-instructions per clock per SM) that size the inner loops.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

template <int MODE>
__global__ void __launch_bounds__(512) k(int iters, float* out, long long* cyc, const float* in) {
  float a[16]; unsigned u[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = in[i] + threadIdx.x; u[i] = __float_as_uint(in[16 + i]) + threadIdx.x; }
  float w = in[40]; unsigned wb = __float_as_uint(in[41]);
  unsigned long long wp; asm("mov.b64 %0, {%1, %1};" : "=l"(wp) : "f"(w));
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {            // FFMA 3-reg
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(__uint_as_float(u[i])), "f"(w));
      } else if (MODE == 1) {     // FFMA2 (pairs a[i], a[i+1]) -> 8 per inner loop
        if (i & 1) continue;
        unsigned long long acc, v;
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a[i]), "f"(a[i + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(u[i]), "r"(u[i + 1]));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(v), "l"(wp));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(acc));
      } else if (MODE == 2) {     // FHFMA.BF16 (mixed precision: bf16 x bf16 + f32)
        unsigned short lo, hi, wl, wh;
        asm("mov.b32 {%0,%1}, %2;" : "=h"(lo), "=h"(hi) : "r"(u[i]));
        asm("mov.b32 {%0,%1}, %2;" : "=h"(wl), "=h"(wh) : "r"(wb));
        if (i & 1) asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(a[i]) : "h"(hi), "h"(wl));
        else       asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(a[i]) : "h"(lo), "h"(wl));
      } else if (MODE == 3) {     // LOP3 (alu)
        asm volatile("and.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(wb + i));
      } else if (MODE == 4) {     // IMAD.U32 shift (fma pipe)
        asm volatile("mul.lo.u32 %0, %0, 65537;" : "+r"(u[i]));
      } else if (MODE == 5) {     // F2I.FLOOR
        int r; asm volatile("cvt.rmi.s32.f32 %0, %1;" : "=r"(r) : "f"(a[i])); a[i] = __int_as_float(r);
      } else if (MODE == 6) {     // FADD.RM
        asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(w));
      } else if (MODE == 7) {     // mix: FHFMA + LOP3 interleaved (dual pipe?)
        unsigned short lo, hi, wl, wh;
        asm("mov.b32 {%0,%1}, %2;" : "=h"(lo), "=h"(hi) : "r"(u[i]));
        asm("mov.b32 {%0,%1}, %2;" : "=h"(wl), "=h"(wh) : "r"(wb));
        if (i & 1) asm volatile("and.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(wb + i));
        else       asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(a[i]) : "h"(lo), "h"(wl));
      } else if (MODE == 8) {     // mix: FFMA + LOP3 interleaved
        if (i & 1) asm volatile("and.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(wb + i));
        else       asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(__uint_as_float(u[i])), "f"(w));
      } else if (MODE == 9) {     // PRMT
        asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(u[i]) : "r"(wb));
      } else if (MODE == 10) {    // SHFL
        a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);
      }
    }
  }
  long long t1 = clock64();
  float s = 0; unsigned su = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { s += a[i]; su ^= u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(su);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// LDS.128 peak: 8-lane groups, address pattern precomputed (no div/mod in loop), conflict-free pairs
__global__ void __launch_bounds__(512) lds_peak(int iters, int npix, float* out, long long* cyc, int mode) {
  extern __shared__ __align__(128) uint8_t smem[];
  for (int i = threadIdx.x; i < npix * 16; i += blockDim.x) ((uint32_t*)smem)[i] = i;
  __syncthreads();
  uint32_t s = (blockIdx.x * 977 + (threadIdx.x >> (mode == 0 ? 3 : 2))) * 2654435761u + 12345u;
  const int sub = mode == 0 ? (threadIdx.x & 7) : (threadIdx.x & 3);
  const uint32_t mask = 511;   // npix >= 513
  uint4 acc = make_uint4(0, 0, 0, 0);
  long long t0 = clock64();
#pragma unroll 8
  for (int it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    uint32_t p = (s >> 10) & mask;
    const uint4 v = *(const uint4*)(smem + p * 64 + sub * 16);
    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc.x ^ acc.y ^ acc.z ^ acc.w);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  float* out; long long* cyc; float* in;
  CK(cudaMalloc(&out, nsm * 512 * 4)); CK(cudaMalloc(&cyc, nsm * 8)); CK(cudaMalloc(&in, 64 * 4));
  float hin[64]; for (int i = 0; i < 64; ++i) hin[i] = 1.0f + i * 0.001f;
  CK(cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice));
  long long hc[1024];
  const char* names[11] = {"FFMA_3reg", "FFMA2", "FHFMA_BF16", "LOP3", "IMAD_U32", "F2I_FLOOR", "FADD_RM", "FHFMA+LOP3", "FFMA+LOP3", "PRMT", "SHFL"};
  int iters = 2048;
  for (int mode = 0; mode < 11; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (mode) {
        case 0: k<0><<<nsm, 512>>>(iters, out, cyc, in); break; case 1: k<1><<<nsm, 512>>>(iters, out, cyc, in); break;
        case 2: k<2><<<nsm, 512>>>(iters, out, cyc, in); break; case 3: k<3><<<nsm, 512>>>(iters, out, cyc, in); break;
        case 4: k<4><<<nsm, 512>>>(iters, out, cyc, in); break; case 5: k<5><<<nsm, 512>>>(iters, out, cyc, in); break;
        case 6: k<6><<<nsm, 512>>>(iters, out, cyc, in); break; case 7: k<7><<<nsm, 512>>>(iters, out, cyc, in); break;
        case 8: k<8><<<nsm, 512>>>(iters, out, cyc, in); break; case 9: k<9><<<nsm, 512>>>(iters, out, cyc, in); break;
        case 10: k<10><<<nsm, 512>>>(iters, out, cyc, in); break;
      }
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(hc, cyc, 8 * nsm, cudaMemcpyDeviceToHost));
    double n = (mode == 1 ? 8.0 : 16.0) * iters * 16;   // warp-instr per CTA (16 warps)
    printf("{\"test\":\"issue_%s\",\"warp_instr_per_clk_per_sm\":%.3f}\n", names[mode], n / hc[0]);
  }
  int npix = 1024, smem = npix * 64;
  CK(cudaFuncSetAttribute(lds_peak, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) { lds_peak<<<nsm, 512, smem>>>(4096, npix, out, cyc, mode); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(hc, cyc, 8 * nsm, cudaMemcpyDeviceToHost));
    printf("{\"test\":\"lds128_peak_%s\",\"B_per_clk_per_sm\":%.1f}\n", mode == 0 ? "8lane_128B" : "4lane_64B_random", 512.0 * 16 * 4096 / hc[0]);
  }
  return 0;
}
