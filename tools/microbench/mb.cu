// Microbenchmarks that size the design choices of the MSDeformAttn kernels on B200 (sm_100a).
// Each test prints one JSON line. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb mb.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// ---- MB1: LDS.128 gather, 8-lane groups read 128 B (two adjacent 64 B pixels) at random pixel pairs
__global__ void __launch_bounds__(512) mb_lds_gather(int iters, int npix, uint32_t* sink, long long* cyc) {
  extern __shared__ __align__(128) uint8_t smem[];
  for (int i = threadIdx.x; i < npix * 16; i += blockDim.x) ((uint32_t*)smem)[i] = i;
  __syncthreads();
  uint32_t s = (blockIdx.x * 977 + (threadIdx.x >> 3)) * 2654435761u + 12345u;
  uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  int sub = threadIdx.x & 7;
  long long t0 = clock64();
  #pragma unroll 4
  for (int it = 0; it < iters; ++it) {
    uint32_t p = lcg(s) % (uint32_t)(npix - 1);
    const uint4 v = *(const uint4*)(smem + p * 64 + sub * 16);
    acc0 ^= v.x; acc1 += v.y; acc2 ^= v.z; acc3 += v.w;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if ((acc0 ^ acc1 ^ acc2 ^ acc3) == 0x12345) sink[0] = acc0;
}

// ---- MB1b: same but 4-lane groups reading 64 B at random pixels (expect ~1.5x conflicts)
__global__ void __launch_bounds__(512) mb_lds_gather4(int iters, int npix, uint32_t* sink, long long* cyc) {
  extern __shared__ __align__(128) uint8_t smem[];
  for (int i = threadIdx.x; i < npix * 16; i += blockDim.x) ((uint32_t*)smem)[i] = i;
  __syncthreads();
  uint32_t s = (blockIdx.x * 977 + (threadIdx.x >> 2)) * 2654435761u + 12345u;
  uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  int sub = threadIdx.x & 3;
  long long t0 = clock64();
  #pragma unroll 4
  for (int it = 0; it < iters; ++it) {
    uint32_t p = lcg(s) % (uint32_t)(npix);
    const uint4 v = *(const uint4*)(smem + p * 64 + sub * 16);
    acc0 ^= v.x; acc1 += v.y; acc2 ^= v.z; acc3 += v.w;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if ((acc0 ^ acc1 ^ acc2 ^ acc3) == 0x12345) sink[0] = acc0;
}

// ---- MB2: shared int atomics at random addresses
__global__ void __launch_bounds__(512) mb_atoms(int iters, int nwords, uint32_t* sink, long long* cyc) {
  extern __shared__ __align__(128) uint8_t smem[];
  int* sm = (int*)smem;
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  uint32_t s = (blockIdx.x * 977 + threadIdx.x) * 2654435761u + 12345u;
  long long t0 = clock64();
  #pragma unroll 4
  for (int it = 0; it < iters; ++it) {
    uint32_t p = lcg(s) % (uint32_t)nwords;
    atomicAdd(&sm[p], 1);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (sm[threadIdx.x] == 0x7fffffff) sink[0] = 1;
}

// ---- MB2b: shared-memory RMW of 128 B rows without atomics (LDS.128 + FADD + STS.128), 8 lanes per row
__global__ void __launch_bounds__(512) mb_smem_rmw(int iters, int nrows, uint32_t* sink, long long* cyc) {
  extern __shared__ __align__(128) uint8_t smem[];
  for (int i = threadIdx.x; i < nrows * 32; i += blockDim.x) ((float*)smem)[i] = 0.f;
  __syncthreads();
  uint32_t s = (blockIdx.x * 977 + (threadIdx.x >> 3)) * 2654435761u + 12345u;
  int sub = threadIdx.x & 7;
  long long t0 = clock64();
  #pragma unroll 4
  for (int it = 0; it < iters; ++it) {
    uint32_t p = lcg(s) % (uint32_t)nrows;
    float4* a = (float4*)(smem + p * 128 + sub * 16);
    float4 v = *a; v.x += 1.f; v.y += 2.f; v.z += 3.f; v.w += 4.f; *a = v;
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (((float*)smem)[threadIdx.x] == -1.f) sink[0] = 1;
}

// ---- MB3/4: global vector reductions, 8 lanes x 16 B = 128 B rows (f32x4) or 4 lanes x 16 B = 64 B rows (bf16x8)
template <int MODE>  // 0: f32x4 (128 B rows), 1: bf16x8 (64 B rows), 2: scalar f32 (32 lanes x 4 B), 3: bf16x2 (16 lanes x 4 B)
__global__ void __launch_bounds__(256) mb_red(int iters, uint32_t nrows, uint8_t* buf, int local_span) {
  constexpr int LPR = MODE == 0 ? 8 : MODE == 1 ? 4 : MODE == 2 ? 32 : 16;   // lanes per row
  constexpr int ROWB = (MODE == 0 || MODE == 2) ? 128 : 64;
  uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  int sub = threadIdx.x % LPR;
  uint32_t s = gid * 2654435761u + 12345u;
  // local_span > 0: rows are drawn from a CTA-local span (spatial locality like an encoder tile)
  uint32_t base = local_span > 0 ? (uint32_t)(((uint64_t)blockIdx.x * 7919u * (uint32_t)local_span) % (nrows - local_span)) : 0u;
  uint32_t span = local_span > 0 ? (uint32_t)local_span : nrows;
  for (int it = 0; it < iters; ++it) {
    uint32_t p = base + lcg(s) % span;
    uint8_t* a = buf + (size_t)p * ROWB;
    if (MODE == 0) {
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(a + sub * 16), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
    } else if (MODE == 1) {
      asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%2,%3,%4};" :: "l"(a + sub * 16), "r"(0x3f803f80u), "r"(0x3f803f80u), "r"(0x3f803f80u), "r"(0x3f803f80u) : "memory");
    } else if (MODE == 2) {
      asm volatile("red.global.add.f32 [%0], %1;" :: "l"(a + sub * 4), "f"(1.f) : "memory");
    } else {
      asm volatile("red.global.add.noftz.bf16x2 [%0], %1;" :: "l"(a + sub * 4), "r"(0x3f803f80u) : "memory");
    }
  }
}

// ---- MB5: TMA bulk reduce smem -> global (.add.f32), chunk bytes per op
__global__ void __launch_bounds__(128) mb_bulk_red(int iters, int chunk, uint32_t nchunks, uint8_t* buf) {
  extern __shared__ __align__(128) uint8_t smem[];
  for (int i = threadIdx.x; i < chunk / 4; i += blockDim.x) ((float*)smem)[i] = 1.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = blockIdx.x * 2654435761u + 12345u;
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    for (int it = 0; it < iters; ++it) {
      uint32_t p = lcg(s) % nchunks;
      uint8_t* g = buf + (size_t)p * chunk;
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" :: "l"(g), "r"(sa), "r"(chunk) : "memory");
      if ((it & 7) == 7) { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// ---- MB6: LDG.128 gather from global: 4-lane groups read 64 B segments at random rows
__global__ void __launch_bounds__(256) mb_ldg_gather(int iters, uint32_t nrows, const uint8_t* __restrict__ buf, uint32_t* sink, int local_span) {
  uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  int sub = threadIdx.x & 3;
  uint32_t s = gid * 2654435761u + 12345u;
  uint32_t base = local_span > 0 ? (uint32_t)(((uint64_t)blockIdx.x * 7919u * (uint32_t)local_span) % (nrows - local_span)) : 0u;
  uint32_t span = local_span > 0 ? (uint32_t)local_span : nrows;
  uint32_t acc = 0;
  #pragma unroll 4
  for (int it = 0; it < iters; ++it) {
    uint32_t p = base + lcg(s) % span;
    const uint4 v = __ldg((const uint4*)(buf + (size_t)p * 64 + sub * 16));
    acc ^= v.x + v.y + v.z + v.w;
  }
  if (acc == 0x12345) sink[0] = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("{\"device\":\"%s\",\"sms\":%d,\"l2_bytes\":%d,\"smem_optin\":%zu,\"clock_khz\":%d}\n", prop.name, nsm, prop.l2CacheSize, prop.sharedMemPerBlockOptin, clk_khz);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  uint32_t* sink; CK(cudaMalloc(&sink, 64));
  long long* cyc; CK(cudaMalloc(&cyc, sizeof(long long) * 4096));
  long long hc[4096];

  {  // MB1
    int npix = 800, iters = 4096, smem = npix * 64;
    CK(cudaFuncSetAttribute(mb_lds_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(mb_lds_gather4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int which = 0; which < 2; ++which) for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      if (which == 0) mb_lds_gather<<<nsm, 512, smem>>>(iters, npix, sink, cyc); else mb_lds_gather4<<<nsm, 512, smem>>>(iters, npix, sink, cyc);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hc, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
      double bytes = 512.0 * 16 * iters;  // per CTA
      if (rep) printf("{\"test\":\"lds128_gather_%dlane\",\"B_per_clk_per_sm\":%.1f,\"ms\":%.3f,\"TBps_chip\":%.2f}\n", which ? 4 : 8, bytes / hc[0], time_ms(e0, e1), bytes * nsm / time_ms(e0, e1) / 1e9);
    }
  }
  {  // MB2
    int nwords = 16384, iters = 2048, smem = nwords * 4;
    CK(cudaFuncSetAttribute(mb_atoms, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int rep = 0; rep < 2; ++rep) {
      mb_atoms<<<nsm, 512, smem>>>(iters, nwords, sink, cyc); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hc, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
      if (rep) printf("{\"test\":\"atoms_add_s32_random\",\"lane_ops_per_clk_per_sm\":%.2f}\n", 512.0 * iters / hc[0]);
    }
    int nrows = 1024; smem = nrows * 128;
    CK(cudaFuncSetAttribute(mb_smem_rmw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int rep = 0; rep < 2; ++rep) {
      mb_smem_rmw<<<nsm, 512, smem>>>(iters, nrows, sink, cyc); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hc, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
      if (rep) printf("{\"test\":\"smem_rmw_128B_rows\",\"rows_per_clk_per_sm\":%.3f,\"B_rmw_per_clk_per_sm\":%.1f}\n", 64.0 * iters / hc[0], 64.0 * iters * 128 / hc[0]);
    }
  }
  {  // MB3/4 global reductions
    size_t sizes[3] = {24u << 20, 96u << 20, 768u << 20};
    uint8_t* buf; CK(cudaMalloc(&buf, sizes[2]));
    CK(cudaMemset(buf, 0, sizes[2]));
    int iters = 256;
    int grid = nsm * 16;
    for (int mode = 0; mode < 4; ++mode) for (int si = 0; si < 3; ++si) for (int loc = 0; loc < 3; ++loc) {
      int rowb = (mode == 0 || mode == 2) ? 128 : 64;
      int lpr = mode == 0 ? 8 : mode == 1 ? 4 : mode == 2 ? 32 : 16;
      uint32_t nrows = (uint32_t)(sizes[si] / rowb);
      int local_span = loc == 0 ? 0 : loc == 1 ? 2048 : 256;
      if (loc == 2 && si != 0) continue;
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        if (mode == 0) mb_red<0><<<grid, 256>>>(iters, nrows, buf, local_span);
        else if (mode == 1) mb_red<1><<<grid, 256>>>(iters, nrows, buf, local_span);
        else if (mode == 2) mb_red<2><<<grid, 256>>>(iters, nrows, buf, local_span);
        else mb_red<3><<<grid, 256>>>(iters, nrows, buf, local_span);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms = time_ms(e0, e1); if (ms < best) best = ms;
      }
      double rows = (double)grid * 256 / lpr * iters;
      printf("{\"test\":\"redg_%s\",\"buf_MB\":%zu,\"local_span_rows\":%d,\"ms\":%.3f,\"Grows_per_s\":%.2f,\"payload_GBps\":%.1f}\n",
             mode == 0 ? "f32x4_128Brow" : mode == 1 ? "bf16x8_64Brow" : mode == 2 ? "f32_scalar_128Brow" : "bf16x2_64Brow",
             sizes[si] >> 20, local_span, best, rows / best / 1e6, rows * rowb / best / 1e6);
    }
    // hot-spot: only 256 rows (coarse level contention)
    for (int mode = 0; mode < 2; ++mode) {
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        if (mode == 0) mb_red<0><<<grid, 256>>>(iters, 256 * 8, buf, 0); else mb_red<1><<<grid, 256>>>(iters, 256 * 8, buf, 0);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms = time_ms(e0, e1); if (ms < best) best = ms;
      }
      int lpr = mode == 0 ? 8 : 4; int rowb = mode == 0 ? 128 : 64;
      double rows = (double)grid * 256 / lpr * iters;
      printf("{\"test\":\"redg_hot2048rows_%s\",\"ms\":%.3f,\"Grows_per_s\":%.2f,\"payload_GBps\":%.1f}\n", mode == 0 ? "f32x4" : "bf16x8", best, rows / best / 1e6, rows * rowb / best / 1e6);
    }
    // MB5 bulk reduce
    int chunks[4] = {128, 1024, 8192, 32768};
    for (int ci = 0; ci < 4; ++ci) {
      int chunk = chunks[ci]; int it2 = chunk <= 1024 ? 2048 : 256;
      CK(cudaFuncSetAttribute(mb_bulk_red, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      uint32_t nchunks = (uint32_t)(sizes[0] / chunk);
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        mb_bulk_red<<<nsm * 2, 128, 65536>>>(it2, chunk, nchunks, buf);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms = time_ms(e0, e1); if (ms < best) best = ms;
      }
      double ops = (double)nsm * 2 * it2;
      printf("{\"test\":\"tma_bulk_reduce_add_f32\",\"chunk_B\":%d,\"ms\":%.3f,\"Mops_per_s\":%.1f,\"payload_GBps\":%.1f}\n", chunk, best, ops / best / 1e3, ops * chunk / best / 1e6);
    }
    // MB6 LDG gather
    for (int si = 0; si < 3; ++si) for (int loc = 0; loc < 2; ++loc) {
      uint32_t nrows = (uint32_t)(sizes[si] / 64);
      int local_span = loc ? 2048 : 0;
      float best = 1e30f; int it2 = 512;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        mb_ldg_gather<<<grid, 256>>>(it2, nrows, buf, sink, local_span);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms = time_ms(e0, e1); if (ms < best) best = ms;
      }
      double rows = (double)grid * 64 * it2;
      printf("{\"test\":\"ldg128_gather_64Brow\",\"buf_MB\":%zu,\"local_span_rows\":%d,\"ms\":%.3f,\"payload_GBps\":%.1f}\n", sizes[si] >> 20, local_span, best, rows * 64 / best / 1e6);
    }
    CK(cudaFree(buf));
  }
  return 0;
}
