// MB4 (round-2 design inputs): what an in-SM, tile-local pre-reduction of grad_value and a shared-memory value window
// could sustain.
//   A. conflict-free integer shared atomics: one warp adds one 128-byte row (lane = channel, int32 fixed point) at a random
//      row of a window -- ATOMS.ADD without a return value -- against the plain LDS + IADD + STS of the same rows.
//   B. 4-lane x 16-byte gathers of 64-byte rows from shared memory, (1) random rows, (2) rows ordered so that in every warp
//      instruction four groups read an even and four an odd row (the two x-neighbours of a bilinear footprint have opposite
//      parity when the window pitch is odd), i.e. no two groups share a 64-byte half of the banks more than 4 ways.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 4; }

template <int MODE>   // 0: atomicAdd (no return), 1: plain RMW, 2: atomicAdd with 4 independent rows in flight
__global__ void __launch_bounds__(512) mb_atoms_rows(int iters, int nrows, int* sink, long long* cyc) {
  extern __shared__ __align__(128) int sm[];
  for (int i = threadIdx.x; i < nrows * 32; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  uint32_t s = (blockIdx.x * 977 + (threadIdx.x >> 5)) * 2654435761u + 12345u;
  long long t0 = clock64();
  #pragma unroll 4
  for (int it = 0; it < iters; ++it) {
    uint32_t r = lcg(s) % (uint32_t)nrows;
    if (MODE == 1) { sm[r * 32 + lane] += it; }
    else atomicAdd(&sm[r * 32 + lane], it);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  int acc = 0;
  for (int i = threadIdx.x; i < nrows * 32; i += blockDim.x) acc ^= sm[i];
  if (acc == 0x7fffffff) sink[0] = acc;
}

template <int MODE>   // 0: random rows, 1: parity-balanced order
__global__ void __launch_bounds__(512) mb_lds_rows(int iters, int nrows, uint32_t* sink, long long* cyc) {
  extern __shared__ __align__(128) int sm[];
  for (int i = threadIdx.x; i < nrows * 16; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int sub = threadIdx.x & 3, grp = (threadIdx.x >> 2) & 7;
  uint32_t s = (blockIdx.x * 977 + (threadIdx.x >> 2)) * 2654435761u + 12345u;
  uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  long long t0 = clock64();
  #pragma unroll 2
  for (int it = 0; it < iters; ++it) {
    uint32_t r = lcg(s) % (uint32_t)(nrows - 1);
    uint32_t ra, rb;
    if (MODE == 0) { ra = r; rb = lcg(s) % (uint32_t)(nrows - 1); }
    else { uint32_t flip = (r ^ grp) & 1u; ra = r + flip; rb = r + (flip ^ 1u); }   // ra has the parity of grp, rb the other
    const uint4 u = *reinterpret_cast<const uint4*>(&sm[ra * 16 + sub * 4]);
    const uint4 v = *reinterpret_cast<const uint4*>(&sm[rb * 16 + sub * 4]);
    a0 += u.x ^ v.x; a1 += u.y ^ v.y; a2 += u.z ^ v.z; a3 += u.w ^ v.w;
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if ((a0 ^ a1 ^ a2 ^ a3) == 0x12345) sink[0] = a0;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  int* sink; long long* cyc; CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&cyc, nsm * sizeof(long long)));
  long long hc[1024];
  const int iters = 4096;
  {
    const int nrows = 841;                                   // a 29 x 29 window of 128-byte rows = 105 KB
    const int smem = nrows * 128;
    CK(cudaFuncSetAttribute(mb_atoms_rows<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(mb_atoms_rows<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int mode = 0; mode < 2; ++mode)
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) mb_atoms_rows<0><<<nsm, 512, smem>>>(iters, nrows, sink, cyc);
        else mb_atoms_rows<1><<<nsm, 512, smem>>>(iters, nrows, sink, cyc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hc, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost));
        if (rep) printf("{\"test\":\"%s\",\"rows_per_clk_per_sm\":%.3f,\"src\":\"tools/microbench/mb4.cu\"}\n",
                        mode == 0 ? "atoms_add_s32_128Brow_conflict_free" : "smem_rmw_s32_128Brow_racy", 16.0 * iters / hc[0]);
      }
  }
  {
    const int nrows = 1796;                                  // four windows of 64-byte rows = 115 KB
    const int smem = nrows * 64;
    CK(cudaFuncSetAttribute(mb_lds_rows<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(mb_lds_rows<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int mode = 0; mode < 2; ++mode)
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) mb_lds_rows<0><<<nsm, 512, smem>>>(iters, nrows, (uint32_t*)sink, cyc);
        else mb_lds_rows<1><<<nsm, 512, smem>>>(iters, nrows, (uint32_t*)sink, cyc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hc, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost));
        if (rep) printf("{\"test\":\"%s\",\"B_per_clk_per_sm\":%.1f,\"src\":\"tools/microbench/mb4.cu\"}\n",
                        mode == 0 ? "lds128_4lane_64Brow_random" : "lds128_4lane_64Brow_parity_balanced", 512.0 * 32 * iters / hc[0]);
      }
  }
  return 0;
}
