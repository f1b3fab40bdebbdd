// MB3: does pairing x-adjacent accumulator rows (8 lanes x 16 B = one 128-byte span at 64-byte alignment) raise the packed
// fp16 reduction rate over 4 lanes x 16 B = one 64-byte row per group?  Same payload per lane in every mode.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 4; }

// MODE 0: 64 B rows, 4 lanes;  1: 128 B spans at random 64 B alignment, 8 lanes;  2: 128 B spans, 128 B aligned, 8 lanes
template <int MODE>
__global__ void __launch_bounds__(256) mb_red16(int iters, uint32_t nrows64, uint8_t* buf) {
  constexpr int LPR = MODE == 0 ? 4 : 8;
  uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  int sub = threadIdx.x % LPR;
  uint32_t s = gid * 2654435761u + 12345u;
  for (int it = 0; it < iters; ++it) {
    uint32_t p = lcg(s) % (nrows64 - 2);
    if (MODE == 2) p &= ~1u;
    uint8_t* a = buf + (size_t)p * 64 + sub * 16;
    asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1,%2,%3,%4};" :: "l"(a), "r"(0x3c003c00u), "r"(0x3c003c00u), "r"(0x3c003c00u), "r"(0x3c003c00u) : "memory");
  }
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const size_t bytes = 48ull << 20;
  uint8_t* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
  const uint32_t nrows64 = bytes / 64;
  const int iters = 256, grid = prop.multiProcessorCount * 16;
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      if (mode == 0) mb_red16<0><<<grid, 256>>>(iters, nrows64, buf);
      else if (mode == 1) mb_red16<1><<<grid, 256>>>(iters, nrows64, buf);
      else mb_red16<2><<<grid, 256>>>(iters, nrows64, buf);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double lanes = (double)grid * 256 * iters;
      const double rows64 = lanes / 4;      // 64-byte row equivalents
      if (rep) printf("{\"test\":\"redg_f16x8\",\"mode\":\"%s\",\"ms\":%.3f,\"G64Brows_per_s\":%.2f,\"payload_GBps\":%.1f}\n",
                      mode == 0 ? "64B rows, 4 lanes" : mode == 1 ? "128B spans @64B align, 8 lanes" : "128B spans @128B align, 8 lanes",
                      ms, rows64 / ms / 1e6, rows64 * 64 / ms / 1e6);
    }
  }
  return 0;
}
