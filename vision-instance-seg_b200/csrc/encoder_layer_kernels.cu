// Memory-bound glue kernels of one pixel-decoder encoder layer (include/msda_encoder_b200.h).
//
// Every kernel here is a single streaming pass over HBM: rows of C (= d_model, 256 in every BASELINE config) fp32 /
// bf16 elements, 128-bit loads and stores, one warp per row for the LayerNorm pair (statistics by warp shuffles, the
// row never leaves registers), column sums accumulated in registers over a persistent grid and reduced
// deterministically in two stages.  They replace what stock torch runs under bf16 autocast as separate add / cast /
// layer_norm / relu-backward / sum kernels around the MSDeformAttn call of upstream
// MSDeformAttnTransformerEncoderLayer.forward.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <algorithm>

#include "../../include/msda_b200.h"
#include "../../include/msda_encoder_b200.h"

namespace msda_enc {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxBlocks = 148 * 8;          // persistent grids: at most 8 CTAs per SM of a B200 (scratch sizing bound)

// Persistent grids are sized to exactly one resident wave: (CTAs of this kernel that fit on an SM) x (SM count), capped by
// `cap` — a grid of 4 CTAs/SM for a kernel that only fits 3 runs a second, one-third-full wave (measured: LayerNorm
// backward 0.345 -> see DESIGN §3.5).  The two device queries are cached per kernel.
struct OccupancyEntry { const void* fn; size_t smem; int per_sm; };

template <typename K>
static int resident_grid(K kernel, size_t dyn_smem, long long wanted, int cap) {
  static thread_local OccupancyEntry cache[32];
  static thread_local int cached = 0, sms = 0;
  const void* key = reinterpret_cast<const void*>(kernel);
  int per_sm = 0;
  for (int i = 0; i < cached; ++i)
    if (cache[i].fn == key && cache[i].smem == dyn_smem) per_sm = cache[i].per_sm;
  if (per_sm == 0) {
    int dev = 0;
    if (sms == 0 && (cudaGetDevice(&dev) != cudaSuccess ||
                     cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1))
      sms = 148;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, dyn_smem) != cudaSuccess || per_sm < 1) per_sm = 2;
    if (cached < 32) cache[cached++] = OccupancyEntry{key, dyn_smem, per_sm};
  }
  const long long g = std::min<long long>(std::min<long long>(wanted, static_cast<long long>(per_sm) * sms), cap);
  return static_cast<int>(std::max<long long>(g, 1));
}

__device__ __forceinline__ float4 bf16x4_to_float4(const uint2& u) {
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ uint2 float4_to_bf16x4(const float4& v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// ---------------------------------------------------------------------------------------------------------
// out16 = bf16(a + b)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
add_cast_kernel(const float4* __restrict__ a, const float4* __restrict__ b, uint4* __restrict__ out, size_t n8) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 a0 = __ldcs(a + 2 * i), a1 = __ldcs(a + 2 * i + 1);
    const float4 b0 = __ldg(b + 2 * i), b1 = __ldg(b + 2 * i + 1);
    const uint2 lo = float4_to_bf16x4(make_float4(a0.x + b0.x, a0.y + b0.y, a0.z + b0.z, a0.w + b0.w));
    const uint2 hi = float4_to_bf16x4(make_float4(a1.x + b1.x, a1.y + b1.y, a1.z + b1.z, a1.w + b1.w));
    out[i] = make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
}

// ---------------------------------------------------------------------------------------------------------
// residual add + LayerNorm, one warp per row; NV = C / 128 float4 chunks per lane
// ---------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(kThreads)
add_layernorm_fwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ delta,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* __restrict__ y, __nv_bfloat16* __restrict__ y16,
                         float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows, float eps) {
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31;
  const long long warp0 = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * kWarps;
  float4 g[NV], bt[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    g[v] = __ldg(reinterpret_cast<const float4*>(gamma) + v * 32 + lane);
    bt[v] = __ldg(reinterpret_cast<const float4*>(beta) + v * 32 + lane);
  }
  for (long long row = warp0; row < rows; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * C);
    float4 t[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) t[v] = __ldcs(xr + v * 32 + lane);
    if (delta != nullptr) {
      const uint2* dr = reinterpret_cast<const uint2*>(delta + row * C);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 d = bf16x4_to_float4(__ldcs(dr + v * 32 + lane));
        t[v].x += d.x; t[v].y += d.y; t[v].z += d.z; t[v].w += d.w;
      }
    }
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) s += (t[v].x + t[v].y) + (t[v].z + t[v].w);
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      t[v].x -= mean; t[v].y -= mean; t[v].z -= mean; t[v].w -= mean;
      q += (t[v].x * t[v].x + t[v].y * t[v].y) + (t[v].z * t[v].z + t[v].w * t[v].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
    float4* yr = reinterpret_cast<float4*>(y + row * C);
    uint2* y16r = y16 ? reinterpret_cast<uint2*>(y16 + row * C) : nullptr;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 o = make_float4(fmaf(t[v].x * rstd, g[v].x, bt[v].x), fmaf(t[v].y * rstd, g[v].y, bt[v].y),
                                   fmaf(t[v].z * rstd, g[v].z, bt[v].z), fmaf(t[v].w * rstd, g[v].w, bt[v].w));
      yr[v * 32 + lane] = o;
      if (y16r) y16r[v * 32 + lane] = float4_to_bf16x4(o);
    }
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  }
}

// backward: dx = rstd * (gg - mean(gg) - xhat * mean(gg * xhat)),  gg = g * gamma;  dgamma += g * xhat;  dbeta += g
template <int NV>
__global__ void __launch_bounds__(kThreads)
add_layernorm_bwd_kernel(const float* __restrict__ gy, const __nv_bfloat16* __restrict__ gy16,
                         const float* __restrict__ x, const __nv_bfloat16* __restrict__ delta,
                         const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                         const float* __restrict__ gamma, float* __restrict__ dx, __nv_bfloat16* __restrict__ ddelta,
                         float* __restrict__ partials, long long rows) {
  constexpr int C = NV * 128;
  extern __shared__ __align__(16) float red[];          // [kWarps][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp0 = static_cast<long long>(blockIdx.x) * kWarps + warp;
  const long long nwarps = static_cast<long long>(gridDim.x) * kWarps;
  float4 gm[NV], dg[NV], db[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    gm[v] = __ldg(reinterpret_cast<const float4*>(gamma) + v * 32 + lane);
    dg[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = warp0; row < rows; row += nwarps) {
    float4 g[NV], xh[NV];
    const float4* xr = reinterpret_cast<const float4*>(x + row * C);
#pragma unroll
    for (int v = 0; v < NV; ++v) xh[v] = __ldcs(xr + v * 32 + lane);
    if (gy != nullptr) {
      const float4* gr = reinterpret_cast<const float4*>(gy + row * C);
#pragma unroll
      for (int v = 0; v < NV; ++v) g[v] = __ldcs(gr + v * 32 + lane);
    } else {
#pragma unroll
      for (int v = 0; v < NV; ++v) g[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (gy16 != nullptr) {
      const uint2* gr = reinterpret_cast<const uint2*>(gy16 + row * C);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 d = bf16x4_to_float4(__ldcs(gr + v * 32 + lane));
        g[v].x += d.x; g[v].y += d.y; g[v].z += d.z; g[v].w += d.w;
      }
    }
    if (delta != nullptr) {
      const uint2* dr = reinterpret_cast<const uint2*>(delta + row * C);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 d = bf16x4_to_float4(__ldcs(dr + v * 32 + lane));
        xh[v].x += d.x; xh[v].y += d.y; xh[v].z += d.z; xh[v].w += d.w;
      }
    }
    const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      xh[v].x = (xh[v].x - mean) * rstd; xh[v].y = (xh[v].y - mean) * rstd;
      xh[v].z = (xh[v].z - mean) * rstd; xh[v].w = (xh[v].w - mean) * rstd;
      dg[v].x = fmaf(g[v].x, xh[v].x, dg[v].x); dg[v].y = fmaf(g[v].y, xh[v].y, dg[v].y);
      dg[v].z = fmaf(g[v].z, xh[v].z, dg[v].z); dg[v].w = fmaf(g[v].w, xh[v].w, dg[v].w);
      db[v].x += g[v].x; db[v].y += g[v].y; db[v].z += g[v].z; db[v].w += g[v].w;
      g[v].x *= gm[v].x; g[v].y *= gm[v].y; g[v].z *= gm[v].z; g[v].w *= gm[v].w;
      s1 += (g[v].x + g[v].y) + (g[v].z + g[v].w);
      s2 += (g[v].x * xh[v].x + g[v].y * xh[v].y) + (g[v].z * xh[v].z + g[v].w * xh[v].w);
    }
    s1 = warp_sum(s1) * (1.0f / C);
    s2 = warp_sum(s2) * (1.0f / C);
    float4* dxr = reinterpret_cast<float4*>(dx + row * C);
    uint2* ddr = ddelta ? reinterpret_cast<uint2*>(ddelta + row * C) : nullptr;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 o = make_float4(rstd * (g[v].x - s1 - xh[v].x * s2), rstd * (g[v].y - s1 - xh[v].y * s2),
                                   rstd * (g[v].z - s1 - xh[v].z * s2), rstd * (g[v].w - s1 - xh[v].w * s2));
      dxr[v * 32 + lane] = o;
      if (ddr) ddr[v * 32 + lane] = float4_to_bf16x4(o);
    }
  }
  // CTA-level reduction of the per-warp dgamma / dbeta partial sums, then one row of `partials` per CTA
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    *reinterpret_cast<float4*>(red + (warp * 2 + 0) * C + (v * 32 + lane) * 4) = dg[v];
    *reinterpret_cast<float4*>(red + (warp * 2 + 1) * C + (v * 32 + lane) * 4) = db[v];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kThreads) {
    const int k = i / C, c = i - k * C;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[(w * 2 + k) * C + c];
    partials[static_cast<size_t>(blockIdx.x) * 2 * C + i] = s;
  }
}

// out[c] = sum_b partials[b][c] for c < n_cols; columns >= split go to out1 (dgamma | dbeta)
__global__ void __launch_bounds__(kThreads)
finalize_partials_kernel(const float* __restrict__ partials, int nb, int n_cols, float* __restrict__ out0,
                         float* __restrict__ out1, int split) {
  __shared__ float red[kWarps][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), lane_b = threadIdx.x >> 5;
  float s = 0.f;
  if (col < n_cols)
    for (int b = lane_b; b < nb; b += kWarps) s += partials[static_cast<size_t>(b) * n_cols + col];
  red[lane_b][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && col < n_cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += red[w][threadIdx.x];
    if (col < split) out0[col] = t; else out1[col - split] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------
// column sums of a (batch, rows_per_batch, C) bf16 tensor over a row range, optionally fused with ReLU backward
// ---------------------------------------------------------------------------------------------------------
// A CTA covers TC = blockDim-tile of 8-column groups (gridDim.y tiles over the columns) and kThreads / TC row lanes;
// gridDim.x CTAs stride over the rows.  Each thread keeps 8 fp32 sums; the CTA writes one partial row.
template <bool RELU>
__global__ void __launch_bounds__(kThreads)
colsum_kernel(__nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ h, float* __restrict__ partials,
              long long total_rows, long long span, long long row_begin, long long rows_per_batch, int C, int TC) {
  extern __shared__ __align__(16) float red[];          // [row lanes][TC * 8]
  const int CG = C / 8;
  const int RL = kThreads / TC;
  const int cg = blockIdx.y * TC + (threadIdx.x % TC);
  const int rl = threadIdx.x / TC;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (cg < CG) {
    for (long long r = static_cast<long long>(blockIdx.x) * RL + rl; r < total_rows; r += static_cast<long long>(gridDim.x) * RL) {
      const long long n = r / span;
      const long long row = n * rows_per_batch + row_begin + (r - n * span);
      uint4* gp = reinterpret_cast<uint4*>(g + row * C) + cg;
      uint4 u = RELU ? *gp : __ldcs(gp);
      if constexpr (RELU) {
        const uint4 hv = __ldcs(reinterpret_cast<const uint4*>(h + row * C) + cg);
        // bf16 > 0  <=>  sign bit clear and magnitude non-zero
        auto mask2 = [](uint32_t gv, uint32_t hh) {
          const uint32_t lo = ((hh & 0x8000u) == 0u && (hh & 0x7fffu) != 0u) ? 0x0000ffffu : 0u;
          const uint32_t hi = ((hh & 0x80000000u) == 0u && (hh & 0x7fff0000u) != 0u) ? 0xffff0000u : 0u;
          return gv & (lo | hi);
        };
        u.x = mask2(u.x, hv.x); u.y = mask2(u.y, hv.y); u.z = mask2(u.z, hv.z); u.w = mask2(u.w, hv.w);
        *gp = u;
      }
      acc[0] += __uint_as_float(u.x << 16); acc[1] += __uint_as_float(u.x & 0xffff0000u);
      acc[2] += __uint_as_float(u.y << 16); acc[3] += __uint_as_float(u.y & 0xffff0000u);
      acc[4] += __uint_as_float(u.z << 16); acc[5] += __uint_as_float(u.z & 0xffff0000u);
      acc[6] += __uint_as_float(u.w << 16); acc[7] += __uint_as_float(u.w & 0xffff0000u);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[(rl * TC + (threadIdx.x % TC)) * 8 + k] = acc[k];
  __syncthreads();
  for (int i = threadIdx.x; i < TC * 8; i += kThreads) {
    const int col = blockIdx.y * TC * 8 + i;
    if (col < C) {
      float s = 0.f;
      for (int w = 0; w < RL; ++w) s += red[w * TC * 8 + i];
      partials[static_cast<size_t>(blockIdx.x) * C + col] = s;
    }
  }
}

// Any column count (rows not 16-byte aligned): scalar bf16 accesses, 8 row lanes x 32 columns per CTA.  Only the bias
// gradients of very small sampling_offsets / attention_weights Linears (heads * levels * points not a multiple of 8) land here.
template <bool RELU>
__global__ void __launch_bounds__(kThreads)
colsum_scalar_kernel(__nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ h, float* __restrict__ partials,
                     long long total_rows, long long span, long long row_begin, long long rows_per_batch, int C) {
  __shared__ float red[kWarps][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.y * 32 + cl;
  float acc = 0.f;
  if (col < C) {
    for (long long r = static_cast<long long>(blockIdx.x) * kWarps + rl; r < total_rows; r += static_cast<long long>(gridDim.x) * kWarps) {
      const long long n = r / span;
      const long long row = n * rows_per_batch + row_begin + (r - n * span);
      float v = __bfloat162float(g[row * C + col]);
      if constexpr (RELU) {
        if (!(__bfloat162float(h[row * C + col]) > 0.f)) { v = 0.f; g[row * C + col] = __float2bfloat16_rn(0.f); }
      }
      acc += v;
    }
  }
  red[rl][cl] = acc;
  __syncthreads();
  if (threadIdx.x < 32 && col < C) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w][threadIdx.x];
    partials[static_cast<size_t>(blockIdx.x) * C + col] = s;
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int NV>
static int launch_ln_fwd(const float* x, const void* delta16, const float* gamma, const float* beta, float* y, void* y16,
                         float* mean, float* rstd, long long rows, float eps, cudaStream_t st) {
  const int grid = resident_grid(add_layernorm_fwd_kernel<NV>, 0, (rows + kWarps - 1) / kWarps, kMaxBlocks);
  add_layernorm_fwd_kernel<NV><<<grid, kThreads, 0, st>>>(x, static_cast<const __nv_bfloat16*>(delta16), gamma, beta, y,
                                                          static_cast<__nv_bfloat16*>(y16), mean, rstd, rows, eps);
  return static_cast<int>(cudaGetLastError());
}

constexpr int kMaxLnBwdBlocks = 148 * 4;     // scratch sizing bound of the dgamma / dbeta partials

template <int NV>
static int launch_ln_bwd(const float* gy, const void* gy16, const float* x, const void* delta16, const float* mean,
                         const float* rstd, const float* gamma, float* dx, void* ddelta16, float* dgamma, float* dbeta,
                         float* partials, long long rows, cudaStream_t st) {
  constexpr int C = NV * 128;
  const size_t smem = static_cast<size_t>(kWarps) * 2 * C * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(add_layernorm_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  const int grid = resident_grid(add_layernorm_bwd_kernel<NV>, smem, (rows + kWarps - 1) / kWarps, kMaxLnBwdBlocks);
  add_layernorm_bwd_kernel<NV><<<grid, kThreads, smem, st>>>(
      gy, static_cast<const __nv_bfloat16*>(gy16), x, static_cast<const __nv_bfloat16*>(delta16), mean, rstd, gamma, dx,
      static_cast<__nv_bfloat16*>(ddelta16), partials, rows);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  finalize_partials_kernel<<<2 * C / 32, kThreads, 0, st>>>(partials, grid, 2 * C, dgamma, dbeta, C);
  return static_cast<int>(cudaGetLastError());
}

static int colsum_tile(int C) {
  const int CG = C / 8;
  int TC = 1;
  while (TC * 2 <= CG && TC * 2 <= kThreads) TC *= 2;
  return TC;
}

template <bool RELU>
static int launch_colsum(void* g16, const void* h16, float* out, float* partials, long long batch, long long rows_per_batch,
                         long long row_begin, long long row_end, int C, cudaStream_t st) {
  const long long span = row_end - row_begin;
  const long long total = batch * span;
  if (C % 8 != 0) {
    const int tiles = (C + 31) / 32;
    const int resident = resident_grid(colsum_scalar_kernel<RELU>, 0, 1ll << 40, kMaxBlocks);
    long long gx = std::max<long long>(1, std::min<long long>((total + kWarps - 1) / kWarps, resident / tiles > 0 ? resident / tiles : 1));
    colsum_scalar_kernel<RELU><<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(tiles)), kThreads, 0, st>>>(
        static_cast<__nv_bfloat16*>(g16), static_cast<const __nv_bfloat16*>(h16), partials, total, span, row_begin, rows_per_batch, C);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
    finalize_partials_kernel<<<(C + 31) / 32, kThreads, 0, st>>>(partials, static_cast<int>(gx), C, out, out, C);
    return static_cast<int>(cudaGetLastError());
  }
  const int TC = colsum_tile(C);
  const int RL = kThreads / TC;
  const int tiles = (C / 8 + TC - 1) / TC;
  const size_t smem = static_cast<size_t>(kThreads) * 8 * sizeof(float);
  const int resident = resident_grid(colsum_kernel<RELU>, smem, 1ll << 40, kMaxBlocks);
  long long gx = (total + RL - 1) / RL;
  gx = std::max<long long>(1, std::min<long long>(gx, resident / tiles > 0 ? resident / tiles : 1));
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(tiles));
  colsum_kernel<RELU><<<grid, kThreads, smem, st>>>(static_cast<__nv_bfloat16*>(g16), static_cast<const __nv_bfloat16*>(h16),
                                                   partials, total, span, row_begin, rows_per_batch, C, TC);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  finalize_partials_kernel<<<(C + 31) / 32, kThreads, 0, st>>>(partials, static_cast<int>(gx), C, out, out, C);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace msda_enc

using namespace msda_enc;

extern "C" int msda_enc_add_cast(const float* a, const float* b, void* out16, size_t n, void* stream) {
  if (!a || !b || !out16) return MSDA_ERR_NULL_POINTER;
  if (n == 0 || (n & 7) != 0) return MSDA_ERR_BAD_SHAPE;
  if (!aligned16(a) || !aligned16(b) || !aligned16(out16)) return MSDA_ERR_MISALIGNED;
  const size_t n8 = n / 8;
  const int grid = resident_grid(add_cast_kernel, 0, static_cast<long long>((n8 + kThreads - 1) / kThreads), kMaxBlocks);
  add_cast_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), static_cast<uint4*>(out16), n8);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int msda_enc_add_layernorm_forward(const float* x, const void* delta16, const float* gamma, const float* beta,
                                              float* y, void* y16, float* mean, float* rstd,
                                              long long rows, int C, float eps, void* stream) {
  if (!x || !gamma || !beta || !y || !mean || !rstd) return MSDA_ERR_NULL_POINTER;
  if (rows <= 0 || C <= 0 || C % 128 != 0 || C > 1024) return MSDA_ERR_BAD_SHAPE;
  if (!aligned16(x) || !aligned16(delta16) || !aligned16(gamma) || !aligned16(beta) || !aligned16(y) || !aligned16(y16))
    return MSDA_ERR_MISALIGNED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (C / 128) {
    case 1: return launch_ln_fwd<1>(x, delta16, gamma, beta, y, y16, mean, rstd, rows, eps, st);
    case 2: return launch_ln_fwd<2>(x, delta16, gamma, beta, y, y16, mean, rstd, rows, eps, st);
    case 3: return launch_ln_fwd<3>(x, delta16, gamma, beta, y, y16, mean, rstd, rows, eps, st);
    case 4: return launch_ln_fwd<4>(x, delta16, gamma, beta, y, y16, mean, rstd, rows, eps, st);
    case 6: return launch_ln_fwd<6>(x, delta16, gamma, beta, y, y16, mean, rstd, rows, eps, st);
    case 8: return launch_ln_fwd<8>(x, delta16, gamma, beta, y, y16, mean, rstd, rows, eps, st);
  }
  return MSDA_ERR_BAD_SHAPE;
}

extern "C" size_t msda_enc_add_layernorm_backward_scratch_bytes(int C) {
  if (C <= 0) return 0;
  return static_cast<size_t>(kMaxLnBwdBlocks) * 2 * C * sizeof(float);
}

extern "C" int msda_enc_add_layernorm_backward(const float* gy, const void* gy16, const float* x, const void* delta16,
                                               const float* mean, const float* rstd, const float* gamma,
                                               float* dx, void* ddelta16, float* dgamma, float* dbeta,
                                               void* partials, size_t partials_bytes,
                                               long long rows, int C, void* stream) {
  if (!x || !mean || !rstd || !gamma || !dx || !dgamma || !dbeta || !partials || (!gy && !gy16)) return MSDA_ERR_NULL_POINTER;
  if (rows <= 0 || C <= 0 || C % 128 != 0 || C > 1024) return MSDA_ERR_BAD_SHAPE;
  if (partials_bytes < msda_enc_add_layernorm_backward_scratch_bytes(C)) return MSDA_ERR_SCRATCH_TOO_SMALL;
  if (!aligned16(gy) || !aligned16(gy16) || !aligned16(x) || !aligned16(delta16) || !aligned16(gamma) || !aligned16(dx) ||
      !aligned16(ddelta16) || !aligned16(partials))
    return MSDA_ERR_MISALIGNED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* p = static_cast<float*>(partials);
  switch (C / 128) {
    case 1: return launch_ln_bwd<1>(gy, gy16, x, delta16, mean, rstd, gamma, dx, ddelta16, dgamma, dbeta, p, rows, st);
    case 2: return launch_ln_bwd<2>(gy, gy16, x, delta16, mean, rstd, gamma, dx, ddelta16, dgamma, dbeta, p, rows, st);
    case 3: return launch_ln_bwd<3>(gy, gy16, x, delta16, mean, rstd, gamma, dx, ddelta16, dgamma, dbeta, p, rows, st);
    case 4: return launch_ln_bwd<4>(gy, gy16, x, delta16, mean, rstd, gamma, dx, ddelta16, dgamma, dbeta, p, rows, st);
    case 6: return launch_ln_bwd<6>(gy, gy16, x, delta16, mean, rstd, gamma, dx, ddelta16, dgamma, dbeta, p, rows, st);
    case 8: return launch_ln_bwd<8>(gy, gy16, x, delta16, mean, rstd, gamma, dx, ddelta16, dgamma, dbeta, p, rows, st);
  }
  return MSDA_ERR_BAD_SHAPE;
}

extern "C" size_t msda_enc_colsum_scratch_bytes(int C) {
  if (C <= 0) return 0;
  return static_cast<size_t>(kMaxBlocks) * ((C + 31) / 32 * 32) * sizeof(float);
}

static int colsum_common(bool relu, void* g16, const void* h16, float* out, void* scratch, size_t scratch_bytes,
                         long long batch, long long rows_per_batch, long long row_begin, long long row_end, int C, void* stream) {
  if (!g16 || !out || !scratch || (relu && !h16)) return MSDA_ERR_NULL_POINTER;
  if (batch <= 0 || rows_per_batch <= 0 || row_begin < 0 || row_end > rows_per_batch || row_end <= row_begin || C <= 0 ||
      C > (1 << 20))
    return MSDA_ERR_BAD_SHAPE;
  if (scratch_bytes < msda_enc_colsum_scratch_bytes(C)) return MSDA_ERR_SCRATCH_TOO_SMALL;
  if (C % 8 == 0 && (!aligned16(g16) || !aligned16(h16))) return MSDA_ERR_MISALIGNED;
  if (!aligned16(scratch)) return MSDA_ERR_MISALIGNED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (relu)
    return launch_colsum<true>(g16, h16, out, static_cast<float*>(scratch), batch, rows_per_batch, row_begin, row_end, C, st);
  return launch_colsum<false>(g16, h16, out, static_cast<float*>(scratch), batch, rows_per_batch, row_begin, row_end, C, st);
}

extern "C" int msda_enc_colsum(const void* g16, float* out, void* scratch, size_t scratch_bytes,
                               long long batch, long long rows_per_batch, long long row_begin, long long row_end, int C,
                               void* stream) {
  return colsum_common(false, const_cast<void*>(g16), nullptr, out, scratch, scratch_bytes, batch, rows_per_batch, row_begin,
                       row_end, C, stream);
}

extern "C" int msda_enc_relu_bwd_colsum(void* g16, const void* h16, float* out, void* scratch, size_t scratch_bytes,
                                        long long rows, int C, void* stream) {
  return colsum_common(true, g16, h16, out, scratch, scratch_bytes, 1, rows, 0, rows, C, stream);
}
