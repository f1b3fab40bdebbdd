// B200-native multi-scale deformable attention: direct-gather kernels + the C ABI (include/msda_b200.h).
//
// Two kernel families live here:
//   * "vec"  — the production path for D in {16, 32, 64, 128}: a group of G = D*sizeof(T)/16 lanes owns one
//              (batch, query, head) pair, every lane moves 16 bytes of a head's channel slice per corner,
//              sampling locations / attention weights of the warp's pairs are staged once in shared memory
//              with coalesced 128-bit loads, the weighted sum completes in registers.  Backward reduces
//              grad_sampling_loc / grad_attn_weight with warp shuffles inside the lane group (each element
//              is written exactly once, through shared memory, with coalesced stores) and scatters
//              grad_value with packed red.global.add (f32x4, or bf16x8 / f16x8 when asked).
//   * "any"  — compatibility path for every other channel count (upstream's test suite uses 30, 71, 1025,
//              2048, 3096) and for float64 (gradcheck): one warp per pair, lanes stride the channels.
//
// Replaces upstream ms_deform_im2col_cuda.cuh (ms_deformable_im2col_gpu_kernel and the six
// ms_deformable_col2im_gpu_kernel_* variants); see SURVEY.md §2b / §8a rows a6, a7.
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <type_traits>
#include <vector>

#include "msda_common.cuh"
#include "../../include/msda_b200.h"

namespace msda {

// =====================================================================================================
// Forward, vector path
// =====================================================================================================
template <typename T, int D>
__global__ void __launch_bounds__(kThreads)
msda_fwd_vec_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                    const float* __restrict__ attn, T* __restrict__ out,
                    int S, int M, int Lq, int L, int P, int total_pairs) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int G = D / VEC;        // lanes per (b, q, m) pair
  constexpr int GPW = 32 / G;       // pairs per warp
  static_assert(G >= 1 && G <= 32 && (G & (G - 1)) == 0, "lane group must be a power of two");

  extern __shared__ __align__(16) float smem[];
  __shared__ LevelMeta meta;
  load_level_meta(meta, shapes, lsi, L);

  const int LP = L * P;
  const int loc_stride = 2 * LP + 4;   // floats; +16 B pad keeps the per-group rows on distinct banks
  const int attn_stride = LP + 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / G, c = lane % G;
  float* wloc = smem + warp * GPW * (loc_stride + attn_stride);
  float* wattn = wloc + GPW * loc_stride;

  const int pair0 = (blockIdx.x * kWarps + warp) * GPW;
  const int nvalid = min(GPW, total_pairs - pair0);
  if (nvalid <= 0) return;

  // ---- stage this warp's sampling locations and attention weights (contiguous in global memory) ----
  {
    const float* gl = loc + static_cast<size_t>(pair0) * (2 * LP);
    const float* ga = attn + static_cast<size_t>(pair0) * LP;
    if ((LP & 3) == 0) {
      const int lv = LP / 2, av = LP / 4;     // float4s per pair row
      for (int i = lane; i < nvalid * lv; i += 32) {
        const int r = i / lv, k = i - r * lv;
        *reinterpret_cast<float4*>(wloc + r * loc_stride + 4 * k) = __ldg(reinterpret_cast<const float4*>(gl) + i);
      }
      for (int i = lane; i < nvalid * av; i += 32) {
        const int r = i / av, k = i - r * av;
        *reinterpret_cast<float4*>(wattn + r * attn_stride + 4 * k) = __ldg(reinterpret_cast<const float4*>(ga) + i);
      }
    } else {
      for (int i = lane; i < nvalid * 2 * LP; i += 32) {
        const int r = i / (2 * LP), k = i - r * 2 * LP;
        wloc[r * loc_stride + k] = __ldg(gl + i);
      }
      for (int i = lane; i < nvalid * LP; i += 32) {
        const int r = i / LP, k = i - r * LP;
        wattn[r * attn_stride + k] = __ldg(ga + i);
      }
    }
  }
  __syncwarp();
  if (g >= nvalid) return;

  const int pair = pair0 + g;
  const int m = pair % M;
  const int b = (pair / M) / Lq;
  const size_t pix_stride = static_cast<size_t>(M) * D;            // elements between neighbouring pixels
  const T* vb = value + static_cast<size_t>(b) * S * pix_stride + m * D + c * VEC;
  const float* myloc = wloc + g * loc_stride;
  const float* myattn = wattn + g * attn_stride;

  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

  for (int l = 0; l < L; ++l) {
    const int H = meta.H[l], W = meta.W[l];
    const T* vl = vb + static_cast<size_t>(meta.start[l]) * pix_stride;
#pragma unroll 4
    for (int p = 0; p < P; ++p) {
      const float2 xy = *reinterpret_cast<const float2*>(myloc + 2 * (l * P + p));
      const float a = myattn[l * P + p];
      const Footprint<float> fp = make_footprint<float>(xy.x, xy.y, H, W);
      const T* p00 = vl + (static_cast<ptrdiff_t>(fp.h0) * W + fp.w0) * static_cast<ptrdiff_t>(pix_stride);
      const uint4 zero = make_uint4(0, 0, 0, 0);
      const uint4 u00 = fp.v00 ? ldg16(p00) : zero;
      const uint4 u01 = fp.v01 ? ldg16(p00 + pix_stride) : zero;
      const uint4 u10 = fp.v10 ? ldg16(p00 + static_cast<size_t>(W) * pix_stride) : zero;
      const uint4 u11 = fp.v11 ? ldg16(p00 + static_cast<size_t>(W + 1) * pix_stride) : zero;
      const float w00 = fp.hh * fp.hw * a, w01 = fp.hh * fp.lw * a, w10 = fp.lh * fp.hw * a, w11 = fp.lh * fp.lw * a;
      float f[VEC];
      unpack16<T>(u00, f);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = fmaf(w00, f[i], acc[i]);
      unpack16<T>(u01, f);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = fmaf(w01, f[i], acc[i]);
      unpack16<T>(u10, f);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = fmaf(w10, f[i], acc[i]);
      unpack16<T>(u11, f);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = fmaf(w11, f[i], acc[i]);
    }
  }
  *reinterpret_cast<uint4*>(out + static_cast<size_t>(pair) * D + c * VEC) = pack16<T>(acc);
}

// =====================================================================================================
// Backward, vector path
// =====================================================================================================
// GV16 = false: grad_value contributions go to an fp32 buffer `gv32` laid out like value
//               (grad_value itself for T = float, the caller's scratch for 16-bit T) with red.v4.f32.
// GV16 = true : 16-bit T only; contributions are rounded to T and added with packed 16-bit red.
template <typename T, int D, bool GV16>
__global__ void __launch_bounds__(kThreads)
msda_bwd_vec_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                    const float* __restrict__ attn, const T* __restrict__ grad_out,
                    float* __restrict__ gv32, T* __restrict__ gv16,
                    float* __restrict__ grad_loc, float* __restrict__ grad_attn,
                    int S, int M, int Lq, int L, int P, int total_pairs) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int G = D / VEC;
  constexpr int GPW = 32 / G;

  extern __shared__ __align__(16) float smem[];
  __shared__ LevelMeta meta;
  load_level_meta(meta, shapes, lsi, L);

  const int LP = L * P;
  const int loc_stride = 2 * LP + 4;
  const int attn_stride = LP + 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / G, c = lane % G;
  float* wloc = smem + warp * GPW * (loc_stride + attn_stride);
  float* wattn = wloc + GPW * loc_stride;

  const int pair0 = (blockIdx.x * kWarps + warp) * GPW;
  const int nvalid = min(GPW, total_pairs - pair0);
  if (nvalid <= 0) return;

  const float* gl = loc + static_cast<size_t>(pair0) * (2 * LP);
  const float* ga = attn + static_cast<size_t>(pair0) * LP;
  const bool vec_rows = (LP & 3) == 0;
  if (vec_rows) {
    const int lv = LP / 2, av = LP / 4;
    for (int i = lane; i < nvalid * lv; i += 32) {
      const int r = i / lv, k = i - r * lv;
      *reinterpret_cast<float4*>(wloc + r * loc_stride + 4 * k) = __ldg(reinterpret_cast<const float4*>(gl) + i);
    }
    for (int i = lane; i < nvalid * av; i += 32) {
      const int r = i / av, k = i - r * av;
      *reinterpret_cast<float4*>(wattn + r * attn_stride + 4 * k) = __ldg(reinterpret_cast<const float4*>(ga) + i);
    }
  } else {
    for (int i = lane; i < nvalid * 2 * LP; i += 32) {
      const int r = i / (2 * LP), k = i - r * 2 * LP;
      wloc[r * loc_stride + k] = __ldg(gl + i);
    }
    for (int i = lane; i < nvalid * LP; i += 32) {
      const int r = i / LP, k = i - r * LP;
      wattn[r * attn_stride + k] = __ldg(ga + i);
    }
  }
  __syncwarp();

  // Lanes of padding groups (tail warp only) stay in the loop so that the full-mask shuffles are legal;
  // they read pair 0 of the warp and never write.
  const bool active = g < nvalid;
  const int pair = pair0 + (active ? g : 0);
  const int m = pair % M;
  const int b = (pair / M) / Lq;
  const size_t pix_stride = static_cast<size_t>(M) * D;
  const size_t img_off = static_cast<size_t>(b) * S * pix_stride + m * D + c * VEC;
  const T* vb = value + img_off;
  float* myloc = wloc + (active ? g : 0) * loc_stride;
  float* myattn = wattn + (active ? g : 0) * attn_stride;

  float go[VEC];
  unpack16<T>(ldg16(grad_out + static_cast<size_t>(pair) * D + c * VEC), go);

  for (int l = 0; l < L; ++l) {
    const int H = meta.H[l], W = meta.W[l];
    const size_t lvl_off = static_cast<size_t>(meta.start[l]) * pix_stride;
    const T* vl = vb + lvl_off;
#pragma unroll 2
    for (int p = 0; p < P; ++p) {
      const int lp = l * P + p;
      const float2 xy = *reinterpret_cast<const float2*>(myloc + 2 * lp);
      const float a = myattn[lp];
      const Footprint<float> fp = make_footprint<float>(xy.x, xy.y, H, W);
      const ptrdiff_t o00 = (static_cast<ptrdiff_t>(fp.h0) * W + fp.w0) * static_cast<ptrdiff_t>(pix_stride);
      const ptrdiff_t o01 = o00 + static_cast<ptrdiff_t>(pix_stride);
      const ptrdiff_t o10 = o00 + static_cast<ptrdiff_t>(W) * static_cast<ptrdiff_t>(pix_stride);
      const ptrdiff_t o11 = o10 + static_cast<ptrdiff_t>(pix_stride);
      const uint4 zero = make_uint4(0, 0, 0, 0);
      const uint4 u00 = fp.v00 ? ldg16(vl + o00) : zero;
      const uint4 u01 = fp.v01 ? ldg16(vl + o01) : zero;
      const uint4 u10 = fp.v10 ? ldg16(vl + o10) : zero;
      const uint4 u11 = fp.v11 ? ldg16(vl + o11) : zero;

      // ---- grad_value: corner_weight * attn * grad_out, scattered with packed reductions ----
      if (active) {
        const float w[4] = {fp.hh * fp.hw * a, fp.hh * fp.lw * a, fp.lh * fp.hw * a, fp.lh * fp.lw * a};
        const bool ok[4] = {fp.v00, fp.v01, fp.v10, fp.v11};
        const ptrdiff_t off[4] = {o00, o01, o10, o11};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (ok[k]) {
            float r[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) r[i] = w[k] * go[i];
            if constexpr (GV16) {
              red_add_16bit_x8<T>(gv16 + img_off + lvl_off + off[k], pack16<T>(r));
            } else {
              float* dst = gv32 + img_off + lvl_off + off[k];
#pragma unroll
              for (int i = 0; i < VEC; i += 4) red_add_f32x4(dst + i, r[i], r[i + 1], r[i + 2], r[i + 3]);
            }
          }
        }
      }

      // ---- per-corner dot products <v_c, grad_out>, reduced over the lane group ----
      float f[VEC];
      float d00 = 0.f, d01 = 0.f, d10 = 0.f, d11 = 0.f;
      unpack16<T>(u00, f);
#pragma unroll
      for (int i = 0; i < VEC; ++i) d00 = fmaf(f[i], go[i], d00);
      unpack16<T>(u01, f);
#pragma unroll
      for (int i = 0; i < VEC; ++i) d01 = fmaf(f[i], go[i], d01);
      unpack16<T>(u10, f);
#pragma unroll
      for (int i = 0; i < VEC; ++i) d10 = fmaf(f[i], go[i], d10);
      unpack16<T>(u11, f);
#pragma unroll
      for (int i = 0; i < VEC; ++i) d11 = fmaf(f[i], go[i], d11);
#pragma unroll
      for (int s = G / 2; s >= 1; s >>= 1) {
        d00 += __shfl_xor_sync(0xffffffffu, d00, s);
        d01 += __shfl_xor_sync(0xffffffffu, d01, s);
        d10 += __shfl_xor_sync(0xffffffffu, d10, s);
        d11 += __shfl_xor_sync(0xffffffffu, d11, s);
      }
      // grad_attn = bilinear(value) . grad_out ; grad_loc = attn * (W, H) * d(bilinear)/d(w, h) . grad_out
      const float g_attn = fp.hh * fp.hw * d00 + fp.hh * fp.lw * d01 + fp.lh * fp.hw * d10 + fp.lh * fp.lw * d11;
      const float g_x = static_cast<float>(W) * a * (fp.hh * (d01 - d00) + fp.lh * (d11 - d10));
      const float g_y = static_cast<float>(H) * a * (fp.hw * (d10 - d00) + fp.lw * (d11 - d01));
      __syncwarp();
      if (active && c == 0) {   // in place: every lane of the group has already read (x, y, a) of this point
        *reinterpret_cast<float2*>(myloc + 2 * lp) = make_float2(g_x, g_y);
        myattn[lp] = g_attn;
      }
    }
  }
  __syncwarp();

  // ---- coalesced write-back of the staged gradients ----
  float* ol = grad_loc + static_cast<size_t>(pair0) * (2 * LP);
  float* oa = grad_attn + static_cast<size_t>(pair0) * LP;
  if (vec_rows) {
    const int lv = LP / 2, av = LP / 4;
    for (int i = lane; i < nvalid * lv; i += 32) {
      const int r = i / lv, k = i - r * lv;
      reinterpret_cast<float4*>(ol)[i] = *reinterpret_cast<const float4*>(wloc + r * loc_stride + 4 * k);
    }
    for (int i = lane; i < nvalid * av; i += 32) {
      const int r = i / av, k = i - r * av;
      reinterpret_cast<float4*>(oa)[i] = *reinterpret_cast<const float4*>(wattn + r * attn_stride + 4 * k);
    }
  } else {
    for (int i = lane; i < nvalid * 2 * LP; i += 32) {
      const int r = i / (2 * LP), k = i - r * 2 * LP;
      ol[i] = wloc[r * loc_stride + k];
    }
    for (int i = lane; i < nvalid * LP; i += 32) {
      const int r = i / LP, k = i - r * LP;
      oa[i] = wattn[r * attn_stride + k];
    }
  }
}

// fp32 accumulation buffer -> 16-bit grad_value (one rounding per element)
template <typename T>
__global__ void __launch_bounds__(256)
msda_round_scratch_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t n8) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    reinterpret_cast<uint4*>(dst)[i] = pack16<T>(f);
  }
}

// =====================================================================================================
// Compatibility path: any D, any dtype (incl. float64).  One warp per (b, q, m) pair.
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(kThreads)
msda_fwd_any_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lsi, const typename Traits<T>::Aux* __restrict__ loc,
                    const typename Traits<T>::Aux* __restrict__ attn, T* __restrict__ out,
                    int S, int M, int D, int Lq, int L, int P, int total_pairs) {
  using A = typename Traits<T>::Acc;
  __shared__ LevelMeta meta;
  load_level_meta(meta, shapes, lsi, L);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * kWarps + warp;
  if (pair >= total_pairs) return;
  const int m = pair % M;
  const int b = (pair / M) / Lq;
  const size_t pix_stride = static_cast<size_t>(M) * D;
  const T* vb = value + static_cast<size_t>(b) * S * pix_stride + m * D;
  const auto* myloc = loc + static_cast<size_t>(pair) * L * P * 2;
  const auto* myattn = attn + static_cast<size_t>(pair) * L * P;
  for (int d0 = 0; d0 < D; d0 += 32) {
    const int d = d0 + lane;
    A acc = 0;
    for (int l = 0; l < L; ++l) {
      const int H = meta.H[l], W = meta.W[l];
      const T* vl = vb + static_cast<size_t>(meta.start[l]) * pix_stride;
      for (int p = 0; p < P; ++p) {
        const A x = myloc[2 * (l * P + p)], y = myloc[2 * (l * P + p) + 1], a = myattn[l * P + p];
        const Footprint<A> fp = make_footprint<A>(x, y, H, W);
        if (!fp.inside || d >= D) continue;
        const T* p00 = vl + (static_cast<ptrdiff_t>(fp.h0) * W + fp.w0) * static_cast<ptrdiff_t>(pix_stride) + d;
        const A v00 = fp.v00 ? to_acc<T>(p00[0]) : A(0);
        const A v01 = fp.v01 ? to_acc<T>(p00[pix_stride]) : A(0);
        const A v10 = fp.v10 ? to_acc<T>(p00[static_cast<size_t>(W) * pix_stride]) : A(0);
        const A v11 = fp.v11 ? to_acc<T>(p00[static_cast<size_t>(W + 1) * pix_stride]) : A(0);
        acc += a * (fp.hh * fp.hw * v00 + fp.hh * fp.lw * v01 + fp.lh * fp.hw * v10 + fp.lh * fp.lw * v11);
      }
    }
    if (d < D) out[static_cast<size_t>(pair) * D + d] = from_acc<T>(acc);
  }
}

// GVT = type of the grad_value accumulation buffer (T itself for float/double, float scratch for 16-bit T)
template <typename T, typename GVT>
__global__ void __launch_bounds__(kThreads)
msda_bwd_any_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lsi, const typename Traits<T>::Aux* __restrict__ loc,
                    const typename Traits<T>::Aux* __restrict__ attn, const T* __restrict__ grad_out,
                    GVT* __restrict__ gv, typename Traits<T>::Aux* __restrict__ grad_loc,
                    typename Traits<T>::Aux* __restrict__ grad_attn,
                    int S, int M, int D, int Lq, int L, int P, int total_pairs) {
  using A = typename Traits<T>::Acc;
  __shared__ LevelMeta meta;
  load_level_meta(meta, shapes, lsi, L);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * kWarps + warp;
  if (pair >= total_pairs) return;       // warp-uniform
  const int m = pair % M;
  const int b = (pair / M) / Lq;
  const size_t pix_stride = static_cast<size_t>(M) * D;
  const size_t img_off = static_cast<size_t>(b) * S * pix_stride + m * D;
  const auto* myloc = loc + static_cast<size_t>(pair) * L * P * 2;
  const auto* myattn = attn + static_cast<size_t>(pair) * L * P;
  const T* mygo = grad_out + static_cast<size_t>(pair) * D;
  for (int l = 0; l < L; ++l) {
    const int H = meta.H[l], W = meta.W[l];
    const size_t lvl_off = img_off + static_cast<size_t>(meta.start[l]) * pix_stride;
    for (int p = 0; p < P; ++p) {
      const int lp = l * P + p;
      const A x = myloc[2 * lp], y = myloc[2 * lp + 1], a = myattn[lp];
      const Footprint<A> fp = make_footprint<A>(x, y, H, W);
      A d00 = 0, d01 = 0, d10 = 0, d11 = 0;
      if (fp.inside) {
        const ptrdiff_t o00 = (static_cast<ptrdiff_t>(fp.h0) * W + fp.w0) * static_cast<ptrdiff_t>(pix_stride);
        const ptrdiff_t o01 = o00 + static_cast<ptrdiff_t>(pix_stride);
        const ptrdiff_t o10 = o00 + static_cast<ptrdiff_t>(W) * static_cast<ptrdiff_t>(pix_stride);
        const ptrdiff_t o11 = o10 + static_cast<ptrdiff_t>(pix_stride);
        for (int d = lane; d < D; d += 32) {
          const A go = to_acc<T>(mygo[d]);
          const T* vp = value + lvl_off + d;
          GVT* gp = gv + lvl_off + d;
          if (fp.v00) { d00 += to_acc<T>(vp[o00]) * go; atomicAdd(gp + o00, static_cast<GVT>(fp.hh * fp.hw * a * go)); }
          if (fp.v01) { d01 += to_acc<T>(vp[o01]) * go; atomicAdd(gp + o01, static_cast<GVT>(fp.hh * fp.lw * a * go)); }
          if (fp.v10) { d10 += to_acc<T>(vp[o10]) * go; atomicAdd(gp + o10, static_cast<GVT>(fp.lh * fp.hw * a * go)); }
          if (fp.v11) { d11 += to_acc<T>(vp[o11]) * go; atomicAdd(gp + o11, static_cast<GVT>(fp.lh * fp.lw * a * go)); }
        }
      }
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        d00 += __shfl_xor_sync(0xffffffffu, d00, s);
        d01 += __shfl_xor_sync(0xffffffffu, d01, s);
        d10 += __shfl_xor_sync(0xffffffffu, d10, s);
        d11 += __shfl_xor_sync(0xffffffffu, d11, s);
      }
      if (lane == 0) {
        grad_attn[static_cast<size_t>(pair) * L * P + lp] =
            fp.hh * fp.hw * d00 + fp.hh * fp.lw * d01 + fp.lh * fp.hw * d10 + fp.lh * fp.lw * d11;
        grad_loc[(static_cast<size_t>(pair) * L * P + lp) * 2] = static_cast<A>(W) * a * (fp.hh * (d01 - d00) + fp.lh * (d11 - d10));
        grad_loc[(static_cast<size_t>(pair) * L * P + lp) * 2 + 1] = static_cast<A>(H) * a * (fp.hw * (d10 - d00) + fp.lw * (d11 - d01));
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
msda_round_scratch_any_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = from_acc<T>(src[i]);
}

// =====================================================================================================
// Host side
// =====================================================================================================
static thread_local int g_last_launches = 0;
static std::atomic<long long> g_total_launches{0};

// ---- optional per-launch timing of the dominant kernels (bench.py's roofline leg) --------------------
struct ProfileRecord { cudaEvent_t start, stop; int kind; };
// process-wide (autograd runs backward on its own thread), guarded by a mutex; off by default
static std::atomic<bool> g_profile_on{false};
static std::mutex g_profile_mu;
static std::vector<ProfileRecord> g_profile;

struct ScopedKernelTimer {
  cudaStream_t st; cudaEvent_t stop = nullptr; bool on = false;
  ScopedKernelTimer(int kind, cudaStream_t s) : st(s) {
    if (!g_profile_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_profile_mu);
    if (g_profile.size() >= 65536) return;
    ProfileRecord r; r.kind = kind;
    if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
    cudaEventRecord(r.start, st);
    stop = r.stop; on = true;
    g_profile.push_back(r);
  }
  ~ScopedKernelTimer() { if (on) cudaEventRecord(stop, st); }
};

struct Problem {
  int N, S, M, D, Lq, L, P;
  int total_pairs;
};

static int validate(const Problem& pr, int dtype, int im2col_step) {
  if (pr.N <= 0 || pr.S <= 0 || pr.M <= 0 || pr.D <= 0 || pr.Lq <= 0 || pr.L <= 0 || pr.P <= 0) return MSDA_ERR_BAD_SHAPE;
  if (pr.L > kMaxLevels) return MSDA_ERR_BAD_SHAPE;
  const long long pairs = static_cast<long long>(pr.N) * pr.Lq * pr.M;
  if (pairs >= (1ll << 31) - 64) return MSDA_ERR_BAD_SHAPE;
  if (static_cast<long long>(pr.L) * pr.P > 4096) return MSDA_ERR_BAD_SHAPE;
  if (dtype < MSDA_F32 || dtype > MSDA_F16) return MSDA_ERR_BAD_DTYPE;
  if (im2col_step <= 0) return MSDA_ERR_IM2COL_STEP;
  const int step = pr.N < im2col_step ? pr.N : im2col_step;
  if (pr.N % step != 0) return MSDA_ERR_IM2COL_STEP;
  return MSDA_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T> static bool vec_supported(int D) { return D == 16 || D == 32 || D == 64 || D == 128; }

template <typename T, int D>
static size_t vec_smem_bytes(int L, int P) {
  constexpr int G = D / (16 / static_cast<int>(sizeof(T)));
  constexpr int GPW = 32 / G;
  return static_cast<size_t>(kWarps) * GPW * (3 * L * P + 8) * sizeof(float);
}

template <typename K>
static cudaError_t allow_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
}

template <typename T, int D>
static int launch_fwd_vec(const Problem& pr, const void* value, const int64_t* shapes, const int64_t* lsi,
                          const void* loc, const void* attn, void* out, cudaStream_t st) {
  constexpr int GPW = 32 / (D / (16 / static_cast<int>(sizeof(T))));
  const size_t smem = vec_smem_bytes<T, D>(pr.L, pr.P);
  if (smem > 200 * 1024) return MSDA_ERR_BAD_SHAPE;
  cudaError_t e = allow_smem(msda_fwd_vec_kernel<T, D>, smem);
  if (e != cudaSuccess) return static_cast<int>(e);
  const int per_cta = kWarps * GPW;
  const int grid = (pr.total_pairs + per_cta - 1) / per_cta;
  ScopedKernelTimer timer(MSDA_KERNEL_FORWARD, st);
  msda_fwd_vec_kernel<T, D><<<grid, kThreads, smem, st>>>(
      static_cast<const T*>(value), shapes, lsi, static_cast<const float*>(loc), static_cast<const float*>(attn),
      static_cast<T*>(out), pr.S, pr.M, pr.Lq, pr.L, pr.P, pr.total_pairs);
  ++g_last_launches, ++g_total_launches;
  return static_cast<int>(cudaGetLastError());
}

template <typename T>
static int launch_fwd(const Problem& pr, const void* value, const int64_t* shapes, const int64_t* lsi,
                      const void* loc, const void* attn, void* out, cudaStream_t st) {
  if constexpr (!std::is_same<T, double>::value) {
    if (vec_supported<T>(pr.D)) {
      switch (pr.D) {
        case 16: return launch_fwd_vec<T, 16>(pr, value, shapes, lsi, loc, attn, out, st);
        case 32: return launch_fwd_vec<T, 32>(pr, value, shapes, lsi, loc, attn, out, st);
        case 64: return launch_fwd_vec<T, 64>(pr, value, shapes, lsi, loc, attn, out, st);
        case 128: return launch_fwd_vec<T, 128>(pr, value, shapes, lsi, loc, attn, out, st);
      }
    }
  }
  using Aux = typename Traits<T>::Aux;
  const int grid = (pr.total_pairs + kWarps - 1) / kWarps;
  msda_fwd_any_kernel<T><<<grid, kThreads, 0, st>>>(
      static_cast<const T*>(value), shapes, lsi, static_cast<const Aux*>(loc), static_cast<const Aux*>(attn),
      static_cast<T*>(out), pr.S, pr.M, pr.D, pr.Lq, pr.L, pr.P, pr.total_pairs);
  ++g_last_launches, ++g_total_launches;
  return static_cast<int>(cudaGetLastError());
}

template <typename T, int D>
static int launch_bwd_vec(const Problem& pr, const void* value, const int64_t* shapes, const int64_t* lsi,
                          const void* loc, const void* attn, const void* go, float* gv32, void* gv16,
                          void* gloc, void* gattn, bool use16, cudaStream_t st) {
  constexpr int GPW = 32 / (D / (16 / static_cast<int>(sizeof(T))));
  const size_t smem = vec_smem_bytes<T, D>(pr.L, pr.P);
  if (smem > 200 * 1024) return MSDA_ERR_BAD_SHAPE;
  const int per_cta = kWarps * GPW;
  const int grid = (pr.total_pairs + per_cta - 1) / per_cta;
  cudaError_t e;
  if constexpr (sizeof(T) == 2) {
    if (use16) {
      e = allow_smem(msda_bwd_vec_kernel<T, D, true>, smem);
      if (e != cudaSuccess) return static_cast<int>(e);
      ScopedKernelTimer timer(MSDA_KERNEL_BACKWARD, st);
      msda_bwd_vec_kernel<T, D, true><<<grid, kThreads, smem, st>>>(
          static_cast<const T*>(value), shapes, lsi, static_cast<const float*>(loc), static_cast<const float*>(attn),
          static_cast<const T*>(go), nullptr, static_cast<T*>(gv16), static_cast<float*>(gloc),
          static_cast<float*>(gattn), pr.S, pr.M, pr.Lq, pr.L, pr.P, pr.total_pairs);
      ++g_last_launches, ++g_total_launches;
      return static_cast<int>(cudaGetLastError());
    }
  }
  e = allow_smem(msda_bwd_vec_kernel<T, D, false>, smem);
  if (e != cudaSuccess) return static_cast<int>(e);
  ScopedKernelTimer timer(MSDA_KERNEL_BACKWARD, st);
  msda_bwd_vec_kernel<T, D, false><<<grid, kThreads, smem, st>>>(
      static_cast<const T*>(value), shapes, lsi, static_cast<const float*>(loc), static_cast<const float*>(attn),
      static_cast<const T*>(go), gv32, nullptr, static_cast<float*>(gloc), static_cast<float*>(gattn),
      pr.S, pr.M, pr.Lq, pr.L, pr.P, pr.total_pairs);
  ++g_last_launches, ++g_total_launches;
  return static_cast<int>(cudaGetLastError());
}

template <typename T>
static int launch_bwd(const Problem& pr, const void* value, const int64_t* shapes, const int64_t* lsi,
                      const void* loc, const void* attn, const void* go, void* gv, void* gloc, void* gattn,
                      void* scratch, int flags, cudaStream_t st) {
  const size_t n_value = static_cast<size_t>(pr.N) * pr.S * pr.M * pr.D;
  constexpr bool k16 = sizeof(T) == 2;
  const bool use16 = k16 && (flags & MSDA_BWD_GRAD_VALUE_16BIT_ATOMICS) && vec_supported<T>(pr.D);
  cudaError_t e = cudaMemsetAsync(gv, 0, n_value * sizeof(T), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (k16 && !use16) {
    e = cudaMemsetAsync(scratch, 0, n_value * sizeof(float), st);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  int rc;
  bool done = false;
  if constexpr (!std::is_same<T, double>::value) {
    if (vec_supported<T>(pr.D)) {
      float* gv32 = k16 ? static_cast<float*>(scratch) : static_cast<float*>(gv);
      switch (pr.D) {
        case 16: rc = launch_bwd_vec<T, 16>(pr, value, shapes, lsi, loc, attn, go, gv32, gv, gloc, gattn, use16, st); break;
        case 32: rc = launch_bwd_vec<T, 32>(pr, value, shapes, lsi, loc, attn, go, gv32, gv, gloc, gattn, use16, st); break;
        case 64: rc = launch_bwd_vec<T, 64>(pr, value, shapes, lsi, loc, attn, go, gv32, gv, gloc, gattn, use16, st); break;
        default: rc = launch_bwd_vec<T, 128>(pr, value, shapes, lsi, loc, attn, go, gv32, gv, gloc, gattn, use16, st); break;
      }
      done = true;
    }
  }
  if (!done) {
    using Aux = typename Traits<T>::Aux;
    const int grid = (pr.total_pairs + kWarps - 1) / kWarps;
    if constexpr (k16) {
      msda_bwd_any_kernel<T, float><<<grid, kThreads, 0, st>>>(
          static_cast<const T*>(value), shapes, lsi, static_cast<const Aux*>(loc), static_cast<const Aux*>(attn),
          static_cast<const T*>(go), static_cast<float*>(scratch), static_cast<Aux*>(gloc), static_cast<Aux*>(gattn),
          pr.S, pr.M, pr.D, pr.Lq, pr.L, pr.P, pr.total_pairs);
    } else {
      msda_bwd_any_kernel<T, T><<<grid, kThreads, 0, st>>>(
          static_cast<const T*>(value), shapes, lsi, static_cast<const Aux*>(loc), static_cast<const Aux*>(attn),
          static_cast<const T*>(go), static_cast<T*>(gv), static_cast<Aux*>(gloc), static_cast<Aux*>(gattn),
          pr.S, pr.M, pr.D, pr.Lq, pr.L, pr.P, pr.total_pairs);
    }
    ++g_last_launches, ++g_total_launches;
    rc = static_cast<int>(cudaGetLastError());
  }
  if (rc != 0) return rc;
  if constexpr (k16) {
    if (!use16) {
      if ((n_value & 7) == 0) {
        const size_t n8 = n_value / 8;
        const int grid = static_cast<int>(std::min<size_t>((n8 + 255) / 256, 148 * 16));
        msda_round_scratch_kernel<T><<<grid, 256, 0, st>>>(static_cast<const float*>(scratch), static_cast<T*>(gv), n8);
      } else {
        const int grid = static_cast<int>(std::min<size_t>((n_value + 255) / 256, 148 * 16));
        msda_round_scratch_any_kernel<T><<<grid, 256, 0, st>>>(static_cast<const float*>(scratch), static_cast<T*>(gv), n_value);
      }
      ++g_last_launches, ++g_total_launches;
      rc = static_cast<int>(cudaGetLastError());
    }
  }
  return rc;
}

}  // namespace msda

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace msda;

extern "C" int msda_abi_version(void) { return 1; }

extern "C" const char* msda_error_string(int code) {
  switch (code) {
    case MSDA_OK: return "success";
    case MSDA_ERR_NULL_POINTER: return "null pointer argument";
    case MSDA_ERR_BAD_SHAPE: return "invalid shape (non-positive dimension, more than 32 levels, or too many pairs)";
    case MSDA_ERR_BAD_DTYPE: return "unsupported value dtype";
    case MSDA_ERR_MISALIGNED: return "buffer is not 16-byte aligned";
    case MSDA_ERR_IM2COL_STEP: return "batch size must be divisible by min(batch, im2col_step)";
    case MSDA_ERR_SCRATCH_TOO_SMALL: return "scratch buffer missing or too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown msda error";
}

extern "C" int msda_last_launch_count(void) { return g_last_launches; }

extern "C" long long msda_total_launch_count(void) { return g_total_launches.load(); }

extern "C" int msda_profile_enable(int on) {
  g_profile_on.store(on != 0);
  return MSDA_OK;
}

extern "C" int msda_profile_collect(float* ms, int* kinds, int max_records) {
  std::lock_guard<std::mutex> lk(g_profile_mu);
  int n = 0;
  for (auto& r : g_profile) {
    float t = 0.f;
    if (cudaEventSynchronize(r.stop) == cudaSuccess && cudaEventElapsedTime(&t, r.start, r.stop) == cudaSuccess &&
        n < max_records && ms && kinds) {
      ms[n] = t; kinds[n] = r.kind; ++n;
    }
    cudaEventDestroy(r.start); cudaEventDestroy(r.stop);
  }
  g_profile.clear();
  return n;
}

extern "C" int msda_forward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                            const void* sampling_loc, const void* attn_weight, void* output,
                            int N, int S, int M, int D, int Lq, int L, int P,
                            int value_dtype, int im2col_step, void* stream) {
  g_last_launches = 0;
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !output) return MSDA_ERR_NULL_POINTER;
  Problem pr{N, S, M, D, Lq, L, P, 0};
  const int v = validate(pr, value_dtype, im2col_step);
  if (v != MSDA_OK) return v;
  pr.total_pairs = N * Lq * M;
  if (!aligned16(value) || !aligned16(sampling_loc) || !aligned16(attn_weight) || !aligned16(output)) return MSDA_ERR_MISALIGNED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (value_dtype) {
    case MSDA_F32: return launch_fwd<float>(pr, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, st);
    case MSDA_F64: return launch_fwd<double>(pr, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, st);
    case MSDA_BF16: return launch_fwd<__nv_bfloat16>(pr, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, st);
    case MSDA_F16: return launch_fwd<__half>(pr, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, st);
  }
  return MSDA_ERR_BAD_DTYPE;
}

extern "C" size_t msda_backward_scratch_bytes(int N, int S, int M, int D, int value_dtype, int flags) {
  if (value_dtype != MSDA_BF16 && value_dtype != MSDA_F16) return 0;
  const bool vec = (D == 16 || D == 32 || D == 64 || D == 128);
  if ((flags & MSDA_BWD_GRAD_VALUE_16BIT_ATOMICS) && vec) return 0;
  return static_cast<size_t>(N) * S * M * D * sizeof(float);
}

extern "C" int msda_backward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                             const void* sampling_loc, const void* attn_weight, const void* grad_output,
                             void* grad_value, void* grad_sampling_loc, void* grad_attn_weight,
                             void* scratch, size_t scratch_bytes,
                             int N, int S, int M, int D, int Lq, int L, int P,
                             int value_dtype, int im2col_step, int flags, void* stream) {
  g_last_launches = 0;
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !grad_output ||
      !grad_value || !grad_sampling_loc || !grad_attn_weight)
    return MSDA_ERR_NULL_POINTER;
  Problem pr{N, S, M, D, Lq, L, P, 0};
  const int v = validate(pr, value_dtype, im2col_step);
  if (v != MSDA_OK) return v;
  pr.total_pairs = N * Lq * M;
  if (!aligned16(value) || !aligned16(sampling_loc) || !aligned16(attn_weight) || !aligned16(grad_output) ||
      !aligned16(grad_value) || !aligned16(grad_sampling_loc) || !aligned16(grad_attn_weight) || !aligned16(scratch))
    return MSDA_ERR_MISALIGNED;
  const size_t need = msda_backward_scratch_bytes(N, S, M, D, value_dtype, flags);
  if (need > 0 && (!scratch || scratch_bytes < need)) return MSDA_ERR_SCRATCH_TOO_SMALL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (value_dtype) {
    case MSDA_F32:
      return launch_bwd<float>(pr, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                               grad_value, grad_sampling_loc, grad_attn_weight, scratch, flags, st);
    case MSDA_F64:
      return launch_bwd<double>(pr, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                                grad_value, grad_sampling_loc, grad_attn_weight, scratch, flags, st);
    case MSDA_BF16:
      return launch_bwd<__nv_bfloat16>(pr, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                                       grad_value, grad_sampling_loc, grad_attn_weight, scratch, flags, st);
    case MSDA_F16:
      return launch_bwd<__half>(pr, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                                grad_value, grad_sampling_loc, grad_attn_weight, scratch, flags, st);
  }
  return MSDA_ERR_BAD_DTYPE;
}
