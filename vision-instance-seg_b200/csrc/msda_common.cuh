// Shared device helpers for the B200 (sm_100a) multi-scale deformable attention kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace msda {

constexpr int kMaxLevels = 32;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---- type traits -----------------------------------------------------------------------------
template <typename T> struct Traits;
template <> struct Traits<float>         { using Acc = float;  using Aux = float;  };
template <> struct Traits<double>        { using Acc = double; using Aux = double; };
template <> struct Traits<__nv_bfloat16> { using Acc = float;  using Aux = float;  };
template <> struct Traits<__half>        { using Acc = float;  using Aux = float;  };

template <typename T> __device__ __forceinline__ typename Traits<T>::Acc to_acc(T v);
template <> __device__ __forceinline__ float  to_acc<float>(float v) { return v; }
template <> __device__ __forceinline__ double to_acc<double>(double v) { return v; }
template <> __device__ __forceinline__ float  to_acc<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float  to_acc<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_acc(typename Traits<T>::Acc v);
template <> __device__ __forceinline__ float  from_acc<float>(float v) { return v; }
template <> __device__ __forceinline__ double from_acc<double>(double v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_acc<__half>(float v) { return __float2half_rn(v); }

// ---- 16-byte vectors of T  <->  float registers -------------------------------------------------
template <typename T> struct Vec16 { static constexpr int kElems = 16 / sizeof(T); };

__device__ __forceinline__ uint4 ldg16(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

template <typename T> __device__ __forceinline__ void unpack16(const uint4& u, float* f);
template <> __device__ __forceinline__ void unpack16<float>(const uint4& u, float* f) {
  f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
}
template <> __device__ __forceinline__ void unpack16<__nv_bfloat16>(const uint4& u, float* f) {
  // a bf16 is the high half of an fp32: exact widening with one shift / one mask per element
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack16<__half>(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

template <typename T> __device__ __forceinline__ uint4 pack16(const float* f);
template <> __device__ __forceinline__ uint4 pack16<float>(const float* f) {
  return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
}
template <> __device__ __forceinline__ uint4 pack16<__nv_bfloat16>(const float* f) {
  uint4 u;
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  return u;
}
template <> __device__ __forceinline__ uint4 pack16<__half>(const float* f) {
  uint4 u;
  __half2 a = __floats2half2_rn(f[0], f[1]), b = __floats2half2_rn(f[2], f[3]);
  __half2 c = __floats2half2_rn(f[4], f[5]), d = __floats2half2_rn(f[6], f[7]);
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  return u;
}

// ---- packed reductions to global memory (no return value: REDG, resolved in L2) -----------------
__device__ __forceinline__ void red_add_f32x4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <typename T> __device__ __forceinline__ void red_add_16bit_x8(T* p, const uint4& u);
template <> __device__ __forceinline__ void red_add_16bit_x8<__nv_bfloat16>(__nv_bfloat16* p, const uint4& u) {
  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
template <> __device__ __forceinline__ void red_add_16bit_x8<__half>(__half* p, const uint4& u) {
  asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

// ---- level metadata: int64 device tensors -> shared int32 -----------------------------------------
struct LevelMeta {
  int H[kMaxLevels];
  int W[kMaxLevels];
  int start[kMaxLevels];
};

__device__ __forceinline__ void load_level_meta(LevelMeta& meta, const int64_t* __restrict__ shapes,
                                                const int64_t* __restrict__ lsi, int L) {
  if (threadIdx.x < L) {
    meta.H[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x]);
    meta.W[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x + 1]);
    meta.start[threadIdx.x] = static_cast<int>(lsi[threadIdx.x]);
  }
  __syncthreads();
}

// Bilinear footprint of one sampling point, in the upstream kernel's convention:
// h_im = y*H - 0.5, w_im = x*W - 0.5; the point contributes only if -1 < h_im < H and -1 < w_im < W;
// each corner is zero-padded on its own.
template <typename F>
struct Footprint {
  int h0, w0;          // top-left corner (may be -1)
  F lh, lw, hh, hw;    // fractional parts and their complements
  bool inside;         // point-level gate
  bool v00, v01, v10, v11;
};

template <typename F>
__device__ __forceinline__ Footprint<F> make_footprint(F x, F y, int H, int W) {
  Footprint<F> fp;
  const F h_im = y * static_cast<F>(H) - static_cast<F>(0.5);
  const F w_im = x * static_cast<F>(W) - static_cast<F>(0.5);
  fp.inside = (h_im > static_cast<F>(-1)) && (w_im > static_cast<F>(-1)) && (h_im < static_cast<F>(H)) && (w_im < static_cast<F>(W));
  const F hf = floor(h_im), wf = floor(w_im);
  fp.h0 = static_cast<int>(hf);
  fp.w0 = static_cast<int>(wf);
  fp.lh = h_im - hf; fp.lw = w_im - wf;
  fp.hh = static_cast<F>(1) - fp.lh; fp.hw = static_cast<F>(1) - fp.lw;
  const bool top = fp.h0 >= 0, bot = fp.h0 + 1 <= H - 1, left = fp.w0 >= 0, right = fp.w0 + 1 <= W - 1;
  fp.v00 = fp.inside && top && left;
  fp.v01 = fp.inside && top && right;
  fp.v10 = fp.inside && bot && left;
  fp.v11 = fp.inside && bot && right;
  return fp;
}

}  // namespace msda
