// Shared device helpers for the B200 (sm_100a) multi-scale deformable attention kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace msda {

constexpr int kMaxLevels = 32;
#ifndef MSDA_SPARSE_FACTOR
#define MSDA_SPARSE_FACTOR 4
#endif
// a level is "sparse" (its grad_value contributions are added directly, see build_accum_layout) when
// kSparseFactor * Lq * P <= H_l * W_l, i.e. it expects at most 4 / kSparseFactor corner rows per pixel row.
// 4 (one expected add per row) rather than 2: with 2 the 64x64 level of the cfg4 decoder shape also qualified, and
// when all 300 queries crowd into 3 % of the image the packed bf16 adds into that level lost up to 15 % of
// grad_value (tests/dev/gpu_sparse_accuracy.py); with 4 the direct mode stays within a few 1e-3 of the fp16 buckets
// over the whole clustering range measured
constexpr long long kSparseFactor = MSDA_SPARSE_FACTOR;
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
  // bucketed fp16 accumulation of grad_value (see AccumLayout below)
  int accK[kMaxLevels];      // number of private copies ("buckets") of the level; 0 = sparse level, added directly
  int accBase[kMaxLevels];   // first pixel row of the level's bucket 0 inside one image of the accumulator
  int accStride;             // pixel rows per image of the accumulator
  int dirOff[kMaxLevels];    // sparse levels: first row of the level among one image's directly accumulated rows
  int dirRows;               // directly accumulated pixel rows per image
  int bktOff[kMaxLevels];    // bucketed levels: first row of the level among one image's bucketed pixel rows
  int bktRows;               // bucketed pixel rows per image (dirRows + bktRows = S)
  uint32_t sortMask;         // hybrid backward: levels whose grad_value comes from the sorting kernel (one copy each)
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

// ---------------------------------------------------------------------------------------------------
// Bucketed fp16 accumulation layout.  An fp16 atomic add rounds the running sum to 11 bits, so the error of an
// accumulator grows like sqrt(number of adds): measured 3e-3 of max|grad_value| at ~20 adds per element but
// 1-2.5e-2 at ~1 360 (the 16x16 level of the 1024^2 encoder shape).  The accumulator therefore keeps
// K_l = ceil(expected adds per element of level l / depth) private copies of level l, a query adds into copy
// (q mod K_l), and the rounding pass sums the copies in fp32.  expected adds = ceil(Lq*P / (H_l*W_l)).
// Everything is derived on the device from the int64 shape tensor; the host only needs the upper bound
// accum_rows_bound() to size the buffer.
//
// Sparse levels (decoder cross-attention: a few hundred queries against 10^4 pixels).  When a level expects at most
// one corner row per pixel row (four corner rows per point: 4*Lq*P <= H_l*W_l) almost every element receives
// zero or one add, and a packed 16-bit add straight into grad_value is then as accurate as accumulating anywhere
// else and rounding once.  Such a level gets K_l = 0: its contributions are added, unscaled, into the (zeroed)
// grad_value rows in the value dtype, it owns no accumulator rows, and the rounding pass skips it -- the dense zero /
// sum / round traffic of the fp16 accumulator (which dominates a decoder layer's backward) is paid only for the
// coarse levels.
// ---------------------------------------------------------------------------------------------------
__host__ __device__ inline long long accum_rows_bound(int S, int L, int Lq, int P, int depth) {
  depth &= 0xffff;                       // see accum_depth_of
  if (depth < 1) depth = 1;
  const long long per_level = (static_cast<long long>(Lq) * P + depth - 1) / depth;
  return static_cast<long long>(L) * (per_level + 1) + 2ll * S;
}

// Hybrid backward (dense call site): `depth_and_split` carries, above the 16 bits of the bucket depth, a threshold on the
// expected adds per element.  Levels above it (coarse levels: hundreds of corner rows per pixel) get their grad_value from
// msda_bwd_scatter_tiled_kernel, which pre-reduces inside the SM and therefore needs a single copy; levels at or below it
// (fine levels: little to pre-reduce) keep the direct reductions of msda_bwd_vec_kernel and their buckets.  0 = no split.
__host__ __device__ inline int accum_depth_of(int depth_and_split) { return depth_and_split & 0xffff; }
__host__ __device__ inline int accum_split_of(int depth_and_split) { return (depth_and_split >> 16) & 0x7fff; }

// call after load_level_meta (thread 0 fills, then a barrier)
__device__ __forceinline__ void build_accum_layout(LevelMeta& meta, int L, int Lq, int P, int depth_and_split, bool sparse_direct) {
  if (threadIdx.x == 0) {
    const int depth = accum_depth_of(depth_and_split), split = accum_split_of(depth_and_split);
    int base = 0, dir = 0, bkt = 0;
    uint32_t sorted = 0u;
    for (int l = 0; l < L; ++l) {
      const long long hw = static_cast<long long>(meta.H[l]) * meta.W[l];
      const long long adds = hw > 0 ? (static_cast<long long>(Lq) * P + hw - 1) / hw : 1;
      int K = static_cast<int>((adds + depth - 1) / depth);
      if (split > 0 && adds > split && l < 32) { K = 1; sorted |= 1u << l; }
      const bool direct = sparse_direct && hw > 0 && kSparseFactor * Lq * P <= hw;
      meta.accK[l] = direct ? 0 : (K < 1 ? 1 : K);
      meta.accBase[l] = base;
      meta.dirOff[l] = dir;
      meta.bktOff[l] = bkt;
      base += meta.accK[l] * static_cast<int>(hw);
      if (direct) dir += static_cast<int>(hw); else bkt += static_cast<int>(hw);
    }
    meta.accStride = base;
    meta.dirRows = dir;
    meta.bktRows = bkt;
    meta.sortMask = sorted;
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

// ---------------------------------------------------------------------------------------------------
// Fast fp32 footprint used by the vector kernels.
//
// Same arithmetic as upstream up to the corner fetch: w_im = x*W - 0.5 (mul, then sub), the (-1, W) gate,
// lw = w_im - floor(w_im).  floor() is taken with one round-down add of 1.5*2^23 (FADD.RM, full rate)
// instead of F2I/FRND (measured 0.5 warp-instr/clk/SM on B200, profiles/microbench_issue_r01.jsonl): for
// |t| < 2^22 the sum lies in [2^23, 2^24) where ulp = 1, so rounding down yields floor(t) + 1.5*2^23 exactly;
// the float floor is recovered by an exact subtraction and the integer from the mantissa bits.
//
// Zero padding is folded into the weights: every corner address is clamped into the level (so all four
// loads are unconditional) and an out-of-range corner gets weight 0.  The clamped address of an invalid
// corner is always another, valid corner of the same footprint, so non-finite values propagate exactly as
// upstream.  A point that fails the gate altogether must not touch `value` (upstream skips it, and 0 * NaN would
// leak a non-finite pixel into the result): the kernels point such a point at kZeroBlock with a zero pixel stride
// (gated_base / gated_stride below), so its four loads stay unconditional and read zeros.
// ---------------------------------------------------------------------------------------------------
// 16 zero bytes: what a sampling point that fails upstream's gate reads instead of `value` (see make_taps)
static __device__ uint4 kZeroBlock = {0u, 0u, 0u, 0u};

struct Taps {
  bool inside;                     // upstream's point-level gate
  int i00, i01, i10, i11;          // pixel indices inside the level, always in range
  float hhm, lhm, hwm, lwm;        // hh/lh/hw/lw with the row / column validity (and the gate) folded in
  float Tm, Bm, Lm, Rm;            // 1.0f / 0.0f validity of top / bottom row and left / right column
};

__device__ __forceinline__ void floor_split(float t, float& frac, int& i) {
  const float r = __fadd_rd(t, 12582912.0f);
  i = __float_as_int(r) - 0x4B400000;
  frac = __fsub_rn(t, __fsub_rn(r, 12582912.0f));
}

__device__ __forceinline__ Taps make_taps(float x, float y, int H, int W, float Hf, float Wf) {
  Taps t;
  const float h_im = __fsub_rn(__fmul_rn(y, Hf), 0.5f);
  const float w_im = __fsub_rn(__fmul_rn(x, Wf), 0.5f);
  const bool inside = (h_im > -1.f) && (w_im > -1.f) && (h_im < Hf) && (w_im < Wf);
  t.inside = inside;
  float lh, lw; int iy, ix;
  floor_split(h_im, lh, iy);
  floor_split(w_im, lw, ix);
  const bool top = inside && iy >= 0, bot = inside && iy < H - 1;
  const bool left = inside && ix >= 0, right = inside && ix < W - 1;
  t.Tm = top ? 1.f : 0.f; t.Bm = bot ? 1.f : 0.f; t.Lm = left ? 1.f : 0.f; t.Rm = right ? 1.f : 0.f;
  t.hhm = top ? 1.f - lh : 0.f; t.lhm = bot ? lh : 0.f;
  t.hwm = left ? 1.f - lw : 0.f; t.lwm = right ? lw : 0.f;
  const int yt = min(max(iy, 0), H - 1), yb = min(max(iy + 1, 0), H - 1);
  const int xl = min(max(ix, 0), W - 1), xr = min(max(ix + 1, 0), W - 1);
  t.i00 = yt * W + xl; t.i01 = yt * W + xr; t.i10 = yb * W + xl; t.i11 = yb * W + xr;
  return t;
}

__device__ __forceinline__ const char* gated_base(const Taps& t, const char* level_base) {
  return t.inside ? level_base : reinterpret_cast<const char*>(&kZeroBlock);
}
__device__ __forceinline__ uint32_t gated_stride(const Taps& t, uint32_t pix_bytes) { return t.inside ? pix_bytes : 0u; }

// ---- acc[i] += w * v[i] over the VEC elements of one 16-byte vector -----------------------------------
// fp32 values: plain FFMA with the fp32 weight.  16-bit values: Blackwell's mixed-precision FMA
// (PTX fma.rn.f32.bf16 / .f16 -> SASS FHFMA): the product of two 16-bit operands is exact in fp32 and the
// accumulation is fp32, one instruction per element and no unpacking.  The operand `w` must then already be
// a 16-bit weight (bits in the low half).
template <typename T> struct WeightT { using type = float; };
template <> struct WeightT<__nv_bfloat16> { using type = uint32_t; };
template <> struct WeightT<__half> { using type = uint32_t; };

template <typename T> __device__ __forceinline__ typename WeightT<T>::type make_weight(float w);
template <> __device__ __forceinline__ float make_weight<float>(float w) { return w; }
template <> __device__ __forceinline__ uint32_t make_weight<__nv_bfloat16>(float w) {
  return static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(w)));
}
template <> __device__ __forceinline__ uint32_t make_weight<__half>(float w) {
  return static_cast<uint32_t>(__half_as_ushort(__float2half_rn(w)));
}

__device__ __forceinline__ void fma_mixed_bf16(float& acc, uint32_t v, bool hi, uint32_t w) {
  unsigned short lo16, hi16, w16;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(lo16), "=h"(hi16) : "r"(v));
  asm("cvt.u16.u32 %0, %1;" : "=h"(w16) : "r"(w));
  if (hi) asm("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc) : "h"(hi16), "h"(w16));
  else    asm("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc) : "h"(lo16), "h"(w16));
}
__device__ __forceinline__ void fma_mixed_f16(float& acc, uint32_t v, bool hi, uint32_t w) {
  unsigned short lo16, hi16, w16;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(lo16), "=h"(hi16) : "r"(v));
  asm("cvt.u16.u32 %0, %1;" : "=h"(w16) : "r"(w));
  if (hi) asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc) : "h"(hi16), "h"(w16));
  else    asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc) : "h"(lo16), "h"(w16));
}

template <typename T> __device__ __forceinline__ void axpy16(float* acc, const uint4& u, typename WeightT<T>::type w);
template <> __device__ __forceinline__ void axpy16<float>(float* acc, const uint4& u, float w) {
  acc[0] = fmaf(w, __uint_as_float(u.x), acc[0]); acc[1] = fmaf(w, __uint_as_float(u.y), acc[1]);
  acc[2] = fmaf(w, __uint_as_float(u.z), acc[2]); acc[3] = fmaf(w, __uint_as_float(u.w), acc[3]);
}
template <> __device__ __forceinline__ void axpy16<__nv_bfloat16>(float* acc, const uint4& u, uint32_t w) {
  fma_mixed_bf16(acc[0], u.x, false, w); fma_mixed_bf16(acc[1], u.x, true, w);
  fma_mixed_bf16(acc[2], u.y, false, w); fma_mixed_bf16(acc[3], u.y, true, w);
  fma_mixed_bf16(acc[4], u.z, false, w); fma_mixed_bf16(acc[5], u.z, true, w);
  fma_mixed_bf16(acc[6], u.w, false, w); fma_mixed_bf16(acc[7], u.w, true, w);
}
template <> __device__ __forceinline__ void axpy16<__half>(float* acc, const uint4& u, uint32_t w) {
  fma_mixed_f16(acc[0], u.x, false, w); fma_mixed_f16(acc[1], u.x, true, w);
  fma_mixed_f16(acc[2], u.y, false, w); fma_mixed_f16(acc[3], u.y, true, w);
  fma_mixed_f16(acc[4], u.z, false, w); fma_mixed_f16(acc[5], u.z, true, w);
  fma_mixed_f16(acc[6], u.w, false, w); fma_mixed_f16(acc[7], u.w, true, w);
}

// same, with the 16-bit weight taken from the low (hi = false) or high half of a packed weight pair
template <typename T> __device__ __forceinline__ void axpy16_packed(float* acc, const uint4& u, uint32_t wpair, bool hi);
__device__ __forceinline__ void fma_mixed_bf16_sel(float& acc, uint32_t v, bool vhi, uint32_t wpair, bool whi) {
  unsigned short v0, v1, w0, w1;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(v0), "=h"(v1) : "r"(v));
  asm("mov.b32 {%0, %1}, %2;" : "=h"(w0), "=h"(w1) : "r"(wpair));
  asm("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc) : "h"(vhi ? v1 : v0), "h"(whi ? w1 : w0));
}
__device__ __forceinline__ void fma_mixed_f16_sel(float& acc, uint32_t v, bool vhi, uint32_t wpair, bool whi) {
  unsigned short v0, v1, w0, w1;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(v0), "=h"(v1) : "r"(v));
  asm("mov.b32 {%0, %1}, %2;" : "=h"(w0), "=h"(w1) : "r"(wpair));
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc) : "h"(vhi ? v1 : v0), "h"(whi ? w1 : w0));
}
template <> __device__ __forceinline__ void axpy16_packed<__nv_bfloat16>(float* acc, const uint4& u, uint32_t wp, bool hi) {
  fma_mixed_bf16_sel(acc[0], u.x, false, wp, hi); fma_mixed_bf16_sel(acc[1], u.x, true, wp, hi);
  fma_mixed_bf16_sel(acc[2], u.y, false, wp, hi); fma_mixed_bf16_sel(acc[3], u.y, true, wp, hi);
  fma_mixed_bf16_sel(acc[4], u.z, false, wp, hi); fma_mixed_bf16_sel(acc[5], u.z, true, wp, hi);
  fma_mixed_bf16_sel(acc[6], u.w, false, wp, hi); fma_mixed_bf16_sel(acc[7], u.w, true, wp, hi);
}
template <> __device__ __forceinline__ void axpy16_packed<__half>(float* acc, const uint4& u, uint32_t wp, bool hi) {
  fma_mixed_f16_sel(acc[0], u.x, false, wp, hi); fma_mixed_f16_sel(acc[1], u.x, true, wp, hi);
  fma_mixed_f16_sel(acc[2], u.y, false, wp, hi); fma_mixed_f16_sel(acc[3], u.y, true, wp, hi);
  fma_mixed_f16_sel(acc[4], u.z, false, wp, hi); fma_mixed_f16_sel(acc[5], u.z, true, wp, hi);
  fma_mixed_f16_sel(acc[6], u.w, false, wp, hi); fma_mixed_f16_sel(acc[7], u.w, true, wp, hi);
}
template <typename T> __device__ __forceinline__ uint32_t pack_weight_pair(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack_weight_pair<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
template <> __device__ __forceinline__ uint32_t pack_weight_pair<__half>(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// ---- dot(v, g) over one 16-byte vector, fp32 accumulate (16-bit: exact products through FHFMA) ----------
template <typename T> __device__ __forceinline__ float dot16(const uint4& v, const uint4& g, float acc);
template <> __device__ __forceinline__ float dot16<float>(const uint4& v, const uint4& g, float acc) {
  acc = fmaf(__uint_as_float(v.x), __uint_as_float(g.x), acc); acc = fmaf(__uint_as_float(v.y), __uint_as_float(g.y), acc);
  acc = fmaf(__uint_as_float(v.z), __uint_as_float(g.z), acc); acc = fmaf(__uint_as_float(v.w), __uint_as_float(g.w), acc);
  return acc;
}
__device__ __forceinline__ void dot_mixed_bf16(float& acc, uint32_t v, uint32_t g) {
  unsigned short v0, v1, g0, g1;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(v0), "=h"(v1) : "r"(v));
  asm("mov.b32 {%0, %1}, %2;" : "=h"(g0), "=h"(g1) : "r"(g));
  asm("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc) : "h"(v0), "h"(g0));
  asm("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc) : "h"(v1), "h"(g1));
}
__device__ __forceinline__ void dot_mixed_f16(float& acc, uint32_t v, uint32_t g) {
  unsigned short v0, v1, g0, g1;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(v0), "=h"(v1) : "r"(v));
  asm("mov.b32 {%0, %1}, %2;" : "=h"(g0), "=h"(g1) : "r"(g));
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc) : "h"(v0), "h"(g0));
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc) : "h"(v1), "h"(g1));
}
template <> __device__ __forceinline__ float dot16<__nv_bfloat16>(const uint4& v, const uint4& g, float acc) {
  float a0 = acc, a1 = 0.f;    // two chains for ILP
  dot_mixed_bf16(a0, v.x, g.x); dot_mixed_bf16(a1, v.y, g.y); dot_mixed_bf16(a0, v.z, g.z); dot_mixed_bf16(a1, v.w, g.w);
  return a0 + a1;
}
template <> __device__ __forceinline__ float dot16<__half>(const uint4& v, const uint4& g, float acc) {
  float a0 = acc, a1 = 0.f;
  dot_mixed_f16(a0, v.x, g.x); dot_mixed_f16(a1, v.y, g.y); dot_mixed_f16(a0, v.z, g.z); dot_mixed_f16(a1, v.w, g.w);
  return a0 + a1;
}

}  // namespace msda
