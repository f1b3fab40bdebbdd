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
#include <cstdlib>
#include <atomic>
#include <mutex>
#include <type_traits>
#include <vector>

#include "msda_common.cuh"
#include "../../include/msda_b200.h"

namespace msda {

#ifndef BWD_MIN_CTAS
#define BWD_MIN_CTAS 3
#endif

// =====================================================================================================
// Shared-memory staging of a warp's sampling locations / attention weights (and, in backward, of the
// gradients that replace them in place).  The warp's pairs are contiguous in global memory, so the copies
// are plain coalesced 128-bit transfers; each pair's row gets 16 bytes of padding so that the G-lane groups
// of a warp read their rows from distinct banks.
// =====================================================================================================
// Row strides of the staged rows, in floats: the payload rounded up to a multiple of 4 plus 4 floats of padding, so that
// every row (and every warp's block of rows) starts 16-byte aligned for any L*P — float2 / float4 reads of a row are
// then always legal — and the G-lane groups of a warp still read their rows from distinct banks.
__host__ __device__ inline int loc_row_stride(int LP) { return ((2 * LP + 3) & ~3) + 4; }
__host__ __device__ inline int attn_row_stride(int LP) { return ((LP + 3) & ~3) + 4; }

struct WarpStage {
  float* loc;        // [GPW][2*LP + 4]
  float* attn;       // [GPW][LP + 4]
  float* gattn;      // [GPW][LP + 4]  fused backward only: grad wrt the softmax output (attn itself is kept)
  int loc_stride, attn_stride;
};

// Rows are moved in quads of 4 elements; AT is float, or the 16-bit value type when the fused kernels consume the
// Linear outputs of an autocast region directly (8-byte loads / stores, converted to / from the fp32 staging rows).
template <typename AT> __device__ __forceinline__ float4 load_quad(const AT* __restrict__ base, int i);
template <> __device__ __forceinline__ float4 load_quad<float>(const float* __restrict__ base, int i) {
  return __ldg(reinterpret_cast<const float4*>(base) + i);
}
template <> __device__ __forceinline__ float4 load_quad<__nv_bfloat16>(const __nv_bfloat16* __restrict__ base, int i) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(base) + i);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}
template <> __device__ __forceinline__ float4 load_quad<__half>(const __half* __restrict__ base, int i) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(base) + i);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename AT> __device__ __forceinline__ void store_quad(AT* __restrict__ base, int i, const float4& v);
template <> __device__ __forceinline__ void store_quad<float>(float* __restrict__ base, int i, const float4& v) {
  __stcs(reinterpret_cast<float4*>(base) + i, v);
}
template <> __device__ __forceinline__ void store_quad<__nv_bfloat16>(__nv_bfloat16* __restrict__ base, int i, const float4& v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  __stcs(reinterpret_cast<uint2*>(base) + i, make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b)));
}
template <> __device__ __forceinline__ void store_quad<__half>(__half* __restrict__ base, int i, const float4& v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  __stcs(reinterpret_cast<uint2*>(base) + i, make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b)));
}

template <typename AT>
__device__ __forceinline__ void stage_in(const WarpStage& ws, const AT* __restrict__ gl, const AT* __restrict__ ga,
                                         int nvalid, int LP, int lane) {
  if ((LP & 3) == 0) {
    const int lv = LP / 2, av = LP / 4;     // quads per pair row
    for (int i = lane; i < nvalid * lv; i += 32) {
      const int r = i / lv, k = i - r * lv;
      *reinterpret_cast<float4*>(ws.loc + r * ws.loc_stride + 4 * k) = load_quad<AT>(gl, i);
    }
    for (int i = lane; i < nvalid * av; i += 32) {
      const int r = i / av, k = i - r * av;
      *reinterpret_cast<float4*>(ws.attn + r * ws.attn_stride + 4 * k) = load_quad<AT>(ga, i);
    }
  } else {
    for (int i = lane; i < nvalid * 2 * LP; i += 32) {
      const int r = i / (2 * LP), k = i - r * 2 * LP;
      ws.loc[r * ws.loc_stride + k] = to_acc<AT>(gl[i]);
    }
    for (int i = lane; i < nvalid * LP; i += 32) {
      const int r = i / LP, k = i - r * LP;
      ws.attn[r * ws.attn_stride + k] = to_acc<AT>(ga[i]);
    }
  }
}

template <typename AT>
__device__ __forceinline__ void stage_out(const WarpStage& ws, const float* __restrict__ attn_src,
                                          AT* __restrict__ gl, AT* __restrict__ ga,
                                          int nvalid, int LP, int lane) {
  if ((LP & 3) == 0) {
    const int lv = LP / 2, av = LP / 4;
    for (int i = lane; i < nvalid * lv; i += 32) {
      const int r = i / lv, k = i - r * lv;
      store_quad<AT>(gl, i, *reinterpret_cast<const float4*>(ws.loc + r * ws.loc_stride + 4 * k));
    }
    for (int i = lane; i < nvalid * av; i += 32) {
      const int r = i / av, k = i - r * av;
      store_quad<AT>(ga, i, *reinterpret_cast<const float4*>(attn_src + r * ws.attn_stride + 4 * k));
    }
  } else {
    for (int i = lane; i < nvalid * 2 * LP; i += 32) {
      const int r = i / (2 * LP), k = i - r * 2 * LP;
      gl[i] = from_acc<AT>(ws.loc[r * ws.loc_stride + k]);
    }
    for (int i = lane; i < nvalid * LP; i += 32) {
      const int r = i / LP, k = i - r * LP;
      ga[i] = from_acc<AT>(attn_src[r * ws.attn_stride + k]);
    }
  }
}

// =====================================================================================================
// Fused pre-op.  In fused mode the kernels consume what MSDeformAttn.forward holds
// *before* it materialises sampling_locations / softmaxed attention_weights: the raw outputs of the
// sampling_offsets and attention_weights Linears plus reference_points (N, Lq, L, R), R = 2 (encoder: points)
// or 4 (decoder: boxes).  After the warp's rows are staged, each lane group rewrites its own row in place,
// with the module's arithmetic (see stage_ref for the rounding):
//     R == 2:  loc = ref + off / (W_l, H_l)
//     R == 4:  loc = ref_xy + ((off / P) * ref_wh) * 0.5
//     attn = softmax over the L*P logits of the pair = exp(x - max) / sum     (ex2.approx-based exp, one reciprocal)
// The backward applies the matching chain rules to grad_loc / grad_attn before they leave shared memory:
//     R == 2:  grad_off = grad_loc / (W_l, H_l);   R == 4:  grad_off = ((grad_loc * 0.5) * ref_wh) / P
//     grad_logit = (grad_attn - sum_j(grad_attn_j * attn_j)) * attn
// so sampling_locations / attention_weights and their gradients never exist in HBM.
// =====================================================================================================
// Per (pair, level) affine map of the raw offsets: loc = (rx, ry) + off * (sx, sy), staged as float4 (rx, ry, sx, sy).
//   R == 2: (sx, sy) = (1 / W_l, 1 / H_l);   R == 4: (sx, sy) = ref_wh * (0.5 / P).
// For power-of-two W_l, H_l and P (every BASELINE config) this is bit-identical to the module's
// `ref + off / (W, H)` / `ref_xy + off / P * ref_wh * 0.5`; otherwise it differs by at most one ulp of the offset term.
// The loads are issued together with the staging loads of the warp's rows, so their latency overlaps.
__device__ __forceinline__ void stage_ref(float4* __restrict__ sref, const LevelMeta& meta, const float* __restrict__ ref,
                                          int R, int pair0, int nvalid, int M, int L, int P, int lane) {
  const float half_over_p = 0.5f / static_cast<float>(P);
  for (int i = lane; i < nvalid * L; i += 32) {
    const int r = i / L, l = i - r * L;
    const float* row = ref + (static_cast<size_t>((pair0 + r) / M) * L + l) * R;
    float4 v;
    if (R == 2) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(row));
      v = make_float4(t.x, t.y, 1.0f / static_cast<float>(meta.W[l]), 1.0f / static_cast<float>(meta.H[l]));
    } else {
      const float4 t = __ldg(reinterpret_cast<const float4*>(row));
      v = make_float4(t.x, t.y, __fmul_rn(t.z, half_over_p), __fmul_rn(t.w, half_over_p));
    }
    sref[r * L + l] = v;
  }
}

// level of point i (i / P) without an integer division: exact for i, P < 2^20
__device__ __forceinline__ int level_of(int i, float inv_p) {
  return __float2int_rz(__fmul_rn(static_cast<float>(i) + 0.5f, inv_p));
}

// Executed by ALL lanes of the warp with full-mask shuffles (per-group masks would leave the warp split into
// independently scheduled groups for the main loop); padding groups of a ragged last warp shadow row 0 and do not write.
template <int G>
__device__ __forceinline__ void fused_prepare(float* myloc, float* myattn, const float4* __restrict__ myref,
                                              int L, int P, int c, bool active) {
  const int LP = L * P;
  const float inv_p = 1.0f / static_cast<float>(P);
  for (int i = c; i < LP; i += G) {
    const float4 rs = myref[level_of(i, inv_p)];
    const float2 off = *reinterpret_cast<const float2*>(myloc + 2 * i);
    if (active)
      *reinterpret_cast<float2*>(myloc + 2 * i) =
          make_float2(__fadd_rn(rs.x, __fmul_rn(off.x, rs.z)), __fadd_rn(rs.y, __fmul_rn(off.y, rs.w)));
  }
  // softmax: the first 4 logits of the lane (all of them when L*P <= 4*G, e.g. 16 points on 4 lanes) stay in
  // registers between the passes; longer rows spill over to shared memory (the shared-memory pipe is what bounds
  // the gather loop that follows, so every LDS/STS saved here is time saved there)
  float lg[4];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = c + k * G;
    lg[k] = i < LP ? myattn[i] : -INFINITY;
    mx = fmaxf(mx, lg[k]);
  }
  for (int i = c + 4 * G; i < LP; i += G) mx = fmaxf(mx, myattn[i]);
#pragma unroll
  for (int s = G / 2; s >= 1; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    lg[k] = __expf(lg[k] - mx);           // exp(-inf) = 0 for the padding slots
    sum += lg[k];
  }
  for (int i = c + 4 * G; i < LP; i += G) sum += __expf(myattn[i] - mx);
#pragma unroll
  for (int s = G / 2; s >= 1; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
  const float inv = 1.0f / sum;
  if (active) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = c + k * G;
      if (i < LP) myattn[i] = lg[k] * inv;
    }
    for (int i = c + 4 * G; i < LP; i += G) myattn[i] = __expf(myattn[i] - mx) * inv;
  }
  __syncwarp();
}

// =====================================================================================================
// Forward, vector path
// =====================================================================================================
// FUSED: `loc` / `attn` hold raw sampling offsets / attention logits and `ref` (N, Lq, L, R) the reference points
// AT   : element type of `loc` / `attn` (float; the fused kernels also take the 16-bit value type)
// Plain forward: 6 CTAs/SM (40 registers, 8 bytes of spill) instead of the 5 the compiler's 43 registers allow:
// the kernel is bound by issue slots and L1TEX wavefronts, one more resident CTA hides more of the gather latency
// (A/B on one box: 0.801 -> 0.780 ms at cfg3).  7 CTAs/SM spill ~100 bytes; the fused variant keeps its registers.
template <typename T, int D, bool FUSED, typename AT>
__global__ void __launch_bounds__(kThreads, FUSED ? 5 : 6)
msda_fwd_vec_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lsi, const AT* __restrict__ loc,
                    const AT* __restrict__ attn, const float* __restrict__ ref, int R, T* __restrict__ out,
                    int S, int M, int Lq, int L, int P, int total_pairs, int vps) {
  // vps: elements between neighbouring pixels of `value` (M*D when dense; larger when the projections of several
  // layers are interleaved per pixel, see msda_forward_strided)
  constexpr int VEC = 16 / sizeof(T);
  constexpr int G = D / VEC;        // lanes per (b, q, m) pair
  constexpr int GPW = 32 / G;       // pairs per warp
  static_assert(G >= 1 && G <= 32 && (G & (G - 1)) == 0, "lane group must be a power of two");

  extern __shared__ __align__(16) float smem[];
  __shared__ LevelMeta meta;
  load_level_meta(meta, shapes, lsi, L);

  const int LP = L * P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / G, c = lane % G;
  WarpStage ws;
  ws.loc_stride = loc_row_stride(LP);
  ws.attn_stride = attn_row_stride(LP);
  ws.loc = smem + warp * GPW * (ws.loc_stride + ws.attn_stride);
  ws.attn = ws.loc + GPW * ws.loc_stride;
  ws.gattn = nullptr;

  const int pair0 = (blockIdx.x * kWarps + warp) * GPW;
  const int nvalid = min(GPW, total_pairs - pair0);
  if (nvalid <= 0) return;
  stage_in<AT>(ws, loc + static_cast<size_t>(pair0) * (2 * LP), attn + static_cast<size_t>(pair0) * LP, nvalid, LP, lane);
  float4* sref = nullptr;
  if constexpr (FUSED) {
    // float4 rows behind the CTA's loc / attn rows (16-byte aligned: every row length is a multiple of 4 floats)
    sref = reinterpret_cast<float4*>(smem + kWarps * GPW * (ws.loc_stride + ws.attn_stride)) + warp * GPW * L;
    stage_ref(sref, meta, ref, R, pair0, nvalid, M, L, P, lane);
  }
  __syncwarp();
  if constexpr (FUSED) {
    const bool active = g < nvalid;
    const int row = active ? g : 0;
    fused_prepare<G>(ws.loc + row * ws.loc_stride, ws.attn + row * ws.attn_stride, sref + row * L, L, P, c, active);
  }
  if (g >= nvalid) return;

  const int pair = pair0 + g;
  const int m = pair % M;
  const int b = (pair / M) / Lq;
  const uint32_t pix_bytes = static_cast<uint32_t>(vps) * sizeof(T);          // bytes between neighbouring pixels
  const char* vb = reinterpret_cast<const char*>(value) + (static_cast<size_t>(b) * S * vps + m * D) * sizeof(T) + c * 16;
  const float* myloc = ws.loc + g * ws.loc_stride;
  const float* myattn = ws.attn + g * ws.attn_stride;

  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

  // one sampling point: four unconditional 16-byte corner loads + the weighted accumulation
  auto sample = [&](float x, float y, float a, int H, int W, float Hf, float Wf, const char* vl) {
    const Taps t = make_taps(x, y, H, W, Hf, Wf);
    // pixel index * pixel stride fits 32 bits (validated on the host): one IMAD.WIDE per corner; a point that fails
    // the gate reads the zero block instead of `value`
    const char* vg = gated_base(t, vl);
    const uint32_t pg = gated_stride(t, pix_bytes);
    const uint4 u00 = ldg16(vg + static_cast<size_t>(static_cast<uint32_t>(t.i00) * pg));
    const uint4 u01 = ldg16(vg + static_cast<size_t>(static_cast<uint32_t>(t.i01) * pg));
    const uint4 u10 = ldg16(vg + static_cast<size_t>(static_cast<uint32_t>(t.i10) * pg));
    const uint4 u11 = ldg16(vg + static_cast<size_t>(static_cast<uint32_t>(t.i11) * pg));
    const float ah = t.hhm * a, al = t.lhm * a;
    axpy16<T>(acc, u00, make_weight<T>(ah * t.hwm));
    axpy16<T>(acc, u01, make_weight<T>(ah * t.lwm));
    axpy16<T>(acc, u10, make_weight<T>(al * t.hwm));
    axpy16<T>(acc, u11, make_weight<T>(al * t.lwm));
  };
  for (int l = 0; l < L; ++l) {
    const int H = meta.H[l], W = meta.W[l];
    const float Hf = static_cast<float>(H), Wf = static_cast<float>(W);
    const char* vl = vb + static_cast<size_t>(meta.start[l]) * pix_bytes;
    if ((P & 1) == 0) {                    // rows are 16-byte aligned (loc_row_stride); l*P + p is even here
      // the shared-memory pipe is the one the gathers saturate: fetch two points' (x, y) / weights per LDS
      for (int p = 0; p < P; p += 2) {
        const float4 xy2 = *reinterpret_cast<const float4*>(myloc + 2 * (l * P + p));
        const float2 a2 = *reinterpret_cast<const float2*>(myattn + l * P + p);
        sample(xy2.x, xy2.y, a2.x, H, W, Hf, Wf, vl);
        sample(xy2.z, xy2.w, a2.y, H, W, Hf, Wf, vl);
      }
      continue;
    }
#pragma unroll 2
    for (int p = 0; p < P; ++p) {
      const float2 xy = *reinterpret_cast<const float2*>(myloc + 2 * (l * P + p));
      const float a = myattn[l * P + p];
      const Taps t = make_taps(xy.x, xy.y, H, W, Hf, Wf);
      const char* vg = gated_base(t, vl);
      const uint32_t pg = gated_stride(t, pix_bytes);
      const uint4 u00 = ldg16(vg + static_cast<size_t>(static_cast<uint32_t>(t.i00) * pg));
      const uint4 u01 = ldg16(vg + static_cast<size_t>(static_cast<uint32_t>(t.i01) * pg));
      const uint4 u10 = ldg16(vg + static_cast<size_t>(static_cast<uint32_t>(t.i10) * pg));
      const uint4 u11 = ldg16(vg + static_cast<size_t>(static_cast<uint32_t>(t.i11) * pg));
      const float ah = t.hhm * a, al = t.lhm * a;
      axpy16<T>(acc, u00, make_weight<T>(ah * t.hwm));
      axpy16<T>(acc, u01, make_weight<T>(ah * t.lwm));
      axpy16<T>(acc, u10, make_weight<T>(al * t.hwm));
      axpy16<T>(acc, u11, make_weight<T>(al * t.lwm));
    }
  }
  __stcs(reinterpret_cast<uint4*>(out + static_cast<size_t>(pair) * D + c * VEC), pack16<T>(acc));
}

// =====================================================================================================
// Backward, vector path
// =====================================================================================================
// Cluster guard (see msda_cluster_density_kernel): the launcher enqueues the 16-bit and the fp32 accumulation pipeline
// back to back; every kernel of a pipeline returns at once unless the flag in the control block selects it.
__device__ __forceinline__ bool gated_off(const uint32_t* __restrict__ gate, int want) {
  return gate != nullptr && ((__ldg(gate) != 0u) != (want != 0));
}

// Reduce-scatter of per-lane partial sums inside a G-lane group: every lane enters with R values for each of
// G consecutive sampling points and leaves with the group totals of the point whose index equals its lane
// position c.  (R*(G-1) shuffles per G points instead of R*log2(G) per point.)
template <int G, int R>
__device__ __forceinline__ void group_reduce_scatter(float (&v)[G][R], int c) {
#pragma unroll
  for (int s = G / 2; s >= 1; s >>= 1) {
    const bool upper = (c & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float keep = upper ? v[j + s][r] : v[j][r];
        const float send = upper ? v[j][r] : v[j + s][r];
        v[j][r] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
  }
}

// Scale of the fp16 accumulation buffer (GV16 mode).  |grad_value[b, s, m, d]| <= sum_q |grad_out[b, q, m, d]| *
// sum_p attn*bilinear <= Lq * max|grad_out| for ANY input (attention weights of a query sum to <= 1 in the
// module; bilinear weights to <= 1), so with scale = the largest power of two such that
// Lq * max|grad_out| * scale <= 2^15 no partial sum can overflow fp16 (max 65504), whatever the sampling
// pattern.  ctrl[0] holds the bits of max|grad_out| (bits of a non-negative float order like unsigned integers: atomicMax
// in msda_zero_f16_buckets_kernel, or in the tiled backward's dots kernel).
__device__ __forceinline__ float f16_accum_scale(const uint32_t* __restrict__ ctrl, int Lq) {
  const float bound = __uint_as_float(__ldg(ctrl)) * static_cast<float>(Lq);
  if (!(bound > 0.f)) return 1.f;
  int e;
  frexpf(bound, &e);                         // bound = f * 2^e, f in [0.5, 1)  =>  bound <= 2^e
  return ldexpf(1.f, max(-100, min(100, 15 - e)));
}

// GV16 = false: grad_value contributions go to an fp32 buffer `gv32` laid out like value
//               (grad_value itself for T = float, the caller's scratch for 16-bit T) with red.v4.f32.
// GV16 = true : 16-bit T only; contributions are scaled (f16_accum_scale), rounded to fp16 and added with
//               packed red.global.add.noftz.v4.f16x2 into an fp16 buffer `gv16` laid out like value: half the
//               reduction bytes of the fp32 path (the SM->L2 reduction path is what bounds this kernel).
// FUSED   : `loc` / `attn` hold raw sampling offsets / attention logits, `ref` (N, Lq, L, R) the reference points;
//           `grad_loc` / `grad_attn` receive the gradients of the raw offsets / logits (see fused_prepare).
// AT      : element type of `loc` / `attn` / `grad_loc` / `grad_attn` (float; fused kernels also the 16-bit value type)
// SPARSE  : GV16 only; some level may be sparse (4*Lq*P <= H_l*W_l <= S, decided on the host from Lq, P, S) and then
//           adds straight into grad_value (build_accum_layout).  A separate instantiation because the extra
//           level-uniform branch and addressing in the reduction loop cost the dense encoder shapes ~8 %.
// HYB     : GV16 only; hybrid backward of the dense call site: levels in meta.sortMask (see build_accum_layout) get their
//           grad_value from msda_bwd_scatter_tiled_kernel, this kernel skips their reductions (and does everything else)
template <typename T, int D, bool GV16, bool FUSED, typename AT, bool SPARSE, bool HYB = false>
__global__ void __launch_bounds__(kThreads, BWD_MIN_CTAS)
msda_bwd_vec_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lsi, const AT* __restrict__ loc,
                    const AT* __restrict__ attn, const float* __restrict__ ref, int R,
                    const T* __restrict__ grad_out,
                    float* __restrict__ gv32, __half* __restrict__ gv16, T* __restrict__ gv_direct,
                    const uint32_t* __restrict__ ctrl,
                    AT* __restrict__ grad_loc, AT* __restrict__ grad_attn,
                    int S, int M, int Lq, int L, int P, int total_pairs, int depth, int vps, int gps,
                    const uint32_t* __restrict__ gate, int gate_want) {
  if (gated_off(gate, gate_want)) return;   // cluster guard: the other accumulation mode handles this call
  // vps: elements between neighbouring pixels of `value`; gps: the same for the buffer laid out like value that receives
  // reductions directly (gv32, or gv_direct for the sparse levels of GV16 mode: see build_accum_layout)
  constexpr int VEC = 16 / sizeof(T);
  constexpr int G = D / VEC;
  constexpr int GPW = 32 / G;
  constexpr bool k16 = sizeof(T) == 2;
  float gv_scale = 1.f;
  if constexpr (GV16) gv_scale = f16_accum_scale(ctrl, Lq);

  extern __shared__ __align__(16) float smem[];
  __shared__ LevelMeta meta;
  load_level_meta(meta, shapes, lsi, L);
  if constexpr (GV16) build_accum_layout(meta, L, Lq, P, depth, SPARSE);

  const int LP = L * P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / G, c = lane % G;
  WarpStage ws;
  ws.loc_stride = loc_row_stride(LP);
  ws.attn_stride = attn_row_stride(LP);
  ws.loc = smem + warp * GPW * (ws.loc_stride + (FUSED ? 2 : 1) * ws.attn_stride);
  ws.attn = ws.loc + GPW * ws.loc_stride;
  ws.gattn = FUSED ? ws.attn + GPW * ws.attn_stride : nullptr;

  const int pair0 = (blockIdx.x * kWarps + warp) * GPW;
  const int nvalid = min(GPW, total_pairs - pair0);
  if (nvalid <= 0) return;
  stage_in<AT>(ws, loc + static_cast<size_t>(pair0) * (2 * LP), attn + static_cast<size_t>(pair0) * LP, nvalid, LP, lane);
  float4* sref = nullptr;
  if constexpr (FUSED) {
    sref = reinterpret_cast<float4*>(smem + kWarps * GPW * (ws.loc_stride + 2 * ws.attn_stride)) + warp * GPW * L;
    stage_ref(sref, meta, ref, R, pair0, nvalid, M, L, P, lane);
  }
  __syncwarp();

  // Lanes of padding groups (tail warp only) stay in the loop so that the full-mask shuffles are legal;
  // they shadow pair 0 of the warp and never write.
  const bool active = g < nvalid;
  const int pair = pair0 + (active ? g : 0);
  const int m = pair % M;
  const int b = (pair / M) / Lq;
  const uint32_t pix_bytes = static_cast<uint32_t>(vps) * sizeof(T);
  const size_t img_pix = static_cast<size_t>(b) * S * M + m;                     // in units of head slices (dense layout)
  const char* vb = reinterpret_cast<const char*>(value) + (static_cast<size_t>(b) * S * vps + m * D) * sizeof(T) + c * 16;
  float* myloc = ws.loc + (active ? g : 0) * ws.loc_stride;
  float* myattn = ws.attn + (active ? g : 0) * ws.attn_stride;
  float* mygattn = nullptr;
  const float4* myref = nullptr;
  if constexpr (FUSED) {
    mygattn = ws.gattn + (active ? g : 0) * ws.attn_stride;
    myref = sref + (active ? g : 0) * L;
    fused_prepare<G>(myloc, myattn, myref, L, P, c, active);     // padding groups shadow row 0 read-only
  }

  // grad_out of this pair: raw 16 bytes in the value layout (for the dot products) and, for the fp32
  // scatter of 16-bit types, the fp32 values of the channels this lane adds: lanes of a group then cover
  // 64 contiguous bytes per red.v4.f32 (whole 32-byte sectors) instead of four half-sectors.
  const T* go_row = grad_out + static_cast<size_t>(pair) * D;
  const uint4 go_raw = ldg16(go_row + c * VEC);
  float go_s[VEC];                       // channels [4c, 4c+4) and, for 16-bit T, [D/2 + 4c, D/2 + 4c + 4)
  if constexpr (k16 && !GV16) {
    const uint2 lo = __ldg(reinterpret_cast<const uint2*>(go_row + 4 * c));
    const uint2 hi = __ldg(reinterpret_cast<const uint2*>(go_row + D / 2 + 4 * c));
    float t8[8];
    unpack16<T>(make_uint4(lo.x, lo.y, hi.x, hi.y), t8);
#pragma unroll
    for (int i = 0; i < VEC; ++i) go_s[i] = t8[i];
  } else {
    unpack16<T>(go_raw, go_s);
  }
  // element offset of this lane's first scatter channel inside a head slice
  const int sc0 = (k16 && !GV16) ? 4 * c : c * VEC;
  float* gv32_base = gv32 + (static_cast<size_t>(b) * S * gps + m * D) + sc0;
  const int q = (pair / M) % Lq;
  // bucketed fp16 accumulator: image b starts at row b*accStride; the head / channel offset is the same
  __half* gv16_base = nullptr;
  T* gvd_base = nullptr;                  // sparse levels: this lane's channels of image b in grad_value itself
  size_t acc_row = 0;                     // first accumulator row of (this query's bucket of) the current level
  bool direct = false;                    // current level is sparse: add straight into grad_value (warp-uniform)
  bool red_here = true;                   // HYB: this kernel reduces the current level's grad_value (warp-uniform)
  float gv_unscale = 1.f;
  if constexpr (GV16) {
    gv16_base = gv16 + (static_cast<size_t>(b) * meta.accStride * M + m) * D + c * VEC;
    if constexpr (SPARSE) {
      gvd_base = gv_direct + (static_cast<size_t>(b) * S * gps + m * D) + c * VEC;
      direct = meta.accK[0] == 0;
    }
    acc_row = static_cast<size_t>(meta.accBase[0]) + static_cast<size_t>(direct ? 0 : q % meta.accK[0]) * (meta.H[0] * meta.W[0]);
    if constexpr (HYB) red_here = (meta.sortMask & 1u) == 0u;
    gv_unscale = 1.f / gv_scale;          // exact: power of two
#pragma unroll
    for (int i = 0; i < VEC; ++i) go_s[i] *= gv_scale;      // exact: power-of-two scale
  }
  const uint32_t pix_elems = static_cast<uint32_t>(M) * D;
  const uint32_t gps_elems = static_cast<uint32_t>(gps);

  // Points are processed in chunks of CH; groups of up to 8 lanes reduce-scatter a chunk of G points at once,
  // wider groups (D = 64 fp32, D = 128) fall back to one butterfly per point to keep registers in check.
  constexpr int CH = (G <= 8) ? G : 1;
  float part[CH][3];                     // per point of the current chunk: (grad_attn, grad_x, grad_y) partials
  int l = 0, p = 0;
  int H = meta.H[0], W = meta.W[0];
  float Hf = static_cast<float>(H), Wf = static_cast<float>(W);
  size_t lvl_pix = static_cast<size_t>(meta.start[0]);

  for (int lp0 = 0; lp0 < LP; lp0 += CH) {
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int lp = lp0 + j;
      if (lp < LP) {                      // warp-uniform
        const float2 xy = *reinterpret_cast<const float2*>(myloc + 2 * lp);
        const float a = myattn[lp];
        const Taps t = make_taps(xy.x, xy.y, H, W, Hf, Wf);
        const char* vl = gated_base(t, vb + lvl_pix * pix_bytes);      // zero block for a point that fails the gate
        const uint32_t pg = gated_stride(t, pix_bytes);
        const uint4 u00 = ldg16(vl + static_cast<size_t>(static_cast<uint32_t>(t.i00) * pg));
        const uint4 u01 = ldg16(vl + static_cast<size_t>(static_cast<uint32_t>(t.i01) * pg));
        const uint4 u10 = ldg16(vl + static_cast<size_t>(static_cast<uint32_t>(t.i10) * pg));
        const uint4 u11 = ldg16(vl + static_cast<size_t>(static_cast<uint32_t>(t.i11) * pg));

        // ---- grad_value: (corner weight * attn) * grad_out, scattered with packed reductions ----
        if (active && (!HYB || red_here)) {
          const float ah = t.hhm * a, al = t.lhm * a;
          const float w[4] = {ah * t.hwm, ah * t.lwm, al * t.hwm, al * t.lwm};
          const int idx[4] = {t.i00, t.i01, t.i10, t.i11};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (w[k] != 0.f) {            // invalid corners (and exact-zero weights) add nothing
              float r[VEC];
#pragma unroll
              for (int i = 0; i < VEC; ++i) r[i] = w[k] * go_s[i];
              if constexpr (GV16) {
                if (SPARSE && direct) {   // sparse level: unscaled, in the value dtype, into grad_value
#pragma unroll
                  for (int i = 0; i < VEC; ++i) r[i] *= gv_unscale;
                  red_add_16bit_x8<T>(gvd_base + (lvl_pix + idx[k]) * gps_elems, pack16<T>(r));
                } else {
                  red_add_16bit_x8<__half>(gv16_base + (acc_row + idx[k]) * pix_elems, pack16<__half>(r));
                }
              } else if constexpr (k16) {
                const size_t e = (lvl_pix + idx[k]) * gps_elems;
                red_add_f32x4(gv32_base + e, r[0], r[1], r[2], r[3]);
                red_add_f32x4(gv32_base + e + D / 2, r[4], r[5], r[6], r[7]);
              } else {
                const size_t e = (lvl_pix + idx[k]) * gps_elems;
                red_add_f32x4(gv32_base + e, r[0], r[1], r[2], r[3]);
              }
            }
          }
        }

        // ---- per-corner <value, grad_out>, then this lane's share of the three gradients ----
        const float d00 = dot16<T>(u00, go_raw, 0.f), d01 = dot16<T>(u01, go_raw, 0.f);
        const float d10 = dot16<T>(u10, go_raw, 0.f), d11 = dot16<T>(u11, go_raw, 0.f);
        part[j][0] = t.hhm * (t.hwm * d00 + t.lwm * d01) + t.lhm * (t.hwm * d10 + t.lwm * d11);
        part[j][1] = Wf * a * (t.hhm * (t.Rm * d01 - t.Lm * d00) + t.lhm * (t.Rm * d11 - t.Lm * d10));
        part[j][2] = Hf * a * (t.hwm * (t.Bm * d10 - t.Tm * d00) + t.lwm * (t.Bm * d11 - t.Tm * d01));
        if (++p == P) {                   // next level
          p = 0; ++l;
          if (l < L) {
            H = meta.H[l]; W = meta.W[l]; Hf = static_cast<float>(H); Wf = static_cast<float>(W);
            lvl_pix = static_cast<size_t>(meta.start[l]);
            if constexpr (GV16) {
              if constexpr (SPARSE) direct = meta.accK[l] == 0;
              if constexpr (HYB) red_here = ((meta.sortMask >> l) & 1u) == 0u;
              acc_row = static_cast<size_t>(meta.accBase[l]) + static_cast<size_t>(direct ? 0 : q % meta.accK[l]) * (H * W);
            }
          }
        }
      } else {
        part[j][0] = 0.f; part[j][1] = 0.f; part[j][2] = 0.f;
      }
    }
    int owner;                            // which point of the chunk this lane writes (-1: none)
    if constexpr (CH == G) {
      group_reduce_scatter<G, 3>(part, c);
      owner = c;
    } else {
#pragma unroll
      for (int s = G / 2; s >= 1; s >>= 1) {
        part[0][0] += __shfl_xor_sync(0xffffffffu, part[0][0], s);
        part[0][1] += __shfl_xor_sync(0xffffffffu, part[0][1], s);
        part[0][2] += __shfl_xor_sync(0xffffffffu, part[0][2], s);
      }
      owner = (c == 0) ? 0 : -1;
    }
    __syncwarp();
    // the owner lane now holds the group totals of point lp0 + owner; every lane of the group has already
    // read that point's (x, y, attn), so the gradients can replace them in place
    if (active && owner >= 0 && lp0 + owner < LP) {
      if constexpr (FUSED) {
        const int lpo = lp0 + owner;
        const float4 rs = myref[level_of(lpo, 1.0f / static_cast<float>(P))];
        *reinterpret_cast<float2*>(myloc + 2 * lpo) = make_float2(__fmul_rn(part[0][1], rs.z), __fmul_rn(part[0][2], rs.w));
        mygattn[lpo] = part[0][0];        // the softmax output stays in myattn for the chain rule below
      } else {
        *reinterpret_cast<float2*>(myloc + 2 * (lp0 + owner)) = make_float2(part[0][1], part[0][2]);
        myattn[lp0 + owner] = part[0][0];
      }
    }
  }
  __syncwarp();
  if constexpr (FUSED) {
    {                                     // softmax backward over the pair's L*P entries (all lanes: full-mask shuffles)
      float dot = 0.f;
      for (int i = c; i < LP; i += G) dot += mygattn[i] * myattn[i];
#pragma unroll
      for (int s = G / 2; s >= 1; s >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, s);
      if (active)
        for (int i = c; i < LP; i += G) mygattn[i] = (mygattn[i] - dot) * myattn[i];
    }
    __syncwarp();
    stage_out<AT>(ws, ws.gattn, grad_loc + static_cast<size_t>(pair0) * (2 * LP), grad_attn + static_cast<size_t>(pair0) * LP, nvalid, LP, lane);
  } else {
    stage_out<AT>(ws, ws.attn, grad_loc + static_cast<size_t>(pair0) * (2 * LP), grad_attn + static_cast<size_t>(pair0) * LP, nvalid, LP, lane);
  }
}

// fp32 accumulation buffer (dense, laid out like value) -> 16-bit grad_value (one rounding per element); gps = elements
// between neighbouring pixels of dst, vpp = 16-byte output vectors per pixel row
template <typename T>
__global__ void __launch_bounds__(256)
msda_round_scratch_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t n8, uint32_t vpp, uint32_t gps,
                          const uint32_t* __restrict__ gate, int gate_want) {
  if (gated_off(gate, gate_want)) return;
  const bool dense = gps == vpp * 8u;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (dense) {
      reinterpret_cast<uint4*>(dst)[i] = pack16<T>(f);
    } else {
      const size_t pix = i / vpp;
      const uint32_t v = static_cast<uint32_t>(i - pix * vpp);
      *reinterpret_cast<uint4*>(dst + pix * gps + v * 8u) = pack16<T>(f);
    }
  }
}

// gated memset (the fp32 accumulation buffer of the cluster guard's second pipeline)
static __global__ void __launch_bounds__(256)
msda_zero_fill_kernel(uint4* __restrict__ p, size_t n16, const uint32_t* __restrict__ gate, int gate_want) {
  if (gated_off(gate, gate_want)) return;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    p[i] = z;
}

// zero the control block, the rows of the bucketed fp16 accumulator that the device-side layout actually uses
// (the host only knows the upper bound accum_rows_bound(); for few queries the layout is ~half of it) and, when
// `gv` is given, the grad_value rows of the sparse levels, which receive their reductions directly
// (16-bit element types only: 8 elements per 16-byte vector; gps = elements between neighbouring pixels of gv).
// When `go` is given the same launch also reduces max|grad_out| into *go_max (which the caller has zeroed with a memset,
// so ctrl_vecs is 0 then): the reads of the one pass overlap the writes of the other, one launch and ~25 us less per
// backward than two kernels back to back.
template <typename T>
__global__ void __launch_bounds__(256)
msda_zero_f16_buckets_kernel(uint4* __restrict__ scratch, size_t ctrl_vecs, const int64_t* __restrict__ shapes,
                             const int64_t* __restrict__ lsi, uint16_t* __restrict__ gv, int N, int S, int M, int D,
                             int Lq, int L, int P, int depth, int gps, const T* __restrict__ go, size_t go_n8,
                             uint32_t* __restrict__ go_max, const uint32_t* __restrict__ gate, int gate_want) {
  if (gated_off(gate, gate_want)) return;
  __shared__ LevelMeta meta;
  load_level_meta(meta, shapes, lsi, L);
  build_accum_layout(meta, L, Lq, P, depth, gv != nullptr);
  const size_t vpp = static_cast<size_t>(M) * D / 8;                 // 16-byte vectors per pixel row
  const size_t total = ctrl_vecs + static_cast<size_t>(N) * meta.accStride * vpp;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthr = static_cast<size_t>(gridDim.x) * blockDim.x;
  if (go != nullptr) {
    // interleaved: every round issues four independent 16-byte loads of grad_out, then this thread's share of the zero
    // stores while they are in flight, then folds the loads into the running maximum
    float m = 0.f;
    const size_t nl = (go_n8 + nthr - 1) / nthr;                     // load rounds of this thread
    const size_t ns = (total + nthr - 1) / nthr;                     // store rounds
    const size_t rounds = (nl + 3) / 4;
    const size_t spr = rounds ? (ns + rounds - 1) / rounds : ns;     // stores per round of four loads
    size_t ks = 0;
    for (size_t kl = 0; kl < nl; kl += 4) {
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t idx = tid + (kl + j) * nthr;
        u[j] = idx < go_n8 ? __ldg(reinterpret_cast<const uint4*>(go) + idx) : z;
      }
      const size_t ks_end = ks + spr < ns ? ks + spr : ns;
      for (; ks < ks_end; ++ks) {
        const size_t idx = tid + ks * nthr;
        if (idx < total) scratch[idx] = z;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
        unpack16<T>(u[j], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) m = fmaxf(m, fabsf(f[k]));
      }
    }
    for (; ks < ns; ++ks) {
      const size_t idx = tid + ks * nthr;
      if (idx < total) scratch[idx] = z;
    }
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sft));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(go_max, __float_as_uint(m));
  } else {
    for (size_t i = tid; i < total; i += nthr) scratch[i] = z;
  }
  if (meta.dirRows > 0) {
    // a sparse level's rows are one contiguous range per image: level by level, (image, vector) flattened, with 32-bit
    // index arithmetic whenever it fits (the 64-bit divisions of a row-by-row walk made this pass instruction-bound:
    // 47 us for the 167 MB of the 300-query decoder shape)
    const bool dense = static_cast<size_t>(gps) == static_cast<size_t>(M) * D;
    for (int l = 0; l < L; ++l) {
      if (meta.accK[l] != 0) continue;
      const size_t n16 = static_cast<size_t>(meta.H[l]) * meta.W[l] * vpp;         // vectors of the level in one image
      const size_t ltotal = static_cast<size_t>(N) * n16;
      uint16_t* lvl = gv + static_cast<size_t>(meta.start[l]) * gps;
      const size_t img_elems = static_cast<size_t>(S) * gps;
      if (ltotal < (1ull << 32)) {
        const uint32_t n16u = static_cast<uint32_t>(n16), vppu = static_cast<uint32_t>(vpp);
        for (size_t i = tid; i < ltotal; i += nthr) {
          const uint32_t iu = static_cast<uint32_t>(i), b = iu / n16u, j = iu - b * n16u;
          uint16_t* img = lvl + b * img_elems;
          if (dense) {
            reinterpret_cast<uint4*>(img)[j] = z;
          } else {
            const uint32_t r = j / vppu, v = j - r * vppu;
            *reinterpret_cast<uint4*>(img + static_cast<size_t>(r) * gps + v * 8) = z;
          }
        }
      } else {
        for (size_t i = tid; i < ltotal; i += nthr) {
          const size_t b = i / n16, j = i - b * n16, r = j / vpp, v = j - r * vpp;
          *reinterpret_cast<uint4*>(lvl + b * img_elems + r * gps + v * 8) = z;
        }
      }
    }
  }
}

// bucketed, scaled fp16 accumulation buffer -> 16-bit grad_value: sum the K_l copies in fp32, unscale, round once.
// blockIdx.y = image; one flattened index over the 16-byte vectors of the image's bucketed rows (a sparse level's
// grad_value rows already hold their sums), so every thread gets the same share whatever the levels' sizes; the
// level of a vector is found by comparing against the levels' first indices, its constants come from a small shared
// table (the first version spent ~160 instructions per vector on divisions, the per-element level search and 64-bit
// index arithmetic: issue-bound at 0.09 ms for the cfg3 shape, 0.03 ms for the few rows of the decoder shape).
// 32-bit index arithmetic: S * M * D * 4 bytes < 2^32 for the vector kernels, and the accumulator of one image holds
// fewer vectors than that.
template <typename T>
__global__ void __launch_bounds__(256)
msda_round_f16_buckets_kernel(const __half* __restrict__ acc, T* __restrict__ dst, const int64_t* __restrict__ shapes,
                              const int64_t* __restrict__ lsi, const uint32_t* __restrict__ ctrl,
                              int N, int S, int M, int D, int Lq, int L, int P, int depth, int sparse_direct, int gps,
                              const uint32_t* __restrict__ gate, int gate_want) {
  if (gated_off(gate, gate_want)) return;
  __shared__ LevelMeta meta;
  struct LevelRow { uint32_t first, K, kstride, src0, drow0; };   // first flattened vector, copies, vectors per copy, ...
  __shared__ LevelRow tab[kMaxLevels];
  __shared__ int s_nlv;
  load_level_meta(meta, shapes, lsi, L);
  build_accum_layout(meta, L, Lq, P, depth, sparse_direct != 0);
  const uint32_t vpp = static_cast<uint32_t>(M * D / 8);
  if (threadIdx.x == 0) {
    int n = 0;
    for (int l = 0; l < L; ++l) {
      if (meta.accK[l] == 0) continue;
      LevelRow t;
      t.first = static_cast<uint32_t>(meta.bktOff[l]) * vpp;
      t.K = static_cast<uint32_t>(meta.accK[l]);
      t.kstride = static_cast<uint32_t>(meta.H[l] * meta.W[l]) * vpp;
      t.src0 = static_cast<uint32_t>(meta.accBase[l]) * vpp - t.first;      // + i: vector of copy 0 inside the image's accumulator
      t.drow0 = static_cast<uint32_t>(meta.start[l] - meta.bktOff[l]);      // + row among the bucketed rows: pixel row in the image
      tab[n++] = t;
    }
    s_nlv = n;
  }
  __syncthreads();
  const int nlv = s_nlv;
  const float inv = 1.f / f16_accum_scale(ctrl, Lq);        // exact: power of two
  const uint32_t per_img = static_cast<uint32_t>(meta.bktRows) * vpp;
  const bool dense = static_cast<uint32_t>(gps) == vpp * 8u;
  for (int b = blockIdx.y; b < N; b += gridDim.y) {
    const uint4* src_img = reinterpret_cast<const uint4*>(acc) + static_cast<size_t>(b) * meta.accStride * vpp;
    T* dst_img = dst + static_cast<size_t>(b) * S * gps;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += gridDim.x * blockDim.x) {
      int li = 0;
      for (int k = 1; k < nlv; ++k) li += (i >= tab[k].first) ? 1 : 0;          // levels are in ascending order of `first`
      const LevelRow t = tab[li];
      const uint4* src = src_img + (t.src0 + i);
      float sum[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) sum[j] = 0.f;
      uint32_t k = 0;
      for (; k + 4 <= t.K; k += 4) {                                             // coarse levels: 4 loads in flight
        uint4 u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) u[q] = __ldcs(src + static_cast<size_t>(k + q) * t.kstride);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float f[8];
          unpack16<__half>(u[q], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) sum[j] += f[j];
        }
      }
      for (; k < t.K; ++k) {
        float f[8];
        unpack16<__half>(__ldcs(src + static_cast<size_t>(k) * t.kstride), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[j] += f[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum[j] *= inv;
      if (dense) {
        reinterpret_cast<uint4*>(dst_img)[t.drow0 * vpp + i] = pack16<T>(sum);
      } else {
        const uint32_t r = i / vpp, v = i - r * vpp;
        *reinterpret_cast<uint4*>(dst_img + static_cast<size_t>(t.drow0 + r) * gps + v * 8) = pack16<T>(sum);
      }
    }
  }
}

// =====================================================================================================
// Cluster guard.  The 16-bit accumulation modes size their precision for evenly spread sampling: an fp16 bucket is
// meant to receive ~depth adds, a sparse level's grad_value element zero or one.  When many queries crowd onto a few
// pixels (measured: > 1 000 sampling points on one pixel of the 300-query decoder shape) both drift past the 2e-2 gate,
// while fp32 accumulation does not care.  For calls small enough that one more pass over sampling_loc is noise
// (kGuardMaxPoints), this kernel counts the sampling points per (run of 4 pixels, head) -- top-left corner, all four
// corners land within one pixel of it -- and raises ctrl[1] as soon as a cell holds more than its level tolerates:
// 32 points for a sparse level (packed adds in the value dtype), 512 per bucket for a bucketed one (fp16, ~11 bits).  Every kernel of the
// two accumulation pipelines then checks the flag (gated_off) and only the selected pipeline does any work.
// =====================================================================================================
constexpr int kGuardSparseLimit = 32;
constexpr int kGuardBucketLimit = 512;

static __global__ void __launch_bounds__(256)
msda_cluster_density_kernel(const float* __restrict__ loc, const int64_t* __restrict__ shapes,
                            const int64_t* __restrict__ lsi, uint32_t* __restrict__ counters, uint32_t* __restrict__ ctrl,
                            long long total_points, int cells, int M, int Lq, int L, int P, int depth, int sparse_direct) {
  __shared__ LevelMeta meta;
  load_level_meta(meta, shapes, lsi, L);
  build_accum_layout(meta, L, Lq, P, depth, sparse_direct != 0);
  bool over = false;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total_points;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int l = static_cast<int>((idx / P) % L);
    const int m = static_cast<int>((idx / (static_cast<long long>(P) * L)) % M);
    const long long b = idx / (static_cast<long long>(P) * L * M * Lq);
    const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + idx);
    const int H = meta.H[l], W = meta.W[l];
    const Taps t = make_taps(xy.x, xy.y, H, W, static_cast<float>(H), static_cast<float>(W));
    if (!t.inside) continue;
    const uint32_t cell = static_cast<uint32_t>(meta.start[l] + t.i00) >> 2;
    const uint32_t c = atomicAdd(counters + (static_cast<size_t>(b) * cells + cell) * M + m, 1u) + 1u;
    const int K = meta.accK[l];
    over |= c > static_cast<uint32_t>(K == 0 ? kGuardSparseLimit : kGuardBucketLimit * K);
  }
  if (__any_sync(0xffffffffu, over) && (threadIdx.x & 31) == 0) atomicOr(ctrl + 1, 1u);
}

// Tiled kernels for the dense encoder call site (Lq == S): shared-memory value windows, sorted grad_value sums
#include "msda_tiled.cuh"

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
constexpr size_t kF16CtrlBytes = 256;      // control block in front of the fp16 accumulation buffer
constexpr int kDefaultAccumDepth = 32;     // expected adds per fp16 accumulator element (see build_accum_layout)

static int accum_depth(int flags) {
  const int d = (flags >> 8) & 0xffff;
  return d > 0 ? d : kDefaultAccumDepth;
}

static size_t f16_scratch_bytes(int N, int S, int M, int D, int Lq, int L, int P, int depth) {
  const long long rows = accum_rows_bound(S, L, Lq, P, depth);
  return kF16CtrlBytes + static_cast<size_t>(N) * static_cast<size_t>(rows) * M * D * sizeof(__half);
}
// ---- build layout -----------------------------------------------------------------------------------------
// This file compiles either as ONE translation unit (no macro: everything below) or, to build in parallel, several
// times (see __graft_entry__.build): -DMSDA_TU_DTYPE=0|1|2 emits the kernels + launchers of float/double | bfloat16 |
// float16 by explicit instantiation of launch_fwd<T> / launch_bwd<T>, and -DMSDA_TU_ABI emits the shared host state
// and the C ABI, which reaches the launchers through `extern template`.
#if defined(MSDA_TU_DTYPE) || defined(MSDA_TU_ABI)
#define MSDA_SPLIT_BUILD 1
#endif
#if defined(MSDA_SPLIT_BUILD) && !defined(MSDA_TU_ABI)
#define MSDA_STATE extern
#else
#define MSDA_STATE
#endif
#ifdef MSDA_SPLIT_BUILD
#define MSDA_LAUNCHER
#else
#define MSDA_LAUNCHER static
#endif

#if defined(MSDA_SPLIT_BUILD) && !defined(MSDA_TU_ABI)
extern thread_local int g_last_launches;
extern std::atomic<long long> g_total_launches;
extern std::atomic<int> g_tiled_mode;
extern std::atomic<int> g_hybrid_split;
#else
thread_local int g_last_launches = 0;
std::atomic<long long> g_total_launches{0};
std::atomic<int> g_tiled_mode{-1};         // msda_set_tiled_mode(); -1 = take MSDA_B200_TILED (default 0) from the environment on first use
std::atomic<int> g_hybrid_split{8};        // msda_set_hybrid_split(): see build_accum_layout
#endif

// SM count of the current device (cached per thread; one process drives one GPU)
static int device_sm_count() {
  static thread_local int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1)
      sms = 148;
  }
  return sms;
}

// 0 = direct kernels, 1 = tiled forward + tiled backward (dots + sort), 2 = hybrid backward (direct kernel for the
// gradients of locations / weights and the fine levels' grad_value, sorting kernel for the coarse levels; direct forward)
static int tiled_mode() {
  int m = g_tiled_mode.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = std::getenv("MSDA_B200_TILED");
    m = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 0;
    g_tiled_mode.store(m, std::memory_order_relaxed);
  }
  return m;
}

// ---- optional per-launch timing of the dominant kernels (bench.py's roofline leg) --------------------
struct ProfileRecord { cudaEvent_t start, stop; int kind; };
// process-wide (autograd runs backward on its own thread), guarded by a mutex; off by default
#if defined(MSDA_SPLIT_BUILD) && !defined(MSDA_TU_ABI)
extern std::atomic<bool> g_profile_on;
extern std::mutex g_profile_mu;
extern std::vector<ProfileRecord> g_profile;
#else
std::atomic<bool> g_profile_on{false};
std::mutex g_profile_mu;
std::vector<ProfileRecord> g_profile;
#endif

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
  const float* ref = nullptr;   // fused pre-op: reference points (N, Lq, L, R); nullptr = plain operator
  int R = 0;
  bool aux16 = false;           // fused pre-op only: offsets / logits (and their gradients) are in the 16-bit value type
  bool hybrid = false;          // hybrid backward: the direct kernel leaves the levels in meta.sortMask to the sorting kernel
  long long vps = 0, gps = 0;   // elements between neighbouring pixels of value / grad_value (0 = dense, M*D)
  // cluster guard: when set, a backward kernel returns at once unless (*gate != 0) == (gate_want != 0)
  const uint32_t* gate = nullptr;
  int gate_want = 0;
  int value_stride() const { return static_cast<int>(vps > 0 ? vps : static_cast<long long>(M) * D); }
  int grad_stride() const { return static_cast<int>(gps > 0 ? gps : static_cast<long long>(M) * D); }
  bool strided() const { return value_stride() != M * D || grad_stride() != M * D; }
};

// pixel strides of the *_strided entry points: at least one dense pixel row, whole 16-byte vectors
static int validate_strides(const Problem& pr, size_t elem_bytes) {
  const long long dense = static_cast<long long>(pr.M) * pr.D;
  for (long long st : {pr.vps, pr.gps}) {
    if (st == 0) continue;
    if (st < dense || st >= (1ll << 31) || (st * static_cast<long long>(elem_bytes)) % 16 != 0) return MSDA_ERR_BAD_STRIDE;
  }
  return MSDA_OK;
}

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

// shared memory the vector kernels stage per CTA (the backward's is the larger one; see vec_smem_bytes): L*P rows of a
// warp's pairs must fit next to each other, which caps L*P at ~130 (16-bit, D = 16) ... ~530 (fp32, D = 32)
constexpr size_t kVecSmemLimit = 200 * 1024;
static size_t vec_smem_bytes_rt(int D, size_t elem_bytes, int L, int P, bool fused) {
  const int G = static_cast<int>(D * elem_bytes / 16);
  const int GPW = 32 / (G > 0 ? G : 1);
  const int LP = L * P;
  return static_cast<size_t>(kWarps) * GPW * (loc_row_stride(LP) + (fused ? 2 : 1) * attn_row_stride(LP) + (fused ? 4 * L : 0)) *
         sizeof(float);
}

// vector kernels: head dims with power-of-two lane groups, per-image byte offsets that fit 32 bits, and staged rows that
// fit shared memory; everything else (upstream has no such limits) runs on the compatibility kernels
static bool vec_supported_rt(int S, int M, int D, int L, int P, size_t elem_bytes, long long row_elems, bool fused) {
  const bool d_ok = D == 16 || D == 32 || D == 64 || D == 128;
  if (!d_ok) return false;
  if (static_cast<unsigned long long>(S) * static_cast<unsigned long long>(row_elems) * sizeof(float) >= (1ull << 32)) return false;
  return vec_smem_bytes_rt(D, elem_bytes, L, P, fused) <= kVecSmemLimit;
}
template <typename T> static bool vec_supported(const Problem& pr) {
  return vec_supported_rt(pr.S, pr.M, pr.D, pr.L, pr.P, sizeof(T), std::max(pr.value_stride(), pr.grad_stride()), pr.ref != nullptr);
}

// per warp: GPW rows of loc (2*LP + 4 floats) and attn (LP + 4); fused backward adds a gattn row per pair; fused kernels
// add the float4 (rx, ry, sx, sy) row of every (pair, level) behind all warps' rows
template <typename T, int D>
static size_t vec_smem_bytes(int L, int P, bool fused = false, bool bwd = false) {
  constexpr int G = D / (16 / static_cast<int>(sizeof(T)));
  constexpr int GPW = 32 / G;
  const bool fb = fused && bwd;
  const int LP = L * P;
  return static_cast<size_t>(kWarps) * GPW * (loc_row_stride(LP) + (fb ? 2 : 1) * attn_row_stride(LP) + (fused ? 4 * L : 0)) *
         sizeof(float);
}

// L1 / shared-memory split of the gather kernels.  Left alone, the driver configured 102 KB of shared memory for the
// backward (3 CTAs x 16 KB needed: ncu launch__shared_mem_config_size), i.e. 126 KB of L1 for a kernel whose gather misses
// each cost a cycle of the SM's crossbar request port -- the unit it saturates.  Ask for just what
// `ctas` resident CTAs need (a hint: the driver rounds up to the next configuration).  MSDA_B200_CARVEOUT=<percent>
// overrides, -1 keeps the driver's choice.
static int carveout_percent(size_t dyn_bytes, int ctas) {
  static const int forced = [] {
    const char* e = std::getenv("MSDA_B200_CARVEOUT");
    return e ? std::atoi(e) : -2;
  }();
  if (forced >= -1) return forced;
  const size_t need = static_cast<size_t>(ctas) * (dyn_bytes + 2048);      // + static LevelMeta + the driver's 1 KB per CTA
  const int pct = static_cast<int>((need * 100 + 233471) / 233472);
  return pct > 100 ? 100 : pct;
}

template <typename K>
static cudaError_t allow_smem(K kernel, size_t bytes, int ctas = 0) {
  if (ctas > 0) {
    struct Entry { const void* fn; int pct; };
    static thread_local Entry cache[32];
    static thread_local int cached = 0;
    const int pct = carveout_percent(bytes, ctas);
    const void* key = reinterpret_cast<const void*>(kernel);
    Entry* slot = nullptr;
    for (int i = 0; i < cached && slot == nullptr; ++i)
      if (cache[i].fn == key) slot = &cache[i];
    if (slot == nullptr && cached < 32) { slot = &cache[cached++]; *slot = Entry{key, -1000}; }
    if ((slot == nullptr || slot->pct != pct) && pct >= 0) {        // set once per kernel (and again if the need changes)
      const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      if (e != cudaSuccess) return e;
    }
    if (slot != nullptr) slot->pct = pct;
  }
  // the kernels also hold a static LevelMeta (< 1 KB): opt in as soon as dynamic + static could pass the 48 KB default
  if (bytes + 1024 <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
}

template <typename T, int D, bool FUSED, typename AT>
static int launch_fwd_vec_impl(const Problem& pr, const void* value, const int64_t* shapes, const int64_t* lsi,
                               const void* loc, const void* attn, void* out, cudaStream_t st) {
  constexpr int GPW = 32 / (D / (16 / static_cast<int>(sizeof(T))));
  const int per_cta = kWarps * GPW;
  const int grid = (pr.total_pairs + per_cta - 1) / per_cta;
  const size_t smem = vec_smem_bytes<T, D>(pr.L, pr.P, FUSED, false);
  if (smem > kVecSmemLimit) return MSDA_ERR_BAD_SHAPE;       // not reached: vec_supported() sends such shapes elsewhere
  cudaError_t e = allow_smem(msda_fwd_vec_kernel<T, D, FUSED, AT>, smem, FUSED ? 5 : 6);
  if (e != cudaSuccess) return static_cast<int>(e);
  ScopedKernelTimer timer(MSDA_KERNEL_FORWARD, st);
  msda_fwd_vec_kernel<T, D, FUSED, AT><<<grid, kThreads, smem, st>>>(
      static_cast<const T*>(value), shapes, lsi, static_cast<const AT*>(loc), static_cast<const AT*>(attn),
      pr.ref, pr.R, static_cast<T*>(out), pr.S, pr.M, pr.Lq, pr.L, pr.P, pr.total_pairs, pr.value_stride());
  ++g_last_launches, ++g_total_launches;
  return static_cast<int>(cudaGetLastError());
}

template <typename T, int D>
static int launch_fwd_vec(const Problem& pr, const void* value, const int64_t* shapes, const int64_t* lsi,
                          const void* loc, const void* attn, void* out, cudaStream_t st) {
  if constexpr (sizeof(T) == 2) {
    if (pr.ref && pr.aux16) return launch_fwd_vec_impl<T, D, true, T>(pr, value, shapes, lsi, loc, attn, out, st);
  }
  if (pr.ref) return launch_fwd_vec_impl<T, D, true, float>(pr, value, shapes, lsi, loc, attn, out, st);
  return launch_fwd_vec_impl<T, D, false, float>(pr, value, shapes, lsi, loc, attn, out, st);
}

// ---- tiled kernels (msda_tiled.cuh): dense call site, 16-bit values, head dim 32 ---------------------------
// What the host can check without reading the shape tensor; everything else (level nesting, halo, window capacity) is
// decided on the device, where a point that does not fit its window takes the slow path.
template <typename T> static bool tiled_shape_ok(const Problem& pr) {
  return sizeof(T) == 2 && pr.D == tiled::kD && pr.Lq == pr.S && pr.ref == nullptr && pr.L <= tiled::kMaxL &&
         pr.L * pr.P <= tiled::kMaxLP && vec_supported<T>(pr);
}
template <typename T> static bool tiled_supported(const Problem& pr) { return tiled_mode() == 1 && tiled_shape_ok<T>(pr); }
// hybrid backward: dense values / gradients only (the sorting kernel writes the accumulator, not grad_value, so strides
// would be fine -- but keep the first version to the shape that is measured)
template <typename T> static bool hybrid_supported(const Problem& pr) {
  return tiled_mode() == 2 && tiled_shape_ok<T>(pr) && g_hybrid_split.load(std::memory_order_relaxed) > 0;
}

// persistent grid: one resident wave of the kernel (the two device queries are cached per kernel and thread)
template <typename K>
static int tiled_grid(K kernel, int threads, size_t dyn_smem) {
  struct Entry { const void* fn; int per_sm; };
  static thread_local Entry cache[16];
  static thread_local int cached = 0, sms = 0;
  const void* key = reinterpret_cast<const void*>(kernel);
  for (int i = 0; i < cached; ++i)
    if (cache[i].fn == key) return cache[i].per_sm * sms;
  int dev = 0;
  if (sms == 0 && (cudaGetDevice(&dev) != cudaSuccess ||
                   cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1))
    sms = 148;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  if (cached < 16) cache[cached++] = Entry{key, per_sm};
  return per_sm * sms;
}

template <typename T>
static int launch_fwd_tiled(const Problem& pr, const void* value, const int64_t* shapes, const int64_t* lsi,
                            const void* loc, const void* attn, void* out, cudaStream_t st) {
  auto kernel = tiled::msda_fwd_tiled_kernel<T>;
  const size_t smem = tiled::kFwdSmemBytes;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int grid = tiled_grid(kernel, tiled::kThreadsT, smem);
  ScopedKernelTimer timer(MSDA_KERNEL_FORWARD, st);
  kernel<<<grid, tiled::kThreadsT, smem, st>>>(static_cast<const T*>(value), shapes, lsi, static_cast<const float*>(loc),
                                               static_cast<const float*>(attn), static_cast<T*>(out), pr.N, pr.S, pr.M,
                                               pr.Lq, pr.L, pr.P, pr.value_stride());
  ++g_last_launches, ++g_total_launches;
  return static_cast<int>(cudaGetLastError());
}

// grad_value of all levels (tiled backward) or of the levels in the layout's sortMask (hybrid backward) into the fp16
// accumulator; ctrl[0] = max|grad_out| must be there already
template <typename T>
static int launch_bwd_scatter(const Problem& pr, const int64_t* shapes, const int64_t* lsi, const void* loc, const void* attn,
                              const void* go, __half* acc16, const uint32_t* ctrl, int depth, cudaStream_t st) {
  auto kernel = tiled::msda_bwd_scatter_tiled_kernel<T>;
  const size_t smem = tiled::kScSmemBytes;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int grid = tiled_grid(kernel, tiled::kScThreads, smem);
  ScopedKernelTimer timer(MSDA_KERNEL_BACKWARD_SCATTER, st);
  kernel<<<grid, tiled::kScThreads, smem, st>>>(shapes, lsi, static_cast<const float*>(loc), static_cast<const float*>(attn),
                                                static_cast<const T*>(go), acc16, ctrl, pr.N, pr.S, pr.M, pr.Lq, pr.L,
                                                pr.P, depth);
  ++g_last_launches, ++g_total_launches;
  return static_cast<int>(cudaGetLastError());
}

// grad_sampling_loc / grad_attn_weight (+ max|grad_out| into ctrl[0]), then grad_value into the fp16 accumulator
template <typename T>
static int launch_bwd_tiled(const Problem& pr, const void* value, const int64_t* shapes, const int64_t* lsi,
                            const void* loc, const void* attn, const void* go, __half* acc16, uint32_t* ctrl,
                            void* gloc, void* gattn, int depth, cudaStream_t st) {
  {
    auto kernel = tiled::msda_bwd_dots_tiled_kernel<T>;
    const size_t smem = tiled::kDotsSmemBytes;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    const int grid = tiled_grid(kernel, tiled::kThreadsT, smem);
    ScopedKernelTimer timer(MSDA_KERNEL_BACKWARD_DOTS, st);
    kernel<<<grid, tiled::kThreadsT, smem, st>>>(static_cast<const T*>(value), shapes, lsi, static_cast<const float*>(loc),
                                                 static_cast<const float*>(attn), static_cast<const T*>(go),
                                                 static_cast<float*>(gloc), static_cast<float*>(gattn), ctrl, pr.N, pr.S,
                                                 pr.M, pr.Lq, pr.L, pr.P, pr.value_stride());
    ++g_last_launches, ++g_total_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  return launch_bwd_scatter<T>(pr, shapes, lsi, loc, attn, go, acc16, ctrl, depth, st);
}

template <typename T>
MSDA_LAUNCHER int launch_fwd(const Problem& pr, const void* value, const int64_t* shapes, const int64_t* lsi,
                      const void* loc, const void* attn, void* out, cudaStream_t st) {
  if constexpr (sizeof(T) == 2) {
    if (tiled_supported<T>(pr)) return launch_fwd_tiled<T>(pr, value, shapes, lsi, loc, attn, out, st);
  }
  if constexpr (!std::is_same<T, double>::value) {
    if (vec_supported<T>(pr)) {
      switch (pr.D) {
        case 16: return launch_fwd_vec<T, 16>(pr, value, shapes, lsi, loc, attn, out, st);
        case 32: return launch_fwd_vec<T, 32>(pr, value, shapes, lsi, loc, attn, out, st);
        case 64: return launch_fwd_vec<T, 64>(pr, value, shapes, lsi, loc, attn, out, st);
        case 128: return launch_fwd_vec<T, 128>(pr, value, shapes, lsi, loc, attn, out, st);
      }
    }
  }
  if (pr.ref) return MSDA_ERR_FUSED_UNSUPPORTED;
  if (pr.strided()) return MSDA_ERR_BAD_STRIDE;         // the compatibility kernels read dense values only
  using Aux = typename Traits<T>::Aux;
  const int grid = (pr.total_pairs + kWarps - 1) / kWarps;
  msda_fwd_any_kernel<T><<<grid, kThreads, 0, st>>>(
      static_cast<const T*>(value), shapes, lsi, static_cast<const Aux*>(loc), static_cast<const Aux*>(attn),
      static_cast<T*>(out), pr.S, pr.M, pr.D, pr.Lq, pr.L, pr.P, pr.total_pairs);
  ++g_last_launches, ++g_total_launches;
  return static_cast<int>(cudaGetLastError());
}

template <typename T, int D, bool FUSED, typename AT>
static int launch_bwd_vec_impl(const Problem& pr, const void* value, const int64_t* shapes, const int64_t* lsi,
                               const void* loc, const void* attn, const void* go, float* gv32, __half* gv16,
                               void* gv_direct, int gv_stride,
                               const uint32_t* ctrl, void* gloc, void* gattn, bool use16, int depth, cudaStream_t st) {
  constexpr int GPW = 32 / (D / (16 / static_cast<int>(sizeof(T))));
  const size_t smem = vec_smem_bytes<T, D>(pr.L, pr.P, FUSED, true);
  if (smem > kVecSmemLimit) return MSDA_ERR_BAD_SHAPE;       // not reached: vec_supported() sends such shapes elsewhere
  const int per_cta = kWarps * GPW;
  const int grid = (pr.total_pairs + per_cta - 1) / per_cta;
  cudaError_t e;
  if constexpr (sizeof(T) == 2) {
    if (use16) {
      auto launch16 = [&](auto kernel) -> int {
        cudaError_t err = allow_smem(kernel, smem, BWD_MIN_CTAS);
        if (err != cudaSuccess) return static_cast<int>(err);
        ScopedKernelTimer timer(MSDA_KERNEL_BACKWARD, st);
        kernel<<<grid, kThreads, smem, st>>>(
            static_cast<const T*>(value), shapes, lsi, static_cast<const AT*>(loc), static_cast<const AT*>(attn),
            pr.ref, pr.R, static_cast<const T*>(go), nullptr, gv16, static_cast<T*>(gv_direct), ctrl, static_cast<AT*>(gloc),
            static_cast<AT*>(gattn), pr.S, pr.M, pr.Lq, pr.L, pr.P, pr.total_pairs, depth, pr.value_stride(), gv_stride,
            pr.gate, pr.gate_want);
        ++g_last_launches, ++g_total_launches;
        return static_cast<int>(cudaGetLastError());
      };
      if constexpr (D == tiled::kD && !FUSED && std::is_same<AT, float>::value) {
        if (pr.hybrid) return launch16(msda_bwd_vec_kernel<T, D, true, false, float, false, true>);
      }
      return gv_direct ? launch16(msda_bwd_vec_kernel<T, D, true, FUSED, AT, true>)
                       : launch16(msda_bwd_vec_kernel<T, D, true, FUSED, AT, false>);
    }
  }
  e = allow_smem(msda_bwd_vec_kernel<T, D, false, FUSED, AT, false>, smem, BWD_MIN_CTAS);
  if (e != cudaSuccess) return static_cast<int>(e);
  ScopedKernelTimer timer(MSDA_KERNEL_BACKWARD, st);
  msda_bwd_vec_kernel<T, D, false, FUSED, AT, false><<<grid, kThreads, smem, st>>>(
      static_cast<const T*>(value), shapes, lsi, static_cast<const AT*>(loc), static_cast<const AT*>(attn),
      pr.ref, pr.R, static_cast<const T*>(go), gv32, nullptr, nullptr, nullptr, static_cast<AT*>(gloc), static_cast<AT*>(gattn),
      pr.S, pr.M, pr.Lq, pr.L, pr.P, pr.total_pairs, depth, pr.value_stride(), gv_stride, pr.gate, pr.gate_want);
  ++g_last_launches, ++g_total_launches;
  return static_cast<int>(cudaGetLastError());
}

template <typename T, int D>
static int launch_bwd_vec(const Problem& pr, const void* value, const int64_t* shapes, const int64_t* lsi,
                          const void* loc, const void* attn, const void* go, float* gv32, __half* gv16,
                          void* gvd, int gvs,
                          const uint32_t* ctrl, void* gloc, void* gattn, bool use16, int depth, cudaStream_t st) {
  if constexpr (sizeof(T) == 2) {
    if (pr.ref && pr.aux16)
      return launch_bwd_vec_impl<T, D, true, T>(pr, value, shapes, lsi, loc, attn, go, gv32, gv16, gvd, gvs, ctrl, gloc, gattn, use16, depth, st);
  }
  if (pr.ref)
    return launch_bwd_vec_impl<T, D, true, float>(pr, value, shapes, lsi, loc, attn, go, gv32, gv16, gvd, gvs, ctrl, gloc, gattn, use16, depth, st);
  return launch_bwd_vec_impl<T, D, false, float>(pr, value, shapes, lsi, loc, attn, go, gv32, gv16, gvd, gvs, ctrl, gloc, gattn, use16, depth, st);
}

// ---- scratch layout of the 16-bit backward ---------------------------------------------------------------
//   [ control block kF16CtrlBytes ][ cluster-guard counters ][ payload ]
// control block: ctrl[0] = bits of max|grad_out|, ctrl[1] = cluster flag.  The counters exist only for calls the guard
// covers (by size alone, so that msda_backward_scratch_bytes and the launcher always agree); the payload is the scaled
// fp16 bucket accumulator or, when the guard may switch to it, the larger of that and a dense fp32 copy of grad_value.
constexpr long long kGuardMaxPoints = 4ll << 20;     // sampling points per call up to which the density pass runs

struct ScratchLayout {
  bool guard;
  int cells;                 // 4-pixel cells per image (+1)
  size_t counters_off, counters_bytes, payload_off, total;
};

static ScratchLayout scratch_layout_16(int N, int S, int M, int D, int Lq, int L, int P, int flags) {
  ScratchLayout sl{};
  const long long points = static_cast<long long>(N) * Lq * M * L * P;
  sl.guard = !(flags & MSDA_BWD_NO_CLUSTER_GUARD) && points <= kGuardMaxPoints;
  sl.cells = (S + 3) / 4 + 1;
  sl.counters_off = kF16CtrlBytes;
  sl.counters_bytes = sl.guard ? ((static_cast<size_t>(N) * sl.cells * M * sizeof(uint32_t) + 255) & ~static_cast<size_t>(255)) : 0;
  sl.payload_off = sl.counters_off + sl.counters_bytes;
  size_t payload = f16_scratch_bytes(N, S, M, D, Lq, L, P, accum_depth(flags)) - kF16CtrlBytes;
  if (sl.guard) payload = std::max(payload, static_cast<size_t>(N) * S * M * D * sizeof(float));
  sl.total = sl.payload_off + payload;
  return sl;
}

// 16-bit values, fp32 accumulation: zero the dense fp32 buffer, backward kernel with red.v4.f32, one rounding pass.
// `gated`: part of the cluster guard's pair of pipelines (kernels check pr.gate; the memset becomes a gated kernel).
template <typename T>
static int run_bwd_fp32_accum(const Problem& pr, const void* value, const int64_t* shapes, const int64_t* lsi,
                              const void* loc, const void* attn, const void* go, void* gv, void* gloc, void* gattn,
                              float* acc32, int depth, bool gated, cudaStream_t st) {
  const size_t n_value = static_cast<size_t>(pr.N) * pr.S * pr.M * pr.D;
  const int sms = device_sm_count();
  cudaError_t e;
  if (gated) {
    const size_t n16 = n_value * sizeof(float) / 16;
    msda_zero_fill_kernel<<<static_cast<int>(std::min<size_t>((n16 + 255) / 256, static_cast<size_t>(sms) * 8)), 256, 0, st>>>(
        reinterpret_cast<uint4*>(acc32), n16, pr.gate, pr.gate_want);
    ++g_last_launches, ++g_total_launches;
    e = cudaGetLastError();
  } else {
    e = cudaMemsetAsync(acc32, 0, n_value * sizeof(float), st);
  }
  if (e != cudaSuccess) return static_cast<int>(e);
  int rc;
  if (vec_supported<T>(pr)) {
    const int gvs = pr.M * pr.D;                         // the fp32 buffer is dense
    switch (pr.D) {
      case 16: rc = launch_bwd_vec<T, 16>(pr, value, shapes, lsi, loc, attn, go, acc32, nullptr, nullptr, gvs, nullptr, gloc, gattn, false, depth, st); break;
      case 32: rc = launch_bwd_vec<T, 32>(pr, value, shapes, lsi, loc, attn, go, acc32, nullptr, nullptr, gvs, nullptr, gloc, gattn, false, depth, st); break;
      case 64: rc = launch_bwd_vec<T, 64>(pr, value, shapes, lsi, loc, attn, go, acc32, nullptr, nullptr, gvs, nullptr, gloc, gattn, false, depth, st); break;
      default: rc = launch_bwd_vec<T, 128>(pr, value, shapes, lsi, loc, attn, go, acc32, nullptr, nullptr, gvs, nullptr, gloc, gattn, false, depth, st); break;
    }
  } else {
    if (pr.ref) return MSDA_ERR_FUSED_UNSUPPORTED;
    using Aux = typename Traits<T>::Aux;
    const int grid = (pr.total_pairs + kWarps - 1) / kWarps;
    msda_bwd_any_kernel<T, float><<<grid, kThreads, 0, st>>>(
        static_cast<const T*>(value), shapes, lsi, static_cast<const Aux*>(loc), static_cast<const Aux*>(attn),
        static_cast<const T*>(go), acc32, static_cast<Aux*>(gloc), static_cast<Aux*>(gattn),
        pr.S, pr.M, pr.D, pr.Lq, pr.L, pr.P, pr.total_pairs);
    ++g_last_launches, ++g_total_launches;
    rc = static_cast<int>(cudaGetLastError());
  }
  if (rc != 0) return rc;
  if ((n_value & 7) == 0 && (pr.M * pr.D) % 8 == 0) {
    const size_t n8 = n_value / 8;
    const int grid = static_cast<int>(std::min<size_t>((n8 + 255) / 256, static_cast<size_t>(sms) * 16));
    msda_round_scratch_kernel<T><<<grid, 256, 0, st>>>(acc32, static_cast<T*>(gv), n8, static_cast<uint32_t>(pr.M * pr.D / 8),
                                                      static_cast<uint32_t>(pr.grad_stride()), pr.gate, pr.gate_want);
  } else {
    const int grid = static_cast<int>(std::min<size_t>((n_value + 255) / 256, static_cast<size_t>(sms) * 16));
    msda_round_scratch_any_kernel<T><<<grid, 256, 0, st>>>(acc32, static_cast<T*>(gv), n_value);
  }
  ++g_last_launches, ++g_total_launches;
  return static_cast<int>(cudaGetLastError());
}

// 16-bit values, scaled fp16 buckets (+ sparse levels added directly): zero + max|grad_out| (one launch), backward, sum + round.
template <typename T>
static int run_bwd_f16_accum(const Problem& pr, const void* value, const int64_t* shapes, const int64_t* lsi,
                             const void* loc, const void* attn, const void* go, void* gv, void* gloc, void* gattn,
                             uint32_t* ctrl, __half* acc16, bool zero_ctrl, int flags, cudaStream_t st) {
  const int sms = device_sm_count();
  const int gstride = pr.grad_stride();
  // tiled kernels: one reduction per (destination row, tile) reaches the accumulator -- at most a few hundred adds per
  // element on the coarsest level of a pyramid -- so a single copy per level (no buckets) keeps the fp16 error at ~2e-3
  const bool tiled = tiled_supported<T>(pr) && pr.gate == nullptr;
  const bool hybrid = !tiled && hybrid_supported<T>(pr) && pr.gate == nullptr;
  const int depth = tiled ? 0xffff
                          : accum_depth(flags) | (hybrid ? std::min(g_hybrid_split.load(std::memory_order_relaxed), 0x7fff) << 16 : 0);
  // a level can only be sparse (4*Lq*P <= H_l*W_l) if 4*Lq*P <= S: decided here, the levels themselves on the device
  const bool sparse_direct = !hybrid && !(flags & MSDA_BWD_NO_SPARSE_DIRECT) && kSparseFactor * pr.Lq * pr.P <= pr.S;
  // Direct kernels: the zero pass also reduces max|grad_out| into ctrl[0], which must then be clear before the launch
  // (a memset node; the kernel cannot order its own clearing against other CTAs' atomicMax).  Tiled kernels: the dots
  // kernel produces the maximum, the zero pass clears the control block itself when it sits right in front of the
  // accumulator (i.e. without the guard's counters in between).
  cudaError_t e = cudaSuccess;
  if (!tiled && zero_ctrl) {
    e = cudaMemsetAsync(ctrl, 0, kF16CtrlBytes, st);
    if (e != cudaSuccess) return static_cast<int>(e);
    zero_ctrl = false;
  }
  const size_t go_n8 = static_cast<size_t>(pr.N) * pr.Lq * pr.M * pr.D / 8;     // D is a multiple of 16 here
  uint4* zero_base = zero_ctrl ? reinterpret_cast<uint4*>(ctrl) : reinterpret_cast<uint4*>(acc16);
  msda_zero_f16_buckets_kernel<T><<<sms * 8, 256, 0, st>>>(zero_base, zero_ctrl ? kF16CtrlBytes / 16 : 0, shapes, lsi,
                                                          sparse_direct ? static_cast<uint16_t*>(gv) : nullptr,
                                                          pr.N, pr.S, pr.M, pr.D, pr.Lq, pr.L, pr.P, depth, gstride,
                                                          tiled ? nullptr : static_cast<const T*>(go), go_n8, ctrl,
                                                          pr.gate, pr.gate_want);
  ++g_last_launches, ++g_total_launches;
  e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  int rc;
  if (tiled) {
    rc = launch_bwd_tiled<T>(pr, value, shapes, lsi, loc, attn, go, acc16, ctrl, gloc, gattn, depth, st);
  } else {
    void* gvd = sparse_direct ? gv : nullptr;                      // sparse levels add straight into grad_value
    Problem prh = pr;
    prh.hybrid = hybrid;
    switch (pr.D) {
      case 16: rc = launch_bwd_vec<T, 16>(prh, value, shapes, lsi, loc, attn, go, nullptr, acc16, gvd, gstride, ctrl, gloc, gattn, true, depth, st); break;
      case 32: rc = launch_bwd_vec<T, 32>(prh, value, shapes, lsi, loc, attn, go, nullptr, acc16, gvd, gstride, ctrl, gloc, gattn, true, depth, st); break;
      case 64: rc = launch_bwd_vec<T, 64>(prh, value, shapes, lsi, loc, attn, go, nullptr, acc16, gvd, gstride, ctrl, gloc, gattn, true, depth, st); break;
      default: rc = launch_bwd_vec<T, 128>(prh, value, shapes, lsi, loc, attn, go, nullptr, acc16, gvd, gstride, ctrl, gloc, gattn, true, depth, st); break;
    }
    if (rc == 0 && hybrid) rc = launch_bwd_scatter<T>(pr, shapes, lsi, loc, attn, go, acc16, ctrl, depth, st);
  }
  if (rc != 0) return rc;
  const size_t n8_img = static_cast<size_t>(pr.S) * pr.M * pr.D / 8;
  const size_t want = (n8_img + 255) / 256, cap = std::max<size_t>(1, (static_cast<size_t>(sms) * 16) / static_cast<size_t>(pr.N));
  const dim3 grid(static_cast<unsigned>(std::min(want, cap)), static_cast<unsigned>(std::min(pr.N, 65535)));
  msda_round_f16_buckets_kernel<T><<<grid, 256, 0, st>>>(acc16, static_cast<T*>(gv), shapes, lsi, ctrl, pr.N, pr.S,
                                                        pr.M, pr.D, pr.Lq, pr.L, pr.P, depth, sparse_direct && !tiled ? 1 : 0,
                                                        gstride, pr.gate, pr.gate_want);
  ++g_last_launches, ++g_total_launches;
  return static_cast<int>(cudaGetLastError());
}

template <typename T>
MSDA_LAUNCHER int launch_bwd(const Problem& pr_in, const void* value, const int64_t* shapes, const int64_t* lsi,
                      const void* loc, const void* attn, const void* go, void* gv, void* gloc, void* gattn,
                      void* scratch, int flags, cudaStream_t st) {
  Problem pr = pr_in;
  const size_t n_value = static_cast<size_t>(pr.N) * pr.S * pr.M * pr.D;
  constexpr bool k16 = sizeof(T) == 2;
  const bool use16 = k16 && !(flags & MSDA_BWD_GRAD_VALUE_FP32_ACCUM) && vec_supported<T>(pr);
  // strided grad_value: the vector kernels only (fp32 values directly, 16-bit values in either accumulation mode)
  if (pr.strided() && (!vec_supported<T>(pr) || std::is_same<T, double>::value)) return MSDA_ERR_BAD_STRIDE;

  if constexpr (k16) {
    if (!use16)      // fp32 accumulation asked for, or a head dim only the compatibility kernels cover
      return run_bwd_fp32_accum<T>(pr, value, shapes, lsi, loc, attn, go, gv, gloc, gattn, static_cast<float*>(scratch),
                                   accum_depth(flags), false, st);
    const ScratchLayout sl = scratch_layout_16(pr.N, pr.S, pr.M, pr.D, pr.Lq, pr.L, pr.P, flags);
    uint32_t* ctrl = static_cast<uint32_t*>(scratch);
    char* payload = static_cast<char*>(scratch) + sl.payload_off;
    // the density pass reads sampling locations: the fused pre-op (raw offsets) and the tiled kernels go unguarded
    const bool guard = sl.guard && pr.ref == nullptr && !tiled_supported<T>(pr) && !hybrid_supported<T>(pr);
    if (!guard) {
      if (sl.counters_bytes != 0) {       // the accumulator does not follow the control block directly: clear it separately
        const cudaError_t e0 = cudaMemsetAsync(scratch, 0, kF16CtrlBytes, st);
        if (e0 != cudaSuccess) return static_cast<int>(e0);
      }
      return run_bwd_f16_accum<T>(pr, value, shapes, lsi, loc, attn, go, gv, gloc, gattn, ctrl,
                                  reinterpret_cast<__half*>(payload), sl.counters_bytes == 0, flags, st);
    }
    // ---- cluster guard: count, then both pipelines; only the one ctrl[1] selects does any work ----
    cudaError_t e = cudaMemsetAsync(scratch, 0, sl.payload_off, st);          // control block + counters
    if (e != cudaSuccess) return static_cast<int>(e);
    const long long points = static_cast<long long>(pr.N) * pr.Lq * pr.M * pr.L * pr.P;
    const bool sparse_direct = !(flags & MSDA_BWD_NO_SPARSE_DIRECT) && kSparseFactor * pr.Lq * pr.P <= pr.S;
    const int dgrid = static_cast<int>(std::min<long long>((points + 255) / 256, static_cast<long long>(device_sm_count()) * 8));
    msda_cluster_density_kernel<<<dgrid, 256, 0, st>>>(static_cast<const float*>(loc), shapes, lsi,
                                                      reinterpret_cast<uint32_t*>(static_cast<char*>(scratch) + sl.counters_off),
                                                      ctrl, points, sl.cells, pr.M, pr.Lq, pr.L, pr.P, accum_depth(flags),
                                                      sparse_direct ? 1 : 0);
    ++g_last_launches, ++g_total_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
    pr.gate = ctrl + 1;
    pr.gate_want = 0;
    int rc = run_bwd_f16_accum<T>(pr, value, shapes, lsi, loc, attn, go, gv, gloc, gattn, ctrl,
                                  reinterpret_cast<__half*>(payload), false, flags, st);
    if (rc != 0) return rc;
    pr.gate_want = 1;
    return run_bwd_fp32_accum<T>(pr, value, shapes, lsi, loc, attn, go, gv, gloc, gattn, reinterpret_cast<float*>(payload),
                                 accum_depth(flags), true, st);
  } else {
    // fp32 / fp64 values: reductions go straight into grad_value
    cudaError_t e;
    const int gstride = pr.grad_stride();
    if (gstride == pr.M * pr.D)
      e = cudaMemsetAsync(gv, 0, n_value * sizeof(T), st);
    else
      e = cudaMemset2DAsync(gv, static_cast<size_t>(gstride) * sizeof(T), 0, static_cast<size_t>(pr.M) * pr.D * sizeof(T),
                            static_cast<size_t>(pr.N) * pr.S, st);
    if (e != cudaSuccess) return static_cast<int>(e);
    if constexpr (!std::is_same<T, double>::value) {
      if (vec_supported<T>(pr)) {
        float* gv32 = static_cast<float*>(gv);
        const int depth = accum_depth(flags);
        switch (pr.D) {
          case 16: return launch_bwd_vec<T, 16>(pr, value, shapes, lsi, loc, attn, go, gv32, nullptr, nullptr, gstride, nullptr, gloc, gattn, false, depth, st);
          case 32: return launch_bwd_vec<T, 32>(pr, value, shapes, lsi, loc, attn, go, gv32, nullptr, nullptr, gstride, nullptr, gloc, gattn, false, depth, st);
          case 64: return launch_bwd_vec<T, 64>(pr, value, shapes, lsi, loc, attn, go, gv32, nullptr, nullptr, gstride, nullptr, gloc, gattn, false, depth, st);
          default: return launch_bwd_vec<T, 128>(pr, value, shapes, lsi, loc, attn, go, gv32, nullptr, nullptr, gstride, nullptr, gloc, gattn, false, depth, st);
        }
      }
    }
    if (pr.ref) return MSDA_ERR_FUSED_UNSUPPORTED;
    using Aux = typename Traits<T>::Aux;
    const int grid = (pr.total_pairs + kWarps - 1) / kWarps;
    msda_bwd_any_kernel<T, T><<<grid, kThreads, 0, st>>>(
        static_cast<const T*>(value), shapes, lsi, static_cast<const Aux*>(loc), static_cast<const Aux*>(attn),
        static_cast<const T*>(go), static_cast<T*>(gv), static_cast<Aux*>(gloc), static_cast<Aux*>(gattn),
        pr.S, pr.M, pr.D, pr.Lq, pr.L, pr.P, pr.total_pairs);
    ++g_last_launches, ++g_total_launches;
    return static_cast<int>(cudaGetLastError());
  }
}

#ifdef MSDA_SPLIT_BUILD
#define MSDA_LAUNCHERS_OF(KW, T)                                                                                      \
  KW template int launch_fwd<T>(const Problem&, const void*, const int64_t*, const int64_t*, const void*, const void*, \
                                void*, cudaStream_t);                                                                 \
  KW template int launch_bwd<T>(const Problem&, const void*, const int64_t*, const int64_t*, const void*, const void*, \
                                const void*, void*, void*, void*, void*, int, cudaStream_t);
#ifdef MSDA_TU_ABI
MSDA_LAUNCHERS_OF(extern, float)
MSDA_LAUNCHERS_OF(extern, double)
MSDA_LAUNCHERS_OF(extern, __nv_bfloat16)
MSDA_LAUNCHERS_OF(extern, __half)
#elif MSDA_TU_DTYPE == 0
MSDA_LAUNCHERS_OF(, float)
MSDA_LAUNCHERS_OF(, double)
#elif MSDA_TU_DTYPE == 1
MSDA_LAUNCHERS_OF(, __nv_bfloat16)
#elif MSDA_TU_DTYPE == 2
MSDA_LAUNCHERS_OF(, __half)
#endif
#endif

}  // namespace msda

#if !defined(MSDA_SPLIT_BUILD) || defined(MSDA_TU_ABI)
// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace msda;

extern "C" int msda_abi_version(void) { return 5; }

extern "C" const char* msda_error_string(int code) {
  switch (code) {
    case MSDA_OK: return "success";
    case MSDA_ERR_NULL_POINTER: return "null pointer argument";
    case MSDA_ERR_BAD_SHAPE: return "invalid shape (non-positive dimension, more than 32 levels, or too many pairs)";
    case MSDA_ERR_BAD_DTYPE: return "unsupported value dtype";
    case MSDA_ERR_MISALIGNED: return "buffer is not 16-byte aligned";
    case MSDA_ERR_IM2COL_STEP: return "batch size must be divisible by min(batch, im2col_step)";
    case MSDA_ERR_SCRATCH_TOO_SMALL: return "scratch buffer missing or too small";
    case MSDA_ERR_BAD_STRIDE:
      return "pixel stride must be >= M*D elements and a multiple of 16 bytes; strided tensors need the vector kernels "
             "(head dim 16/32/64/128, not float64)";
    case MSDA_ERR_FUSED_UNSUPPORTED:
      return "fused pre-op needs float32/bfloat16/float16 values with head dim 16/32/64/128 and reference points of width 2 or 4";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown msda error";
}

extern "C" int msda_set_tiled_mode(int mode) {
  const int prev = tiled_mode();
  g_tiled_mode.store(mode < 0 || mode > 2 ? 0 : mode);
  return prev;
}

extern "C" int msda_set_hybrid_split(int adds) {
  const int prev = g_hybrid_split.load();
  g_hybrid_split.store(adds < 0 ? 0 : adds);
  return prev;
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

static size_t dtype_bytes(int value_dtype) {
  return value_dtype == MSDA_F64 ? 8 : value_dtype == MSDA_F32 ? 4 : 2;
}

extern "C" int msda_forward_strided(const void* value, long long value_pixel_stride,
                                    const int64_t* spatial_shapes, const int64_t* level_start_index,
                                    const void* sampling_loc, const void* attn_weight, void* output,
                                    int N, int S, int M, int D, int Lq, int L, int P,
                                    int value_dtype, int im2col_step, void* stream) {
  g_last_launches = 0;
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !output) return MSDA_ERR_NULL_POINTER;
  Problem pr{N, S, M, D, Lq, L, P, 0};
  int v = validate(pr, value_dtype, im2col_step);
  if (v != MSDA_OK) return v;
  pr.vps = value_pixel_stride;
  v = validate_strides(pr, dtype_bytes(value_dtype));
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

extern "C" int msda_forward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                            const void* sampling_loc, const void* attn_weight, void* output,
                            int N, int S, int M, int D, int Lq, int L, int P,
                            int value_dtype, int im2col_step, void* stream) {
  return msda_forward_strided(value, 0, spatial_shapes, level_start_index, sampling_loc, attn_weight, output,
                              N, S, M, D, Lq, L, P, value_dtype, im2col_step, stream);
}

extern "C" size_t msda_backward_scratch_bytes(int N, int S, int M, int D, int Lq, int L, int P, int value_dtype, int flags) {
  if (value_dtype != MSDA_BF16 && value_dtype != MSDA_F16) return 0;
  if (N <= 0 || S <= 0 || M <= 0 || D <= 0 || Lq <= 0 || L <= 0 || P <= 0) return 0;
  const bool vec = vec_supported_rt(S, M, D, L, P, 2, static_cast<long long>(M) * D, false);
  if (!(flags & MSDA_BWD_GRAD_VALUE_FP32_ACCUM) && vec) return scratch_layout_16(N, S, M, D, Lq, L, P, flags).total;
  return static_cast<size_t>(N) * S * M * D * sizeof(float);
}

extern "C" int msda_backward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                             const void* sampling_loc, const void* attn_weight, const void* grad_output,
                             void* grad_value, void* grad_sampling_loc, void* grad_attn_weight,
                             void* scratch, size_t scratch_bytes,
                             int N, int S, int M, int D, int Lq, int L, int P,
                             int value_dtype, int im2col_step, int flags, void* stream) {
  return msda_backward_strided(value, 0, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                               grad_value, 0, grad_sampling_loc, grad_attn_weight, scratch, scratch_bytes,
                               N, S, M, D, Lq, L, P, value_dtype, im2col_step, flags, stream);
}

extern "C" int msda_backward_strided(const void* value, long long value_pixel_stride,
                                     const int64_t* spatial_shapes, const int64_t* level_start_index,
                                     const void* sampling_loc, const void* attn_weight, const void* grad_output,
                                     void* grad_value, long long grad_value_pixel_stride,
                                     void* grad_sampling_loc, void* grad_attn_weight,
                                     void* scratch, size_t scratch_bytes,
                                     int N, int S, int M, int D, int Lq, int L, int P,
                                     int value_dtype, int im2col_step, int flags, void* stream) {
  g_last_launches = 0;
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !grad_output ||
      !grad_value || !grad_sampling_loc || !grad_attn_weight)
    return MSDA_ERR_NULL_POINTER;
  Problem pr{N, S, M, D, Lq, L, P, 0};
  int v = validate(pr, value_dtype, im2col_step);
  if (v != MSDA_OK) return v;
  pr.vps = value_pixel_stride;
  pr.gps = grad_value_pixel_stride;
  v = validate_strides(pr, dtype_bytes(value_dtype));
  if (v != MSDA_OK) return v;
  pr.total_pairs = N * Lq * M;
  if (!aligned16(value) || !aligned16(sampling_loc) || !aligned16(attn_weight) || !aligned16(grad_output) ||
      !aligned16(grad_value) || !aligned16(grad_sampling_loc) || !aligned16(grad_attn_weight) || !aligned16(scratch))
    return MSDA_ERR_MISALIGNED;
  const size_t need = msda_backward_scratch_bytes(N, S, M, D, Lq, L, P, value_dtype, flags);
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

// ---- fused pre-op entry points ---------------------------------------------------
extern "C" int msda_fused_supported(int D, int value_dtype) {
  const bool d_ok = D == 16 || D == 32 || D == 64 || D == 128;
  return (d_ok && (value_dtype == MSDA_F32 || value_dtype == MSDA_BF16 || value_dtype == MSDA_F16)) ? 1 : 0;
}


static int fused_validate(const Problem& pr, int value_dtype, int ref_dim, int aux_dtype) {
  if (ref_dim != 2 && ref_dim != 4) return MSDA_ERR_FUSED_UNSUPPORTED;
  if (aux_dtype != MSDA_F32 && !(aux_dtype == value_dtype && (value_dtype == MSDA_BF16 || value_dtype == MSDA_F16)))
    return MSDA_ERR_FUSED_UNSUPPORTED;
  if (!msda_fused_supported(pr.D, value_dtype)) return MSDA_ERR_FUSED_UNSUPPORTED;
  if (static_cast<unsigned long long>(pr.S) * pr.M * pr.D * sizeof(float) >= (1ull << 32)) return MSDA_ERR_FUSED_UNSUPPORTED;
  return MSDA_OK;
}

extern "C" int msda_fused_forward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                  const void* reference_points, int ref_dim, const void* sampling_offsets,
                                  const void* attn_logits, void* output,
                                  int N, int S, int M, int D, int Lq, int L, int P,
                                  int value_dtype, int aux_dtype, int im2col_step, void* stream) {
  g_last_launches = 0;
  if (!value || !spatial_shapes || !level_start_index || !reference_points || !sampling_offsets || !attn_logits || !output)
    return MSDA_ERR_NULL_POINTER;
  Problem pr{N, S, M, D, Lq, L, P, 0};
  int v = validate(pr, value_dtype, im2col_step);
  if (v != MSDA_OK) return v;
  v = fused_validate(pr, value_dtype, ref_dim, aux_dtype);
  if (v != MSDA_OK) return v;
  pr.total_pairs = N * Lq * M;
  pr.ref = static_cast<const float*>(reference_points);
  pr.R = ref_dim;
  pr.aux16 = aux_dtype != MSDA_F32;
  if (!aligned16(value) || !aligned16(sampling_offsets) || !aligned16(attn_logits) || !aligned16(output) ||
      !aligned16(reference_points))
    return MSDA_ERR_MISALIGNED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (value_dtype) {
    case MSDA_F32: return launch_fwd<float>(pr, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, output, st);
    case MSDA_BF16: return launch_fwd<__nv_bfloat16>(pr, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, output, st);
    case MSDA_F16: return launch_fwd<__half>(pr, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, output, st);
  }
  return MSDA_ERR_BAD_DTYPE;
}

extern "C" int msda_fused_backward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                   const void* reference_points, int ref_dim, const void* sampling_offsets,
                                   const void* attn_logits, const void* grad_output,
                                   void* grad_value, void* grad_sampling_offsets, void* grad_attn_logits,
                                   void* scratch, size_t scratch_bytes,
                                   int N, int S, int M, int D, int Lq, int L, int P,
                                   int value_dtype, int aux_dtype, int im2col_step, int flags, void* stream) {
  g_last_launches = 0;
  if (!value || !spatial_shapes || !level_start_index || !reference_points || !sampling_offsets || !attn_logits ||
      !grad_output || !grad_value || !grad_sampling_offsets || !grad_attn_logits)
    return MSDA_ERR_NULL_POINTER;
  Problem pr{N, S, M, D, Lq, L, P, 0};
  int v = validate(pr, value_dtype, im2col_step);
  if (v != MSDA_OK) return v;
  v = fused_validate(pr, value_dtype, ref_dim, aux_dtype);
  if (v != MSDA_OK) return v;
  pr.total_pairs = N * Lq * M;
  pr.ref = static_cast<const float*>(reference_points);
  pr.R = ref_dim;
  pr.aux16 = aux_dtype != MSDA_F32;
  if (!aligned16(value) || !aligned16(sampling_offsets) || !aligned16(attn_logits) || !aligned16(grad_output) ||
      !aligned16(grad_value) || !aligned16(grad_sampling_offsets) || !aligned16(grad_attn_logits) ||
      !aligned16(scratch) || !aligned16(reference_points))
    return MSDA_ERR_MISALIGNED;
  const size_t need = msda_backward_scratch_bytes(N, S, M, D, Lq, L, P, value_dtype, flags);
  if (need > 0 && (!scratch || scratch_bytes < need)) return MSDA_ERR_SCRATCH_TOO_SMALL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (value_dtype) {
    case MSDA_F32:
      return launch_bwd<float>(pr, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, grad_output,
                               grad_value, grad_sampling_offsets, grad_attn_logits, scratch, flags, st);
    case MSDA_BF16:
      return launch_bwd<__nv_bfloat16>(pr, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, grad_output,
                                       grad_value, grad_sampling_offsets, grad_attn_logits, scratch, flags, st);
    case MSDA_F16:
      return launch_bwd<__half>(pr, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, grad_output,
                                grad_value, grad_sampling_offsets, grad_attn_logits, scratch, flags, st);
  }
  return MSDA_ERR_BAD_DTYPE;
}
#endif  // C ABI translation unit
