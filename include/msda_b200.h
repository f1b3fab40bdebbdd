/*
 * msda_b200.h — C ABI of the B200-native multi-scale deformable attention operator.
 *
 * Drop-in boundary.  These entry points replace the two functions the upstream
 * `MultiScaleDeformableAttention` extension exports (IDEA-Research/MaskDINO,
 * maskdino/modeling/pixel_decoder/ops/src/vision.cpp: `ms_deform_attn_forward`,
 * `ms_deform_attn_backward`; dispatch in src/ms_deform_attn.h; host wrappers in
 * src/cuda/ms_deform_attn_cuda.cu).  That checkout is not vendored by the reference repository:
 * it is put on sys.path at /root/reference/training/maskdino/train_full.py:15-16 and reached through
 * `from maskdino import add_maskdino_config` (train_full.py:28) + `build_model(cfg)`
 * (train_full.py:308, evaluate.py:109, visualize.py:251).
 *
 * Conventions (all entry points)
 *   - plain pointers and sizes only; no torch / pybind types.
 *   - every buffer is allocated, owned and kept alive by the caller; the library never allocates,
 *     frees or synchronises.  All work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - all device buffers are contiguous, on the current CUDA device, 16-byte aligned.
 *   - return value: 0 = success; < 0 = argument validation failure (MSDA_ERR_*); > 0 = cudaError_t
 *     of the failing runtime call / launch.  No printf-and-continue.
 *   - re-entrant; no global mutable state except a one-time, thread-safe kernel attribute set-up.
 *
 * Tensor layouts (row-major, identical to upstream)
 *   value              (N, S, M, D)          value dtype          S = sum_l H_l*W_l
 *   spatial_shapes     (L, 2)  int64         rows are (H_l, W_l)      [device memory]
 *   level_start_index  (L,)    int64         prefix sums of H_l*W_l   [device memory]
 *   sampling_loc       (N, Lq, M, L, P, 2)   float32 (float64 when value is float64); last dim (x, y)
 *   attn_weight        (N, Lq, M, L, P)      same dtype as sampling_loc
 *   output             (N, Lq, M*D)          value dtype
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* value dtype codes */
enum { MSDA_F32 = 0, MSDA_F64 = 1, MSDA_BF16 = 2, MSDA_F16 = 3 };

/* validation errors (negative so they never collide with cudaError_t) */
enum {
  MSDA_OK = 0,
  MSDA_ERR_NULL_POINTER = -1,
  MSDA_ERR_BAD_SHAPE = -2,       /* a dimension is <= 0 or too large (L > 32, overflow of 2^31 pairs) */
  MSDA_ERR_BAD_DTYPE = -3,
  MSDA_ERR_MISALIGNED = -4,      /* a buffer is not 16-byte aligned */
  MSDA_ERR_IM2COL_STEP = -5,     /* N % min(N, im2col_step) != 0 (upstream assert kept) */
  MSDA_ERR_SCRATCH_TOO_SMALL = -6,
  MSDA_ERR_FUSED_UNSUPPORTED = -7, /* fused pre-op: head dim not in {16,32,64,128}, float64 values, or ref_dim not 2/4 */
  MSDA_ERR_BAD_STRIDE = -8        /* *_strided entry points: stride < M*D, not a multiple of 16 bytes, or a path without stride support */
};

/* backward flags */
enum {
  /* Default.  fp32 / fp64 values: red.global.add straight into grad_value.  16-bit values (vector kernels):
   * grad_value is accumulated with packed red.global.add.noftz.v4.f16x2 into a *scaled, bucketed fp16* buffer
   * in `scratch` — half the reduction bytes of fp32 accumulation, and the SM->L2 reduction path is what bounds
   * the backward kernel — then the buckets are summed in fp32, unscaled and rounded once.
   *   scale  : power of two derived on the device from max|grad_output| and Lq such that no partial sum can
   *            overflow fp16 for any input;
   *   buckets: level l keeps ceil(ceil(Lq*P / (H_l*W_l)) / depth) private copies and query q adds into copy
   *            q mod K_l, so that an fp16 accumulator receives ~depth adds (default 32; bits 8..23 of `flags`
   *            override it).  Measured error of grad_value at the 1024^2 encoder shape: ~3e-3 of its max, the
   *            same as fp32 accumulation followed by bf16 rounding. */
  MSDA_BWD_DEFAULT = 0,
  /* 16-bit values: accumulate in an fp32 scratch with red.global.add.v4.f32 instead (2x the reduction bytes). */
  MSDA_BWD_GRAD_VALUE_FP32_ACCUM = 2,
  /* Default mode, sparse levels.  A level with 4*Lq*P <= H_l*W_l (decoder cross-attention: a few hundred queries
   * against 10^4 pixels; at most one expected add per element) owns no accumulator rows: its contributions are
   * added, unscaled, with packed 16-bit reductions in the value dtype straight into the zeroed grad_value rows, so the
   * dense zero / sum / round passes over the accumulator (most of a decoder layer's backward) are paid for the coarse
   * levels only.  This flag switches that off (every level goes through the fp16 buckets). */
  MSDA_BWD_NO_SPARSE_DIRECT = 4,
  /* Default mode, cluster guard.  Both 16-bit mechanisms above size their precision for evenly spread sampling.  For
   * calls of at most 4 Mi sampling points (decoder cross-attention; the dense encoder shapes are far larger and spread
   * by construction) msda_backward first counts the sampling points per (run of 4 pixels, head) on the device; if a cell
   * holds more than its level tolerates (32 for a sparse level, 512 per bucket otherwise -- e.g. every query looking at
   * the same object) it raises a flag in the control block of `scratch`.  The 16-bit pipeline and an fp32-accumulation
   * pipeline are both enqueued, and every kernel returns at once unless the flag selects its pipeline: no host
   * synchronisation, same results as MSDA_BWD_GRAD_VALUE_FP32_ACCUM when the flag is up.  Not applied to the fused
   * pre-op entry points (they never see sampling locations).  This flag switches the guard off. */
  MSDA_BWD_NO_CLUSTER_GUARD = 8
};
#define MSDA_BWD_ACCUM_DEPTH(depth) (((depth) & 0xffff) << 8)

/* Library ABI version (bumped on any signature change). */
int msda_abi_version(void);

/* Human-readable text for a return code of this library (static storage). */
const char* msda_error_string(int code);

/*
 * Forward.  Replaces upstream `ms_deform_attn_forward(value, spatial_shapes, level_start_index,
 * sampling_loc, attn_weight, im2col_step) -> output` (vision.cpp / ms_deform_attn_cuda.cu
 * `ms_deform_attn_cuda_forward`).  `output` is fully overwritten.  `im2col_step` is validated the
 * way upstream does (N % min(N, im2col_step) == 0) and otherwise ignored: one launch covers the
 * whole batch, which is semantically identical to upstream's per-chunk loop.
 */
int msda_forward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                 const void* sampling_loc, const void* attn_weight, void* output,
                 int N, int S, int M, int D, int Lq, int L, int P,
                 int value_dtype, int im2col_step, void* stream);

/* Bytes of scratch `msda_backward` needs for this problem (0 if none).  The bound depends only on sizes the
 * host knows (the int64 shape tensor stays on the device). */
size_t msda_backward_scratch_bytes(int N, int S, int M, int D, int Lq, int L, int P, int value_dtype, int flags);

/*
 * Backward.  Replaces upstream `ms_deform_attn_backward(value, spatial_shapes, level_start_index,
 * sampling_loc, attn_weight, grad_output, im2col_step) -> [grad_value, grad_sampling_loc,
 * grad_attn_weight]` (`ms_deform_attn_cuda_backward`).  grad_output is (N, Lq, M*D) in the value
 * dtype.  grad_sampling_loc and grad_attn_weight are fully overwritten (no pre-zeroing needed).
 * The buffer that receives the reductions (grad_value for fp32/fp64, `scratch` for 16-bit values) is zeroed by
 * this call (cudaMemsetAsync on `stream`); `scratch` holds msda_backward_scratch_bytes() bytes (NULL when 0).
 */
int msda_backward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                  const void* sampling_loc, const void* attn_weight, const void* grad_output,
                  void* grad_value, void* grad_sampling_loc, void* grad_attn_weight,
                  void* scratch, size_t scratch_bytes,
                  int N, int S, int M, int D, int Lq, int L, int P,
                  int value_dtype, int im2col_step, int flags, void* stream);

/*
 * Strided variants (decoder `value_proj` reuse).  The nine decoder layers of MaskDINO each run
 * their own `value_proj` Linear over the same encoder memory; one GEMM with the nine weights stacked produces
 * (N, S, layers, M*D), in which layer i's `value` is the view [:, :, i] -- dense per pixel row, but with
 * `value_pixel_stride = layers*M*D` elements between neighbouring pixels (and S*value_pixel_stride between images).
 * These entry points read such a view in place and write `grad_value` into the matching view of one shared
 * (N, S, layers, M*D) gradient buffer, so the stacked projection needs no per-layer copies in either direction.
 * Strides are in elements, >= M*D, multiples of 16 bytes; 0 means dense.  Pointers address element [0, 0, 0, 0] of
 * the view.  Vector kernels only (head dim 16/32/64/128, not float64); otherwise MSDA_ERR_BAD_STRIDE.  `msda_forward` / `msda_backward` are these with both strides 0.
 */
int msda_forward_strided(const void* value, long long value_pixel_stride,
                         const int64_t* spatial_shapes, const int64_t* level_start_index,
                         const void* sampling_loc, const void* attn_weight, void* output,
                         int N, int S, int M, int D, int Lq, int L, int P,
                         int value_dtype, int im2col_step, void* stream);

int msda_backward_strided(const void* value, long long value_pixel_stride,
                          const int64_t* spatial_shapes, const int64_t* level_start_index,
                          const void* sampling_loc, const void* attn_weight, const void* grad_output,
                          void* grad_value, long long grad_value_pixel_stride,
                          void* grad_sampling_loc, void* grad_attn_weight,
                          void* scratch, size_t scratch_bytes,
                          int N, int S, int M, int D, int Lq, int L, int P,
                          int value_dtype, int im2col_step, int flags, void* stream);

/*
 * Tiled kernels for the dense call site (csrc/msda_tiled.cuh) -- opt-in.  When every value pixel is also a query
 * (Lq == S: the pixel-decoder encoder layers), values are 16-bit and the head dim is 32, msda_forward / msda_backward can
 * keep per-tile windows of `value` in shared memory (cp.async, zero-filled outside the level) and sum grad_value per
 * destination row in registers (counting sort by destination, no shared-memory read-modify-write) before one packed
 * fp16 reduction per row and tile leaves the SM.  Results do not depend on the mode (points that leave their window take a
 * global-memory path).  Mode 0 (default; environment MSDA_B200_TILED=0|1 is read on first use) runs the direct kernels,
 * mode 1 the tiled ones.  Measured at the 1024^2 encoder shape the tiled kernels send 8x fewer reductions to the L2 but
 * are issue / latency bound and slower than the direct kernels, hence opt-in.  Returns the previous
 * mode.
 */
int msda_set_tiled_mode(int mode);

/*
 * Mode 2 of msda_set_tiled_mode (MSDA_B200_TILED=2): hybrid backward for the same call site, forward unchanged.  The
 * direct backward kernel computes grad_sampling_loc / grad_attn_weight and the grad_value reductions of the fine levels;
 * the coarse levels -- those that expect more than `adds` corner rows per pixel row, ceil(Lq*P / (H_l*W_l)) > adds,
 * where pre-reducing inside the SM removes most of the reductions the L2 would otherwise resolve -- get their
 * grad_value from the sorting kernel of the tiled backward.  msda_set_hybrid_split sets `adds` (default 8; 0 switches
 * the hybrid path off) and returns the previous value.
 */
int msda_set_hybrid_split(int adds);

/* Number of kernel launches (not memsets) the last forward/backward call on this thread enqueued;
 * used by bench.py to report `gpu_launches`. */
int msda_last_launch_count(void);

/* Process-wide, monotonically increasing count of kernel launches enqueued by this library (all threads,
 * including autograd's backward thread). */
long long msda_total_launch_count(void);

/*
 * Optional per-launch timing of the dominant kernels, for bench.py's roofline leg.  While enabled
 * (process-wide) every forward / backward call records a CUDA event pair on `stream` immediately around
 * its main kernel (kind MSDA_KERNEL_FORWARD / MSDA_KERNEL_BACKWARD; memsets and the fp32->16-bit rounding
 * pass are outside the pair).  msda_profile_collect() waits for the recorded events, writes up to
 * max_records (duration in ms, kind) pairs in call order, frees them and returns how many it wrote.
 * These two calls are the only ones in the library that create events or block the host.
 */
enum { MSDA_KERNEL_FORWARD = 1, MSDA_KERNEL_BACKWARD = 2,
       /* tiled backward (dense call site): the two kernels that replace MSDA_KERNEL_BACKWARD */
       MSDA_KERNEL_BACKWARD_DOTS = 3, MSDA_KERNEL_BACKWARD_SCATTER = 4 };
int msda_profile_enable(int on);
int msda_profile_collect(float* ms, int* kinds, int max_records);

/*
 * Fused pre-op (opt-in, `MSDeformAttnFunction.apply` is unchanged).
 *
 * These two entry points fold into the sampling kernels the arithmetic upstream `MSDeformAttn.forward`
 * (maskdino/modeling/pixel_decoder/ops/modules/ms_deform_attn.py) performs between its Linears and
 * `MSDeformAttnFunction.apply`:
 *     attention_weights = softmax(attn_logits over L*P)
 *     sampling_locations = reference_points[:, :, None, :, None, :] + sampling_offsets / (W_l, H_l)      ref_dim == 2
 *                        = reference_points[..., :2] + sampling_offsets / P * reference_points[..., 2:] * 0.5   ref_dim == 4
 * with the same fp32 rounding order, so that sampling_locations / attention_weights and their gradients are
 * never written to or read from HBM (12 of the 20 forward and 24 of the 36 backward algorithmic bytes per
 * sampled point at the encoder shape).
 *
 *   reference_points   (N, Lq, L, ref_dim)   float32
 *   sampling_offsets   (N, Lq, M, L, P, 2)   aux dtype   raw output of the sampling_offsets Linear
 *   attn_logits        (N, Lq, M, L*P)       aux dtype   raw output of the attention_weights Linear
 *
 * aux_dtype is MSDA_F32, or — for 16-bit values — the value dtype itself: inside a torch.autocast region the two
 * Linears hand over bfloat16 / float16, which the kernels then read (and whose gradients they write) directly instead
 * of through fp32 copies; all arithmetic stays fp32.
 *
 * The backward returns the gradients of sampling_offsets and attn_logits (softmax and offset chain rules applied
 * in shared memory); reference_points receives no gradient here (callers that need one compose the plain operator).
 * Scratch: msda_backward_scratch_bytes().  Vector kernels only: msda_fused_supported() tells whether (D, dtype)
 * is covered; otherwise MSDA_ERR_FUSED_UNSUPPORTED is returned and nothing is enqueued.
 */
int msda_fused_supported(int D, int value_dtype);

int msda_fused_forward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                       const void* reference_points, int ref_dim, const void* sampling_offsets,
                       const void* attn_logits, void* output,
                       int N, int S, int M, int D, int Lq, int L, int P,
                       int value_dtype, int aux_dtype, int im2col_step, void* stream);

int msda_fused_backward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                        const void* reference_points, int ref_dim, const void* sampling_offsets,
                        const void* attn_logits, const void* grad_output,
                        void* grad_value, void* grad_sampling_offsets, void* grad_attn_logits,
                        void* scratch, size_t scratch_bytes,
                        int N, int S, int M, int D, int Lq, int L, int P,
                        int value_dtype, int aux_dtype, int im2col_step, int flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */
