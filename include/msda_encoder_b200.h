/*
 * msda_encoder_b200.h — C ABI of the memory-bound glue kernels of one pixel-decoder encoder layer (SURVEY.md §8f rank 3:
 * "the step either side of the op").
 *
 * Upstream `MSDeformAttnTransformerEncoderLayer.forward` / `forward_ffn`
 * (IDEA-Research/MaskDINO maskdino/modeling/pixel_decoder/maskdino_encoder.py, reached by the reference only through
 * `build_model(cfg)`, /root/reference/training/maskdino/train_full.py:308) is, around the MSDeformAttn call,
 *
 *     q    = src + pos                                  -> msda_enc_add_cast            (bf16 copy for the three Linears)
 *     src  = norm1(src + dropout1(self_attn(q, ..)))    -> msda_enc_add_layernorm_*     (residual add + LayerNorm, fp32 + bf16 out)
 *     src2 = linear2(dropout2(relu(linear1(src))))      -> GEMMs stay library GEMMs; relu backward + bias gradient:
 *                                                          msda_enc_relu_bwd_colsum, bias gradients: msda_enc_colsum
 *     src  = norm2(src + dropout3(src2))                -> msda_enc_add_layernorm_*
 *
 * Under bf16 autocast stock torch runs each of these as separate elementwise / reduction kernels with fp32<->bf16 casts in
 * between (55 % of the cfg3 training step, profiles/train_step_breakdown_r01.txt); the entry points below do each of them
 * in one pass over HBM.  Conventions as in msda_b200.h: plain pointers, caller-owned buffers (16-byte aligned,
 * contiguous), work enqueued on `stream`, 0 / negative MSDA_ERR_* / positive cudaError_t returned, no allocation, no sync.
 * The 16-bit type is bfloat16 throughout (the autocast dtype of the path).  C must be a multiple of 128, at most 1024,
 * for the LayerNorm kernels; any C for the column sums (vector path when C is a multiple of 8).
 */
#ifndef MSDA_ENCODER_B200_H_
#define MSDA_ENCODER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* out16[i] = bf16(a[i] + b[i]);  a, b float32 (n elements, n % 8 == 0).  `with_pos_embed` + the autocast cast. */
int msda_enc_add_cast(const float* a, const float* b, void* out16, size_t n, void* stream);

/* Residual add + LayerNorm over the last dimension (eps as torch.nn.LayerNorm):
 *     y = LayerNorm(x + float(delta16)) * gamma + beta,    rows x C
 * x float32; delta16 bfloat16 or NULL (plain LayerNorm); y float32; y16 = bf16(y) (NULL to skip); mean / rstd float32
 * per row (saved for the backward). */
int msda_enc_add_layernorm_forward(const float* x, const void* delta16, const float* gamma, const float* beta,
                                   float* y, void* y16, float* mean, float* rstd,
                                   long long rows, int C, float eps, void* stream);

/* Bytes of `partials` scratch the backward needs for dgamma / dbeta. */
size_t msda_enc_add_layernorm_backward_scratch_bytes(int C);

/* Backward of the above.  The incoming gradient of y is gy (float32, may be NULL) + float(gy16) (bfloat16, may be NULL);
 * x / delta16 / mean / rstd / gamma as in the forward.  Writes dx (float32, gradient of x), ddelta16 (bfloat16, the same
 * values rounded once; NULL to skip), dgamma / dbeta (float32, fully overwritten; deterministic two-stage reduction). */
int msda_enc_add_layernorm_backward(const float* gy, const void* gy16, const float* x, const void* delta16,
                                    const float* mean, const float* rstd, const float* gamma,
                                    float* dx, void* ddelta16, float* dgamma, float* dbeta,
                                    void* partials, size_t partials_bytes,
                                    long long rows, int C, void* stream);

/* Bytes of scratch the two column-sum entry points need. */
size_t msda_enc_colsum_scratch_bytes(int C);

/* out[c] = sum over the selected rows of g16[row, c]  (bias gradient of a Linear; level_embed gradient).
 * g16 is (batch, rows_per_batch, C) bfloat16; rows row_begin <= r < row_end of every batch entry are summed
 * (0, rows_per_batch = everything).  out float32 [C], fully overwritten; deterministic. */
int msda_enc_colsum(const void* g16, float* out, void* scratch, size_t scratch_bytes,
                    long long batch, long long rows_per_batch, long long row_begin, long long row_end, int C, void* stream);

/* ReLU backward fused with the bias gradient of the Linear in front of it:
 *     g16[i] = h16[i] > 0 ? g16[i] : 0   (in place),   out[c] = sum_rows g16[row, c]
 * g16, h16 (rows, C) bfloat16; out float32 [C]. */
int msda_enc_relu_bwd_colsum(void* g16, const void* h16, float* out, void* scratch, size_t scratch_bytes,
                             long long rows, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_ENCODER_B200_H_ */
