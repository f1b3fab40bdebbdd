/* TEST INFRASTRUCTURE ONLY — scalar C restatement of the multi-scale deformable attention path.
 * PARITY STATUS: unpinned against the reference's own tests (it has none); see oracle/__init__.py.
 *
 * Follows the published algorithm of the upstream CUDA kernels
 * (IDEA-Research/MaskDINO maskdino/modeling/pixel_decoder/ops/src/cuda/ms_deform_im2col_cuda.cuh:
 *  ms_deform_attn_im2col_bilinear, ms_deform_attn_col2im_bilinear), which the reference reaches through
 * build_model(cfg) (/root/reference/training/maskdino/train_full.py:308).  That checkout is not vendored,
 * so this is a restatement, not a build of the reference.  double precision, OpenMP over (b, q) for the
 * forward; the backward is serial over queries inside one (b, m) plane so that grad_value needs no atomics.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/libmsda_oracle.so oracle/msda_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline int inside(double h, double w, int H, int W) { return h > -1 && w > -1 && h < H && w < W; }

/* value (N,S,M,D), shapes (L,2) rows (H,W), lsi (L), loc (N,Lq,M,L,P,2) (x,y), attn (N,Lq,M,L,P), out (N,Lq,M*D) */
void msda_oracle_forward(const double* value, const int64_t* shapes, const int64_t* lsi, const double* loc,
                         const double* attn, double* out, int N, int S, int M, int D, int Lq, int L, int P) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < N; ++b)
    for (int q = 0; q < Lq; ++q)
      for (int m = 0; m < M; ++m) {
        double* o = out + (((size_t)b * Lq + q) * M + m) * D;
        for (int d = 0; d < D; ++d) o[d] = 0.0;
        for (int l = 0; l < L; ++l) {
          const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
          const double* vl = value + ((size_t)b * S + lsi[l]) * M * D + (size_t)m * D;
          for (int p = 0; p < P; ++p) {
            const size_t idx = ((((size_t)b * Lq + q) * M + m) * L + l) * P + p;
            const double x = loc[2 * idx], y = loc[2 * idx + 1], a = attn[idx];
            const double h_im = y * H - 0.5, w_im = x * W - 0.5;
            if (!inside(h_im, w_im, H, W)) continue;
            const int h0 = (int)floor(h_im), w0 = (int)floor(w_im);
            const double lh = h_im - h0, lw = w_im - w0, hh = 1 - lh, hw = 1 - lw;
            const double wt[4] = {hh * hw, hh * lw, lh * hw, lh * lw};
            const int hy[4] = {h0, h0, h0 + 1, h0 + 1}, wx[4] = {w0, w0 + 1, w0, w0 + 1};
            for (int c = 0; c < 4; ++c) {
              if (hy[c] < 0 || hy[c] >= H || wx[c] < 0 || wx[c] >= W) continue;
              const double* v = vl + ((size_t)hy[c] * W + wx[c]) * M * D;
              for (int d = 0; d < D; ++d) o[d] += a * wt[c] * v[d];
            }
          }
        }
      }
}

void msda_oracle_backward(const double* value, const int64_t* shapes, const int64_t* lsi, const double* loc,
                          const double* attn, const double* grad_out, double* grad_value, double* grad_loc,
                          double* grad_attn, int N, int S, int M, int D, int Lq, int L, int P) {
  memset(grad_value, 0, sizeof(double) * (size_t)N * S * M * D);
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < N; ++b)
    for (int m = 0; m < M; ++m)
      for (int q = 0; q < Lq; ++q) {
        const double* go = grad_out + (((size_t)b * Lq + q) * M + m) * D;
        for (int l = 0; l < L; ++l) {
          const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
          const size_t lvl = ((size_t)b * S + lsi[l]) * M * D + (size_t)m * D;
          for (int p = 0; p < P; ++p) {
            const size_t idx = ((((size_t)b * Lq + q) * M + m) * L + l) * P + p;
            const double x = loc[2 * idx], y = loc[2 * idx + 1], a = attn[idx];
            const double h_im = y * H - 0.5, w_im = x * W - 0.5;
            grad_attn[idx] = 0.0; grad_loc[2 * idx] = 0.0; grad_loc[2 * idx + 1] = 0.0;
            if (!inside(h_im, w_im, H, W)) continue;
            const int h0 = (int)floor(h_im), w0 = (int)floor(w_im);
            const double lh = h_im - h0, lw = w_im - w0, hh = 1 - lh, hw = 1 - lw;
            const double wt[4] = {hh * hw, hh * lw, lh * hw, lh * lw};
            const double dh[4] = {-hw, -lw, hw, lw}, dw[4] = {-hh, hh, -lh, lh};
            const int hy[4] = {h0, h0, h0 + 1, h0 + 1}, wx[4] = {w0, w0 + 1, w0, w0 + 1};
            double g_a = 0, g_h = 0, g_w = 0;
            for (int c = 0; c < 4; ++c) {
              if (hy[c] < 0 || hy[c] >= H || wx[c] < 0 || wx[c] >= W) continue;
              const size_t off = lvl + ((size_t)hy[c] * W + wx[c]) * M * D;
              double dot = 0;
              for (int d = 0; d < D; ++d) {
                dot += value[off + d] * go[d];
                grad_value[off + d] += wt[c] * a * go[d];
              }
              g_a += wt[c] * dot; g_h += dh[c] * dot; g_w += dw[c] * dot;
            }
            grad_attn[idx] = g_a;
            grad_loc[2 * idx] = W * a * g_w;
            grad_loc[2 * idx + 1] = H * a * g_h;
          }
        }
      }
}
