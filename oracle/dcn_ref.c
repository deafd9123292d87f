/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C, fp32) of the reference's
 * deformable-convolution forward.  Not product code: only tests/, the smoke
 * check and bench.py's cpu_baseline / reference arm may load this library.
 *
 * Restates (read, not copied; the reference is a scalar CUDA grid-stride
 * kernel + per-sample cuBLAS SGEMM, this is a direct nested loop):
 *   - corner-wise zero-padded bilinear sample
 *       ops/dcn/src/deform_conv_cuda_kernel.cu:467-496 (v2), :83-114 (v1)
 *   - DCNv2 column value  val*mask, inside-test (h>-1 && w>-1 && h<H && w<W)
 *       ops/dcn/src/deform_conv_cuda_kernel.cu:570-632
 *   - DCNv1 column value
 *       ops/dcn/src/deform_conv_cuda_kernel.cu:190-242
 *   - weight contraction per sample / per weight-group and bias add
 *       ops/dcn/src/deform_conv_cuda.cpp:486-564 (v2), :151-258 (v1)
 *
 * Pinned by: ops/dcn/simple_check.py:11-22 (DCNv1 known-answer vector) in
 * tests/test_oracle_dcn.py, and against torchvision.ops.deform_conv2d CPU
 * (the op the model's DCN alignment calls, arch/SIDECVSR_our.py:3352).
 *
 * Layouts (all contiguous NCHW fp32, exactly the reference's):
 *   x      [B, C, H, W]
 *   offset [B, dg*2*kh*kw, Ho, Wo]   channel 2*(g*kh*kw + i*kw + j) + {0:dy, 1:dx}
 *   mask   [B, dg*kh*kw,   Ho, Wo]   (NULL => DCNv1, mask == 1)
 *   weight [Co, C/groups, kh, kw]
 *   bias   [Co] or NULL
 *   y      [B, Co, Ho, Wo]
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float bilinear_zero_pad(const float *plane, int H, int W, float h, float w) {
  int h_low = (int)floorf(h);
  int w_low = (int)floorf(w);
  int h_high = h_low + 1;
  int w_high = w_low + 1;
  float lh = h - (float)h_low;
  float lw = w - (float)w_low;
  float hh = 1.0f - lh, hw = 1.0f - lw;
  float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
  if (h_low >= 0 && w_low >= 0) v1 = plane[(size_t)h_low * W + w_low];
  if (h_low >= 0 && w_high <= W - 1) v2 = plane[(size_t)h_low * W + w_high];
  if (h_high <= H - 1 && w_low >= 0) v3 = plane[(size_t)h_high * W + w_low];
  if (h_high <= H - 1 && w_high <= W - 1) v4 = plane[(size_t)h_high * W + w_high];
  float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  /* same association as the reference: ((w1*v1 + w2*v2) + w3*v3) + w4*v4 */
  volatile float t1 = w1 * v1, t2 = w2 * v2, t3 = w3 * v3, t4 = w4 * v4;
  return ((t1 + t2) + t3) + t4;
}

/* Output size exactly as deform_conv.py:98-111 / :174-183. */
static inline int out_size(int in, int pad, int dil, int k, int stride) {
  return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;
}

/*
 * Returns 0 on success, negative on a shape error (the reference raises).
 * Also optionally records the integer sample indices floor(h_im), floor(w_im)
 * for every (b, g, tap, ho, wo) into idx_out [B, dg*kh*kw, Ho, Wo, 2] (int32)
 * so that the CUDA path's MV-to-offset indexing can be checked bit-exactly.
 */
int oracle_dcn_forward(const float *x, const float *offset, const float *mask,
                       const float *weight, const float *bias, float *y,
                       int B, int C, int H, int W, int Co, int kh, int kw,
                       int stride_h, int stride_w, int pad_h, int pad_w,
                       int dil_h, int dil_w, int groups, int dg, int32_t *idx_out) {
  if (groups <= 0 || dg <= 0 || C % groups || Co % groups || C % dg) return -1;
  const int Ho = out_size(H, pad_h, dil_h, kh, stride_h);
  const int Wo = out_size(W, pad_w, dil_w, kw, stride_w);
  if (Ho < 1 || Wo < 1) return -2;
  const int Cg = C / groups;    /* input channels per weight group */
  const int Cog = Co / groups;  /* output channels per weight group */
  const int cpdg = C / dg;      /* channels per deformable group */
  const int KK = kh * kw;
  const size_t P = (size_t)Ho * Wo;

#pragma omp parallel
  {
    float *col = (float *)malloc(sizeof(float) * (size_t)C * KK);
#pragma omp for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
      for (int ho = 0; ho < Ho; ++ho) {
        for (int wo = 0; wo < Wo; ++wo) {
          const int h_in = ho * stride_h - pad_h;
          const int w_in = wo * stride_w - pad_w;
          /* columns for this output pixel: col[c*KK + tap] */
          for (int c = 0; c < C; ++c) {
            const int g = c / cpdg;
            const float *plane = x + ((size_t)b * C + c) * H * W;
            const float *off_g = offset + ((size_t)b * dg + g) * 2 * KK * P;
            const float *msk_g = mask ? mask + ((size_t)b * dg + g) * KK * P : NULL;
            for (int i = 0; i < kh; ++i) {
              for (int j = 0; j < kw; ++j) {
                const int tap = i * kw + j;
                const float off_h = off_g[(size_t)(2 * tap) * P + (size_t)ho * Wo + wo];
                const float off_w = off_g[(size_t)(2 * tap + 1) * P + (size_t)ho * Wo + wo];
                const float h_im = (float)(h_in + i * dil_h) + off_h;
                const float w_im = (float)(w_in + j * dil_w) + off_w;
                float val = 0.f;
                if (h_im > -1 && w_im > -1 && h_im < H && w_im < W)
                  val = bilinear_zero_pad(plane, H, W, h_im, w_im);
                if (msk_g) val = val * msk_g[(size_t)tap * P + (size_t)ho * Wo + wo];
                col[c * KK + tap] = val;
                if (idx_out && c % cpdg == 0) {
                  int32_t *o = idx_out + ((((size_t)b * dg + g) * KK + tap) * P + (size_t)ho * Wo + wo) * 2;
                  o[0] = (int32_t)floorf(h_im);
                  o[1] = (int32_t)floorf(w_im);
                }
              }
            }
          }
          for (int co = 0; co < Co; ++co) {
            const int wg = co / Cog;
            const float *wrow = weight + (size_t)co * Cg * KK;
            const float *crow = col + (size_t)wg * Cg * KK;
            float acc = 0.f;
            for (int k = 0; k < Cg * KK; ++k) acc += wrow[k] * crow[k];
            if (bias) acc += bias[co];
            y[(((size_t)b * Co + co) * Ho + ho) * Wo + wo] = acc;
          }
        }
      }
    }
    free(col);
  }
  return 0;
}

/* flow_warp restatement: arch/SIDECVSR_our.py:3068-3099 with F.grid_sample
 * (bilinear, zeros, align_corners=True).  Replays the normalise /
 * un-normalise round trip in fp32, one rounding per op:
 *   vx = (w + flow_x);  gx = 2.0f*vx / max(W-1,1) - 1.0f       (:3091-3094)
 *   ix = ((gx + 1) / 2) * (W-1)                                 (ATen grid_sampler_unnormalize, align_corners)
 * formula = 0 : ((g+1)/2)*(size-1)   -- ATen scalar / CUDA kernel
 * formula = 1 : (g+1)*((size-1)/2)   -- ATen vectorised CPU kernel
 * x [B,C,H,W], flow [B,H,W,2] (x,y), y [B,C,H,W]; idx_out [B,H,W,2] = (iy_nw, ix_nw) or NULL.
 */
int oracle_flow_warp(const float *x, const float *flow, float *y, int B, int C, int H, int W,
                     int formula, int32_t *idx_out) {
  const float dw = (float)(W - 1 > 1 ? W - 1 : 1);
  const float dh = (float)(H - 1 > 1 ? H - 1 : 1);
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int h = 0; h < H; ++h) {
      for (int w = 0; w < W; ++w) {
        const float *f = flow + (((size_t)b * H + h) * W + w) * 2;
        volatile float vx = (float)w + f[0];
        volatile float vy = (float)h + f[1];
        volatile float ax = 2.0f * vx; volatile float bx = ax / dw; volatile float gx = bx - 1.0f;
        volatile float ay = 2.0f * vy; volatile float by = ay / dh; volatile float gy = by - 1.0f;
        float ix, iy;
        if (formula == 0) {
          volatile float px = gx + 1.0f; volatile float qx = px / 2.0f; ix = qx * (float)(W - 1);
          volatile float py = gy + 1.0f; volatile float qy = py / 2.0f; iy = qy * (float)(H - 1);
        } else {
          volatile float px = gx + 1.0f; ix = px * ((float)(W - 1) / 2.0f);
          volatile float py = gy + 1.0f; iy = py * ((float)(H - 1) / 2.0f);
        }
        const float fx = floorf(ix), fy = floorf(iy);
        const int ix_nw = (int)fx, iy_nw = (int)fy;
        const int ix_ne = ix_nw + 1, iy_sw = iy_nw + 1;
        /* ATen weights: nw = (ix_se-ix)*(iy_se-iy) ... */
        const float ex = (float)ix_ne, ey = (float)iy_sw;
        volatile float tx1 = ex - ix, tx0 = ix - fx, ty1 = ey - iy, ty0 = iy - fy;
        const float nw = tx1 * ty1, ne = tx0 * ty1, sw = tx1 * ty0, se = tx0 * ty0;
        if (idx_out) {
          int32_t *o = idx_out + (((size_t)b * H + h) * W + w) * 2;
          o[0] = iy_nw; o[1] = ix_nw;
        }
        for (int c = 0; c < C; ++c) {
          const float *plane = x + ((size_t)b * C + c) * H * W;
          float acc = 0.f;
          if (iy_nw >= 0 && iy_nw < H && ix_nw >= 0 && ix_nw < W) acc += plane[(size_t)iy_nw * W + ix_nw] * nw;
          if (iy_nw >= 0 && iy_nw < H && ix_ne >= 0 && ix_ne < W) acc += plane[(size_t)iy_nw * W + ix_ne] * ne;
          if (iy_sw >= 0 && iy_sw < H && ix_nw >= 0 && ix_nw < W) acc += plane[(size_t)iy_sw * W + ix_nw] * sw;
          if (iy_sw >= 0 && iy_sw < H && ix_ne >= 0 && ix_ne < W) acc += plane[(size_t)iy_sw * W + ix_ne] * se;
          y[(((size_t)b * C + c) * H + h) * W + w] = acc;
        }
      }
    }
  }
  return 0;
}
