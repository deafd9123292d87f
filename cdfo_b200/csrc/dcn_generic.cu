// Deformable convolution forward (DCNv1 / DCNv2), reference semantics, any shape.
//
// Semantics follow the reference op (fp32 sampling arithmetic, corner-wise zero padding, inside test
// h_im > -1 && w_im > -1 && h_im < H && w_im < W): ops/dcn/src/deform_conv_cuda_kernel.cu:467-496,570-632 (v2),
// :83-114,190-242 (v1); contraction + bias: ops/dcn/src/deform_conv_cuda.cpp:486-564, :151-258.
//
// Design (not the reference's im2col+SGEMM): one CTA owns 64 output pixels x 64 output channels of one sample
// and one weight group; the sampled columns for a chunk of K = (channel, tap) rows live only in shared memory
// (never in HBM) and are contracted against a shared-memory weight chunk with fp32 FMAs.  This is the
// catch-all path (any C, groups, dg, stride, dilation, dtype); the model's hot shape goes to dcn_sm100.cu.
#include "cdfo_common.cuh"

namespace cdfo {

constexpr int kTP = 64;    // output pixels per CTA
constexpr int kCOB = 64;   // output channels per CTA
constexpr int kKCH = 72;   // K rows (channel x tap) per shared-memory chunk
constexpr int kThreads = 256;

struct DcnGenericParams {
  const void *x, *offset, *mask, *weight, *bias;
  void *y;
  int B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, groups, dg, Ho, Wo;
};

// Corner-wise zero-padded bilinear sample of one fp32-converted plane.
template <typename T>
__device__ __forceinline__ float dcn_bilinear(const T *__restrict__ plane, int H, int W, float h, float w) {
  const int h_low = (int)floorf(h), w_low = (int)floorf(w);
  const int h_high = h_low + 1, w_high = w_low + 1;
  const float lh = h - (float)h_low, lw = w - (float)w_low;
  const float hh = 1.f - lh, hw = 1.f - lw;
  float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
  if (h_low >= 0 && w_low >= 0) v1 = to_f32(plane[(size_t)h_low * W + w_low]);
  if (h_low >= 0 && w_high <= W - 1) v2 = to_f32(plane[(size_t)h_low * W + w_high]);
  if (h_high <= H - 1 && w_low >= 0) v3 = to_f32(plane[(size_t)h_high * W + w_low]);
  if (h_high <= H - 1 && w_high <= W - 1) v4 = to_f32(plane[(size_t)h_high * W + w_high]);
  const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) dcn_generic_kernel(DcnGenericParams p) {
  __shared__ float col[kKCH][kTP];
  __shared__ float wsm[kCOB][kKCH + 1];

  const T *__restrict__ x = (const T *)p.x;
  const T *__restrict__ offset = (const T *)p.offset;
  const T *__restrict__ mask = (const T *)p.mask;
  const T *__restrict__ weight = (const T *)p.weight;
  const T *__restrict__ bias = (const T *)p.bias;
  T *__restrict__ y = (T *)p.y;

  const int KK = p.kh * p.kw;
  const int Cg = p.C / p.groups, Cog = p.Co / p.groups, cpdg = p.C / p.dg;
  const int P = p.Ho * p.Wo;
  const int co_chunks = ceil_div(Cog, kCOB);
  const int wg = blockIdx.y / co_chunks;
  const int co0 = (blockIdx.y % co_chunks) * kCOB;  // within the weight group
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * kTP;
  const int tid = threadIdx.x;
  const int Ktot = Cg * KK;

  const int my_p = tid % kTP, my_cq = tid / kTP;  // my_cq in [0,4)
  float acc[kCOB / 4];
#pragma unroll
  for (int m = 0; m < kCOB / 4; ++m) acc[m] = 0.f;

  for (int k0 = 0; k0 < Ktot; k0 += kKCH) {
    const int rows = min(kKCH, Ktot - k0);
    // ---- phase A: sampled (and modulated) columns of this K chunk, shared memory only ----
    for (int e = tid; e < rows * kTP; e += kThreads) {
      const int r = e / kTP, pp = e % kTP;
      const int pix = p0 + pp;
      float val = 0.f;
      if (pix < P) {
        const int k = k0 + r;
        const int c = wg * Cg + k / KK, tap = k % KK;
        const int i = tap / p.kw, j = tap % p.kw;
        const int g = c / cpdg;
        const int ho = pix / p.Wo, wo = pix % p.Wo;
        const T *off_g = offset + ((size_t)b * p.dg + g) * 2 * KK * P;
        const float off_h = to_f32(off_g[(size_t)(2 * tap) * P + pix]);
        const float off_w = to_f32(off_g[(size_t)(2 * tap + 1) * P + pix]);
        const float h_im = (float)(ho * p.sh - p.ph + i * p.dh) + off_h;
        const float w_im = (float)(wo * p.sw - p.pw + j * p.dw) + off_w;
        if (h_im > -1.f && w_im > -1.f && h_im < (float)p.H && w_im < (float)p.W)
          val = dcn_bilinear(x + ((size_t)b * p.C + c) * p.H * p.W, p.H, p.W, h_im, w_im);
        if (mask) val *= to_f32(mask[(((size_t)b * p.dg + g) * KK + tap) * P + pix]);
      }
      col[r][pp] = val;
    }
    // ---- weight chunk ----
    for (int e = tid; e < kCOB * rows; e += kThreads) {
      const int co = e / rows, r = e % rows;
      float wv = 0.f;
      if (co0 + co < Cog) wv = to_f32(weight[((size_t)(wg * Cog + co0 + co)) * Ktot + k0 + r]);
      wsm[co][r] = wv;
    }
    __syncthreads();
    // ---- phase B: contraction ----
    for (int r = 0; r < rows; ++r) {
      const float v = col[r][my_p];
#pragma unroll
      for (int m = 0; m < kCOB / 4; ++m) acc[m] = fmaf(wsm[my_cq + 4 * m][r], v, acc[m]);
    }
    __syncthreads();
  }

  const int pix = p0 + my_p;
  if (pix < P) {
#pragma unroll
    for (int m = 0; m < kCOB / 4; ++m) {
      const int co = co0 + my_cq + 4 * m;
      if (co < Cog) {
        const int cog = wg * Cog + co;
        float v = acc[m];
        if (bias) v += to_f32(bias[cog]);
        y[((size_t)b * p.Co + cog) * P + pix] = from_f32<T>(v);
      }
    }
  }
}

__global__ void dcn_sample_index_kernel(const float *__restrict__ offset, int32_t *__restrict__ idx, int B,
                                        int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, int dh,
                                        int dw, int dg, int Ho, int Wo) {
  const int KK = kh * kw;
  const size_t P = (size_t)Ho * Wo;
  const size_t total = (size_t)B * dg * KK * P;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = e % P;
    const int tap = (int)((e / P) % KK);
    const size_t bg = e / P / KK;  // b*dg + g
    const int ho = (int)(pix / Wo), wo = (int)(pix % Wo);
    const int i = tap / kw, j = tap % kw;
    const float off_h = offset[(bg * 2 * KK + 2 * tap) * P + pix];
    const float off_w = offset[(bg * 2 * KK + 2 * tap + 1) * P + pix];
    const float h_im = (float)(ho * sh - ph + i * dh) + off_h;
    const float w_im = (float)(wo * sw - pw + j * dw) + off_w;
    idx[e * 2 + 0] = (int32_t)floorf(h_im);
    idx[e * 2 + 1] = (int32_t)floorf(w_im);
  }
}

}  // namespace cdfo

using namespace cdfo;

extern "C" int cdfo_dcn_fwd(const void *x, const void *offset, const void *mask, const void *weight,
                            const void *bias, void *y, int B, int C, int H, int W, int Co, int kh, int kw,
                            int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                            int groups, int dg, int dtype, void *stream) {
  CDFO_REQUIRE(x && offset && weight && y, CDFO_ERR_NULL, "cdfo_dcn_fwd: x/offset/weight/y must be non-NULL");
  CDFO_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && Co > 0 && kh > 0 && kw > 0, CDFO_ERR_SHAPE,
               "cdfo_dcn_fwd: non-positive size");
  CDFO_REQUIRE(stride_h > 0 && stride_w > 0 && dil_h > 0 && dil_w > 0 && pad_h >= 0 && pad_w >= 0, CDFO_ERR_SHAPE,
               "cdfo_dcn_fwd: stride/dilation must be > 0, padding >= 0");
  CDFO_REQUIRE(groups > 0 && dg > 0 && C % groups == 0 && Co % groups == 0, CDFO_ERR_SHAPE,
               "cdfo_dcn_fwd: channels (%d in, %d out) not divisible by groups %d", C, Co, groups);
  CDFO_REQUIRE(C % dg == 0, CDFO_ERR_SHAPE, "cdfo_dcn_fwd: input channels must divide deformable group size");
  const int Ho = conv_out_size(H, pad_h, dil_h, kh, stride_h), Wo = conv_out_size(W, pad_w, dil_w, kw, stride_w);
  CDFO_REQUIRE(Ho >= 1 && Wo >= 1, CDFO_ERR_SHAPE, "cdfo_dcn_fwd: output size is too small (%d x %d)", Ho, Wo);
  if (B == 0) return CDFO_OK;
  DcnGenericParams p{x, offset, mask, weight, bias, y, B, C, H, W, Co, kh, kw, stride_h, stride_w,
                     pad_h, pad_w, dil_h, dil_w, groups, dg, Ho, Wo};
  const int Cog = Co / groups;
  dim3 grid(ceil_div(Ho * Wo, kTP), groups * ceil_div(Cog, kCOB), B);
  CDFO_REQUIRE(grid.y <= 65535 && grid.z <= 65535, CDFO_ERR_UNSUPPORTED, "cdfo_dcn_fwd: grid too large");
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case CDFO_F32: dcn_generic_kernel<float><<<grid, kThreads, 0, s>>>(p); break;
    case CDFO_F16: dcn_generic_kernel<__half><<<grid, kThreads, 0, s>>>(p); break;
    case CDFO_BF16: dcn_generic_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(p); break;
    default: return fail(CDFO_ERR_UNSUPPORTED, "cdfo_dcn_fwd: unknown dtype %d", dtype);
  }
  return check_launch("cdfo_dcn_fwd");
}

extern "C" int cdfo_dcn_sample_index(const float *offset, int32_t *idx, int B, int H, int W, int kh, int kw,
                                     int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                                     int dg, void *stream) {
  CDFO_REQUIRE(offset && idx, CDFO_ERR_NULL, "cdfo_dcn_sample_index: NULL pointer");
  const int Ho = conv_out_size(H, pad_h, dil_h, kh, stride_h), Wo = conv_out_size(W, pad_w, dil_w, kw, stride_w);
  CDFO_REQUIRE(B > 0 && dg > 0 && Ho >= 1 && Wo >= 1, CDFO_ERR_SHAPE, "cdfo_dcn_sample_index: bad shape");
  const size_t total = (size_t)B * dg * kh * kw * Ho * Wo;
  const int blocks = (int)min((size_t)kNumSMs * 8, (total + 255) / 256);
  dcn_sample_index_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(offset, idx, B, H, W, kh, kw, stride_h,
                                                                   stride_w, pad_h, pad_w, dil_h, dil_w, dg, Ho, Wo);
  return check_launch("cdfo_dcn_sample_index");
}
