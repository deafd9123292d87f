// Two memory-bound pieces of the feature extraction (the rest of it is still cuDNN):
//   LayerNorm over channels per pixel ("WithBias"), arch/SIDECVSR_our.py:1169-1198 (used at :1441-1475)
//   the 3x3 depthwise convolution qkv_dwconv of the MDTA attention, arch:1545-1576
// NCHW, fp32 or bf16 storage, fp32 arithmetic.  ATen ran the LayerNorm as ~10 elementwise / reduce launches per call and the
// depthwise convolution at < 10 % of the HBM roofline.
#include "cdfo_common.cuh"

namespace cdfo {
namespace feat {

template <typename T> __device__ __forceinline__ float ldf(const T *p);
template <> __device__ __forceinline__ float ldf<float>(const float *p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16 *p) {
  return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short *>(p)) << 16);
}
template <typename T> __device__ __forceinline__ void stf(T *p, float v);
template <> __device__ __forceinline__ void stf<float>(float *p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// y[b][c][p] = (x - mean_c) * rsqrt(var_c + eps) * gamma[c] + beta[c]; one thread per pixel, C <= 64 values in registers
template <typename T, int C>
__global__ void __launch_bounds__(128) layernorm_c_kernel(const T *__restrict__ x, const float *__restrict__ gamma,
                                                          const float *__restrict__ beta, T *__restrict__ y, int HW, float eps) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (p >= HW) return;
  const T *xp = x + (size_t)b * C * HW + p;
  float v[C], s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) { v[c] = ldf(xp + (size_t)c * HW); s += v[c]; }
  const float mu = s * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) { const float d = v[c] - mu; q = fmaf(d, d, q); }
  const float r = rsqrtf(q * (1.f / C) + eps);
  T *yp = y + (size_t)b * C * HW + p;
#pragma unroll
  for (int c = 0; c < C; ++c) stf(yp + (size_t)c * HW, (v[c] - mu) * r * __ldg(gamma + c) + __ldg(beta + c));
}

// depthwise 3x3, stride 1, padding 1, no bias: thread per pixel of one (b, c) plane
template <typename T>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const T *__restrict__ x, const float *__restrict__ w, T *__restrict__ y, int C,
                                                        int H, int W) {
  const int HW = H * W, p = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y, b = blockIdx.z;
  if (p >= HW) return;
  const int h = p / W, wq = p - h * W;
  const T *xp = x + ((size_t)b * C + c) * HW;
  float k[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int hh = h + i - 1, ww = wq + j - 1;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) acc = fmaf(k[i * 3 + j], ldf(xp + hh * W + ww), acc);
    }
  stf(y + ((size_t)b * C + c) * HW + p, acc);
}

// Vectorised bf16 variant (W % 8 == 0): one thread = 8 consecutive pixels of a row (one 16-byte load per input row + the two
// horizontal neighbours), 9x fewer load instructions than the scalar kernel (which is bound by LSU issue, not by HBM).
__global__ void __launch_bounds__(256) dwconv3x3_bf16x8_kernel(const __nv_bfloat16 *__restrict__ x, const float *__restrict__ w,
                                                               __nv_bfloat16 *__restrict__ y, int C, int H, int W) {
  const int W8 = W >> 3, n = H * W8;
  const int q = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y, b = blockIdx.z;
  if (q >= n) return;
  const int h = q / W8, w0 = (q - h * W8) * 8;
  const __nv_bfloat16 *xp = x + ((size_t)b * C + c) * H * W;
  float k[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int hh = h + i - 1;
    if (hh < 0 || hh >= H) continue;
    const __nv_bfloat16 *row = xp + (size_t)hh * W + w0;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row));
    float r[10];
    r[0] = w0 > 0 ? ldf(row - 1) : 0.f;
    r[9] = w0 + 8 < W ? ldf(row + 8) : 0.f;
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      r[1 + 2 * j] = __uint_as_float(u[j] << 16);
      r[2 + 2 * j] = __uint_as_float(u[j] & 0xffff0000u);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(k[i * 3 + 2], r[j + 2], fmaf(k[i * 3 + 1], r[j + 1], fmaf(k[i * 3], r[j], acc[j])));
  }
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 t = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
    o[j] = *reinterpret_cast<uint32_t *>(&t);
  }
  *reinterpret_cast<uint4 *>(y + ((size_t)b * C + c) * H * W + (size_t)h * W + w0) = make_uint4(o[0], o[1], o[2], o[3]);
}

}  // namespace feat
}  // namespace cdfo

using namespace cdfo;

extern "C" int cdfo_layernorm_c_fwd(const void *x, const float *gamma, const float *beta, void *y, int B, int C, int H, int W, float eps,
                                    int dtype, void *stream) {
  CDFO_REQUIRE(x && gamma && beta && y, CDFO_ERR_NULL, "cdfo_layernorm_c_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, CDFO_ERR_SHAPE, "cdfo_layernorm_c_fwd: bad shape");
  CDFO_REQUIRE(C == 64, CDFO_ERR_UNSUPPORTED, "cdfo_layernorm_c_fwd: 64 channels (got %d)", C);
  const int HW = H * W;
  dim3 grid(ceil_div(HW, 128), B);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CDFO_F32) feat::layernorm_c_kernel<float, 64><<<grid, 128, 0, s>>>((const float *)x, gamma, beta, (float *)y, HW, eps);
  else if (dtype == CDFO_BF16)
    feat::layernorm_c_kernel<__nv_bfloat16, 64><<<grid, 128, 0, s>>>((const __nv_bfloat16 *)x, gamma, beta, (__nv_bfloat16 *)y, HW, eps);
  else return fail(CDFO_ERR_UNSUPPORTED, "cdfo_layernorm_c_fwd: dtype must be fp32 or bf16");
  return check_launch("cdfo_layernorm_c_fwd");
}

extern "C" int cdfo_dwconv3x3_fwd(const void *x, const float *w, void *y, int B, int C, int H, int W, int dtype, void *stream) {
  CDFO_REQUIRE(x && w && y, CDFO_ERR_NULL, "cdfo_dwconv3x3_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && C > 0 && C <= 65535 && H > 0 && W > 0, CDFO_ERR_SHAPE, "cdfo_dwconv3x3_fwd: bad shape");
  dim3 grid(ceil_div(H * W, 256), C, B);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CDFO_F32) feat::dwconv3x3_kernel<float><<<grid, 256, 0, s>>>((const float *)x, w, (float *)y, C, H, W);
  else if (dtype == CDFO_BF16 && W % 8 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0)
    feat::dwconv3x3_bf16x8_kernel<<<dim3(ceil_div(H * (W / 8), 256), C, B), 256, 0, s>>>((const __nv_bfloat16 *)x, w, (__nv_bfloat16 *)y, C, H, W);
  else if (dtype == CDFO_BF16)
    feat::dwconv3x3_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16 *)x, w, (__nv_bfloat16 *)y, C, H, W);
  else return fail(CDFO_ERR_UNSUPPORTED, "cdfo_dwconv3x3_fwd: dtype must be fp32 or bf16");
  return check_launch("cdfo_dwconv3x3_fwd");
}
