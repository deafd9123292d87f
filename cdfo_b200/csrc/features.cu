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


// Per-head Gram matrix and squared norms of the MDTA self-attention (Attention.forward, arch/SIDECVSR_our.py:1545-1576):
//   G[b][hd][i][j] = sum_p q[b][8 hd + i][p] k[b][8 hd + j][p],   nq[b][c] = sum_p q[b][c][p]^2,   nk likewise,
// q = channels [0, 64) and k = channels [64, 128) of the depthwise-convolved qkv tensor [B][Ctot][HW].  With them the whole attention is a
// 64 x 64 matrix per sample (softmax(G / (|q| |k|) T) folded with project_out), applied to v as one batched GEMM: q and k are read ONCE
// (the ATen chain reads / writes them ~10 times: two norms, two divisions, fp32 copies, a K = H*W skinny GEMM).
// A CTA reduces a pixel range; partial sums [B][parts][640] = 512 Gram entries + 64 + 64 are added by the caller in fixed order.
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int kN = 4;
  static __device__ __forceinline__ void load(const float *p, float *o) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int kN = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float *o) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
};

constexpr int kGramLd = 68;

template <typename T>
__global__ void __launch_bounds__(256) mdta_gram_kernel(const T *__restrict__ qk, float *__restrict__ partial, int Ctot, int HW,
                                                        int px_per_part, int vec_ok) {
  __shared__ __align__(16) float tile[128 * kGramLd];     // [128 channels][64 pixels]
  const int tid = threadIdx.x, b = blockIdx.y, part = blockIdx.x, parts = gridDim.x;
  const int p_begin = part * px_per_part, p_end = min(HW, p_begin + px_per_part);
  const T *src = qk + (size_t)b * Ctot * HW;
  // thread -> Gram entries e = tid and tid + 256: head = e / 64, i = (e / 8) % 8, j = e % 8
  const int e0 = tid, e1 = tid + 256;
  const float *q0 = tile + ((e0 >> 6) * 8 + ((e0 >> 3) & 7)) * kGramLd, *k0 = tile + (64 + (e0 >> 6) * 8 + (e0 & 7)) * kGramLd;
  const float *q1 = tile + ((e1 >> 6) * 8 + ((e1 >> 3) & 7)) * kGramLd, *k1 = tile + (64 + (e1 >> 6) * 8 + (e1 & 7)) * kGramLd;
  float g0 = 0.f, g1 = 0.f, nn = 0.f;
  constexpr int kN = Vec<T>::kN, kPerRow = 64 / kN, kIters = 128 * kPerRow / 256;
  for (int p0 = p_begin; p0 < p_end; p0 += 64) {
    const int npx = min(64, p_end - p0);
    __syncthreads();
    if (vec_ok) {      // all of a thread's vector loads are issued before the first store
      float r[kIters][kN];
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int e = tid + it * 256, c = e / kPerRow, q = (e % kPerRow) * kN;
        if (q < npx) Vec<T>::load(src + (size_t)c * HW + p0 + q, r[it]);
        else
#pragma unroll
          for (int i = 0; i < kN; ++i) r[it][i] = 0.f;
      }
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int e = tid + it * 256, c = e / kPerRow, q = (e % kPerRow) * kN;
#pragma unroll
        for (int i = 0; i < kN; i += 4) *reinterpret_cast<float4 *>(tile + c * kGramLd + q + i) = make_float4(r[it][i], r[it][i + 1], r[it][i + 2], r[it][i + 3]);
      }
    } else {
      for (int e = tid; e < 128 * 64; e += 256) {
        const int c = e >> 6, q = e & 63;
        tile[c * kGramLd + q] = q < npx ? ldf(src + (size_t)c * HW + p0 + q) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int q = 0; q < 64; q += 4) {
      const float4 a = *reinterpret_cast<const float4 *>(q0 + q), c = *reinterpret_cast<const float4 *>(k0 + q);
      const float4 d = *reinterpret_cast<const float4 *>(q1 + q), f = *reinterpret_cast<const float4 *>(k1 + q);
      g0 += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
      g1 += d.x * f.x + d.y * f.y + d.z * f.z + d.w * f.w;
    }
    if (tid < 128) {
      const float *row = tile + tid * kGramLd;
#pragma unroll 4
      for (int q = 0; q < 64; q += 4) {
        const float4 a = *reinterpret_cast<const float4 *>(row + q);
        nn += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
      }
    }
  }
  float *out = partial + ((size_t)b * parts + part) * 640;
  out[e0] = g0;
  out[e1] = g1;
  if (tid < 128) out[512 + tid] = nn;
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

extern "C" int cdfo_mdta_gram_fwd(const void *qk, float *partial, int B, int Ctot, int H, int W, int parts, int dtype, void *stream) {
  CDFO_REQUIRE(qk && partial, CDFO_ERR_NULL, "cdfo_mdta_gram_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && Ctot >= 128 && H > 0 && W > 0 && parts > 0, CDFO_ERR_SHAPE, "cdfo_mdta_gram_fwd: bad shape");
  const int HW = H * W;
  int per = ceil_div(ceil_div(HW, parts), 64) * 64;          // pixel ranges are multiples of the 64-pixel tile
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(parts, B);
  if (dtype == CDFO_F32) {
    const int vec = HW % 4 == 0 && ((uintptr_t)qk & 15) == 0;
    feat::mdta_gram_kernel<float><<<grid, 256, 0, s>>>((const float *)qk, partial, Ctot, HW, per, vec);
  } else if (dtype == CDFO_BF16) {
    const int vec = HW % 8 == 0 && ((uintptr_t)qk & 15) == 0;
    feat::mdta_gram_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16 *)qk, partial, Ctot, HW, per, vec);
  } else {
    return fail(CDFO_ERR_UNSUPPORTED, "cdfo_mdta_gram_fwd: dtype must be fp32 or bf16");
  }
  return check_launch("cdfo_mdta_gram_fwd");
}
