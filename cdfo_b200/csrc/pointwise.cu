// 1x1 convolutions of the hot path on the tensor cores (warp-level mma.sync m16n8k8 TF32, fp32 accumulate), fp32 I/O:
//   LLongRangAttention.input_conv   64 -> 128 on x = fea + residual prior (arch/SIDECVSR_our.py:2206, :4449: the sum is formed
//                                   while the tile is loaded and never written)
//   LLongRangAttention.conv_du_re.0 64 -> 64 + ReLU on the residual prior (arch:2183)
//   LLongRangAttention.fuse         128 -> 64 on cat[long, local] (pixel-major inputs) + bias + x (arch:2246-2249)
// y[b][co][p] = act(bias[co] + sum_k W[co][k] x[b][k][p]) + resid1[b][co][p] + resid2[b][co][p]
//   mode 0: x = in1 + in2 (in2 optional), both [B][K][HW] (NCHW);   mode 1: x = cat(in1, in2), both pixel-major [B][HW][64], K = 128.
// Persistent CTAs (two or three per SM) walk 64-pixel tiles of one sample with the weights resident in shared memory as TF32;
// HBM-bound by construction: (K + Co [+ residuals]) * 4 bytes per pixel.
#include "cdfo_common.cuh"

namespace cdfo {
namespace pw {

constexpr int kTP = 64, kThreads = 256;
constexpr int kLdT = 72;    // [k][px] tiles: 72 % 32 == 8 -> conflict-free B-fragment loads (4 k rows x 8 pixels)

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct Params {
  const float *in1, *in2, *w, *bias, *resid1, *resid2;
  float *out;
  int HW, act;
};

template <int K, int CO, int MODE>
__global__ void __launch_bounds__(kThreads, (K == 64 && CO == 128) ? 2 : 3) pointwise_kernel(const Params p) {
  constexpr int kLdW = K + 4;                   // weight rows / pixel-major tile rows: (K + 4) % 32 == 4
  constexpr int kMT = CO / 32;                  // 16-channel m-tiles per warp (two warp rows)
  constexpr int kInFloats = MODE == 0 ? K * kLdT : kTP * kLdW;
  extern __shared__ __align__(16) float sm[];
  float *Wm = sm;                               // [CO][kLdW] TF32 bits
  float *In = Wm + CO * kLdW;                   // mode 0: [K][kLdT]; mode 1: [64 px][kLdW]
  float *Out = In;                              // [CO][kLdT]: staged over the input tile once every warp is done reading it
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y, part = blockIdx.x, parts = gridDim.x, HW = p.HW;
  const int ntiles = (HW + kTP - 1) / kTP;
  const int t0 = (int)((long long)part * ntiles / parts), t1 = (int)((long long)(part + 1) * ntiles / parts);
  for (int e = tid; e < CO * K; e += kThreads) Wm[(e / K) * kLdW + (e % K)] = __uint_as_float(to_tf32(p.w[e]));
  const uint32_t *Wu = reinterpret_cast<const uint32_t *>(Wm);
  const int px0 = (warp & 3) * 16, m0 = (warp >> 2) * kMT;

  for (int tile = t0; tile < t1; ++tile) {
    const int p0 = tile * kTP, npx = min(kTP, HW - p0);
    __syncthreads();     // previous tile's Out readers are done (and Wm is complete on the first pass)
    if (MODE == 0) {
      const float *a = p.in1 + (size_t)b * K * HW, *c = p.in2 ? p.in2 + (size_t)b * K * HW : nullptr;
      if ((HW & 3) == 0 && npx == kTP) {
        for (int e = tid; e < K * (kTP / 4); e += kThreads) {
          const int k = e / (kTP / 4), q = (e % (kTP / 4)) * 4;
          float4 v = __ldg(reinterpret_cast<const float4 *>(a + (size_t)k * HW + p0 + q));
          if (c) {
            const float4 u = __ldg(reinterpret_cast<const float4 *>(c + (size_t)k * HW + p0 + q));
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
          }
          *reinterpret_cast<float4 *>(In + k * kLdT + q) = v;
        }
      } else {
        for (int e = tid; e < K * kTP; e += kThreads) {
          const int k = e / kTP, q = e % kTP;
          float v = 0.f;
          if (q < npx) v = __ldg(a + (size_t)k * HW + p0 + q) + (c ? __ldg(c + (size_t)k * HW + p0 + q) : 0.f);
          In[k * kLdT + q] = v;
        }
      }
    } else {
      const float *a = p.in1 + ((size_t)b * HW + p0) * 64, *c = p.in2 + ((size_t)b * HW + p0) * 64;
      for (int e = tid; e < kTP * 32; e += kThreads) {       // 16 float4 per pixel and input
        const int q = e >> 5, j = e & 31;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < npx) v = __ldg(reinterpret_cast<const float4 *>((j < 16 ? a : c) + (size_t)q * 64) + (j & 15));
        *reinterpret_cast<float4 *>(In + q * kLdW + j * 4) = v;
      }
    }
    __syncthreads();
    float acc[kMT][2][4];
#pragma unroll
    for (int m = 0; m < kMT; ++m)
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
#pragma unroll 4
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t bf[2][2];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        if (MODE == 0) {
          bf[n][0] = to_tf32(In[(k0 + t) * kLdT + px0 + n * 8 + g]);
          bf[n][1] = to_tf32(In[(k0 + t + 4) * kLdT + px0 + n * 8 + g]);
        } else {
          bf[n][0] = to_tf32(In[(px0 + n * 8 + g) * kLdW + k0 + t]);
          bf[n][1] = to_tf32(In[(px0 + n * 8 + g) * kLdW + k0 + t + 4]);
        }
      }
#pragma unroll
      for (int m = 0; m < kMT; ++m) {
        const uint32_t *wr = Wu + ((m0 + m) * 16 + g) * kLdW + k0 + t;
        const uint32_t af[4] = {wr[0], wr[8 * kLdW], wr[4], wr[8 * kLdW + 4]};
        mma_tf32(acc[m][0], af, bf[0][0], bf[0][1]);
        mma_tf32(acc[m][1], af, bf[1][0], bf[1][1]);
      }
    }
    __syncthreads();     // all warps are done with In: Out aliases it
#pragma unroll
    for (int m = 0; m < kMT; ++m)
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        *reinterpret_cast<float2 *>(Out + ((m0 + m) * 16 + g) * kLdT + px0 + n * 8 + 2 * t) = make_float2(acc[m][n][0], acc[m][n][1]);
        *reinterpret_cast<float2 *>(Out + ((m0 + m) * 16 + g + 8) * kLdT + px0 + n * 8 + 2 * t) = make_float2(acc[m][n][2], acc[m][n][3]);
      }
    __syncthreads();
    for (int e = tid; e < CO * kTP; e += kThreads) {
      const int co = e / kTP, q = e % kTP;
      if (q >= npx) continue;
      float v = Out[co * kLdT + q] + (p.bias ? __ldg(p.bias + co) : 0.f);
      if (p.act == 1) v = fmaxf(v, 0.f);
      const size_t o = ((size_t)b * CO + co) * HW + p0 + q;
      if (p.resid1) v += __ldg(p.resid1 + o);
      if (p.resid2) v += __ldg(p.resid2 + o);
      p.out[o] = v;
    }
  }
}

template <int K, int CO, int MODE>
static int launch(const Params &p, int B, cudaStream_t s) {
  constexpr int kLdW = K + 4;
  const size_t in_f = MODE == 0 ? K * kLdT : kTP * kLdW, out_f = CO * kLdT;
  const size_t smem = (size_t)(CO * kLdW + (in_f > out_f ? in_f : out_f)) * 4;
  auto kern = pointwise_kernel<K, CO, MODE>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(CDFO_ERR_CUDA, "cudaFuncSetAttribute(pointwise): %s", cudaGetErrorString(e));
    attr = true;
  }
  int parts = (((K == 64 && CO == 128) ? 2 : 3) * kNumSMs) / B;
  if (parts < 1) parts = 1;
  const int ntiles = (p.HW + kTP - 1) / kTP;
  if (parts > ntiles) parts = ntiles;
  kern<<<dim3(parts, B), kThreads, smem, s>>>(p);
  return check_launch("cdfo_pointwise_conv_fwd");
}

}  // namespace pw

int pointwise_conv(const float *in1, const float *in2, const float *w, const float *bias, const float *resid1, const float *resid2,
                   float *out, int B, int K, int Co, int HW, int act, int mode, cudaStream_t s) {
  pw::Params p{in1, in2, w, bias, resid1, resid2, out, HW, act};
  if (mode == 0 && K == 64 && Co == 64) return pw::launch<64, 64, 0>(p, B, s);
  if (mode == 0 && K == 64 && Co == 128) return pw::launch<64, 128, 0>(p, B, s);
  if (mode == 1 && K == 128 && Co == 64) return pw::launch<128, 64, 1>(p, B, s);
  return fail(CDFO_ERR_UNSUPPORTED, "cdfo_pointwise_conv_fwd: unsupported (mode %d, %d -> %d channels)", mode, K, Co);
}

}  // namespace cdfo

using namespace cdfo;

extern "C" int cdfo_pointwise_conv_fwd(const float *in1, const float *in2, const float *w, const float *bias, const float *resid1,
                                       const float *resid2, float *out, int B, int K, int Co, int H, int W, int act, int mode,
                                       void *stream) {
  CDFO_REQUIRE(in1 && w && out && (mode == 0 || in2), CDFO_ERR_NULL, "cdfo_pointwise_conv_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, CDFO_ERR_SHAPE, "cdfo_pointwise_conv_fwd: bad shape");
  CDFO_REQUIRE(act == 0 || act == 1, CDFO_ERR_UNSUPPORTED, "cdfo_pointwise_conv_fwd: act %d", act);
  CDFO_REQUIRE((((uintptr_t)in1 | (uintptr_t)in2 | (uintptr_t)out) & 15) == 0, CDFO_ERR_SHAPE, "cdfo_pointwise_conv_fwd: 16-byte alignment");
  return pointwise_conv(in1, in2, w, bias, resid1, resid2, out, B, K, Co, H * W, act, mode, (cudaStream_t)stream);
}
