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
  float *out;          // NCHW fp32, or nullptr when out8 is set
  uint4 *out8;         // c8 bf16 [B][out_chunks][HW][8]: the CO / 8 chunks go to [chunk0, chunk0 + CO / 8)
  int out_chunks, chunk0;
  int HW, act;
};

__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = valid ? 16 : 0;      // src-size 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}

// HW % 4 == 0 (16-byte rows).  Input tiles arrive by cp.async into a double buffer (tile i + 1 is in flight while tile i is
// multiplied and stored); mode 0 keeps in1 and in2 apart in shared memory and adds them when the B fragment is formed.  The output
// tile is staged over the consumed input buffer and leaves as 16-byte stores: NCHW fp32 rows, or -- out8 -- c8 bf16 chunks written
// straight into a channel range of the consumer's tensor (the model's cat([fea, x_n]) never exists in fp32).
template <int K, int CO, int MODE, bool IN2>
__global__ void __launch_bounds__(kThreads, (K == 64 && CO == 128) ? 2 : 3) pointwise_kernel(const Params p) {
  constexpr int kLdW = K + 4;                   // weight rows / pixel-major tile rows: (K + 4) % 32 == 4
  constexpr int kMT = CO / 32;                  // 16-channel m-tiles per warp (two warp rows)
  constexpr int kIn1 = MODE == 0 ? K * kLdT : kTP * kLdW;           // floats of one input tile
  constexpr int kInAll = (MODE == 0 && IN2) ? 2 * kIn1 : kIn1;
  constexpr int kBuf = kInAll > CO * kLdT ? kInAll : CO * kLdT;     // one pipeline stage (the output tile is staged over it)
  extern __shared__ __align__(16) float sm[];
  float *Wm = sm;                               // [CO][kLdW] TF32 bits
  float *Buf = Wm + CO * kLdW;                  // [2][kBuf]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y, part = blockIdx.x, parts = gridDim.x, HW = p.HW;
  const int ntiles = (HW + kTP - 1) / kTP;
  const int t0 = (int)((long long)part * ntiles / parts), t1 = (int)((long long)(part + 1) * ntiles / parts);
  const uint32_t *Wu = reinterpret_cast<const uint32_t *>(Wm);
  const int px0 = (warp & 3) * 16, m0 = (warp >> 2) * kMT;

  auto prefetch = [&](int tile, float *dst) {
    const int p0 = tile * kTP, npx = min(kTP, HW - p0);
    if (MODE == 0) {
      const float *a = p.in1 + (size_t)b * K * HW + p0, *c = IN2 ? p.in2 + (size_t)b * K * HW + p0 : nullptr;
      for (int e = tid; e < K * (kTP / 4); e += kThreads) {
        const int k = e / (kTP / 4), q = (e % (kTP / 4)) * 4;
        const bool ok = q < npx;
        cp_async16(dst + k * kLdT + q, a + (size_t)k * HW + (ok ? q : 0), ok);
        if (IN2) cp_async16(dst + kIn1 + k * kLdT + q, c + (size_t)k * HW + (ok ? q : 0), ok);
      }
    } else {
      const float *a = p.in1 + ((size_t)b * HW + p0) * 64, *c = p.in2 + ((size_t)b * HW + p0) * 64;
      for (int e = tid; e < kTP * 32; e += kThreads) {       // 16 float4 per pixel and input
        const int q = e >> 5, j = e & 31;
        const bool ok = q < npx;
        cp_async16(dst + q * kLdW + j * 4, (j < 16 ? a : c) + (size_t)(ok ? q : 0) * 64 + (j & 15) * 4, ok);
      }
    }
    cp_async_commit();
  };

  if (t0 < t1) prefetch(t0, Buf);
  for (int e = tid; e < CO * K; e += kThreads) Wm[(e / K) * kLdW + (e % K)] = __uint_as_float(to_tf32(p.w[e]));

  for (int tile = t0; tile < t1; ++tile) {
    const int p0 = tile * kTP, npx = min(kTP, HW - p0);
    float *In = Buf + ((tile - t0) & 1) * kBuf;
    float *Out = In;                            // [CO][kLdT]: staged over the input stage once every warp is done reading it
    if (tile + 1 < t1) {
      prefetch(tile + 1, Buf + ((tile + 1 - t0) & 1) * kBuf);   // that stage's last readers passed the barrier that ended tile - 1
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();     // this tile's bytes (every thread's copies) are visible; Wm is complete on the first pass
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
          const int i0 = (k0 + t) * kLdT + px0 + n * 8 + g, i1 = i0 + 4 * kLdT;
          bf[n][0] = to_tf32(IN2 ? In[i0] + In[kIn1 + i0] : In[i0]);
          bf[n][1] = to_tf32(IN2 ? In[i1] + In[kIn1 + i1] : In[i1]);
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
    if (p.out8) {
      // one thread = 8 channels of one pixel = one 16-byte chunk; lanes run along the pixels (512 contiguous bytes per warp)
      for (int e = tid; e < (CO / 8) * kTP; e += kThreads) {
        const int c8 = e / kTP, q = e % kTP;
        if (q >= npx) continue;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int co = c8 * 8 + i;
          float a = Out[co * kLdT + q] + (p.bias ? __ldg(p.bias + co) : 0.f);
          if (p.act == 1) a = fmaxf(a, 0.f);
          const size_t o = ((size_t)b * CO + co) * HW + p0 + q;
          if (p.resid1) a += __ldg(p.resid1 + o);
          if (p.resid2) a += __ldg(p.resid2 + o);
          v[i] = a;
        }
        p.out8[((size_t)b * p.out_chunks + p.chunk0 + c8) * HW + p0 + q] =
            make_uint4(pack_bf2(v[0], v[1]), pack_bf2(v[2], v[3]), pack_bf2(v[4], v[5]), pack_bf2(v[6], v[7]));
      }
    } else {
      for (int e = tid; e < CO * (kTP / 4); e += kThreads) {
        const int co = e / (kTP / 4), q = (e % (kTP / 4)) * 4;
        if (q >= npx) continue;
        float4 v = *reinterpret_cast<const float4 *>(Out + co * kLdT + q);
        const float bb = p.bias ? __ldg(p.bias + co) : 0.f;
        v.x += bb; v.y += bb; v.z += bb; v.w += bb;
        if (p.act == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        const size_t o = ((size_t)b * CO + co) * HW + p0 + q;
        if (p.resid1) { const float4 r = __ldg(reinterpret_cast<const float4 *>(p.resid1 + o)); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
        if (p.resid2) { const float4 r = __ldg(reinterpret_cast<const float4 *>(p.resid2 + o)); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
        *reinterpret_cast<float4 *>(p.out + o) = v;
      }
    }
    __syncthreads();     // Out is consumed: the next iteration may prefetch tile + 2 into this stage
  }
}

// any HW (rows not 16-byte aligned): synchronous scalar loads and stores, NCHW fp32 output only
template <int K, int CO, int MODE>
__global__ void __launch_bounds__(kThreads, (K == 64 && CO == 128) ? 2 : 3) pointwise_ragged_kernel(const Params p) {
  constexpr int kLdW = K + 4;
  constexpr int kMT = CO / 32;
  extern __shared__ __align__(16) float sm[];
  float *Wm = sm;
  float *In = Wm + CO * kLdW;
  float *Out = In;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y, part = blockIdx.x, parts = gridDim.x, HW = p.HW;
  const int ntiles = (HW + kTP - 1) / kTP;
  const int t0 = (int)((long long)part * ntiles / parts), t1 = (int)((long long)(part + 1) * ntiles / parts);
  for (int e = tid; e < CO * K; e += kThreads) Wm[(e / K) * kLdW + (e % K)] = __uint_as_float(to_tf32(p.w[e]));
  const uint32_t *Wu = reinterpret_cast<const uint32_t *>(Wm);
  const int px0 = (warp & 3) * 16, m0 = (warp >> 2) * kMT;
  for (int tile = t0; tile < t1; ++tile) {
    const int p0 = tile * kTP, npx = min(kTP, HW - p0);
    __syncthreads();
    if (MODE == 0) {
      const float *a = p.in1 + (size_t)b * K * HW, *c = p.in2 ? p.in2 + (size_t)b * K * HW : nullptr;
      for (int e = tid; e < K * kTP; e += kThreads) {
        const int k = e / kTP, q = e % kTP;
        float v = 0.f;
        if (q < npx) v = __ldg(a + (size_t)k * HW + p0 + q) + (c ? __ldg(c + (size_t)k * HW + p0 + q) : 0.f);
        In[k * kLdT + q] = v;
      }
    } else {
      const float *a = p.in1 + ((size_t)b * HW + p0) * 64, *c = p.in2 + ((size_t)b * HW + p0) * 64;
      for (int e = tid; e < kTP * 32; e += kThreads) {
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
    __syncthreads();
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
  constexpr int kPerSM = (K == 64 && CO == 128) ? 2 : 3;
  int parts = (kPerSM * kNumSMs) / B;
  if (parts < 1) parts = 1;
  const int ntiles = (p.HW + kTP - 1) / kTP;
  if (parts > ntiles) parts = ntiles;
  const size_t in1 = MODE == 0 ? K * kLdT : kTP * kLdW, out_f = CO * kLdT;
  if (p.HW % 4 != 0) {
    if (p.out8) return fail(CDFO_ERR_UNSUPPORTED, "cdfo_pointwise_conv_c8_fwd: H * W must be a multiple of 4");
    const size_t smem = (size_t)(CO * kLdW + (in1 > out_f ? in1 : out_f)) * 4;
    auto kern = pointwise_ragged_kernel<K, CO, MODE>;
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(CDFO_ERR_CUDA, "cudaFuncSetAttribute(pointwise): %s", cudaGetErrorString(e));
      attr = true;
    }
    kern<<<dim3(parts, B), kThreads, smem, s>>>(p);
    return check_launch("cdfo_pointwise_conv_fwd");
  }
  const bool in2 = MODE == 0 && p.in2 != nullptr;
  const size_t in_all = in2 ? 2 * in1 : in1;
  const size_t smem = (size_t)(CO * kLdW + 2 * (in_all > out_f ? in_all : out_f)) * 4;
  if (in2) {
    auto kern = pointwise_kernel<K, CO, MODE, true>;
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(CDFO_ERR_CUDA, "cudaFuncSetAttribute(pointwise): %s", cudaGetErrorString(e));
      attr = true;
    }
    kern<<<dim3(parts, B), kThreads, smem, s>>>(p);
  } else {
    auto kern = pointwise_kernel<K, CO, MODE, false>;
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(CDFO_ERR_CUDA, "cudaFuncSetAttribute(pointwise): %s", cudaGetErrorString(e));
      attr = true;
    }
    kern<<<dim3(parts, B), kThreads, smem, s>>>(p);
  }
  return check_launch("cdfo_pointwise_conv_fwd");
}

}  // namespace pw

int pointwise_conv(const float *in1, const float *in2, const float *w, const float *bias, const float *resid1, const float *resid2,
                   float *out, int B, int K, int Co, int HW, int act, int mode, cudaStream_t s, void *out_c8, int out_channels,
                   int channel0) {
  pw::Params p{in1, in2, w, bias, resid1, resid2, out_c8 ? nullptr : out, (uint4 *)out_c8, out_channels / 8, channel0 / 8, HW, act};
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
  CDFO_REQUIRE((((uintptr_t)in1 | (uintptr_t)in2 | (uintptr_t)out | (uintptr_t)resid1 | (uintptr_t)resid2) & 15) == 0, CDFO_ERR_SHAPE,
               "cdfo_pointwise_conv_fwd: 16-byte alignment");
  return pointwise_conv(in1, in2, w, bias, resid1, resid2, out, B, K, Co, H * W, act, mode, (cudaStream_t)stream, nullptr, 0, 0);
}

extern "C" int cdfo_pointwise_conv_c8_fwd(const float *in1, const float *in2, const float *w, const float *bias, const float *resid1,
                                          const float *resid2, void *out_c8, int B, int K, int Co, int H, int W, int act, int mode,
                                          int out_channels, int channel0, void *stream) {
  CDFO_REQUIRE(in1 && w && out_c8 && (mode == 0 || in2), CDFO_ERR_NULL, "cdfo_pointwise_conv_c8_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, CDFO_ERR_SHAPE, "cdfo_pointwise_conv_c8_fwd: bad shape");
  CDFO_REQUIRE(act == 0 || act == 1, CDFO_ERR_UNSUPPORTED, "cdfo_pointwise_conv_c8_fwd: act %d", act);
  CDFO_REQUIRE(out_channels % 8 == 0 && channel0 % 8 == 0 && channel0 >= 0 && channel0 + Co <= out_channels, CDFO_ERR_SHAPE,
               "cdfo_pointwise_conv_c8_fwd: channels [%d, %d) do not fit %d output channels (multiples of 8)", channel0, channel0 + Co, out_channels);
  CDFO_REQUIRE((((uintptr_t)in1 | (uintptr_t)in2 | (uintptr_t)out_c8 | (uintptr_t)resid1 | (uintptr_t)resid2) & 15) == 0, CDFO_ERR_SHAPE,
               "cdfo_pointwise_conv_c8_fwd: 16-byte alignment");
  return pointwise_conv(in1, in2, w, bias, resid1, resid2, nullptr, B, K, Co, H * W, act, mode, (cudaStream_t)stream, out_c8, out_channels, channel0);
}
