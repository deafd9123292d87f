// Feature extraction of CVSR_V8 on channel-chunked bf16 ("c8" = [B][C/8][H][W][8]), the layout of the tcgen05
// convolution kernels: everything of PAItransformerSA_2 / PartitionTransformerSA_2 (arch/SIDECVSR_our.py:1441-1475, :1643-1653)
// that is not a 64-channel 1x1 / 3x3 convolution (those run in conv3x3_sm100.cu) lives here, so that a steady-state step launches
// no cuDNN / cuBLAS / ATen kernel for this stage and every reduction has a fixed order (bit-identical reruns):
//   cdfo_prior_conv_c8_fwd     conv_first / conv_second: Conv2d(1, 64, 3, 1, 1) [+ lrelu 0.1] (arch:4376-4377, :4417-4419)
//   cdfo_layernorm_c8_fwd      WithBias LayerNorm over the 64 channels of a pixel (arch:1169-1198)
//   cdfo_dwconv3x3_c8_fwd      qkv_dwconv: depthwise 3x3 on the 192-channel qkv tensor (arch:1558, :1564)
//   cdfo_mdta_gram_c8_fwd      per-head q k^T over H*W and the squared norms of q / k rows (arch:1567-1572): one head = one chunk
//   cdfo_mdta_fold_fwd         softmax(normalised Gram * temperature) folded with project_out into one 64x64 matrix per sample
//   cdfo_mdta_apply_c8_fwd     x1 + M v  [and x1 + M v + x2]  (arch:1574-1576, :1470-1472)
//   cdfo_conv16_c8_fwd         the 16-channel side branch (side_to_feaoneUDSA_2, arch:1815-1832): 3x3 stride-2 padding-2 convolutions
//                              and 3x3 stride-2 padding-2 transposed convolutions (output_padding 0 / 1) + lrelu, direct
//   cdfo_spatial_gate_c8_fwd   SpatialAttention (arch:1883-1899): x * sigmoid(conv7x7([max_c x, mean_c x]))
// fp32 arithmetic throughout, bf16 storage between kernels.
#include "cdfo_common.cuh"

namespace cdfo {
namespace fc8 {

__device__ __forceinline__ void unpack8(const uint4 q, float (&v)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf2(v[0], v[1]), pack_bf2(v[2], v[3]), pack_bf2(v[4], v[5]), pack_bf2(v[6], v[7]));
}
__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : 0.1f * x; }

// ---- Conv2d(1, Co, 3, 1, 1) on a one-channel fp32 map -> c8 bf16 (same tap order as cdfo_prior_conv_fwd) ----
__global__ void __launch_bounds__(128) prior_conv_c8_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                            const float *__restrict__ bias, uint4 *__restrict__ y, int Co, int H, int W,
                                                            int act) {
  extern __shared__ float ws[];   // [Co][9] weights + [Co] bias
  for (int e = threadIdx.x; e < Co * 9; e += blockDim.x) ws[e] = w[e];
  for (int e = threadIdx.x; e < Co; e += blockDim.x) ws[Co * 9 + e] = bias ? bias[e] : 0.f;
  __syncthreads();
  const int HW = H * W, p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (p >= HW) return;
  const int h = p / W, wq = p - h * W;
  const float *xp = x + (size_t)b * HW;
  float v[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int hh = h + i - 1, ww = wq + j - 1;
      v[i * 3 + j] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xp + hh * W + ww) : 0.f;
    }
  uint4 *yp = y + (size_t)b * (Co / 8) * HW + p;
  for (int kc = 0; kc < Co / 8; ++kc) {
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = kc * 8 + e;
      float acc = ws[Co * 9 + c];
#pragma unroll
      for (int k = 0; k < 9; ++k) acc = fmaf(ws[c * 9 + k], v[k], acc);
      o[e] = act ? lrelu(acc) : acc;
    }
    yp[(size_t)kc * HW] = pack8(o);
  }
}

// ---- LayerNorm over the 64 channels of a pixel: thread per pixel, 8 x 16-byte loads / stores, coalesced per chunk plane ----
__global__ void __launch_bounds__(128) layernorm_c8_kernel(const uint4 *__restrict__ x, const float *__restrict__ gamma,
                                                           const float *__restrict__ beta, uint4 *__restrict__ y, int HW, float eps) {
  __shared__ float gs[64], bs[64];
  if (threadIdx.x < 64) { gs[threadIdx.x] = gamma[threadIdx.x]; bs[threadIdx.x] = beta[threadIdx.x]; }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (p >= HW) return;
  const uint4 *xp = x + (size_t)b * 8 * HW + p;
  float v[64], s = 0.f;
#pragma unroll
  for (int kc = 0; kc < 8; ++kc) {
    float t[8];
    unpack8(__ldg(xp + (size_t)kc * HW), t);
#pragma unroll
    for (int e = 0; e < 8; ++e) { v[kc * 8 + e] = t[e]; s += t[e]; }
  }
  const float mu = s * (1.f / 64.f);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < 64; ++c) { const float d = v[c] - mu; q = fmaf(d, d, q); }
  const float r = rsqrtf(q * (1.f / 64.f) + eps);
  uint4 *yp = y + (size_t)b * 8 * HW + p;
#pragma unroll
  for (int kc = 0; kc < 8; ++kc) {
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (v[kc * 8 + e] - mu) * r * gs[kc * 8 + e] + bs[kc * 8 + e];
    yp[(size_t)kc * HW] = pack8(o);
  }
}

// ---- depthwise 3x3, stride 1, padding 1, no bias ----
// Thread = (column, 8-channel chunk, strip of kDwRows rows): the three input rows an output row needs slide through registers (24 values
// unpacked per new row instead of 72 per output, 3 loads instead of 9, no per-tap bounds tests: rows / columns outside the frame enter
// the window as zeros), the 72 weights of the chunk live in registers.  Same fp32 tap order (i, j row-major) as the direct form.
constexpr int kDwRows = 16;
__global__ void __launch_bounds__(128) dwconv3x3_c8_kernel(const uint4 *__restrict__ x, const float *__restrict__ w, uint4 *__restrict__ y,
                                                           int C8, int H, int W, int strips) {
  const int kc = blockIdx.y / strips, strip = blockIdx.y - kc * strips, b = blockIdx.z;
  const int wq = blockIdx.x * blockDim.x + threadIdx.x;
  if (wq >= W) return;
  float wr[8][9];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[e][k] = __ldg(w + kc * 72 + e * 9 + k);
  const uint4 *xp = x + ((size_t)b * C8 + kc) * H * W;
  uint4 *yp = y + ((size_t)b * C8 + kc) * H * W;
  const int r0 = strip * kDwRows, r1 = min(r0 + kDwRows, H);
  const bool has_l = wq > 0, has_r = wq + 1 < W;
  float win[3][3][8];                           // [row slot][column -1, 0, +1][channel]
  uint4 raw[2][3];                              // rows r + 1 and r + 2 in flight (two rows of look-ahead: the kernel is latency-bound)
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  auto load_raw = [&](int r, uint4 (&dst)[3]) {
    dst[0] = dst[1] = dst[2] = z;
    if (r >= 0 && r < H) {
      const uint4 *row = xp + (size_t)r * W + wq;
      if (has_l) dst[0] = __ldg(row - 1);
      dst[1] = __ldg(row);
      if (has_r) dst[2] = __ldg(row + 1);
    }
  };
  auto unpack_row = [&](const uint4 (&src)[3], float (&dst)[3][8]) {
    unpack8(src[0], dst[0]);
    unpack8(src[1], dst[1]);
    unpack8(src[2], dst[2]);
  };
  {
    uint4 t0[3], t1[3];
    load_raw(r0 - 1, t0);
    load_raw(r0, t1);
    load_raw(r0 + 1, raw[0]);
    load_raw(r0 + 2, raw[1]);
    unpack_row(t0, win[0]);
    unpack_row(t1, win[1]);
  }
  for (int rb = r0; rb < r1; rb += 6) {
#pragma unroll
    for (int u = 0; u < 6; ++u) {               // output row rb + u: window slots (u, u + 1, u + 2) mod 3 hold rows r - 1, r, r + 1
      const int r = rb + u;
      if (r < r1) {
        unpack_row(raw[u % 2], win[(u + 2) % 3]);          // row r + 1, loaded two steps ago
        load_raw(r + 3, raw[u % 2]);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(wr[e][i * 3 + j], win[(u + i) % 3][j][e], acc[e]);
        yp[(size_t)r * W + wq] = pack8(acc);
      }
    }
  }
}

// ---- per-head Gram q k^T and squared row norms: head h = chunk h of q (chunks 0..7) and of k (chunks 8..15) ----
// partial [B][parts][640] = (G [8 heads][8][8] | |q|^2 [64] | |k|^2 [64]); block = 256 threads over a contiguous pixel range
constexpr int kGramThreads = 256;
__global__ void __launch_bounds__(kGramThreads) mdta_gram_c8_kernel(const uint4 *__restrict__ qkv, float *__restrict__ partial, int C8,
                                                                    int HW, int parts) {
  const int part = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int p0 = (int)((long long)part * HW / parts), p1 = (int)((long long)(part + 1) * HW / parts);
  const uint4 *qp = qkv + ((size_t)b * C8 + head) * HW, *kp = qkv + ((size_t)b * C8 + 8 + head) * HW;
  float g[64], nq[8], nk[8];
#pragma unroll
  for (int i = 0; i < 64; ++i) g[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) nq[i] = nk[i] = 0.f;
  for (int p = p0 + threadIdx.x; p < p1; p += kGramThreads) {
    float q[8], k[8];
    unpack8(__ldg(qp + p), q);
    unpack8(__ldg(kp + p), k);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      nq[i] = fmaf(q[i], q[i], nq[i]);
      nk[i] = fmaf(k[i], k[i], nk[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[i * 8 + j] = fmaf(q[i], k[j], g[i * 8 + j]);
    }
  }
  // fixed-order block reduction: lanes by xor-shuffle, then the 8 warps in order
  __shared__ float red[kGramThreads / 32][80];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 80; ++i) {
    float v = i < 64 ? g[i] : (i < 72 ? nq[i - 64] : nk[i - 72]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 80) {
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < kGramThreads / 32; ++wv) v += red[wv][threadIdx.x];
    float *out = partial + ((size_t)b * parts + part) * 640;
    const int i = threadIdx.x;
    if (i < 64) out[head * 64 + i] = v;
    else if (i < 72) out[512 + head * 8 + (i - 64)] = v;
    else out[576 + head * 8 + (i - 72)] = v;
  }
}

// ---- M[b] = project_out . blockdiag_h softmax_j( G_h[i][j] / (max(|q_i|, eps) max(|k_j|, eps)) * T_h ): one CTA per sample ----
__global__ void __launch_bounds__(256) mdta_fold_kernel(const float *__restrict__ partial, const float *__restrict__ temperature,
                                                        const float *__restrict__ proj, float *__restrict__ M, int parts) {
  __shared__ float s[640], attn[512];
  const int b = blockIdx.x, t = threadIdx.x;
  for (int i = t; i < 640; i += blockDim.x) {
    const float *src = partial + (size_t)b * parts * 640 + i;
    float v = 0.f;
#pragma unroll 8
    for (int pt = 0; pt < parts; ++pt) v += src[(size_t)pt * 640];     // fixed order; independent loads, eight in flight
    s[i] = v;
  }
  __syncthreads();
  if (t < 64) {   // row t = (head, i): softmax over j
    const int head = t >> 3, i = t & 7;
    const float nq = fmaxf(sqrtf(s[512 + head * 8 + i]), 1e-12f);      // F.normalize: x / max(|x|, eps)
    float l[8], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float nk = fmaxf(sqrtf(s[576 + head * 8 + j]), 1e-12f);
      l[j] = s[head * 64 + i * 8 + j] / (nq * nk) * temperature[head];
      mx = fmaxf(mx, l[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { l[j] = __expf(l[j] - mx); sum += l[j]; }
#pragma unroll
    for (int j = 0; j < 8; ++j) attn[head * 64 + i * 8 + j] = l[j] / sum;
  }
  __syncthreads();
  // M[o][8 head + j] = sum_i proj[o][8 head + i] attn[head][i][j]
  for (int e = t; e < 4096; e += blockDim.x) {
    const int o = e >> 6, c = e & 63, head = c >> 3, j = c & 7;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(proj[o * 64 + head * 8 + i], attn[head * 64 + i * 8 + j], acc);
    M[(size_t)b * 4096 + e] = acc;
  }
}

// ---- out1 = x1 + M_b v, optionally out2 = out1 + x2, on the tensor cores (warp-level mma.sync m16n8k16 bf16, fp32 accumulate) ----
// Pixels are the M dimension: the A fragments (v: 16 pixels x 16 input channels) and the C fragments (16 pixels x 8 output channels) are
// 4-byte words of the c8 chunks, loaded from / stored to global memory directly (a warp instruction covers 8 pixels x 16 bytes = 128
// contiguous bytes per chunk) -- no shared-memory staging; the B fragments (M_b as bf16, 64 registers) are loaded once per CTA.  The SIMT
// form (thread per pixel, 4096 FMAs + 1024 broadcast LDS.128) was bound by both the fp32 pipe and the shared-memory reads at ~60 us each for
// a job that moves 30 us of HBM traffic.
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
constexpr int kApplyLd = 36;      // words per row of M in shared memory: 36 % 32 == 4 -> conflict-free B-fragment loads
__global__ void __launch_bounds__(128) mdta_apply_c8_kernel(const uint4 *__restrict__ qkv, int C8, int v_chunk0, const float *__restrict__ M,
                                                            const uint4 *__restrict__ x1, const uint4 *__restrict__ x2,
                                                            uint4 *__restrict__ out1, uint4 *__restrict__ out2, int HW, int tiles_per_cta) {
  __shared__ uint32_t Ms[64 * kApplyLd];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const float2 *Mb = reinterpret_cast<const float2 *>(M + (size_t)b * 4096);
  for (int e = tid; e < 2048; e += blockDim.x) {
    const float2 m = Mb[e];
    Ms[(e >> 5) * kApplyLd + (e & 31)] = pack_bf2(m.x, m.y);
  }
  __syncthreads();
  uint32_t bfr[8][4][2];            // [output-channel tile nt][k step s]: M[8 nt + g][16 s + 2 t, +1], M[8 nt + g][16 s + 8 + 2 t, +1]
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int sk = 0; sk < 4; ++sk) {
      bfr[nt][sk][0] = Ms[(nt * 8 + g) * kApplyLd + 8 * sk + t];
      bfr[nt][sk][1] = Ms[(nt * 8 + g) * kApplyLd + 8 * sk + 4 + t];
    }
  const uint32_t *vw = reinterpret_cast<const uint32_t *>(qkv + ((size_t)b * C8 + v_chunk0) * HW);     // chunk kc, pixel p, word t: (kc HW + p) 4 + t
  const uint32_t *x1w = reinterpret_cast<const uint32_t *>(x1 + (size_t)b * 8 * HW);
  const uint32_t *x2w = x2 ? reinterpret_cast<const uint32_t *>(x2 + (size_t)b * 8 * HW) : nullptr;
  uint32_t *o1w = reinterpret_cast<uint32_t *>(out1 + (size_t)b * 8 * HW);
  uint32_t *o2w = out2 ? reinterpret_cast<uint32_t *>(out2 + (size_t)b * 8 * HW) : nullptr;
  const int tile0 = (blockIdx.x * 4 + warp) * tiles_per_cta;       // 16-pixel tiles of this warp
  for (int tl = 0; tl < tiles_per_cta; ++tl) {
    const int px = (tile0 + tl) * 16;
    if (px >= HW) break;
    const int pa = px + g, pb = px + g + 8;
    const bool va = pa < HW, vb = pb < HW;
    uint32_t a[4][4];
#pragma unroll
    for (int sk = 0; sk < 4; ++sk) {
      a[sk][0] = va ? __ldg(vw + ((size_t)(2 * sk) * HW + pa) * 4 + t) : 0u;
      a[sk][1] = vb ? __ldg(vw + ((size_t)(2 * sk) * HW + pb) * 4 + t) : 0u;
      a[sk][2] = va ? __ldg(vw + ((size_t)(2 * sk + 1) * HW + pa) * 4 + t) : 0u;
      a[sk][3] = vb ? __ldg(vw + ((size_t)(2 * sk + 1) * HW + pb) * 4 + t) : 0u;
    }
    uint32_t ra[8], rb[8], sa[8], sb[8];      // residual words: x1 (and x2) at (pixel pa / pb, chunk nt, word t)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      ra[nt] = va ? __ldg(x1w + ((size_t)nt * HW + pa) * 4 + t) : 0u;
      rb[nt] = vb ? __ldg(x1w + ((size_t)nt * HW + pb) * 4 + t) : 0u;
      sa[nt] = (x2w && va) ? __ldg(x2w + ((size_t)nt * HW + pa) * 4 + t) : 0u;
      sb[nt] = (x2w && vb) ? __ldg(x2w + ((size_t)nt * HW + pb) * 4 + t) : 0u;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int sk = 0; sk < 4; ++sk) mma_bf16_16816(c, a[sk], bfr[nt][sk][0], bfr[nt][sk][1]);
      // c[0..1] = (pixel pa, channels 8 nt + 2 t, + 1), c[2..3] = pixel pb
      const uint32_t oa = pack_bf2(c[0] + __uint_as_float(ra[nt] << 16), c[1] + __uint_as_float(ra[nt] & 0xffff0000u));
      const uint32_t ob = pack_bf2(c[2] + __uint_as_float(rb[nt] << 16), c[3] + __uint_as_float(rb[nt] & 0xffff0000u));
      if (va) o1w[((size_t)nt * HW + pa) * 4 + t] = oa;
      if (vb) o1w[((size_t)nt * HW + pb) * 4 + t] = ob;
      if (o2w) {
        // the second output continues from the bf16-rounded first one, exactly as a separate add kernel reading out1 would
        if (va) o2w[((size_t)nt * HW + pa) * 4 + t] = pack_bf2(__uint_as_float(oa << 16) + __uint_as_float(sa[nt] << 16),
                                                                __uint_as_float(oa & 0xffff0000u) + __uint_as_float(sa[nt] & 0xffff0000u));
        if (vb) o2w[((size_t)nt * HW + pb) * 4 + t] = pack_bf2(__uint_as_float(ob << 16) + __uint_as_float(sb[nt] << 16),
                                                                __uint_as_float(ob & 0xffff0000u) + __uint_as_float(sb[nt] & 0xffff0000u));
      }
    }
  }
}

// ---- 16 -> 16 channels, 3x3, stride 2, padding 2: forward convolution (mode 0) or transposed convolution (mode 1), + lrelu 0.1 ----
// x [B][2][Hi][Wi][8], y [B][y_chunks][Ho][Wo][8] (chunks 0 and 1 written); w fp32: mode 0 [co][ci][3][3], mode 1 [ci][co][3][3] (torch layouts)
__global__ void __launch_bounds__(128) conv16_c8_kernel(const uint4 *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                                                        uint4 *__restrict__ y, int Hi, int Wi, int Ho, int Wo, int y_chunks, int mode) {
  __shared__ float ws[9 * 16 * 16];      // [tap][ci][co]
  __shared__ float bs[16];
  for (int e = threadIdx.x; e < 2304; e += blockDim.x) {
    const int co = e & 15, ci = (e >> 4) & 15, tap = e >> 8;
    ws[e] = mode == 0 ? w[(co * 16 + ci) * 9 + tap] : w[(ci * 16 + co) * 9 + tap];
  }
  if (threadIdx.x < 16) bs[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (p >= Ho * Wo) return;
  const int oy = p / Wo, ox = p - oy * Wo;
  const size_t HWi = (size_t)Hi * Wi;
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = bs[c];
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const int i = tap / 3, j = tap - i * 3;
    int iy, ix;
    if (mode == 0) {                 // y[oy] = sum_i w[i] x[2 oy - 2 + i]
      iy = 2 * oy - 2 + i;
      ix = 2 * ox - 2 + j;
    } else {                         // transposed: oy = 2 iy - 2 + i  <=>  iy = (oy + 2 - i) / 2 when that is an integer
      const int ny = oy + 2 - i, nx = ox + 2 - j;
      if ((ny | nx) & 1) continue;
      iy = ny >> 1;
      ix = nx >> 1;
    }
    if (iy < 0 || iy >= Hi || ix < 0 || ix >= Wi) continue;
    float v[16];
    {
      float t[8];
      unpack8(__ldg(x + ((size_t)b * 2 + 0) * HWi + (size_t)iy * Wi + ix), t);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = t[e];
      unpack8(__ldg(x + ((size_t)b * 2 + 1) * HWi + (size_t)iy * Wi + ix), t);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[8 + e] = t[e];
    }
    const float4 *wt = reinterpret_cast<const float4 *>(ws + tap * 256);
#pragma unroll
    for (int ci = 0; ci < 16; ++ci) {
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 m = wt[ci * 4 + c4];
        acc[c4 * 4 + 0] = fmaf(m.x, v[ci], acc[c4 * 4 + 0]);
        acc[c4 * 4 + 1] = fmaf(m.y, v[ci], acc[c4 * 4 + 1]);
        acc[c4 * 4 + 2] = fmaf(m.z, v[ci], acc[c4 * 4 + 2]);
        acc[c4 * 4 + 3] = fmaf(m.w, v[ci], acc[c4 * 4 + 3]);
      }
    }
  }
  float o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = lrelu(acc[e]);
  y[((size_t)b * y_chunks + 0) * Ho * Wo + p] = pack8(o);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = lrelu(acc[8 + e]);
  y[((size_t)b * y_chunks + 1) * Ho * Wo + p] = pack8(o);
}

// ---- SpatialAttention on a 16-channel map: pooled = (max_c, mean_c) -> 7x7 conv (2 -> 1, padding 3) -> sigmoid -> x * gate ----
__global__ void __launch_bounds__(128) spatial_pool16_kernel(const uint4 *__restrict__ x, float2 *__restrict__ pooled, int HW) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (p >= HW) return;
  float a[8], c[8];
  unpack8(__ldg(x + ((size_t)b * 2 + 0) * HW + p), a);
  unpack8(__ldg(x + ((size_t)b * 2 + 1) * HW + p), c);
  float mx = a[0], s = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) { mx = fmaxf(mx, fmaxf(a[e], c[e])); s += a[e]; }
#pragma unroll
  for (int e = 0; e < 8; ++e) s += c[e];
  pooled[(size_t)b * HW + p] = make_float2(mx, s * (1.f / 16.f));
}
__global__ void __launch_bounds__(128) spatial_gate16_kernel(const uint4 *__restrict__ x, const float2 *__restrict__ pooled,
                                                             const float *__restrict__ w, const float *__restrict__ bias, uint4 *__restrict__ y,
                                                             int H, int W) {
  __shared__ float ws[98];
  if (threadIdx.x < 98) ws[threadIdx.x] = w[threadIdx.x];       // [1][2][7][7]: channel 0 = max, 1 = mean (torch.cat order, arch:1896)
  __syncthreads();
  const int HW = H * W, p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (p >= HW) return;
  const int h = p / W, wq = p - h * W;
  float acc = bias ? bias[0] : 0.f;
  for (int i = 0; i < 7; ++i) {
    const int hh = h + i - 3;
    if (hh < 0 || hh >= H) continue;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int ww = wq + j - 3;
      if (ww < 0 || ww >= W) continue;
      const float2 v = __ldg(pooled + (size_t)b * HW + hh * W + ww);
      acc = fmaf(ws[i * 7 + j], v.x, acc);
      acc = fmaf(ws[49 + i * 7 + j], v.y, acc);
    }
  }
  const float gate = 1.f / (1.f + __expf(-acc));
#pragma unroll
  for (int kc = 0; kc < 2; ++kc) {
    float t[8];
    unpack8(__ldg(x + ((size_t)b * 2 + kc) * HW + p), t);
#pragma unroll
    for (int e = 0; e < 8; ++e) t[e] *= gate;
    y[((size_t)b * 2 + kc) * HW + p] = pack8(t);
  }
}

}  // namespace fc8
}  // namespace cdfo

using namespace cdfo;

extern "C" int cdfo_prior_conv_c8_fwd(const float *x, const float *w, const float *bias, void *y_c8, int B, int Co, int H, int W, int lrelu,
                                      void *stream) {
  CDFO_REQUIRE(x && w && y_c8, CDFO_ERR_NULL, "cdfo_prior_conv_c8_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && Co > 0 && Co <= 1024 && Co % 8 == 0 && H > 0 && W > 0, CDFO_ERR_SHAPE, "cdfo_prior_conv_c8_fwd: bad shape");
  fc8::prior_conv_c8_kernel<<<dim3(ceil_div(H * W, 128), B), 128, (size_t)Co * 10 * 4, (cudaStream_t)stream>>>(x, w, bias, (uint4 *)y_c8, Co, H, W,
                                                                                                              lrelu);
  return check_launch("cdfo_prior_conv_c8_fwd");
}

extern "C" int cdfo_layernorm_c8_fwd(const void *x_c8, const float *gamma, const float *beta, void *y_c8, int B, int H, int W, float eps,
                                     void *stream) {
  CDFO_REQUIRE(x_c8 && gamma && beta && y_c8, CDFO_ERR_NULL, "cdfo_layernorm_c8_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, CDFO_ERR_SHAPE, "cdfo_layernorm_c8_fwd: bad shape");
  fc8::layernorm_c8_kernel<<<dim3(ceil_div(H * W, 128), B), 128, 0, (cudaStream_t)stream>>>((const uint4 *)x_c8, gamma, beta, (uint4 *)y_c8, H * W, eps);
  return check_launch("cdfo_layernorm_c8_fwd");
}

extern "C" int cdfo_dwconv3x3_c8_fwd(const void *x_c8, const float *w, void *y_c8, int B, int C, int H, int W, void *stream) {
  CDFO_REQUIRE(x_c8 && w && y_c8, CDFO_ERR_NULL, "cdfo_dwconv3x3_c8_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && C > 0 && C % 8 == 0 && C / 8 <= 65535 && H > 0 && W > 0, CDFO_ERR_SHAPE, "cdfo_dwconv3x3_c8_fwd: bad shape");
  const int strips = ceil_div(H, fc8::kDwRows);
  CDFO_REQUIRE((long long)(C / 8) * strips <= 65535, CDFO_ERR_SHAPE, "cdfo_dwconv3x3_c8_fwd: too many row strips");
  fc8::dwconv3x3_c8_kernel<<<dim3(ceil_div(W, 128), (C / 8) * strips, B), 128, 0, (cudaStream_t)stream>>>((const uint4 *)x_c8, w, (uint4 *)y_c8, C / 8,
                                                                                                           H, W, strips);
  return check_launch("cdfo_dwconv3x3_c8_fwd");
}

extern "C" int cdfo_mdta_gram_c8_fwd(const void *qkv_c8, float *partial, int B, int C, int H, int W, int parts, void *stream) {
  CDFO_REQUIRE(qkv_c8 && partial, CDFO_ERR_NULL, "cdfo_mdta_gram_c8_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && C >= 128 && C % 8 == 0 && H > 0 && W > 0 && parts > 0 && parts <= 65535, CDFO_ERR_SHAPE,
               "cdfo_mdta_gram_c8_fwd: bad shape");
  fc8::mdta_gram_c8_kernel<<<dim3(parts, 8, B), fc8::kGramThreads, 0, (cudaStream_t)stream>>>((const uint4 *)qkv_c8, partial, C / 8, H * W, parts);
  return check_launch("cdfo_mdta_gram_c8_fwd");
}

extern "C" int cdfo_mdta_fold_fwd(const float *partial, const float *temperature, const float *project_out, float *M, int B, int parts,
                                  void *stream) {
  CDFO_REQUIRE(partial && temperature && project_out && M, CDFO_ERR_NULL, "cdfo_mdta_fold_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && parts > 0, CDFO_ERR_SHAPE, "cdfo_mdta_fold_fwd: bad shape");
  fc8::mdta_fold_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(partial, temperature, project_out, M, parts);
  return check_launch("cdfo_mdta_fold_fwd");
}

extern "C" int cdfo_mdta_apply_c8_fwd(const void *qkv_c8, int C, int v_channel0, const float *M, const void *x1_c8, const void *x2_c8,
                                      void *out1_c8, void *out2_c8, int B, int H, int W, void *stream) {
  CDFO_REQUIRE(qkv_c8 && M && x1_c8 && out1_c8 && (!out2_c8 || x2_c8), CDFO_ERR_NULL, "cdfo_mdta_apply_c8_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && C % 8 == 0 && v_channel0 % 8 == 0 && v_channel0 >= 0 && v_channel0 + 64 <= C && H > 0 && W > 0,
               CDFO_ERR_SHAPE, "cdfo_mdta_apply_c8_fwd: bad shape");
  // a warp walks `tiles` 16-pixel tiles (its 64 B-fragment registers are loaded once): about two waves of four-warp CTAs over the SMs
  const int n_tiles = ceil_div(H * W, 16);
  int tiles = ceil_div(n_tiles * B, 4 * 2 * 6 * kNumSMs);
  if (tiles < 1) tiles = 1;
  fc8::mdta_apply_c8_kernel<<<dim3(ceil_div(n_tiles, 4 * tiles), B), 128, 0, (cudaStream_t)stream>>>(
      (const uint4 *)qkv_c8, C / 8, v_channel0 / 8, M, (const uint4 *)x1_c8, (const uint4 *)x2_c8, (uint4 *)out1_c8, (uint4 *)out2_c8, H * W, tiles);
  return check_launch("cdfo_mdta_apply_c8_fwd");
}

extern "C" int cdfo_conv16_c8_fwd(const void *x_c8, const float *w, const float *bias, void *y_c8, int B, int Hi, int Wi, int Ho, int Wo,
                                  int y_channels, int transposed, void *stream) {
  CDFO_REQUIRE(x_c8 && w && y_c8, CDFO_ERR_NULL, "cdfo_conv16_c8_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && y_channels >= 16 && y_channels % 8 == 0, CDFO_ERR_SHAPE,
               "cdfo_conv16_c8_fwd: bad shape");
  if (transposed) {
    CDFO_REQUIRE((Ho == 2 * Hi - 3 || Ho == 2 * Hi - 2) && (Wo == 2 * Wi - 3 || Wo == 2 * Wi - 2), CDFO_ERR_SHAPE,
                 "cdfo_conv16_c8_fwd: transposed output must be (in - 1) * 2 - 4 + 3 (+ 1 with output_padding)");
  } else {
    CDFO_REQUIRE(Ho == (Hi + 1) / 2 + 1 && Wo == (Wi + 1) / 2 + 1, CDFO_ERR_SHAPE, "cdfo_conv16_c8_fwd: output must be (in + 4 - 3) / 2 + 1");
  }
  fc8::conv16_c8_kernel<<<dim3(ceil_div(Ho * Wo, 128), B), 128, 0, (cudaStream_t)stream>>>((const uint4 *)x_c8, w, bias, (uint4 *)y_c8, Hi, Wi, Ho, Wo,
                                                                                          y_channels / 8, transposed ? 1 : 0);
  return check_launch("cdfo_conv16_c8_fwd");
}

extern "C" int cdfo_spatial_gate_c8_fwd(const void *x_c8, const float *w, const float *bias, void *pooled_ws, void *y_c8, int B, int H, int W,
                                        void *stream) {
  CDFO_REQUIRE(x_c8 && w && pooled_ws && y_c8, CDFO_ERR_NULL, "cdfo_spatial_gate_c8_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, CDFO_ERR_SHAPE, "cdfo_spatial_gate_c8_fwd: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  fc8::spatial_pool16_kernel<<<dim3(ceil_div(H * W, 128), B), 128, 0, s>>>((const uint4 *)x_c8, (float2 *)pooled_ws, H * W);
  fc8::spatial_gate16_kernel<<<dim3(ceil_div(H * W, 128), B), 128, 0, s>>>((const uint4 *)x_c8, (const float2 *)pooled_ws, w, bias, (uint4 *)y_c8, H, W);
  return check_launch("cdfo_spatial_gate_c8_fwd");
}
