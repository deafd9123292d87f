// Prior-guided long-range attention (LLongRangAttention.forward, arch/SIDECVSR_our.py:2179-2249) without ever
// materialising a score tensor (the reference writes [B*H,W,W] + [B*W,H,H] + [B*H*W/64,64,64] fp32 scores).
//
// Observation that shapes the kernels: the residual mask is 1[softmax_c(v_c + gumbel_c) >= 0.5] (arch:2188-2195), so
// at most ONE channel per pixel is set.  The row-attention query/key  sq = conv1x9_channels(mask * q) + beta  is then
// "beta everywhere + one 9-tap bump", and the row score  sq_w . sq_w'  has the closed form
//     64 beta^2 + beta (s_w + s_w') + q_w q_w' R[c_w][c_w'],   s_w = q_w * K1[c_w]
// (K1 / R: 64 / 64x64 tables of tap sums, built on the host from directW1_conv).  Terms that do not depend on w'
// cancel in the softmax, so all unmasked queries of a row share ONE softmax distribution; masked queries get their own.
// Column attention has dense 64-d queries (9-tap mix along H of those rows) and is done as a flash-style pass per column.
//
//   lra_mask_kernel   u (uniform noise), v_max, q      -> per pixel: masked channel index (or 255) + its q value
//   lra_row_kernel    per (b, h): v -> conv1x9 over channels -> softmax(W) -> vrow, stored column-major for the next pass
//   lra_col_kernel    per (b, w): Q from the compact mask info (9-tap along H), softmax(H) . vrow -> long_out (NHWC)
//   lra_win_kernel    per 8x8 window: q with the masked channel zeroed, softmax(64) . v -> loc_out (NHWC)
//   lra_fuse_kernel   1x1 conv(128 -> 64) over [long_out, loc_out] + bias + x -> NCHW fp32
// All arithmetic fp32 (the 0.5 threshold makes the mask discontinuous: keep it and the softmax exponents exact).
#include "cdfo_common.cuh"

namespace cdfo {

struct LraTables {       // device pointers, built on the host from directW1_conv / directH1_conv
  const float *kw;       // [9] taps along channels
  const float *kh;       // [9] taps along H
  const float *k1;       // [64]    K1[c]      = sum of in-range taps of a bump centred at channel c
  const float *r;        // [64*64] R[c1][c2]  = sum_c kw[c1-c+4] kw[c2-c+4] over in-range c
  float beta, bh;        // biases of directW1_conv / directH1_conv
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------ mask
// u [B,64,H,W], vmax [B,64], q = qv[:, :64] of input_conv's output [B,128,H,W] -> midx [B,H,W] uint8, qsel [B,H,W] fp32
__global__ void lra_mask_kernel(const float *__restrict__ u, const float *__restrict__ vmax, const float *__restrict__ qv,
                                uint8_t *__restrict__ midx, float *__restrict__ qsel, int HW) {
  __shared__ float vm[64];
  const int b = blockIdx.y;
  if (threadIdx.x < 64) vm[threadIdx.x] = vmax[b * 64 + threadIdx.x];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float *up = u + (size_t)b * 64 * HW + p;
  float lmax = -INFINITY, second_sum = 0.f;
  int cmax = 0;
  float l[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) {
    const float g = -logf(-logf(up[(size_t)c * HW]));   // gumbel_softmax, arch:2173
    l[c] = vm[c] + g;
    if (l[c] > lmax) { lmax = l[c]; cmax = c; }
  }
#pragma unroll
  for (int c = 0; c < 64; ++c) second_sum += expf(l[c] - lmax);   // includes exp(0) = 1 of the max channel
  const bool on = (1.0f / second_sum) >= 0.5f;                    // softmax of the max channel >= 0.5 (arch:2194-2195)
  midx[(size_t)b * HW + p] = on ? (uint8_t)cmax : (uint8_t)255;
  qsel[(size_t)b * HW + p] = on ? qv[((size_t)b * 128 + cmax) * HW + p] : 0.f;
}

// ------------------------------------------------------------------------------------------------ rows
// CTA per (b, h).  v = qv[:, 64:]; output vrow_t [B][W][H][64] (token-major per column for the column pass).
constexpr int kRowThreads = 256;
__global__ void __launch_bounds__(kRowThreads) lra_row_kernel(const float *__restrict__ qv, const uint8_t *__restrict__ midx,
                                                             const float *__restrict__ qsel, float *__restrict__ vrow_t,
                                                             LraTables t, int H, int W) {
  extern __shared__ float sm[];
  const int b = blockIdx.y, h = blockIdx.x;
  const int HW = H * W;
  float *vr = sm;                     // [W][65]
  float *sw = vr + W * 65;            // [W]  beta * s_w
  float *e0 = sw + W;                 // [W]  exp(beta s_w - m0)
  float *qs = e0 + W;                 // [W]  q of the masked channel
  float *n0 = qs + W;                 // [64] common numerator
  float *red = n0 + 64;               // [256] reduction scratch
  float *ew = red + 256;              // [8][W] per-warp exponent scratch
  int *cs = reinterpret_cast<int *>(ew + 8 * W);  // [W] masked channel or -1
  int *mlist = cs + W;                // [W] indices of masked tokens
  __shared__ int mcount;
  __shared__ float m0_s, d0_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float kw[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) kw[i] = t.kw[i];

  // v_r[w][c] = beta + sum_t kw[t] v[c + t - 4][w]   (conv along the channel axis, arch:2219)
  const float *vbase = qv + ((size_t)b * 128 + 64) * HW + (size_t)h * W;
  for (int e = tid; e < 64 * W; e += kRowThreads) {
    const int c = e / W, w = e - c * W;
    float acc = t.beta;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int cc = c + i - 4;
      if (cc >= 0 && cc < 64) acc = fmaf(kw[i], __ldg(vbase + (size_t)cc * HW + w), acc);
    }
    vr[w * 65 + c] = acc;
  }
  if (tid == 0) mcount = 0;
  __syncthreads();
  for (int w = tid; w < W; w += kRowThreads) {
    const int c = midx[(size_t)b * HW + h * W + w];
    const float q = qsel[(size_t)b * HW + h * W + w];
    const bool on = c != 255;
    cs[w] = on ? c : -1;
    qs[w] = q;
    sw[w] = on ? t.beta * q * t.k1[c] : 0.f;
    if (on) mlist[atomicAdd(&mcount, 1)] = w;
  }
  __syncthreads();
  // common distribution of every unmasked query: softmax_w'(beta * s_w')
  float m = -INFINITY;
  for (int w = tid; w < W; w += kRowThreads) m = fmaxf(m, sw[w]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float mm = red[0];
    for (int i = 1; i < kRowThreads / 32; ++i) mm = fmaxf(mm, red[i]);
    m0_s = mm;
  }
  __syncthreads();
  const float m0 = m0_s;
  float d = 0.f;
  for (int w = tid; w < W; w += kRowThreads) {
    const float e = expf(sw[w] - m0);
    e0[w] = e;
    d += e;
  }
  d = warp_sum(d);
  __syncthreads();
  if (lane == 0) red[warp] = d;
  __syncthreads();
  if (tid == 0) {
    float dd = 0.f;
    for (int i = 0; i < kRowThreads / 32; ++i) dd += red[i];
    d0_s = dd;
  }
  {  // N0[c] = sum_w' e0[w'] v_r[w'][c]: thread = (c, quarter of the row)
    const int c = tid & 63, part = tid >> 6;
    float acc = 0.f;
    for (int w = part; w < W; w += 4) acc = fmaf(e0[w], vr[w * 65 + c], acc);
    __syncthreads();
    red[tid] = acc;
    __syncthreads();
    if (tid < 64) n0[tid] = (red[tid] + red[tid + 64] + red[tid + 128] + red[tid + 192]) / d0_s;
  }
  __syncthreads();
  // unmasked queries: the common vector
  for (int e = tid; e < W * 64; e += kRowThreads) {
    const int w = e >> 6, c = e & 63;
    if (cs[w] < 0) vrow_t[(((size_t)b * W + w) * H + h) * 64 + c] = n0[c];
  }
  // masked queries: their own softmax over w' (one warp per query)
  float *my_e = ew + warp * W;
  const int nm = mcount;
  for (int qi = warp; qi < nm; qi += kRowThreads / 32) {
    const int w = mlist[qi];
    const int c1 = cs[w];
    const float q1 = qs[w];
    float mx = -INFINITY;
    for (int w2 = lane; w2 < W; w2 += 32) {
      float x = sw[w2];
      const int c2 = cs[w2];
      if (c2 >= 0) x = fmaf(q1 * qs[w2], t.r[c1 * 64 + c2], x);
      my_e[w2] = x;
      mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    float ds = 0.f;
    for (int w2 = lane; w2 < W; w2 += 32) {
      const float e = expf(my_e[w2] - mx);
      my_e[w2] = e;
      ds += e;
    }
    ds = warp_sum(ds);
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
    for (int w2 = 0; w2 < W; ++w2) {
      const float e = my_e[w2];
      a0 = fmaf(e, vr[w2 * 65 + lane], a0);
      a1 = fmaf(e, vr[w2 * 65 + 32 + lane], a1);
    }
    float *o = vrow_t + (((size_t)b * W + w) * H + h) * 64;
    o[lane] = a0 / ds;
    o[32 + lane] = a1 / ds;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ columns
// CTA per (b, w): q_c[h] = bh + sum_t kh[t] sq[h+t-4] (zero rows outside the frame), sq[h'] = beta + bump(h').
constexpr int kColThreads = 256;
__global__ void __launch_bounds__(kColThreads) lra_col_kernel(const float *__restrict__ vrow_t, const uint8_t *__restrict__ midx,
                                                             const float *__restrict__ qsel, float *__restrict__ long_out,
                                                             LraTables t, int H, int W) {
  extern __shared__ float sm[];
  const int b = blockIdx.y, w = blockIdx.x;
  const int HW = H * W;
  float *bufA = sm;               // [H][65]: sq, later V
  float *Q = bufA + H * 65;       // [H][65]
  float *pw = Q + H * 65;         // [8][H] per-warp probabilities
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float kh[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) kh[i] = t.kh[i];

  for (int e = tid; e < H * 64; e += kColThreads) {   // sq[h][c] = beta + bump
    const int h = e >> 6, c = e & 63;
    const int cm = midx[(size_t)b * HW + h * W + w];
    float v = t.beta;
    if (cm != 255) {
      const int ti = cm - c + 4;
      if (ti >= 0 && ti <= 8) v = fmaf(__ldg(t.kw + ti), qsel[(size_t)b * HW + h * W + w], v);
    }
    bufA[h * 65 + c] = v;
  }
  __syncthreads();
  for (int e = tid; e < H * 64; e += kColThreads) {   // Q = conv9 along H
    const int h = e >> 6, c = e & 63;
    float acc = t.bh;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int hh = h + i - 4;
      if (hh >= 0 && hh < H) acc = fmaf(kh[i], bufA[hh * 65 + c], acc);
    }
    Q[h * 65 + c] = acc;
  }
  __syncthreads();
  const float *vsrc = vrow_t + ((size_t)b * W + w) * H * 64;   // contiguous [H][64]
  for (int e = tid; e < H * 64; e += kColThreads) bufA[(e >> 6) * 65 + (e & 63)] = vsrc[e];
  __syncthreads();

  float *p = pw + warp * H;
  for (int h = warp; h < H; h += kColThreads / 32) {
    const float *qh = Q + h * 65;
    float mx = -INFINITY;
    for (int h2 = lane; h2 < H; h2 += 32) {
      const float *q2 = Q + h2 * 65;
      float s = 0.f;
#pragma unroll 16
      for (int c = 0; c < 64; ++c) s = fmaf(qh[c], q2[c], s);
      p[h2] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float ds = 0.f;
    for (int h2 = lane; h2 < H; h2 += 32) {
      const float e = expf(p[h2] - mx);
      p[h2] = e;
      ds += e;
    }
    ds = warp_sum(ds);
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
    for (int h2 = 0; h2 < H; ++h2) {
      const float e = p[h2];
      a0 = fmaf(e, bufA[h2 * 65 + lane], a0);
      a1 = fmaf(e, bufA[h2 * 65 + 32 + lane], a1);
    }
    float *o = long_out + (((size_t)b * H + h) * W + w) * 64;   // NHWC
    o[lane] = a0 / ds;
    o[32 + lane] = a1 / ds;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ 8x8 windows
// CTA per window; tokens = 64 pixels, sq = q with the masked channel zeroed ((1 - mask) * q, arch:2236-2239).
constexpr int kWinThreads = 256;
__global__ void __launch_bounds__(kWinThreads) lra_win_kernel(const float *__restrict__ qv, const uint8_t *__restrict__ midx,
                                                             float *__restrict__ loc_out, int H, int W) {
  __shared__ float q[64 * 65];
  __shared__ float v[64 * 65];
  __shared__ float pr[8][64];
  const int b = blockIdx.z, wy = blockIdx.y, wx = blockIdx.x;
  const int HW = H * W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < 64 * 64; e += kWinThreads) {
    const int c = e >> 6, tok = e & 63;                      // tok = dh*8 + dw
    const int h = wy * 8 + (tok >> 3), w = wx * 8 + (tok & 7);
    const size_t pix = (size_t)h * W + w;
    float qq = qv[((size_t)b * 128 + c) * HW + pix];
    if (midx[(size_t)b * HW + pix] == c) qq = 0.f;
    q[tok * 65 + c] = qq;
    v[tok * 65 + c] = qv[((size_t)b * 128 + 64 + c) * HW + pix];
  }
  __syncthreads();
  for (int tok = warp; tok < 64; tok += kWinThreads / 32) {
    const float *qt = q + tok * 65;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 16
    for (int c = 0; c < 64; ++c) {
      s0 = fmaf(qt[c], q[lane * 65 + c], s0);
      s1 = fmaf(qt[c], q[(lane + 32) * 65 + c], s1);
    }
    const float mx = warp_max(fmaxf(s0, s1));
    const float e0 = expf(s0 - mx), e1 = expf(s1 - mx);
    const float ds = warp_sum(e0 + e1);
    pr[warp][lane] = e0;
    pr[warp][lane + 32] = e1;
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 16
    for (int t2 = 0; t2 < 64; ++t2) {
      const float e = pr[warp][t2];
      a0 = fmaf(e, v[t2 * 65 + lane], a0);
      a1 = fmaf(e, v[t2 * 65 + 32 + lane], a1);
    }
    const int h = wy * 8 + (tok >> 3), w = wx * 8 + (tok & 7);
    float *o = loc_out + (((size_t)b * H + h) * W + w) * 64;
    o[lane] = a0 / ds;
    o[32 + lane] = a1 / ds;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ fuse
// out[b][co][p] = bias[co] + sum_c Wf[co][c] long[p][c] + sum_c Wf[co][64+c] loc[p][c] + x[b][co][p]   (arch:2246-2249)
constexpr int kFuseThreads = 256, kFusePix = 64;
__global__ void __launch_bounds__(kFuseThreads) lra_fuse_kernel(const float *__restrict__ long_out, const float *__restrict__ loc_out,
                                                               const float *__restrict__ wf, const float *__restrict__ bf,
                                                               const float *__restrict__ x, float *__restrict__ out, int HW) {
  extern __shared__ float sm[];
  float *ws = sm;                    // [128][65]: ws[k][co] = Wf[co][k]
  float *in = ws + 128 * 65;         // [kFusePix][129]
  const int b = blockIdx.y, p0 = blockIdx.x * kFusePix;
  const int tid = threadIdx.x;
  for (int e = tid; e < 64 * 128; e += kFuseThreads) {
    const int co = e >> 7, k = e & 127;
    ws[k * 65 + co] = wf[e];
  }
  for (int e = tid; e < kFusePix * 64; e += kFuseThreads) {
    const int pp = e >> 6, c = e & 63;
    const int p = p0 + pp;
    float a = 0.f, l = 0.f;
    if (p < HW) {
      a = long_out[((size_t)b * HW + p) * 64 + c];
      l = loc_out[((size_t)b * HW + p) * 64 + c];
    }
    in[pp * 129 + c] = a;
    in[pp * 129 + 64 + c] = l;
  }
  __syncthreads();
  const int pp = tid & 63, cq = tid >> 6;   // thread: pixel pp, output channels cq*16 .. cq*16+15
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = bf[cq * 16 + i];
  for (int k = 0; k < 128; ++k) {
    const float a = in[pp * 129 + k];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(ws[k * 65 + cq * 16 + i], a, acc[i]);
  }
  const int p = p0 + pp;
  if (p < HW) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const size_t o = ((size_t)b * 64 + cq * 16 + i) * HW + p;
      out[o] = acc[i] + x[o];
    }
  }
}

}  // namespace cdfo

using namespace cdfo;

// tables: float[9 kw | 9 kh | 64 k1 | 4096 r] on the device; scratch sizes are the caller's (see cdfo_lra_workspace_bytes).
extern "C" size_t cdfo_lra_workspace_bytes(int B, int H, int W) {
  const size_t P = (size_t)B * H * W;
  return P * (1 + 4) + 3 * P * 64 * 4 + 256;   // midx + qsel + vrow_t + long_out + loc_out
}

extern "C" int cdfo_lra_fwd(const float *qv, const float *u, const float *vmax, const float *x, const float *tables, float beta,
                            float bh, const float *fuse_w, const float *fuse_b, float *out, void *workspace, int B, int H,
                            int W, void *stream) {
  CDFO_REQUIRE(qv && u && vmax && x && tables && fuse_w && fuse_b && out && workspace, CDFO_ERR_NULL, "cdfo_lra_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0, CDFO_ERR_SHAPE,
               "cdfo_lra_fwd: H and W must be multiples of the window size 8 (got %d x %d)", H, W);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t P = (size_t)B * H * W;
  const int HW = H * W;
  uint8_t *ws = (uint8_t *)workspace;
  float *qsel = (float *)ws;
  float *vrow_t = qsel + P;
  float *long_out = vrow_t + P * 64;
  float *loc_out = long_out + P * 64;
  uint8_t *midx = (uint8_t *)(loc_out + P * 64);
  LraTables t{tables, tables + 9, tables + 18, tables + 82, beta, bh};

  lra_mask_kernel<<<dim3(ceil_div(HW, 128), B), 128, 0, s>>>(u, vmax, qv, midx, qsel, HW);

  const size_t row_smem = ((size_t)W * 65 + 3 * W + 64 + 256 + 8 * W) * 4 + 2 * (size_t)W * 4;
  const size_t col_smem = ((size_t)2 * H * 65 + 8 * H) * 4;
  const size_t fuse_smem = ((size_t)128 * 65 + kFusePix * 129) * 4;
  const int kDynMax = 224 * 1024;  // 227 KB opt-in limit minus the kernels' small static shared memory
  CDFO_REQUIRE(row_smem <= (size_t)kDynMax && col_smem <= (size_t)kDynMax, CDFO_ERR_UNSUPPORTED,
               "cdfo_lra_fwd: frame %d x %d exceeds the shared-memory row/column buffers (W <= 730, H <= 410)", H, W);
  static bool attr = false;
  if (!attr) {
    cudaError_t e1 = cudaFuncSetAttribute(lra_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynMax);
    cudaError_t e2 = cudaFuncSetAttribute(lra_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynMax);
    cudaError_t e3 = cudaFuncSetAttribute(lra_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fuse_smem);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
      return fail(CDFO_ERR_CUDA, "cdfo_lra_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    attr = true;
  }
  lra_row_kernel<<<dim3(H, B), kRowThreads, row_smem, s>>>(qv, midx, qsel, vrow_t, t, H, W);
  lra_col_kernel<<<dim3(W, B), kColThreads, col_smem, s>>>(vrow_t, midx, qsel, long_out, t, H, W);
  lra_win_kernel<<<dim3(W / 8, H / 8, B), kWinThreads, 0, s>>>(qv, midx, loc_out, H, W);
  lra_fuse_kernel<<<dim3(ceil_div(HW, kFusePix), B), kFuseThreads, fuse_smem, s>>>(long_out, loc_out, fuse_w, fuse_b, x, out, HW);
  return check_launch("cdfo_lra_fwd");
}
