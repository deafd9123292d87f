// Dual channel attention ("MDTA") of the alignment modules, fused:
//   MVDualAttAlignment.forward  arch/SIDECVSR_our.py:3303-3337  (8 heads x 8 channels, fusion_out without ReLU)   mode 0
//   DualAttAlignment.forward    arch/SIDECVSR_our.py:3455-3492  (4 heads x 16 channels, ReLU after fusion_out)    mode 1
//
// Reference chain per call (~30 ATen launches): flow_warp, cat, 1x1 conv, two global average pools + 64->4->64 gates,
// L2-normalise q and k over H*W, per-head q k^T (contraction over H*W), * temperature, softmax over 8/16 channels,
// attn @ v (twice, v = warped*g and pred*g), project_out 1x1 (twice).  Algebra used here:
//   * q = x and k = fused are the same in both passes, so the softmax matrix A (block diagonal, 64 x hc) is shared;
//   * attn @ (v * gate) followed by project_out is linear per pixel:  o1 = M1 warped,  o2 = M2 pred  with
//     M = P . blockdiag(A) . diag(gate)  (64 x 64 per sample);  mode 1 additionally folds the second fusion_out
//     (arch:3492):  out = ReLU(Wa (o1 + o2) + Wb x) = ReLU((Wa M1) warped + (Wa M2) pred + Wb x).
// Three kernels; the per-pixel 64 x 128 / 64 x 64 linear maps run on the tensor cores (warp-level mma.sync m16n8k8 TF32,
// fp32 accumulate), the statistics and the softmax in fp32:
//   1. mdta_stats_kernel   gathers warped (flow_warp index arithmetic of csrc/priors.cu), computes fused = Wf [warped; pred]
//                          per 64-pixel tile with a TF32 tensor-core GEMM from shared memory, and accumulates in registers,
//                          across the tiles of a persistent CTA: sum warped, sum pred, sum x^2, sum fused^2 and the
//                          per-head Gram x_c . fused_c'.  Writes warped (needed again in 3) and per-CTA partials.
//   2. mdta_attn_kernel    per sample: fixed-order reduction of the partials (deterministic), gates, normalisation,
//                          temperature, softmax by warp shuffles, the M matrices (transposed for kernel 3).
//   3. mdta_apply_kernel   per pixel o1 = M1 warped, o2 = M2 pred -> c8 bf16 [2B] (mode 0: the input of conv_offset.0)
//                          or ReLU(MA warped + MB pred + MC x) -> NCHW fp32 + channel sums for CALayer (mode 1).
// Bytes per pixel and call (fp32 NCHW inputs): 1 reads extra (gather) 256 + pred 256 + x 256 + flow 8, writes warped 256;
// 3 reads warped 256 + pred 256 (+ x 256), writes 256: ~1.8-2.3 kB/px -> 0.04 ms per call at c3 at HBM peak.
#include "cdfo_common.cuh"

namespace cdfo {
namespace mdta {

constexpr int kTP = 64;        // pixels per tile: small enough that two CTAs fit an SM, so one CTA's load phase (global latency)
                               // overlaps the other's GEMM / statistics phases
constexpr int kTPShift = 6;
constexpr int kLd = 72;        // row stride in floats of a [64][kTP] tile: 16-byte aligned rows; 72 % 32 == 8 makes the
                               // B-fragment loads of the tensor-core GEMM (4 k rows x 8 pixels per warp) conflict-free
constexpr int kThreads = 256;
constexpr int kMaxParts = 64;

__host__ __device__ inline int stats_len(int hc) { return 256 + 64 * hc; }

struct StatsParams {
  const float *x, *extra, *pred, *flow, *wf;   // wf = fusion_out weight [64][128]
  void *warped;                                // NCHW fp32 [B][64][HW], or c8 bf16 [B][8][HW][8] (C8)
  float *partial;                              // partial [B][parts][stats_len]
  int H, W, x_batch, hc, relu;
};

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

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

// Per-pixel linear maps on the tensor cores (warp-level mma.sync m16n8k8 TF32, fp32 accumulate):
//   acc[32 out x 16 px of this warp] += Wm[64 out][K] * in[K][64 px]
// Wm: TF32 bit patterns, row stride ldw with ldw % 32 == 4 (A-fragment loads conflict-free); in: fp32 tile [K][kLd], rounded
// to TF32 at load.  Warp w owns pixels 16 (w & 3) .. +15 (two 8-pixel n-tiles) and the two 16-channel m-tiles 2 (w >> 2), +1:
// acc[m][n][0..1] = (channel 16 (2 (w >> 2) + m) + g, pixels 16 (w & 3) + 8n + 2t, +1), acc[m][n][2..3] = that channel + 8.
__device__ __forceinline__ void tile_gemm_mma(float (&acc)[2][2][4], const float *__restrict__ Wm, int ldw, const float *__restrict__ in,
                                              int K, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3, px0 = (warp & 3) * 16, m0 = (warp >> 2) * 2;
  const uint32_t *Wu = reinterpret_cast<const uint32_t *>(Wm);
#pragma unroll 4
  for (int k0 = 0; k0 < K; k0 += 8) {
    uint32_t bf[2][2];
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      bf[n][0] = to_tf32(in[(k0 + t) * kLd + px0 + n * 8 + g]);
      bf[n][1] = to_tf32(in[(k0 + t + 4) * kLd + px0 + n * 8 + g]);
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const uint32_t *wr = Wu + ((m0 + m) * 16 + g) * ldw + k0 + t;
      const uint32_t af[4] = {wr[0], wr[8 * ldw], wr[4], wr[8 * ldw + 4]};
      mma_tf32(acc[m][0], af, bf[0][0], bf[0][1]);
      mma_tf32(acc[m][1], af, bf[1][0], bf[1][1]);
    }
  }
}
// accumulator fragments of this warp -> tile [64][kLd] in shared memory (optionally ReLU)
__device__ __forceinline__ void stage_frags(const float (&acc)[2][2][4], float *__restrict__ dst, int warp, int lane, bool relu) {
  const int g = lane >> 2, t = lane & 3, px0 = (warp & 3) * 16, m0 = (warp >> 2) * 2;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      float2 lo = make_float2(acc[m][n][0], acc[m][n][1]), hi = make_float2(acc[m][n][2], acc[m][n][3]);
      if (relu) { lo.x = fmaxf(lo.x, 0.f); lo.y = fmaxf(lo.y, 0.f); hi.x = fmaxf(hi.x, 0.f); hi.y = fmaxf(hi.y, 0.f); }
      *reinterpret_cast<float2 *>(dst + ((m0 + m) * 16 + g) * kLd + px0 + n * 8 + 2 * t) = lo;
      *reinterpret_cast<float2 *>(dst + ((m0 + m) * 16 + g + 8) * kLd + px0 + n * 8 + 2 * t) = hi;
    }
}
__device__ __forceinline__ void zero_frags(float (&acc)[2][2][4]) {
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
}

__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = valid ? 16 : 0;      // src-size 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void unpack8(const uint4 q, float (&v)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}

// C8 = true: x, extra, pred are c8 bf16 [B][8][HW][8] (what their producers write: conv_expand_fea_r's tcgen05 epilogue, the prior
// convolution, the centre feature packed for the stack) and warped leaves as c8 bf16 -- a corner of the gather is ONE 16-byte load per
// 8 channels (32 loads per pixel instead of 256 scalar ones) and the kernel moves half the bytes; arithmetic stays fp32 / TF32.
// Warp-specialised since round 2: threads [0, 256) are LOADERS (gather + tile loads of tile i + 1 into the other shared-memory stage, the
// warped features to global memory), threads [256, 512) are CONSUMERS (row sums, fused = Wf [warped; pred] on the tensor cores, Gram) of
// tile i; two mbarriers per stage hand the tiles over.  One CTA per SM.  Before, every thread did both and each 64-pixel tile paid one
// exposed global round trip plus five block barriers (4.3 us per tile and SM for 33 KB of traffic).
template <bool C8>
__global__ void __launch_bounds__(2 * kThreads, 1) mdta_stats_kernel(const StatsParams p) {
  extern __shared__ __align__(16) float sm[];
  float *Wm = sm;                    // [64][132] fusion_out weight as TF32 bits
  float *stage0 = Wm + 64 * 132;     // 2 stages x { Wt [64][kLd] warped, later fused | Pt [64][kLd] pred | Xt [64][kLd] x (query) }
  uint64_t *bars = reinterpret_cast<uint64_t *>(stage0 + 2 * 3 * 64 * kLd);      // full[2] | empty[2]
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  const bool loader = threadIdx.x < kThreads;
  const int tid = threadIdx.x & (kThreads - 1);       // index inside the group
  const int b = blockIdx.y, part = blockIdx.x, parts = gridDim.x;
  const int HW = p.H * p.W;
  const int ntiles = (HW + kTP - 1) / kTP;
  const int t0 = (int)((long long)part * ntiles / parts), t1 = (int)((long long)(part + 1) * ntiles / parts);
  for (int e = threadIdx.x; e < 64 * 128; e += 2 * kThreads) Wm[(e >> 7) * 132 + (e & 127)] = __uint_as_float(to_tf32(p.wf[e]));
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * i), "r"(kThreads) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto bar_wait = [&](int i, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 20000;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar0 + 8u * i), "r"(parity)
          : "memory");
    }
  };
  auto bar_arrive = [&](int i) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8u * i) : "memory"); };
  const float *xs = p.x + (size_t)(b % p.x_batch) * 64 * HW;
  const float *ex = p.extra + (size_t)b * 64 * HW;
  const float *pr = p.pred + (size_t)b * 64 * HW;
  float *wo = reinterpret_cast<float *>(p.warped) + (size_t)b * 64 * HW;
  const int hc = p.hc;
  float rs = 0.f, gacc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};      // Gram fragments of consumer warps 0..3

  if (loader) {
  float nflow_x = 0.f, nflow_y = 0.f;
  if (t0 < t1 && t0 * kTP + (tid & (kTP - 1)) < HW) {
    nflow_x = __ldg(p.flow + ((size_t)b * 2 + 0) * HW + t0 * kTP + (tid & (kTP - 1)));
    nflow_y = __ldg(p.flow + ((size_t)b * 2 + 1) * HW + t0 * kTP + (tid & (kTP - 1)));
  }
  for (int tile = t0, it = 0; tile < t1; ++tile, ++it) {
    const int p0 = tile * kTP, npx = min(kTP, HW - p0);
    const int st = it & 1;
    float *Wt = stage0 + st * 3 * 64 * kLd, *Pt = Wt + 64 * kLd, *Xt = Pt + 64 * kLd;
    bar_wait(2 + st, ((it >> 1) & 1) ^ 1);     // the consumers are done with this stage
    const bool vec = !C8 && (HW & 3) == 0;
    uint4 pq[2] = {}, xq[2] = {};
    if (C8) {    // pred and x chunks of this thread (pixel tid & 63, chunks tid >> 6 and + 4): in flight underneath the gather
      const uint4 *pr8 = reinterpret_cast<const uint4 *>(p.pred) + (size_t)b * 8 * HW;
      const uint4 *xs8 = reinterpret_cast<const uint4 *>(p.x) + (size_t)(b % p.x_batch) * 8 * HW;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int ch = (tid >> kTPShift) + 4 * j, q = tid & (kTP - 1);
        if (q < npx) {
          pq[j] = __ldg(pr8 + (size_t)ch * HW + p0 + q);
          xq[j] = __ldg(xs8 + (size_t)ch * HW + p0 + q);
        }
      }
    }
    if (vec) {   // pred and x tiles: asynchronous 16-byte copies, in flight while this thread gathers its 16 channels below
      for (int e = tid; e < 64 * (kTP / 4); e += kThreads) {
        const int c = e / (kTP / 4), q = (e % (kTP / 4)) * 4;
        const bool ok = q < npx;
        const size_t o = (size_t)c * HW + p0 + (ok ? q : 0);
        cp_async16(Pt + c * kLd + q, pr + o, ok);
        cp_async16(Xt + c * kLd + q, xs + o, ok);
      }
      cp_async_commit();
    }
    {  // ---- A: gather warped, load pred and x
      const int px = tid & (kTP - 1), part = tid >> kTPShift;   // 4 parts x 16 channels
      const int pp = p0 + px;
      const bool valid = px < npx;
      // this tile's flow was loaded one tile ago (the gather addresses depend on it: two dependent global round trips per tile otherwise)
      const float flow_x = nflow_x, flow_y = nflow_y;
      if (tile + 1 < t1 && pp + kTP < HW) {
        nflow_x = __ldg(p.flow + ((size_t)b * 2 + 0) * HW + pp + kTP);
        nflow_y = __ldg(p.flow + ((size_t)b * 2 + 1) * HW + pp + kTP);
      }
      int o00 = 0, o01 = 0, o10 = 0, o11 = 0;
      float nw = 0.f, ne = 0.f, sw = 0.f, se = 0.f;
      if (valid) {
        const int h = pp / p.W, w = pp - h * p.W;
        const float ix = warp_src_coord(w, flow_x, p.W);
        const float iy = warp_src_coord(h, flow_y, p.H);
        const float fx = floorf(ix), fy = floorf(iy);
        // clamp before the int conversion: +-inf / NaN flows (mv2mvs keeps x/0 = inf) must not index out of range
        const int x0 = (int)fminf(fmaxf(fx, -2.f), (float)p.W), y0 = (int)fminf(fmaxf(fy, -2.f), (float)p.H);
        const int x1 = x0 + 1, y1 = y0 + 1;
        const float tx1 = __fsub_rn((float)x1, ix), tx0 = __fsub_rn(ix, fx);
        const float ty1 = __fsub_rn((float)y1, iy), ty0 = __fsub_rn(iy, fy);
        const bool oy0 = y0 >= 0 && y0 < p.H, oy1 = y1 >= 0 && y1 < p.H, ox0 = x0 >= 0 && x0 < p.W, ox1 = x1 >= 0 && x1 < p.W;
        const bool fin = fx == fx && fy == fy && fabsf(fx) < 1e9f && fabsf(fy) < 1e9f;
        nw = (fin && oy0 && ox0) ? tx1 * ty1 : 0.f; ne = (fin && oy0 && ox1) ? tx0 * ty1 : 0.f;
        sw = (fin && oy1 && ox0) ? tx1 * ty0 : 0.f; se = (fin && oy1 && ox1) ? tx0 * ty0 : 0.f;
        const int cy0 = min(max(y0, 0), p.H - 1), cy1 = min(max(y1, 0), p.H - 1);
        const int cx0 = min(max(x0, 0), p.W - 1), cx1 = min(max(x1, 0), p.W - 1);
        o00 = cy0 * p.W + cx0; o01 = cy0 * p.W + cx1; o10 = cy1 * p.W + cx0; o11 = cy1 * p.W + cx1;
      }
      if (C8) {
        // 16 channels = two chunks: eight 16-byte loads in flight per thread
        const uint4 *ex8 = reinterpret_cast<const uint4 *>(p.extra) + ((size_t)b * 8 + part * 2) * HW;
        uint4 *wo8 = reinterpret_cast<uint4 *>(p.warped) + ((size_t)b * 8 + part * 2) * HW;
        uint4 t[2][4] = {};
        if (valid) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint4 *plane = ex8 + (size_t)half * HW;
            t[half][0] = __ldg(plane + o00);
            t[half][1] = __ldg(plane + o01);
            t[half][2] = __ldg(plane + o10);
            t[half][3] = __ldg(plane + o11);
          }
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float c00[8], c01[8], c10[8], c11[8], v[8];
          unpack8(t[half][0], c00); unpack8(t[half][1], c01); unpack8(t[half][2], c10); unpack8(t[half][3], c11);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // same accumulation order as grid_sample: nw, ne, sw, se
            float a = c00[i] * nw;
            a += c01[i] * ne;
            a += c10[i] * sw;
            a += c11[i] * se;
            v[i] = a;
            Wt[(part * 16 + half * 8 + i) * kLd + px] = a;
          }
          if (valid) wo8[(size_t)half * HW + pp] = make_uint4(bf2(v[0], v[1]), bf2(v[2], v[3]), bf2(v[4], v[5]), bf2(v[6], v[7]));
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int ch = part + 4 * j;
          float a[8], c[8];
          unpack8(pq[j], a);
          unpack8(xq[j], c);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            Pt[(ch * 8 + i) * kLd + px] = a[i];
            Xt[(ch * 8 + i) * kLd + px] = c[i];
          }
        }
      } else {
      // two batches of 8 channels: 32 independent loads in flight per thread (the kernel is latency-bound: 16 warps per SM)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float c00[8], c01[8], c10[8], c11[8];
        const float *plane0 = ex + (size_t)(part * 16 + half * 8) * HW;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float *plane = plane0 + (size_t)i * HW;
          c00[i] = valid ? __ldg(plane + o00) : 0.f;
          c01[i] = valid ? __ldg(plane + o01) : 0.f;
          c10[i] = valid ? __ldg(plane + o10) : 0.f;
          c11[i] = valid ? __ldg(plane + o11) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = part * 16 + half * 8 + i;
          // same accumulation order as grid_sample: nw, ne, sw, se
          float v = c00[i] * nw;
          v += c01[i] * ne;
          v += c10[i] * sw;
          v += c11[i] * se;
          if (valid) wo[(size_t)c * HW + pp] = v;
          Wt[c * kLd + px] = v;
        }
      }
      if (!vec) {
        for (int e = tid; e < 64 * kTP; e += kThreads) {
          const int c = e >> kTPShift, q = e & (kTP - 1);
          const bool ok = q < npx;
          Pt[c * kLd + q] = ok ? __ldg(pr + (size_t)c * HW + p0 + q) : 0.f;
          Xt[c * kLd + q] = ok ? __ldg(xs + (size_t)c * HW + p0 + q) : 0.f;
        }
      }
      }
    }
    cp_async_wait<0>();
    bar_arrive(st);                            // this thread's part of the stage is written
  }
  return;
  }
  // ------------------------------------------------ consumers
  for (int tile = t0, it = 0; tile < t1; ++tile, ++it) {
    const int st = it & 1;
    float *Wt = stage0 + st * 3 * 64 * kLd, *Pt = Wt + 64 * kLd, *Xt = Pt + 64 * kLd;
    bar_wait(st, (it >> 1) & 1);
    // ---- B: row sums of warped / pred / x^2, then fused = Wf [warped; pred]
    if (tid < 192) {
      const int role = tid >> 6, c = tid & 63;
      const float *row = (role == 0 ? Wt : (role == 1 ? Pt : Xt)) + c * kLd;
      float s = 0.f;
      if (role == 2) {
#pragma unroll 8
        for (int q = 0; q < kTP; q += 4) { const float4 v = ld4(row + q); s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
      } else {
#pragma unroll 8
        for (int q = 0; q < kTP; q += 4) { const float4 v = ld4(row + q); s += (v.x + v.y) + (v.z + v.w); }
      }
      rs += s;
    }
    float acc[2][2][4];
    zero_frags(acc);
    tile_gemm_mma(acc, Wm, 132, Wt, 64, tid >> 5, tid & 31);
    tile_gemm_mma(acc, Wm + 64, 132, Pt, 64, tid >> 5, tid & 31);
    asm volatile("bar.sync 1, 256;" ::: "memory");       // every consumer warp is done reading the warped tile
    stage_frags(acc, Wt, tid >> 5, tid & 31, p.relu != 0);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // ---- C: sum fused^2 and the per-head Gram
    if (tid >= 192) {
      const float *row = Wt + (tid - 192) * kLd;
      float s = 0.f;
#pragma unroll 8
      for (int q = 0; q < kTP; q += 4) { const float4 v = ld4(row + q); s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
      rs += s;
    }
    // per-head Gram x_c . fused_c' over the pixels of the tile on the tensor cores: warp rb < 4 owns the 16 x 16 block of channels
    // [16 rb, 16 rb + 16) (one head of 16 channels, or two heads of 8 on its diagonal), K = the 64 pixels.  (As fp32 dot products from
    // shared memory this phase alone cost 2 048 of the ~4 600 shared-memory wavefronts per tile that bound the kernel.)
    if ((tid >> 5) < 4) {
      const int rb = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
      const float *xa = Xt + (16 * rb + gq) * kLd + tq, *fa = Wt + (16 * rb + gq) * kLd + tq;
#pragma unroll
      for (int k0 = 0; k0 < kTP; k0 += 8) {
        const uint32_t af[4] = {to_tf32(xa[k0]), to_tf32(xa[8 * kLd + k0]), to_tf32(xa[k0 + 4]), to_tf32(xa[8 * kLd + k0 + 4])};
        mma_tf32(gacc[0], af, to_tf32(fa[k0]), to_tf32(fa[k0 + 4]));
        mma_tf32(gacc[1], af, to_tf32(fa[8 * kLd + k0]), to_tf32(fa[8 * kLd + k0 + 4]));
      }
    }
    bar_arrive(2 + st);                        // the stage may be refilled
  }
  float *out = p.partial + ((size_t)b * parts + part) * stats_len(hc);
  out[tid] = rs;   // [0,64) sum warped, [64,128) sum pred, [128,192) sum x^2, [192,256) sum fused^2
  if ((tid >> 5) < 4) {
    // C fragments: gacc[nt][0..1] = (channel 16 rb + gq, fused channel 16 rb + 8 nt + 2 tq, + 1), gacc[nt][2..3] = channel + 8.
    // out[256 + c * hc + j] = x_c . fused_{head(c) * hc + j}: with 8-channel heads only the diagonal 8 x 8 blocks are pairs of one head
    const int rb = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int c = 16 * rb + gq + 8 * r, c2 = 16 * rb + 8 * nt + 2 * tq;
        if (c / hc == c2 / hc) {
          out[256 + c * hc + (c2 % hc)] = gacc[nt][2 * r];
          out[256 + c * hc + (c2 % hc) + 1] = gacc[nt][2 * r + 1];
        }
      }
  }
}

struct AttnParams {
  const float *partial;             // [B][parts][stats_len]
  const float *du_w1, *du_b1, *du_w2, *du_b2;   // conv_du: [4][64], [4], [64][4], [64]
  const float *temperature;         // [heads]
  const float *proj;                // project_out weight [64][64]
  const float *fold;                // mode 1: fusion_out weight [64][128] (Wa | Wb); mode 0: nullptr
  float *mats;                      // [B][3][64][64]: mats[b][m][o][k]
  int parts, hc, HW;
};

__global__ void __launch_bounds__(256) mdta_attn_kernel(const AttnParams p) {
  __shared__ float st[256 + 64 * 16];
  __shared__ float gate[2][64];
  __shared__ float hid[2][4];
  __shared__ float A[64 * 16];       // A[c][j]: softmax over the hc channels c' = head(c)*hc + j
  __shared__ float M[2][64 * 65];    // M[m][o][c']
  const int tid = threadIdx.x, b = blockIdx.x, hc = p.hc, S = stats_len(hc);
  for (int e = tid; e < S; e += 256) {
    float s = 0.f;
    for (int q = 0; q < p.parts; ++q) s += p.partial[((size_t)b * p.parts + q) * S + e];   // fixed order
    st[e] = s;
  }
  __syncthreads();
  const float inv_hw = 1.f / (float)p.HW;
  if (tid < 8) {   // conv_du.0 + ReLU on the two pooled vectors
    const int which = tid >> 2, j = tid & 3;
    float s = p.du_b1[j];
    for (int c = 0; c < 64; ++c) s = fmaf(p.du_w1[j * 64 + c], st[which * 64 + c] * inv_hw, s);
    hid[which][j] = fmaxf(s, 0.f);
  }
  __syncthreads();
  if (tid < 128) {
    const int which = tid >> 6, c = tid & 63;
    float s = p.du_b2[c];
#pragma unroll
    for (int j = 0; j < 4; ++j) s = fmaf(p.du_w2[c * 4 + j], hid[which][j], s);
    gate[which][c] = 1.f / (1.f + expf(-s));
  }
  // logits -> softmax over j (rows of hc entries); one thread per row is plenty (64 rows)
  if (tid >= 128 && tid < 192) {
    const int c = tid - 128, head = c / hc;
    const float nq = fmaxf(sqrtf(st[128 + c]), 1e-12f), t = p.temperature[head];
    float lg[16], mx = -INFINITY;
    for (int j = 0; j < hc; ++j) {
      const float nk = fmaxf(sqrtf(st[192 + head * hc + j]), 1e-12f);
      lg[j] = st[256 + c * hc + j] / (nq * nk) * t;
      mx = fmaxf(mx, lg[j]);
    }
    float ds = 0.f;
    for (int j = 0; j < hc; ++j) { lg[j] = expf(lg[j] - mx); ds += lg[j]; }
    for (int j = 0; j < hc; ++j) A[c * hc + j] = lg[j] / ds;
  }
  __syncthreads();
  // M_m[o][c'] = gate_m[c'] * sum_{c in head(c')} P[o][c] A[c][c' - head*hc]
  for (int e = tid; e < 2 * 4096; e += 256) {
    const int m = e >> 12, o = (e >> 6) & 63, c2 = e & 63, head = c2 / hc, j = c2 - head * hc;
    float s = 0.f;
    for (int i = 0; i < hc; ++i) s = fmaf(p.proj[o * 64 + head * hc + i], A[(head * hc + i) * hc + j], s);
    M[m][o * 65 + c2] = s * gate[m][c2];
  }
  __syncthreads();
  float *mats = p.mats + (size_t)b * 3 * 4096;
  if (!p.fold) {
    for (int e = tid; e < 2 * 4096; e += 256) {
      const int m = e >> 12, o = (e >> 6) & 63, k = e & 63;
      mats[m * 4096 + o * 64 + k] = M[m][o * 65 + k];
    }
  } else {
    for (int e = tid; e < 3 * 4096; e += 256) {
      const int m = e >> 12, o = (e >> 6) & 63, k = e & 63;
      float s;
      if (m == 2) {
        s = p.fold[o * 128 + 64 + k];                       // Wb
      } else {
        s = 0.f;
        for (int i = 0; i < 64; ++i) s = fmaf(p.fold[o * 128 + i], M[m][i * 65 + k], s);   // (Wa M_m)[o][k]
      }
      mats[m * 4096 + o * 64 + k] = s;
    }
  }
}

struct ApplyParams {
  const void *warped, *pred, *x;   // NCHW fp32, or c8 bf16 (mdta_apply_c8_kernel)
  const float *mats;      // [B][3][64][64] (out, in)
  void *out;              // mode 0: c8 bf16 [2B][8][HW][8]; mode 1: NCHW fp32 [B][64][HW]
  float *ca_partial;      // mode 1: [B][parts][64] channel sums of out
  int H, W, B, x_batch, mode;
  int nstages;            // 2: input tiles double-buffered by cp.async (H * W % 4 == 0, mode 0); 1: one stage
};

__global__ void __launch_bounds__(kThreads, 2) mdta_apply_kernel(const ApplyParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, b = blockIdx.y, part = blockIdx.x, parts = gridDim.x;
  const int HW = p.H * p.W;
  const int ntiles = (HW + kTP - 1) / kTP;
  const int t0 = (int)((long long)part * ntiles / parts), t1 = (int)((long long)(part + 1) * ntiles / parts);
  const int nm = p.mode == 1 ? 3 : 2;
  float *Mm = sm;                    // [nm][64][68] TF32 bits
  float *St = Mm + nm * 64 * 68;     // [nstages][nm][64][kLd] input tiles (warped, pred, x); the outputs are staged over them
  const int stage_f = nm * 64 * kLd;
  for (int e = tid; e < nm * 4096; e += kThreads)
    Mm[(e >> 12) * 64 * 68 + ((e >> 6) & 63) * 68 + (e & 63)] = __uint_as_float(to_tf32(p.mats[(size_t)b * 3 * 4096 + e]));
  const float *wp = reinterpret_cast<const float *>(p.warped) + (size_t)b * 64 * HW, *pr = reinterpret_cast<const float *>(p.pred) + (size_t)b * 64 * HW;
  const float *xs = p.x ? reinterpret_cast<const float *>(p.x) + (size_t)(b % p.x_batch) * 64 * HW : nullptr;
  const int cb = tid >> 5, pg = tid & 31;
  const bool vec = (HW & 3) == 0;
  // input tiles arrive by cp.async (16-byte rows); with two stages tile i + 1 is in flight while tile i is multiplied and stored
  auto load_tile = [&](int tile, float *dst) {
    const int p0 = tile * kTP, npx = min(kTP, HW - p0);
    if (vec) {
      for (int e = tid; e < 64 * (kTP / 4); e += kThreads) {
        const int c = e / (kTP / 4), q = (e % (kTP / 4)) * 4;
        const bool ok = q < npx;
        const size_t o = (size_t)c * HW + p0 + (ok ? q : 0);
        cp_async16(dst + c * kLd + q, wp + o, ok);
        cp_async16(dst + 64 * kLd + c * kLd + q, pr + o, ok);
        if (p.mode == 1) cp_async16(dst + 128 * kLd + c * kLd + q, xs + o, ok);
      }
    } else {
      for (int e = tid; e < 64 * kTP; e += kThreads) {
        const int c = e >> kTPShift, q = e & (kTP - 1);
        const bool ok = q < npx;
        dst[c * kLd + q] = ok ? __ldg(wp + (size_t)c * HW + p0 + q) : 0.f;
        dst[64 * kLd + c * kLd + q] = ok ? __ldg(pr + (size_t)c * HW + p0 + q) : 0.f;
        if (p.mode == 1) dst[128 * kLd + c * kLd + q] = ok ? __ldg(xs + (size_t)c * HW + p0 + q) : 0.f;
      }
    }
    cp_async_commit();
  };
  const bool two = p.nstages == 2;
  if (two && t0 < t1) load_tile(t0, St);
  float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int tile = t0; tile < t1; ++tile) {
    const int p0 = tile * kTP, npx = min(kTP, HW - p0);
    float *Wt = St + (two ? ((tile - t0) & 1) * stage_f : 0), *Pt = Wt + 64 * kLd, *Xt = Pt + 64 * kLd;
    __syncthreads();       // the stage about to be refilled (two: the other one, used by tile - 1) has no readers left; Mm is complete
    if (two) {
      if (tile + 1 < t1) {
        load_tile(tile + 1, St + ((tile + 1 - t0) & 1) * stage_f);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
    } else {
      load_tile(tile, St);
      cp_async_wait<0>();
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    float a1[2][2][4];
    zero_frags(a1);
    tile_gemm_mma(a1, Mm, 68, Wt, 64, warp, lane);
    if (p.mode == 0) {
      float a2[2][2][4];
      zero_frags(a2);
      tile_gemm_mma(a2, Mm + 64 * 68, 68, Pt, 64, warp, lane);
      __syncthreads();                       // every warp is done reading the input tiles
      stage_frags(a1, Wt, warp, lane, false);
      stage_frags(a2, Pt, warp, lane, false);
      __syncthreads();
      uint4 *z = reinterpret_cast<uint4 *>(p.out);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int q = pg * 2 + j;
        if (q < npx) {
          const float *r1 = Wt + cb * 8 * kLd + q, *r2 = Pt + cb * 8 * kLd + q;
          z[((size_t)b * 8 + cb) * HW + p0 + q] = make_uint4(bf2(r1[0], r1[kLd]), bf2(r1[2 * kLd], r1[3 * kLd]),
                                                              bf2(r1[4 * kLd], r1[5 * kLd]), bf2(r1[6 * kLd], r1[7 * kLd]));
          z[((size_t)(p.B + b) * 8 + cb) * HW + p0 + q] = make_uint4(bf2(r2[0], r2[kLd]), bf2(r2[2 * kLd], r2[3 * kLd]),
                                                                      bf2(r2[4 * kLd], r2[5 * kLd]), bf2(r2[6 * kLd], r2[7 * kLd]));
        }
      }
    } else {
      tile_gemm_mma(a1, Mm + 64 * 68, 68, Pt, 64, warp, lane);
      tile_gemm_mma(a1, Mm + 2 * 64 * 68, 68, Xt, 64, warp, lane);
      __syncthreads();
      stage_frags(a1, Wt, warp, lane, true);
      __syncthreads();
      float *o = reinterpret_cast<float *>(p.out) + (size_t)b * 64 * HW;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 v = *reinterpret_cast<const float2 *>(Wt + (cb * 8 + i) * kLd + pg * 2);
        const float vv[2] = {v.x, v.y};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int q = pg * 2 + j;
          if (q < npx) {
            o[(size_t)(cb * 8 + i) * HW + p0 + q] = vv[j];
            csum[i] += vv[j];
          }
        }
      }
    }
  }
  if (p.mode == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = csum[i];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (pg == 0) p.ca_partial[((size_t)b * parts + part) * 64 + cb * 8 + i] = s;
    }
  }
}

// mode 0 with c8 bf16 inputs (warped as mdta_stats_kernel<true> wrote it, pred): o1 = M1 warped, o2 = M2 pred -> c8 bf16 [2B].
// A thread's four 16-byte chunks of tile i + 1 are loaded into registers before the GEMMs of tile i (one shared-memory stage).
__global__ void __launch_bounds__(kThreads, 2) mdta_apply_c8_kernel(const ApplyParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, b = blockIdx.y, part = blockIdx.x, parts = gridDim.x;
  const int HW = p.H * p.W;
  const int ntiles = (HW + kTP - 1) / kTP;
  const int t0 = (int)((long long)part * ntiles / parts), t1 = (int)((long long)(part + 1) * ntiles / parts);
  float *Mm = sm;                    // [2][64][68] TF32 bits
  float *Wt = Mm + 2 * 64 * 68;      // [64][kLd] warped, then o1
  float *Pt = Wt + 64 * kLd;         // [64][kLd] pred, then o2
  for (int e = tid; e < 2 * 4096; e += kThreads)
    Mm[(e >> 12) * 64 * 68 + ((e >> 6) & 63) * 68 + (e & 63)] = __uint_as_float(to_tf32(p.mats[(size_t)b * 3 * 4096 + e]));
  const uint4 *wp8 = reinterpret_cast<const uint4 *>(p.warped) + (size_t)b * 8 * HW;
  const uint4 *pr8 = reinterpret_cast<const uint4 *>(p.pred) + (size_t)b * 8 * HW;
  const int cb = tid >> 5, pg = tid & 31;
  const int q_ld = tid & (kTP - 1), ch_ld = tid >> kTPShift;      // this thread's pixel and chunks (ch_ld, ch_ld + 4) of the tile loads
  uint4 wq[2], pq[2];
  auto fetch = [&](int tile) {
    const int p0 = tile * kTP, npx = min(kTP, HW - p0);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      wq[j] = pq[j] = make_uint4(0u, 0u, 0u, 0u);
      if (q_ld < npx) {
        wq[j] = __ldg(wp8 + (size_t)(ch_ld + 4 * j) * HW + p0 + q_ld);
        pq[j] = __ldg(pr8 + (size_t)(ch_ld + 4 * j) * HW + p0 + q_ld);
      }
    }
  };
  if (t0 < t1) fetch(t0);
  for (int tile = t0; tile < t1; ++tile) {
    const int p0 = tile * kTP, npx = min(kTP, HW - p0);
    __syncthreads();       // the previous tile's staged outputs have been read; Mm is complete on the first pass
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float a[8], c[8];
      unpack8(wq[j], a);
      unpack8(pq[j], c);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        Wt[((ch_ld + 4 * j) * 8 + i) * kLd + q_ld] = a[i];
        Pt[((ch_ld + 4 * j) * 8 + i) * kLd + q_ld] = c[i];
      }
    }
    __syncthreads();
    if (tile + 1 < t1) fetch(tile + 1);
    const int warp = tid >> 5, lane = tid & 31;
    float a1[2][2][4], a2[2][2][4];
    zero_frags(a1);
    zero_frags(a2);
    tile_gemm_mma(a1, Mm, 68, Wt, 64, warp, lane);
    tile_gemm_mma(a2, Mm + 64 * 68, 68, Pt, 64, warp, lane);
    __syncthreads();                       // every warp is done reading the input tiles
    stage_frags(a1, Wt, warp, lane, false);
    stage_frags(a2, Pt, warp, lane, false);
    __syncthreads();
    uint4 *z = reinterpret_cast<uint4 *>(p.out);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int q = pg * 2 + j;
      if (q < npx) {
        const float *r1 = Wt + cb * 8 * kLd + q, *r2 = Pt + cb * 8 * kLd + q;
        z[((size_t)b * 8 + cb) * HW + p0 + q] = make_uint4(bf2(r1[0], r1[kLd]), bf2(r1[2 * kLd], r1[3 * kLd]),
                                                            bf2(r1[4 * kLd], r1[5 * kLd]), bf2(r1[6 * kLd], r1[7 * kLd]));
        z[((size_t)(p.B + b) * 8 + cb) * HW + p0 + q] = make_uint4(bf2(r2[0], r2[kLd]), bf2(r2[2 * kLd], r2[3 * kLd]),
                                                                    bf2(r2[4 * kLd], r2[5 * kLd]), bf2(r2[6 * kLd], r2[7 * kLd]));
      }
    }
  }
}

// gate[b][c] = sigmoid(W2 relu(W1 mean + b1) + b2), mean = sum of per-part channel sums / HW   (CALayer, arch:2032-2043)
__global__ void channel_gate_kernel(const float *__restrict__ partial, int parts, float inv_hw, const float *__restrict__ w1,
                                    const float *__restrict__ b1, const float *__restrict__ w2, const float *__restrict__ b2,
                                    int C, int Cmid, float *__restrict__ gate) {
  extern __shared__ float s[];
  float *mean = s, *hid = s + C;
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int q = 0; q < parts; ++q) a += partial[((size_t)b * parts + q) * C + c];
    mean[c] = a * inv_hw;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Cmid; j += blockDim.x) {
    float a = b1 ? b1[j] : 0.f;
    for (int c = 0; c < C; ++c) a = fmaf(w1[j * C + c], mean[c], a);
    hid[j] = fmaxf(a, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = b2 ? b2[c] : 0.f;
    for (int j = 0; j < Cmid; ++j) a = fmaf(w2[c * Cmid + j], hid[j], a);
    gate[(size_t)b * C + c] = 1.f / (1.f + expf(-a));
  }
}

// NCHW fp32 * scale[b][c] -> c8 bf16
__global__ void pack_c8_scaled_kernel(const float *__restrict__ x, const float *__restrict__ scale, uint4 *__restrict__ out, int C,
                                      int HW) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int c8 = blockIdx.y, b = blockIdx.z;
  const float *src = x + ((size_t)b * C + c8 * 8) * HW + p;
  const float *sc = scale + (size_t)b * C + c8 * 8;
  uint32_t v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = bf2(src[(size_t)(2 * i) * HW] * sc[2 * i], src[(size_t)(2 * i + 1) * HW] * sc[2 * i + 1]);
  out[((size_t)b * (C / 8) + c8) * HW + p] = make_uint4(v[0], v[1], v[2], v[3]);
}

// c8 bf16 -> NCHW fp32, + add[b % add_batch] (NCHW fp32)
__global__ void unpack_c8_add_kernel(const uint4 *__restrict__ in, const float *__restrict__ add, int add_batch, float *__restrict__ y,
                                     int C, int HW) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int c8 = blockIdx.y, b = blockIdx.z;
  const uint4 raw = in[((size_t)b * (C / 8) + c8) * HW + p];
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  float *dst = y + ((size_t)b * C + c8 * 8) * HW + p;
  const float *a = add + ((size_t)(b % add_batch) * C + c8 * 8) * HW + p;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    dst[(size_t)(2 * i) * HW] = __uint_as_float(w[i] << 16) + a[(size_t)(2 * i) * HW];
    dst[(size_t)(2 * i + 1) * HW] = __uint_as_float(w[i] & 0xffff0000u) + a[(size_t)(2 * i + 1) * HW];
  }
}

static int parts_for(int B) {
  int parts = (2 * kNumSMs) / (B > 0 ? B : 1);
  if (parts < 1) parts = 1;
  if (parts > kMaxParts) parts = kMaxParts;
  return parts;
}

}  // namespace mdta
}  // namespace cdfo

using namespace cdfo;

// workspace layout (floats): partial [B][parts][256 + 64*hc] | mats [B][3][4096] | warped [B][64][HW]
extern "C" size_t cdfo_mdta_workspace_bytes(int B, int H, int W, int heads) {
  if (B <= 0 || H <= 0 || W <= 0 || (heads != 4 && heads != 8)) return 0;
  const int parts = mdta::parts_for(B), hc = 64 / heads;
  return ((size_t)B * parts * mdta::stats_len(hc) + (size_t)B * 3 * 4096 + (size_t)B * 64 * H * W) * 4;
}

static int mdta_run(const void *x, int x_batch, const void *extra, const void *pred, const float *flow, const float *fusion_w,
                    const float *du_w1, const float *du_b1, const float *du_w2, const float *du_b2, const float *temperature,
                    const float *proj_w, int heads, int mode, void *out, float *ca_sums, void *workspace, int B, int H, int W, void *stream,
                    bool c8);

extern "C" int cdfo_mdta_fwd(const float *x, int x_batch, const float *extra, const float *pred, const float *flow,
                             const float *fusion_w, const float *du_w1, const float *du_b1, const float *du_w2,
                             const float *du_b2, const float *temperature, const float *proj_w, int heads, int mode,
                             void *out, float *ca_sums, void *workspace, int B, int H, int W, void *stream) {
  return mdta_run(x, x_batch, extra, pred, flow, fusion_w, du_w1, du_b1, du_w2, du_b2, temperature, proj_w, heads, mode, out, ca_sums, workspace,
                  B, H, W, stream, false);
}

extern "C" int cdfo_mdta_c8_fwd(const void *x_c8, int x_batch, const void *extra_c8, const void *pred_c8, const float *flow,
                                const float *fusion_w, const float *du_w1, const float *du_b1, const float *du_w2, const float *du_b2,
                                const float *temperature, const float *proj_w, int heads, void *out_c8, void *workspace, int B, int H, int W,
                                void *stream) {
  return mdta_run(x_c8, x_batch, extra_c8, pred_c8, flow, fusion_w, du_w1, du_b1, du_w2, du_b2, temperature, proj_w, heads, 0, out_c8, nullptr,
                  workspace, B, H, W, stream, true);
}

static int mdta_run(const void *x, int x_batch, const void *extra, const void *pred, const float *flow, const float *fusion_w,
                    const float *du_w1, const float *du_b1, const float *du_w2, const float *du_b2, const float *temperature,
                    const float *proj_w, int heads, int mode, void *out, float *ca_sums, void *workspace, int B, int H, int W, void *stream,
                    bool c8) {
  CDFO_REQUIRE(x && extra && pred && flow && fusion_w && du_w1 && du_b1 && du_w2 && du_b2 && temperature && proj_w && out && workspace,
               CDFO_ERR_NULL, "cdfo_mdta_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535, CDFO_ERR_SHAPE, "cdfo_mdta_fwd: bad shape");
  CDFO_REQUIRE(heads == 4 || heads == 8, CDFO_ERR_UNSUPPORTED, "cdfo_mdta_fwd: 4 or 8 heads over 64 channels (got %d)", heads);
  CDFO_REQUIRE(mode == 0 || mode == 1, CDFO_ERR_UNSUPPORTED, "cdfo_mdta_fwd: mode %d", mode);
  CDFO_REQUIRE(mode == 0 || ca_sums, CDFO_ERR_NULL, "cdfo_mdta_fwd: mode 1 needs ca_sums");
  if (x_batch <= 0) x_batch = B;
  CDFO_REQUIRE(B % x_batch == 0, CDFO_ERR_SHAPE, "cdfo_mdta_fwd: B (%d) must be a multiple of x_batch (%d)", B, x_batch);
  CDFO_REQUIRE((((uintptr_t)x | (uintptr_t)extra | (uintptr_t)pred | (uintptr_t)out | (uintptr_t)workspace) & 15) == 0, CDFO_ERR_SHAPE,
               "cdfo_mdta_fwd: tensors must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int parts = mdta::parts_for(B), hc = 64 / heads, HW = H * W;
  float *ws = (float *)workspace;
  float *partial = ws;
  float *mats = partial + (size_t)B * parts * mdta::stats_len(hc);
  float *warped = mats + (size_t)B * 3 * 4096;
  const int nm = mode == 1 ? 3 : 2;
  const int nstages = (mode == 0 && (H * W) % 4 == 0) ? 2 : 1;
  const size_t smem1 = (size_t)(2 * 3 * 64 * mdta::kLd + 64 * 132) * 4 + 64;
  const size_t smem3 = (size_t)(nstages * nm * 64 * mdta::kLd + nm * 64 * 68) * 4, smem3_max = (size_t)(4 * 64 * mdta::kLd + 3 * 64 * 68) * 4;
  const size_t smem3_c8 = (size_t)(2 * 64 * mdta::kLd + 2 * 64 * 68) * 4;
  static bool attr = false;
  if (!attr) {
    cudaError_t e1 = cudaFuncSetAttribute(mdta::mdta_stats_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    cudaError_t e2 = cudaFuncSetAttribute(mdta::mdta_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3_max);
    cudaError_t e3 = cudaFuncSetAttribute(mdta::mdta_stats_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    cudaError_t e4 = cudaFuncSetAttribute(mdta::mdta_apply_c8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3_c8);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess)
      return fail(CDFO_ERR_CUDA, "cdfo_mdta_fwd: cudaFuncSetAttribute: %s",
                  cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : (e3 != cudaSuccess ? e3 : e4))));
    attr = true;
  }
  mdta::StatsParams sp{(const float *)x, (const float *)extra, (const float *)pred, flow, fusion_w, warped, partial, H, W, x_batch, hc, mode == 1 ? 1 : 0};
  if (c8) mdta::mdta_stats_kernel<true><<<dim3(parts, B), 2 * mdta::kThreads, smem1, s>>>(sp);
  else mdta::mdta_stats_kernel<false><<<dim3(parts, B), 2 * mdta::kThreads, smem1, s>>>(sp);
  mdta::AttnParams ap{partial, du_w1, du_b1, du_w2, du_b2, temperature, proj_w, mode == 1 ? fusion_w : nullptr, mats, parts, hc, HW};
  mdta::mdta_attn_kernel<<<B, 256, 0, s>>>(ap);
  mdta::ApplyParams pp{warped, pred, x, mats, out, ca_sums, H, W, B, x_batch, mode, nstages};
  if (c8) mdta::mdta_apply_c8_kernel<<<dim3(parts, B), mdta::kThreads, smem3_c8, s>>>(pp);
  else mdta::mdta_apply_kernel<<<dim3(parts, B), mdta::kThreads, smem3, s>>>(pp);
  return check_launch("cdfo_mdta_fwd");
}

extern "C" int cdfo_mdta_parts(int B) { return mdta::parts_for(B); }

extern "C" int cdfo_channel_gate_fwd(const float *partial_sums, int parts, const float *w1, const float *b1, const float *w2,
                                     const float *b2, float *gate, int B, int C, int Cmid, int HW, void *stream) {
  CDFO_REQUIRE(partial_sums && w1 && w2 && gate, CDFO_ERR_NULL, "cdfo_channel_gate_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && C > 0 && Cmid > 0 && parts > 0 && HW > 0 && C + Cmid <= 8192, CDFO_ERR_SHAPE, "cdfo_channel_gate_fwd: bad shape");
  mdta::channel_gate_kernel<<<B, 128, (size_t)(C + Cmid) * 4, (cudaStream_t)stream>>>(partial_sums, parts, 1.f / (float)HW, w1, b1, w2, b2,
                                                                                     C, Cmid, gate);
  return check_launch("cdfo_channel_gate_fwd");
}

extern "C" int cdfo_pack_c8_scaled(const float *x_nchw, const float *scale, void *x_c8, int B, int C, int H, int W, void *stream) {
  CDFO_REQUIRE(x_nchw && scale && x_c8, CDFO_ERR_NULL, "cdfo_pack_c8_scaled: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && C > 0 && C % 8 == 0 && H > 0 && W > 0, CDFO_ERR_SHAPE, "cdfo_pack_c8_scaled: bad shape");
  dim3 grid(ceil_div(H * W, 128), C / 8, B);
  mdta::pack_c8_scaled_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x_nchw, scale, (uint4 *)x_c8, C, H * W);
  return check_launch("cdfo_pack_c8_scaled");
}

extern "C" int cdfo_unpack_c8_add(const void *x_c8, const float *add_nchw, int add_batch, float *y_nchw, int B, int C, int H, int W,
                                  void *stream) {
  CDFO_REQUIRE(x_c8 && add_nchw && y_nchw, CDFO_ERR_NULL, "cdfo_unpack_c8_add: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && C > 0 && C % 8 == 0 && H > 0 && W > 0, CDFO_ERR_SHAPE, "cdfo_unpack_c8_add: bad shape");
  if (add_batch <= 0) add_batch = B;
  CDFO_REQUIRE(B % add_batch == 0, CDFO_ERR_SHAPE, "cdfo_unpack_c8_add: B must be a multiple of add_batch");
  dim3 grid(ceil_div(H * W, 128), C / 8, B);
  mdta::unpack_c8_add_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const uint4 *)x_c8, add_nchw, add_batch, y_nchw, C, H * W);
  return check_launch("cdfo_unpack_c8_add");
}
