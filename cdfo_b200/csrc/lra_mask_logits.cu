// Mask logits of LLongRangAttention (arch/SIDECVSR_our.py:2183-2186) after the first 1x1 convolution:
//   v_max = ReLU(conv_du_re2( mean_{h,w} ReLU(conv_du_re.2(v)) )),   conv_du_re.2 = Conv2d(64, 64, 3, stride 2, padding 2)
// (the bilinear up-sampling of the 1x1 map that follows in the reference is a broadcast).  Round 1 ran the strided convolution, the
// mean and the 64 x 64 matrix-vector product as cuDNN / ATen calls; here:
//   lra_logit_conv_kernel    the stride-2 convolution as an implicit GEMM on the tensor cores (warp-level mma.sync m16n8k8 TF32, fp32
//                            accumulate -- the 0.5 threshold on softmax(v_max + gumbel) makes the mask discontinuous, so this path keeps
//                            TF32 operands like the cuDNN call it replaces, not bf16), ReLU and the spatial sum of every 4 x 16-pixel
//                            output tile in its epilogue: the [B, 64, H/2+1, W/2+1] activation is never written
//   lra_logit_finish_kernel  fixed-order sum of the tile sums / (Ho Wo) -> 64 x 64 matrix-vector product + bias + ReLU -> v_max [B, 64]
// Layout: v NCHW fp32 (what the pointwise conv_du_re.0 kernel writes).  All 9 x 64 x 64 weights stay resident in shared memory as
// TF32; the input patch of a tile (9 rows x 33 columns) streams through a double buffer in four 16-channel chunks with cp.async,
// de-interleaved by column parity so that the stride-2 B-fragment loads are conflict-free.
#include <cuda.h>

#include "cdfo_common.cuh"
#include "sm100_ptx.cuh"

namespace cdfo {
namespace lml {

constexpr int kTH = 4, kTW = 16;                  // output tile
constexpr int kPR = 2 * kTH + 1, kPC = 2 * kTW + 1;   // input patch: 9 rows x 33 columns
constexpr int kHalf = 20;                         // columns of one parity, padded (17 used)
constexpr int kPlane = kPR * 2 * kHalf;           // floats per input channel of the patch: 360 = 8 (mod 32): conflict-free k x n fragments
constexpr int kChunk = 16;                        // input channels per pipeline stage
constexpr int kLdW = 68;                          // weight rows [co][ci + 4]: 68 = 4 (mod 32)
constexpr int kThreads = 256;
constexpr size_t kSmem = (size_t)(9 * 64 * kLdW + 2 * kChunk * kPlane) * 4;

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
__device__ __forceinline__ void cp_async4(void *dst, const void *src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = valid ? 4 : 0;      // src-size 0: zero-filled, nothing read (the convolution's zero padding)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// partial [B][tiles_per_img][64]: sum over the tile's valid output pixels of ReLU(conv + bias)
__global__ void __launch_bounds__(kThreads, 1)
lra_logit_conv_kernel(const float *__restrict__ v, const float *__restrict__ w, const float *__restrict__ bias, float *__restrict__ partial,
                      int B, int H, int W, int Ho, int Wo, int tiles_x, int tiles_per_img) {
  extern __shared__ __align__(16) float sm[];
  uint32_t *Ws = reinterpret_cast<uint32_t *>(sm);             // [9][64][kLdW] TF32 bits
  float *Ps = sm + 9 * 64 * kLdW;                              // [2][kChunk][kPR][2][kHalf]
  __shared__ float red[2][64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int mw = warp & 3, nh = warp >> 2;
  const int num_tiles = B * tiles_per_img;
  const size_t HW = (size_t)H * W;

  for (int e = tid; e < 9 * 64 * 64; e += kThreads) {
    const int ci = e & 63, co = (e >> 6) & 63, tap = e >> 12;
    Ws[(tap * 64 + co) * kLdW + ci] = to_tf32(w[(co * 64 + ci) * 9 + tap]);
  }

  auto prefetch = [&](int tile, int chunk, float *dst) {
    const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
    const int iy0 = 2 * (r / tiles_x) * kTH - 2, ix0 = 2 * (r % tiles_x) * kTW - 2;
    const float *src = v + ((size_t)b * 64 + chunk * kChunk) * HW;
    for (int e = tid; e < kChunk * kPR * kPC; e += kThreads) {
      const int col = e % kPC, row = (e / kPC) % kPR, ci = e / (kPC * kPR);
      const int iy = iy0 + row, ix = ix0 + col;
      const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
      cp_async4(dst + ci * kPlane + (row * 2 + (col & 1)) * kHalf + (col >> 1), src + (size_t)ci * HW + (ok ? (size_t)iy * W + ix : 0), ok);
    }
    cp_async_commit();
  };

  int tile = blockIdx.x;
  if (tile < num_tiles) prefetch(tile, 0, Ps);
  int buf = 0;
  for (; tile < num_tiles; tile += gridDim.x) {
    float acc[4][4];
#pragma unroll
    for (int nn = 0; nn < 4; ++nn)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nn][i] = 0.f;
    for (int chunk = 0; chunk < 4; ++chunk) {
      // next stage: the following chunk of this tile, or the first chunk of this CTA's next tile
      const int ntile = chunk == 3 ? tile + gridDim.x : tile, nchunk = chunk == 3 ? 0 : chunk + 1;
      if (ntile < num_tiles) {
        prefetch(ntile, nchunk, Ps + (buf ^ 1) * kChunk * kPlane);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();                 // this stage's bytes are visible (and Ws on the first pass)
      const float *P = Ps + buf * kChunk * kPlane;
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        const int i = tap / 3, j = tap - i * 3;
#pragma unroll
        for (int c8 = 0; c8 < 2; ++c8) {
          uint32_t a[4];
          const uint32_t *wr = Ws + (tap * 64 + mw * 16 + g) * kLdW + chunk * kChunk + c8 * 8 + t;
          a[0] = wr[0];
          a[1] = wr[8 * kLdW];
          a[2] = wr[4];
          a[3] = wr[8 * kLdW + 4];
#pragma unroll
          for (int nn = 0; nn < 4; ++nn) {
            const int q = nh * 4 + nn, oy = q >> 1, ox = (q & 1) * 8 + g;          // output pixel of this B column
            const float *pp = P + (c8 * 8 + t) * kPlane + ((2 * oy + i) * 2 + (j & 1)) * kHalf + ox + (j >> 1);
            mma_tf32(acc[nn], a, to_tf32(pp[0]), to_tf32(pp[4 * kPlane]));
          }
        }
      }
      __syncthreads();                 // every warp is done with this stage before it is refilled two prefetches from now
      buf ^= 1;
    }
    // epilogue: + bias, ReLU, sum over the valid pixels of the tile; C fragment: rows (co) g, g + 8, columns (pixels) 2t, 2t + 1
    const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
    const int oy0 = (r / tiles_x) * kTH, ox0 = (r % tiles_x) * kTW;
    const float b_lo = bias[mw * 16 + g], b_hi = bias[mw * 16 + g + 8];
    float s_lo = 0.f, s_hi = 0.f;
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
      const int q = nh * 4 + nn, oy = oy0 + (q >> 1), oxb = ox0 + (q & 1) * 8 + 2 * t;
      if (oy < Ho) {
        if (oxb < Wo) { s_lo += fmaxf(acc[nn][0] + b_lo, 0.f); s_hi += fmaxf(acc[nn][2] + b_hi, 0.f); }
        if (oxb + 1 < Wo) { s_lo += fmaxf(acc[nn][1] + b_lo, 0.f); s_hi += fmaxf(acc[nn][3] + b_hi, 0.f); }
      }
    }
    s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 1);
    s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 2);
    s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 1);
    s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 2);
    if (t == 0) {
      red[nh][mw * 16 + g] = s_lo;
      red[nh][mw * 16 + g + 8] = s_hi;
    }
    __syncthreads();
    if (tid < 64) partial[((size_t)b * tiles_per_img + r) * 64 + tid] = red[0][tid] + red[1][tid];
    // (red is rewritten only after the next tile's four stage barriers)
  }
}

// ---- TMA variant (the one launched when W % 4 == 0): the 4-byte cp.async copies above cost the LSU one request per element (2 cycles per
// element and SM: the kernel's bound); here a producer warp issues ONE tiled TMA load per stage -- a box of 40 columns x 10 rows x 8
// channels of the NCHW fp32 map, zero-filled outside the frame = the convolution's padding -- into a four-stage mbarrier ring, and the
// eight MMA warps never touch the copies.  (TMA ignores an element stride on dimension 0, so the columns land interleaved.)  The stride-2
// B fragments are read as 8-byte pairs (input columns 2 ox, 2 ox + 1 = taps j = 0, 1) plus one word (j = 2); the channel stride of
// 400 floats = 16 (mod 32) makes the 8-byte loads conflict-free, the single-word loads are 2-way conflicts (4 wavefronts per 3 taps).
constexpr int kTCols = 40, kTRows = 10, kTChunk = 8, kTStages = 4;
constexpr int kTPlane = kTRows * kTCols;                  // 400 floats per channel
constexpr int kTStageBytes = kTChunk * kTPlane * 4;       // 12 800
constexpr int kTThreads = 288;                            // 8 MMA warps + 1 producer warp
constexpr size_t kTSmem = (size_t)9 * 64 * kLdW * 4 + (size_t)kTStages * kTStageBytes + 2 * kTStages * 8 + 16;

__global__ void __launch_bounds__(kTThreads, 1)
lra_logit_conv_tma_kernel(const __grid_constant__ CUtensorMap vmap, const float *__restrict__ w, const float *__restrict__ bias,
                          float *__restrict__ partial, int B, int Ho, int Wo, int tiles_x, int tiles_per_img) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint32_t *Ws = reinterpret_cast<uint32_t *>(smem_raw);                                    // [9][64][kLdW] TF32 bits
  float *Ps = reinterpret_cast<float *>(smem_raw + (size_t)9 * 64 * kLdW * 4);              // [kTStages][8 channels][10 rows][40 columns]
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)9 * 64 * kLdW * 4 + (size_t)kTStages * kTStageBytes);
  __shared__ float red[2][2][64];
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto FULL = [&](int st) { return bar0 + 8u * st; };
  auto EMPTY = [&](int st) { return bar0 + 8u * (kTStages + st); };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int num_tiles = B * tiles_per_img;
  constexpr int kChunks = 64 / kTChunk;
  const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  for (int e = tid; e < 9 * 64 * 64; e += kTThreads) {
    const int ci = e & 63, co = (e >> 6) & 63, tap = e >> 12;
    Ws[(tap * 64 + co) * kLdW + ci] = to_tf32(w[(co * 64 + ci) * 9 + tap]);
  }
  if (tid == 0) {
    for (int st = 0; st < kTStages; ++st) {
      ptx::mbar_init(FULL(st), 1);
      ptx::mbar_init(EMPTY(st), 8);
    }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&vmap);
  }
  __syncthreads();

  if (warp == 8) {
    // ---------------- producer: one box per (tile, 8-channel chunk)
    if (lane == 0) {
      int item = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int tile = blockIdx.x + lt * gridDim.x;
        const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
        // the box starts two columns left of the patch: an inner start coordinate that is not a multiple of 16 bytes faults (UTMALDG ->
        // illegal instruction, measured), and 32 tx - 4 is
        const int iy0 = 2 * (r / tiles_x) * kTH - 2, ix0 = 2 * (r % tiles_x) * kTW - 4;
        for (int chunk = 0; chunk < kChunks; ++chunk, ++item) {
          const int st = item % kTStages;
          ptx::mbar_wait_parked(EMPTY(st), ((item / kTStages) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(FULL(st), kTStageBytes);
          const uint32_t dst = ptx::smem_u32(Ps) + st * kTStageBytes;
          ptx::tma_load_3d(dst, &vmap, FULL(st), ix0, iy0, b * 64 + chunk * kTChunk);
        }
      }
    }
    return;
  }

  // ---------------- eight MMA warps: warp = (16 output channels mw, half of the tile's pixels nh)
  const int mw = warp & 3, nh = warp >> 2;
  int item = 0;
  for (int lt = 0; lt < my_tiles; ++lt) {
    const int tile = blockIdx.x + lt * gridDim.x;
    float acc[4][4];
#pragma unroll
    for (int nn = 0; nn < 4; ++nn)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nn][i] = 0.f;
    for (int chunk = 0; chunk < kChunks; ++chunk, ++item) {
      const int st = item % kTStages;
      ptx::mbar_wait_parked(FULL(st), (item / kTStages) & 1);
      const float *P = Ps + st * (kTStageBytes / 4);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        uint32_t a[3][4];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const uint32_t *wr = Ws + ((i * 3 + j) * 64 + mw * 16 + g) * kLdW + chunk * kTChunk + t;
          a[j][0] = wr[0];
          a[j][1] = wr[8 * kLdW];
          a[j][2] = wr[4];
          a[j][3] = wr[8 * kLdW + 4];
        }
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) {
          const int q = nh * 4 + nn, oy = q >> 1, ox = (q & 1) * 8 + g;          // output pixel of this B column
          const float *pp = P + t * kTPlane + (2 * oy + i) * kTCols + 2 * ox + 2;       // + 2: the box's two extra columns
          const float2 lo = *reinterpret_cast<const float2 *>(pp), hi = *reinterpret_cast<const float2 *>(pp + 4 * kTPlane);
          const float lo2 = pp[2], hi2 = pp[4 * kTPlane + 2];
          mma_tf32(acc[nn], a[0], to_tf32(lo.x), to_tf32(hi.x));
          mma_tf32(acc[nn], a[1], to_tf32(lo.y), to_tf32(hi.y));
          mma_tf32(acc[nn], a[2], to_tf32(lo2), to_tf32(hi2));
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(EMPTY(st));
    }
    // epilogue: + bias, ReLU, sum over the valid pixels of the tile; C fragment: rows (co) g, g + 8, columns (pixels) 2t, 2t + 1
    const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
    const int oy0 = (r / tiles_x) * kTH, ox0 = (r % tiles_x) * kTW;
    const float b_lo = bias[mw * 16 + g], b_hi = bias[mw * 16 + g + 8];
    float s_lo = 0.f, s_hi = 0.f;
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
      const int q = nh * 4 + nn, oy = oy0 + (q >> 1), oxb = ox0 + (q & 1) * 8 + 2 * t;
      if (oy < Ho) {
        if (oxb < Wo) { s_lo += fmaxf(acc[nn][0] + b_lo, 0.f); s_hi += fmaxf(acc[nn][2] + b_hi, 0.f); }
        if (oxb + 1 < Wo) { s_lo += fmaxf(acc[nn][1] + b_lo, 0.f); s_hi += fmaxf(acc[nn][3] + b_hi, 0.f); }
      }
    }
    s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 1);
    s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 2);
    s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 1);
    s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 2);
    float (*rd)[64] = red[lt & 1];       // two copies: a warp would have to run two tiles ahead of warps 0 / 1 to overwrite a live one,
    if (t == 0) {                        // and the four-stage ring keeps all eight warps within one tile of each other
      rd[nh][mw * 16 + g] = s_lo;
      rd[nh][mw * 16 + g + 8] = s_hi;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid < 64) partial[((size_t)b * tiles_per_img + r) * 64 + tid] = rd[0][tid] + rd[1][tid];
  }
}

// v_max[b][o] = ReLU(b3[o] + sum_c W3[o][c] * mean[c]),  mean[c] = (sum over tiles, fixed order) / (Ho Wo)
__global__ void __launch_bounds__(64) lra_logit_finish_kernel(const float *__restrict__ partial, const float *__restrict__ w3,
                                                              const float *__restrict__ b3, float *__restrict__ vmax, int tiles_per_img,
                                                              float inv_count) {
  __shared__ float mean[64];
  const int b = blockIdx.x, c = threadIdx.x;
  float s = 0.f;
  for (int tl = 0; tl < tiles_per_img; ++tl) s += partial[((size_t)b * tiles_per_img + tl) * 64 + c];
  mean[c] = s * inv_count;
  __syncthreads();
  float acc = b3 ? b3[c] : 0.f;
#pragma unroll 8
  for (int k = 0; k < 64; ++k) acc = fmaf(w3[c * 64 + k], mean[k], acc);
  vmax[(size_t)b * 64 + c] = fmaxf(acc, 0.f);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace lml
}  // namespace cdfo

using namespace cdfo;

static bool g_logit_tma = true;     // cdfo_lra_set_logit_tma: A/B switch (tests, tools)
extern "C" int cdfo_lra_set_logit_tma(int on) {
  g_logit_tma = on != 0;
  return CDFO_OK;
}

extern "C" size_t cdfo_lra_mask_logits_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  const int Ho = (H + 1) / 2 + 1, Wo = (W + 1) / 2 + 1;
  return (size_t)B * ceil_div(Ho, lml::kTH) * ceil_div(Wo, lml::kTW) * 64 * sizeof(float);
}

extern "C" int cdfo_lra_mask_logits_fwd(const float *v, const float *w2, const float *b2, const float *w3, const float *b3, float *vmax,
                                        void *workspace, int B, int H, int W, void *stream) {
  CDFO_REQUIRE(v && w2 && b2 && w3 && vmax && workspace, CDFO_ERR_NULL, "cdfo_lra_mask_logits_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && H > 0 && W > 0, CDFO_ERR_SHAPE, "cdfo_lra_mask_logits_fwd: bad shape");
  const int Ho = (H + 1) / 2 + 1, Wo = (W + 1) / 2 + 1;                     // (H + 2 * 2 - 3) / 2 + 1
  const int tiles_x = ceil_div(Wo, lml::kTW), tiles_per_img = tiles_x * ceil_div(Ho, lml::kTH);
  const long long nt = (long long)B * tiles_per_img;
  CDFO_REQUIRE(nt < (1ll << 31), CDFO_ERR_UNSUPPORTED, "cdfo_lra_mask_logits_fwd: too many tiles");
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(lml::lra_logit_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lml::kSmem);
    if (e != cudaSuccess) return fail(CDFO_ERR_CUDA, "cudaFuncSetAttribute(lra_logit_conv): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (int)(nt < kNumSMs ? nt : kNumSMs);
  bool tma_done = false;
  if (g_logit_tma && W % 4 == 0 && ((uintptr_t)v & 15) == 0 && (long long)B * 64 < (1ll << 31)) {
    static lml::EncodeTiledFn enc = lml::encode_tiled_fn();
    static bool tattr = false;
    if (enc && !tattr) {
      cudaError_t e = cudaFuncSetAttribute(lml::lra_logit_conv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lml::kTSmem);
      if (e != cudaSuccess) return fail(CDFO_ERR_CUDA, "cudaFuncSetAttribute(lra_logit_conv_tma): %s", cudaGetErrorString(e));
      tattr = true;
    }
    if (enc) {
      CUtensorMap vm;
      const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * 64};
      const cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
      const cuuint32_t box[3] = {lml::kTCols, lml::kTRows, lml::kTChunk};
      const cuuint32_t estr[3] = {1, 1, 1};
      CUresult cr = enc(&vm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(v), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CDFO_REQUIRE(cr == CUDA_SUCCESS, CDFO_ERR_CUDA, "cuTensorMapEncodeTiled(v) failed with CUresult %d", (int)cr);
      lml::lra_logit_conv_tma_kernel<<<grid, lml::kTThreads, lml::kTSmem, s>>>(vm, w2, b2, (float *)workspace, B, Ho, Wo, tiles_x, tiles_per_img);
      tma_done = true;
    }
  }
  if (!tma_done)
    lml::lra_logit_conv_kernel<<<grid, lml::kThreads, lml::kSmem, s>>>(v, w2, b2, (float *)workspace, B, H, W, Ho, Wo, tiles_x, tiles_per_img);
  lml::lra_logit_finish_kernel<<<B, 64, 0, s>>>((const float *)workspace, w3, b3, vmax, tiles_per_img, 1.0f / (float)((long long)Ho * Wo));
  return check_launch("cdfo_lra_mask_logits_fwd");
}
