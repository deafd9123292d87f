// Modulated deformable convolution (DCNv2) forward at the model's hot shape as ONE sm_100a kernel:
// C = Co = 64, 3x3, stride = pad = dil = 1, groups = 1, dg | 16 (the model: dg = 16).
//
// Semantics: ops/dcn/src/deform_conv_cuda_kernel.cu:467-496,570-632 + deform_conv_cuda.cpp:486-564
// (= torchvision.ops.deform_conv2d at arch/SIDECVSR_our.py:3352), optionally with the decoded MV prior added
// to the learned offset residual inside the kernel (arch/SIDECVSR_our.py:3347: offset + flow.flip(1).repeat(...)).
//
// Implicit GEMM, no im2col buffer in HBM:   D[128 px, 64 co] = sum_tap A_tap[128 px, 64 ci] * W_tap[64 co, 64 ci]^T
//   * persistent CTAs (one per SM) walk 4x32-pixel tiles (one warp lane per column: offset/mask reads are
//     full 128-byte lines, and the gather footprint of a tile fits the L1 that is left beside shared memory);
//   * producer warps: per (pixel, tap, channel quad) read offset/mask (streaming, prefetched one tap ahead into
//     registers), compute the fp32 sample position exactly like the reference, gather the 4 corners (8 bytes =
//     4 bf16 channels each) from the zero-bordered quad-planar input x_q4p, blend in fp32, and write the bf16
//     A operand straight into shared memory in the tcgen05 canonical K-major layout (ring of 4 tap stages);
//     the reference's inside test + per-corner zero padding are realised by clamping the sample position to
//     [-1, H] x [-1, W] and reading a zero border (1 pixel before, 2 after), which is value-identical;
//   * one thread issues tcgen05.mma (M128 N64 K16, bf16 -> fp32 in TMEM), 4 per tap, 36 per tile, and commits
//     to mbarriers (stage free / accumulator ready);
//   * 4 consumer warps: tcgen05.ld, + bias, coalesced store (NCHW fp32 or c8 bf16).
// The weights (64 x 576 bf16 = 72 KB) are fetched once per CTA with bulk async copies.
#include "cdfo_common.cuh"
#include "sm100_ptx.cuh"

namespace cdfo {

constexpr int kTileM = 128;
constexpr int kTileH = 4, kTileW = 32;     // kTileM pixels as 4 rows x 32 columns
constexpr int kStages = 4;
constexpr int kEpiWarps = 4;
constexpr int kWBytes = 9 * 64 * 64 * 2;       // 73728
constexpr int kABytes = kTileM * 64 * 2;       // 16384 per tap stage
constexpr int kALbo = kTileM * 16;             // K-adjacent core matrices of A
constexpr int kBLbo = 64 * 16;                 // K-adjacent core matrices of W
constexpr int kSbo = 128;                      // 8-row groups are contiguous
constexpr int kTapWBytes = 64 * 64 * 2;        // 8192 per tap

struct DcnSm100Params {
  const uint2 *x;      // [B][16][H+3][W+3] channel quads (4 bf16 = 8 bytes), zero border 1 before / 2 after
  const void *offset;  // [B][dg*18][H*W]
  const void *mask;    // [B][dg*9][H*W]
  const float *mv;     // [B][2][H*W] (x, y) or nullptr
  const uint8_t *wpk;  // [9][8][64][8] bf16
  const float *bias;   // [64] or nullptr
  void *y;
  int B, H, W, dg, out_mode;
  int gshift;  // deformable group of channel quad q is q >> gshift
  int x_batch;               // x holds x_batch samples; output sample b reads x[b % x_batch]
  long long off_bstride, msk_bstride;  // elements between consecutive samples of offset / mask
  int tiles_x, tiles_per_img, num_tiles;
};

constexpr size_t dcn_sm100_smem_bytes() { return kWBytes + kStages * kABytes + 256 + 16 * 8 + 16; }

// Tag for offset/mask given as packed fields [B][9 taps][dg/gp][H*W][gp] x fp16x4 (dy, dx, mask, 0) -- what the fused head writes.
struct FieldsH4 { uint2 v; };

template <typename OffT> __device__ __forceinline__ float ld_stream(const OffT *p);
template <> __device__ __forceinline__ float ld_stream<float>(const float *p) { return __ldcs(p); }
template <> __device__ __forceinline__ float ld_stream<__half>(const __half *p) {
  return __half2float(__ushort_as_half(__ldcs(reinterpret_cast<const unsigned short *>(p))));
}

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}

// (b, h0, w0) of a tile
struct TileCoord { int b, h0, w0; };
__device__ __forceinline__ TileCoord tile_coord(const DcnSm100Params &p, int tile) {
  TileCoord t;
  t.b = tile / p.tiles_per_img;
  const int r = tile - t.b * p.tiles_per_img;
  const int ty = r / p.tiles_x;
  t.h0 = ty * kTileH;
  t.w0 = (r - ty * p.tiles_x) * kTileW;
  return t;
}

template <typename OffT, int kProdWarps>
__global__ void __launch_bounds__((kEpiWarps + kProdWarps) * 32, 1) dcn_sm100_kernel(const DcnSm100Params p) {
  constexpr int kProdThreads = kProdWarps * 32;
  constexpr int kQuads = 16 / (kProdThreads / kTileM);  // channel quads handled per thread per tap (4)
  constexpr bool kPacked = sizeof(OffT) == sizeof(uint2);
  static_assert(kQuads == 4, "producer mapping assumes 512 producer threads");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *wsm = smem;
  uint8_t *asmem = smem + kWBytes;
  float *bias_s = reinterpret_cast<float *>(smem + kWBytes + kStages * kABytes);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kWBytes + kStages * kABytes + 256);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 16);
  // barrier map: [0,4) A stage full, [4,8) A stage empty, 8 accumulator full, 12 weights
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = p.H * p.W;

  if (tid < 64) bias_s[tid] = p.bias ? p.bias[tid] : 0.f;
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init(BAR(s), kProdThreads);
        ptx::mbar_init(BAR(4 + s), 1);
      }
      ptx::mbar_init(BAR(8), 1);
      ptx::mbar_init(BAR(12), 1);
      ptx::fence_mbar_init();
      ptx::mbar_arrive_expect_tx(BAR(12), kWBytes);
      for (int t = 0; t < 9; ++t)
        ptx::bulk_g2s(ptx::smem_u32(wsm) + t * kTapWBytes, p.wpk + t * kTapWBytes, kTapWBytes, BAR(12));
    }
    __syncwarp();
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarps) {
    // ============ consumers: warp 0 lane 0 issues the MMAs of a tile, then all 4 warps drain TMEM ============
    // (the MMA stream of a tile is ~1.2k cycles, the producers need several times that: sharing a warp between
    //  issue and epilogue costs nothing and keeps the CTA at 20 warps = 96 registers per thread)
    const uint32_t idesc = ptx::make_idesc_bf16(kTileM, 64);
    if (warp == 0) ptx::mbar_wait(BAR(12), 0);  // weights have landed (written by the async proxy)
    int stage = 0, phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      if (warp == 0) {
        for (int tap = 0; tap < 9; ++tap) {
          ptx::mbar_wait(BAR(stage), phase);  // producers filled this A stage
          ptx::tc_fence_after();
          if (lane == 0) {
            const uint32_t a0 = ptx::smem_u32(asmem) + stage * kABytes;
            const uint32_t b0 = ptx::smem_u32(wsm) + tap * kTapWBytes;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t ad = ptx::make_smem_desc(a0 + j * 2 * kALbo, kALbo, kSbo);
              const uint64_t bd = ptx::make_smem_desc(b0 + j * 2 * kBLbo, kBLbo, kSbo);
              ptx::umma_f16(tmem_base, ad, bd, idesc, (tap | j) != 0);
            }
            ptx::umma_commit(BAR(4 + stage));          // stage reusable once these MMAs retire
            if (tap == 8) ptx::umma_commit(BAR(8));    // accumulator complete
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      const TileCoord tc = tile_coord(p, tile);
      const int h = tc.h0 + warp, w = tc.w0 + lane;  // accumulator row = warp * 32 + lane = ty * 32 + tx
      const bool live = h < p.H && w < p.W;
      const int pix = h * p.W + w;
      ptx::mbar_wait(BAR(8), acc_phase);
      acc_phase ^= 1;
      ptx::tc_fence_after();
      uint32_t r[32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        ptx::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + half * 32, r);
        ptx::tmem_ld_wait();
        if (half == 1) {
          // every consumer thread has its accumulator rows in registers before warp 0 may overwrite TMEM
          ptx::tc_fence_before();
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        if (live) {
          if (p.out_mode == 0) {
            float *y = reinterpret_cast<float *>(p.y) + ((size_t)tc.b * 64 + half * 32) * P + pix;
#pragma unroll
            for (int n = 0; n < 32; ++n) y[(size_t)n * P] = __uint_as_float(r[n]) + bias_s[half * 32 + n];
          } else {
            uint4 *y = reinterpret_cast<uint4 *>(p.y) + ((size_t)tc.b * 8 + half * 4) * P + pix;
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
              uint4 v;
              const float *bs = bias_s + half * 32 + kc * 8;
              v.x = pack_bf2(__uint_as_float(r[kc * 8 + 0]) + bs[0], __uint_as_float(r[kc * 8 + 1]) + bs[1]);
              v.y = pack_bf2(__uint_as_float(r[kc * 8 + 2]) + bs[2], __uint_as_float(r[kc * 8 + 3]) + bs[3]);
              v.z = pack_bf2(__uint_as_float(r[kc * 8 + 4]) + bs[4], __uint_as_float(r[kc * 8 + 5]) + bs[5]);
              v.w = pack_bf2(__uint_as_float(r[kc * 8 + 6]) + bs[6], __uint_as_float(r[kc * 8 + 7]) + bs[7]);
              y[(size_t)kc * P] = v;
            }
          }
        }
      }
    }
  } else {
    // =========================== producers: offsets -> gather -> bilinear -> A operand ===========================
    const OffT *__restrict__ offset = reinterpret_cast<const OffT *>(p.offset);
    const OffT *__restrict__ mask = reinterpret_cast<const OffT *>(p.mask);
    const int ptid = tid - kEpiWarps * 32;
    const int row = ptid & (kTileM - 1);       // A-operand row = ty * 32 + tx
    const int ty = row >> 5, tx = row & 31;
    const int quad0 = (ptid / kTileM) * kQuads;  // this thread's 4 channel quads = c8 chunks quad0/2, quad0/2 + 1
    const int Wp = p.W + 3;
    const int plane_q = (p.H + 3) * Wp;  // quads per (b, quad) plane
    const float Hf = (float)p.H, Wf = (float)p.W, Wpf = (float)Wp;
    int goff[kQuads], gmsk[kQuads];  // element offsets of tap 0 of each quad's deformable group
#pragma unroll
    for (int qi = 0; qi < kQuads; ++qi) {
      const int g = (quad0 + qi) >> p.gshift;
      goff[qi] = g * 18 * P;
      gmsk[qi] = g * 9 * P;
    }

    // per-tile state of the pixel this thread owns
    struct PixState { const OffT *off; const OffT *msk; const uint2 *x; float hb, wb, mvx, mvy, live; };
    auto pix_state = [&](int tile) {
      PixState s;
      const TileCoord tc = tile_coord(p, tile);
      const int h = tc.h0 + ty, w = tc.w0 + tx;
      const bool live = h < p.H && w < p.W;
      const int pixc = min(h, p.H - 1) * p.W + min(w, p.W - 1);  // ragged tiles: clamp the address, zero the mask
      s.off = offset + (size_t)tc.b * p.off_bstride + (kPacked ? (size_t)pixc * (p.dg == 16 ? 2 : 1) : (size_t)pixc);
      s.msk = mask + (size_t)tc.b * p.msk_bstride + pixc;
      s.x = p.x + ((size_t)(tc.b % p.x_batch) * 16 + quad0) * plane_q + Wp + 1;  // + border shift
      s.hb = (float)(h - 1);
      s.wb = (float)(w - 1);
      s.live = live ? 1.f : 0.f;
      s.mvx = p.mv ? __ldg(p.mv + ((size_t)tc.b * 2 + 0) * P + pixc) : 0.f;
      s.mvy = p.mv ? __ldg(p.mv + ((size_t)tc.b * 2 + 1) * P + pixc) : 0.f;
      return s;
    };
    auto load_tap = [&](const PixState &s, int tap, float (&o)[kQuads * 3]) {
#pragma unroll
      for (int qi = 0; qi < kQuads; ++qi) {
        if constexpr (kPacked) {
          // fields [B][9 taps][dg/gp][H*W][gp] x (dy, dx, m, 0), gp = 2 for dg = 16 else 1; s.off points at (b, pixel * gp)
          const int g = (quad0 + qi) >> p.gshift, gp = p.dg == 16 ? 2 : 1;
          const uint2 raw = __ldcs(reinterpret_cast<const uint2 *>(s.off) + ((size_t)tap * (p.dg / gp) + g / gp) * P * gp + g % gp);
          const float2 d = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
          o[qi * 3 + 0] = d.x;
          o[qi * 3 + 1] = d.y;
          o[qi * 3 + 2] = __low2float(*reinterpret_cast<const __half2 *>(&raw.y));
        } else {
          const OffT *po = s.off + goff[qi] + tap * 2 * P;
          o[qi * 3 + 0] = ld_stream(po);
          o[qi * 3 + 1] = ld_stream(po + P);
          o[qi * 3 + 2] = ld_stream(s.msk + gmsk[qi] + tap * P);
        }
      }
    };

    int stage = 0, phase = 0;
    int tile = blockIdx.x;
    if (tile < p.num_tiles) {
      PixState cur = pix_state(tile);
      float o_cur[kQuads * 3], o_nxt[kQuads * 3];
      load_tap(cur, 0, o_cur);
      while (true) {
        const int next_tile = tile + gridDim.x;
        PixState nxt = cur;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          // ---- prefetch the next tap's (or next tile's first tap's) offsets / mask into registers
          if (tap < 8) {
            load_tap(cur, tap + 1, o_nxt);
          } else if (next_tile < p.num_tiles) {
            nxt = pix_state(next_tile);
            load_tap(nxt, 0, o_nxt);
          }
          const int ti = tap / 3;
          const float hb = cur.hb + (float)ti, wb = cur.wb + (float)(tap - 3 * ti);
          ptx::mbar_wait(BAR(4 + stage), phase ^ 1);  // MMA released this stage
          uint8_t *a_dst = asmem + stage * kABytes + (quad0 >> 1) * kALbo + row * 16;
#pragma unroll
          for (int pair = 0; pair < kQuads / 2; ++pair) {
            uint2 v[2][4];
            float wgt[2][4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int qi = pair * 2 + e;
              // reference order: offset = residual + flow (arch :3347), then h_im = base + offset (.cu:614-615)
              const float h_im = __fadd_rn(hb, __fadd_rn(o_cur[qi * 3 + 0], cur.mvy));
              const float w_im = __fadd_rn(wb, __fadd_rn(o_cur[qi * 3 + 1], cur.mvx));
              const float m = o_cur[qi * 3 + 2] * cur.live;
              // inside test + corner zero padding == clamp to [-1, H] x [-1, W] + zero border (NaN -> -1 -> 0)
              const float hc = fminf(fmaxf(h_im, -1.f), Hf), wc = fminf(fmaxf(w_im, -1.f), Wf);
              const float hf = floorf(hc), wf = floorf(wc);
              const float lh = hc - hf, lw = wc - wf;
              const float a = m - m * lh, bb = m * lh;
              wgt[e][0] = a - a * lw; wgt[e][1] = a * lw; wgt[e][2] = bb - bb * lw; wgt[e][3] = bb * lw;
              const int idx = (int)fmaf(hf, Wpf, wf);  // exact: integers < 2^24
              const uint2 *q = cur.x + (size_t)qi * plane_q + idx;
              v[e][0] = __ldg(q); v[e][1] = __ldg(q + 1); v[e][2] = __ldg(q + Wp); v[e][3] = __ldg(q + Wp + 1);
            }
            uint32_t out[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              float c0 = wgt[e][0] * bf_lo(v[e][0].x), c1 = wgt[e][0] * bf_hi(v[e][0].x);
              float c2 = wgt[e][0] * bf_lo(v[e][0].y), c3 = wgt[e][0] * bf_hi(v[e][0].y);
#pragma unroll
              for (int k = 1; k < 4; ++k) {
                c0 = fmaf(wgt[e][k], bf_lo(v[e][k].x), c0); c1 = fmaf(wgt[e][k], bf_hi(v[e][k].x), c1);
                c2 = fmaf(wgt[e][k], bf_lo(v[e][k].y), c2); c3 = fmaf(wgt[e][k], bf_hi(v[e][k].y), c3);
              }
              out[e * 2 + 0] = pack_bf2(c0, c1);
              out[e * 2 + 1] = pack_bf2(c2, c3);
            }
            *reinterpret_cast<uint4 *>(a_dst + pair * kALbo) = make_uint4(out[0], out[1], out[2], out[3]);
          }
          ptx::fence_proxy_async_smem();  // my generic-proxy stores -> visible to tcgen05.mma (async proxy)
          ptx::mbar_arrive(BAR(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
#pragma unroll
          for (int i = 0; i < kQuads * 3; ++i) o_cur[i] = o_nxt[i];
        }
        if (next_tile >= p.num_tiles) break;
        tile = next_tile;
        cur = nxt;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 64);
}

// W [64 co][64 ci][3][3] fp32 -> [tap][kc][co][8] bf16 (B operand, canonical K-major core matrices)
__global__ void dcn_sm100_pack_weight_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 9 * 64 * 64) return;
  const int j = e % 8, co = (e / 8) % 64, kc = (e / 512) % 8, tap = e / 4096;
  out[e] = __float2bfloat16_rn(w[((size_t)co * 64 + kc * 8 + j) * 9 + tap]);
}

// NCHW fp32 -> [B][C/4][H+3][W+3] channel quads (4 bf16) with a zero border: 1 pixel before, 2 after
__global__ void pack_q4p_kernel(const float *__restrict__ x, uint2 *__restrict__ out, int C, int H, int W) {
  const int Wp = W + 3, Hp = H + 3;
  const int pp = blockIdx.x * blockDim.x + threadIdx.x;
  if (pp >= Hp * Wp) return;
  const int q = blockIdx.y, b = blockIdx.z;
  const int h = pp / Wp - 1, w = pp % Wp - 1;
  uint2 v = make_uint2(0, 0);
  if (h >= 0 && h < H && w >= 0 && w < W) {
    const float *src = x + (((size_t)b * C + q * 4) * H + h) * W + w;
    const size_t HW = (size_t)H * W;
    v.x = pack_bf2(src[0], src[HW]);
    v.y = pack_bf2(src[2 * HW], src[3 * HW]);
  }
  out[((size_t)b * (C / 4) + q) * Hp * Wp + pp] = v;
}

}  // namespace cdfo

using namespace cdfo;

extern "C" int cdfo_dcn_sm100_pack_weight(const float *w, void *wpk, void *stream) {
  CDFO_REQUIRE(w && wpk, CDFO_ERR_NULL, "cdfo_dcn_sm100_pack_weight: NULL pointer");
  dcn_sm100_pack_weight_kernel<<<ceil_div(9 * 64 * 64, 256), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16 *)wpk);
  return check_launch("cdfo_dcn_sm100_pack_weight");
}

extern "C" int cdfo_pack_q4p(const float *x_nchw, void *x_q4p, int B, int C, int H, int W, void *stream) {
  CDFO_REQUIRE(x_nchw && x_q4p, CDFO_ERR_NULL, "cdfo_pack_q4p: NULL pointer");
  CDFO_REQUIRE(B > 0 && B <= 65535 && C > 0 && C % 4 == 0 && H > 0 && W > 0, CDFO_ERR_SHAPE, "cdfo_pack_q4p: bad shape");
  dim3 grid(ceil_div((H + 3) * (W + 3), 128), C / 4, B);
  pack_q4p_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x_nchw, (uint2 *)x_q4p, C, H, W);
  return check_launch("cdfo_pack_q4p");
}

template <typename OffT, int PW>
static int launch_dcn_sm100(const DcnSm100Params &p, int grid, cudaStream_t s) {
  auto kern = dcn_sm100_kernel<OffT, PW>;
  static bool attr_done = false;  // per instantiation
  const size_t smem = dcn_sm100_smem_bytes();
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(CDFO_ERR_CUDA, "cudaFuncSetAttribute(dcn_sm100): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  kern<<<grid, (kEpiWarps + PW) * 32, smem, s>>>(p);
  return check_launch("cdfo_dcn_sm100_fwd");
}

extern "C" int cdfo_dcn_sm100_fwd(const void *x_q4p, const void *offset, const void *mask, const float *mv,
                                  const void *wpk, const float *bias, void *y, int B, int H, int W, int dg,
                                  int off_dtype, int out_mode, int num_ctas, int x_batch, long long off_bstride,
                                  long long msk_bstride, void *stream) {
  CDFO_REQUIRE(x_q4p && offset && (mask || off_dtype == CDFO_FIELDS_F16X4) && wpk && y, CDFO_ERR_NULL, "cdfo_dcn_sm100_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && H > 0 && W > 0, CDFO_ERR_SHAPE, "cdfo_dcn_sm100_fwd: bad shape");
  CDFO_REQUIRE(dg == 1 || dg == 2 || dg == 4 || dg == 8 || dg == 16, CDFO_ERR_UNSUPPORTED,
               "cdfo_dcn_sm100_fwd: deformable groups must divide 16 (got %d)", dg);
  CDFO_REQUIRE(out_mode == 0 || out_mode == 1, CDFO_ERR_UNSUPPORTED, "cdfo_dcn_sm100_fwd: out_mode %d", out_mode);
  CDFO_REQUIRE(((uintptr_t)wpk & 15) == 0 && ((uintptr_t)x_q4p & 7) == 0 && ((uintptr_t)y & 15) == 0, CDFO_ERR_SHAPE,
               "cdfo_dcn_sm100_fwd: wpk, y must be 16-byte aligned (x 8-byte)");
  DcnSm100Params p;
  p.x = (const uint2 *)x_q4p; p.offset = offset; p.mask = mask; p.mv = mv; p.wpk = (const uint8_t *)wpk;
  p.bias = bias; p.y = y; p.B = B; p.H = H; p.W = W; p.dg = dg; p.out_mode = out_mode;
  p.x_batch = x_batch > 0 ? x_batch : B;
  p.off_bstride = off_bstride > 0 ? off_bstride : (long long)dg * 18 * H * W;
  p.msk_bstride = msk_bstride > 0 ? msk_bstride : (long long)dg * 9 * H * W;
  CDFO_REQUIRE(B % p.x_batch == 0, CDFO_ERR_SHAPE, "cdfo_dcn_sm100_fwd: B (%d) must be a multiple of x_batch (%d)", B, p.x_batch);
  p.gshift = 0;
  while ((16 >> p.gshift) > dg) ++p.gshift;  // quads per deformable group = 16 / dg
  p.tiles_x = ceil_div(W, kTileW);
  p.tiles_per_img = p.tiles_x * ceil_div(H, kTileH);
  CDFO_REQUIRE((long long)dg * 18 * H * W < (1ll << 31), CDFO_ERR_UNSUPPORTED, "cdfo_dcn_sm100_fwd: offset planes too large for 32-bit indexing");
  const long long nt = (long long)p.tiles_per_img * B;
  CDFO_REQUIRE(nt < (1ll << 31), CDFO_ERR_UNSUPPORTED, "cdfo_dcn_sm100_fwd: too many tiles");
  p.num_tiles = (int)nt;
  int grid = num_ctas > 0 ? num_ctas : kNumSMs;
  if (grid > p.num_tiles) grid = p.num_tiles;
  cudaStream_t s = (cudaStream_t)stream;
  if (off_dtype == CDFO_F32) return launch_dcn_sm100<float, 16>(p, grid, s);
  if (off_dtype == CDFO_F16) return launch_dcn_sm100<__half, 16>(p, grid, s);
  if (off_dtype == CDFO_FIELDS_F16X4) {
    if (off_bstride <= 0) p.off_bstride = (long long)dg * 9 * H * W;
    return launch_dcn_sm100<FieldsH4, 16>(p, grid, s);
  }
  return fail(CDFO_ERR_UNSUPPORTED, "cdfo_dcn_sm100_fwd: offset dtype must be fp32, fp16 or packed fp16x4 fields");
}
