// 3x3 / stride 1 / padding 1 convolution as a tcgen05 implicit GEMM (sm_100a), channel-chunked bf16 activations.
//
// Replaces the cuDNN calls behind the reference's nn.Conv2d(…, 3, 1, 1) layers on the hot path:
//   conv_offset.0 (64->64) and conv_offset.2 (64->432) of MVDualAttAlignment   arch/SIDECVSR_our.py:3271-3275
//   ResidualBlock_noBN.conv1/conv2 (64->64) of DualAttAlignment                arch/SIDECVSR_our.py:254-271, :3452-3453
//   conv_expand_fea_r (128->64)                                                arch/SIDECVSR_our.py:4382, :4454
// (and, as the "next" row, the trunk's 64->256 / 256->64 pairs, arch:378-406).
//
// D[128 px, NT co] = sum over (64-channel block kb, tap) of A_{kb,tap}[128 px, 64 ci] * W_{kb,tap}[NT co, 64 ci]^T
//   * activations live in HBM as "c8" = [B][C/8][H][W][8] bf16.  One TMA 5-D box per (tile, kb) lands the
//     (16+2) x (8+2) pixel halo of 64 channels in shared memory as [ci/8][18][10][8]: exactly the tcgen05
//     canonical K-major layout (8 pixels x 16 bytes = one 128-byte core matrix), with the convolution's zero
//     padding supplied by TMA out-of-bounds fill.  The A operand of tap (i, j) is the SAME shared memory at a
//     byte offset of (i*10 + j)*16: no im2col, no data movement between taps (SBO = 160 B, LBO = 2880 B).
//   * a CTA is persistent, owns one N tile (NT output channels) whose weights stay resident in shared memory,
//     and walks 16x8-pixel M tiles; warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue (two per TMEM lane
//     quarter, alternating column chunks)
//     (tcgen05.ld -> +bias -> activation -> +residual -> bf16 c8 or fp32 NCHW); TMEM accumulator double buffered.
#include <cuda.h>
#include <stdlib.h>

#include "cdfo_common.cuh"
#include "mv_head_math.cuh"
#include "sm100_ptx.cuh"

namespace cdfo {

constexpr int kCvTileH = 16, kCvTileW = 8;                 // 128 output pixels
// geometry of a KS x KS convolution (KS = 3: 18 x 10 halo, 2880 B per 8-channel chunk, 160 B between 8-pixel rows; KS = 1: the tile itself)
template <int KS> struct CvGeom {
  static constexpr int kHaloH = kCvTileH + KS - 1, kHaloW = kCvTileW + KS - 1;
  static constexpr int kPlane = kHaloH * kHaloW * 16;
  static constexpr int kSbo = kHaloW * 16;
  static constexpr int kTaps = KS * KS;
};
constexpr int kCvThreads = 320;                            // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
constexpr int kCvEpiThreads = 256;

struct Conv3x3Params {
  const uint8_t *wpk;   // [n_tiles][9 taps][Cin/8][NT][8] bf16
  const float *bias;    // [Cout] or nullptr
  const uint4 *resid;   // c8 bf16 [B][Cout/8][H][W][8] or nullptr: added after the activation
  void *y;
  int B, Cin, Cout, H, W;
  int tma_wide;  // 1: the tensor map merges the pixel and channel axes (one request per halo row instead of one per 16-byte pixel chunk)
  int act;       // 0 none, 1 relu, 2 leaky relu 0.1
  int out_mode;  // 0: NCHW fp32, 1: c8 bf16, 2: c8 bf16 through PixelShuffle(2): output channels arrive ordered
                 //    n' = (2i+j)*(Cout/4) + c and leave at [B][Cout/32][2H][2W][8], pixel (2h+i, 2w+j)
  // epi 0: bias -> act -> +resid.
  // Offset/mask head of MVDualAttAlignment (arch/SIDECVSR_our.py:3341-3350): output channels arrive permuted as triples
  // (dy_k, dx_k, m_k), k' = tap*dg + g, and leave as one fp16x4 "field" (dy, dx, m, 0) per (tap, pixel, group):
  // y [B][9][dg/gp][H*W][gp] x 8 B, gp = 2 for dg = 16 else 1 (what the DCN producer reads with two coalesced 16-byte loads)
  // epi 1: (mag*tanh(dy), mag*tanh(dx), m)                                   (first head evaluation)
  // epi 2: (aux.dy + mag*tanh(dy), aux.dx + mag*tanh(dx), sigmoid(aux.m + m)) (second evaluation; aux = epi-1 output)
  // epi 3: channel 0 only (+ bias) + bilinear x4 skip of a 1-channel LR image (aux = float [B][H/4][W/4], align_corners=False):
  //        conv_last + base of the tail, arch/SIDECVSR_our.py:4477-4480; y = fp32 [B][1][H][W]
  int epi;
  float mag;
  const uint2 *aux;
  int n_tiles, tiles_x, tiles_y, m_tiles;
  // 1: the weights of the N tile do not fit shared memory next to the A stages (Cin >= 256 at NT = 64): the 9 x NT x 64
  // weights of each 64-channel block travel with the block's A halo through the same stage (L2-resident, re-read per tile)
  int stream_w;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ uint32_t cv_pack_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}

// DUAL (offset / mask head only, NT = 144): the input holds 2 B samples, sample b and b + B are the two hidden maps of
// MVDualAttAlignment (arch/SIDECVSR_our.py:3339-3350).  A CTA computes the SAME pixel tile of both back to back; the epilogue keeps
// the first evaluation (mag tanh(dy1), mag tanh(dx1), m1) in registers, rounded to fp16 exactly as the two-launch path stores it, and
// combines it with the second: the intermediate fields (1152 B per pixel written and read again) never reach HBM.
template <int NT, int kStages, int KBLK, int KS, bool DUAL = false>
__global__ void __launch_bounds__(kCvThreads, 1)
conv3x3_sm100_kernel(const __grid_constant__ CUtensorMap tmap, const Conv3x3Params p) {
  static_assert(!DUAL || NT == 144, "the dual head uses the 144-channel N tile");
  constexpr int kEvals = DUAL ? 2 : 1;
  constexpr int kAccCols = NT <= 32 ? 32 : (NT <= 64 ? 64 : (NT <= 128 ? 128 : 256));  // per accumulator buffer
  constexpr int kTmemCols = 2 * kAccCols;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int KB = p.Cin / KBLK;                  // K blocks of KBLK channels (one pipeline stage each)
  using G = CvGeom<KS>;
  constexpr int kTaps = G::kTaps;
  const int w_bytes = kTaps * p.Cin * NT * 2;
  constexpr int kChunks = KBLK / 8;             // 8-channel chunks per stage
  constexpr int kABytes = kChunks * G::kPlane;  // activation halo of one stage
  constexpr int kPiece = NT * KBLK * 2;         // weights of one (tap, K block)
  const int stage_stride = p.stream_w ? ((kABytes + kTaps * kPiece + 1023) & ~1023) : kABytes;
  uint8_t *wsm = smem;
  uint8_t *asmem = smem + (p.stream_w ? 0 : ((w_bytes + 1023) & ~1023));
  uint64_t *bars = reinterpret_cast<uint64_t *>(asmem + kStages * stage_stride);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 16);
  float *bias_s = reinterpret_cast<float *>(bars + 18);
  // barrier map: [0,kStages) A full, [4,4+kStages) A empty, 8/9 accumulator full, 10/11 accumulator empty, 12 weights
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_first = blockIdx.x / p.n_tiles, m_step = gridDim.x / p.n_tiles;
  const int n0 = n_tile * NT;

  for (int i = tid; i < NT; i += kCvThreads) bias_s[i] = (p.bias && n0 + i < p.Cout) ? p.bias[n0 + i] : 0.f;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(BAR(s), 1);
      ptx::mbar_init(BAR(4 + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(BAR(8 + a), 1);
      ptx::mbar_init(BAR(10 + a), kCvEpiThreads);
    }
    ptx::mbar_init(BAR(12), 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tmap);
    if (!p.stream_w) {
      // resident weights of this N tile: taps * KB pieces of NT x 64 bf16
      ptx::mbar_arrive_expect_tx(BAR(12), w_bytes);
      const uint8_t *src = p.wpk + (size_t)n_tile * w_bytes;
      for (int i = 0; i < kTaps * KB; ++i) ptx::bulk_g2s(ptx::smem_u32(wsm) + i * kPiece, src + (size_t)i * kPiece, kPiece, BAR(12));   // [tap][Cin/8][NT][8] as is
    }
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    int stage = 0, phase = 0;
    for (int mt = m_first; mt < p.m_tiles; mt += m_step) {
      const int b = mt / (p.tiles_x * p.tiles_y);
      const int r = mt - b * p.tiles_x * p.tiles_y;
      const int h0 = (r / p.tiles_x) * kCvTileH, w0 = (r % p.tiles_x) * kCvTileW;
      for (int ev = 0; ev < kEvals; ++ev)
      for (int kb = 0; kb < KB; ++kb) {
        if (lane == 0) {
          ptx::mbar_wait(BAR(4 + stage), phase ^ 1);
          const uint32_t dst = ptx::smem_u32(asmem) + stage * stage_stride;
          ptx::mbar_arrive_expect_tx(BAR(stage), kABytes + (p.stream_w ? kTaps * kPiece : 0));
          const int bi = b + ev * p.B;       // DUAL: the second hidden map of the pair
          if (p.tma_wide) ptx::tma_load_5d(dst, &tmap, BAR(stage), (w0 - KS / 2) * 8, h0 - KS / 2, kb * kChunks, bi, 0);
          else ptx::tma_load_5d(dst, &tmap, BAR(stage), 0, w0 - KS / 2, h0 - KS / 2, kb * kChunks, bi);
          if (p.stream_w) {
            // streamed weights are packed [n_tile][K block][tap][chunk][NT][8]: one bulk copy per stage
            ptx::bulk_g2s(dst + kABytes, p.wpk + (size_t)n_tile * w_bytes + (size_t)kb * kTaps * kPiece, kTaps * kPiece, BAR(stage));
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = ptx::make_idesc_bf16(128, NT);
    if (!p.stream_w) ptx::mbar_wait(BAR(12), 0);
    int stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int mt = m_first; mt < p.m_tiles; mt += m_step)
    for (int ev = 0; ev < kEvals; ++ev) {
      ptx::mbar_wait(BAR(10 + acc), acc_phase ^ 1);
      ptx::tc_fence_after();
      for (int kb = 0; kb < KB; ++kb) {
        ptx::mbar_wait(BAR(stage), phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {       // one elected lane: no per-thread waterfall around the uniform-operand MMAs
          const uint32_t a0 = ptx::smem_u32(asmem) + stage * stage_stride;
          // descriptors once per stage; every MMA adds constants to the low words (at N <= 64 the issuing thread is the critical path)
          const uint64_t ad0 = ptx::make_smem_desc(a0, G::kPlane, G::kSbo);
          const uint64_t bd0 = ptx::make_smem_desc(p.stream_w ? a0 + kABytes : ptx::smem_u32(wsm) + kb * kPiece, NT * 16, 128);
          const uint32_t a_lo0 = (uint32_t)ad0, a_hi = (uint32_t)(ad0 >> 32), b_lo0 = (uint32_t)bd0, b_hi = (uint32_t)(bd0 >> 32);
          const uint32_t b_tap_step = (uint32_t)((p.stream_w ? 1 : KB) * kPiece) >> 4;
          const uint32_t tmem_d = tmem_base + acc * kAccCols;
#pragma unroll
          for (int tap = 0; tap < kTaps; ++tap) {
#pragma unroll
            for (int j = 0; j < KBLK / 16; ++j) {
              const uint32_t a_lo = a_lo0 + ((((tap / KS) * G::kHaloW + (tap % KS)) * 16 + j * 2 * G::kPlane) >> 4);
              const uint32_t b_lo = b_lo0 + tap * b_tap_step + ((j * 2 * (NT * 16)) >> 4);
              ptx::umma_f16_w(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, (tap | j) != 0 ? 1u : (uint32_t)(kb != 0));
            }
          }
          ptx::umma_commit(BAR(4 + stage));
          if (kb == KB - 1) ptx::umma_commit(BAR(8 + acc));
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // =========================== epilogue ===========================
    const int quarter = warp & 3;                 // TMEM lanes this warp may read
    const int ehalf = (warp - 2) >> 2;            // the two warps of a quarter take alternate column chunks
    const int row = quarter * 32 + lane;          // tile pixel: ty = row / 8, tx = row % 8
    const int ty = row >> 3, tx = row & 7;
    const size_t HW = (size_t)p.H * p.W;
    int acc = 0, acc_phase = 0;
    for (int mt = m_first; mt < p.m_tiles; mt += m_step) {
      const int b = mt / (p.tiles_x * p.tiles_y);
      const int r = mt - b * p.tiles_x * p.tiles_y;
      const int h = (r / p.tiles_x) * kCvTileH + ty, w = (r % p.tiles_x) * kCvTileW + tx;
      const bool live = h < p.H && w < p.W;
      const size_t pix = (size_t)h * p.W + w;
      if constexpr (DUAL) {
        uint2 stash[2][16];      // first evaluation of this thread's (up to) two 48-channel chunks, as the fp16 fields of epi 1
        const int dgn = p.Cout / 27;
        const bool vec = dgn == 16;
#pragma unroll
        for (int ev = 0; ev < 2; ++ev) {
          ptx::mbar_wait(BAR(8 + acc), acc_phase);
          ptx::tc_fence_after();
#pragma unroll
          for (int ci = 0; ci < 2; ++ci) {
            const int c0 = ehalf * 48 + ci * 96;
            if (c0 < NT) {
              uint32_t r0[16], r1[16], r2[16];
              const uint32_t ta = tmem_base + acc * kAccCols + ((uint32_t)(quarter * 32) << 16) + c0;
              tmem_ld16(ta, r0);
              tmem_ld16(ta + 16, r1);
              tmem_ld16(ta + 32, r2);
              ptx::tmem_ld_wait();
              if (c0 + 96 >= NT) {           // this warp's last read of the accumulator buffer
                ptx::tc_fence_before();
                ptx::mbar_arrive(BAR(10 + acc));
              }
              if (live) {
                float v[48];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  v[i] = __fadd_rn(__uint_as_float(r0[i]), bias_s[c0 + i]);
                  v[16 + i] = __fadd_rn(__uint_as_float(r1[i]), bias_s[c0 + 16 + i]);
                  v[32 + i] = __fadd_rn(__uint_as_float(r2[i]), bias_s[c0 + 32 + i]);
                }
                const int k0 = (n0 + c0) / 3;
                const size_t base16 = (((size_t)b * 9 + k0 / 16) * 8 * HW + pix) * 2;   // dg == 16: pair plane 0 of this tap
                uint2 outv[16];
#pragma unroll
                for (int t = 0; t < 16; ++t) {   // offset_1 + offset_2, sigmoid(mask_1 + mask_2): arch:3347,3350 (mv_head_math.cuh)
                  if (ev == 0) stash[ci][t] = head::first(v[3 * t], v[3 * t + 1], v[3 * t + 2], p.mag);
                  else outv[t] = head::second(stash[ci][t], v[3 * t], v[3 * t + 1], v[3 * t + 2], p.mag);
                }
                if (ev == 1) {
                  uint2 *y = reinterpret_cast<uint2 *>(p.y);
                  if (vec) {
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                      *reinterpret_cast<uint4 *>(y + base16 + (size_t)t * 2 * HW) = make_uint4(outv[2 * t].x, outv[2 * t].y, outv[2 * t + 1].x, outv[2 * t + 1].y);
                  } else {
#pragma unroll
                    for (int t = 0; t < 16; ++t) y[((size_t)b * 9 * dgn + k0 + t) * HW + pix] = outv[t];
                  }
                }
              }
            }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
        continue;
      }
      ptx::mbar_wait(BAR(8 + acc), acc_phase);
      ptx::tc_fence_after();
      if (p.epi == 3) {
        uint32_t rr[16];
        if (ehalf == 0) {
          tmem_ld16(tmem_base + acc * kAccCols + ((uint32_t)(quarter * 32) << 16), rr);
          ptx::tmem_ld_wait();
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(BAR(10 + acc));
        if (live && ehalf == 0) {
          // F.interpolate(scale_factor=4, bilinear, align_corners=False): src = max((dst + 0.5) / 4 - 0.5, 0)
          const int Hl = p.H >> 2, Wl = p.W >> 2;
          const float sy = fmaxf(((float)h + 0.5f) * 0.25f - 0.5f, 0.f), sx = fmaxf(((float)w + 0.5f) * 0.25f - 0.5f, 0.f);
          const int y0 = (int)sy, x0 = (int)sx;
          const int y1 = min(y0 + 1, Hl - 1), x1 = min(x0 + 1, Wl - 1);
          const float ly = sy - (float)y0, lx = sx - (float)x0;
          const float *lr = reinterpret_cast<const float *>(p.aux) + (size_t)b * Hl * Wl;
          const float base = (1.f - ly) * ((1.f - lx) * __ldg(lr + y0 * Wl + x0) + lx * __ldg(lr + y0 * Wl + x1)) +
                             ly * ((1.f - lx) * __ldg(lr + y1 * Wl + x0) + lx * __ldg(lr + y1 * Wl + x1));
          reinterpret_cast<float *>(p.y)[(size_t)b * HW + pix] = __uint_as_float(rr[0]) + bias_s[0] + base;
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      if (p.epi != 0) {
        if constexpr (NT % 48 == 0) {
          bool arrived = false;
#pragma unroll 1
          for (int c0 = ehalf * 48; c0 < NT; c0 += 96) {
            uint32_t r0[16], r1[16], r2[16];
            const uint32_t ta = tmem_base + acc * kAccCols + ((uint32_t)(quarter * 32) << 16) + c0;
            tmem_ld16(ta, r0);
            tmem_ld16(ta + 16, r1);
            tmem_ld16(ta + 32, r2);
            ptx::tmem_ld_wait();
            if (c0 + 96 >= NT) {           // this warp's last read of the accumulator buffer
              ptx::tc_fence_before();
              ptx::mbar_arrive(BAR(10 + acc));
              arrived = true;
            }
            if (!live) continue;
            float v[48];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v[i] = __fadd_rn(__uint_as_float(r0[i]), bias_s[c0 + i]);
              v[16 + i] = __fadd_rn(__uint_as_float(r1[i]), bias_s[c0 + 16 + i]);
              v[32 + i] = __fadd_rn(__uint_as_float(r2[i]), bias_s[c0 + 32 + i]);
            }
            // fields layout [B][9 taps][dg/gp][H*W][gp] (gp = 2 for dg == 16, else 1): triple k' = tap * dg + g.  The 16
            // consecutive triples of one thread are the 16 groups of one tap when dg == 16: 8 pair planes, 16 bytes each,
            // consecutive lanes = consecutive pixels (coalesced); otherwise they are placed one by one (8 bytes each).
            const int dgn = p.Cout / 27, k0 = (n0 + c0) / 3;
            uint2 prior[16], outv[16];
            const bool vec = dgn == 16;
            const size_t base16 = (((size_t)b * 9 + k0 / 16) * 8 * HW + pix) * 2;   // dg == 16: pair plane 0 of this tap
            if (p.epi == 2) {
              if (vec) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                  const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p.aux + base16 + (size_t)t * 2 * HW));
                  prior[2 * t] = make_uint2(q.x, q.y);
                  prior[2 * t + 1] = make_uint2(q.z, q.w);
                }
              } else {
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                  const int k = k0 + t;
                  prior[t] = __ldg(p.aux + ((size_t)b * 9 * dgn + k) * HW + pix);
                }
              }
            }
#pragma unroll
            for (int t = 0; t < 16; ++t)     // offset_1 + offset_2, sigmoid(mask_1 + mask_2): arch:3347,3350 (mv_head_math.cuh)
              outv[t] = p.epi == 2 ? head::second(prior[t], v[3 * t], v[3 * t + 1], v[3 * t + 2], p.mag)
                                   : head::first(v[3 * t], v[3 * t + 1], v[3 * t + 2], p.mag);
            uint2 *y = reinterpret_cast<uint2 *>(p.y);
            if (vec) {
#pragma unroll
              for (int t = 0; t < 8; ++t)
                *reinterpret_cast<uint4 *>(y + base16 + (size_t)t * 2 * HW) = make_uint4(outv[2 * t].x, outv[2 * t].y, outv[2 * t + 1].x, outv[2 * t + 1].y);
            } else {
#pragma unroll
              for (int t = 0; t < 16; ++t) {
                const int k = k0 + t;
                y[((size_t)b * 9 * dgn + k) * HW + pix] = outv[t];
              }
            }
          }
          if (!arrived) {
            ptx::tc_fence_before();
            ptx::mbar_arrive(BAR(10 + acc));
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      bool arrived = false;
#pragma unroll 1
      for (int c0 = ehalf * 16; c0 < NT; c0 += 32) {
        uint32_t rr[16];
        tmem_ld16(tmem_base + acc * kAccCols + ((uint32_t)(quarter * 32) << 16) + c0, rr);
        ptx::tmem_ld_wait();
        if (c0 + 32 >= NT) {  // this warp's last read of this accumulator buffer
          ptx::tc_fence_before();
          ptx::mbar_arrive(BAR(10 + acc));
          arrived = true;
        }
        if (!live || n0 + c0 >= p.Cout) continue;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float t = __uint_as_float(rr[i]) + bias_s[c0 + i];
          if (p.act == 1) t = fmaxf(t, 0.f);
          else if (p.act == 2) t = t > 0.f ? t : 0.1f * t;
          v[i] = t;
        }
        if (p.resid) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint4 q = __ldg(p.resid + ((size_t)b * (p.Cout / 8) + (n0 + c0) / 8 + half) * HW + pix);
            const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              v[half * 8 + 2 * i] += __uint_as_float(qq[i] << 16);
              v[half * 8 + 2 * i + 1] += __uint_as_float(qq[i] & 0xffff0000u);
            }
          }
        }
        if (p.out_mode == 0) {
          float *y = reinterpret_cast<float *>(p.y) + ((size_t)b * p.Cout + n0 + c0) * HW + pix;
#pragma unroll
          for (int i = 0; i < 16; ++i) y[(size_t)i * HW] = v[i];
        } else if (p.out_mode == 2) {
          const int cps = p.Cout >> 2;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int n = n0 + c0 + half * 8, q = n / cps, kc = (n - q * cps) >> 3;
            uint4 *y = reinterpret_cast<uint4 *>(p.y) +
                       (((size_t)b * (cps >> 3) + kc) * (2 * p.H) + 2 * h + (q >> 1)) * (size_t)(2 * p.W) + 2 * w + (q & 1);
            *y = make_uint4(cv_pack_bf2(v[half * 8 + 0], v[half * 8 + 1]), cv_pack_bf2(v[half * 8 + 2], v[half * 8 + 3]),
                            cv_pack_bf2(v[half * 8 + 4], v[half * 8 + 5]), cv_pack_bf2(v[half * 8 + 6], v[half * 8 + 7]));
          }
        } else {
          uint4 *y = reinterpret_cast<uint4 *>(p.y) + ((size_t)b * (p.Cout / 8) + (n0 + c0) / 8) * HW + pix;
          y[0] = make_uint4(cv_pack_bf2(v[0], v[1]), cv_pack_bf2(v[2], v[3]), cv_pack_bf2(v[4], v[5]), cv_pack_bf2(v[6], v[7]));
          y[HW] = make_uint4(cv_pack_bf2(v[8], v[9]), cv_pack_bf2(v[10], v[11]), cv_pack_bf2(v[12], v[13]),
                             cv_pack_bf2(v[14], v[15]));
        }
      }
      if (!arrived) {
        ptx::tc_fence_before();
        ptx::mbar_arrive(BAR(10 + acc));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// weight [Cout][Cin][3][3] fp32 -> [n_tile][tap][Cin/8][NT][8] bf16 (rows beyond Cout are zero)
__global__ void conv3x3_pack_weight_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ out, int Cout, int Cin,
                                           int NT, int n_tiles, int streamed, int taps) {
  const size_t total = (size_t)n_tiles * taps * Cin * NT;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int j = e % 8;
    const int n = (e / 8) % NT;
    int kc, tap;
    if (streamed) {   // [n_tile][kb (64 ch)][tap][8 chunks][NT][8]
      const int c8 = (e / 8 / NT) % 8;
      tap = (e / 8 / NT / 8) % taps;
      kc = (int)((e / 8 / NT / 8 / taps) % (Cin / 64)) * 8 + c8;
    } else {          // [n_tile][tap][Cin/8][NT][8]
      kc = (e / 8 / NT) % (Cin / 8);
      tap = (e / 8 / NT / (Cin / 8)) % taps;
    }
    const int nt = e / 8 / NT / (Cin / 8) / taps;
    const int co = nt * NT + n, ci = kc * 8 + j;
    out[e] = __float2bfloat16_rn(co < Cout ? w[((size_t)co * Cin + ci) * taps + tap] : 0.f);
  }
}

int conv3x3_pack_weight_raw(const float *w, void *out, int Cout, int Cin, int NT, int n_tiles, int streamed, int taps, cudaStream_t s) {
  conv3x3_pack_weight_kernel<<<kNumSMs * 2, 256, 0, s>>>(w, (__nv_bfloat16 *)out, Cout, Cin, NT, n_tiles, streamed, taps);
  return check_launch("conv3x3_pack_weight");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// N-tile size used for a given Cout / Cin: the largest tile that divides Cout and whose weights (9 * Cin * NT bf16)
// stay resident in shared memory next to the A stages.
static int conv3x3_ntile_resident(int Cout, int Cin, int taps = 9) {
  const int budget = (taps == 9 ? 168 : 128) * 1024;     // 1x1: four 16 KB stages instead of two 23 KB ones
  const int cands[5] = {144, 128, 64, 32, 16};
  for (int c = taps == 9 ? 0 : 1; c < 5; ++c) {
    const int nt = cands[c];
    if (taps * Cin * nt * 2 <= budget && Cout % nt == 0) return nt;
  }
  return 0;
}
// Wide inputs (Cin >= 256: the trunk's 256 -> 64, tsa_fusion's 448 -> 64) would leave only a 16/32-channel resident N tile,
// where every A tile is re-read per N tile and the MMA is bound by the shared-memory reads of A; stream the weights instead.
bool conv3x3_streams(int Cout, int Cin, int taps = 9) {
  return taps == 9 && conv3x3_ntile_resident(Cout, Cin) < 64 && Cout % 64 == 0 && Cin >= 256;
}
int conv3x3_ntile(int Cout, int Cin, int taps = 9) { return conv3x3_streams(Cout, Cin, taps) ? 64 : conv3x3_ntile_resident(Cout, Cin, taps); }

template <int NT, int kStages, int KBLK, int KS, bool DUAL = false>
static int launch_conv3x3(const CUtensorMap &tm, const Conv3x3Params &p, int grid, cudaStream_t s) {
  auto kern = conv3x3_sm100_kernel<NT, kStages, KBLK, KS, DUAL>;
  constexpr int kTaps = KS * KS;
  const int w_bytes = kTaps * p.Cin * NT * 2;
  constexpr int kABytes = (KBLK / 8) * CvGeom<KS>::kPlane;
  const size_t smem = (p.stream_w ? (size_t)kStages * ((kABytes + kTaps * NT * KBLK * 2 + 1023) & ~1023)
                                  : (size_t)((w_bytes + 1023) & ~1023) + kStages * kABytes) + 18 * 8 + NT * 4 + 64;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(CDFO_ERR_CUDA, "cudaFuncSetAttribute(conv3x3_sm100<%d>, %zu): %s", NT, smem, cudaGetErrorString(e));
    attr_smem = smem;
  }
  kern<<<grid, kCvThreads, smem, s>>>(tm, p);
  return check_launch("cdfo_conv3x3_sm100_fwd");
}

}  // namespace cdfo

using namespace cdfo;

extern "C" int cdfo_conv3x3_sm100_ntile(int Cout, int Cin) { return conv3x3_ntile(Cout, Cin); }

extern "C" size_t cdfo_conv_sm100_weight_bytes(int Cout, int Cin, int ksize) {
  if (ksize != 1 && ksize != 3) return 0;
  const int nt = conv3x3_ntile(Cout, Cin, ksize * ksize);
  if (!nt) return 0;
  return (size_t)ceil_div(Cout, nt) * ksize * ksize * Cin * nt * 2;
}
extern "C" size_t cdfo_conv3x3_sm100_weight_bytes(int Cout, int Cin) { return cdfo_conv_sm100_weight_bytes(Cout, Cin, 3); }

extern "C" int cdfo_conv_sm100_pack_weight(const float *w, void *wpk, int Cout, int Cin, int ksize, void *stream) {
  CDFO_REQUIRE(w && wpk, CDFO_ERR_NULL, "cdfo_conv_sm100_pack_weight: NULL pointer");
  CDFO_REQUIRE(ksize == 1 || ksize == 3, CDFO_ERR_UNSUPPORTED, "cdfo_conv_sm100: kernel size 1 or 3 (got %d)", ksize);
  const int taps = ksize * ksize, nt = conv3x3_ntile(Cout, Cin, taps);
  CDFO_REQUIRE(nt && Cin % 64 == 0, CDFO_ERR_UNSUPPORTED, "cdfo_conv_sm100: unsupported channels %d -> %d", Cin, Cout);
  const int n_tiles = ceil_div(Cout, nt);
  conv3x3_pack_weight_kernel<<<kNumSMs * 2, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16 *)wpk, Cout, Cin, nt, n_tiles,
                                                                             conv3x3_streams(Cout, Cin, taps) ? 1 : 0, taps);
  return check_launch("cdfo_conv_sm100_pack_weight");
}
extern "C" int cdfo_conv3x3_sm100_pack_weight(const float *w, void *wpk, int Cout, int Cin, void *stream) {
  return cdfo_conv_sm100_pack_weight(w, wpk, Cout, Cin, 3, stream);
}

static int conv3x3_run(const void *x_c8, const void *wpk, const float *bias, const void *resid_c8, void *y, int B, int Cin,
                       int Cout, int H, int W, int act, int out_mode, int epi, float mag, const void *aux, void *stream, int ksize = 3);

extern "C" int cdfo_conv3x3_sm100_fwd(const void *x_c8, const void *wpk, const float *bias, const void *resid_c8, void *y,
                                      int B, int Cin, int Cout, int H, int W, int act, int out_mode, void *stream) {
  return conv3x3_run(x_c8, wpk, bias, resid_c8, y, B, Cin, Cout, H, W, act, out_mode, 0, 0.f, nullptr, stream);
}

extern "C" int cdfo_conv_sm100_fwd(const void *x_c8, const void *wpk, const float *bias, const void *resid_c8, void *y, int B,
                                   int Cin, int Cout, int H, int W, int ksize, int act, int out_mode, void *stream) {
  CDFO_REQUIRE(ksize == 1 || ksize == 3, CDFO_ERR_UNSUPPORTED, "cdfo_conv_sm100_fwd: kernel size 1 or 3 (got %d)", ksize);
  return conv3x3_run(x_c8, wpk, bias, resid_c8, y, B, Cin, Cout, H, W, act, out_mode, 0, 0.f, nullptr, stream, ksize);
}

extern "C" int cdfo_mv_offset_head_sm100_fwd(const void *z_c8, const void *wpk, const float *bias, const void *first,
                                             void *out, int B, int Cin, int dg, int H, int W, float magnitude,
                                             void *stream) {
  CDFO_REQUIRE(dg > 0 && (dg * 27) % 144 == 0, CDFO_ERR_UNSUPPORTED,
               "cdfo_mv_offset_head_sm100_fwd: deformable_groups * 27 must be a multiple of 144 (got dg = %d)", dg);
  CDFO_REQUIRE(conv3x3_ntile(dg * 27, Cin) == 144, CDFO_ERR_UNSUPPORTED,
               "cdfo_mv_offset_head_sm100_fwd: needs the 144-channel N tile (Cin = %d too large)", Cin);
  CDFO_REQUIRE(((uintptr_t)out & 15) == 0 && ((uintptr_t)first & 15) == 0, CDFO_ERR_SHAPE, "cdfo_mv_offset_head_sm100_fwd: fields must be 16-byte aligned");
  return conv3x3_run(z_c8, wpk, bias, nullptr, out, B, Cin, dg * 27, H, W, 0, 0, first ? 2 : 1, magnitude, first, stream);
}

extern "C" int cdfo_mv_offset_head_dual_sm100_fwd(const void *z_c8, const void *wpk, const float *bias, void *out, int B, int Cin, int dg,
                                                  int H, int W, float magnitude, void *stream) {
  CDFO_REQUIRE(dg > 0 && (dg * 27) % 144 == 0, CDFO_ERR_UNSUPPORTED,
               "cdfo_mv_offset_head_dual_sm100_fwd: deformable_groups * 27 must be a multiple of 144 (got dg = %d)", dg);
  CDFO_REQUIRE(conv3x3_ntile(dg * 27, Cin) == 144, CDFO_ERR_UNSUPPORTED,
               "cdfo_mv_offset_head_dual_sm100_fwd: needs the 144-channel N tile (Cin = %d too large)", Cin);
  CDFO_REQUIRE(((uintptr_t)out & 15) == 0, CDFO_ERR_SHAPE, "cdfo_mv_offset_head_dual_sm100_fwd: fields must be 16-byte aligned");
  return conv3x3_run(z_c8, wpk, bias, nullptr, out, B, Cin, dg * 27, H, W, 0, 0, 4, magnitude, nullptr, stream);
}

extern "C" int cdfo_conv_last_skip_sm100_fwd(const void *x_c8, const void *wpk, const float *bias, const float *lr, float *y, int B,
                                             int Cin, int H, int W, void *stream) {
  CDFO_REQUIRE(lr, CDFO_ERR_NULL, "cdfo_conv_last_skip_sm100_fwd: NULL pointer");
  return conv3x3_run(x_c8, wpk, bias, nullptr, y, B, Cin, 16, H, W, 0, 0, 3, 0.f, lr, stream);
}

static int conv3x3_run(const void *x_c8, const void *wpk, const float *bias, const void *resid_c8, void *y, int B, int Cin,
                       int Cout, int H, int W, int act, int out_mode, int epi, float mag, const void *aux, void *stream, int ksize) {
  CDFO_REQUIRE(x_c8 && wpk && y, CDFO_ERR_NULL, "cdfo_conv3x3_sm100_fwd: NULL pointer");
  CDFO_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, CDFO_ERR_SHAPE, "cdfo_conv3x3_sm100_fwd: bad shape");
  CDFO_REQUIRE(Cin % 64 == 0 && Cout % 16 == 0, CDFO_ERR_UNSUPPORTED,
               "cdfo_conv3x3_sm100_fwd: Cin must be a multiple of 64 and Cout of 16 (got %d -> %d)", Cin, Cout);
  CDFO_REQUIRE(act >= 0 && act <= 2 && out_mode >= 0 && out_mode <= 2, CDFO_ERR_UNSUPPORTED, "cdfo_conv3x3_sm100_fwd: act/out_mode");
  CDFO_REQUIRE(out_mode != 2 || (Cout % 32 == 0 && resid_c8 == nullptr), CDFO_ERR_UNSUPPORTED,
               "cdfo_conv3x3_sm100_fwd: pixel-shuffle output needs Cout %% 32 == 0 and no residual");
  CDFO_REQUIRE(epi != 3 || (H % 4 == 0 && W % 4 == 0 && aux), CDFO_ERR_SHAPE, "cdfo_conv_last_skip_sm100_fwd: H, W must be multiples of 4");
  CDFO_REQUIRE(((uintptr_t)x_c8 & 15) == 0 && ((uintptr_t)wpk & 15) == 0 && ((uintptr_t)y & (epi ? 7 : 15)) == 0, CDFO_ERR_SHAPE,
               "cdfo_conv3x3_sm100_fwd: pointers must be 16-byte aligned");
  const int taps = ksize * ksize;
  const int nt = conv3x3_ntile(Cout, Cin, taps);
  CDFO_REQUIRE(nt, CDFO_ERR_UNSUPPORTED, "cdfo_conv3x3_sm100_fwd: no N tile fits shared memory for %d -> %d", Cin, Cout);
  EncodeTiledFn enc = encode_tiled_fn();
  CDFO_REQUIRE(enc, CDFO_ERR_CUDA, "cdfo_conv3x3_sm100_fwd: cuTensorMapEncodeTiled not available from the driver");
  CUtensorMap tm;
  static const bool wide = getenv("CDFO_TMA_WIDE1") == nullptr || getenv("CDFO_TMA_WIDE1")[0] != '0';
  const int Bx = epi == 4 ? 2 * B : B;          // dual head: the input holds both hidden maps, [2 B] samples
  cuuint64_t gdim[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(Cin / 8), (cuuint64_t)Bx};
  cuuint64_t gstr[4] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)(Cin / 8) * H * W * 16};
  cuuint32_t box[5] = {8, (cuuint32_t)(kCvTileW + ksize - 1), (cuuint32_t)(kCvTileH + ksize - 1), 8, 1};
  if (wide) {   // same bytes, same landing order; TMA issues one request per halo row (up to 160 bytes) instead of one per 16-byte chunk
    gdim[0] = (cuuint64_t)W * 8; gdim[1] = H; gdim[2] = Cin / 8; gdim[3] = Bx; gdim[4] = 1;
    gstr[0] = (cuuint64_t)W * 16; gstr[1] = (cuuint64_t)H * W * 16; gstr[2] = (cuuint64_t)(Cin / 8) * H * W * 16; gstr[3] = gstr[2] * Bx;
    box[0] = 8 * (kCvTileW + ksize - 1); box[1] = kCvTileH + ksize - 1; box[2] = 8; box[3] = 1; box[4] = 1;
  }
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(x_c8), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CDFO_REQUIRE(cr == CUDA_SUCCESS, CDFO_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
  Conv3x3Params p;
  p.wpk = (const uint8_t *)wpk; p.bias = bias; p.resid = (const uint4 *)resid_c8; p.y = y;
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.H = H; p.W = W; p.act = act; p.out_mode = out_mode;
  p.epi = epi; p.mag = mag; p.aux = (const uint2 *)aux; p.tma_wide = wide ? 1 : 0;
  p.n_tiles = ceil_div(Cout, nt);
  p.stream_w = conv3x3_streams(Cout, Cin, taps) ? 1 : 0;
  p.tiles_x = ceil_div(W, kCvTileW); p.tiles_y = ceil_div(H, kCvTileH);
  const long long mt = (long long)B * p.tiles_x * p.tiles_y;
  CDFO_REQUIRE(mt < (1ll << 31), CDFO_ERR_UNSUPPORTED, "cdfo_conv3x3_sm100_fwd: too many tiles");
  p.m_tiles = (int)mt;
  int groups = kNumSMs / p.n_tiles;             // CTAs per N tile
  if (groups > p.m_tiles) groups = p.m_tiles;
  const int grid = groups * p.n_tiles;
  cudaStream_t s = (cudaStream_t)stream;
  if (ksize == 1) {   // 1x1: the A operand is the tile itself, four 16 KB stages
    switch (nt) {
      case 16: return launch_conv3x3<16, 4, 64, 1>(tm, p, grid, s);
      case 32: return launch_conv3x3<32, 4, 64, 1>(tm, p, grid, s);
      case 64: return launch_conv3x3<64, 4, 64, 1>(tm, p, grid, s);
      case 128: return launch_conv3x3<128, 4, 64, 1>(tm, p, grid, s);
    }
    return fail(CDFO_ERR_UNSUPPORTED, "cdfo_conv_sm100_fwd: N tile %d", nt);
  }
  if (epi == 4) {
    CDFO_REQUIRE(nt == 144, CDFO_ERR_UNSUPPORTED, "cdfo_mv_offset_head_dual_sm100_fwd: needs the 144-channel N tile");
    return launch_conv3x3<144, 2, 64, 3, true>(tm, p, grid, s);
  }
  // resident weights: with room for four 23 KB halo stages the TMA latency of the next tiles hides behind the current one (a 64 -> 64
  // convolution has ONE K block per tile, so two stages meant one tile of look-ahead)
  const bool deep = !p.stream_w && (size_t)taps * Cin * nt * 2 + 4 * (size_t)(64 / 8) * CvGeom<3>::kPlane + 2048 <= 227 * 1024;
  if (deep) {
    switch (nt) {
      case 16: return launch_conv3x3<16, 4, 64, 3>(tm, p, grid, s);
      case 32: return launch_conv3x3<32, 4, 64, 3>(tm, p, grid, s);
      case 64: return launch_conv3x3<64, 4, 64, 3>(tm, p, grid, s);
    }
  }
  switch (nt) {
    case 16: return launch_conv3x3<16, 2, 64, 3>(tm, p, grid, s);
    case 32: return launch_conv3x3<32, 2, 64, 3>(tm, p, grid, s);
    case 64: return launch_conv3x3<64, 2, 64, 3>(tm, p, grid, s);
    case 128: return launch_conv3x3<128, 2, 64, 3>(tm, p, grid, s);
    case 144: return launch_conv3x3<144, 2, 64, 3>(tm, p, grid, s);
  }
  return fail(CDFO_ERR_UNSUPPORTED, "cdfo_conv3x3_sm100_fwd: N tile %d", nt);
}
