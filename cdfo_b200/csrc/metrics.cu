// Frame I/O conversions and the on-GPU PSNR / SSIM of the evaluation loop.
//
//   planes -> unit fp32      test_LD_37.py:19-29  generate_input: img.astype(float32) / 255.0, 270-row frames get two zero rows
//   SR fp32 -> uint8 frame   test_LD_37.py:172-180 crop of the padded rows, clamp(0, 1) * 255.0, astype(uint8) (truncation)
//   PSNR / SSIM              metric/psnr_ssim.py:278-317 calculate_psnr, :320-350 _ssim, :353-399 calculate_ssim,
//                            :446-484 cal_psnr_ssim (crop_border 4, test_y_channel: the float32 / 255 * 255 round trip of :210-214)
//
// All three are HBM-bound (1-5 bytes per HR pixel); the SSIM kernel evaluates the 11x11 Gaussian (sigma 1.5) separably in
// fp64 from a shared-memory tile, never materialising the five filtered maps cv2.filter2D writes in the reference.
// Reductions are two-pass with a fixed summation order, so results are bit-reproducible from run to run.
#include "cdfo_common.cuh"

namespace cdfo {
namespace met {

// ---------------------------------------------------------------- planes -> fp32 / 255
template <typename T>
__global__ void planes_to_unit_kernel(const T *__restrict__ src, float *__restrict__ dst, int H_in, int W, int H_out, size_t total) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(e % W);
    const size_t r = e / W;
    const int y = (int)(r % H_out);
    const size_t n = r / H_out;
    // correctly rounded k / 255 like numpy / ATen's CPU true division; padded rows are zero (test_LD_37.py:24-26)
    dst[e] = y < H_in ? __fdiv_rn((float)src[(n * H_in + y) * W + x], 255.0f) : 0.f;
  }
}

// ---------------------------------------------------------------- SR -> uint8
__global__ void sr_to_u8_kernel(const float *__restrict__ sr, uint8_t *__restrict__ out, int H_in, int W, int H_out, size_t total) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(e % W);
    const size_t r = e / W;
    const int y = (int)(r % H_out);
    const size_t n = r / H_out;
    float v = sr[(n * H_in + y) * W + x];
    v = v != v ? 0.f : fminf(fmaxf(v, 0.f), 1.f);          // clamp(0, 1); NaN -> 0
    out[e] = (uint8_t)(int)__fmul_rn(v, 255.0f);            // numpy astype(uint8): truncation toward zero
  }
}

// vectorised variant: 4 pixels per thread (W % 4 == 0, 16-byte aligned rows)
__global__ void sr_to_u8_vec4_kernel(const float4 *__restrict__ sr, uchar4 *__restrict__ out, int H_in, int W4, int H_out, size_t total) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(e % W4);
    const size_t r = e / W4;
    const int y = (int)(r % H_out);
    const size_t n = r / H_out;
    const float4 v = __ldcs(sr + (n * H_in + y) * W4 + x);
    auto q = [](float a) {
      a = a != a ? 0.f : fminf(fmaxf(a, 0.f), 1.f);
      return (unsigned char)(int)__fmul_rn(a, 255.0f);
    };
    out[e] = make_uchar4(q(v.x), q(v.y), q(v.z), q(v.w));
  }
}

// ---------------------------------------------------------------- PSNR / SSIM
constexpr int kTW = 32, kTH = 16, kR = 5;           // output tile, Gaussian radius
constexpr int kIW = kTW + 2 * kR, kIH = kTH + 2 * kR;

struct GaussTaps { double g[11]; };

// to_y_channel of a single-channel image (metric/psnr_ssim.py:210-214): float32(k) / 255 * 255, in float32
__device__ __forceinline__ float y_round_trip(uint8_t k) { return __fmul_rn(__fdiv_rn((float)k, 255.0f), 255.0f); }

__device__ __forceinline__ double block_sum(double v, double *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < nw; ++w) s += red[w];      // fixed order
  __syncthreads();
  return s;
}

// one block = one 32x16 tile of the SSIM map of one image; partial[b][tile] = sum of the tile's SSIM values
__global__ void __launch_bounds__(kTW * kTH) ssim_tile_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b,
                                                              double *__restrict__ partial, int H, int W, int border, int tiles_x,
                                                              int tiles, GaussTaps taps) {
  __shared__ float ia[kIH][kIW], ib[kIH][kIW];
  __shared__ double hz[5][kIH][kTW];
  __shared__ double red[kTW * kTH / 32];
  const int img = blockIdx.y, tile = blockIdx.x;
  const int ty0 = (tile / tiles_x) * kTH, tx0 = (tile % tiles_x) * kTW;
  const int Hc = H - 2 * border, Wc = W - 2 * border;      // cropped image
  const int Hs = Hc - 2 * kR, Ws = Wc - 2 * kR;            // SSIM map = valid part of the filtered cropped image
  const uint8_t *pa = a + (size_t)img * H * W, *pb = b + (size_t)img * H * W;
  const int tid = threadIdx.x;
  for (int e = tid; e < kIH * kIW; e += kTW * kTH) {
    const int r = e / kIW, c = e % kIW;
    const int y = ty0 + r, x = tx0 + c;                     // cropped-image coordinates
    float va = 0.f, vb = 0.f;
    if (y < Hc && x < Wc) {
      va = y_round_trip(pa[(size_t)(y + border) * W + x + border]);
      vb = y_round_trip(pb[(size_t)(y + border) * W + x + border]);
    }
    ia[r][c] = va;
    ib[r][c] = vb;
  }
  __syncthreads();
  const int tx = tid % kTW, ty = tid / kTW;
  for (int r = ty; r < kIH; r += kTH) {
    double s1 = 0, s2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const double p = (double)ia[r][tx + k], q = (double)ib[r][tx + k], g = taps.g[k];
      s1 += g * p; s2 += g * q; s11 += g * (p * p); s22 += g * (q * q); s12 += g * (p * q);
    }
    hz[0][r][tx] = s1; hz[1][r][tx] = s2; hz[2][r][tx] = s11; hz[3][r][tx] = s22; hz[4][r][tx] = s12;
  }
  __syncthreads();
  double val = 0.0;
  if (ty0 + ty < Hs && tx0 + tx < Ws) {
    double m1 = 0, m2 = 0, e11 = 0, e22 = 0, e12 = 0;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const double g = taps.g[k];
      m1 += g * hz[0][ty + k][tx]; m2 += g * hz[1][ty + k][tx];
      e11 += g * hz[2][ty + k][tx]; e22 += g * hz[3][ty + k][tx]; e12 += g * hz[4][ty + k][tx];
    }
    const double C1 = (0.01 * 255) * (0.01 * 255), C2 = (0.03 * 255) * (0.03 * 255);
    const double m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
    const double v1 = e11 - m11, v2 = e22 - m22, cv = e12 - m12;
    val = ((2 * m12 + C1) * (2 * cv + C2)) / ((m11 + m22 + C1) * (v1 + v2 + C2));
  }
  const double s = block_sum(val, red);
  if (tid == 0) partial[(size_t)img * tiles + tile] = s;
}

// squared error over the cropped image; partial[b][chunk]
__global__ void sse_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, double *__restrict__ partial, int H, int W,
                           int border, int chunks) {
  __shared__ double red[8];
  const int img = blockIdx.y, chunk = blockIdx.x;
  const int Hc = H - 2 * border, Wc = W - 2 * border;
  const size_t n = (size_t)Hc * Wc;
  const uint8_t *pa = a + (size_t)img * H * W, *pb = b + (size_t)img * H * W;
  double s = 0.0;
  for (size_t e = (size_t)chunk * blockDim.x + threadIdx.x; e < n; e += (size_t)chunks * blockDim.x) {
    const int y = (int)(e / Wc) + border, x = (int)(e % Wc) + border;
    const float d = __fsub_rn(y_round_trip(pa[(size_t)y * W + x]), y_round_trip(pb[(size_t)y * W + x]));
    s += (double)__fmul_rn(d, d);                           // (img1 - img2) ** 2 is a float32 array in the reference
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[(size_t)img * chunks + chunk] = s;
}

// one block per image: fixed-order sums of the partials -> (psnr, ssim); optional accumulation for the sequence average
__global__ void finalize_kernel(const double *__restrict__ ssim_part, const double *__restrict__ sse_part, int tiles, int chunks,
                                double n_sse, double n_ssim, double *__restrict__ frame_out, double *__restrict__ accum) {
  __shared__ double red[8];
  const int img = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < tiles; i += blockDim.x) s += ssim_part[(size_t)img * tiles + i];
  const double ssim_sum = block_sum(s, red);
  s = 0.0;
  for (int i = threadIdx.x; i < chunks; i += blockDim.x) s += sse_part[(size_t)img * chunks + i];
  const double sse = block_sum(s, red);
  if (threadIdx.x == 0) {
    const double mse = sse / n_sse;
    const double psnr = mse == 0.0 ? __longlong_as_double(0x7ff0000000000000ll) : 20.0 * log10(255.0 / sqrt(mse));
    const double ssim = ssim_sum / n_ssim;
    if (frame_out) { frame_out[img * 2 + 0] = psnr; frame_out[img * 2 + 1] = ssim; }
    if (accum) { accum[img * 3 + 0] += psnr; accum[img * 3 + 1] += ssim; accum[img * 3 + 2] += 1.0; }
  }
}

static int grid_for(size_t total, int threads) {
  const size_t blocks = (total + threads - 1) / threads;
  const size_t cap = (size_t)kNumSMs * 16;
  return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}
constexpr int kSseChunks = kNumSMs * 2;

}  // namespace met
}  // namespace cdfo

using namespace cdfo;

extern "C" int cdfo_planes_to_unit_f32(const void *src, int src_kind, float *dst, int n_planes, int H_in, int W, int H_out, void *stream) {
  CDFO_REQUIRE(src && dst, CDFO_ERR_NULL, "cdfo_planes_to_unit_f32: NULL pointer");
  CDFO_REQUIRE(n_planes > 0 && H_in > 0 && W > 0 && H_out >= H_in, CDFO_ERR_SHAPE, "cdfo_planes_to_unit_f32: bad shape (H_out must be >= H_in)");
  const size_t total = (size_t)n_planes * H_out * W;
  const int grid = met::grid_for(total, 256);
  cudaStream_t s = (cudaStream_t)stream;
  switch (src_kind) {
    case 0: met::planes_to_unit_kernel<uint8_t><<<grid, 256, 0, s>>>((const uint8_t *)src, dst, H_in, W, H_out, total); break;
    case 1: met::planes_to_unit_kernel<int8_t><<<grid, 256, 0, s>>>((const int8_t *)src, dst, H_in, W, H_out, total); break;
    case 2: met::planes_to_unit_kernel<int16_t><<<grid, 256, 0, s>>>((const int16_t *)src, dst, H_in, W, H_out, total); break;
    case 3: met::planes_to_unit_kernel<int32_t><<<grid, 256, 0, s>>>((const int32_t *)src, dst, H_in, W, H_out, total); break;
    default: return fail(CDFO_ERR_UNSUPPORTED, "cdfo_planes_to_unit_f32: source kind %d (0 u8, 1 i8, 2 i16, 3 i32)", src_kind);
  }
  return check_launch("cdfo_planes_to_unit_f32");
}

extern "C" int cdfo_sr_to_u8(const float *sr, uint8_t *out, int n_planes, int H_in, int W, int H_out, void *stream) {
  CDFO_REQUIRE(sr && out, CDFO_ERR_NULL, "cdfo_sr_to_u8: NULL pointer");
  CDFO_REQUIRE(n_planes > 0 && H_out > 0 && W > 0 && H_out <= H_in, CDFO_ERR_SHAPE, "cdfo_sr_to_u8: bad shape (H_out must be <= H_in)");
  cudaStream_t s = (cudaStream_t)stream;
  if (W % 4 == 0 && ((uintptr_t)sr & 15) == 0 && ((uintptr_t)out & 3) == 0) {
    const size_t total = (size_t)n_planes * H_out * (W / 4);
    met::sr_to_u8_vec4_kernel<<<met::grid_for(total, 256), 256, 0, s>>>((const float4 *)sr, (uchar4 *)out, H_in, W / 4, H_out, total);
  } else {
    const size_t total = (size_t)n_planes * H_out * W;
    met::sr_to_u8_kernel<<<met::grid_for(total, 256), 256, 0, s>>>(sr, out, H_in, W, H_out, total);
  }
  return check_launch("cdfo_sr_to_u8");
}

static int ssim_tiles(int H, int W, int border, int *tiles_x) {
  const int Hs = H - 2 * border - 2 * met::kR, Ws = W - 2 * border - 2 * met::kR;
  if (Hs <= 0 || Ws <= 0) return 0;
  *tiles_x = ceil_div(Ws, met::kTW);
  return *tiles_x * ceil_div(Hs, met::kTH);
}

extern "C" size_t cdfo_psnr_ssim_workspace_bytes(int B, int H, int W, int border) {
  int tx = 0;
  const int tiles = ssim_tiles(H, W, border, &tx);
  if (B <= 0 || tiles <= 0) return 0;
  return (size_t)B * (tiles + met::kSseChunks) * sizeof(double);
}

extern "C" int cdfo_psnr_ssim_u8(const uint8_t *res, const uint8_t *gt, int B, int H, int W, int border, double *frame_out,
                                 double *accum, void *workspace, void *stream) {
  CDFO_REQUIRE(res && gt && workspace, CDFO_ERR_NULL, "cdfo_psnr_ssim_u8: NULL pointer");
  CDFO_REQUIRE(frame_out || accum, CDFO_ERR_NULL, "cdfo_psnr_ssim_u8: neither frame_out nor accum given");
  CDFO_REQUIRE(B > 0 && B <= 65535 && border >= 0, CDFO_ERR_SHAPE, "cdfo_psnr_ssim_u8: bad batch / border");
  int tiles_x = 0;
  const int tiles = ssim_tiles(H, W, border, &tiles_x);
  CDFO_REQUIRE(tiles > 0, CDFO_ERR_SHAPE, "cdfo_psnr_ssim_u8: %dx%d image with border %d is smaller than the 11x11 SSIM window", H, W, border);
  // cv2.getGaussianKernel(11, 1.5): exp(-(i - 5)^2 / (2 sigma^2)) normalised to sum 1, in double
  met::GaussTaps taps;
  double sum = 0.0;
  for (int i = 0; i < 11; ++i) {
    taps.g[i] = exp(-((double)(i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5));
    sum += taps.g[i];
  }
  for (int i = 0; i < 11; ++i) taps.g[i] /= sum;
  double *ssim_part = (double *)workspace, *sse_part = ssim_part + (size_t)B * tiles;
  cudaStream_t s = (cudaStream_t)stream;
  met::ssim_tile_kernel<<<dim3(tiles, B), met::kTW * met::kTH, 0, s>>>(res, gt, ssim_part, H, W, border, tiles_x, tiles, taps);
  met::sse_kernel<<<dim3(met::kSseChunks, B), 256, 0, s>>>(res, gt, sse_part, H, W, border, met::kSseChunks);
  const int Hc = H - 2 * border, Wc = W - 2 * border;
  met::finalize_kernel<<<B, 256, 0, s>>>(ssim_part, sse_part, tiles, met::kSseChunks, (double)Hc * Wc,
                                         (double)(Hc - 2 * met::kR) * (Wc - 2 * met::kR), frame_out, accum);
  return check_launch("cdfo_psnr_ssim_u8");
}
