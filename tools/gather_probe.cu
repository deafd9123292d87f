// Gather-path probe for the DCN producer (design experiment, not product code):
// how fast can one SM fetch the 2x2 bilinear footprint of 4 fp16 channels per (pixel, group, tap) sample?
//   mode 0: 4 x LDG.64 from a quad-planar plane (8-byte texels)                      -- what dcn_sm100 v2 does
//   mode 1: 2 x LDG.128 from a pair-duplicated plane (16-byte texel = texels w, w+1)
//   mode 2: 1 x tex2DLayered<float4>, hardware bilinear on half4 texels, border mode
//   mode 3: 4 x tex2DLayered<uint2>-style point fetches (raw 8-byte texels through the TEX pipe)
//   mode 4: 2 x LDG.128 pair-duplicated + fp16x2 blend (HFMA2), the candidate v3 inner loop
// Sample positions: block-constant MV (8x8 blocks, +-1.5 px) + N(0, jitter) per sample, like tools/bench_dcn.py.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe tools/gather_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int H = 272, W = 480, Q = 16, TAPS = 9;

struct Params {
  const uint2 *xq;       // [B][Q][H+3][W+3] 8-byte texels, zero border 1 before / 2 after
  const uint4 *xp;       // [B][Q][H+3][W+3] 16-byte pair texels (w, w+1)
  cudaTextureObject_t tex_lin, tex_pt, tex_p2d;  // p2d: pitch-linear 2D over xq, planes folded into rows
  const __half *fields;  // [B][Q*9][H*W][4] (dy, dx, m, 0)
  float *out;            // [B][H*W] checksum per pixel
  int B;
};

__device__ __forceinline__ float bf(uint32_t u, int hi) {
  __half2 h = *reinterpret_cast<__half2 *>(&u);
  return hi ? __high2float(h) : __low2float(h);
}

template <int MODE>
__global__ void __launch_bounds__(512, 2) probe(Params p) {
  // persistent: CTA walks 4x32 tiles; thread = (pixel row in tile, 4 quads)
  const int tiles_x = W / 32, tiles_y = H / 4, per_img = tiles_x * tiles_y;
  const int ntiles = per_img * p.B;
  const int tid = threadIdx.x, row = tid & 127, ty = row >> 5, tx = row & 31, quad0 = (tid >> 7) * 4;
  const int Wp = W + 3, plane = (H + 3) * Wp, P = H * W;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / per_img, r = tile % per_img, h = (r / tiles_x) * 4 + ty, w = (r % tiles_x) * 32 + tx;
    const int pix = h * W + w;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    __half2 hacc0 = __float2half2_rn(0.f), hacc1 = hacc0;
    uint2 nxt[4];
#pragma unroll
    for (int qi = 0; qi < 4; ++qi)
      nxt[qi] = __ldcs(reinterpret_cast<const uint2 *>(p.fields) + ((size_t)b * Q * 9 + (quad0 + qi) * 9) * P + pix);
#pragma unroll 1
    for (int tap = 0; tap < TAPS; ++tap) {
      uint2 cur[4];
#pragma unroll
      for (int qi = 0; qi < 4; ++qi) {
        cur[qi] = nxt[qi];
        if (tap < 8) nxt[qi] = __ldcs(reinterpret_cast<const uint2 *>(p.fields) + ((size_t)b * Q * 9 + (quad0 + qi) * 9 + tap + 1) * P + pix);
      }
#pragma unroll
      for (int qi = 0; qi < 4; ++qi) {
        const int q = quad0 + qi;
        const uint2 raw = cur[qi];
        const float2 d = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
        const float m = __low2float(*reinterpret_cast<const __half2 *>(&raw.y));
        const float h_im = (float)(h - 1 + tap / 3) + d.x, w_im = (float)(w - 1 + tap % 3) + d.y;
        if (MODE == 5 || MODE == 6 || MODE == 7) {
          // mixed: quads with qi < NT go through the texture unit (pitch-2D, planes folded into the row axis),
          // the others through the LSU (pair texels + HFMA2)
          constexpr int NT = MODE == 5 ? 2 : (MODE == 6 ? 3 : 4);
          if (qi < NT) {
            const float hcl = fminf(fmaxf(h_im, -1.f), (float)H);
            const float4 v = tex2D<float4>(p.tex_p2d, w_im + 1.5f, hcl + 1.5f + (float)((b * Q + q) * (H + 3)));
            const __half2 m2 = __float2half2_rn(m);
            hacc0 = __hfma2(m2, __floats2half2_rn(v.x, v.y), hacc0);
            hacc1 = __hfma2(m2, __floats2half2_rn(v.z, v.w), hacc1);
            continue;
          }
        }
        if (MODE == 2) {
          const float4 v = tex2DLayered<float4>(p.tex_lin, w_im + 1.5f, h_im + 1.5f, b * Q + q);
          acc0 = fmaf(m, v.x, acc0); acc1 = fmaf(m, v.y, acc1); acc2 = fmaf(m, v.z, acc2); acc3 = fmaf(m, v.w, acc3);
          continue;
        }
        const float hc = fminf(fmaxf(h_im, -1.f), (float)H), wc = fminf(fmaxf(w_im, -1.f), (float)W);
        // floor without the XU pipe: round-down add of 1.5 * 2^23
        const float th = __fadd_rd(hc, 12582912.f), tw = __fadd_rd(wc, 12582912.f);
        const float hf = th - 12582912.f, wf = tw - 12582912.f;
        const float lh = hc - hf, lw = wc - wf;
        const float a = m - m * lh, bb = m * lh;
        const float w0 = a - a * lw, w1 = a * lw, w2 = bb - bb * lw, w3 = bb * lw;
        const int hi = __float_as_int(th) - 0x4B400000, wi = __float_as_int(tw) - 0x4B400000;
        const int idx = (hi + 1) * Wp + wi + 1;
        if (MODE == 0) {
          const uint2 *qp = p.xq + ((size_t)b * Q + q) * plane + idx;
          const uint2 v0 = __ldg(qp), v1 = __ldg(qp + 1), v2 = __ldg(qp + Wp), v3 = __ldg(qp + Wp + 1);
          acc0 += w0 * bf(v0.x, 0) + w1 * bf(v1.x, 0) + w2 * bf(v2.x, 0) + w3 * bf(v3.x, 0);
          acc1 += w0 * bf(v0.x, 1) + w1 * bf(v1.x, 1) + w2 * bf(v2.x, 1) + w3 * bf(v3.x, 1);
          acc2 += w0 * bf(v0.y, 0) + w1 * bf(v1.y, 0) + w2 * bf(v2.y, 0) + w3 * bf(v3.y, 0);
          acc3 += w0 * bf(v0.y, 1) + w1 * bf(v1.y, 1) + w2 * bf(v2.y, 1) + w3 * bf(v3.y, 1);
        } else if (MODE == 1) {
          const uint4 *qp = p.xp + ((size_t)b * Q + q) * plane + idx;
          const uint4 v0 = __ldg(qp), v2 = __ldg(qp + Wp);
          acc0 += w0 * bf(v0.x, 0) + w1 * bf(v0.z, 0) + w2 * bf(v2.x, 0) + w3 * bf(v2.z, 0);
          acc1 += w0 * bf(v0.x, 1) + w1 * bf(v0.z, 1) + w2 * bf(v2.x, 1) + w3 * bf(v2.z, 1);
          acc2 += w0 * bf(v0.y, 0) + w1 * bf(v0.w, 0) + w2 * bf(v2.y, 0) + w3 * bf(v2.w, 0);
          acc3 += w0 * bf(v0.y, 1) + w1 * bf(v0.w, 1) + w2 * bf(v2.y, 1) + w3 * bf(v2.w, 1);
        } else if (MODE >= 4) {
          const uint4 *qp = p.xp + ((size_t)b * Q + q) * plane + idx;
          const uint4 v0 = __ldg(qp), v2 = __ldg(qp + Wp);
          const __half2 h0 = __float2half2_rn(w0), h1 = __float2half2_rn(w1), h2 = __float2half2_rn(w2), h3 = __float2half2_rn(w3);
          hacc0 = __hfma2(h0, *reinterpret_cast<const __half2 *>(&v0.x), hacc0);
          hacc1 = __hfma2(h0, *reinterpret_cast<const __half2 *>(&v0.y), hacc1);
          hacc0 = __hfma2(h1, *reinterpret_cast<const __half2 *>(&v0.z), hacc0);
          hacc1 = __hfma2(h1, *reinterpret_cast<const __half2 *>(&v0.w), hacc1);
          hacc0 = __hfma2(h2, *reinterpret_cast<const __half2 *>(&v2.x), hacc0);
          hacc1 = __hfma2(h2, *reinterpret_cast<const __half2 *>(&v2.y), hacc1);
          hacc0 = __hfma2(h3, *reinterpret_cast<const __half2 *>(&v2.z), hacc0);
          hacc1 = __hfma2(h3, *reinterpret_cast<const __half2 *>(&v2.w), hacc1);
        } else if (MODE == 3) {
          const int layer = b * Q + q;
          const float fx = (float)(wi + 1) + 0.5f, fy = (float)(hi + 1) + 0.5f;
          const float4 v0 = tex2DLayered<float4>(p.tex_pt, fx, fy, layer), v1 = tex2DLayered<float4>(p.tex_pt, fx + 1.f, fy, layer);
          const float4 v2 = tex2DLayered<float4>(p.tex_pt, fx, fy + 1.f, layer), v3 = tex2DLayered<float4>(p.tex_pt, fx + 1.f, fy + 1.f, layer);
          acc0 += w0 * v0.x + w1 * v1.x + w2 * v2.x + w3 * v3.x;
          acc1 += w0 * v0.y + w1 * v1.y + w2 * v2.y + w3 * v3.y;
          acc2 += w0 * v0.z + w1 * v1.z + w2 * v2.z + w3 * v3.z;
          acc3 += w0 * v0.w + w1 * v1.w + w2 * v2.w + w3 * v3.w;
        }
      }
    }
    const float2 f0 = __half22float2(hacc0), f1 = __half22float2(hacc1);
    atomicAdd(p.out + (size_t)b * P + pix, acc0 + acc1 + acc2 + acc3 + f0.x + f0.y + f1.x + f1.y);
  }
}

int main(int argc, char **argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 6;
  const float jitter = argc > 2 ? atof(argv[2]) : 0.3f;
  const int grid = 148 * (argc > 3 ? atoi(argv[3]) : 2);
  const int Hp = H + 3, Wp = W + 3, P = H * W;
  srand(1);
  auto rnd = []() { return (rand() & 0xffff) / 65536.f; };
  auto gauss = [&]() { return sqrtf(-2.f * logf(rnd() + 1e-7f)) * cosf(6.2831853f * rnd()); };
  // x
  std::vector<__half> xq((size_t)B * Q * Hp * Wp * 4, __float2half(0.f)), xp((size_t)B * Q * Hp * Wp * 8, __float2half(0.f));
  for (int bq = 0; bq < B * Q; ++bq)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w)
        for (int c = 0; c < 4; ++c) xq[(((size_t)bq * Hp + h + 1) * Wp + w + 1) * 4 + c] = __float2half(gauss());
  for (int bq = 0; bq < B * Q; ++bq)
    for (int h = 0; h < Hp; ++h)
      for (int w = 0; w < Wp; ++w)
        for (int c = 0; c < 4; ++c) {
          xp[(((size_t)bq * Hp + h) * Wp + w) * 8 + c] = xq[(((size_t)bq * Hp + h) * Wp + w) * 4 + c];
          xp[(((size_t)bq * Hp + h) * Wp + w) * 8 + 4 + c] = w + 1 < Wp ? xq[(((size_t)bq * Hp + h) * Wp + w + 1) * 4 + c] : __float2half(0.f);
        }
  std::vector<__half> fields((size_t)B * Q * 9 * P * 4);
  for (int b = 0; b < B; ++b) {
    std::vector<float> mvy((H / 8) * (W / 8)), mvx((H / 8) * (W / 8));
    for (auto &v : mvy) v = (rnd() - 0.5f) * 3.f;
    for (auto &v : mvx) v = (rnd() - 0.5f) * 3.f;
    for (int k = 0; k < Q * 9; ++k)
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          const size_t i = (((size_t)b * Q * 9 + k) * P + h * W + w) * 4;
          const int blk = (h / 8) * (W / 8) + w / 8;
          fields[i + 0] = __float2half(mvy[blk] + jitter * gauss());
          fields[i + 1] = __float2half(mvx[blk] + jitter * gauss());
          fields[i + 2] = __float2half(rnd());
          fields[i + 3] = __float2half(0.f);
        }
  }
  Params p;
  p.B = B;
  void *d_xq, *d_xp, *d_f;
  CK(cudaMalloc(&d_xq, xq.size() * 2)); CK(cudaMalloc(&d_xp, xp.size() * 2)); CK(cudaMalloc(&d_f, fields.size() * 2));
  CK(cudaMemcpy(d_xq, xq.data(), xq.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_xp, xp.data(), xp.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_f, fields.data(), fields.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&p.out, (size_t)B * P * 4));
  p.xq = (const uint2 *)d_xq; p.xp = (const uint4 *)d_xp; p.fields = (const __half *)d_f;
  // layered 2D texture of half4 texels
  cudaChannelFormatDesc cd = cudaCreateChannelDescHalf4();
  cudaArray_t arr;
  CK(cudaMalloc3DArray(&arr, &cd, make_cudaExtent(Wp, Hp, B * Q), cudaArrayLayered));
  cudaMemcpy3DParms cp = {};
  cp.srcPtr = make_cudaPitchedPtr(xq.data(), Wp * 8, Wp, Hp);
  cp.dstArray = arr; cp.extent = make_cudaExtent(Wp, Hp, B * Q); cp.kind = cudaMemcpyHostToDevice;
  CK(cudaMemcpy3D(&cp));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  td.filterMode = cudaFilterModeLinear;
  CK(cudaCreateTextureObject(&p.tex_lin, &rd, &td, nullptr));
  td.filterMode = cudaFilterModePoint;
  CK(cudaCreateTextureObject(&p.tex_pt, &rd, &td, nullptr));

  {
    const size_t pitch = (((size_t)Wp * 8 + 127) / 128) * 128;
    void *d_xt; CK(cudaMalloc(&d_xt, pitch * Hp * B * Q));
    CK(cudaMemcpy2D(d_xt, pitch, d_xq, (size_t)Wp * 8, (size_t)Wp * 8, (size_t)Hp * B * Q, cudaMemcpyDeviceToDevice));
    cudaResourceDesc r2 = {}; r2.resType = cudaResourceTypePitch2D; r2.res.pitch2D.devPtr = d_xt; r2.res.pitch2D.desc = cd;
    r2.res.pitch2D.width = Wp; r2.res.pitch2D.height = (size_t)Hp * B * Q; r2.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc t2 = {};
    t2.addressMode[0] = t2.addressMode[1] = cudaAddressModeBorder; t2.readMode = cudaReadModeElementType; t2.normalizedCoords = 0;
    t2.filterMode = cudaFilterModeLinear;
    cudaError_t e = cudaCreateTextureObject(&p.tex_p2d, &r2, &t2, nullptr);
    printf("pitch2D texture %d x %zu pitch %d: %s\n", Wp, (size_t)Hp * B * Q, Wp * 8, cudaGetErrorString(e));
    if (e != cudaSuccess) { p.tex_p2d = p.tex_lin; cudaGetLastError(); }
  }
  std::vector<float> ref((size_t)B * P), got((size_t)B * P);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run = [&](int mode, const char *name) {
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      CK(cudaMemset(p.out, 0, (size_t)B * P * 4));
      CK(cudaEventRecord(e0));
      switch (mode) {
        case 0: probe<0><<<grid, 512>>>(p); break;
        case 1: probe<1><<<grid, 512>>>(p); break;
        case 2: probe<2><<<grid, 512>>>(p); break;
        case 3: probe<3><<<grid, 512>>>(p); break;
        case 4: probe<4><<<grid, 512>>>(p); break;
        case 5: probe<5><<<grid, 512>>>(p); break;
        case 6: probe<6><<<grid, 512>>>(p); break;
        case 7: probe<7><<<grid, 512>>>(p); break;
      }
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpy(got.data(), p.out, got.size() * 4, cudaMemcpyDeviceToHost));
    double err = 0, mag = 0;
    if (mode == 0) ref = got;
    for (size_t i = 0; i < got.size(); ++i) { err = fmax(err, fabs((double)got[i] - ref[i])); mag = fmax(mag, fabs((double)ref[i])); }
    printf("mode %d %-34s %8.1f us   %.1f Gsamples/s   max|diff vs mode0| %.4g (max|ref| %.3g)\n", mode, name, best * 1e3,
           (double)B * P * 144 / best / 1e6, err, mag);
  };
  printf("B=%d jitter=%.2f  samples=%.1fM\n", B, jitter, (double)B * P * 144 / 1e6);
  run(0, "4x LDG.64 quad-planar");
  run(1, "2x LDG.128 pair texels");
  run(4, "2x LDG.128 pair + HFMA2 blend");
  run(2, "1x TEX linear half4");
  run(3, "4x TEX point half4");
  run(5, "mixed 2 TEX(p2d) + 2 LSU pair/HFMA2");
  run(6, "mixed 3 TEX(p2d) + 1 LSU pair/HFMA2");
  run(7, "4 TEX(p2d folded) + HFMA2 mask");
  return 0;
}
