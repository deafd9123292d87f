// Offset / mask head arithmetic of MVDualAttAlignment (arch/SIDECVSR_our.py:3341-3350), shared by every kernel that evaluates it
// (conv3x3_sm100.cu: two-launch and dual-launch heads; mv_dcn_fused_sm100.cu: head + DCN in one kernel) so that they agree BIT FOR
// BIT: every step is an explicit round-to-nearest op (no contraction into FMAs that the compiler might apply in one kernel and not
// in the other), the hardware approximations (tanh.approx, ex2.approx, rcp.approx) are deterministic.
//   offset = mag * tanh(o_1) + mag * tanh(o_2)  (+ flow, added by the DCN kernel in the reference's order)      arch:3345-3347
//   mask   = sigmoid(m_1 + m_2)                                                                                  arch:3350
// Between the two evaluations the first one is held as fp16 (dy, dx | m, 0) -- the "field" the two-launch path stores in HBM.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace cdfo {
namespace head {

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint2 pack_field(float dy, float dx, float m) {
  const __half2 h0 = __floats2half2_rn(dy, dx), h1 = __floats2half2_rn(m, 0.f);
  return make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
}
// first evaluation: (mag tanh(dy), mag tanh(dx), m) from the biased conv outputs
__device__ __forceinline__ uint2 first(float vdy, float vdx, float vm, float mag) {
  return pack_field(__fmul_rn(tanh_approx(vdy), mag), __fmul_rn(tanh_approx(vdx), mag), vm);
}
// second evaluation combined with the stored first one: (dy_1 + mag tanh(dy), dx_1 + mag tanh(dx), sigmoid(m_1 + m))
__device__ __forceinline__ uint2 second(uint2 prior, float vdy, float vdx, float vm, float mag) {
  const float2 pd = __half22float2(*reinterpret_cast<const __half2 *>(&prior.x));
  const float pm = __low2float(*reinterpret_cast<const __half2 *>(&prior.y));
  const float dy = __fadd_rn(__fmul_rn(tanh_approx(vdy), mag), pd.x);
  const float dx = __fadd_rn(__fmul_rn(tanh_approx(vdx), mag), pd.y);
  const float m = __fdividef(1.f, __fadd_rn(1.f, __expf(-__fadd_rn(pm, vm))));
  return pack_field(dy, dx, m);
}

}  // namespace head
}  // namespace cdfo
