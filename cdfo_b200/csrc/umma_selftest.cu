// Hardware self-test of the tcgen05 plumbing every sm_100a kernel of this library relies on:
// canonical SWIZZLE_NONE K-major operand layout, shared-memory / instruction descriptors, TMEM alloc,
// tcgen05.mma + commit -> mbarrier, tcgen05.ld.  One CTA computes D[128,64] = A[128,64] * B[64,64]^T.
// tests/test_sm100_gpu.py compares D with a host matmul; a wrong descriptor convention shows up here,
// in isolation, instead of inside the gather-fed DCN kernel.
#include "cdfo_common.cuh"
#include "sm100_ptx.cuh"

namespace cdfo {

__global__ void __launch_bounds__(128) umma_selftest_kernel(const __nv_bfloat16 *__restrict__ A,
                                                            const __nv_bfloat16 *__restrict__ Bm, float *__restrict__ D,
                                                            int swap_lbo_sbo) {
  __shared__ __align__(128) uint8_t sA[128 * 64 * 2];  // [kc=8][m=128][8]
  __shared__ __align__(128) uint8_t sB[64 * 64 * 2];   // [kc=8][n=64][8]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid / 32;

  for (int e = tid; e < 128 * 8; e += 128) {  // 16-byte chunks of A
    const int m = e / 8, kc = e % 8;
    *reinterpret_cast<uint4 *>(sA + (kc * 128 + m) * 16) = *reinterpret_cast<const uint4 *>(A + m * 64 + kc * 8);
  }
  for (int e = tid; e < 64 * 8; e += 128) {
    const int n = e / 8, kc = e % 8;
    *reinterpret_cast<uint4 *>(sB + (kc * 64 + n) * 16) = *reinterpret_cast<const uint4 *>(Bm + n * 64 + kc * 8);
  }
  ptx::fence_proxy_async_smem();
  if (tid == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_d = tmem_base_slot;

  if (tid == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, 64);
    const uint32_t a_lbo = 128 * 16, a_sbo = 128, b_lbo = 64 * 16, b_sbo = 128;
    for (int j = 0; j < 4; ++j) {  // K = 16 per instruction = 2 core matrices along K
      const uint32_t a_addr = ptx::smem_u32(sA) + j * 2 * a_lbo;
      const uint32_t b_addr = ptx::smem_u32(sB) + j * 2 * b_lbo;
      const uint64_t ad = swap_lbo_sbo ? ptx::make_smem_desc(a_addr, a_sbo, a_lbo) : ptx::make_smem_desc(a_addr, a_lbo, a_sbo);
      const uint64_t bd = swap_lbo_sbo ? ptx::make_smem_desc(b_addr, b_sbo, b_lbo) : ptx::make_smem_desc(b_addr, b_lbo, b_sbo);
      ptx::umma_f16(tmem_d, ad, bd, idesc, j > 0);
    }
    ptx::umma_commit(ptx::smem_u32(&bar));
  }
  ptx::mbar_wait(ptx::smem_u32(&bar), 0);
  ptx::tc_fence_after();

  uint32_t r[32];
  for (int half = 0; half < 2; ++half) {
    ptx::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + half * 32, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[tid * 64 + half * 32 + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_d, 64);
}

}  // namespace cdfo

extern "C" int cdfo_umma_selftest(const void *A, const void *B, float *D, int swap_lbo_sbo, void *stream) {
  CDFO_REQUIRE(A && B && D, CDFO_ERR_NULL, "cdfo_umma_selftest: NULL pointer");
  cdfo::umma_selftest_kernel<<<1, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)A, (const __nv_bfloat16 *)B, D,
                                                                 swap_lbo_sbo);
  return cdfo::check_launch("cdfo_umma_selftest");
}
