// Host launcher for the fp64 DMMA GEMM (see gemm_dmma.cuh).
#include "gemm_dmma.cuh"

#include "common.h"

namespace dlaf_b200 {

void launch_gemm_nt_f64(const GemmArgs& a, cudaStream_t stream) {
  using Cfg = GemmCfg128;
  if (a.M <= 0 || a.N <= 0)
    return;
  DLAF_B200_ASSERT(a.M % Cfg::BM == 0 && a.N % Cfg::BN == 0 && a.K % Cfg::BK == 0 && a.K > 0,
                   "gemm shape must be a multiple of the CTA tile");
  DLAF_B200_ASSERT(a.lda % 2 == 0 && a.ldb % 2 == 0 && a.ldc % 2 == 0,
                   "16-byte aligned operand columns");
  DLAF_B200_ASSERT((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(a.B) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(a.C) & 15) == 0,
                   "16-byte aligned operands");
  DLAF_B200_ASSERT(a.a_ts % 2 == 0 && a.b_ts % 2 == 0, "16-byte aligned panel tiles");
  static bool configured = false;
  if (!configured) {
    DLAF_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_f64_kernel<Cfg>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES));
    configured = true;
  }
  dim3 grid(a.M / Cfg::BM, a.N / Cfg::BN);
  gemm_nt_f64_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(a);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

template <>
void launch_gemm_nt<double>(const GemmArgsT<double>& a, cudaStream_t stream) {
  launch_gemm_nt_f64(a, stream);
}

}  // namespace dlaf_b200
