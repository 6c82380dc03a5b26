// Host launcher for the fp64 DMMA GEMM (see gemm_dmma.cuh).
#include "gemm_dmma.cuh"

#include <cstdlib>

#include "common.h"

namespace dlaf_b200 {

namespace {

template <class Cfg>
void launch_cfg(const GemmArgs& a, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    DLAF_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_f64_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES));
    configured = true;
  }
  dim3 grid(a.M / Cfg::BM, a.N / Cfg::BN);
  gemm_nt_f64_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(a);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

void launch_gemm_nt_f64_cfg(const GemmArgs& a, int cfg, cudaStream_t stream) {
  switch (cfg) {
    case 0: launch_cfg<GemmCfg128>(a, stream); break;
    case 1: launch_cfg<GemmCfg64x128>(a, stream); break;
    case 2: launch_cfg<GemmCfg64>(a, stream); break;
    case 3: launch_cfg<GemmCfg128k32>(a, stream); break;
    case 4: launch_cfg<GemmCfg128x64>(a, stream); break;
    case 5: launch_cfg<GemmCfg64w4s3>(a, stream); break;
    case 6: launch_cfg<GemmCfg64w4s4>(a, stream); break;
    case 7: launch_cfg<GemmCfg32x128w4>(a, stream); break;
    default: DLAF_B200_ASSERT(false, "unknown gemm configuration");
  }
}

void launch_gemm_nt_f64(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0)
    return;
  DLAF_B200_ASSERT(a.M % 128 == 0 && a.N % 128 == 0 && a.K % 16 == 0 && a.K > 0,
                   "gemm shape must be a multiple of the CTA tile");
  DLAF_B200_ASSERT(a.lda % 2 == 0 && a.ldb % 2 == 0 && a.ldc % 2 == 0,
                   "16-byte aligned operand columns");
  DLAF_B200_ASSERT((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(a.B) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(a.C) & 15) == 0,
                   "16-byte aligned operands");
  DLAF_B200_ASSERT(a.a_ts % 2 == 0 && a.b_ts % 2 == 0, "16-byte aligned panel tiles");
  // Tile choice (measured, profiles/r01_kernel_test_6_4warp_ctas.log): 64x64 tiles on 4-warp CTAs, four
  // resident per SM, beat the classic 128x128 / 8-warp tile (33.9 vs 29.4 TFLOP/s): barriers, pipeline
  // fill and epilogue of one CTA are covered by the other three. The same kernel serves the small
  // latency-critical GEMMs of the panel chain (4x more CTAs than 128x128 tiles).
  static const int bulk_cfg = [] {
    const char* e = std::getenv("DLAF_B200_GEMM_BULK_CFG");
    return e ? std::atoi(e) : 5;
  }();
  const bool in_place = (static_cast<const void*>(a.A) == static_cast<const void*>(a.C));
  DLAF_B200_ASSERT(!in_place || a.N == 128, "in-place product needs one CTA column");
  if (in_place)
    // C aliases A: one CTA must own all N columns of its rows (N == 128 == BN)
    launch_cfg<GemmCfg32x128w4>(a, stream);
  else
    launch_gemm_nt_f64_cfg(a, bulk_cfg, stream);
}

void launch_gemm_nt_f64_if(const GemmArgs& a, const int* flag, cudaStream_t stream) {
  using Cfg = GemmCfg64w4s3;
  if (a.M <= 0 || a.N <= 0)
    return;
  DLAF_B200_ASSERT(a.M % Cfg::BM == 0 && a.N % Cfg::BN == 0 && a.K % Cfg::BK == 0 && a.K > 0 && flag != nullptr,
                   "guarded gemm shape");
  DLAF_B200_ASSERT(a.lda % 2 == 0 && a.ldb % 2 == 0 && a.ldc % 2 == 0 && a.a_ts % 2 == 0 && a.b_ts % 2 == 0,
                   "16-byte aligned operand columns");
  static int ctas = 0;
  if (ctas == 0) {
    DLAF_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_f64_if_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES));
    int dev = 0, nsm = 0;
    DLAF_CUDA_CHECK(cudaGetDevice(&dev));
    DLAF_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    ctas = nsm * Cfg::MINB;
  }
  const long tiles = static_cast<long>(a.M / Cfg::BM) * (a.N / Cfg::BN);
  const unsigned grid = static_cast<unsigned>(tiles < ctas ? tiles : ctas);
  gemm_nt_f64_if_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(a, flag);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

namespace {
template <class Cfg>
void launch_trsm_cfg(const TrsmFusedArgs& a, int m, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    DLAF_CUDA_CHECK(cudaFuncSetAttribute(trsm_fused_f64_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES));
    configured = true;
  }
  trsm_fused_f64_kernel<Cfg><<<m / Cfg::BM, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(a);
  DLAF_CUDA_CHECK(cudaGetLastError());
}
}  // namespace

void launch_trsm_fused_f64(const TrsmFusedArgs& a, int m, cudaStream_t stream) {
  if (m <= 0 || a.ns <= 0)
    return;
  DLAF_B200_ASSERT(m % 32 == 0, "fused TRSM: rows must be a multiple of 32");
  DLAF_B200_ASSERT(a.ldb % 2 == 0 && a.ldt % 2 == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(a.T) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0,
                   "fused TRSM: 16-byte aligned operands");
  // Short panels (the single critical tile of the two-chain schedule: 512 rows) are bound by the DMMA rate of the few SMs
  // they occupy (16 CTAs x 4 sequential substitution phases = 75 us): 16-row CTAs put them on twice as many SMs.
  // DLAF_B200_TRSM_BM=32 keeps the 32-row CTAs everywhere.
  static const bool small_ok = [] {
    const char* e = std::getenv("DLAF_B200_TRSM_BM");
    return e == nullptr || std::atoi(e) != 32;
  }();
  if (small_ok && m <= 2048)
    launch_trsm_cfg<GemmCfg16x128w4>(a, m, stream);
  else
    launch_trsm_cfg<GemmCfg32x128w4>(a, m, stream);
}

template <>
void launch_gemm_nt<double>(const GemmArgsT<double>& a, cudaStream_t stream) {
  launch_gemm_nt_f64(a, stream);
}

}  // namespace dlaf_b200
