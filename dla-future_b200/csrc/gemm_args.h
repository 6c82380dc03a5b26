// Argument block shared by all NT-GEMM kernels of the POTRF path (fp64 DMMA, fp32, complex).
//
//   C(MxN) = beta * C + alpha * A(MxK) * B(NxK)^H        all column-major
#pragma once

namespace dlaf_b200 {

enum GemmMask : int { kMaskNone = 0, kMaskLower = 1 };

template <class T>
struct GemmArgsT {
  const T* A;  // M x K, column-major
  long lda;
  const T* B;  // N x K, column-major
  long ldb;
  T* C;  // M x N, column-major (may alias A for the in-place TRSM step: N == one CTA column)
  long ldc;
  int M, N, K;
  double alpha, beta;  // real scalars (+-1, 0 on this path; herk takes real alpha/beta anyway)
  // Lower-triangular mask on GLOBAL tile coordinates (block-cyclic, tiles of nbp x nbp):
  // local tile row li -> global gi = li * P + prow, local tile col lj -> gj = lj * Q + pcol.
  int mask;
  int nbp;
  int P, Q, prow, pcol;
  int ti0, tj0;  // local tile index of C(0,0)
  int ri0, ci0;  // element offset of C(0,0) inside that tile (sub-tile calls)
  // Panel workspaces are stored tile by tile (each nbp x nbp tile contiguous, like the reference's
  // Panel<..> with AllocationLayout::Tiles, matrix/panel.h:392): row r of A lives at
  // A + (r / nbp) * a_ts + r % nbp. 0 = plain column-major operand.
  long a_ts, b_ts;
  int dbg_stagger_ns;  // measurement aid: delay the second resident wave of CTAs by this many ns
};

// Tile classification against the lower-triangular mask.
//  0: skip (entirely above the diagonal)   1: full   2: straddles the diagonal (element mask)
template <class T>
__host__ __device__ inline int classify_tile(const GemmArgsT<T>& p, int row0, int col0, int BM, int BN,
                                             long& grow0, long& gcol0) {
  if (p.mask == kMaskNone) {
    grow0 = row0;
    gcol0 = col0;
    return 1;
  }
  const int r = p.ri0 + row0, c = p.ci0 + col0;
  const long gi = static_cast<long>(p.ti0 + r / p.nbp) * p.P + p.prow;
  const long gj = static_cast<long>(p.tj0 + c / p.nbp) * p.Q + p.pcol;
  grow0 = gi * p.nbp + r % p.nbp;
  gcol0 = gj * p.nbp + c % p.nbp;
  if (grow0 + BM - 1 < gcol0)
    return 0;  // last row above first column -> nothing in the lower triangle
  if (grow0 >= gcol0 + BN - 1)
    return 1;
  return 2;
}

}  // namespace dlaf_b200

#include <cuda_runtime.h>

namespace dlaf_b200 {
// Fused panel TRSM (fp64, gemm_dmma.cuh: trsm_fused_f64_kernel): B (m x ns*128 row panel) <- B * L^-T in ONE launch.
struct TrsmFusedArgs {
  double* B;        // m x ns*G row panel, column-major
  long ldb;
  const double* T;  // factored diagonal tile (lower), column-major
  long ldt;
  const double* W;  // ns inverted diagonal blocks, each G x G contiguous (ld = G)
  int ns;
};
// m % 32 == 0, 16-byte aligned operands with even leading dimensions.
void launch_trsm_fused_f64(const TrsmFusedArgs& args, int m, cudaStream_t stream);

// Native fp64 (DMMA) product executed only if *flag != 0 when the kernel starts (gemm_dmma.cu): the fallback of the
// int8-digit trailing update when its guard fires (gemm_ozaki.h).
void launch_gemm_nt_f64_if(const GemmArgsT<double>& args, const int* flag, cudaStream_t stream);

// complex<double> flavour (gemm_zdmma.cu: trsm_fused_z_kernel), 64 x 64 diagonal blocks, m % 64 == 0.
void launch_trsm_fused_z(double2* b, long ldb, int m, const double2* t, long ldt, const double2* w, int ns,
                         cudaStream_t stream);

// Kernel entry point per element type (explicit specialisations live next to the kernels).
template <class T>
void launch_gemm_nt(const GemmArgsT<T>& args, cudaStream_t stream);
}  // namespace dlaf_b200
