// Reduction of a Hermitian-definite generalized eigenproblem to standard form on the GPU grid:
//     A <- inv(L) A inv(L)^H   (uplo 'L', B = L L^H)       A <- inv(U)^H A inv(U)   (uplo 'U', B = U^H U)
//
// Replaces dlaf::eigensolver::internal::generalized_to_standard<Backend::GPU, Device::GPU, T>
// (include/dlaf/eigensolver/gen_to_std.h:50-75, :101-127; GenToStd<B,D,T>::call_L / call_U local and distributed,
// eigensolver/gen_to_std/impl.h:238-281, :283-505, :507-568, :570-769): the blocked xHEGST (itype 1) as tile loops of
// cusolver hegst / cublas trsm / hemm / her2k / gemm calls, whose last part — the solve of each column panel against the
// trailing factor (impl.h:270-280) — is a chain of nt-k dependent tile TRSMs + GEMMs per step.
//
// Here (lower case; 'U' runs the same on the conjugate-transposed problem with the grid roles swapped):
//   phase 1, k = 0 .. nt-1: the diagonal tile  A_kk <- inv(L_kk) A_kk inv(L_kk)^H  (two panel substitutions on the full
//       Hermitian tile), the column panel  P = A(i>k,k) inv(L_kk)^H - 1/2 L(i>k,k) A_kk,  the trailing matrix
//       A(i,j) -= P(i) L(j,k)^H + L(i,k) P(j)^H  (i >= j > k) as TWO lower-masked launches on the POTRF update engine, then
//       the second half  P -= 1/2 L(i>k,k) A_kk.
//   phase 2: the panel solves do not feed back into later steps, so ALL of them are deferred and done as ONE
//       strictly-lower triangular sweep  X = inv(L) C  restricted to the blocks below the diagonal: for j = 1 .. nt-1
//       X(j, :j) = inv(L_jj) C(j, :j)  (one fused panel substitution on the transposed row),
//       C(t > j, :j) -= L(t,j) X(j, :j)  ONE launch — n^3/3 of the n^3 flops moved from a latency-bound chain to the bulk engine.
// fp64 products run on tcgen05 as exact int8 digit planes with the data-dependent guard of POTRF (gemm_ozaki.h), fp32 as
// 3xTF32, complex on the native kernels. L is never modified (the reference may modify its diagonal tiles temporarily).
//
// Communication per step on a P x Q grid (NCCL, device-direct; reference: impl.h:337-343, :370-376, :424-445):
//   phase 1: [L_kk | inverted blocks | A_kk] down the process column of block column k; the pair of column panels (P, L)
//   along the process rows; their tiles j to process column j % Q from the rank holding the diagonal tile (j,j);
//   phase 2: [L_jj | inverted blocks] along process row j % P; the solved row down the process columns; L(t>j, j) along
//   the process rows.
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>

#include "types.h"

namespace dlaf_b200 {

struct HegstProblem {
  char uplo = 'L';
  long n = 0;
  int nb = 1;
  // USER grid and my source-adjusted ("virtual") coordinates in it (A and L are distributed identically)
  int P = 1, Q = 1, prow = 0, pcol = 0;
  int src_row = 0, src_col = 0;  // NCCL ranks of virtual coordinate 0 inside col_comm (size P) / row_comm (size Q)
};

// In place on the DEVICE copy of the local part of A (user layout, column-major, lda): only the `uplo` triangle is read
// and written. l: local part of the Cholesky factor (read only, `uplo` triangle). Collective over the grid;
// asynchronous on `stream` except for workspace allocation / release. Returns the number of kernels launched;
// *guard_steps (may be null): steps whose fp64 update ran on the native kernel because the int8 digit guard fired.
template <class T>
long generalized_to_standard_device(const HegstProblem& p, T* a, long lda, const T* l, long ldl, ncclComm_t row_comm,
                                    ncclComm_t col_comm, cudaStream_t stream, int* guard_steps = nullptr);

}  // namespace dlaf_b200
