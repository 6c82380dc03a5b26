// Inverse of a triangular matrix and inverse of a Hermitian positive definite matrix from its Cholesky factor on the GPU
// grid.
//
// Replaces dlaf::triangular_inverse<Backend::GPU, Device::GPU, T> (include/dlaf/inverse/triangular.h:38-76; loop nests
// Triangular<B,D,T>::call_L / call_U, inverse/triangular/impl.h:183-229, :231-365, :367-413, :415-549) and
// dlaf::inverse_from_cholesky_factor<Backend::GPU, Device::GPU, T> (include/dlaf/inverse/cholesky.h:38-83 = the triangular
// inverse followed by AssembleCholeskyInverse<B,D,T>::call_L / call_U, inverse/cholesky/impl.h:180-224, :226-359,
// :361-405, :407-540) — tile loops of cublas?trsm / ?gemm / ?trmm / ?herk calls plus trtri / lauum tile kernels — by two
// sweeps that consist of ONE GEMM launch per step on the engines of the POTRF trailing update:
//
//   lower case (the upper one is the same on the conjugate-transposed problem with the process-grid roles swapped,
//   like uplo == 'U' of the factorization):
//
//   (1) W = L^-1, k = nt-1 .. 0 (the reference's order, impl.h:199-228): with Wh_k = L_kk^-H (all diagonal tiles are
//       inverted up front — they are never modified before their own step — by the panel substitution kernel applied to
//       an identity tile),
//           column panel   A(i,k) <- -A(i,k) L_kk^-1              i > k      one GEMM against Wh_k
//           trailing       A(i,j) <- A(i,j) + A(i,k) L(k,j)       i > k > j
//           row panel      A(k,j) <- L_kk^-1 L(k,j)               j < k
//           diagonal       A(k,k) <- L_kk^-1
//       where trailing update and row panel are ONE launch  C(i >= k, j < k) -= V B^H  with the extended column panel
//       V = [L_kk^-1 ; A(i>k,k)] and B(j) = -L(k,j)^H, after row k has been packed into B and zeroed.
//
//   (2) A^-1 = W^H W, k = 0 .. nt-1 (impl.h:195-223): with P(j) = W(k,j)^H, j <= k (the diagonal tile included), packed and
//       row k zeroed,  C(i <= k, j <= i) += P(i) P(j)^H  — the herk / gemm updates of the leading triangle, the trmm of
//       row k and the lauum of the diagonal tile in ONE lower-masked launch.
//
// fp64 runs these launches on tcgen05 as exact int8 digit products (gemm_ozaki.h) with the same guard and native
// fallback as POTRF; fp32 as 3xTF32 on tcgen05 (gemm_tf32.h); complex on the native DMMA / SIMT kernels.
//
// Communication per step on a P x Q grid (NCCL, device-direct; the reference's panel broadcasts
// impl.h:268-281, :296-309 and cholesky/impl.h:266-276): (1) Wh_k down the process column of block column k, the extended
// column panel along the process rows, the packed row-k tiles down the process columns; (2) the packed row-k tiles down
// the process columns, then each one from the rank that holds the diagonal tile of its block column along that rank's
// process row (the transposed-panel pattern of POTRF).
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>

#include "types.h"

namespace dlaf_b200 {

struct InverseProblem {
  char uplo = 'L', diag = 'N';  // diag only matters for the triangular inverse
  long n = 0;
  int nb = 1;
  // USER grid and my source-adjusted ("virtual") coordinates in it
  int P = 1, Q = 1, prow = 0, pcol = 0;
  int src_row = 0, src_col = 0;  // NCCL ranks of virtual coordinate 0 inside col_comm (size P) / row_comm (size Q)
};

enum InversePhases : int { kTriangularInverse = 1, kAssembleFromInverseFactor = 2, kInverseFromCholeskyFactor = 3 };

// In place on the DEVICE copy of the local part (user layout, column-major, lda): only the `uplo` triangle is read and
// written (with diag == 'U' the diagonal is neither read nor written). row_comm: ranks of my process row (size Q);
// col_comm: ranks of my process column (size P). Collective over the grid; asynchronous on `stream` except for
// workspace allocation / release. Returns the number of kernels launched; *guard_steps (may be null) receives the number
// of steps whose fp64 update ran on the native kernel because the int8 digit guard fired.
template <class T>
long inverse_device(const InverseProblem& p, int phases, T* a, long lda, ncclComm_t row_comm, ncclComm_t col_comm,
                    cudaStream_t stream, int* guard_steps = nullptr);

}  // namespace dlaf_b200
