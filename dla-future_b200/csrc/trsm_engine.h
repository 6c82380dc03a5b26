// Distributed triangular solve on the GPU grid — the step right after POTRF in every consumer.
//
// Replaces dlaf::triangular_solver<Backend::GPU, Device::GPU, T> local and distributed
// (include/dlaf/solver/triangular.h:31-134; the eight loop nests call_LLN .. call_RUT of solver/triangular/impl.h:236-1205,
// each a tile loop of cublas?trsm + cublas?gemm calls, blas/tile.h:238-261, :337-349) by ONE schedule:
//
//   every (side, uplo, op) combination is the same problem   Y <- c Y M^-1   with a triangular M:
//       Right:  X op(A) = alpha B      ->  Y = B   (m x n),  M = op(A),    c = alpha
//       Left :  op(A) X = alpha B      ->  Y = B^H (n x m),  M = op(A)^H,  c = conj(alpha)   (X = Y^H at the end)
//   and with G := M^H (one of A, conj(A), A^T, A^H) every product has the NT shape the POTRF kernels provide:
//       for k forward (G lower) or backward (G upper):
//           Y_k <- Y_k G_kk^-H                      panel TRSM against pre-inverted 128-blocks (like POTRF's panel)
//           Y_t <- Y_t - Y_k G(t,k)^H   for the remaining block columns t      ONE GEMM launch per step
//   The transposition of B for Side::Left happens at the boundary together with a swap of the process-grid roles (the
//   local part of B^H is the local transpose of the local part of B on the transposed grid), exactly like uplo == 'U' in
//   the factorization; the tiles of G are taken from the caller's A where they lie and transposed / conjugated while they
//   are packed for the broadcast.
//
// Communication per step on a P x Q grid (NCCL, device-direct): the diagonal tile + its inverted blocks down the process
// column that holds Y_k; the solved Y_k along the process rows; the tiles G(t,k) to the process columns that hold Y_t —
// either straight down the column (when the stored tile already sits in that column) or along the row first and then
// down the column (the POTRF "transposed panel" pattern, communication/broadcast_panel.h:107-188).
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>

#include "gemm_args.h"
#include "types.h"

namespace dlaf_b200 {

struct TrsmProblem {
  char side = 'L', uplo = 'L', op = 'N', diag = 'N';
  long m = 0, n = 0;   // B is m x n; A is m x m (Left) or n x n (Right)
  int mb = 1, nb = 1;  // block sizes of B (rows, columns); A's blocks are mb x mb (Left) or nb x nb (Right)
  // USER grid and my source-adjusted ("virtual") coordinates in it
  int P = 1, Q = 1, prow = 0, pcol = 0;
  int src_row = 0, src_col = 0;  // NCCL ranks of virtual coordinate 0 inside col_comm (size P) / row_comm (size Q)
};

// Solves in place on DEVICE copies of the local parts (user layout, column-major): a (lda), b (ldb); alpha by value as
// (re, im). row_comm: ranks of my process row (size Q); col_comm: ranks of my process column (size P). Collective over
// the grid; asynchronous on `stream` except for workspace allocation. Returns the number of kernels launched.
template <class T>
long triangular_solve_device(const TrsmProblem& p, double alpha_re, double alpha_im, const T* a, long lda, T* b, long ldb,
                             ncclComm_t row_comm, ncclComm_t col_comm, cudaStream_t stream);

}  // namespace dlaf_b200
