// dlaf::triangular_solver — mirror of include/dlaf/solver/triangular.h:31-134 of the reference over this library's C ABI
// (dlaf_b200_triangular_solver_*, include/dlaf_c/b200_ext.h). Same argument meaning:
//   op(A) X = alpha B  (side == Left)   or   X op(A) = alpha B  (side == Right),   B overwritten with X,
// A triangular (uplo, diag), square with square blocks, distributed on the same grid as B. The host matrices are the
// caller's local parts; like the reference's C entry points for the other algorithms the call is synchronous.
#pragma once

#include <complex>

#include <dlaf/communication/communicator_grid.h>
#include <dlaf/matrix/matrix.h>
#include <dlaf/types.h>
#include <dlaf_c/b200_ext.h>

namespace dlaf {
namespace blas_like {
enum class Side : char { Left = 'L', Right = 'R' };
enum class Op : char { NoTrans = 'N', Trans = 'T', ConjTrans = 'C' };
enum class Diag : char { NonUnit = 'N', Unit = 'U' };
}  // namespace blas_like

namespace internal {
inline int trsm_call(int c, char s, char u, char o, char d, const float* al, const float* a, DLAF_descriptor da, float* b, DLAF_descriptor db) { return dlaf_b200_triangular_solver_s(c, s, u, o, d, al, a, da, b, db); }
inline int trsm_call(int c, char s, char u, char o, char d, const double* al, const double* a, DLAF_descriptor da, double* b, DLAF_descriptor db) { return dlaf_b200_triangular_solver_d(c, s, u, o, d, al, a, da, b, db); }
inline int trsm_call(int c, char s, char u, char o, char d, const std::complex<float>* al, const std::complex<float>* a, DLAF_descriptor da, std::complex<float>* b, DLAF_descriptor db) { return dlaf_b200_triangular_solver_c(c, s, u, o, d, al, a, da, b, db); }
inline int trsm_call(int c, char s, char u, char o, char d, const std::complex<double>* al, const std::complex<double>* a, DLAF_descriptor da, std::complex<double>* b, DLAF_descriptor db) { return dlaf_b200_triangular_solver_z(c, s, u, o, d, al, a, da, b, db); }
}  // namespace internal

/// Distributed (or 1 x 1) triangular solve on host-resident local matrices. `uplo_char` is 'L' or 'U'.
template <class T>
void triangular_solver(comm::CommunicatorGrid& grid, blas_like::Side side, char uplo_char, blas_like::Op op,
                       blas_like::Diag diag, T alpha, const T* a_local, DLAF_descriptor desc_a, T* b_local,
                       DLAF_descriptor desc_b) {
  internal::trsm_call(grid.context(), static_cast<char>(side), uplo_char, static_cast<char>(op), static_cast<char>(diag), &alpha,
                      a_local, desc_a, b_local, desc_b);
}
}  // namespace dlaf
