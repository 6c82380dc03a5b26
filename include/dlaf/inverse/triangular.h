// dlaf/inverse/triangular.h — same names and template parameters as the reference's include/dlaf/inverse/triangular.h:38-76:
//
//   template <Backend B, Device D, class T> void triangular_inverse(blas::Uplo, blas::Diag, Matrix<T, D>&);
//   template <Backend B, Device D, class T> void triangular_inverse(comm::CommunicatorGrid&, blas::Uplo, blas::Diag, Matrix<T, D>&);
//
// The `uplo` triangle is overwritten with the inverse of that triangular matrix (Diag::Unit: the diagonal is assumed to
// be 1 and is neither read nor written). Backend::GPU / Device::GPU only. Returns when the result is complete.
#pragma once

#include <dlaf/inverse/cholesky.h>

namespace blas {
enum class Diag : char { NonUnit = 'N', Unit = 'U' };  // blaspp's blas::Diag, next to blas::Uplo (dlaf/types.h)
}

namespace dlaf {

template <Backend B, Device D, class T>
void triangular_inverse(comm::CommunicatorGrid& grid, const blas::Uplo uplo, const blas::Diag diag, Matrix<T, D>& mat_a) {
  static_assert(B == Backend::GPU && D == Device::GPU, "this build provides Backend::GPU / Device::GPU only");
  internal::require_same_grid(grid, mat_a, "triangular_inverse");
  internal::call_inverse(mat_a.context(), 1, internal::uplo_char(uplo), static_cast<char>(diag), mat_a.ptr(), mat_a.descriptor(),
                         mat_a.stream());
}

template <Backend B, Device D, class T>
void triangular_inverse(const blas::Uplo uplo, const blas::Diag diag, Matrix<T, D>& mat_a) {
  static_assert(B == Backend::GPU && D == Device::GPU, "this build provides Backend::GPU / Device::GPU only");
  internal::call_inverse(mat_a.context(), 1, internal::uplo_char(uplo), static_cast<char>(diag), mat_a.ptr(), mat_a.descriptor(),
                         mat_a.stream());
}
}  // namespace dlaf
