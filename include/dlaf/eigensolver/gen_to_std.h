// dlaf/eigensolver/gen_to_std.h — same names and template parameters as the reference's
// include/dlaf/eigensolver/gen_to_std.h:50-75 (local) and :101-127 (distributed):
//
//   template <Backend B, Device D, class T>
//   void eigensolver::internal::generalized_to_standard(blas::Uplo, Matrix<T, D>& mat_a, Matrix<T, D>& mat_b);
//   ... (comm::CommunicatorGrid&, blas::Uplo, Matrix<T, D>& mat_a, Matrix<T, D>& mat_b);
//
// mat_a: Hermitian A, its `uplo` triangle is overwritten with that of inv(L) A inv(L)^H / inv(U)^H A inv(U); mat_b: the
// Cholesky factor of B in its `uplo` triangle — never modified here (the reference may modify its diagonal tiles
// temporarily, hence the non-const reference in the signature, kept for source compatibility). Backend::GPU / Device::GPU
// only. Returns when the result is complete.
#pragma once

#include <dlaf/inverse/cholesky.h>

namespace dlaf::eigensolver::internal {
namespace detail {
inline int call(int c, char u, float* a, DLAF_descriptor da, const float* b, DLAF_descriptor db, cudaStream_t s) { return dlaf_b200_generalized_to_standard_device_s(c, u, a, da, b, db, s); }
inline int call(int c, char u, double* a, DLAF_descriptor da, const double* b, DLAF_descriptor db, cudaStream_t s) { return dlaf_b200_generalized_to_standard_device_d(c, u, a, da, b, db, s); }
inline int call(int c, char u, std::complex<float>* a, DLAF_descriptor da, const std::complex<float>* b, DLAF_descriptor db, cudaStream_t s) { return dlaf_b200_generalized_to_standard_device_c(c, u, a, da, b, db, s); }
inline int call(int c, char u, std::complex<double>* a, DLAF_descriptor da, const std::complex<double>* b, DLAF_descriptor db, cudaStream_t s) { return dlaf_b200_generalized_to_standard_device_z(c, u, a, da, b, db, s); }
}  // namespace detail

template <Backend B, Device D, class T>
void generalized_to_standard(comm::CommunicatorGrid& grid, const blas::Uplo uplo, Matrix<T, D>& mat_a, Matrix<T, D>& mat_b) {
  static_assert(B == Backend::GPU && D == Device::GPU, "this build provides Backend::GPU / Device::GPU only");
  dlaf::internal::require_same_grid(grid, mat_a, "generalized_to_standard");
  dlaf::internal::require_same_grid(grid, mat_b, "generalized_to_standard");
  mat_b.waitLocalTiles();  // the factor may still be in flight on its own stream
  detail::call(mat_a.context(), dlaf::internal::uplo_char(uplo), mat_a.ptr(), mat_a.descriptor(), mat_b.ptr(), mat_b.descriptor(),
               mat_a.stream());
}

template <Backend B, Device D, class T>
void generalized_to_standard(const blas::Uplo uplo, Matrix<T, D>& mat_a, Matrix<T, D>& mat_b) {
  static_assert(B == Backend::GPU && D == Device::GPU, "this build provides Backend::GPU / Device::GPU only");
  mat_b.waitLocalTiles();
  detail::call(mat_a.context(), dlaf::internal::uplo_char(uplo), mat_a.ptr(), mat_a.descriptor(), mat_b.ptr(), mat_b.descriptor(),
               mat_a.stream());
}
}  // namespace dlaf::eigensolver::internal
