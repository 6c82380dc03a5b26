// dlaf/inverse/cholesky.h — same names and template parameters as the reference's include/dlaf/inverse/cholesky.h:38-83:
//
//   template <Backend B, Device D, class T> void inverse_from_cholesky_factor(blas::Uplo, Matrix<T, D>&);
//   template <Backend B, Device D, class T> void inverse_from_cholesky_factor(comm::CommunicatorGrid&, blas::Uplo, Matrix<T, D>&);
//
// On entry the matrix holds the Cholesky factor in the `uplo` triangle, on exit that triangle of inv(A); only that
// triangle is accessed. Backend::GPU / Device::GPU only (no MC backend, no CPU fallback). Unlike the factorization the
// call returns when the result is complete (the algorithm allocates and releases its workspaces per call).
#pragma once

#include <complex>
#include <cstdio>
#include <cstdlib>

#include <dlaf/communication/communicator_grid.h>
#include <dlaf/factorization/cholesky.h>
#include <dlaf/matrix/matrix.h>
#include <dlaf/types.h>
#include <dlaf_c/b200_ext.h>

namespace dlaf {
namespace internal {
inline int call_inverse(int c, int ph, char u, char d, float* a, DLAF_descriptor ds, cudaStream_t s) { return dlaf_b200_inverse_device_s(c, ph, u, d, a, ds, s); }
inline int call_inverse(int c, int ph, char u, char d, double* a, DLAF_descriptor ds, cudaStream_t s) { return dlaf_b200_inverse_device_d(c, ph, u, d, a, ds, s); }
inline int call_inverse(int c, int ph, char u, char d, std::complex<float>* a, DLAF_descriptor ds, cudaStream_t s) { return dlaf_b200_inverse_device_c(c, ph, u, d, a, ds, s); }
inline int call_inverse(int c, int ph, char u, char d, std::complex<double>* a, DLAF_descriptor ds, cudaStream_t s) { return dlaf_b200_inverse_device_z(c, ph, u, d, a, ds, s); }
template <class M>
void require_same_grid(comm::CommunicatorGrid& grid, M& mat, const char* what) {
  if (grid.context() != mat.context()) {  // equal_process_grid of the reference
    std::fprintf(stderr, "[dlaf] %s: the matrix is not distributed on the given communicator grid\n", what);
    std::abort();
  }
}
}  // namespace internal

template <Backend B, Device D, class T>
void inverse_from_cholesky_factor(comm::CommunicatorGrid& grid, const blas::Uplo uplo, Matrix<T, D>& mat_a) {
  static_assert(B == Backend::GPU && D == Device::GPU, "this build provides Backend::GPU / Device::GPU only");
  internal::require_same_grid(grid, mat_a, "inverse_from_cholesky_factor");
  internal::call_inverse(mat_a.context(), 3, internal::uplo_char(uplo), 'N', mat_a.ptr(), mat_a.descriptor(), mat_a.stream());
}

template <Backend B, Device D, class T>
void inverse_from_cholesky_factor(const blas::Uplo uplo, Matrix<T, D>& mat_a) {
  static_assert(B == Backend::GPU && D == Device::GPU, "this build provides Backend::GPU / Device::GPU only");
  internal::call_inverse(mat_a.context(), 3, internal::uplo_char(uplo), 'N', mat_a.ptr(), mat_a.descriptor(), mat_a.stream());
}
}  // namespace dlaf
