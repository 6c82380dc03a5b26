// dlaf/factorization/cholesky.h — the public C++ entry points of the path, same names and template
// parameters as the reference (include/dlaf/factorization/cholesky.h:41-52 local, :71-83 distributed):
//
//   template <Backend B, Device D, class T> void cholesky_factorization(blas::Uplo, Matrix<T, D>&);
//   template <Backend B, Device D, class T> void cholesky_factorization(comm::CommunicatorGrid&, blas::Uplo, Matrix<T, D>&);
//
// Asynchronous like the reference: the call returns once the work is enqueued (on the matrix's stream);
// completion is observed with mat.waitLocalTiles(). Only the `uplo` triangle is read and written.
// Backend::GPU / Device::GPU is the product. Backend::MC (CPU) does not exist in this build — asking for it is
// a compile-time error, not a silent fallback. A matrix that is not positive definite is reported through
// cholesky_info(mat) (LAPACK-style), where the reference asserts / traps.
#pragma once

#include <complex>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include <dlaf/communication/communicator_grid.h>
#include <dlaf/matrix/matrix.h>
#include <dlaf/types.h>
#include <dlaf_c/b200_ext.h>

namespace dlaf {

namespace internal {
inline int call_device(int ctx, char uplo, float* a, DLAF_descriptor d, cudaStream_t s) {
  return dlaf_b200_cholesky_factorization_device_s(ctx, uplo, a, d, s);
}
inline int call_device(int ctx, char uplo, double* a, DLAF_descriptor d, cudaStream_t s) {
  return dlaf_b200_cholesky_factorization_device_d(ctx, uplo, a, d, s);
}
inline int call_device(int ctx, char uplo, std::complex<float>* a, DLAF_descriptor d, cudaStream_t s) {
  return dlaf_b200_cholesky_factorization_device_c(ctx, uplo, a, d, s);
}
inline int call_device(int ctx, char uplo, std::complex<double>* a, DLAF_descriptor d, cudaStream_t s) {
  return dlaf_b200_cholesky_factorization_device_z(ctx, uplo, a, d, s);
}
inline char uplo_char(blas::Uplo uplo) {
  return uplo == blas::Uplo::Lower ? 'L' : 'U';
}
}  // namespace internal

// Distributed (reference cholesky.h:71-83). Preconditions as in the reference (square matrix, square
// blocks, matrix distributed on `grid`) are checked by the library and terminate on violation.
template <Backend B, Device D, class T>
void cholesky_factorization(comm::CommunicatorGrid& grid, const blas::Uplo uplo, Matrix<T, D>& mat_a) {
  static_assert(B == Backend::GPU && D == Device::GPU,
                "this build provides Backend::GPU / Device::GPU only (no MC backend, no CPU fallback)");
  static_assert(std::is_same_v<T, float> || std::is_same_v<T, double> || std::is_same_v<T, std::complex<float>> ||
                    std::is_same_v<T, std::complex<double>>,
                "element types: float, double, std::complex<float>, std::complex<double>");
  // equal_process_grid(mat_a, grid) of the reference (cholesky.h:76): the matrix must have been created on this grid
  if (grid.context() != mat_a.context()) {
    std::fprintf(stderr, "[dlaf] cholesky_factorization: the matrix is not distributed on the given communicator grid\n");
    std::abort();
  }
  internal::call_device(mat_a.context(), internal::uplo_char(uplo), mat_a.ptr(), mat_a.descriptor(), mat_a.stream());
}

// Local (reference cholesky.h:41-52): the matrix lives on a 1 x 1 grid.
template <Backend B, Device D, class T>
void cholesky_factorization(const blas::Uplo uplo, Matrix<T, D>& mat_a) {
  static_assert(B == Backend::GPU && D == Device::GPU,
                "this build provides Backend::GPU / Device::GPU only (no MC backend, no CPU fallback)");
  internal::call_device(mat_a.context(), internal::uplo_char(uplo), mat_a.ptr(), mat_a.descriptor(), mat_a.stream());
}

// LAPACK-style info of the last factorization of `mat_a` (0 = success); synchronises the matrix' stream.
template <class T, Device D>
int cholesky_info(Matrix<T, D>& mat_a) {
  return dlaf_b200_wait(mat_a.context(), mat_a.stream());
}

}  // namespace dlaf
