// dlaf/util_matrix.h — the input generator of the miniapp (reference: include/dlaf/util_matrix.h:410-453,
// :529-531): random Hermitian positive definite matrix, values independent of the distribution.
#pragma once

#include <complex>

#include <dlaf/matrix/matrix.h>
#include <dlaf_c/b200_ext.h>

namespace dlaf::matrix::util {

inline void set_random_hermitian_positive_definite(Matrix<float, Device::CPU>& m) {
  dlaf_b200_set_random_hermitian_positive_definite_s(m.context(), m.ptr(), m.descriptor());
}
inline void set_random_hermitian_positive_definite(Matrix<double, Device::CPU>& m) {
  dlaf_b200_set_random_hermitian_positive_definite_d(m.context(), m.ptr(), m.descriptor());
}
inline void set_random_hermitian_positive_definite(Matrix<std::complex<float>, Device::CPU>& m) {
  dlaf_b200_set_random_hermitian_positive_definite_c(m.context(), m.ptr(), m.descriptor());
}
inline void set_random_hermitian_positive_definite(Matrix<std::complex<double>, Device::CPU>& m) {
  dlaf_b200_set_random_hermitian_positive_definite_z(m.context(), m.ptr(), m.descriptor());
}

}  // namespace dlaf::matrix::util
