// Host-side helpers of the miniapp (input generation). See util_matrix.cpp.
#pragma once

#include <complex>

namespace dlaf_b200 {

// This rank's local part of an n x n block-cyclic matrix (tile size nb) on a P x Q grid; (vrow, vcol)
// are the rank's coordinates relative to the source rank ((rank - src) mod grid).
template <class T>
struct LocalMatrixView {
  T* data;
  long ld;
  long n;
  int nb;
  int P, Q, vrow, vcol;
};

template <class T>
void set_random_hermitian_positive_definite_local(const LocalMatrixView<T>& m);

}  // namespace dlaf_b200
