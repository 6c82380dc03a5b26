// dlaf/types.h — B200 build. Mirrors the part of the reference's include/dlaf/types.h:25-61, :121-162 that
// the Cholesky path exposes (Backend / Device selectors, SizeType, flop accounting).
#pragma once

#include <complex>
#include <cstddef>

namespace dlaf {

using SizeType = std::ptrdiff_t;  // types.h:25

enum class Device { CPU, GPU, Default = GPU };   // types.h:31-43 (built "WITH_GPU")
enum class Backend { MC, GPU, Default = GPU };   // types.h:45-61

template <Backend B>
struct DefaultDevice;
template <>
struct DefaultDevice<Backend::MC> {
  static constexpr Device value = Device::CPU;
};
template <>
struct DefaultDevice<Backend::GPU> {
  static constexpr Device value = Device::GPU;
};
template <Backend B>
inline constexpr Device DefaultDevice_v = DefaultDevice<B>::value;

template <class T>
struct TypeInfo {
  using BaseType = T;
  static constexpr int ops_add = 1, ops_mul = 1;
};
template <class T>
struct TypeInfo<std::complex<T>> {
  using BaseType = T;
  static constexpr int ops_add = 2, ops_mul = 6;  // types.h:121-132
};
template <class T>
using BaseType = typename TypeInfo<T>::BaseType;

// types.h:159-162
template <class T>
constexpr double total_ops(const double add, const double mul) {
  return TypeInfo<T>::ops_add * add + TypeInfo<T>::ops_mul * mul;
}

}  // namespace dlaf

// blaspp's enum used in the reference's signatures (blas::Uplo); taken from <blas.hh> when available.
#if __has_include(<blas.hh>)
#include <blas.hh>
#else
namespace blas {
enum class Uplo : char { Upper = 'U', Lower = 'L', General = 'G' };
}
#endif
