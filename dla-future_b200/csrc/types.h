// Element-type plumbing: host-side std::complex <-> device-side vector types, per-type kernel
// granularity.
#pragma once

#include <cuda_runtime.h>

#include <complex>

namespace dlaf_b200 {

template <class T>
struct DevType {
  using type = T;
};
template <>
struct DevType<std::complex<float>> {
  using type = float2;
};
template <>
struct DevType<std::complex<double>> {
  using type = double2;
};
template <class T>
using devtype_t = typename DevType<T>::type;

template <class T>
struct BaseOf {
  using type = T;
};
template <class T>
struct BaseOf<std::complex<T>> {
  using type = T;
};
template <>
struct BaseOf<float2> {
  using type = float;
};
template <>
struct BaseOf<double2> {
  using type = double;
};
template <class T>
using base_t = typename BaseOf<T>::type;

// Granularity of every kernel on the path for element type T (device type): padded tile edges are
// multiples of it. Real types: 128 (DMMA CTA tile / diagonal block). Complex: 64 (a 64x64 complex
// block is the same flop volume and shared-memory footprint class as a 128x128 real one).
template <class T>
struct Gran {
  static constexpr int value = 128;
};
template <>
struct Gran<float2> {
  static constexpr int value = 64;
};
template <>
struct Gran<double2> {
  static constexpr int value = 64;
};

__host__ __device__ inline float conj_val(float v) {
  return v;
}
__host__ __device__ inline double conj_val(double v) {
  return v;
}
__host__ __device__ inline float2 conj_val(float2 v) {
  return make_float2(v.x, -v.y);
}
__host__ __device__ inline double2 conj_val(double2 v) {
  return make_double2(v.x, -v.y);
}

__host__ __device__ inline float re_part(float v) {
  return v;
}
__host__ __device__ inline double re_part(double v) {
  return v;
}
__host__ __device__ inline float re_part(float2 v) {
  return v.x;
}
__host__ __device__ inline double re_part(double2 v) {
  return v.x;
}

template <class T>
__host__ __device__ inline T make_real(base_t<T> v);
template <>
__host__ __device__ inline float make_real<float>(float v) {
  return v;
}
template <>
__host__ __device__ inline double make_real<double>(double v) {
  return v;
}
template <>
__host__ __device__ inline float2 make_real<float2>(float v) {
  return make_float2(v, 0.f);
}
template <>
__host__ __device__ inline double2 make_real<double2>(double v) {
  return make_double2(v, 0.0);
}

}  // namespace dlaf_b200
