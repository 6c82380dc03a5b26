// Common host/device helpers for the B200 POTRF engine.
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace dlaf_b200 {

// Kernel granularity: every padded tile edge (nbp) is a multiple of this, so no kernel on the
// hot path ever sees a ragged tile. Ragged user tiles are padded with an identity block at the
// boundary (see layout.cu); chol([A 0; 0 I]) = [L 0; 0 I] keeps results identical.
constexpr int kGran = 128;

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
  if (e != cudaSuccess) {
    std::fprintf(stderr, "[dlaf_b200] CUDA error %s (%d) at %s:%d: %s\n", cudaGetErrorString(e),
                 static_cast<int>(e), file, line, what);
    std::fflush(stderr);
    std::abort();  // same convention as the reference: failures terminate (src/c_api/utils.cpp:56-69)
  }
}

#define DLAF_CUDA_CHECK(expr) ::dlaf_b200::cuda_check((expr), #expr, __FILE__, __LINE__)

#define DLAF_B200_ASSERT(cond, msg)                                                            \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      std::fprintf(stderr, "[dlaf_b200] assertion failed: %s (%s) at %s:%d\n", #cond, msg,     \
                   __FILE__, __LINE__);                                                        \
      std::fflush(stderr);                                                                     \
      std::abort();                                                                            \
    }                                                                                          \
  } while (0)

inline int ceil_div(long a, long b) {
  return static_cast<int>((a + b - 1) / b);
}
inline long round_up(long a, long b) {
  return (a + b - 1) / b * b;
}

}  // namespace dlaf_b200
