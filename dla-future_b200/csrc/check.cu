// Result check of the miniapp, on the GPU: max|A - L L^H| / max|A| over the referenced triangle
// (reference: check_cholesky / cholesky_diff, miniapp/miniapp_cholesky.cpp:262-446, which builds L L^H
// tile by tile with blas::gemm on the host and reduces max norms with MPI). Single-rank grids only.
#include <cuda_runtime.h>

#include <complex>
#include <cstring>

#include "common.h"
#include "gemm_args.h"
#include "layout.cuh"
#include "types.h"

namespace dlaf_b200 {

namespace {

inline double __longlong_as_double_host(unsigned long long v) {
  double d;
  static_assert(sizeof(d) == sizeof(v), "size");
  memcpy(&d, &v, sizeof(d));
  return d;
}

__device__ __forceinline__ double abs_of(float v) { return fabsf(v); }
__device__ __forceinline__ double abs_of(double v) { return fabs(v); }
__device__ __forceinline__ double abs_of(float2 v) { return hypotf(v.x, v.y); }
__device__ __forceinline__ double abs_of(double2 v) { return hypot(v.x, v.y); }

// max |x(i,j)| over i >= j of the leading n x n part; result accumulated with an integer atomicMax on the
// bit pattern (monotonic for non-negative doubles)
template <class T>
__global__ void max_abs_lower_kernel(const T* __restrict__ x, long ld, long n, unsigned long long* out) {
  double m = 0.0;
  const long j = blockIdx.x;
  for (long i = j + threadIdx.x; i < n; i += blockDim.x)
    m = fmax(m, abs_of(x[i + j * ld]));
  for (int o = 16; o > 0; o >>= 1)
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ double wm[32];
  if ((threadIdx.x & 31) == 0)
    wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (unsigned w = 1; w < blockDim.x / 32; ++w)
      m = fmax(m, wm[w]);
    atomicMax(out, static_cast<unsigned long long>(__double_as_longlong(m)));
  }
}

}  // namespace

template <class T>
double check_cholesky_single_rank(char uplo, long n, int nb, const T* a_host, long lda, const T* f_host, long ldf) {
  if (n == 0)
    return 0.0;
  constexpr int G = Gran<T>::value;
  const bool upper = (uplo == 'U' || uplo == 'u');
  LayoutParams p;
  p.n = n;
  p.nb = nb;
  p.nbp = static_cast<int>(round_up(nb, G));
  p.nt = ceil_div(n, nb);
  p.P = p.Q = 1;
  p.prow = p.pcol = 0;
  p.ltr = p.ltc = p.nt;
  p.ld = static_cast<long>(p.nt) * p.nbp;
  p.transposed = upper ? 1 : 0;
  const long npad = p.ld;
  const long lds = round_up(n, 2);
  T *stage = nullptr, *az = nullptr, *lz = nullptr;
  unsigned long long* d_max = nullptr;
  DLAF_CUDA_CHECK(cudaMalloc(&stage, sizeof(T) * lds * n));
  DLAF_CUDA_CHECK(cudaMalloc(&az, sizeof(T) * npad * npad));
  DLAF_CUDA_CHECK(cudaMalloc(&lz, sizeof(T) * npad * npad));
  DLAF_CUDA_CHECK(cudaMalloc(&d_max, 2 * sizeof(unsigned long long)));
  DLAF_CUDA_CHECK(cudaMemset(d_max, 0, 2 * sizeof(unsigned long long)));
  cudaStream_t s = nullptr;
  p.ldu = lds;
  // A -> az (referenced triangle only, rest and padding zero)
  DLAF_CUDA_CHECK(cudaMemcpy2DAsync(stage, sizeof(T) * lds, a_host, sizeof(T) * lda, sizeof(T) * n, n, cudaMemcpyHostToDevice, s));
  DLAF_CUDA_CHECK(cudaMemsetAsync(az, 0, sizeof(T) * npad * npad, s));
  launch_to_slab<T>(az, stage, p, s);
  max_abs_lower_kernel<T><<<static_cast<unsigned>(npad), 256, 0, s>>>(az, npad, npad, d_max);
  // factor -> lz (lower triangle incl. diagonal, strictly upper zero: setUpperToZeroForDiagonalTiles)
  DLAF_CUDA_CHECK(cudaMemcpy2DAsync(stage, sizeof(T) * lds, f_host, sizeof(T) * ldf, sizeof(T) * n, n, cudaMemcpyHostToDevice, s));
  DLAF_CUDA_CHECK(cudaMemsetAsync(lz, 0, sizeof(T) * npad * npad, s));
  launch_to_slab<T>(lz, stage, p, s);
  // az <- az - lz lz^H on the lower triangle
  GemmArgsT<T> g{};
  g.A = lz;
  g.lda = npad;
  g.B = lz;
  g.ldb = npad;
  g.C = az;
  g.ldc = npad;
  g.M = g.N = g.K = static_cast<int>(npad);
  g.alpha = -1.0;
  g.beta = 1.0;
  g.mask = kMaskLower;
  g.nbp = p.nbp;
  g.P = g.Q = 1;
  launch_gemm_nt<T>(g, s);
  max_abs_lower_kernel<T><<<static_cast<unsigned>(npad), 256, 0, s>>>(az, npad, npad, d_max + 1);
  unsigned long long h[2];
  DLAF_CUDA_CHECK(cudaMemcpyAsync(h, d_max, sizeof(h), cudaMemcpyDeviceToHost, s));
  DLAF_CUDA_CHECK(cudaStreamSynchronize(s));
  cudaFree(stage);
  cudaFree(az);
  cudaFree(lz);
  cudaFree(d_max);
  const double max_a = __longlong_as_double_host(h[0]);
  const double max_d = __longlong_as_double_host(h[1]);
  return max_d / max_a;
}

#define INST(T) template double check_cholesky_single_rank<T>(char, long, int, const T*, long, const T*, long);
INST(float)
INST(double)
INST(float2)
INST(double2)

}  // namespace dlaf_b200
