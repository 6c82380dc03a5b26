// Type-generic SIMT NT GEMM (C = beta C + alpha A B^H) for the element types that do not have a
// tensor-core kernel yet: float, complex<float>, complex<double>. Same contract, mask and tile-stride
// operands as the fp64 DMMA kernel (gemm_dmma.cuh). CTA tile G x G (G = Gran<T>), 16 x 16 threads,
// each thread a strided (G/16) x (G/16) micro-tile so that shared-memory reads are conflict free / broadcast
// and global accesses of C are coalesced along the rows.
//
// Replaces, for s/c/z: cublas{S,C,Z}gemm / {Ssyrk,Cherk,Zherk} / {S,C,Z}trsm tile calls of the reference
// (include/dlaf/blas/tile.h:249-349).
#include <cstdint>
#include <cstdlib>

#include "common.h"
#include "gemm_args.h"
#include "types.h"

namespace dlaf_b200 {

namespace {

__device__ __forceinline__ float mac_conj(float c, float a, float b) { return fmaf(a, b, c); }
__device__ __forceinline__ float2 mac_conj(float2 c, float2 a, float2 b) {
  return make_float2(fmaf(a.y, b.y, fmaf(a.x, b.x, c.x)), fmaf(-a.x, b.y, fmaf(a.y, b.x, c.y)));
}
__device__ __forceinline__ double2 mac_conj(double2 c, double2 a, double2 b) {
  return make_double2(fma(a.y, b.y, fma(a.x, b.x, c.x)), fma(-a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ float axpby(double alpha, float acc, double beta, float c) {
  return static_cast<float>(alpha) * acc + static_cast<float>(beta) * c;
}
__device__ __forceinline__ float2 axpby(double alpha, float2 acc, double beta, float2 c) {
  const float a = static_cast<float>(alpha), b = static_cast<float>(beta);
  return make_float2(a * acc.x + b * c.x, a * acc.y + b * c.y);
}
__device__ __forceinline__ double2 axpby(double alpha, double2 acc, double beta, double2 c) {
  return make_double2(alpha * acc.x + beta * c.x, alpha * acc.y + beta * c.y);
}

template <class T, int G>
__global__ void __launch_bounds__(256) gemm_nt_simt_kernel(const GemmArgsT<T> p) {
  constexpr int BM = G, BN = G, BK = 8, TM = G / 16, TN = G / 16;
  __shared__ T As[BK][BM];
  __shared__ T Bs[BK][BN];

  const int row0 = blockIdx.x * BM, col0 = blockIdx.y * BN;
  long grow0, gcol0;
  const int cls = classify_tile(p, row0, col0, BM, BN, grow0, gcol0);
  if (cls == 0)
    return;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const T* Ag = p.A + (p.a_ts ? (row0 / p.nbp) * p.a_ts + row0 % p.nbp : row0);
  const T* Bg = p.B + (p.b_ts ? (col0 / p.nbp) * p.b_ts + col0 % p.nbp : col0);

  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j)
      acc[i][j] = make_real<T>(0);

  for (int k0 = 0; k0 < p.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < BK * BM / 256; ++i) {
      const int idx = tid + i * 256;
      const int m = idx % BM, k = idx / BM;
      As[k][m] = Ag[m + static_cast<long>(k0 + k) * p.lda];
    }
#pragma unroll
    for (int i = 0; i < BK * BN / 256; ++i) {
      const int idx = tid + i * 256;
      const int n = idx % BN, k = idx / BN;
      Bs[k][n] = Bg[n + static_cast<long>(k0 + k) * p.ldb];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i)
        a[i] = As[kk][tx + 16 * i];
#pragma unroll
      for (int j = 0; j < TN; ++j)
        b[j] = Bs[kk][ty + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j)
          acc[i][j] = mac_conj(acc[i][j], a[i], b[j]);
    }
    __syncthreads();  // also orders the last operand reads before the (possibly in-place) stores
  }

  const bool use_beta = (p.beta != 0.0);
  T* Cg = p.C + row0 + static_cast<long>(col0) * p.ldc;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int c = ty + 16 * j;
    T cv[TM];
    if (use_beta) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
        cv[i] = Cg[tx + 16 * i + static_cast<long>(c) * p.ldc];
    }
    else {
#pragma unroll
      for (int i = 0; i < TM; ++i)
        cv[i] = make_real<T>(0);
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int r = tx + 16 * i;
      if (cls == 2 && (grow0 + r) < (gcol0 + c))
        continue;
      T v = axpby(p.alpha, acc[i][j], p.beta, cv[i]);
      // the diagonal of a Hermitian update stays real (zherk semantics): nothing to do here because
      // a * conj(a) accumulates an exactly zero imaginary part term by term only in exact arithmetic,
      // so force it for elements on the global diagonal.
      if (cls == 2 && (grow0 + r) == (gcol0 + c) && p.mask == kMaskLower)
        v = make_real<T>(re_part(v));
      Cg[r + static_cast<long>(c) * p.ldc] = v;
    }
  }
}

template <class T>
void launch_simt(const GemmArgsT<T>& a, cudaStream_t stream) {
  constexpr int G = Gran<T>::value;
  if (a.M <= 0 || a.N <= 0)
    return;
  DLAF_B200_ASSERT(a.M % G == 0 && a.N % G == 0 && a.K % 8 == 0 && a.K > 0, "gemm shape must be a multiple of the CTA tile");
  dim3 grid(a.M / G, a.N / G);
  gemm_nt_simt_kernel<T, G><<<grid, 256, 0, stream>>>(a);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

template <>
void launch_gemm_nt<float>(const GemmArgsT<float>& a, cudaStream_t s) {
  launch_simt<float>(a, s);
}
template <>
void launch_gemm_nt<float2>(const GemmArgsT<float2>& a, cudaStream_t s) {
  launch_simt<float2>(a, s);
}
void launch_gemm_nt_z_dmma(const GemmArgsT<double2>& a, cudaStream_t stream);  // gemm_zdmma.cu

template <>
void launch_gemm_nt<double2>(const GemmArgsT<double2>& a, cudaStream_t s) {
  // DLAF_B200_Z_SIMT=1 keeps the first (SIMT) kernel for A/B measurements
  static const bool simt = std::getenv("DLAF_B200_Z_SIMT") != nullptr;
  if (simt)
    launch_simt<double2>(a, s);
  else
    launch_gemm_nt_z_dmma(a, s);
}

}  // namespace dlaf_b200
