// Distributed result check of the miniapp on the GPU grid — PotrfEngine<T>::residual (see engine.h).
//
// Reference: check_cholesky / cholesky_diff (miniapp/miniapp_cholesky.cpp:262-446): the factor's upper part is
// zeroed, A - L L^H is rebuilt tile by tile (one k at a time, column k of L broadcast along rows and, transposed,
// along columns), and max|diff| / max|A| over the referenced triangle is reduced over all ranks.
// Here: the same k loop on the device. Column k of L (tiles on or below the diagonal) is packed, broadcast along
// the process rows, re-broadcast transposed along the process columns, and ONE masked native GEMM per k subtracts
// L(:,k) L(:,k)^H from the local copy of A; the two max norms are reduced with ncclAllReduce(max).
#include <cstring>

#include "comm.h"
#include "common.h"
#include "engine.h"

namespace dlaf_b200 {

namespace {

__device__ __forceinline__ double abs_of(float v) { return fabsf(v); }
__device__ __forceinline__ double abs_of(double v) { return fabs(v); }
__device__ __forceinline__ double abs_of(float2 v) { return hypotf(v.x, v.y); }
__device__ __forceinline__ double abs_of(double2 v) { return hypot(v.x, v.y); }

// max |x(i,j)| over the local elements whose GLOBAL row index >= GLOBAL column index (block-cyclic slab of nbp tiles);
// blockIdx.x = local column. Accumulated with an integer atomicMax on the bit pattern (monotonic for doubles >= 0).
template <class T>
__global__ void max_abs_lower_dist_kernel(const T* __restrict__ x, long ld, int nbp, int ltr, int P, int Q, int prow,
                                          int pcol, unsigned long long* out) {
  const long c = blockIdx.x;
  const long gcol = (static_cast<long>(c / nbp) * Q + pcol) * nbp + c % nbp;
  double m = 0.0;
  const long rows = static_cast<long>(ltr) * nbp;
  for (long r = threadIdx.x; r < rows; r += blockDim.x) {
    const long grow = (static_cast<long>(r / nbp) * P + prow) * nbp + r % nbp;
    if (grow >= gcol)
      m = fmax(m, abs_of(x[r + c * ld]));
  }
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
double PotrfEngine<T>::residual(const T* a_user, long lda, const T* f_user, long ldf, bool transposed,
                                ncclComm_t grid_comm, cudaStream_t s) {
  using NT = NcclType<T>;
  if (nt_ == 0)
    return 0.0;
  const int P = geo_.P, Q = geo_.Q;
  const size_t tsz = static_cast<size_t>(nbp_) * nbp_;
  const bool have = ltr_ > 0 && ltc_ > 0;
  const long ldc = own_ld_;
  T *az = nullptr, *lz = nullptr, *pan = nullptr, *panT = nullptr;
  double* d_max = nullptr;  // {max|A|, max|A - L L^H|}
  DLAF_CUDA_CHECK(cudaMalloc(&d_max, 2 * sizeof(double)));
  DLAF_CUDA_CHECK(cudaMemsetAsync(d_max, 0, 2 * sizeof(double), s));
  DLAF_CUDA_CHECK(cudaMalloc(&pan, sizeof(T) * tsz * (ltr_ > 0 ? ltr_ : 1)));
  if (P > 1)
    DLAF_CUDA_CHECK(cudaMalloc(&panT, sizeof(T) * tsz * (ltc_ > 0 ? ltc_ : 1)));
  if (have) {
    // zero-filled slabs + the referenced triangle only: padding and the unreferenced triangle (diagonal tiles
    // included: setUpperToZeroForDiagonalTiles, miniapp_cholesky.cpp:268-286) stay zero in A and in the factor
    const size_t bytes = sizeof(T) * static_cast<size_t>(ldc) * ltc_ * nbp_;
    DLAF_CUDA_CHECK(cudaMalloc(&az, bytes));
    DLAF_CUDA_CHECK(cudaMalloc(&lz, bytes));
    DLAF_CUDA_CHECK(cudaMemsetAsync(az, 0, bytes, s));
    DLAF_CUDA_CHECK(cudaMemsetAsync(lz, 0, bytes, s));
    LayoutParams p = layout(lda, transposed);
    p.ld = ldc;
    launch_to_slab<T>(az, a_user, p, s);
    p.ldu = ldf;
    launch_to_slab<T>(lz, f_user, p, s);
    max_abs_lower_dist_kernel<T><<<static_cast<unsigned>(ltc_ * nbp_), 256, 0, s>>>(az, ldc, nbp_, ltr_, P, Q, geo_.prow,
                                                                                  geo_.pcol,
                                                                                  reinterpret_cast<unsigned long long*>(d_max));
    DLAF_CUDA_CHECK(cudaGetLastError());
  }
  for (int k = 0; k < nt_; ++k) {
    const int li0 = cnt_rows(k), lj0 = cnt_cols(k);  // first local row / column tile with global index >= k
    const int mt = ltr_ - li0, nc = ltc_ - lj0;
    const int owner_c = k % Q;
    if (mt > 0) {
      if (geo_.pcol == owner_c)
        launch_pack_panel<T>(lz + static_cast<long>(li0) * nbp_ + static_cast<long>(k / Q) * nbp_ * ldc, ldc, pan, nbp_,
                             mt, s);
      if (Q > 1)
        DLAF_NCCL_CHECK(ncclBroadcast(pan, pan, tsz * mt * NT::mult, NT::value, row_comm_rank(owner_c), row_comm_, s));
    }
    if (P > 1 && nc > 0) {
      DLAF_NCCL_CHECK(ncclGroupStart());
      for (int lj = lj0; lj < ltc_; ++lj) {
        const long gj = static_cast<long>(lj) * Q + geo_.pcol;
        const int root_v = static_cast<int>(gj % P);
        T* recv = panT + tsz * (lj - lj0);
        const T* send = recv;
        if (root_v == geo_.prow)
          send = pan + tsz * (gj / P - li0);
        DLAF_NCCL_CHECK(ncclBroadcast(send, recv, tsz * NT::mult, NT::value, col_comm_rank(root_v), col_comm_, s));
      }
      DLAF_NCCL_CHECK(ncclGroupEnd());
    }
    if (mt <= 0 || nc <= 0)
      continue;
    const long gj0 = static_cast<long>(lj0) * Q + geo_.pcol;
    const int ri0 = cnt_rows(gj0);
    const int mrows = (ltr_ - ri0) * nbp_;
    if (mrows <= 0)
      continue;
    GemmArgsT<T> a{};
    a.C = az + static_cast<long>(ri0) * nbp_ + static_cast<long>(lj0) * nbp_ * ldc;
    a.ldc = ldc;
    a.M = mrows;
    a.N = nc * nbp_;
    a.K = nbp_;
    a.alpha = -1.0;
    a.beta = 1.0;
    a.mask = kMaskLower;
    a.nbp = nbp_;
    a.P = P;
    a.Q = Q;
    a.prow = geo_.prow;
    a.pcol = geo_.pcol;
    a.ti0 = ri0;
    a.tj0 = lj0;
    a.A = pan + tsz * (ri0 - li0);
    a.lda = nbp_;
    a.a_ts = static_cast<long>(tsz);
    if (P == 1) {
      a.B = pan + tsz * (gj0 - k);  // every row is local: tile (gj, k) sits at index gj - k
      a.b_ts = static_cast<long>(tsz) * Q;
    }
    else {
      a.B = panT;
      a.b_ts = static_cast<long>(tsz);
    }
    a.ldb = nbp_;
    launch_gemm_nt<T>(a, s);
  }
  if (have) {
    max_abs_lower_dist_kernel<T><<<static_cast<unsigned>(ltc_ * nbp_), 256, 0, s>>>(
        az, ldc, nbp_, ltr_, P, Q, geo_.prow, geo_.pcol, reinterpret_cast<unsigned long long*>(d_max + 1));
    DLAF_CUDA_CHECK(cudaGetLastError());
  }
  if (P * Q > 1 && grid_comm != nullptr)
    DLAF_NCCL_CHECK(ncclAllReduce(d_max, d_max, 2, ncclDouble, ncclMax, grid_comm, s));
  double h[2] = {0.0, 0.0};
  DLAF_CUDA_CHECK(cudaMemcpyAsync(h, d_max, sizeof(h), cudaMemcpyDeviceToHost, s));
  DLAF_CUDA_CHECK(cudaStreamSynchronize(s));
  cudaFree(az);
  cudaFree(lz);
  cudaFree(pan);
  cudaFree(panT);
  cudaFree(d_max);
  return h[0] > 0.0 ? h[1] / h[0] : h[1];
}

#define INST(T)                                                                                                  \
  template double PotrfEngine<T>::residual(const T*, long, const T*, long, bool, ncclComm_t, cudaStream_t);
INST(float)
INST(double)
INST(float2)
INST(double2)

}  // namespace dlaf_b200
