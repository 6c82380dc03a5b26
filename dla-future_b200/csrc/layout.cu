// See layout.cuh.
#include "layout.cuh"

#include "common.h"
#include "types.h"

namespace dlaf_b200 {

namespace {

constexpr int TS = 32;  // sub-block edge handled by one CTA (32 x 8 threads)

// blockIdx.x = engine local tile (li + lj * ltr); blockIdx.y = 32x32 sub-block of the tile.
template <class T, bool TO_SLAB>
__global__ void __launch_bounds__(TS * 8) convert_kernel(T* __restrict__ slab, T* __restrict__ user,
                                                         const LayoutParams p) {
  __shared__ T tile[TS][TS + 1];
  const int li = blockIdx.x % p.ltr, lj = blockIdx.x / p.ltr;
  const long gi = static_cast<long>(li) * p.P + p.prow, gj = static_cast<long>(lj) * p.Q + p.pcol;
  if (gi < gj || gi >= p.nt || gj >= p.nt)
    return;
  const int nsb = (p.nb + TS - 1) / TS;
  const int bi = blockIdx.y % nsb, bj = blockIdx.y / nsb;
  if (gi == gj && bi < bj)
    return;
  const int rows = static_cast<int>(min(static_cast<long>(p.nb), p.n - gi * p.nb));
  const int cols = static_cast<int>(min(static_cast<long>(p.nb), p.n - gj * p.nb));
  const int tx = threadIdx.x, ty = threadIdx.y;
  const bool diag = (gi == gj);

  T* s_tile = slab + static_cast<long>(li) * p.nbp + static_cast<long>(lj) * p.nbp * p.ld;
  if (!p.transposed) {
    T* u_tile = user + static_cast<long>(li) * p.nb + static_cast<long>(lj) * p.nb * p.ldu;
    const int r = bi * TS + tx;
#pragma unroll
    for (int k = 0; k < TS; k += 8) {
      const int c = bj * TS + ty + k;
      if (r < rows && c < cols && (!diag || r >= c)) {
        if (TO_SLAB)
          s_tile[r + static_cast<long>(c) * p.ld] = u_tile[r + static_cast<long>(c) * p.ldu];
        else
          u_tile[r + static_cast<long>(c) * p.ldu] = s_tile[r + static_cast<long>(c) * p.ld];
      }
    }
  }
  else {
    // engine tile (li, lj) element (r, c)  <->  conj(user tile (lj, li) element (c, r)); the user's
    // local ROW tile index is the engine's local COLUMN tile index.
    T* u_tile = user + static_cast<long>(lj) * p.nb + static_cast<long>(li) * p.nb * p.ldu;
    if (TO_SLAB) {
      // read user coalesced along c (its row index), write slab coalesced along r
      const int c = bj * TS + tx;
#pragma unroll
      for (int k = 0; k < TS; k += 8) {
        const int r = bi * TS + ty + k;
        if (r < rows && c < cols)
          tile[ty + k][tx] = u_tile[c + static_cast<long>(r) * p.ldu];
      }
      __syncthreads();
      const int r2 = bi * TS + tx;
#pragma unroll
      for (int k = 0; k < TS; k += 8) {
        const int c2 = bj * TS + ty + k;
        if (r2 < rows && c2 < cols && (!diag || r2 >= c2))
          s_tile[r2 + static_cast<long>(c2) * p.ld] = conj_val(tile[tx][ty + k]);
      }
    }
    else {
      const int r = bi * TS + tx;
#pragma unroll
      for (int k = 0; k < TS; k += 8) {
        const int c = bj * TS + ty + k;
        if (r < rows && c < cols)
          tile[ty + k][tx] = s_tile[r + static_cast<long>(c) * p.ld];
      }
      __syncthreads();
      const int c2 = bj * TS + tx;
#pragma unroll
      for (int k = 0; k < TS; k += 8) {
        const int r2 = bi * TS + ty + k;
        if (r2 < rows && c2 < cols && (!diag || r2 >= c2))
          u_tile[c2 + static_cast<long>(r2) * p.ldu] = conj_val(tile[tx][ty + k]);
      }
    }
  }
}

template <class T>
__global__ void pad_identity_kernel(T* __restrict__ slab, const LayoutParams p) {
  const int li = blockIdx.x;
  const long gi = static_cast<long>(li) * p.P + p.prow;
  if (gi >= p.nt || (gi - p.pcol) % p.Q != 0 || gi < p.pcol)
    return;
  const int lj = static_cast<int>((gi - p.pcol) / p.Q);
  const int valid = static_cast<int>(min(static_cast<long>(p.nb), p.n - gi * p.nb));
  T* t = slab + static_cast<long>(li) * p.nbp + static_cast<long>(lj) * p.nbp * p.ld;
  for (int i = valid + threadIdx.x; i < p.nbp; i += blockDim.x)
    t[i + static_cast<long>(i) * p.ld] = make_real<T>(1);
}

template <class T>
__global__ void pack_panel_kernel(const T* __restrict__ src, long ld, T* __restrict__ dst, int nbp) {
  // blockIdx.x = tile, blockIdx.y = column of the tile
  const T* s = src + static_cast<long>(blockIdx.x) * nbp + static_cast<long>(blockIdx.y) * ld;
  T* d = dst + static_cast<long>(blockIdx.x) * nbp * nbp + static_cast<long>(blockIdx.y) * nbp;
  for (int r = threadIdx.x; r < nbp; r += blockDim.x)
    d[r] = s[r];
}

}  // namespace

template <class T>
void launch_to_slab(T* slab, const T* user, const LayoutParams& p, cudaStream_t s) {
  if (p.ltr == 0 || p.ltc == 0)
    return;
  const int nsb = (p.nb + TS - 1) / TS;
  dim3 grid(p.ltr * p.ltc, nsb * nsb), block(TS, 8);
  convert_kernel<T, true><<<grid, block, 0, s>>>(slab, const_cast<T*>(user), p);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

template <class T>
void launch_from_slab(const T* slab, T* user, const LayoutParams& p, cudaStream_t s) {
  if (p.ltr == 0 || p.ltc == 0)
    return;
  const int nsb = (p.nb + TS - 1) / TS;
  dim3 grid(p.ltr * p.ltc, nsb * nsb), block(TS, 8);
  convert_kernel<T, false><<<grid, block, 0, s>>>(const_cast<T*>(slab), user, p);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

template <class T>
void launch_pad_identity(T* slab, const LayoutParams& p, cudaStream_t s) {
  if (p.ltr == 0 || p.ltc == 0)
    return;
  DLAF_CUDA_CHECK(cudaMemsetAsync(slab, 0, sizeof(T) * static_cast<size_t>(p.ld) * p.ltc * p.nbp, s));
  pad_identity_kernel<T><<<p.ltr, 128, 0, s>>>(slab, p);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

template <class T>
void launch_pack_panel(const T* src, long ld, T* dst, int nbp, int ntiles, cudaStream_t s) {
  if (ntiles <= 0)
    return;
  dim3 grid(ntiles, nbp);
  pack_panel_kernel<T><<<grid, 128, 0, s>>>(src, ld, dst, nbp);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

#define INSTANTIATE(T)                                                                         \
  template void launch_to_slab<T>(T*, const T*, const LayoutParams&, cudaStream_t);            \
  template void launch_from_slab<T>(const T*, T*, const LayoutParams&, cudaStream_t);          \
  template void launch_pad_identity<T>(T*, const LayoutParams&, cudaStream_t);                 \
  template void launch_pack_panel<T>(const T*, long, T*, int, int, cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(double)
INSTANTIATE(float2)
INSTANTIATE(double2)

}  // namespace dlaf_b200
