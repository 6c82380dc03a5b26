// See potrf_tile.cuh. Algorithm (validated against numpy in tools/proto_potrf_inv.py):
//
// The 128x128 block lives in shared memory as S (column-major, odd leading dimension so both
// S(r, c) with r running and S(c, r) with r running are bank-conflict free). The lower triangle
// holds A -> L. The strictly upper triangle holds M^T where M converges to inv(L): applying the
// elementary eliminations L_j^-1 to the identity is, element for element, the SAME rank-1 rule as
// the right-looking Cholesky update:
//
//     for column j:   d = sqrt(S(j,j));  S(:,j) /= d  (all rows but j)
//                     for s > j, for r with (r >= s) or (r < j):   S(r,s) -= S(r,j) * S(s,j)
//                     S(j,s) = -S(s,j) / d                                   (new M(s,j))
//
// so one sweep with two barriers per column yields both L and inv(L).
#include "potrf_tile.cuh"

#include "common.h"

namespace dlaf_b200 {

namespace {

constexpr int PB = kPotrfBlock;
constexpr int PLD = PB + 1;
constexpr int kPotrfThreads = 256;
constexpr int kPotrfSmem = (PB * PLD + 2 * PB) * 8;

__global__ void __launch_bounds__(kPotrfThreads, 1)
    potrf128_inv_f64_kernel(double* __restrict__ T, long ldt, double* __restrict__ W, long ldw,
                            int* info, int info_offset) {
  extern __shared__ double S[];
  double* dd = S + PB * PLD;  // L diagonal
  double* dinv = dd + PB;     // 1 / L diagonal = diag(inv(L))
  const int tid = threadIdx.x;

  for (int idx = tid; idx < PB * PB; idx += kPotrfThreads) {
    const int r = idx & (PB - 1), c = idx >> 7;
    S[c * PLD + r] = (r >= c) ? T[r + c * ldt] : 0.0;
  }
  __syncthreads();

  const int r = tid & (PB - 1), half = tid >> 7;
  int fail = 0;
  for (int j = 0; j < PB; ++j) {
    const double ajj = S[j * PLD + j];  // nobody writes S(j,j) during or after step j
    if (!(ajj > 0.0)) {                 // also catches NaN; uniform across the CTA
      fail = j + 1;
      break;
    }
    // Only the reciprocal is on the critical path: rsqrt + one Newton step (correct to ~1 ulp); the
    // diagonal entry itself is the correctly rounded sqrt, computed by its owner thread only.
    double inv_d = rsqrt(ajj);
    inv_d = fma(inv_d * 0.5, fma(-ajj * inv_d, inv_d, 1.0), inv_d);
    if (half == 0) {
      if (r == j) {
        dd[j] = sqrt(ajj);
        dinv[j] = inv_d;
      }
      else {
        S[j * PLD + r] *= inv_d;
      }
    }
    __syncthreads();
    const double* colj = S + j * PLD;
    if (r == j) {
      for (int s = j + 1 + half; s < PB; s += 2)
        S[s * PLD + j] = -colj[s] * inv_d;
    }
    else {
      // rows above the pivot (inverse part) see every remaining column; rows below (Cholesky part)
      // only columns up to their own index.
      const int s_end = (r < j) ? PB : r + 1;
      const double srj = colj[r];
      double* col = S + r;
      int s = j + 1 + half;
      for (; s + 6 < s_end; s += 8) {
        const double b0 = colj[s], b1 = colj[s + 2], b2 = colj[s + 4], b3 = colj[s + 6];
        double v0 = col[s * PLD], v1 = col[(s + 2) * PLD], v2 = col[(s + 4) * PLD],
               v3 = col[(s + 6) * PLD];
        v0 = fma(-srj, b0, v0);
        v1 = fma(-srj, b1, v1);
        v2 = fma(-srj, b2, v2);
        v3 = fma(-srj, b3, v3);
        col[s * PLD] = v0;
        col[(s + 2) * PLD] = v1;
        col[(s + 4) * PLD] = v2;
        col[(s + 6) * PLD] = v3;
      }
      for (; s < s_end; s += 2)
        col[s * PLD] = fma(-srj, colj[s], col[s * PLD]);
    }
    __syncthreads();
  }
  if (fail) {
    if (tid == 0)
      atomicCAS(info, 0, info_offset + fail);
    // Leave T untouched past the failure like a trapped reference run would; W gets zeros so that
    // downstream GEMMs stay finite.
    for (int idx = tid; idx < PB * PB; idx += kPotrfThreads)
      W[(idx & (PB - 1)) + (idx >> 7) * ldw] = 0.0;
    return;
  }

  for (int idx = tid; idx < PB * PB; idx += kPotrfThreads) {
    const int rr = idx & (PB - 1), c = idx >> 7;
    if (rr > c) {
      T[rr + c * ldt] = S[c * PLD + rr];
      W[rr + c * ldw] = S[rr * PLD + c];  // M(rr,c) is stored transposed at S(c,rr)
    }
    else if (rr == c) {
      T[rr + c * ldt] = dd[rr];
      W[rr + c * ldw] = dinv[rr];
    }
    else {
      W[rr + c * ldw] = 0.0;
    }
  }
}

}  // namespace

void launch_potrf128_inv_f64(double* T, long ldt, double* W, long ldw, int* info, int info_offset,
                             cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    DLAF_CUDA_CHECK(cudaFuncSetAttribute(potrf128_inv_f64_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kPotrfSmem));
    configured = true;
  }
  potrf128_inv_f64_kernel<<<1, kPotrfThreads, kPotrfSmem, stream>>>(T, ldt, W, ldw, info,
                                                                    info_offset);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

template <>
void launch_potrf_inv<double>(double* t, long ldt, double* w, long ldw, int* info, int info_offset,
                              cudaStream_t stream) {
  launch_potrf128_inv_f64(t, ldt, w, ldw, info, info_offset, stream);
}

}  // namespace dlaf_b200
