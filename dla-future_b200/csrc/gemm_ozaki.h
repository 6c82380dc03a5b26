// fp64 trailing update on the 5th-generation tensor cores (gemm_ozaki_i8.cu): host-side handles.
//
// tcgen05 has no f64 kind, so the fp64 contraction C -= A B^T is carried by EXACT int8 tensor-core products of a
// fixed-point slicing of the operands (Ozaki scheme): every panel row x is scaled by a power of two, rounded to a 55-bit
// signed integer and cut into S = 7 balanced radix-256 digits,
//     x = 2^e * sum_t d_t 2^(-7-8t) + r,   d_t in [-128, 127] (|d_0| <= 65),   |r| <= 2^(e-56),   2^(e-2) <= max|x| < 2^(e-1)
// (entries within a factor 4 of their row's maximum are represented exactly). The digit-plane products D^A_t (D^B_u)^T
// are exact in the int32 TMEM accumulators (|sum| <= 7 * 2^14 * 512 < 2^26), the 28 pairs with g = t + u <= 6 are kept
// and recombined exactly in the epilogue (one rounding when the result is added to C).
// Error model (per element of the K-term product, rows i of A and j of B):
//     |delta| <= K * 2^-53.2 * max_k|a_ik| * max_k|b_jk|        (digit rounding 2 * 2^-54 + dropped pairs 8 * 2^-57, first order)
// i.e. relative to the ROW MAXIMA, not to sum_k |a_ik b_jk| like a native fp64 dot product (K * 2^-53 * sum_k|a_ik b_jk|):
// for the Cholesky trailing update, where max_k |l_ik| <= sqrt(a_ii), this is the classical scaled backward-error bound
// |dA_ij| <= c eps sqrt(a_ii a_jj). On ordinary data the result is CLOSER to the exact product than a native fp64
// GEMM (no accumulation rounding; tests/ozaki_model.py, tools/gpu_ozaki_test); on rows spanning many binades the
// componentwise error can exceed the native one, therefore the GUARD: the split raises a flag when a nonzero entry
// is rounded and keeps fewer than DLAF_B200_OZAKI_MIN_BITS (default 16) significant bits — i.e. it is more than
// ~2^40 below its row's maximum — and the update of that step then runs on the native fp64 (DMMA) kernel instead.
#pragma once

#include <cuda_runtime.h>

#include "gemm_args.h"

namespace dlaf_b200 {

constexpr int kOzakiSlices = 7;
constexpr int kOzakiPairs = 28;  // digit-plane products (t + u <= 6) = int8 MACs per fp64 MAC

// A panel sliced into 7 int8 planes, stored K-major (plane, row, k contiguous), with per-row power-of-two scales
// and the two TMA tensor maps (A-side box: 128 rows, B-side box: 32 rows; CUtensorMap is opaque here).
struct OzakiSplit {
  signed char* q = nullptr;  // [7][rows][kdim]
  double* scale = nullptr;   // [rows]  2^e of the row
  long rows = 0;
  int kdim = 0;
  alignas(64) unsigned char map_a[128];
  alignas(64) unsigned char map_b[128];

  void allocate(long rows_max, int kdim);
  void release();
  // x: nrows x kdim, column-major (leading dimension ld) -> planes / scales of rows [0, nrows).
  // tile_rows / tile_stride describe tile-contiguous panel workspaces: row r of x lives at
  // x + (r / tile_rows) * tile_stride + r % tile_rows (tile_stride == 0: plain column-major).
  // flag (device int, may be null): OR-ed with 1 when the guard criterion above fires for any entry of these rows.
  // dst_row0: first row of the split buffer to fill (x then points at the source of that row).
  void split(const double* x, long ld, long nrows, cudaStream_t s, int tile_rows = 0, long tile_stride = 0,
             int* flag = nullptr, long dst_row0 = 0);
};

// C = C + alpha * A B^T (alpha = +-2^e: an exact rescaling of the row scales; beta = 1) with A = rows [a_row, a_row + M) of `sa`, B = rows [b_row, b_row + N)
// of `sb`; mask / geometry / C taken from `a` (its A, B pointers are ignored). M % 128 == 0, N % 64 == 0,
// K == kdim <= 512, K % 64 == 0. b_tile_rows: see launch_gemm_tf32x3. guard (device int, may be null): the kernel returns
// without touching C when *guard != 0 (pair it with launch_gemm_nt_f64_if(..., guard) on the same stream).
void launch_gemm_ozaki_i8(const GemmArgsT<double>& a, const OzakiSplit& sa, long a_row, const OzakiSplit& sb, long b_row,
                          cudaStream_t stream, long b_tile_rows = 0, const int* guard = nullptr);
// significant bits an entry must keep before the guard fires (DLAF_B200_OZAKI_MIN_BITS, default 16; 0 = guard off)
int ozaki_min_bits();

// Measurement aid (tools/): device buffer of 4096 x 8 clock64 stamps (first 4096 CTAs of a launch); nullptr = off.
// stamps: 0 entry, 1 set-up done, 2 first stage landed, 3 last MMA issued, 4 accumulators complete, 5 epilogue done, 6 exit
void ozaki_set_clock_trace(long long* dev_buffer);

}  // namespace dlaf_b200
