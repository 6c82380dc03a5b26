// fp64 trailing update on the 5th-generation tensor cores (gemm_ozaki_i8.cu): host-side handles.
//
// tcgen05 has no f64 kind, so the fp64 contraction C -= A B^T is carried by EXACT int8 tensor-core products of an
// error-free slicing of the operands (Ozaki scheme): every panel row x is scaled by a power of two and cut into
// S = 8 signed 7-bit digits,   x = 2^e * sum_t q_t 128^-(t+1) + r,  |q_t| <= 64,  |r| <= 2^(e-57),
// the slice products Q^A_t (Q^B_u)^T are exact in the int32 TMEM accumulators, and the anti-diagonal groups
// g = t + u < S are recombined in fp64 in the epilogue. Measured (tools/proto_ozaki_i8.py, tools/gpu_ozaki_test):
// the result is closer to the exact product than a native fp64 GEMM (no accumulation rounding, truncation
// 2^-55 relative to |row| |col|), so this is NOT a reduced-precision path.
#pragma once

#include <cuda_runtime.h>

#include "gemm_args.h"

namespace dlaf_b200 {

constexpr int kOzakiSlices = 8;

// A panel sliced into 8 int8 planes, stored K-major (plane, row, k contiguous), with per-row power-of-two scales
// and the two TMA tensor maps (A-side box: 128 rows, B-side box: 64 rows; CUtensorMap is opaque here).
struct OzakiSplit {
  signed char* q = nullptr;  // [8][rows][kdim]
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
  void split(const double* x, long ld, long nrows, cudaStream_t s, int tile_rows = 0, long tile_stride = 0);
};

// C = C + alpha * A B^T (alpha = +-1, beta = 1) with A = rows [a_row, a_row + M) of `sa`, B = rows [b_row, b_row + N)
// of `sb`; mask / geometry / C taken from `a` (its A, B pointers are ignored). M % 128 == 0, N % 64 == 0,
// K == kdim, K % 64 == 0. b_tile_rows: see launch_gemm_tf32x3.
void launch_gemm_ozaki_i8(const GemmArgsT<double>& a, const OzakiSplit& sa, long a_row, const OzakiSplit& sb, long b_row,
                          cudaStream_t stream, long b_tile_rows = 0);

// Measurement aid (tools/): device buffer of 4096 x 8 clock64 stamps (first 4096 CTAs of a launch); nullptr = off.
// stamps: 0 entry, 1 set-up done, 2 first stage landed, 3 last MMA issued, 4 accumulators complete, 5 epilogue done, 6 exit
void ozaki_set_clock_trace(long long* dev_buffer);

}  // namespace dlaf_b200
