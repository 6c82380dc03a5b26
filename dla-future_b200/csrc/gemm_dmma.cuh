// fp64 tensor-core (DMMA.8x8x4) NT GEMM used by every dense contraction of the POTRF path:
//
//   C(MxN) = beta * C + alpha * A(MxK) * B(NxK)^T        all column-major
//
// It replaces the reference's three BLAS tile calls (include/dlaf/factorization/cholesky/impl.h):
//   * gemmTrailingMatrixTile  (impl.h:82-94,  blas/tile.h:249-261)  A_ij -= A_ik A_jk^H
//   * herkTrailingDiagTile    (impl.h:69-80,  blas/tile.h:293-304)  A_jj -= A_jk A_jk^H (lower only)
//   * trsmPanelTile           (impl.h:55-67,  blas/tile.h:337-349)  A_ik <- A_ik L_kk^-H, as block
//     substitution with pre-inverted 128x128 diagonal blocks (potrf_tile.cu) -> pure GEMMs.
// One launch covers the WHOLE local trailing matrix of a step (not one launch per tile): the
// lower-triangular structure is a tile-level mask evaluated on global (block-cyclic) indices.
//
// sm_100a notes: tcgen05 has no f64 kind, so the fp64 tensor path is the warp-level
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; all wider f64 shapes lower to it on sm_100a).
// Operands are staged global->shared with 16-byte cp.async (LDGSTS) in a 4-stage ring.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "gemm_args.h"

namespace dlaf_b200 {

using GemmArgs = GemmArgsT<double>;

namespace gemm_detail {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;\n" ::);
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

}  // namespace gemm_detail

template <int BM_, int BN_, int BK_, int STAGES_, int MINB_ = 1, int WARPS_M_ = 2, int WARPS_N_ = 4>
struct GemmCfg {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, STAGES = STAGES_, MINB = MINB_;
  static constexpr int THREADS = 32 * WARPS_M_ * WARPS_N_;
  // +4 doubles of padding: the DMMA fragment read (k = lane&3, m = lane>>2) then hits 16 distinct
  // 8-byte banks per half-warp (row stride == 4 mod 16 doubles) -> conflict-free LDS.64.
  static constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
  static constexpr int A_STAGE = BK * LDA_S, B_STAGE = BK * LDB_S;
  static constexpr int LDC_S = BM + 2;  // epilogue staging tile (reuses the operand ring)
  static constexpr int RING_BYTES = STAGES * (A_STAGE + B_STAGE) * 8;
  static constexpr int SMEM_BYTES = RING_BYTES > BN * LDC_S * 8 ? RING_BYTES : BN * LDC_S * 8;
  static constexpr int WARPS_M = WARPS_M_, WARPS_N = WARPS_N_;
  static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;  // warp tile
  static constexpr int FM = WM / 8, FN = WN / 8;              // 8x8 fragments per warp
};

// One CTA tile of the NT product, already resolved to this CTA's operands (the body shared by the plain GEMM
// kernel and the fused panel-TRSM kernel below).
struct GemmTileOp {
  const double* Ag;  // first row of this CTA's A rows (k = 0)
  long lda;
  const double* Bg;  // first row of this CTA's B rows (k = 0)
  long ldb;
  double* Cg;        // C(row0, col0)
  long ldc;
  int KT;            // K / BK
  double alpha, beta;
  int cls;           // 1 = full tile, 2 = straddles the diagonal (element mask on grow0 / gcol0)
  long grow0, gcol0;
};

template <class Cfg>
__device__ __forceinline__ void gemm_nt_f64_tile(const GemmTileOp& t, double* smem) {
  using namespace gemm_detail;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES;
  constexpr int FM = Cfg::FM, FN = Cfg::FN;

  double* As = smem;
  double* Bs = smem + STAGES * Cfg::A_STAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int wm0 = (warp % Cfg::WARPS_M) * Cfg::WM;
  const int wn0 = (warp / Cfg::WARPS_M) * Cfg::WN;

  const double* Ag = t.Ag;
  const double* Bg = t.Bg;
  const long lda = t.lda, ldb = t.ldb, ldc = t.ldc;
  const int KT = t.KT;

  if (t.beta != 0.0) {
    // Pull the C tile towards L2 now; the epilogue reads it ~K/16 stages later.
#pragma unroll
    for (int i = 0; i < (BM / 16) * BN / Cfg::THREADS; ++i) {
      const int l = tid + i * Cfg::THREADS;
      const int col = l / (BM / 16), seg = l % (BM / 16);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(t.Cg + seg * 16 + static_cast<long>(col) * ldc));
    }
  }

  auto load_stage = [&](int slot, int kt) {
    const int k0 = kt * BK;
    double* as = As + slot * Cfg::A_STAGE;
    double* bs = Bs + slot * Cfg::B_STAGE;
#pragma unroll
    for (int i = 0; i < (BK * BM / 2) / Cfg::THREADS; ++i) {
      const int c = tid + i * Cfg::THREADS;
      const int k = c / (BM / 2), m2 = c % (BM / 2);
      cp_async16(as + k * Cfg::LDA_S + 2 * m2, Ag + static_cast<long>(k0 + k) * lda + 2 * m2);
    }
#pragma unroll
    for (int i = 0; i < (BK * BN / 2) / Cfg::THREADS; ++i) {
      const int c = tid + i * Cfg::THREADS;
      const int k = c / (BN / 2), n2 = c % (BN / 2);
      cp_async16(bs + k * Cfg::LDB_S + 2 * n2, Bg + static_cast<long>(k0 + k) * ldb + 2 * n2);
    }
  };

  double acc[FM][FN][2];
#pragma unroll
  for (int i = 0; i < FM; ++i)
#pragma unroll
    for (int j = 0; j < FN; ++j)
      acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT)
      load_stage(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + STAGES - 1;
      if (nk < KT)
        load_stage(nk % STAGES, nk);
      cp_async_commit();
    }
    const double* as = As + (kt % STAGES) * Cfg::A_STAGE + wm0 + g;
    const double* bs = Bs + (kt % STAGES) * Cfg::B_STAGE + wn0 + g;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double af[FM], bf[FN];
#pragma unroll
      for (int i = 0; i < FM; ++i)
        af[i] = as[(kk * 4 + tig) * Cfg::LDA_S + 8 * i];
#pragma unroll
      for (int j = 0; j < FN; ++j)
        bf[j] = bs[(kk * 4 + tig) * Cfg::LDB_S + 8 * j];
#pragma unroll
      for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j)
          dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
  // In-place use (C aliases A, N == BN): every read of this CTA's A rows has landed above, and no
  // other CTA touches these rows, so the epilogue may overwrite them.
  __syncthreads();

  // Epilogue: stage the accumulator tile through the (now idle) operand ring so that C is read and
  // written with 16-byte accesses, 1 KB contiguous per column (coalesced), all loads of a batch in
  // flight before the first store. Cs is column-major with LDC_S == 2 (mod 8): the fragment store
  // (row = g, col = 2*tig+e) then covers 16 distinct 8-byte banks per half-warp.
  constexpr int LDC_S = Cfg::LDC_S;
  double* Cs = smem;
#pragma unroll
  for (int i = 0; i < FM; ++i)
#pragma unroll
    for (int j = 0; j < FN; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        Cs[(wn0 + 8 * j + 2 * tig + e) * LDC_S + wm0 + 8 * i + g] = acc[i][j][e];
  __syncthreads();

  const bool use_beta = (t.beta != 0.0);
  const double alpha = t.alpha, beta = t.beta;
  constexpr int CHUNKS = BM * BN / 2 / Cfg::THREADS;  // 16-byte chunks per thread
  constexpr int BATCH = 8;
  double* Cg = t.Cg;
#pragma unroll 1
  for (int b0 = 0; b0 < CHUNKS; b0 += BATCH) {
    double2 cv[BATCH];
    if (use_beta) {
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const int q = tid + (b0 + b) * Cfg::THREADS;
        const int col = q / (BM / 2), m2 = q % (BM / 2);
        cv[b] = *reinterpret_cast<const double2*>(Cg + 2 * m2 + static_cast<long>(col) * ldc);
      }
    }
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const int q = tid + (b0 + b) * Cfg::THREADS;
      const int col = q / (BM / 2), m2 = q % (BM / 2);
      const double2 a = *reinterpret_cast<const double2*>(Cs + col * LDC_S + 2 * m2);
      double2 v;
      v.x = alpha * a.x;
      v.y = alpha * a.y;
      if (use_beta) {
        v.x += beta * cv[b].x;
        v.y += beta * cv[b].y;
      }
      double* dst = Cg + 2 * m2 + static_cast<long>(col) * ldc;
      if (t.cls == 1) {
        *reinterpret_cast<double2*>(dst) = v;
      }
      else {  // tile straddles the diagonal: element mask, never touch the other triangle
        const long gr = t.grow0 + 2 * m2, gc = t.gcol0 + col;
        if (gr >= gc)
          dst[0] = v.x;
        if (gr + 1 >= gc)
          dst[1] = v.y;
      }
    }
  }
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) gemm_nt_f64_kernel(const GemmArgs p) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK;
  extern __shared__ __align__(16) double smem[];

  const int row0 = blockIdx.x * BM, col0 = blockIdx.y * BN;
  GemmTileOp t;
  t.cls = classify_tile(p, row0, col0, BM, BN, t.grow0, t.gcol0);
  if (t.cls == 0)
    return;

  if (p.dbg_stagger_ns > 0) {
    const unsigned lin = blockIdx.x + blockIdx.y * gridDim.x;
    if (lin >= 148 && lin < 296)
      __nanosleep(p.dbg_stagger_ns);
  }
  t.Ag = p.A + (p.a_ts ? (row0 / p.nbp) * p.a_ts + row0 % p.nbp : row0);
  t.lda = p.lda;
  t.Bg = p.B + (p.b_ts ? (col0 / p.nbp) * p.b_ts + col0 % p.nbp : col0);
  t.ldb = p.ldb;
  t.Cg = p.C + row0 + static_cast<long>(col0) * p.ldc;
  t.ldc = p.ldc;
  t.KT = p.K / BK;
  t.alpha = p.alpha;
  t.beta = p.beta;
  gemm_nt_f64_tile<Cfg>(t, smem);
}

// Guarded flavour of the same GEMM: runs only when *flag != 0, on a persistent-style grid (a few CTAs per SM looping over
// the tiles) so that the usual case — flag clear, the int8-digit kernel did the update — costs one tiny launch.
// This is the native-fp64 fallback of the int8 trailing update (gemm_ozaki.h: guard).
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) gemm_nt_f64_if_kernel(const GemmArgs p, const int* flag) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK;
  extern __shared__ __align__(16) double smem[];
  if (*flag == 0)
    return;
  const int gx = p.M / BM, gy = p.N / BN;
  for (int l = blockIdx.x; l < gx * gy; l += gridDim.x) {
    const int row0 = (l % gx) * BM, col0 = (l / gx) * BN;
    GemmTileOp t;
    t.cls = classify_tile(p, row0, col0, BM, BN, t.grow0, t.gcol0);
    if (t.cls == 0)
      continue;
    t.Ag = p.A + (p.a_ts ? (row0 / p.nbp) * p.a_ts + row0 % p.nbp : row0);
    t.lda = p.lda;
    t.Bg = p.B + (p.b_ts ? (col0 / p.nbp) * p.b_ts + col0 % p.nbp : col0);
    t.ldb = p.ldb;
    t.Cg = p.C + row0 + static_cast<long>(col0) * p.ldc;
    t.ldc = p.ldc;
    t.KT = p.K / BK;
    t.alpha = p.alpha;
    t.beta = p.beta;
    gemm_nt_f64_tile<Cfg>(t, smem);
    __syncthreads();  // the staging ring is reused by the next tile
  }
}

// Panel TRSM in ONE launch:  B <- B * L^-T  for a row panel B (m x ns*G) against the factored diagonal tile
// L (ns*G x ns*G, lower) whose G x G diagonal blocks come pre-inverted (W, from potrf_inv). Rows are
// independent in a right-side triangular solve, so each CTA owns BM rows of B and runs the whole block
// substitution for them, phase after phase, with the tile body above:
//     for j = 0 .. ns-1:   B_j -= sum_{i<j} X_i L_ji^T   (K = j*G;  X_i are this CTA's own finished rows)
//                          X_j  = B_j inv(L_jj)^T         (K = G, in place)
// It replaces 2*ns-1 dependent launches per panel on the critical path (reference: one cublas?trsm per
// tile, include/dlaf/factorization/cholesky/impl.h:55-67, blas/tile.h:337-349). Same shared-memory and
// register footprint as one in-place product, so it slots in next to the bulk update's CTAs.
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) trsm_fused_f64_kernel(const TrsmFusedArgs p) {
  constexpr int G = Cfg::BN;  // one CTA column == one diagonal block
  extern __shared__ __align__(16) double smem[];
  double* rows = p.B + static_cast<long>(blockIdx.x) * Cfg::BM;
  GemmTileOp t;
  t.cls = 1;
  t.grow0 = t.gcol0 = 0;
  t.lda = p.ldb;
  t.ldc = p.ldb;
  for (int j = 0; j < p.ns; ++j) {
    double* bj = rows + static_cast<long>(j) * G * p.ldb;
    if (j > 0) {
      t.Ag = rows;
      t.Bg = p.T + static_cast<long>(j) * G;  // row block j of L, columns [0, j*G)
      t.ldb = p.ldt;
      t.Cg = bj;
      t.KT = j * G / Cfg::BK;
      t.alpha = -1.0;
      t.beta = 1.0;
      gemm_nt_f64_tile<Cfg>(t, smem);
      __threadfence_block();
      __syncthreads();  // B_j (global) and the staging ring are handed to the next phase
    }
    t.Ag = bj;
    t.Bg = p.W + static_cast<long>(j) * G * G;
    t.ldb = G;
    t.Cg = bj;
    t.KT = G / Cfg::BK;
    t.alpha = 1.0;
    t.beta = 0.0;
    gemm_nt_f64_tile<Cfg>(t, smem);
    __threadfence_block();
    __syncthreads();  // X_j is read as an A operand by the following phases
  }
}

// Throughput configuration (bulk trailing update) and two latency configurations for the small
// GEMMs on the critical path (in-tile steps, panel TRSM of the last steps): 2x / 4x more CTAs per
// problem, two CTAs resident per SM.
using GemmCfg128 = GemmCfg<128, 128, 16, 4, 1>;
using GemmCfg64x128 = GemmCfg<64, 128, 16, 4, 2>;
using GemmCfg64 = GemmCfg<64, 64, 16, 4, 2>;
using GemmCfg128k32 = GemmCfg<128, 128, 32, 3, 1>;  // half the barriers per tile
using GemmCfg128x64 = GemmCfg<128, 64, 16, 4, 2>;
// 4-warp CTAs: 4 (3 stages) or 3 (4 stages) independent CTAs per SM -> barriers, prologues and epilogues
// of one CTA are covered by the three others.
using GemmCfg64w4s3 = GemmCfg<64, 64, 16, 3, 4, 2, 2>;
using GemmCfg64w4s4 = GemmCfg<64, 64, 16, 4, 3, 2, 2>;
using GemmCfg32x128w4 = GemmCfg<32, 128, 16, 3, 4, 1, 4>;  // in-place products (one CTA owns all 128 columns)
using GemmCfg16x128w4 = GemmCfg<16, 128, 16, 3, 4, 1, 4>;  // same, half the rows per CTA: twice the CTAs for short panels

// Host launcher (defined in gemm_dmma.cu): picks the tile configuration.
void launch_gemm_nt_f64(const GemmArgs& args, cudaStream_t stream);
// Same product, executed only if *flag != 0 when the kernel starts (flag: device int).
void launch_gemm_nt_f64_if(const GemmArgs& args, const int* flag, cudaStream_t stream);
// Explicit configuration (tools / A-B measurements): 0 = 128x128x16x4, 1 = 64x128 (2 CTA/SM), 2 = 64x64,
// 3 = 128x128x32x3, 4 = 128x64 (2 CTA/SM), 5 = 64x64/4 warps/3 stages (4 CTA/SM), 6 = same/4 stages (3 CTA/SM),
// 7 = 32x128/4 warps (in-place products).
void launch_gemm_nt_f64_cfg(const GemmArgs& args, int cfg, cudaStream_t stream);

}  // namespace dlaf_b200
