// complex<double> NT GEMM on the fp64 tensor pipe:  C = beta C + alpha A B^H,  all column-major,
// interleaved (re, im) storage. A complex multiply-accumulate with a conjugated B operand is four real
// DMMA.8x8x4:   Cr += Ar Br^T + Ai Bi^T,   Ci += Ai Br^T + Ar (-Bi)^T,
// so each 8x8x4 fragment pair costs 4 DMMAs and the flop rate (8 flop per complex mac) equals the real
// DMMA rate. Replaces cublasZgemm / cublasZherk / cublasZtrsm tile calls (include/dlaf/blas/tile.h:249-349)
// with ONE masked launch per step, like the real kernel (gemm_dmma.cuh), same argument block.
//
// CTA: 64 x 64 complex tile (= Gran<double2>), 4 warps as 2 x 2, warp tile 32 x 32 complex, BK = 16,
// 3-stage 16-byte cp.async ring; two CTAs per SM. Shared rows are padded by 2 complex so the LDS.128
// fragment reads (row = lane>>2, k = lane&3) are conflict free per quarter warp.
#include <cstdint>

#include "common.h"
#include "gemm_args.h"
#include "types.h"

namespace dlaf_b200 {

namespace {

constexpr int ZBM = 64, ZBN = 64, ZBK = 16, ZSTAGES = 3, ZTHREADS = 128;
constexpr int ZLD = ZBM + 2;  // complex elements per k-row of a stage (A and B alike)
constexpr int ZSTAGE = ZBK * ZLD;  // complex elements per operand per stage
constexpr int ZLDC = ZBM + 1;      // epilogue staging tile (complex), odd -> conflict-free 16-byte column access
constexpr int ZRING_BYTES = ZSTAGES * 2 * ZSTAGE * 16;
constexpr int ZSMEM_BYTES = ZRING_BYTES > ZBN * ZLDC * 16 ? ZRING_BYTES : ZBN * ZLDC * 16;

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

// One CTA tile, already resolved to this CTA's operands (shared by the plain kernel and the fused panel TRSM).
struct ZTileOp {
  const double2* Ag;
  long lda;
  const double2* Bg;
  long ldb;
  double2* Cg;
  long ldc;
  int KT;
  double alpha, beta;
  int cls;         // 1 = full tile, 2 = straddles the diagonal
  int herm_diag;   // force a real diagonal (zherk)
  long grow0, gcol0;
};

__device__ __forceinline__ void gemm_nt_z_tile(const ZTileOp& t, double2* zsmem) {
  double2* As = zsmem;
  double2* Bs = zsmem + ZSTAGES * ZSTAGE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int wm0 = (warp & 1) * 32, wn0 = (warp >> 1) * 32;

  const double2* Ag = t.Ag;
  const double2* Bg = t.Bg;
  const long lda = t.lda, ldb = t.ldb, ldc = t.ldc;
  const int KT = t.KT;

  auto load_stage = [&](int slot, int kt) {
    const int k0 = kt * ZBK;
    double2* as = As + slot * ZSTAGE;
    double2* bs = Bs + slot * ZSTAGE;
#pragma unroll
    for (int i = 0; i < ZBK * ZBM / ZTHREADS; ++i) {
      const int c = tid + i * ZTHREADS;
      const int k = c / ZBM, m = c % ZBM;
      cp_async16(as + k * ZLD + m, Ag + static_cast<long>(k0 + k) * lda + m);
    }
#pragma unroll
    for (int i = 0; i < ZBK * ZBN / ZTHREADS; ++i) {
      const int c = tid + i * ZTHREADS;
      const int k = c / ZBN, n = c % ZBN;
      cp_async16(bs + k * ZLD + n, Bg + static_cast<long>(k0 + k) * ldb + n);
    }
  };

  // accumulators: [frag row][frag col][c0,c1] for the real and the imaginary part
  double cr[4][4][2], ci[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < ZSTAGES - 1; ++s) {
    if (s < KT)
      load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<ZSTAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + ZSTAGES - 1;
      if (nk < KT)
        load_stage(nk % ZSTAGES, nk);
      cp_async_commit();
    }
    const double2* as = As + (kt % ZSTAGES) * ZSTAGE + wm0 + g;
    const double2* bs = Bs + (kt % ZSTAGES) * ZSTAGE + wn0 + g;
#pragma unroll
    for (int kk = 0; kk < ZBK / 4; ++kk) {
      double2 af[4], bf[4];
      double nbi[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        af[i] = as[(kk * 4 + tig) * ZLD + 8 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bf[j] = bs[(kk * 4 + tig) * ZLD + 8 * j];
        nbi[j] = -bf[j].y;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dmma884(cr[i][j][0], cr[i][j][1], af[i].x, bf[j].x);
          dmma884(cr[i][j][0], cr[i][j][1], af[i].y, bf[j].y);
          dmma884(ci[i][j][0], ci[i][j][1], af[i].y, bf[j].x);
          dmma884(ci[i][j][0], ci[i][j][1], af[i].x, nbi[j]);
        }
    }
  }
  cp_async_wait<0>();
  __syncthreads();  // operand ring is dead from here on (also covers the in-place case, see gemm_dmma.cuh)

  double2* Cs = zsmem;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        Cs[(wn0 + 8 * j + 2 * tig + e) * ZLDC + wm0 + 8 * i + g] = make_double2(cr[i][j][e], ci[i][j][e]);
  __syncthreads();

  const bool use_beta = (t.beta != 0.0);
  const double alpha = t.alpha, beta = t.beta;
  const int cls = t.cls;
  const long grow0 = t.grow0, gcol0 = t.gcol0;
  double2* Cg = t.Cg;
  constexpr int CHUNKS = ZBM * ZBN / ZTHREADS;  // complex elements per thread
  constexpr int BATCH = 8;
#pragma unroll 1
  for (int b0 = 0; b0 < CHUNKS; b0 += BATCH) {
    double2 cv[BATCH];
    if (use_beta) {
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const int q = tid + (b0 + b) * ZTHREADS;
        cv[b] = Cg[(q % ZBM) + static_cast<long>(q / ZBM) * ldc];
      }
    }
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const int q = tid + (b0 + b) * ZTHREADS;
      const int r = q % ZBM, c = q / ZBM;
      if (cls == 2 && (grow0 + r) < (gcol0 + c))
        continue;
      const double2 a = Cs[c * ZLDC + r];
      double2 v = make_double2(alpha * a.x, alpha * a.y);
      if (use_beta) {
        v.x += beta * cv[b].x;
        v.y += beta * cv[b].y;
      }
      if (cls == 2 && t.herm_diag && (grow0 + r) == (gcol0 + c))
        v.y = 0.0;  // the diagonal of a Hermitian update is real (zherk)
      Cg[r + static_cast<long>(c) * ldc] = v;
    }
  }
}

__global__ void __launch_bounds__(ZTHREADS, 2) gemm_nt_z_kernel(const GemmArgsT<double2> p) {
  extern __shared__ __align__(16) double2 zsmem[];
  const int row0 = blockIdx.x * ZBM, col0 = blockIdx.y * ZBN;
  ZTileOp t;
  t.cls = classify_tile(p, row0, col0, ZBM, ZBN, t.grow0, t.gcol0);
  if (t.cls == 0)
    return;
  t.herm_diag = (p.mask == kMaskLower);
  t.Ag = p.A + (p.a_ts ? (row0 / p.nbp) * p.a_ts + row0 % p.nbp : row0);
  t.lda = p.lda;
  t.Bg = p.B + (p.b_ts ? (col0 / p.nbp) * p.b_ts + col0 % p.nbp : col0);
  t.ldb = p.ldb;
  t.Cg = p.C + row0 + static_cast<long>(col0) * p.ldc;
  t.ldc = p.ldc;
  t.KT = p.K / ZBK;
  t.alpha = p.alpha;
  t.beta = p.beta;
  gemm_nt_z_tile(t, zsmem);
}

// Panel TRSM in one launch (complex<double>), see trsm_fused_f64_kernel in gemm_dmma.cuh: each CTA owns ZBM rows
// of the row panel B (m x ns*64) and runs the whole block substitution against the factored diagonal tile T
// (lower) and its ns pre-inverted 64 x 64 diagonal blocks W:   X_j = (B_j - sum_{i<j} X_i L_ji^H) inv(L_jj)^H.
struct ZTrsmFusedArgs {
  double2* B;
  long ldb;
  const double2* T;
  long ldt;
  const double2* W;
  int ns;
};

__global__ void __launch_bounds__(ZTHREADS, 2) trsm_fused_z_kernel(const ZTrsmFusedArgs p) {
  constexpr int G = ZBN;
  extern __shared__ __align__(16) double2 zsmem[];
  double2* rows = p.B + static_cast<long>(blockIdx.x) * ZBM;
  ZTileOp t;
  t.cls = 1;
  t.herm_diag = 0;
  t.grow0 = t.gcol0 = 0;
  t.lda = p.ldb;
  t.ldc = p.ldb;
  for (int j = 0; j < p.ns; ++j) {
    double2* bj = rows + static_cast<long>(j) * G * p.ldb;
    if (j > 0) {
      t.Ag = rows;
      t.Bg = p.T + static_cast<long>(j) * G;  // row block j of L, columns [0, j*G)
      t.ldb = p.ldt;
      t.Cg = bj;
      t.KT = j * G / ZBK;
      t.alpha = -1.0;
      t.beta = 1.0;
      gemm_nt_z_tile(t, zsmem);
      __threadfence_block();
      __syncthreads();
    }
    t.Ag = bj;
    t.Bg = p.W + static_cast<long>(j) * G * G;
    t.ldb = G;
    t.Cg = bj;
    t.KT = G / ZBK;
    t.alpha = 1.0;
    t.beta = 0.0;
    gemm_nt_z_tile(t, zsmem);
    __threadfence_block();
    __syncthreads();
  }
}

}  // namespace

void launch_gemm_nt_z_dmma(const GemmArgsT<double2>& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0)
    return;
  DLAF_B200_ASSERT(a.M % ZBM == 0 && a.N % ZBN == 0 && a.K % ZBK == 0 && a.K > 0, "gemm shape must be a multiple of the CTA tile");
  const bool in_place = (static_cast<const void*>(a.A) == static_cast<const void*>(a.C));
  DLAF_B200_ASSERT(!in_place || a.N == ZBN, "in-place product needs one CTA column");
  static bool configured = false;
  if (!configured) {
    DLAF_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ZSMEM_BYTES));
    configured = true;
  }
  dim3 grid(a.M / ZBM, a.N / ZBN);
  gemm_nt_z_kernel<<<grid, ZTHREADS, ZSMEM_BYTES, stream>>>(a);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

}  // namespace dlaf_b200

namespace dlaf_b200 {
void launch_trsm_fused_z(double2* b, long ldb, int m, const double2* t, long ldt, const double2* w, int ns,
                         cudaStream_t stream) {
  if (m <= 0 || ns <= 0)
    return;
  DLAF_B200_ASSERT(m % ZBM == 0, "fused TRSM: rows must be a multiple of the CTA row block");
  static bool configured = false;
  if (!configured) {
    DLAF_CUDA_CHECK(cudaFuncSetAttribute(trsm_fused_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ZSMEM_BYTES));
    configured = true;
  }
  const ZTrsmFusedArgs a{b, ldb, t, ldt, w, ns};
  trsm_fused_z_kernel<<<m / ZBM, ZTHREADS, ZSMEM_BYTES, stream>>>(a);
  DLAF_CUDA_CHECK(cudaGetLastError());
}
}  // namespace dlaf_b200
