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

__global__ void __launch_bounds__(ZTHREADS, 2) gemm_nt_z_kernel(const GemmArgsT<double2> p) {
  extern __shared__ __align__(16) double2 zsmem[];
  double2* As = zsmem;
  double2* Bs = zsmem + ZSTAGES * ZSTAGE;

  const int row0 = blockIdx.x * ZBM, col0 = blockIdx.y * ZBN;
  long grow0, gcol0;
  const int cls = classify_tile(p, row0, col0, ZBM, ZBN, grow0, gcol0);
  if (cls == 0)
    return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int wm0 = (warp & 1) * 32, wn0 = (warp >> 1) * 32;

  const double2* Ag = p.A + (p.a_ts ? (row0 / p.nbp) * p.a_ts + row0 % p.nbp : row0);
  const double2* Bg = p.B + (p.b_ts ? (col0 / p.nbp) * p.b_ts + col0 % p.nbp : col0);
  const int KT = p.K / ZBK;

  auto load_stage = [&](int slot, int kt) {
    const int k0 = kt * ZBK;
    double2* as = As + slot * ZSTAGE;
    double2* bs = Bs + slot * ZSTAGE;
#pragma unroll
    for (int i = 0; i < ZBK * ZBM / ZTHREADS; ++i) {
      const int c = tid + i * ZTHREADS;
      const int k = c / ZBM, m = c % ZBM;
      cp_async16(as + k * ZLD + m, Ag + static_cast<long>(k0 + k) * p.lda + m);
    }
#pragma unroll
    for (int i = 0; i < ZBK * ZBN / ZTHREADS; ++i) {
      const int c = tid + i * ZTHREADS;
      const int k = c / ZBN, n = c % ZBN;
      cp_async16(bs + k * ZLD + n, Bg + static_cast<long>(k0 + k) * p.ldb + n);
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

  const bool use_beta = (p.beta != 0.0);
  double2* Cg = p.C + row0 + static_cast<long>(col0) * p.ldc;
  constexpr int CHUNKS = ZBM * ZBN / ZTHREADS;  // complex elements per thread
  constexpr int BATCH = 8;
#pragma unroll 1
  for (int b0 = 0; b0 < CHUNKS; b0 += BATCH) {
    double2 cv[BATCH];
    if (use_beta) {
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const int q = tid + (b0 + b) * ZTHREADS;
        cv[b] = Cg[(q % ZBM) + static_cast<long>(q / ZBM) * p.ldc];
      }
    }
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const int q = tid + (b0 + b) * ZTHREADS;
      const int r = q % ZBM, c = q / ZBM;
      if (cls == 2 && (grow0 + r) < (gcol0 + c))
        continue;
      const double2 a = Cs[c * ZLDC + r];
      double2 v = make_double2(p.alpha * a.x, p.alpha * a.y);
      if (use_beta) {
        v.x += p.beta * cv[b].x;
        v.y += p.beta * cv[b].y;
      }
      if (cls == 2 && p.mask == kMaskLower && (grow0 + r) == (gcol0 + c))
        v.y = 0.0;  // the diagonal of a Hermitian update is real (zherk)
      Cg[r + static_cast<long>(c) * p.ldc] = v;
    }
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
