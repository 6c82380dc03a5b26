// fp32 trailing update on the 5th-generation tensor cores:  C = beta C + alpha A B^T  with 3xTF32
// error compensation, tcgen05.mma (kind::tf32) fed by TMA, accumulators in TMEM.
//
// Replaces cublasSgemm / cublasSsyrk tile calls of the reference (include/dlaf/blas/tile.h:249-304) for
// the bulk (~95 % of the flops) of SPOTRF. tcgen05 reads fp32 words as TF32 (10-bit mantissa), so every
// panel is first split (split_tf32_kernel) into two exactly-TF32-representable parts
//     x = hi + lo (+ r, |r| <= 2^-22 |x|),     hi = x with the 13 low mantissa bits cleared,
// stored K-major (row = panel row, 32 consecutive k per 128-byte line) so that a TMA box lands in shared
// memory in the canonical SWIZZLE_128B K-major UMMA layout. The product is accumulated as
//     hi*hi  +  (hi*lo + lo*hi)      (three MMAs per k-step, two fp32 TMEM accumulators added in the epilogue),
// at 1/3 of the TF32 rate. Measured accuracy (tools/gpu_kernel_test): between fp32 and 1xTF32 — the
// tensor core truncates when adding into its fp32 accumulator, which dominates the 2^-21 algorithmic error;
// see DESIGN.md for the stated tolerance.
//
// CTA = one 128 x 128 tile of C, 6 warps:  warp 0 = TMA producer (one elected lane),
// warp 1 = TMEM allocator + MMA issuer (one elected lane), warps 2..5 = epilogue (TMEM lane quadrant
// = warp % 4): tcgen05.ld 32 columns at a time, masked read-modify-write of column-major C (lane = row, so
// every column access is a coalesced 128-byte line).
// Pipeline: 3 stages x {A_hi, A_lo, B_hi, B_lo} x (128 rows x 32 k) fp32 = 64 KB per stage, full/empty
// mbarriers; tcgen05.commit releases a stage / signals the epilogue. Every wait is bounded: a pipeline bug
// traps instead of hanging the GPU.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "gemm_args.h"
#include "gemm_tf32.h"
#include "pool.h"

namespace dlaf_b200 {

namespace {

constexpr int TBM = 128, TBN = 128, TBK = 32, TSTAGES = 3;
constexpr int TILE_BYTES = TBM * TBK * 4;         // 16 KB, one operand part of one stage
constexpr int STAGE_BYTES = 4 * TILE_BYTES;       // A_hi, A_lo, B_hi, B_lo
constexpr int TTHREADS = 192;
constexpr int TSMEM_BYTES = TSTAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr uint32_t kTmemCols = 256;  // [0,128): sum hi*hi   [128,256): sum (hi*lo + lo*hi)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done)
      return;
    if (spin > (1u << 26))
      __trap();  // a broken pipeline must not hang the device
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  // UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_128B, 128-byte rows,
  // 8-row groups 1024 bytes apart. start [0,14) (>>4), LBO [16,30) = 1 (unused for swizzled K-major),
  // SBO [32,46) = 1024 >> 4, version [46,48) = 1 (Blackwell), layout [61,64) = 2 (SWIZZLE_128B).
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128 (cute::UMMA::InstrDescriptor bit layout)
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((TBN >> 3) << 17) | ((TBM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kInstrDesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Tf32Params {
  float* C;
  long ldc;
  int K;
  float alpha, beta;
  GemmArgsT<float> g;  // mask / geometry (A, B, C pointers of g are unused here)
  int a_row, b_row;    // row of A(0,:) / B(0,:) inside the split arrays
  int nbp;             // tile edge
  int b_tile_rows;     // rows of the B split array between consecutive tiles (== nbp when contiguous)
};

__global__ void __launch_bounds__(TTHREADS, 1)
    gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                       const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl,
                       const Tf32Params p) {
  const int row0 = blockIdx.x * TBM, col0 = blockIdx.y * TBN;
  long grow0, gcol0;
  const int cls = classify_tile(p.g, row0, col0, TBM, TBN, grow0, gcol0);
  if (cls == 0)
    return;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + TSTAGES * STAGE_BYTES);
  uint64_t* empty = full + TSTAGES;
  uint64_t* tmem_full = empty + TSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.K / TBK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mAh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mAl)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mBh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mBl)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % TSTAGES;
        const uint32_t ph = (kb / TSTAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        uint8_t* st = tiles + s * STAGE_BYTES;
        tma_load_2d(st, &mAh, &full[s], kb * TBK, p.a_row + row0);
        tma_load_2d(st + TILE_BYTES, &mAl, &full[s], kb * TBK, p.a_row + row0);
        const int brow = p.b_row + (col0 / p.nbp) * p.b_tile_rows + col0 % p.nbp;
        tma_load_2d(st + 2 * TILE_BYTES, &mBh, &full[s], kb * TBK, brow);
        tma_load_2d(st + 3 * TILE_BYTES, &mBl, &full[s], kb * TBK, brow);
      }
    }
  }
  else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % TSTAGES;
        const uint32_t ph = (kb / TSTAGES) & 1;
        mbar_wait(&full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem_u32(tiles + s * STAGE_BYTES);
        const uint64_t ah = make_kmajor_sw128_desc(st), al = make_kmajor_sw128_desc(st + TILE_BYTES);
        const uint64_t bh = make_kmajor_sw128_desc(st + 2 * TILE_BYTES), bl = make_kmajor_sw128_desc(st + 3 * TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < TBK / 8; ++ks) {
          const uint64_t adv = static_cast<uint64_t>((ks * 32) >> 4);  // 8 tf32 = 32 bytes along K inside the swizzle atom
          // The tensor core truncates when it adds into the fp32 accumulator, so the small correction terms
          // get their own accumulator (measured: one shared accumulator loses ~3x accuracy) and the two are
          // added in fp32 in the epilogue.
          umma_tf32(tmem_base, ah + adv, bh + adv, (kb | ks) != 0);
          umma_tf32(tmem_base + TBN, ah + adv, bl + adv, (kb | ks) != 0);
          umma_tf32(tmem_base + TBN, al + adv, bh + adv, 1);
        }
        umma_commit(&empty[s]);  // frees the stage once these MMAs have read it
      }
      umma_commit(tmem_full);  // accumulator complete
    }
  }
  else {
    // ===== epilogue: warps 2..5, TMEM lanes 32*(warp%4) .. +31 =====
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of the tile held by this thread
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* Cg = p.C + row0 + r + static_cast<long>(col0) * p.ldc;
    const bool use_beta = (p.beta != 0.f);
#pragma unroll 1
    for (int c0 = 0; c0 < TBN; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr)
          : "memory");
      uint32_t w[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
            "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]),
            "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]),
            "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
          : "r"(taddr + TBN)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float cv[32];
      if (use_beta) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const bool on = (cls == 1) || (grow0 + r >= gcol0 + c0 + j);
          cv[j] = on ? Cg[static_cast<long>(c0 + j) * p.ldc] : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (cls != 1 && (grow0 + r < gcol0 + c0 + j))
          continue;
        float o = p.alpha * (__uint_as_float(v[j]) + __uint_as_float(w[j]));
        if (use_beta)
          o += p.beta * cv[j];
        Cg[static_cast<long>(c0 + j) * p.ldc] = o;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// x (rows x k, column-major with leading dimension ld) -> hi, lo (rows x k, ROW-major: k contiguous)
__global__ void split_tf32_kernel(const float* __restrict__ x, long ld, int rows, int kdim, float* __restrict__ hi,
                                  float* __restrict__ lo, int tile_rows, long tile_stride) {
  __shared__ float t[32][33];
  const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + tx, k = k0 + ty + i;
    const long roff = tile_stride ? (r / tile_rows) * tile_stride + r % tile_rows : r;
    t[ty + i][tx] = (r < rows && k < kdim) ? x[roff + static_cast<long>(k) * ld] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, k = k0 + tx;
    if (r < rows && k < kdim) {
      const float v = t[tx][ty + i];
      const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      const float l = __uint_as_float(__float_as_uint(v - h) & 0xFFFFE000u);
      hi[static_cast<long>(r) * kdim + k] = h;
      lo[static_cast<long>(r) * kdim + k] = l;
    }
  }
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    DLAF_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    DLAF_B200_ASSERT(f != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    return reinterpret_cast<EncodeFn>(f);
  }();
  return fn;
}

}  // namespace

void Tf32Split::allocate(long rows_max, int kdim_) {
  release();
  rows = rows_max;
  kdim = kdim_;
  hi = pool_alloc<float>(rows * kdim);
  lo = pool_alloc<float>(rows * kdim);
  // 2D map: dim0 = k (contiguous), dim1 = row; box = 32 k x 128 rows; 128-byte swizzle
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kdim), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kdim) * 4};
  const cuuint32_t box[2] = {TBK, TBM};
  const cuuint32_t estr[2] = {1, 1};
  for (int i = 0; i < 2; ++i) {
    CUtensorMap* m = (i == 0) ? reinterpret_cast<CUtensorMap*>(map_hi) : reinterpret_cast<CUtensorMap*>(map_lo);
    const CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, i == 0 ? hi : lo, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DLAF_B200_ASSERT(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
  }
}

void Tf32Split::release() {
  pool_free(hi);
  pool_free(lo);
  hi = lo = nullptr;
}

void Tf32Split::split(const float* x, long ld, long nrows, cudaStream_t s, int tile_rows, long tile_stride) {
  DLAF_B200_ASSERT(nrows <= rows, "split buffer too small");
  if (nrows <= 0)
    return;
  dim3 grid(static_cast<unsigned>((nrows + 31) / 32), static_cast<unsigned>((kdim + 31) / 32)), block(32, 8);
  split_tf32_kernel<<<grid, block, 0, s>>>(x, ld, static_cast<int>(nrows), kdim, hi, lo, tile_rows > 0 ? tile_rows : 1,
                                           tile_stride);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

static_assert(sizeof(CUtensorMap) == 128, "tensor map size");

void launch_gemm_tf32x3(const GemmArgsT<float>& a, const Tf32Split& sa, long a_row, const Tf32Split& sb, long b_row,
                        cudaStream_t stream, long b_tile_rows) {
  if (a.M <= 0 || a.N <= 0)
    return;
  DLAF_B200_ASSERT(a.M % TBM == 0 && a.N % TBN == 0 && a.K % TBK == 0 && a.K == sa.kdim && a.K == sb.kdim,
                   "tf32 gemm shape");
  static bool configured = false;
  if (!configured) {
    DLAF_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TSMEM_BYTES));
    configured = true;
  }
  Tf32Params p;
  p.C = a.C;
  p.ldc = a.ldc;
  p.K = a.K;
  p.alpha = static_cast<float>(a.alpha);
  p.beta = static_cast<float>(a.beta);
  p.g = a;
  p.a_row = static_cast<int>(a_row);
  p.b_row = static_cast<int>(b_row);
  p.nbp = a.nbp;
  p.b_tile_rows = static_cast<int>(b_tile_rows > 0 ? b_tile_rows : a.nbp);
  dim3 grid(a.M / TBM, a.N / TBN);
  gemm_tf32x3_kernel<<<grid, TTHREADS, TSMEM_BYTES, stream>>>(
      *reinterpret_cast<const CUtensorMap*>(sa.map_hi), *reinterpret_cast<const CUtensorMap*>(sa.map_lo),
      *reinterpret_cast<const CUtensorMap*>(sb.map_hi), *reinterpret_cast<const CUtensorMap*>(sb.map_lo), p);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

}  // namespace dlaf_b200
