// fp64 trailing update on tcgen05 via exact int8 digit products (Ozaki scheme), see gemm_ozaki.h.
//
//   C(MxN, fp64, column-major) += alpha * A(MxK) B(NxK)^T          alpha = +-2^e, lower-triangular tile mask
//
// Replaces the cublasDgemm / cublasDsyrk tile calls of the reference's trailing update
// (include/dlaf/factorization/cholesky/impl.h:69-94, include/dlaf/blas/tile.h:249-304) for the bulk (~95 % of the
// flops) of DPOTRF, above the 37 TFLOP/s DMMA/DFMA roofline of B200: 28 int8 MMAs (digit pairs t + u <= 6) per fp64
// product at the int8 tensor rate.
//
// Operands: OzakiSplit (split_i8_kernel below): 7 balanced radix-256 digit planes per panel (int8, K-major) + one
// power-of-two scale per row. Round 2 layout (round 1: 8 radix-128 digits, 36 pairs):
//   * CTA tile 128 x BN of C. BN = 64 (default): the 7 anti-diagonal group accumulators take 7 x 64 = 448 of the 512 TMEM
//     columns. BN = 32 (DLAF_B200_OZAKI_BN=32): 224 columns, TWO accumulator sets, the epilogue of tile i overlaps the
//     MMAs of tile i + 1 — measured slower, because the MMA phase is bound by shared-memory operand reads (~98 B/clk)
//     and narrow tiles re-read the A planes twice as often per MAC (see ozaki_bn() below);
//   * digit plane t of A meets planes u = 0 .. 6-t of B, whose accumulators are adjacent in TMEM and whose smem
//     planes are adjacent rows -> ONE tcgen05.mma.kind::i8 with N = BN (7 - t), split at the N <= 256 limit: 10 (BN = 64) or 7 (BN = 32) MMAs per k-step;
//   * 18 warps: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane), warps 2..17 =
//     epilogue (TMEM lane quadrant = warp % 4, BN / 4 columns each);
//   * 2 (BN = 64: 84 KB per stage) or 3 (BN = 32: 70 KB) stages x {7 A planes (128 rows x 64 k), 7 B planes (BN rows x 64 k)}, two 3-D TMA boxes
//     (k, row, plane) per stage in SWIZZLE_64B K-major UMMA layout; full/empty mbarriers per stage, tmem_full /
//     tmem_empty per accumulator set; every mbarrier wait is bounded (trap instead of hang);
//   * a CTA handles a few consecutive tiles (DLAF_B200_OZAKI_TPC) so that it stays short-lived next to the
//     high-priority panel-chain kernels.
// Epilogue: per row (= TMEM lane) the 7 int32 group sums are folded EXACTLY into two integers (< 2^48, < 2^43),
// converted to fp64 without I2F, combined with one rounding, scaled by 2^(e_row + e_col - 38) and added to the C values
// that were fetched while the MMAs ran; coalesced column accesses; the masked variant is chosen per warp.
// Guard: split_i8_kernel raises a per-step flag when a nonzero entry would keep fewer than `min_bits` significant bits
// (a row of the panel spans more than ~40 binades); this kernel then returns at once and the guarded native kernel
// (gemm_dmma.cu: launch_gemm_nt_f64_if) performs the update in fp64 DMMA instead.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "gemm_args.h"
#include "gemm_ozaki.h"
#include "pool.h"

namespace dlaf_b200 {

namespace {

constexpr int S = kOzakiSlices;
constexpr int OBM = 128, OBK = 64 /* int8 k per stage */;
constexpr int A_PLANE_BYTES = OBM * OBK;  // 8 KB
constexpr int A_STAGE_BYTES = S * A_PLANE_BYTES;
constexpr int OEPI_WARPS = 16;
constexpr int OTHREADS = 64 + 32 * OEPI_WARPS;  // TMA warp, MMA warp, 16 epilogue warps
constexpr int kRowChunk = 64;  // row tiles per rasterization chunk
constexpr uint32_t kTmemCols = 512;

template <int BN>
struct OzCfg {
  static constexpr int B_PLANE_BYTES = BN * OBK;
  static constexpr int B_STAGE_BYTES = S * B_PLANE_BYTES;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (3 * STAGE_BYTES + 3072 <= 227 * 1024) ? 3 : 2;
  static constexpr int SET_COLS = (S * BN <= 256) ? 256 : 512;  // TMEM columns per accumulator set
  static constexpr int NSETS = 512 / SET_COLS;
  static constexpr int PLANES_PER_MMA = (256 / BN) < S ? (256 / BN) : S;  // N <= 256 per instruction
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 2048 /*barriers + column scales*/;
  static constexpr int HC = BN / (OEPI_WARPS / 4);  // columns per epilogue thread
};

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
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ uint64_t make_kmajor_sw64_desc(uint32_t smem_addr) {
  // UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_64B, 64-byte rows, 8-row groups
  // 512 bytes apart. start [0,14) (>>4), LBO [16,30) = 1 (unused for swizzled K-major), SBO [32,46) = 512 >> 4,
  // version [46,48) = 1 (Blackwell), layout [61,64) = 4 (SWIZZLE_64B).
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}
// kind::i8: D = S32 (c_format 2), A = B = signed 8 bit (format 1), both K-major, M = 128, N = n
// (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t instr_desc_i8(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(OBM >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Measurement aid: when non-null, CTAs with linear index < kDbgCtas record clock64() at 7 phase boundaries.
__device__ long long* g_ozaki_clock_trace = nullptr;
constexpr int kDbgCtas = 4096;
#define OZ_TRACE(slot)                                                        \
  do {                                                                        \
    if (trace)                                                                \
      trace[static_cast<long>(lin) * 8 + (slot)] = clock64();                 \
  } while (0)

struct OzakiParams {
  double* C;
  long ldc;
  int K;         // int8 k per row (= kdim)
  double alpha;  // +-2^e (applied with the power-of-two row scale: exact)
  GemmArgsT<double> g;  // mask / geometry (A, B, C pointers of g are unused here)
  int a_row, b_row;     // row of A(0,:) / B(0,:) inside the split arrays
  int nbp;              // tile edge
  int b_tile_rows;      // rows of the B split array between consecutive tiles (== nbp when contiguous)
  const double* scale_a;
  const double* scale_b;
  int gx, gy;          // tile grid (M / 128, N / BN)
  int tiles_per_cta;
  const int* guard;    // non-null: return at once when *guard != 0 (the native fp64 kernel takes the step)
};

// signed 64-bit integer (|v| < 2^51) -> double, exactly, on the integer + fp64-add pipes (no I2F)
__device__ __forceinline__ double i64_to_f64_exact(long long v) {
  return __longlong_as_double(0x4330000000000000LL + (v + (1LL << 51))) - 0x1.8p52;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

// Fold + store of one thread's row segment (HC columns) of a finished tile. FULL = the tile lies entirely in the
// lower triangle (no element mask, branch-free); otherwise `lim` = number of leading columns of the segment that are
// on or below the diagonal for this row.
template <bool FULL, int HC, int BN>
__device__ __forceinline__ void ozaki_fold_store(uint32_t taddr, double* Cg, long ldc, const double (&cv)[HC],
                                                 double row_scale, const double* cs, int lim) {
#pragma unroll
  for (int c0 = 0; c0 < HC; c0 += 8) {
    int acc[S][8];
#pragma unroll
    for (int g = 0; g < S; ++g)
      tmem_ld8(taddr + static_cast<uint32_t>(g * BN + c0), acc[g]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // sum_g acc_g 256^-g = (hi + lo 2^-24) 2^-24 with two exact integers (|hi| < 2^48, |lo| < 2^43)
      const long long hi = (static_cast<long long>(acc[0][j]) << 24) + (static_cast<long long>(acc[1][j]) << 16) +
                           (static_cast<long long>(acc[2][j]) << 8) + static_cast<long long>(acc[3][j]);
      const long long lo = (static_cast<long long>(acc[4][j]) << 16) + (static_cast<long long>(acc[5][j]) << 8) +
                           static_cast<long long>(acc[6][j]);
      const double v = fma(i64_to_f64_exact(lo), 0x1p-24, i64_to_f64_exact(hi));
      const double o = fma(v, row_scale * cs[c0 + j], cv[c0 + j]);
      if (FULL || c0 + j < lim)
        Cg[static_cast<long>(c0 + j) * ldc] = o;
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(OTHREADS, 1)
    gemm_ozaki_i8_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB, const OzakiParams p) {
  using Cfg = OzCfg<BN>;
  constexpr int STAGES = Cfg::STAGES, NSETS = Cfg::NSETS;
  if (p.guard != nullptr && *p.guard != 0)
    return;  // this step runs on the native fp64 kernel (uniform across the grid)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [NSETS]
  uint64_t* tmem_empty = tmem_full + 2;   // [NSETS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  double* col_scale = reinterpret_cast<double*>(tiles + STAGES * Cfg::STAGE_BYTES + 128);  // 2 x BN doubles (tile parity)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lin = blockIdx.x;
  long long* trace = (g_ozaki_clock_trace != nullptr && lin < kDbgCtas) ? g_ozaki_clock_trace : nullptr;
  if (threadIdx.x == 0)
    OZ_TRACE(0);
  const int KB = p.K / OBK;
  // This CTA owns the tiles with linear index [first, last) (row tile fastest: consecutive tiles share their B rows).
  const int first = blockIdx.x * p.tiles_per_cta;
  const int last = min(first + p.tiles_per_cta, p.gx * p.gy);
  // Rasterization: row tiles in chunks of kRowChunk (8192 rows = 29 MB of A digit planes, L2 resident), all column
  // tiles of a chunk before the next chunk, row tile fastest inside a column — so the A planes are read from HBM
  // once per chunk instead of once per column tile when the panel (115 MB at 32768 rows) exceeds L2.
  auto tile_of = [&](int l, int& row0, int& col0, long& grow0, long& gcol0) {
    const int per_chunk = kRowChunk * p.gy;
    const int chunk = l / per_chunk, rem = l - chunk * per_chunk;
    const int rows_here = min(kRowChunk, p.gx - chunk * kRowChunk);
    const int by = rem / rows_here, bx = chunk * kRowChunk + rem - by * rows_here;
    row0 = bx * OBM;
    col0 = by * BN;
    return classify_tile(p.g, row0, col0, OBM, BN, grow0, gcol0);
  };
  {
    bool any = false;
    for (int l = first; l < last && !any; ++l) {
      int r0, c0;
      long gr, gc;
      any = tile_of(l, r0, c0, gr, gc) != 0;
    }
    if (!any)
      return;  // uniform across the CTA
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], OEPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mB)) : "memory");
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
      // ===== TMA producer: runs ahead into the next tiles while MMAs / epilogues of the current ones are busy =====
      int it = 0;  // stage uses so far
      for (int l = first; l < last; ++l) {
        int row0, col0;
        long gr, gc;
        if (tile_of(l, row0, col0, gr, gc) == 0)
          continue;
        const int brow = p.b_row + (col0 / p.nbp) * p.b_tile_rows + col0 % p.nbp;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
          uint8_t* st = tiles + s * Cfg::STAGE_BYTES;
          tma_load_3d(st, &mA, &full[s], kb * OBK, p.a_row + row0, 0);
          tma_load_3d(st + A_STAGE_BYTES, &mB, &full[s], kb * OBK, brow, 0);
        }
      }
    }
  }
  else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      OZ_TRACE(1);
      int it = 0, tcount = 0;
      for (int l = first; l < last; ++l) {
        int row0, col0;
        long gr, gc;
        if (tile_of(l, row0, col0, gr, gc) == 0)
          continue;
        const int set = tcount % NSETS, use = tcount / NSETS;
        if (use > 0) {  // the epilogue must have drained this accumulator set (its previous use)
          mbar_wait(&tmem_empty[set], (use - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tset = tmem_base + static_cast<uint32_t>(set * Cfg::SET_COLS);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          if (it == 0)
            OZ_TRACE(2);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t stA = smem_u32(tiles + s * Cfg::STAGE_BYTES);
          const uint32_t stB = stA + A_STAGE_BYTES;
#pragma unroll
          for (int ks = 0; ks < OBK / 32; ++ks) {
            const uint64_t adv = static_cast<uint64_t>((ks * 32) >> 4);  // 32 int8 = 32 bytes along K inside the swizzle atom
#pragma unroll
            for (int t = 0; t < S; ++t) {
              const uint64_t ad = make_kmajor_sw64_desc(stA + t * A_PLANE_BYTES) + adv;
              // Digit plane t of A meets planes u = 0 .. 6-t of B; their groups g = t + u sit side by side in TMEM
              // (BN columns each) and the B planes side by side in shared memory (BN rows each), so ONE instruction
              // with N = BN (7 - t) covers them all — split only at the N <= 256 limit of the instruction. The first
              // product of every group (kb = ks = t = 0) overwrites its accumulator.
#pragma unroll
              for (int u0 = 0; u0 < S - t; u0 += Cfg::PLANES_PER_MMA) {
                const int planes = (S - t - u0) < Cfg::PLANES_PER_MMA ? (S - t - u0) : Cfg::PLANES_PER_MMA;
                const uint64_t bd = make_kmajor_sw64_desc(stB + u0 * Cfg::B_PLANE_BYTES) + adv;
                umma_i8(tset + static_cast<uint32_t>((t + u0) * BN), ad, bd, instr_desc_i8(planes * BN), (kb | ks | t) != 0);
              }
            }
          }
          umma_commit(&empty[s]);  // frees the stage once these MMAs have read it
        }
        if (tcount == 0)
          OZ_TRACE(3);
        umma_commit(&tmem_full[set]);  // accumulators of this tile complete
        ++tcount;
      }
    }
  }
  else {
    // ===== epilogue: 16 warps; TMEM lane quadrant = warp % 4, column quarter = (warp - 2) / 4 =====
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int r = q * 32 + lane;  // row of the tile held by this thread
    const int e = threadIdx.x - 64;  // 0 .. 32 * OEPI_WARPS - 1
    constexpr int HC = Cfg::HC;
    const int cb = part * HC;
    int tcount = 0;
    for (int l = first; l < last; ++l) {
      int row0, col0;
      long grow0, gcol0;
      const int cls = tile_of(l, row0, col0, grow0, gcol0);
      if (cls == 0)
        continue;
      const int set = tcount % NSETS, use = tcount / NSETS;
      double* cs = col_scale + (tcount & 1) * BN;
      if (e < BN)
        cs[e] = p.scale_b[p.b_row + (col0 / p.nbp) * p.b_tile_rows + col0 % p.nbp + e];
      const double row_scale = p.scale_a[p.a_row + row0 + r] * p.alpha * 0x1p-38;  // 2^-14 digits, 2^-24 from the fold
      double* Cg = p.C + row0 + r + static_cast<long>(col0 + cb) * p.ldc;
      // columns [0, lim) of this thread's segment are on or below the diagonal
      const long room = (grow0 + r) - (gcol0 + cb) + 1;
      const int lim = (cls == 1) ? HC : (room < 0 ? 0 : (room > HC ? HC : static_cast<int>(room)));
      // The C row segment of this thread is fetched while the MMAs run: after the accumulators complete only TMEM
      // loads, the integer fold, two exact conversions and the stores remain.
      double cv[HC];
#pragma unroll
      for (int j = 0; j < HC; ++j)
        cv[j] = (j < lim) ? Cg[static_cast<long>(j) * p.ldc] : 0.0;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * OEPI_WARPS) : "memory");  // column scales of this tile are in place
      mbar_wait(&tmem_full[set], use & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (threadIdx.x == 64 && tcount == 0)
        OZ_TRACE(4);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(set * Cfg::SET_COLS + cb);
      // tcgen05.ld is warp-collective (.sync.aligned): the path must be chosen per WARP, never per thread
      if (__all_sync(0xffffffffu, lim == HC))
        ozaki_fold_store<true, HC, BN>(taddr, Cg, p.ldc, cv, row_scale, cs + cb, lim);
      else
        ozaki_fold_store<false, HC, BN>(taddr, Cg, p.ldc, cv, row_scale, cs + cb, lim);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[set])) : "memory");
      if (threadIdx.x == 64 && tcount == 0)
        OZ_TRACE(5);
      ++tcount;
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
  if (threadIdx.x == 0)
    OZ_TRACE(6);
}

// x (rows x kdim, column-major with leading dimension ld) -> 7 int8 digit planes (K-major) + 2^e per row.
// Block = 32 rows x 8 k-groups (256 threads); thread (r, kg) owns k = kg * PER .. of its row in registers (loads of a
// warp = 32 consecutive rows of one column: coalesced).
// Digits: |x| 2^-e < 1/2 (e = ilogb(row max) + 2); M = rn(x 2^(55-e)) is a 55-bit signed integer, cut from the LOW end
// into balanced radix-256 digits d_6 .. d_1 in [-128, 127] (d = ((M + 128) & 255) - 128, M <- (M - d) / 256) and a top
// digit |d_0| <= 65:   x = 2^e sum_t d_t 2^(-7-8t) + r,  |r| <= 2^(e-56)  (entries within a factor 4 of the row maximum
// are exact). `flag` (may be null): raised when a nonzero entry is rounded AND keeps fewer than min_bits significant bits.
template <int KDIM>
__global__ void __launch_bounds__(256) split_i8_kernel(const double* __restrict__ x, long ld, int rows,
                                                       signed char* __restrict__ q, long plane_stride,
                                                       double* __restrict__ scale, int tile_rows, long tile_stride,
                                                       int* __restrict__ flag, int min_bits, int dst_row0) {
  constexpr int KG = 8, PER = KDIM / KG;  // k values per thread, contiguous chunk [kg * PER, (kg + 1) * PER)
  __shared__ double smax[KG][33];
  const int tr = threadIdx.x & 31, kg = threadIdx.x >> 5;
  const int r = blockIdx.x * 32 + tr;
  const bool live = r < rows;
  const long roff = live ? (tile_stride ? (r / tile_rows) * tile_stride + r % tile_rows : r) : 0;
  double v[PER];
  double m = 0.0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    v[i] = live ? x[roff + static_cast<long>(kg * PER + i) * ld] : 0.0;
    m = isfinite(v[i]) ? fmax(m, fabs(v[i])) : INFINITY;  // Inf / NaN poison the whole row (see the scale below)
  }
  smax[kg][tr] = m;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < KG; ++i)
    m = fmax(m, smax[i][tr]);
  // |v| * 2^-e < 0.5   (e = ilogb(max) + 2);   all-zero (or non-finite) rows: e = 0
  int e = 0;
  if (m > 0.0 && m < 1.0e300)
    e = ilogb(m) + 2;
  e = e < -1000 ? -1000 : (e > 1000 ? 1000 : e);
  const double down = __hiloint2double((1023 - e) << 20, 0);  // 2^-e
  // A row holding Inf / NaN gets a NaN scale: every C entry it contributes to becomes NaN, like in a native fp64 update
  // (the digits themselves cannot carry non-finite values).
  if (kg == 0 && live)
    scale[dst_row0 + r] = (m < INFINITY) ? __hiloint2double((1023 + e) << 20, 0) : __longlong_as_double(0x7FF8000000000000LL);
  long long M[PER];
  bool starved = false;
  const long long keep = (min_bits > 0) ? (1LL << (min_bits - 1)) : 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const double sv = (m < INFINITY) ? (v[i] * down) * 0x1p55 : 0.0;  // exact scaling (two steps: 2^(55-e) may overflow)
    M[i] = __double2ll_rn(sv);
    const long long am = M[i] < 0 ? -M[i] : M[i];
    starved |= (static_cast<double>(M[i]) != sv) && (am < keep);
  }
  if (flag != nullptr && starved && live)
    atomicOr(flag, 1);
  if (!live)
    return;
  signed char* dst = q + static_cast<long>(dst_row0 + r) * KDIM + kg * PER;
#pragma unroll 1
  for (int t = S - 1; t >= 0; --t) {
    uint32_t w[PER / 4];
#pragma unroll
    for (int i = 0; i < PER; i += 4) {
      uint32_t pack = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const long long mm = M[i + b];
        const int d = (t > 0) ? static_cast<int>((mm + 128) & 255) - 128 : static_cast<int>(mm);
        M[i + b] = (mm - d) >> 8;
        pack |= (static_cast<uint32_t>(d) & 0xFFu) << (8 * b);
      }
      w[i / 4] = pack;
    }
    uint4* o = reinterpret_cast<uint4*>(dst + t * plane_stride);
#pragma unroll
    for (int i = 0; i < PER / 16; ++i)
      o[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  }
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    DLAF_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr));
    DLAF_B200_ASSERT(f != nullptr && qr == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    return reinterpret_cast<EncodeFn>(f);
  }();
  return fn;
}

// C tile width of the int8 kernel: 64 (one TMEM accumulator set, default) or 32 (two sets: the epilogue overlaps the next
// tile's MMAs; DLAF_B200_OZAKI_BN=32). Measured (profiles/r02_ozaki_i8_v5_*.log): the MMA phase of BOTH variants runs at
// ~98 B/clk of shared-memory operand reads, not at the tensor pipe's rate — 128 x 32 tiles read 56 KB per k-step
// (585 clk instead of 448), 128 x 64 tiles 96 KB (986 clk instead of 896) — so the wider tile wins (93 vs 85 TFLOP/s
// fp64-equivalent) although its epilogue is exposed.
int ozaki_bn() {
  static const int v = [] {
    const char* e = std::getenv("DLAF_B200_OZAKI_BN");
    const int b = e ? std::atoi(e) : 64;
    return b == 32 ? 32 : 64;
  }();
  return v;
}

}  // namespace

int ozaki_min_bits() {
  static const int v = [] {
    const char* e = std::getenv("DLAF_B200_OZAKI_MIN_BITS");
    const int b = e ? std::atoi(e) : 16;
    return b < 0 ? 0 : (b > 53 ? 53 : b);
  }();
  return v;
}

void OzakiSplit::allocate(long rows_max, int kdim_) {
  release();
  rows = rows_max;
  kdim = kdim_;
  DLAF_B200_ASSERT(kdim % 128 == 0, "Ozaki split: k must be a multiple of 128");
  q = pool_alloc<signed char>(static_cast<size_t>(S) * rows * kdim);
  scale = pool_alloc<double>(rows);
  DLAF_CUDA_CHECK(cudaMemset(q, 0, static_cast<size_t>(S) * rows * kdim));
  // 3-D maps: dim0 = k (contiguous, bytes), dim1 = row, dim2 = digit plane; box = 64 k x {128, BN} rows x 7 planes;
  // 64-byte swizzle
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(kdim), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(S)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(kdim), static_cast<cuuint64_t>(kdim) * static_cast<cuuint64_t>(rows)};
  const cuuint32_t estr[3] = {1, 1, 1};
  for (int i = 0; i < 2; ++i) {
    const cuuint32_t box[3] = {OBK, static_cast<cuuint32_t>(i == 0 ? OBM : ozaki_bn()), S};
    CUtensorMap* m = reinterpret_cast<CUtensorMap*>(i == 0 ? map_a : map_b);
    const CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, q, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DLAF_B200_ASSERT(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
  }
}

void OzakiSplit::release() {
  pool_free(q);
  pool_free(scale);
  q = nullptr;
  scale = nullptr;
}

void OzakiSplit::split(const double* x, long ld, long nrows, cudaStream_t s, int tile_rows, long tile_stride, int* flag,
                       long dst_row0) {
  DLAF_B200_ASSERT(dst_row0 + nrows <= rows, "split buffer too small");
  if (nrows <= 0)
    return;
  const unsigned grid = static_cast<unsigned>((nrows + 31) / 32);
  const long plane_stride = rows * static_cast<long>(kdim);
  const int tr = tile_rows > 0 ? tile_rows : 1;
  const int mb = ozaki_min_bits();
  switch (kdim) {
    case 128: split_i8_kernel<128><<<grid, 256, 0, s>>>(x, ld, static_cast<int>(nrows), q, plane_stride, scale, tr, tile_stride, flag, mb, static_cast<int>(dst_row0)); break;
    case 256: split_i8_kernel<256><<<grid, 256, 0, s>>>(x, ld, static_cast<int>(nrows), q, plane_stride, scale, tr, tile_stride, flag, mb, static_cast<int>(dst_row0)); break;
    case 384: split_i8_kernel<384><<<grid, 256, 0, s>>>(x, ld, static_cast<int>(nrows), q, plane_stride, scale, tr, tile_stride, flag, mb, static_cast<int>(dst_row0)); break;
    case 512: split_i8_kernel<512><<<grid, 256, 0, s>>>(x, ld, static_cast<int>(nrows), q, plane_stride, scale, tr, tile_stride, flag, mb, static_cast<int>(dst_row0)); break;
    default: DLAF_B200_ASSERT(false, "Ozaki split: unsupported k (128, 256, 384 or 512)");
  }
  DLAF_CUDA_CHECK(cudaGetLastError());
}

namespace {
// int8 tensor-pipe peak of this device: every SM issues `iters` back-to-back tcgen05.mma.kind::i8 (M128 N256 K32) on
// one resident shared-memory tile (no loads in the loop) — the roofline denominator of the Ozaki kernel.
__global__ void __launch_bounds__(64, 1) i8_peak_kernel(int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* done = reinterpret_cast<uint64_t*>(tiles + 24576);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  for (int i = threadIdx.x; i < 24576 / 4; i += 64)
    reinterpret_cast<uint32_t*>(tiles)[i] = 0x01010101u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy tile writes -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1 && lane == 0) {
    const uint64_t ad = make_kmajor_sw64_desc(smem_u32(tiles));
    const uint64_t bd = make_kmajor_sw64_desc(smem_u32(tiles + 8192));
    for (int i = 0; i < iters; ++i)
      umma_i8(tmem_base + ((i & 1) ? 256u : 0u), ad + ((i & 2) ? 2 : 0), bd + ((i & 2) ? 2 : 0), instr_desc_i8(256), i >= 2);
    umma_commit(done);
    mbar_wait(done, 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}
}  // namespace

void ozaki_set_clock_trace(long long* dev_buffer) {
  DLAF_CUDA_CHECK(cudaMemcpyToSymbol(g_ozaki_clock_trace, &dev_buffer, sizeof(dev_buffer)));
}

static_assert(sizeof(CUtensorMap) == 128, "tensor map size");

namespace {
template <int BN>
void launch_ozaki_bn(const OzakiParams& p0, const GemmArgsT<double>& a, const OzakiSplit& sa, const OzakiSplit& sb,
                     cudaStream_t stream) {
  using Cfg = OzCfg<BN>;
  static bool configured = false;
  if (!configured) {
    DLAF_CUDA_CHECK(cudaFuncSetAttribute(gemm_ozaki_i8_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES));
    configured = true;
  }
  OzakiParams p = p0;
  p.gx = a.M / OBM;
  p.gy = a.N / BN;
  // Tiles per CTA: large launches amortise set-up over up to 16 tiles (~60-120 us CTAs), small ones keep every SM
  // busy. DLAF_B200_OZAKI_TPC overrides.
  static const int tpc_env = [] {
    const char* e = std::getenv("DLAF_B200_OZAKI_TPC");
    return e ? std::atoi(e) : 0;
  }();
  const long ntiles = static_cast<long>(p.gx) * p.gy;
  const long unit = ntiles * BN / 64;  // in round-1 tile units (128 x 64)
  int tpc = tpc_env;
  if (tpc <= 0)  // trailing matrix >= 16K: 8 units, >= 8K: 4, >= 4K: 2
    tpc = (unit >= 32768 ? 8 : (unit >= 8192 ? 4 : (unit >= 2048 ? 2 : 1))) * (64 / BN);
  p.tiles_per_cta = tpc;
  const unsigned grid = static_cast<unsigned>((ntiles + tpc - 1) / tpc);
  gemm_ozaki_i8_kernel<BN><<<grid, OTHREADS, Cfg::SMEM_BYTES, stream>>>(*reinterpret_cast<const CUtensorMap*>(sa.map_a),
                                                                        *reinterpret_cast<const CUtensorMap*>(sb.map_b), p);
  DLAF_CUDA_CHECK(cudaGetLastError());
}
}  // namespace

void launch_gemm_ozaki_i8(const GemmArgsT<double>& a, const OzakiSplit& sa, long a_row, const OzakiSplit& sb, long b_row,
                          cudaStream_t stream, long b_tile_rows, const int* guard) {
  if (a.M <= 0 || a.N <= 0)
    return;
  DLAF_B200_ASSERT(a.M % OBM == 0 && a.N % 64 == 0 && a.K % OBK == 0 && a.K == sa.kdim && a.K == sb.kdim && a.K <= 512,
                   "ozaki gemm shape");
  {
    int ex = 0;
    const double mant = std::frexp(a.alpha < 0 ? -a.alpha : a.alpha, &ex);
    DLAF_B200_ASSERT(mant == 0.5 && a.beta == 1.0, "ozaki gemm: C += alpha A B^T with alpha = +-2^e only (exact scaling)");
  }
  OzakiParams p;
  p.C = a.C;
  p.ldc = a.ldc;
  p.K = a.K;
  p.alpha = a.alpha;
  p.g = a;
  p.a_row = static_cast<int>(a_row);
  p.b_row = static_cast<int>(b_row);
  p.nbp = a.nbp;
  p.b_tile_rows = static_cast<int>(b_tile_rows > 0 ? b_tile_rows : a.nbp);
  p.scale_a = sa.scale;
  p.scale_b = sb.scale;
  p.gx = p.gy = 0;
  p.tiles_per_cta = 1;
  p.guard = guard;
  if (ozaki_bn() == 64)
    launch_ozaki_bn<64>(p, a, sa, sb, stream);
  else
    launch_ozaki_bn<32>(p, a, sa, sb, stream);
}

}  // namespace dlaf_b200

extern "C" double dlaf_b200_measure_int8_tensor_peak_tops(void) {
  using namespace dlaf_b200;
  int dev = 0, nsm = 0;
  DLAF_CUDA_CHECK(cudaGetDevice(&dev));
  DLAF_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const int smem = 24576 + 1024 + 64, iters = 40000;
  DLAF_CUDA_CHECK(cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  DLAF_CUDA_CHECK(cudaEventCreate(&e0));
  DLAF_CUDA_CHECK(cudaEventCreate(&e1));
  i8_peak_kernel<<<nsm, 64, smem>>>(400);
  double best = 0;
  for (int rep = 0; rep < 3; ++rep) {
    DLAF_CUDA_CHECK(cudaEventRecord(e0));
    i8_peak_kernel<<<nsm, 64, smem>>>(iters);
    DLAF_CUDA_CHECK(cudaEventRecord(e1));
    DLAF_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    DLAF_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double ops = 2.0 * 128 * 256 * 32 * double(iters) * nsm;
    best = ops / ms / 1e9 > best ? ops / ms / 1e9 : best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;  // TOP/s (1e12 int8 multiply-adds x 2 per second)
}
