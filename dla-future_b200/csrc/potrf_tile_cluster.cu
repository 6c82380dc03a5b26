// Whole diagonal tile (up to 512 x 512, fp64) in ONE launch on a thread-block cluster: Cholesky factor of the tile plus
// the inverses of its 128 x 128 diagonal blocks (what turns the panel TRSM into tensor-core GEMMs).
//
// Replaces potrfDiagTile -> tile::potrf -> cusolverDnDpotrf (+ bufferSize query, workspace, assert_info<<<1,1>>>) of the
// reference (include/dlaf/factorization/cholesky/impl.h:46-53, include/dlaf/lapack/tile.h:696-725) and round 1's own
// sequence of 4 single-CTA block kernels + 6 small GEMM launches per tile (engine.cu: factor_diag_tile; 342 us per
// 512-tile, 0.35 % of the fp64 peak: one SM did the arithmetic and ten dependent launches the rest).
//
// Algorithm: right-looking, panels of 8 columns (nbp / 8 <= 64 panels), on a cluster of 8 CTAs x 16 warps.
//   * The lower triangle (1 MB) lives in the SHARED MEMORY of the 8 SMs as 8 x 8 DMMA accumulator fragments stored
//     lane-major (one LDS.128 / STS.128 per lane and fragment, conflict-free); block column Jc (8 columns) belongs to CTA
//     Jc % 8, fragment rows are dealt to the warps round-robin. (Registers cannot hold it: 133-147 KB per CTA = 64-72
//     of the 128 registers a thread of a 512-thread CTA may use, before any working set.) Column-cyclic ownership means
//     a panel is factorised entirely inside its owner CTA — no exchange inside a panel.
//   * Panel J, on its owner: fragments -> shared memory (raw block column), ONE thread factorises the 8 x 8 pivot block
//     (a latency-bound chain of rsqrt + dependent FMAs, ~100 clk per column, the floor of any Cholesky), every other
//     thread then solves ONE panel row against it (independent rows), and the finished columns are written straight
//     into the tile in global memory — which is both the result and the broadcast medium: after a cluster barrier
//     (release / acquire) the other CTAs read the panel back from L2 (ld.global.cg) into their own shared memory.
//   * Trailing update: fragment(I, Jc) -= P(I) P(Jc)^T with two DMMA.8x8x4 per fragment and panel, operands from the
//     k-major panel copy (leading dimension = 8 mod 16 doubles: conflict-free fragment loads).
//   * Look-ahead for free: the owner of panel J first brings ITS block column J up to date, factorises and publishes
//     it, and only then applies panel J-1 to the rest of its fragments; the other CTAs arrive at the cluster barrier
//     before their own update (split arrive / wait), so the critical path per panel is
//         factor + row solve + barrier + panel read-back  (~2.5k clk),   not   ... + trailing update.
//   * Phase 2: inv(L_bb) of the ns = nbp / 128 diagonal blocks, one CTA per block, one warp per 8 columns of the inverse:
//     block forward substitution on 8 x 8 fragments (DMMA), the 8 x 8 diagonal inverses computed first by 16 threads.
// Contract: only the lower triangle of T is read or written; W receives ns blocks of 128 x 128 (column-major, ld 128,
// lower = inverse, strictly upper = 0); non-SPD -> *info = info_offset + 1-based column (first failure wins), no trap,
// the remaining entries are then unspecified (NaN), like a failed LAPACK potrf.
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>

#include "common.h"
#include "potrf_tile.cuh"

namespace dlaf_b200 {

namespace {

constexpr int CL = 8;                  // CTAs per cluster
constexpr int NW = 16;                 // warps per CTA
constexpr int NTHR = NW * 32;          // 512
constexpr int MAXNB = 512;             // largest tile
constexpr int NFC = MAXNB / 8;         // block columns / fragment rows of the largest tile (64)
constexpr int PLD = MAXNB + 8;         // panel copies are k-major: P[k * PLD + row]; PLD = 8 (mod 16) doubles
constexpr int JL = NFC / CL;           // block columns per CTA (8)
constexpr int GB = 128;                // diagonal blocks whose inverses the TRSM wants
// fragments of CTA 0 (the most): block columns 0, 8, .., 56 with 64, 56, .., 8 fragment rows
constexpr int MAXFRAGS = JL * (NFC + CL) / 2;  // 288

// Shared memory of one CTA (214 KB): its part of the lower triangle as 8 x 8 DMMA accumulator fragments (64 doubles each,
// lane-major: lane l owns doubles 2l, 2l+1 = row l/4, columns 2 (l%4), +1), two k-major panel buffers, the pivot factor.
struct __align__(16) TileSmem {
  double frag[MAXFRAGS * 64];
  double P[2][8 * PLD];   // finished panels J (parity J % 2); phase 2 reuses them (minv, scratch)
  double dfL[64];         // pivot block: strictly lower part of L_D (row-major a * 8 + b)
  double dfinv[8];        // 1 / diag(L_D)
  double dfpiv[8];        // the pivots themselves (their square roots are taken off the critical chain)
  // Written by the NEXT rank of the cluster (DSMEM): the highest panel index whose landing buffer over there is free,
  // i.e. this CTA may push panel J into P[J & 1] of rank (J + 1) % CL once free_for >= J.
  int free_for;
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  // (not volatile: independent fragments may be interleaved by the scheduler)
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned map_to_rank(const void* local_smem, unsigned rank) {
  const unsigned la = static_cast<unsigned>(__cvta_generic_to_shared(local_smem));
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  return ra;
}
__device__ __forceinline__ void st_cluster_f64(unsigned raddr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(raddr), "d"(v) : "memory");
}
__device__ __forceinline__ void st_release_cluster_s32(unsigned raddr, int v) {
  asm volatile("st.release.cluster.shared::cluster.s32 [%0], %1;" ::"r"(raddr), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cluster_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.cluster.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(static_cast<unsigned>(__cvta_generic_to_shared(p))) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ double rsqrt_nr(double a) {
  const double y = rsqrt(a);
  return fma(y * 0.5, fma(-a * y, y, 1.0), y);
}

// Measurement aid: clock64 stamps of CTA 0 / thread 0 per panel step (tools/gpu_diag_tile_test): 8 per step
__device__ long long* g_tile_clock_trace = nullptr;

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NTHR, 1)
    potrf_tile_cluster_kernel(double* __restrict__ T, long ldt, double* __restrict__ W, int nbp, int* info,
                              int info_offset) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem& sm = *reinterpret_cast<TileSmem*>(smem_raw);
  unsigned rank_u;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank_u));
  const int rank = static_cast<int>(rank_u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int nfc = nbp / 8;  // block columns / fragment rows of the tile
  long long* trace = (g_tile_clock_trace != nullptr && rank == 0 && tid == 0) ? g_tile_clock_trace : nullptr;

  // fragment (I, Jc = rank + CL * jl), I >= Jc, lives at frag[(foff(jl) + I - Jc) * 64 + 2 * lane]
  auto foff = [&](int jl) {  // fragments of my block columns before local column jl
    int o = 0;
    for (int q = 0; q < jl; ++q)
      o += nfc - (rank + CL * q);
    return o;
  };
  int fo[JL];
#pragma unroll
  for (int jl = 0; jl < JL; ++jl)
    fo[jl] = foff(jl);

  // ---- load my fragments (lower triangle only; the strictly upper part of diagonal fragments is zero-filled)
  for (int jl = 0; jl < JL; ++jl) {
    const int Jc = rank + CL * jl;
    if (Jc >= nfc)
      break;
    for (int I = Jc + warp; I < nfc; I += NW) {
      const int r = 8 * I + g, s = 8 * Jc + 2 * tig;
      double2 v;
      v.x = (r >= s) ? T[r + static_cast<long>(s) * ldt] : 0.0;
      v.y = (r >= s + 1) ? T[r + static_cast<long>(s + 1) * ldt] : 0.0;
      *reinterpret_cast<double2*>(&sm.frag[(fo[jl] + I - Jc) * 64 + 2 * lane]) = v;
    }
  }
  __syncthreads();

  // applies panel `pb` (block column Jp) to my fragments of local block columns jl in [jl_lo, jl_hi] with Jc > Jp;
  // rows are dealt to the warps round-robin (I % NW), so a warp loads the A operand of a fragment row once
  auto update = [&](const double* __restrict__ pb, int Jp, int jl_lo, int jl_hi) {
    double b0[JL], b1[JL];
#pragma unroll
    for (int jl = 0; jl < JL; ++jl) {
      const int Jc = rank + CL * jl;
      const bool on = (jl >= jl_lo && jl <= jl_hi && Jc > Jp && Jc < nfc);
      b0[jl] = on ? pb[tig * PLD + 8 * Jc + g] : 0.0;
      b1[jl] = on ? pb[(4 + tig) * PLD + 8 * Jc + g] : 0.0;
    }
    // two fragment rows per trip: two independent load -> DMMA -> DMMA -> store chains in flight per warp
    for (int I = Jp + 1 + ((warp - (Jp + 1)) & (NW - 1)); I < nfc; I += 2 * NW) {  // I > Jp, I % NW == warp
      const int I2 = I + NW;
      const bool two = I2 < nfc;
      const double a0 = -pb[tig * PLD + 8 * I + g];
      const double a1 = -pb[(4 + tig) * PLD + 8 * I + g];
      const double a2 = two ? -pb[tig * PLD + 8 * I2 + g] : 0.0;
      const double a3 = two ? -pb[(4 + tig) * PLD + 8 * I2 + g] : 0.0;
#pragma unroll
      for (int jl = 0; jl < JL; ++jl) {
        const int Jc = rank + CL * jl;
        if (jl >= jl_lo && jl <= jl_hi && Jc > Jp && Jc <= I2 && Jc < nfc) {  // warp-uniform
          const bool first = (Jc <= I), second = two;
          double2* f = reinterpret_cast<double2*>(&sm.frag[(fo[jl] + I - Jc) * 64 + 2 * lane]);
          double2* f2 = reinterpret_cast<double2*>(&sm.frag[(fo[jl] + I2 - Jc) * 64 + 2 * lane]);
          double2 c = first ? *f : make_double2(0.0, 0.0);
          double2 d = second ? *f2 : make_double2(0.0, 0.0);
          dmma884(c.x, c.y, a0, b0[jl]);
          dmma884(d.x, d.y, a2, b0[jl]);
          dmma884(c.x, c.y, a1, b1[jl]);
          dmma884(d.x, d.y, a3, b1[jl]);
          if (first)
            *f = c;
          if (second)
            *f2 = d;
        }
      }
    }
  };

  // ================================ phase 1: the factor, panel by panel ================================
  if (tid == 0)
    sm.free_for = 1;  // the landing buffers of panels 0 and 1 have never been used
  cluster_arrive_release();
  cluster_wait_acquire();  // every CTA's flag is initialised before a peer may write it
  for (int J = 0; J < nfc; ++J) {
    const bool own = (rank == J % CL);
    const bool has_next = (J + 1 < nfc);
    const int next_rank = (J + 1) % CL;
    const int jlJ = J / CL;
    if (trace)
      trace[J * 8 + 0] = clock64();
    if (own) {
      double* pn = sm.P[J & 1];  // (holds panel J-2: dead)
      if (J > 0)
        update(sm.P[(J - 1) & 1], J - 1, jlJ, jlJ);  // my block column J first: it is the critical path
      // fragments of block column J -> raw panel, k-major, straight into the buffer the finished panel will occupy; same
      // row -> warp mapping as the update, so no barrier in between
      const int foJ = foff(jlJ);  // (fo[] stays statically indexed: registers, not local memory)
      for (int I = J + ((warp - J) & (NW - 1)); I < nfc; I += NW) {
        const double2 c = *reinterpret_cast<const double2*>(&sm.frag[(foJ + I - J) * 64 + 2 * lane]);
        pn[(2 * tig) * PLD + 8 * I + g] = c.x;
        pn[(2 * tig + 1) * PLD + 8 * I + g] = c.y;
      }
      if (trace)
        trace[J * 8 + 1] = clock64();
      if (warp == (J & (NW - 1))) {
        // this warp holds the diagonal fragment: its lane 0 factorises the 8 x 8 pivot block right away (a chain of
        // rsqrt + dependent FMAs, everything in registers with static indices) while the other warps finish their rows
        __syncwarp();
        if (lane == 0) {
          double D[8][8];
#pragma unroll
          for (int b = 0; b < 8; ++b)
#pragma unroll
            for (int a = 0; a < 8; ++a)
              D[a][b] = (a >= b) ? pn[b * PLD + 8 * J + a] : 0.0;
          int fail = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const double ajj = D[j][j];
            if (!(ajj > 0.0) && fail == 0)
              fail = j + 1;
            const double inv = rsqrt_nr(ajj);
            sm.dfinv[j] = inv;
            sm.dfpiv[j] = ajj;
#pragma unroll
            for (int a = j + 1; a < 8; ++a) {
              D[a][j] *= inv;
              sm.dfL[a * 8 + j] = D[a][j];
            }
#pragma unroll
            for (int s = j + 1; s < 8; ++s)
#pragma unroll
              for (int a = s; a < 8; ++a)
                D[a][s] = fma(-D[a][j], D[s][j], D[a][s]);
          }
          if (fail)
            atomicCAS(info, 0, info_offset + 8 * J + fail);
        }
      }
      else if (warp == ((J + 1) & (NW - 1)) && lane == 0 && has_next) {
        // the landing buffer of panel J on the next rank must have been released (it held panel J-2)
        for (unsigned spin = 0; ld_acquire_cluster_s32(&sm.free_for) < J; ++spin)
          if (spin > (1u << 24))
            __trap();
      }
      if (trace)
        trace[J * 8 + 2] = clock64();
      __syncthreads();
      if (trace)
        trace[J * 8 + 3] = clock64();
      {
        // one panel row per thread: x <- x L_D^-T (forward substitution, columns left to right), in place; the finished
        // row goes to my own panel buffer, into the tile in global memory (result, and the medium the other CTAs read it
        // back from) and straight into the panel buffer of the NEXT owner (DSMEM), who therefore never waits for L2
        const int r = 8 * (J + 1) + tid;
        if (r < nbp) {
          double x[8], Lm[8][8], invd[8];
#pragma unroll
          for (int a = 0; a < 8; ++a) {
            invd[a] = sm.dfinv[a];
#pragma unroll
            for (int b = 0; b < 8; ++b)
              Lm[a][b] = (a > b) ? sm.dfL[a * 8 + b] : 0.0;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            x[k] = pn[k * PLD + r];
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            x[s] *= invd[s];
#pragma unroll
            for (int t = s + 1; t < 8; ++t)
              x[t] = fma(-x[s], Lm[t][s], x[t]);
          }
          const unsigned remote = has_next ? map_to_rank(&pn[r], static_cast<unsigned>(next_rank)) : 0u;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            pn[k * PLD + r] = x[k];
            if (has_next)
              st_cluster_f64(remote + static_cast<unsigned>(k * PLD * sizeof(double)), x[k]);
            T[r + static_cast<long>(8 * J + k) * ldt] = x[k];  // coalesced along r
          }
        }
        if (tid < 64) {  // the pivot block itself: strictly lower part from the factor, diagonal = correctly rounded sqrt
          const int a = tid >> 3, b = tid & 7;
          if (a > b)
            T[(8 * J + a) + static_cast<long>(8 * J + b) * ldt] = sm.dfL[a * 8 + b];
          else if (a == b)
            T[static_cast<long>(8 * J + a) * (1 + ldt)] = sqrt(sm.dfpiv[a]);
        }
      }
    }
    if (trace)
      trace[J * 8 + 4] = clock64();
    cluster_arrive_release();  // (release at cluster scope: the global and DSMEM stores above are visible after the wait)
    if (trace)
      trace[J * 8 + 5] = clock64();
    if (J > 0) {
      update(sm.P[(J - 1) & 1], J - 1, own ? jlJ + 1 : 0, JL - 1);  // the rest (overlaps the owner's critical part)
      if (rank == (J + 2) % CL && J + 1 < nfc) {
        // I will RECEIVE panel J+1 in P[(J+1) & 1] = the buffer of panel J-1, which nobody here reads any more: tell its owner
        __syncthreads();
        if (tid == 0)
          st_release_cluster_s32(map_to_rank(&sm.free_for, static_cast<unsigned>((J + 1) % CL)), J + 1);
      }
    }
    if (trace)
      trace[J * 8 + 6] = clock64();
    cluster_wait_acquire();  // panel J is published
    if (!own && rank != next_rank && has_next) {  // (the next owner got it by DSMEM)
      const int r = 8 * (J + 1) + tid;
      if (r < nbp) {
        double* pn = sm.P[J & 1];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          pn[k * PLD + r] = __ldcg(T + r + static_cast<long>(8 * J + k) * ldt);
      }
    }
    __syncthreads();
    if (trace)
      trace[J * 8 + 7] = clock64();
  }

  // ================================ phase 2: inverses of the 128 x 128 diagonal blocks ================================
  const int ns = nbp / GB;
  if (rank < ns) {
    double* minv = &sm.P[0][0];           // 16 x 64: inverses of the 8 x 8 diagonal blocks of my 128-block (row-major)
    double* scratch = &sm.P[1][0];        // NW x 64: per-warp layout-conversion buffer
    const double* Lb = T + static_cast<long>(rank) * GB * (1 + ldt);
    double* Wb = W + static_cast<long>(rank) * GB * GB;
    if (tid < 16) {
      const double* Li = Lb + static_cast<long>(8 * tid) * (1 + ldt);
      double Lm[8][8], M[8][8];
#pragma unroll
      for (int b = 0; b < 8; ++b)
#pragma unroll
        for (int a = 0; a < 8; ++a)
          Lm[a][b] = (a >= b) ? __ldcg(Li + a + static_cast<long>(b) * ldt) : 0.0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int a = 0; a < 8; ++a)
          M[a][c] = 0.0;
        M[c][c] = 1.0 / Lm[c][c];
#pragma unroll
        for (int s = c + 1; s < 8; ++s) {
          double v = 0.0;
#pragma unroll
          for (int j = c; j < s; ++j)
            v = fma(-Lm[s][j], M[j][c], v);
          M[s][c] = v / Lm[s][s];
        }
      }
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b)
          minv[tid * 64 + a * 8 + b] = M[a][b];
    }
    __syncthreads();
    // warp = fragment column jc of X = inv(L_bb): X(i, jc) = inv(L_ii) (delta_{i,jc} I - sum_{jc <= k < i} L(i,k) X(k,jc))
    const int jc = warp;
    double* sc = scratch + warp * 64;
    double xb0[16], xb1[16];  // finished X(k, jc) as DMMA B operands (k = tig / 4 + tig, n = g)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      xb0[i] = xb1[i] = 0.0;
      const long wo = (8 * i + g) + static_cast<long>(8 * jc + 2 * tig) * GB;
      if (i < jc) {  // strictly upper part of the inverse: zeros (warp-uniform)
        Wb[wo] = 0.0;
        Wb[wo + GB] = 0.0;
        continue;
      }
      double t0 = (i == jc && g == 2 * tig) ? 1.0 : 0.0;
      double t1 = (i == jc && g == 2 * tig + 1) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (k < i && k >= jc) {  // warp-uniform
          const double a0 = -__ldcg(Lb + (8 * i + g) + static_cast<long>(8 * k + tig) * ldt);
          const double a1 = -__ldcg(Lb + (8 * i + g) + static_cast<long>(8 * k + 4 + tig) * ldt);
          dmma884(t0, t1, a0, xb0[k]);
          dmma884(t0, t1, a1, xb1[k]);
        }
      }
      // accumulator layout (row g, columns 2 tig, 2 tig + 1) -> B operand layout (k = tig, n = g) through shared memory
      sc[g * 8 + 2 * tig] = t0;
      sc[g * 8 + 2 * tig + 1] = t1;
      __syncwarp();
      const double b0 = sc[tig * 8 + g], b1 = sc[(4 + tig) * 8 + g];
      __syncwarp();
      const double m0 = minv[i * 64 + g * 8 + tig], m1 = minv[i * 64 + g * 8 + 4 + tig];
      double r0 = 0.0, r1 = 0.0;
      dmma884(r0, r1, m0, b0);
      dmma884(r0, r1, m1, b1);
      Wb[wo] = r0;
      Wb[wo + GB] = r1;
      sc[g * 8 + 2 * tig] = r0;
      sc[g * 8 + 2 * tig + 1] = r1;
      __syncwarp();
      xb0[i] = sc[tig * 8 + g];
      xb1[i] = sc[(4 + tig) * 8 + g];
      __syncwarp();
    }
  }
  // no CTA may leave while a peer could still be inside a cluster barrier phase: all of them passed the last wait above
}

bool tile_kernel_enabled() {
  // DLAF_B200_POTRF_TILE=blocks keeps round 1's per-128-block sequence (A/B measurements)
  static const bool on = [] {
    const char* e = std::getenv("DLAF_B200_POTRF_TILE");
    return e == nullptr || std::string(e) != "blocks";
  }();
  return on;
}

}  // namespace

void potrf_tile_set_clock_trace(long long* dev_buffer) {
  DLAF_CUDA_CHECK(cudaMemcpyToSymbol(g_tile_clock_trace, &dev_buffer, sizeof(dev_buffer)));
}

bool potrf_tile_cluster_supported(int nbp) {
  return tile_kernel_enabled() && nbp >= GB && nbp <= MAXNB && nbp % GB == 0;
}

void launch_potrf_tile_cluster_f64(double* T, long ldt, double* W, int nbp, int* info, int info_offset,
                                   cudaStream_t stream) {
  DLAF_B200_ASSERT(nbp >= GB && nbp <= MAXNB && nbp % GB == 0, "cluster tile kernel: 128, 256, 384 or 512");
  static bool configured = false;
  if (!configured) {
    DLAF_CUDA_CHECK(cudaFuncSetAttribute(potrf_tile_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(sizeof(TileSmem))));
    configured = true;
  }
  potrf_tile_cluster_kernel<<<CL, NTHR, sizeof(TileSmem), stream>>>(T, ldt, W, nbp, info, info_offset);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

}  // namespace dlaf_b200
