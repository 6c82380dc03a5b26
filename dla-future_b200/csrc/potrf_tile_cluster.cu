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
// Panel copies are row-major, 64 bytes per row, entries permuted to [x0 x4 x1 x5 x2 x6 x3 x7] and the four 16-byte chunks of
// a row XOR-swizzled with (row >> 1) & 3:
//   * lane (g, tig) of a warp fetches BOTH DMMA operand halves of its row — (x_tig, x_{4+tig}) — with one LDS.128, and a
//     quarter-warp (2 rows x 4 chunks) covers 128 contiguous bytes: conflict-free;
//   * the thread that solves / reads back ONE row touches its 4 chunks at 64-byte stride; the swizzle spreads 8
//     consecutive rows over all 8 16-byte bank groups (without it: 4-way conflicts, measured 4.5 clk per row);
//   * rows 8 (J+1) .. of a panel are one contiguous byte range: ONE bulk DSMEM copy hands it to the next owner.
__host__ __device__ constexpr int ppos(int k) { return (k & 3) * 2 + (k >> 2); }
// index (in doubles) of chunk q (= entries x_q, x_{4+q}) of row `row`
__host__ __device__ constexpr int pchunk(int row, int q) { return row * 8 + ((q ^ ((row >> 1) & 3)) << 1); }
// index of entry k of row `row`
__host__ __device__ constexpr int pidx(int row, int k) { return pchunk(row, ppos(k) >> 1) + (ppos(k) & 1); }
constexpr int JL = NFC / CL;           // block columns per CTA (8)
constexpr int GB = 128;                // diagonal blocks whose inverses the TRSM wants
// fragments of CTA 0 (the most): block columns 0, 8, .., 56 with 64, 56, .., 8 fragment rows
constexpr int MAXFRAGS = JL * (NFC + CL) / 2;  // 288

// Shared memory of one CTA (209 KB): its part of the lower triangle as 8 x 8 DMMA accumulator fragments (64 doubles each,
// lane-major: lane l owns doubles 2l, 2l+1 = row l/4, columns 2 (l%4), +1), two panel buffers, the pivot factor.
struct __align__(16) TileSmem {
  double frag[MAXFRAGS * 64];
  double P[2][8 * MAXNB];  // finished panels J (parity J % 2), layout above (pchunk / pidx); phase 2 reuses them
  double dfL[64];         // pivot block: strictly lower part of L_D (row-major a * 8 + b)
  double dfinv[8];        // 1 / diag(L_D)
  double dfpiv[8];        // the pivots themselves (their square roots are taken off the critical chain)
  // Written by the NEXT rank of the cluster (DSMEM): the highest panel index whose landing buffer over there is free,
  // i.e. this CTA may push panel J into P[J & 1] of rank (J + 1) % CL once free_for >= J.
  int free_for;
  // mbarrier the PREVIOUS rank's bulk DSMEM copy of a panel into my panel buffer completes on (complete_tx): what the
  // next owner waits for instead of the cluster barrier. (Per-thread st.shared::cluster pushes: 2016 16-byte remote
  // stores cost 4.4-7k clk per panel; one cp.async.bulk moves the same 32 KB.)
  unsigned long long mbar;
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
__device__ __forceinline__ unsigned smem_addr(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
// one bulk copy: my shared memory -> the same place in a peer CTA's shared memory, completion (bytes) on the peer's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(unsigned dst_cluster, unsigned src_cta, unsigned bytes, unsigned mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(mbar_cluster)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(unsigned long long* bar, unsigned parity) {
  const unsigned addr = smem_addr(bar);
  for (unsigned spin = 0;; ++spin) {
    unsigned done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done)
      return;
    if (spin > (1u << 24))
      __trap();
  }
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
__device__ __forceinline__ void cluster_arrive_relaxed() {  // nothing to publish: no MEMBAR.ALL.GPU (~1k clk) in front
  asm volatile("barrier.cluster.arrive.relaxed;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ double rsqrt_nr(double a) {
  const double y = rsqrt(a);
  return fma(y * 0.5, fma(-a * y, y, 1.0), y);
}

// Measurement aid: clock64 stamps of CTA 0 / thread 0 per panel step (tools/gpu_diag_tile_test): 8 per step (+ 4 x 64 extra)
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
  // rows are dealt to the warps round-robin (I % NW), so a warp loads the A operands of a fragment row once
  auto update = [&](const double* __restrict__ pb, int Jp, int jl_lo, int jl_hi) {
    double b0[JL], b1[JL];
#pragma unroll
    for (int jl = 0; jl < JL; ++jl) {
      const int Jc = rank + CL * jl;
      const bool on = (jl >= jl_lo && jl <= jl_hi && Jc > Jp && Jc < nfc);
      const double2 b = on ? *reinterpret_cast<const double2*>(&pb[pchunk(8 * Jc + g, tig)]) : make_double2(0.0, 0.0);
      b0[jl] = b.x;
      b1[jl] = b.y;
    }
    // two fragment rows per trip: two independent load -> DMMA -> DMMA -> store chains in flight per warp
    // (measured and dropped: loading 8 fragments, then 16 DMMAs, then 8 stores per batch — slower, 3.3k vs 2.1k clk)
    for (int I = Jp + 1 + ((warp - (Jp + 1)) & (NW - 1)); I < nfc; I += 2 * NW) {  // I > Jp, I % NW == warp
      const int I2 = I + NW;
      const bool two = I2 < nfc;
      const double2 a = *reinterpret_cast<const double2*>(&pb[pchunk(8 * I + g, tig)]);
      const double2 a2 = two ? *reinterpret_cast<const double2*>(&pb[pchunk(8 * I2 + g, tig)]) : make_double2(0.0, 0.0);
      // software-pipelined over the local columns: the fragments of column jl+1 are loaded BEFORE those of column jl are
      // stored (the compiler keeps shared-memory loads behind earlier stores to the fragment array), so the LDS latency
      // overlaps the two dependent DMMAs of the current column
      auto act = [&](int jl) { const int Jc = rank + CL * jl; return jl < JL && jl >= jl_lo && jl <= jl_hi && Jc > Jp && Jc <= I2 && Jc < nfc; };
      auto ldf = [&](int jl, double2& c, double2& d) {
        const int Jc = rank + CL * jl;
        c = (Jc <= I) ? *reinterpret_cast<const double2*>(&sm.frag[(fo[jl] + I - Jc) * 64 + 2 * lane]) : make_double2(0.0, 0.0);
        d = two ? *reinterpret_cast<const double2*>(&sm.frag[(fo[jl] + I2 - Jc) * 64 + 2 * lane]) : make_double2(0.0, 0.0);
      };
      double2 c = make_double2(0.0, 0.0), d = c, cn = c, dn = c;
      if (act(0))
        ldf(0, c, d);
#pragma unroll
      for (int jl = 0; jl < JL; ++jl) {
        const bool cur = act(jl);         // warp-uniform
        if (jl + 1 < JL && act(jl + 1))
          ldf(jl + 1, cn, dn);
        if (cur) {
          const int Jc = rank + CL * jl;
          dmma884(c.x, c.y, -a.x, b0[jl]);
          dmma884(d.x, d.y, -a2.x, b0[jl]);
          dmma884(c.x, c.y, -a.y, b1[jl]);
          dmma884(d.x, d.y, -a2.y, b1[jl]);
          if (Jc <= I)
            *reinterpret_cast<double2*>(&sm.frag[(fo[jl] + I - Jc) * 64 + 2 * lane]) = c;
          if (two)
            *reinterpret_cast<double2*>(&sm.frag[(fo[jl] + I2 - Jc) * 64 + 2 * lane]) = d;
        }
        c = cn;
        d = dn;
      }
    }
  };

  // ================================ phase 1: the factor, panel by panel ================================
  // Hand-over of panel J:  owner J --(one bulk DSMEM copy completing on the receiver's mbarrier)--> next owner (critical path);
  //                        owner J --(global memory + cluster barrier J)--> everybody else (reads it back from L2).
  // Barrier phases: every CTA arrives at barrier J in iteration J; the next owner postpones its WAIT on barrier J until
  // just before it arrives at barrier J+1 (it does not need barrier J: its copy of panel J came by DSMEM).
  if (tid == 0) {
    sm.free_for = 1;  // the landing buffers of panels 0 and 1 have never been used
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sm.mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_arrive_release();
  cluster_wait_acquire();  // every CTA's flag / mbarrier is initialised before a peer may touch it
  for (int J = 0; J < nfc; ++J) {
    const bool own = (rank == J % CL);
    const bool has_next = (J + 1 < nfc);
    const int next_rank = (J + 1) % CL;
    const bool is_next = has_next && rank == next_rank;
    const int jlJ = J / CL;
    if (trace)
      trace[J * 8 + 0] = clock64();
    if (own) {
      double* pn = sm.P[J & 1];  // (holds panel J-2: dead)
      const double* pp = sm.P[(J - 1) & 1];
      // My block column J: apply panel J-1 (it is the critical path) and write the raw column straight into the buffer the
      // finished panel will occupy — the fragments of a factorised column are never needed again.
      {
        const int foJ = foff(jlJ);  // (fo[] stays statically indexed: registers, not local memory)
        double2 bJ = make_double2(0.0, 0.0);
        if (J > 0)
          bJ = *reinterpret_cast<const double2*>(&pp[pchunk(8 * J + g, tig)]);
        for (int I = J + ((warp - J) & (NW - 1)); I < nfc; I += NW) {
          double2 c = *reinterpret_cast<const double2*>(&sm.frag[(foJ + I - J) * 64 + 2 * lane]);
          if (J > 0) {
            const double2 a = *reinterpret_cast<const double2*>(&pp[pchunk(8 * I + g, tig)]);
            dmma884(c.x, c.y, -a.x, bJ.x);
            dmma884(c.x, c.y, -a.y, bJ.y);
          }
          pn[pidx(8 * I + g, 2 * tig)] = c.x;
          pn[pidx(8 * I + g, 2 * tig + 1)] = c.y;
        }
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
              D[a][b] = (a >= b) ? pn[pidx(8 * J + a, b)] : 0.0;
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
      // one panel row per thread: x <- x L_D^-T (forward substitution, columns left to right), in place in my own panel
      // buffer; the whole panel then goes to the NEXT owner's buffer with one bulk DSMEM copy
      const int r = 8 * (J + 1) + tid;
      double x[8];
      if (r < nbp) {
        double Lm[8][8], invd[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          invd[a] = sm.dfinv[a];
#pragma unroll
          for (int b = 0; b < 8; ++b)
            Lm[a][b] = (a > b) ? sm.dfL[a * 8 + b] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 v = *reinterpret_cast<const double2*>(&pn[pchunk(r, q)]);
          x[q] = v.x;
          x[4 + q] = v.y;
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          x[s] *= invd[s];
#pragma unroll
          for (int t = s + 1; t < 8; ++t)
            x[t] = fma(-x[s], Lm[t][s], x[t]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<double2*>(&pn[pchunk(r, q)]) = make_double2(x[q], x[4 + q]);
      }
      if (trace)
        trace[512 + J * 4 + 0] = clock64();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my rows are visible to the bulk-copy engine
      if (trace)
        trace[512 + J * 4 + 1] = clock64();
      __syncthreads();
      if (trace)
        trace[512 + J * 4 + 2] = clock64();
      if (tid == 0 && has_next) {
        // rows 8 (J+1) .. nbp-1 of the panel are contiguous: ONE bulk copy into the same place of the next owner's buffer
        double* first = &pn[8 * (J + 1) * 8];
        bulk_copy_to_peer(map_to_rank(first, static_cast<unsigned>(next_rank)), smem_addr(first),
                          static_cast<unsigned>((nbp - 8 * (J + 1)) * 8 * sizeof(double)),
                          map_to_rank(&sm.mbar, static_cast<unsigned>(next_rank)));
      }
      if (trace)
        trace[J * 8 + 4] = clock64();
      // ... and only now, off the critical path, into the tile in global memory: the result, and the medium the other
      // CTAs read the panel back from after the cluster barrier
      if (r < nbp) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          T[r + static_cast<long>(8 * J + k) * ldt] = x[k];  // coalesced along r
      }
      if (tid < 64) {  // the pivot block itself: strictly lower part from the factor, diagonal = correctly rounded sqrt
        const int a = tid >> 3, b = tid & 7;
        if (a > b)
          T[(8 * J + a) + static_cast<long>(8 * J + b) * ldt] = sm.dfL[a * 8 + b];
        else if (a == b)
          T[static_cast<long>(8 * J + a) * (1 + ldt)] = sqrt(sm.dfpiv[a]);
      }
      if (J > 0)
        cluster_wait_acquire();  // my postponed wait on barrier J-1 (I was the "next owner" of iteration J-1)
    }
    if (own)
      cluster_arrive_release();  // barrier J (release at cluster scope: my global stores are visible after the wait)
    else
      cluster_arrive_relaxed();
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
    if (is_next) {
      // next owner: panel J arrives by bulk DSMEM copy; barrier J is waited for later (see above)
      if (tid == 0) {
        const unsigned bytes = static_cast<unsigned>((nbp - 8 * (J + 1)) * 8 * sizeof(double));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&sm.mbar)), "r"(bytes) : "memory");
      }
      mbar_wait_parity(&sm.mbar, static_cast<unsigned>((J / CL) & 1));  // every thread observes the completion itself
    }
    else {
      cluster_wait_acquire();  // panel J is published in global memory
      if (!own && has_next) {
        const int r = 8 * (J + 1) + tid;
        if (r < nbp) {
          double* pn = sm.P[J & 1];
          double x[8];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            x[k] = __ldcg(T + r + static_cast<long>(8 * J + k) * ldt);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<double2*>(&pn[pchunk(r, q)]) = make_double2(x[q], x[4 + q]);
        }
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
