// Two-CTA (thread-block cluster) variant of the blocked diagonal-block kernel (potrf_block.cuh): Cholesky + inverse
// of one G x G diagonal block, spread over TWO SMs.
//
// Why: potrf_inv_blocked_kernel is bound by the fp64 FMA rate of one SM (64 FMA/clk on B200): per rank-BS step every
// thread applies 8*8*8 FMAs to its register block, 3.4k of the 4.7k clk of a step, 58 us per 128-block — the largest
// single item on the POTRF critical path (4 blocks per 512-tile, profiles/r01_chain_diag_tile_isolated.log). Here the
// 16 x 16 grid of register blocks is dealt out by block COLUMN parity: CTA c of the pair owns block columns tj with
// tj % 2 == c (16 x 8 blocks = 128 threads = one warp per SM sub-partition), which halves the FMA work per SM at every
// step of the sweep (column-cyclic, so the shrinking trailing part stays balanced). Panel column J lives entirely in
// CTA J % 2 — so does the thread that factors its pivot block — which therefore runs phases A1/A2 alone and
// then PUSHES the solved panel (8 KB) and the pivot rows' diagonal data into the peer's shared memory over DSMEM
// (st.shared::cluster); one cluster barrier per step publishes it. Panel buffers are double buffered on J & 1: the
// barrier of step J+1 is what guarantees that nobody still reads the buffer that step J+2 overwrites.
//
// The orchestration below is written once against a small context policy (cta / thread index, local shared-memory
// pointers, cta_sync, cluster_sync, push-to-peer): potrf_tile.cu instantiates it with PTX (mapa, st.shared::cluster,
// barrier.cluster), tools/potrf_cluster_emu.cu with 256 host threads and std::barrier — the SAME code, checked on the
// CPU against host loops (there is no GPU in the build container).
// Status: selected with DLAF_B200_POTRF_KERNEL=cluster2. One GPU run so far (tools/potrf_variant_probe.py): correct at
// once, 326 vs 342 us per 512-tile — less than the model above predicts, so the default stays the single-CTA kernel
// until the per-phase clocks of this one have been looked at.
#pragma once

#include "potrf_block.cuh"

namespace dlaf_b200 {
namespace pblock {

constexpr int kClusterCtas = 2;
constexpr int kClusterThreads = 128;  // per CTA: 16 block rows x 8 block columns

// Ctx requirements:
//   int cta, tid;  T* panel[2];  R *dd, *dinv, *dfinv, *dfsq;  T *dfL, *msc;  int* sfail (4 ints);   (local shared memory)
//   void cta_sync(); void cluster_sync(); template <class V> void push(V* local_address, V value);
// Failure flags (sfail[0..3]): cand[p] = sfail[p] is the verdict of the pivot block of a step with parity p, written by
// the ONE thread that factors that block (at the end of the step before, inside the CTA that owns the step); pub[p] =
// sfail[2 + p] is what the owning CTA publishes to both CTAs BEFORE the step's cluster barrier and what everybody
// tests after it. Two slots each because a fast CTA may already write the flags of step J+1 while a slow thread of
// the other one has not yet tested those of step J (the emulator found exactly that race with a single flag).
template <class C, class T, class Ctx>
PB_HD void potrf_inv_cluster2_body(Ctx& cx, T* Tm, long ldt, T* W, long ldw, int* info, int info_offset) {
  constexpr int BS = C::BS, PB = C::PB, NT = C::NT;
  const int ti = cx.tid % NT, tj = kClusterCtas * (cx.tid / NT) + cx.cta;
  int* cand = cx.sfail;
  int* pub = cx.sfail + 2;
  T reg[BS][BS];
  load_block<C, T>(reg, Tm, ldt, ti, tj);
  if (cx.tid == 0)
    cand[0] = cand[1] = pub[0] = pub[1] = 0;
  cx.cluster_sync();
  if (ti == 0 && tj == 0)  // CTA 0
    cand[0] = factor_pivot_block<C, T>(reg, cx.dfL, cx.dfinv, cx.dfsq);
  int fail = 0;
  for (int J = 0; J < NT; ++J) {
    T* panel = cx.panel[J & 1];
    if (cx.cta == J % kClusterCtas) {
      // A1: the owners of block column J publish it (k-major, padded) to the shared memory of THIS CTA
      if (tj == J)
        write_panel<C, T>(reg, panel, ti);
      cx.cta_sync();  // panel J and the factor of its pivot block (written by a thread of this CTA) are visible
      const int f = cand[J & 1];
      if (cx.tid == 0) {
        pub[J & 1] = f;
        cx.push(pub + (J & 1), f);
      }
      if (f == 0 && cx.tid < PB) {
        // A2: one thread per panel row, then the row goes to the peer
        const int r = cx.tid;
        solve_panel_row<C, T>(panel, cx.dfL, cx.dfinv, cx.dfsq, cx.msc, cx.dd, cx.dinv, J, r);
#pragma unroll
        for (int k = 0; k < BS; ++k) {
          T* p = panel + k * C::PROW + C::poff(r);
          cx.push(p, *p);
        }
        if (r / BS == J) {
          cx.push(cx.dd + r, cx.dd[r]);
          cx.push(cx.dinv + r, cx.dinv[r]);
        }
      }
    }
    cx.cluster_sync();  // panel J (or the failure verdict) complete in both CTAs
    fail = pub[J & 1];
    if (fail)
      break;
    // B: rank-BS update of the register blocks, then the owner of the next pivot block factors it (in the CTA that
    // runs A1/A2 of step J+1) while everybody else is still in the update
    update_block<C, T>(reg, panel, cx.dinv, J, ti, tj);
    if (J + 1 < NT && ti == J + 1 && tj == J + 1) {
      const int f = factor_pivot_block<C, T>(reg, cx.dfL, cx.dfinv, cx.dfsq);
      cand[(J + 1) & 1] = f ? (J + 1) * BS + f : 0;
    }
  }
  if (fail) {
    if (cx.cta == 0 && cx.tid == 0) {
#ifdef __CUDA_ARCH__
      atomicCAS(info, 0, info_offset + fail);
#else
      if (*info == 0)
        *info = info_offset + fail;
#endif
    }
    for (int idx = cx.cta * kClusterThreads + cx.tid; idx < PB * PB; idx += kClusterCtas * kClusterThreads)
      W[(idx % PB) + static_cast<long>(idx / PB) * ldw] = make_real<T>(0);
    return;
  }
  store_block<C, T>(reg, Tm, ldt, W, ldw, cx.dd, cx.dinv, ti, tj);
}

// shared-memory footprint (elements of the pieces, in this order) — used by the kernel and by the emulator
template <class C, class T>
struct ClusterSmem {
  using R = base_t<T>;
  static constexpr size_t bytes = 2 * C::PANEL_ELEMS * sizeof(T) + (2 * C::PB + 2 * C::BS) * sizeof(R) +
                                  2 * C::BS * C::BS * sizeof(T) + 16;
  template <class Ctx>
  PB_HD static void carve(Ctx& cx, unsigned char* base) {
    T* p = reinterpret_cast<T*>(base);
    cx.panel[0] = p;
    cx.panel[1] = p + C::PANEL_ELEMS;
    R* r = reinterpret_cast<R*>(p + 2 * C::PANEL_ELEMS);
    cx.dd = r;
    cx.dinv = r + C::PB;
    cx.dfinv = cx.dinv + C::PB;
    cx.dfsq = cx.dfinv + C::BS;
    cx.dfL = reinterpret_cast<T*>(cx.dfsq + C::BS);
    cx.msc = cx.dfL + C::BS * C::BS;
    cx.sfail = reinterpret_cast<int*>(cx.msc + C::BS * C::BS);
  }
};

}  // namespace pblock
}  // namespace dlaf_b200
