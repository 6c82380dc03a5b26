// Register-resident blocked Cholesky + triangular inverse of one G x G diagonal block (G = 128 real,
// 64 complex) by a single CTA of 256 threads — the per-thread phase functions. They are
// __host__ __device__ so that tools/potrf_block_emu.cu can run the very same code for all 256 "threads"
// on the CPU (there is no GPU in the build container).
//
// Data layout: the block S (lower: A -> L, strictly upper: X = inv(L)^H under construction, see
// potrf_tile.cu for the rank-1 form of the rule) is spread over the register files: thread (ti, tj) of a
// 16 x 16 grid owns the BS x BS sub-block rows ti*BS.., columns tj*BS.. (BS = G/16).
// Per panel step J (BS columns at a time):
//   A1  owners of block column J publish it to shared memory (k-major, padded per row block)
//   A2  one thread per row solves its own panel row x <- x L_D^-H against the factor of the BS x BS pivot
//       block D (in shared memory); the BS pivot rows instead receive L_D and X_D = inv(L_D)^H
//   B   every thread applies the rank-BS update to its register block:
//         S(r,s) -= sum_k P(r,k) conj(P(s,k))   for block columns right of J, rows r >= s or r < J*BS
//       and the pivot block row gets the freshly created X entries  -(X_D-weighted combination);
//       the ONE thread owning the next pivot block then factorises it (a latency-bound chain of rsqrt and
//       dependent FMAs) while all the others are still in their update — measured: with every thread
//       re-deriving the pivot factor inside A2 that chain was 55 % of the kernel time.
#pragma once

#include <cuda_runtime.h>

#include <cmath>

#include "types.h"

namespace dlaf_b200 {
namespace pblock {

#define PB_HD __host__ __device__ __forceinline__

PB_HD float re_of(float v) { return v; }
PB_HD double re_of(double v) { return v; }
PB_HD float re_of(float2 v) { return v.x; }
PB_HD double re_of(double2 v) { return v.x; }

PB_HD float scale_r(float v, float s) { return v * s; }
PB_HD double scale_r(double v, double s) { return v * s; }
PB_HD float2 scale_r(float2 v, float s) { return make_float2(v.x * s, v.y * s); }
PB_HD double2 scale_r(double2 v, double s) { return make_double2(v.x * s, v.y * s); }

// v - a * conj(b)
PB_HD float sub_mul_conj(float v, float a, float b) { return fmaf(-a, b, v); }
PB_HD double sub_mul_conj(double v, double a, double b) { return fma(-a, b, v); }
PB_HD float2 sub_mul_conj(float2 v, float2 a, float2 b) {
  return make_float2(fmaf(-a.y, b.y, fmaf(-a.x, b.x, v.x)), fmaf(a.x, b.y, fmaf(-a.y, b.x, v.y)));
}
PB_HD double2 sub_mul_conj(double2 v, double2 a, double2 b) {
  return make_double2(fma(-a.y, b.y, fma(-a.x, b.x, v.x)), fma(a.x, b.y, fma(-a.y, b.x, v.y)));
}
// v - a * b
PB_HD float sub_mul(float v, float a, float b) { return fmaf(-a, b, v); }
PB_HD double sub_mul(double v, double a, double b) { return fma(-a, b, v); }
PB_HD float2 sub_mul(float2 v, float2 a, float2 b) {
  return make_float2(fmaf(a.y, b.y, fmaf(-a.x, b.x, v.x)), fmaf(-a.x, b.y, fmaf(-a.y, b.x, v.y)));
}
PB_HD double2 sub_mul(double2 v, double2 a, double2 b) {
  return make_double2(fma(a.y, b.y, fma(-a.x, b.x, v.x)), fma(-a.x, b.y, fma(-a.y, b.x, v.y)));
}

PB_HD float fast_rsqrt(float a) {
#ifdef __CUDA_ARCH__
  float y = rsqrtf(a);
  return fmaf(y * 0.5f, fmaf(-a * y, y, 1.0f), y);
#else
  return 1.0f / std::sqrt(a);
#endif
}
PB_HD double fast_rsqrt(double a) {
#ifdef __CUDA_ARCH__
  double y = rsqrt(a);
  return fma(y * 0.5, fma(-a * y, y, 1.0), y);
#else
  return 1.0 / std::sqrt(a);
#endif
}
PB_HD float full_sqrt(float a) { return sqrtf(a); }
PB_HD double full_sqrt(double a) { return sqrt(a); }

template <class T, int PB_>
struct Cfg {
  static constexpr int PB = PB_;
  static constexpr int NT = 16;        // thread grid edge
  static constexpr int BS = PB / NT;   // register block edge = panel width
  static constexpr int PAD = (16 / sizeof(T)) > 0 ? (16 / sizeof(T)) : 1;  // 16 bytes per row block
  static constexpr int RS = BS + PAD;  // padded row-block stride: spreads the 16 row blocks over the banks
  static constexpr int PROW = NT * RS;
  static constexpr int PANEL_ELEMS = BS * PROW;
  PB_HD static int poff(int r) { return (r / BS) * RS + (r % BS); }
};

template <class C, class T>
PB_HD void load_block(T (&reg)[C::BS][C::BS], const T* Tm, long ldt, int ti, int tj) {
#pragma unroll
  for (int b = 0; b < C::BS; ++b)
#pragma unroll
    for (int a = 0; a < C::BS; ++a) {
      const int r = ti * C::BS + a, s = tj * C::BS + b;
      reg[a][b] = (r >= s) ? Tm[r + s * ldt] : make_real<T>(0);
    }
}

// A1 (threads with tj == J)
template <class C, class T>
PB_HD void write_panel(const T (&reg)[C::BS][C::BS], T* panel, int ti) {
#pragma unroll
  for (int b = 0; b < C::BS; ++b)
#pragma unroll
    for (int a = 0; a < C::BS; ++a)
      panel[b * C::PROW + C::poff(ti * C::BS + a)] = reg[a][b];
}

// Pivot-block factorization, executed by ONE thread: the owner of the diagonal register block (ti == tj == J)
// right after that block received its last update — i.e. overlapped with the rank-BS update every other
// thread is still busy with. Publishes L_D (scaled strictly-lower part), 1/diag and diag^2 to shared memory.
// Returns 0 or the 1-based in-block column of a non-positive pivot.
template <class C, class T>
PB_HD int factor_pivot_block(const T (&reg)[C::BS][C::BS], T* dfL, base_t<T>* dfinv, base_t<T>* dfsq) {
  using R = base_t<T>;
  constexpr int BS = C::BS;
  T Dm[BS][BS];
#pragma unroll
  for (int b = 0; b < BS; ++b)
#pragma unroll
    for (int a = 0; a < BS; ++a)
      Dm[a][b] = (a >= b) ? reg[a][b] : make_real<T>(0);
  int fail = 0;
#pragma unroll
  for (int j = 0; j < BS; ++j) {
    const R ajj = re_of(Dm[j][j]);
    if (!(ajj > R(0)) && fail == 0)
      fail = j + 1;
    const R inv = fast_rsqrt(ajj);
    dfinv[j] = inv;
    dfsq[j] = ajj;
#pragma unroll
    for (int a = j + 1; a < BS; ++a) {
      Dm[a][j] = scale_r(Dm[a][j], inv);
      dfL[a * BS + j] = Dm[a][j];
    }
#pragma unroll
    for (int s = j + 1; s < BS; ++s)
#pragma unroll
      for (int a = s; a < BS; ++a)
        Dm[a][s] = sub_mul_conj(Dm[a][s], Dm[a][j], Dm[s][j]);
  }
  return fail;
}

// A2 (threads t < PB, one per panel row): x <- x L_D^-H with the pivot factor read from shared memory.
// Every thread reads and writes only its own row, the pivot factor lives in its own buffer: no barrier
// inside this phase.
template <class C, class T>
PB_HD void solve_panel_row(T* panel, const T* dfL, const base_t<T>* dfinv, const base_t<T>* dfsq, T* msc,
                           base_t<T>* dd, base_t<T>* dinv, int J, int t) {
  using R = base_t<T>;
  constexpr int BS = C::BS;
  const int r = t, rb = r / BS, rr = r % BS;
  T Lm[BS][BS];
  R invd[BS];
#pragma unroll
  for (int a = 0; a < BS; ++a) {
    invd[a] = dfinv[a];
#pragma unroll
    for (int b = 0; b < BS; ++b)
      Lm[a][b] = (a > b) ? dfL[a * BS + b] : make_real<T>(0);
  }
  T x[BS];
  if (rb != J) {
#pragma unroll
    for (int k = 0; k < BS; ++k)
      x[k] = panel[k * C::PROW + C::poff(r)];
    // Cholesky rows below the pivot block AND inverse rows above it obey the same rule
#pragma unroll
    for (int s = 0; s < BS; ++s) {
      T v = x[s];
#pragma unroll
      for (int j = 0; j < s; ++j)
        v = sub_mul_conj(v, x[j], Lm[s][j]);
      x[s] = scale_r(v, invd[s]);
    }
  }
  else {
    // pivot row rr: [ L_D(rr, 0..rr-1) | diag | X_D(rr, s) = conj(M_D(s, rr)), s > rr ],  M_D = inv(L_D).
    // All BS columns of M_D in registers with compile-time indices: the columns are independent chains, so
    // the FP64 pipe latency (8 clk per dependent op) is hidden by instruction-level parallelism across them;
    // a run-time-bounded loop over shared memory for the single needed column measured ~3k clk per step.
    (void) msc;
    T M[BS][BS];
#pragma unroll
    for (int c = 0; c < BS; ++c) {
      M[c][c] = make_real<T>(invd[c]);
#pragma unroll
      for (int s = c + 1; s < BS; ++s) {
        T v = make_real<T>(0);
#pragma unroll
        for (int j = c; j < s; ++j)
          v = sub_mul(v, Lm[s][j], M[j][c]);
        M[s][c] = scale_r(v, invd[s]);
      }
    }
    R dsel = R(0), isel = R(0);
#pragma unroll
    for (int a = 0; a < BS; ++a)
      if (a == rr) {
        dsel = dfsq[a];
        isel = invd[a];
      }
    const R d = full_sqrt(dsel);
#pragma unroll
    for (int s = 0; s < BS; ++s) {
      T v = make_real<T>(0);
#pragma unroll
      for (int a = 0; a < BS; ++a) {
        if (a == rr && s < a)
          v = Lm[a][s];              // L_D(rr, s)
        if (a == rr && s > a)
          v = conj_val(M[s][a]);     // X_D(rr, s) = conj(M_D(s, rr))
      }
      x[s] = (s == rr) ? make_real<T>(d) : v;
    }
    dd[r] = d;
    dinv[r] = isel;
  }
#pragma unroll
  for (int k = 0; k < BS; ++k)
    panel[k * C::PROW + C::poff(r)] = x[k];
}

// B (all threads)
template <class C, class T>
PB_HD void update_block(T (&reg)[C::BS][C::BS], const T* panel, const base_t<T>* dinv, int J, int ti, int tj) {
  using R = base_t<T>;
  constexpr int BS = C::BS;
  if (tj < J)
    return;
  if (tj == J) {  // take the final panel values back into the registers
#pragma unroll
    for (int b = 0; b < BS; ++b)
#pragma unroll
      for (int a = 0; a < BS; ++a)
        reg[a][b] = panel[b * C::PROW + C::poff(ti * BS + a)];
    return;
  }
  const bool full = (ti > tj) || (ti < J);
  const bool diag = (ti == tj);
  const bool piv = (ti == J);
  if (!(full || diag || piv))
    return;  // rows strictly between the pivot block and the column's own diagonal block: still zero
  // One code path for all three cases (every warp holds a lane of each kind, a branch here would make every
  // warp execute the update twice). The pivot block row (ti == J) receives the freshly created
  //   X(r,s) = -( sum_{k>a} X_D(a,k) conj(P(s,k)) + dinv_a conj(P(s,a)) ),
  // which is the general rank-BS update applied to a zero block with the row operand patched:
  //   rk'[a] = X_D(a,k) for k > a,  dinv_a for k == a,  0 for k < a   (its registers are still zero).
  R dv[BS];
#pragma unroll
  for (int a = 0; a < BS; ++a)
    dv[a] = piv ? dinv[J * BS + a] : R(0);
#pragma unroll
  for (int k = 0; k < BS; ++k) {
    T rk[BS], ck[BS];
#pragma unroll
    for (int a = 0; a < BS; ++a) {
      rk[a] = panel[k * C::PROW + C::poff(ti * BS + a)];
      if (piv)
        rk[a] = (k > a) ? rk[a] : (k == a ? make_real<T>(dv[a]) : make_real<T>(0));
    }
#pragma unroll
    for (int b = 0; b < BS; ++b)
      ck[b] = panel[k * C::PROW + C::poff(tj * BS + b)];
#pragma unroll
    for (int a = 0; a < BS; ++a)
#pragma unroll
      for (int b = 0; b < BS; ++b)
        if (!diag || a >= b)
          reg[a][b] = sub_mul_conj(reg[a][b], rk[a], ck[b]);
  }
}

// final store: L into the lower triangle of T (never touching the strictly upper part), inv(L) into W
template <class C, class T>
PB_HD void store_block(const T (&reg)[C::BS][C::BS], T* Tm, long ldt, T* W, long ldw, const base_t<T>* dd,
                       const base_t<T>* dinv, int ti, int tj) {
  constexpr int BS = C::BS;
#pragma unroll
  for (int b = 0; b < BS; ++b)
#pragma unroll
    for (int a = 0; a < BS; ++a) {
      const int r = ti * BS + a, s = tj * BS + b;
      if (r > s) {
        Tm[r + s * ldt] = reg[a][b];
        W[s + r * ldw] = make_real<T>(0);  // strictly upper part of W
      }
      else if (r == s) {
        Tm[r + s * ldt] = make_real<T>(dd[r]);
        W[r + s * ldw] = make_real<T>(dinv[r]);
      }
      else {
        // r < s: this register holds X(r, s) = conj(inv(L)(s, r))
        W[s + r * ldw] = conj_val(reg[a][b]);
      }
    }
}

}  // namespace pblock
}  // namespace dlaf_b200
