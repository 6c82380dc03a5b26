// The one-launch-per-step update  C = C + alpha A B^H  of the sweeps that follow POTRF (inverse_engine.cu, hegst_engine.cu)
// on the engine of the element type: fp64 -> exact int8 digit planes on tcgen05 + guarded native fallback
// (gemm_ozaki.h), fp32 -> 3xTF32 on tcgen05 (gemm_tf32.h), complex -> native DMMA / SIMT kernels (gemm_args.h).
#pragma once

#include <cuda_runtime.h>

#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>

#include "common.h"
#include "gemm_args.h"
#include "gemm_ozaki.h"
#include "gemm_tf32.h"
#include "pool.h"

namespace dlaf_b200 {

// One operand of a bulk update: `rows` rows x nbp columns, either plain column-major (tile_stride == 0, leading
// dimension ld) or tile-contiguous (nbp x nbp tiles with leading dimension ld = nbp, tile_stride elements apart).
template <class T>
struct Operand {
  const T* x;
  long ld;
  long rows;
  long tile_stride;
};

template <class T>
struct BulkUpdate {
  static constexpr int kSlots = 2;
  bool oz = false, tf = false;
  int slots = 1;
  OzakiSplit oa[kSlots], ob[kSlots];
  Tf32Split ta[kSlots], tb[kSlots];
  int* flags = nullptr;
  int nflags = 0, used = 0;
  int* cur = nullptr;  // guard flag of the current step
  OzakiSplit ox;       // one extra small B-side operand per step (a single tile: the diagonal tile of the reduction)
  Tf32Split tx;

  // rows_a / rows_b: largest number of rows of an A-side / B-side operand; nslots operands per side and step
  void init(long rows_a, long rows_b, int nbp, int nsteps, cudaStream_t s, int nslots = 1) {
    slots = nslots;
    if constexpr (std::is_same_v<T, double>) {
      const char* e = std::getenv("DLAF_B200_D_BULK");
      oz = (e == nullptr || std::string(e) == "ozaki") && nbp <= 512 && rows_a > 0 && rows_b > 0;
      if (oz) {
        for (int i = 0; i < slots; ++i) {
          oa[i].allocate(rows_a, nbp);
          ob[i].allocate(rows_b, nbp);
        }
        nflags = nsteps;
        flags = pool_alloc<int>(nflags);
        DLAF_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(int) * nflags, s));
      }
    }
    if constexpr (std::is_same_v<T, float>) {
      tf = std::getenv("DLAF_B200_S_SIMT") == nullptr && rows_a > 0 && rows_b > 0;
      if (tf) {
        for (int i = 0; i < slots; ++i) {
          ta[i].allocate(rows_a, nbp);
          tb[i].allocate(rows_b, nbp);
        }
      }
    }
    (void) rows_a, (void) rows_b, (void) nbp, (void) nsteps, (void) s;
  }

  void init_extra(long rows, int nbp) {
    if (oz)
      ox.allocate(rows, nbp);
    if (tf)
      tx.allocate(rows, nbp);
  }
  long split_extra(const Operand<T>& o, int nbp, cudaStream_t s) {
    if constexpr (std::is_same_v<T, double>) {
      if (oz) {
        ox.split(o.x, o.ld, o.rows, s, o.tile_stride ? nbp : 0, o.tile_stride, cur);
        return 1;
      }
    }
    if constexpr (std::is_same_v<T, float>) {
      if (tf) {
        tx.split(o.x, o.ld, o.rows, s, o.tile_stride ? nbp : 0, o.tile_stride);
        return 1;
      }
    }
    (void) o, (void) nbp, (void) s;
    return 0;
  }
  // C += alpha A B^H with A = slot ia of the A side and B = the extra operand (g.alpha = +-2^e for fp64)
  long gemm_extra(GemmArgsT<T> g, const Operand<T>& a, int ia, const Operand<T>& b, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0)
      return 0;
    g.A = a.x;
    g.lda = a.ld;
    g.a_ts = a.tile_stride;
    g.B = b.x;
    g.ldb = b.ld;
    g.b_ts = b.tile_stride;
    g.beta = 1.0;
    if constexpr (std::is_same_v<T, double>) {
      if (oz) {
        launch_gemm_ozaki_i8(g, oa[ia], 0, ox, 0, s, 0, cur);
        launch_gemm_nt_f64_if(g, cur, s);
        return 2;
      }
    }
    if constexpr (std::is_same_v<T, float>) {
      if (tf) {
        launch_gemm_tf32x3(g, ta[ia], 0, tx, 0, s);
        return 1;
      }
    }
    (void) ia;
    launch_gemm_nt<T>(g, s);
    return 1;
  }

  // ---- step-wise interface: begin_step, split the operands of the step once, then any number of products
  void begin_step() {
    if (oz) {
      DLAF_B200_ASSERT(used < nflags, "guard flags exhausted");
      cur = flags + used++;
    }
  }
  long split(bool b_side, int slot, const Operand<T>& o, int nbp, cudaStream_t s) {
    if (o.rows <= 0)
      return 0;
    if constexpr (std::is_same_v<T, double>) {
      if (oz) {
        (b_side ? ob[slot] : oa[slot]).split(o.x, o.ld, o.rows, s, o.tile_stride ? nbp : 0, o.tile_stride, cur);
        return 1;
      }
    }
    if constexpr (std::is_same_v<T, float>) {
      if (tf) {
        (b_side ? tb[slot] : ta[slot]).split(o.x, o.ld, o.rows, s, o.tile_stride ? nbp : 0, o.tile_stride);
        return 1;
      }
    }
    (void) b_side, (void) slot, (void) nbp, (void) s;
    return 0;
  }
  // g: C, ldc, M, N, K, alpha, mask geometry filled in. A = operand a (split: slot ia of the A side), B = operand b
  // (split: slot ib of the B side, or of the A side when b_on_a_side — the 1 x 1 grid, where both are the same panels).
  long gemm(GemmArgsT<T> g, const Operand<T>& a, int ia, const Operand<T>& b, int ib, bool b_on_a_side, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0)
      return 0;
    g.A = a.x;
    g.lda = a.ld;
    g.a_ts = a.tile_stride;
    g.B = b.x;
    g.ldb = b.ld;
    g.b_ts = b.tile_stride;
    g.beta = 1.0;
    if constexpr (std::is_same_v<T, double>) {
      if (oz) {
        launch_gemm_ozaki_i8(g, oa[ia], 0, b_on_a_side ? oa[ib] : ob[ib], 0, s, 0, cur);
        launch_gemm_nt_f64_if(g, cur, s);
        return 2;
      }
    }
    if constexpr (std::is_same_v<T, float>) {
      if (tf) {
        launch_gemm_tf32x3(g, ta[ia], 0, b_on_a_side ? ta[ib] : tb[ib], 0, s);
        return 1;
      }
    }
    (void) ia, (void) ib, (void) b_on_a_side;
    launch_gemm_nt<T>(g, s);
    return 1;
  }

  // ---- one product per step. same: B is the same panel as A.
  long run(const GemmArgsT<T>& g, const Operand<T>& a, const Operand<T>& b, bool same, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0)
      return 0;
    begin_step();
    long n = split(false, 0, a, g.nbp, s);
    if (!same)
      n += split(true, 0, b, g.nbp, s);
    return n + gemm(g, a, 0, b, 0, same, s);
  }

  // number of steps whose guard fired (synchronises the stream); releases everything
  int finish(cudaStream_t s) {
    int fired = 0;
    if (oz) {
      DLAF_CUDA_CHECK(cudaStreamSynchronize(s));
      if (used > 0) {
        std::vector<int> h(used);
        DLAF_CUDA_CHECK(cudaMemcpy(h.data(), flags, sizeof(int) * used, cudaMemcpyDeviceToHost));
        for (int v : h)
          fired += (v != 0);
      }
      for (int i = 0; i < slots; ++i) {
        oa[i].release();
        ob[i].release();
      }
      ox.release();
      pool_free(flags);
    }
    if (tf) {
      DLAF_CUDA_CHECK(cudaStreamSynchronize(s));
      for (int i = 0; i < slots; ++i) {
        ta[i].release();
        tb[i].release();
      }
      tx.release();
    }
    return fired;
  }
};

}  // namespace dlaf_b200
