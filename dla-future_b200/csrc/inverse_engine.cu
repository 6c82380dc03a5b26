// See inverse_engine.h.
#include "inverse_engine.h"

#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>

#include "comm.h"
#include "common.h"
#include "distribution.h"
#include "pool.h"
#include "bulk_update.cuh"
#include "tri_kernels.cuh"

namespace dlaf_b200 {

namespace {

using namespace trik;

}  // namespace

template <class T>
long inverse_device(const InverseProblem& p, int phases, T* a_user, long lda, ncclComm_t row_comm, ncclComm_t col_comm,
                    cudaStream_t s, int* guard_steps) {
  using NT = NcclType<T>;
  constexpr int G = Gran<T>::value;
  long launches = 0;
  if (guard_steps)
    *guard_steps = 0;
  const bool transposed = (p.uplo == 'U' || p.uplo == 'u');
  DLAF_B200_ASSERT(transposed || p.uplo == 'L' || p.uplo == 'l', "uplo must be L or U");
  const bool unit = (p.diag == 'U' || p.diag == 'u');
  const bool do_trtri = (phases & kTriangularInverse) != 0, do_assemble = (phases & kAssembleFromInverseFactor) != 0;
  DLAF_B200_ASSERT(!(unit && do_assemble), "the Cholesky factor has a non-unit diagonal");
  if (p.n == 0)
    return 0;
  const int nbp = static_cast<int>(round_up(p.nb, G));
  const int ns = nbp / G;
  const int nt = ceil_div(p.n, p.nb);
  const size_t tsz = static_cast<size_t>(nbp) * nbp, wsz = static_cast<size_t>(ns) * G * G;
  // ---- engine grid: the lower-triangular problem; uplo == 'U' works on A^H with the grid roles swapped
  const int Pe = transposed ? p.Q : p.P, Qe = transposed ? p.P : p.Q;
  const int erow = transposed ? p.pcol : p.prow, ecol = transposed ? p.prow : p.pcol;
  ncclComm_t e_row_comm = transposed ? col_comm : row_comm;  // ranks of my ENGINE row (size Qe)
  ncclComm_t e_col_comm = transposed ? row_comm : col_comm;  // ranks of my ENGINE column (size Pe)
  const int e_src_in_col = transposed ? p.src_col : p.src_row, e_src_in_row = transposed ? p.src_row : p.src_col;
  auto col_rank = [&](int v_erow) { return (v_erow + e_src_in_col) % Pe; };
  auto row_rank = [&](int v_ecol) { return (v_ecol + e_src_in_row) % Qe; };
  DLAF_B200_ASSERT(Pe == 1 || e_col_comm != nullptr, "communicator required");
  DLAF_B200_ASSERT(Qe == 1 || e_row_comm != nullptr, "communicator required");

  const int ltr = cnt(nt, erow, Pe), ltc = cnt(nt, ecol, Qe);
  const long lds = static_cast<long>(ltr > 0 ? ltr : 1) * nbp;
  const bool have = ltr > 0 && ltc > 0;
  T* slab = nullptr;
  if (have) {
    slab = pool_alloc<T>(lds * ltc * nbp);
    DLAF_CUDA_CHECK(cudaMemsetAsync(slab, 0, sizeof(T) * lds * ltc * nbp, s));
    dim3 grid(nbp / 32, nbp / 32, ltr * ltc), block(32, 8);
    inv_convert_kernel<T, true><<<grid, block, 0, s>>>(a_user, lda, slab, lds, p.n, p.nb, nbp, Pe, Qe, erow, ecol, ltr, transposed,
                                                       unit, true);
    DLAF_CUDA_CHECK(cudaGetLastError());
    ++launches;
  }
  auto tile = [&](long gi, long gj) { return slab + (gi / Pe) * nbp + (gj / Qe) * nbp * lds; };  // my stored tile (gi, gj)
  auto pack = [&](const T* src, long ld_src, T* dst, long ld_dst, int ntiles, long src_stride, long dst_stride, bool tr, bool cj,
                  bool neg) {
    if (ntiles <= 0)
      return;
    dim3 grid(nbp / 32, nbp / 32, ntiles), block(32, 8);
    trsm_pack_tile_kernel<T><<<grid, block, 0, s>>>(src, ld_src, dst, nbp, tr, cj, src_stride, dst_stride, ld_dst, neg);
    DLAF_CUDA_CHECK(cudaGetLastError());
    ++launches;
  };
  const bool is_complex = sizeof(T) == 2 * sizeof(base_t<T>);

  // ---- workspaces
  T *colp = nullptr, *panA = nullptr, *panB = nullptr, *dbuf = nullptr, *dloc = nullptr;
  colp = pool_alloc<T>(static_cast<size_t>(ltr > 0 ? ltr : 1) * nbp * nbp);
  panB = pool_alloc<T>(tsz * (ltc > 0 ? ltc : 1));
  if (do_assemble && Qe > 1)
    panA = pool_alloc<T>(tsz * (ltr > 0 ? ltr : 1));
  dbuf = pool_alloc<T>(tsz);
  BulkUpdate<T> bulk;
  bulk.init(static_cast<long>(ltr) * nbp, static_cast<long>(ltc) * nbp, nbp, 3 * nt + 2, s);
  bulk.init_extra(nbp, nbp);

  // =================================================================================================================
  // (1) W = L^-1
  if (do_trtri) {
    // my diagonal tiles: [L_kk packed | inverted G-blocks | Wh_k = L_kk^-H], all up front (off the critical path)
    std::vector<int> my_diag;
    for (int k = 0; k < nt; ++k)
      if (k % Pe == erow && k % Qe == ecol)
        my_diag.push_back(k);
    const size_t dsz = 2 * tsz + wsz;
    if (!my_diag.empty()) {
      dloc = pool_alloc<T>(dsz * my_diag.size());
      for (size_t i = 0; i < my_diag.size(); ++i)
        pack(tile(my_diag[i], my_diag[i]), lds, dloc + dsz * i, nbp, 1, 0, 0, false, false, false);
      static bool configured = false;
      if (!configured) {
        DLAF_CUDA_CHECK(cudaFuncSetAttribute(trsm_trtri_blocks_kernel<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(T) * G * (G + 1))));
        configured = true;
      }
      dim3 gridw(ns, static_cast<unsigned>(my_diag.size()));
      trsm_trtri_blocks_kernel<T, G><<<gridw, G, sizeof(T) * G * (G + 1), s>>>(dloc, static_cast<long>(dsz), nbp, dloc + tsz,
                                                                              static_cast<long>(dsz), true);
      DLAF_CUDA_CHECK(cudaGetLastError());
      dim3 gridi(nbp, static_cast<unsigned>(my_diag.size()));
      inv_identity_kernel<T><<<gridi, 128, 0, s>>>(dloc + tsz + wsz, nbp, static_cast<long>(dsz));
      DLAF_CUDA_CHECK(cudaGetLastError());
      launches += 2;
      for (size_t i = 0; i < my_diag.size(); ++i)  // I L_kk^-H by the panel substitution
        launches += solve_rows_against_tile<T>(dloc + dsz * i + tsz + wsz, nbp, nbp, dloc + dsz * i, nbp, dloc + dsz * i + tsz, ns,
                                               true, s);
    }
    for (int k = nt - 1; k >= 0; --k) {
      const int owner_r = k % Pe, owner_c = k % Qe;
      const bool in_col = (ecol == owner_c), in_row = (erow == owner_r);
      const int li_k = cnt(k, erow, Pe), li_k1 = cnt(k + 1, erow, Pe);  // my first local row tile with global index >= k, > k
      const int ncols = cnt(k, ecol, Qe);                                // my local block columns left of k
      const long mk = static_cast<long>(ltr - li_k) * nbp;               // rows of the extended column panel on this rank
      const long vrows = static_cast<long>(ltr - li_k1) * nbp;           // of which strictly below row k
      if (in_col) {  // (a rank of this column without local rows still takes part in the broadcast)
        const T* whk = nullptr;
        if (in_row) {
          size_t idx = 0;
          while (my_diag[idx] != k)
            ++idx;
          whk = dloc + dsz * idx + tsz + wsz;
        }
        if (k < nt - 1 && Pe > 1) {
          DLAF_NCCL_CHECK(ncclBroadcast(in_row ? whk : dbuf, dbuf, tsz * NT::mult, NT::value, col_rank(owner_r), e_col_comm, s));
          whk = dbuf;
        }
        const long lck = k / Qe;
        if (vrows > 0) {
          // column panel: W(i,k) = -A(i,k) L_kk^-1 = -A(i,k) Wh_k^H, into the (zeroed) compact panel on the update engine,
          // then back into the matrix
          GemmArgsT<T> g{};
          g.C = colp + (li_k1 - li_k) * static_cast<long>(nbp);
          g.ldc = mk;
          g.M = static_cast<int>(vrows);
          g.N = nbp;
          g.K = nbp;
          g.alpha = -1.0;
          g.mask = kMaskNone;
          g.nbp = nbp;
          g.P = g.Q = 1;
          DLAF_CUDA_CHECK(cudaMemset2DAsync(g.C, sizeof(T) * mk, 0, sizeof(T) * vrows, nbp, s));
          const Operand<T> opA{slab + static_cast<long>(li_k1) * nbp + lck * nbp * lds, lds, vrows, 0}, opW{whk, nbp, nbp, 0};
          bulk.begin_step();
          launches += bulk.split(false, 0, opA, nbp, s);
          launches += bulk.split_extra(opW, nbp, s);
          launches += bulk.gemm_extra(g, opA, 0, opW, s);
          DLAF_CUDA_CHECK(cudaMemcpy2DAsync(slab + static_cast<long>(li_k1) * nbp + lck * nbp * lds, sizeof(T) * lds, g.C,
                                            sizeof(T) * mk, sizeof(T) * vrows, nbp, cudaMemcpyDeviceToDevice, s));
        }
        if (in_row) {
          // W_kk = Wh_k^H: first tile of the extended panel and the final diagonal tile
          size_t idx = 0;
          while (my_diag[idx] != k)
            ++idx;
          pack(dloc + dsz * idx + tsz + wsz, nbp, colp, mk, 1, 0, 0, true, is_complex, false);
          DLAF_CUDA_CHECK(cudaMemcpy2DAsync(tile(k, k), sizeof(T) * lds, colp, sizeof(T) * mk, sizeof(T) * nbp, nbp,
                                            cudaMemcpyDeviceToDevice, s));
        }
      }
      if (k == 0)
        break;
      // extended column panel along the process rows
      if (Qe > 1 && mk > 0)
        DLAF_NCCL_CHECK(ncclBroadcast(colp, colp, static_cast<size_t>(mk) * nbp * NT::mult, NT::value, row_rank(owner_c), e_row_comm, s));
      // row k: B(j) = -L(k,j)^H packed, row k zeroed, down the process columns
      if (ncols > 0) {
        if (in_row && have) {
          pack(slab + static_cast<long>(k / Pe) * nbp, lds, panB, nbp, ncols, static_cast<long>(nbp) * lds, static_cast<long>(tsz),
               true, is_complex, true);
          DLAF_CUDA_CHECK(cudaMemset2DAsync(slab + static_cast<long>(k / Pe) * nbp, sizeof(T) * lds, 0, sizeof(T) * nbp,
                                            static_cast<size_t>(ncols) * nbp, s));
        }
        if (Pe > 1)
          DLAF_NCCL_CHECK(ncclBroadcast(panB, panB, tsz * ncols * NT::mult, NT::value, col_rank(owner_r), e_col_comm, s));
      }
      // trailing update + row panel: C(i >= k, j < k) -= V B^H
      if (mk > 0 && ncols > 0) {
        GemmArgsT<T> g{};
        g.C = slab + static_cast<long>(li_k) * nbp;
        g.ldc = lds;
        g.M = static_cast<int>(mk);
        g.N = ncols * nbp;
        g.K = nbp;
        g.alpha = -1.0;
        g.mask = kMaskNone;
        g.nbp = nbp;
        g.P = g.Q = 1;
        launches += bulk.run(g, Operand<T>{colp, mk, mk, 0}, Operand<T>{panB, nbp, static_cast<long>(ncols) * nbp, static_cast<long>(tsz)},
                             false, s);
      }
    }
  }

  // =================================================================================================================
  // (2) A^-1 = W^H W (lower triangle)
  if (do_assemble) {
    for (int k = 0; k < nt; ++k) {
      const int owner_r = k % Pe;
      const bool in_row = (erow == owner_r);
      const int ncols = cnt(k + 1, ecol, Qe), nrows = cnt(k + 1, erow, Pe);
      // P(j) = W(k,j)^H for my block columns j <= k, row k zeroed, down the process columns
      if (ncols > 0) {
        if (in_row && have) {
          pack(slab + static_cast<long>(k / Pe) * nbp, lds, panB, nbp, ncols, static_cast<long>(nbp) * lds, static_cast<long>(tsz),
               true, is_complex, false);
          DLAF_CUDA_CHECK(cudaMemset2DAsync(slab + static_cast<long>(k / Pe) * nbp, sizeof(T) * lds, 0, sizeof(T) * nbp,
                                            static_cast<size_t>(ncols) * nbp, s));
        }
        if (Pe > 1)
          DLAF_NCCL_CHECK(ncclBroadcast(panB, panB, tsz * ncols * NT::mult, NT::value, col_rank(owner_r), e_col_comm, s));
      }
      // P(i) for my block rows i <= k: from the rank of my process row that sits in process column i % Qe
      Operand<T> opa{};
      if (Qe > 1) {
        if (nrows > 0) {
          DLAF_NCCL_CHECK(ncclGroupStart());
          for (int li = 0; li < nrows; ++li) {
            const long i = static_cast<long>(li) * Pe + erow;
            const int root_v = static_cast<int>(i % Qe);
            T* recv = panA + tsz * li;
            const T* send = (root_v == ecol) ? panB + tsz * (i / Qe) : recv;
            DLAF_NCCL_CHECK(ncclBroadcast(send, recv, tsz * NT::mult, NT::value, row_rank(root_v), e_row_comm, s));
          }
          DLAF_NCCL_CHECK(ncclGroupEnd());
        }
        opa = Operand<T>{panA, nbp, static_cast<long>(nrows) * nbp, static_cast<long>(tsz)};
      }
      else {
        // one process column: every tile of row k is already here; my block rows are every Pe-th of them
        opa = Operand<T>{panB + tsz * erow, nbp, static_cast<long>(nrows) * nbp, static_cast<long>(tsz) * Pe};
      }
      if (nrows > 0 && ncols > 0) {
        GemmArgsT<T> g{};
        g.C = slab;
        g.ldc = lds;
        g.M = nrows * nbp;
        g.N = ncols * nbp;
        g.K = nbp;
        g.alpha = 1.0;
        g.mask = kMaskLower;
        g.nbp = nbp;
        g.P = Pe;
        g.Q = Qe;
        g.prow = erow;
        g.pcol = ecol;
        launches += bulk.run(g, opa, Operand<T>{panB, nbp, static_cast<long>(ncols) * nbp, static_cast<long>(tsz)}, Pe * Qe == 1, s);
      }
    }
  }

  if (have) {
    dim3 grid(nbp / 32, nbp / 32, ltr * ltc), block(32, 8);
    inv_convert_kernel<T, false><<<grid, block, 0, s>>>(a_user, lda, slab, lds, p.n, p.nb, nbp, Pe, Qe, erow, ecol, ltr,
                                                        transposed, unit, true);
    DLAF_CUDA_CHECK(cudaGetLastError());
    ++launches;
  }
  const int fired = bulk.finish(s);
  if (guard_steps)
    *guard_steps = fired;
  DLAF_CUDA_CHECK(cudaStreamSynchronize(s));
  pool_free(slab);
  pool_free(colp);
  pool_free(panA);
  pool_free(panB);
  pool_free(dbuf);
  pool_free(dloc);
  return launches;
}

#define INST(T) \
  template long inverse_device<T>(const InverseProblem&, int, T*, long, ncclComm_t, ncclComm_t, cudaStream_t, int*);
INST(float)
INST(double)
INST(float2)
INST(double2)

}  // namespace dlaf_b200
