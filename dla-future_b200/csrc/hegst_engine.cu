// See hegst_engine.h.
#include "hegst_engine.h"

#include <vector>

#include "bulk_update.cuh"
#include "comm.h"
#include "common.h"
#include "distribution.h"
#include "pool.h"
#include "tri_kernels.cuh"

namespace dlaf_b200 {

namespace {

using namespace trik;

// dst (ldd) <- full Hermitian tile from the LOWER triangle of src (lds), real diagonal. One CTA per column.
template <class T>
__global__ void hegst_herm_fill_kernel(const T* __restrict__ src, long lds, T* __restrict__ dst, long ldd, int nbp) {
  const int c = blockIdx.x;
  for (int r = threadIdx.x; r < nbp; r += blockDim.x) {
    T v;
    if (r > c)
      v = src[r + static_cast<long>(c) * lds];
    else if (r < c)
      v = conj_val(src[c + static_cast<long>(r) * lds]);
    else
      v = make_real<T>(re_part(src[r + static_cast<long>(c) * lds]));
    dst[r + static_cast<long>(c) * ldd] = v;
  }
}

}  // namespace

template <class T>
long generalized_to_standard_device(const HegstProblem& p, T* a_user, long lda, const T* l_user, long ldl, ncclComm_t row_comm,
                                    ncclComm_t col_comm, cudaStream_t s, int* guard_steps) {
  using NT = NcclType<T>;
  constexpr int G = Gran<T>::value;
  long launches = 0;
  if (guard_steps)
    *guard_steps = 0;
  const bool transposed = (p.uplo == 'U' || p.uplo == 'u');
  DLAF_B200_ASSERT(transposed || p.uplo == 'L' || p.uplo == 'l', "uplo must be L or U");
  if (p.n == 0)
    return 0;
  const int nbp = static_cast<int>(round_up(p.nb, G));
  const int ns = nbp / G;
  const int nt = ceil_div(p.n, p.nb);
  const size_t tsz = static_cast<size_t>(nbp) * nbp, wsz = static_cast<size_t>(ns) * G * G;
  const bool is_complex = sizeof(T) == 2 * sizeof(base_t<T>);
  // ---- engine grid: the lower-triangular problem; uplo == 'U' works on A^H, L = U^H with the grid roles swapped
  const int Pe = transposed ? p.Q : p.P, Qe = transposed ? p.P : p.Q;
  const int erow = transposed ? p.pcol : p.prow, ecol = transposed ? p.prow : p.pcol;
  ncclComm_t e_row_comm = transposed ? col_comm : row_comm;  // ranks of my ENGINE row (size Qe)
  ncclComm_t e_col_comm = transposed ? row_comm : col_comm;  // ranks of my ENGINE column (size Pe)
  const int e_src_in_col = transposed ? p.src_col : p.src_row, e_src_in_row = transposed ? p.src_row : p.src_col;
  auto col_rank = [&](int v_erow) { return (v_erow + e_src_in_col) % Pe; };
  auto row_rank = [&](int v_ecol) { return (v_ecol + e_src_in_row) % Qe; };
  DLAF_B200_ASSERT(Pe == 1 || e_col_comm != nullptr, "communicator required");
  DLAF_B200_ASSERT(Qe == 1 || e_row_comm != nullptr, "communicator required");
  const bool single = (Pe * Qe == 1);

  const int ltr = cnt(nt, erow, Pe), ltc = cnt(nt, ecol, Qe);
  const long lds = static_cast<long>(ltr > 0 ? ltr : 1) * nbp;
  const bool have = ltr > 0 && ltc > 0;
  T *sa = nullptr, *sl = nullptr;  // engine slabs of A and L
  if (have) {
    sa = pool_alloc<T>(lds * ltc * nbp);
    sl = pool_alloc<T>(lds * ltc * nbp);
    DLAF_CUDA_CHECK(cudaMemsetAsync(sa, 0, sizeof(T) * lds * ltc * nbp, s));
    DLAF_CUDA_CHECK(cudaMemsetAsync(sl, 0, sizeof(T) * lds * ltc * nbp, s));
    dim3 grid(nbp / 32, nbp / 32, ltr * ltc), block(32, 8);
    inv_convert_kernel<T, true><<<grid, block, 0, s>>>(a_user, lda, sa, lds, p.n, p.nb, nbp, Pe, Qe, erow, ecol, ltr, transposed,
                                                       false, false);
    inv_convert_kernel<T, true><<<grid, block, 0, s>>>(const_cast<T*>(l_user), ldl, sl, lds, p.n, p.nb, nbp, Pe, Qe, erow, ecol,
                                                       ltr, transposed, false, true);
    DLAF_CUDA_CHECK(cudaGetLastError());
    launches += 2;
  }
  auto pack = [&](const T* src, long ld_src, T* dst, long ld_dst, int ntiles, long src_stride, long dst_stride, bool tr, bool cj) {
    if (ntiles <= 0)
      return;
    dim3 grid(nbp / 32, nbp / 32, ntiles), block(32, 8);
    trsm_pack_tile_kernel<T><<<grid, block, 0, s>>>(src, ld_src, dst, nbp, tr, cj, src_stride, dst_stride, ld_dst, false);
    DLAF_CUDA_CHECK(cudaGetLastError());
    ++launches;
  };
  auto herm_fill = [&](const T* src, long ld_src, T* dst, long ld_dst) {
    hegst_herm_fill_kernel<T><<<nbp, 128, 0, s>>>(src, ld_src, dst, ld_dst, nbp);
    DLAF_CUDA_CHECK(cudaGetLastError());
    ++launches;
  };

  // ---- my diagonal tiles of L: [L_kk packed | inverted G-blocks]
  std::vector<int> my_diag;
  for (int k = 0; k < nt; ++k)
    if (k % Pe == erow && k % Qe == ecol)
      my_diag.push_back(k);
  const size_t dsz = tsz + wsz;
  T* dloc = nullptr;
  if (!my_diag.empty()) {
    dloc = pool_alloc<T>(dsz * my_diag.size());
    for (size_t i = 0; i < my_diag.size(); ++i) {
      const int k = my_diag[i];
      pack(sl + static_cast<long>(k / Pe) * nbp + static_cast<long>(k / Qe) * nbp * lds, lds, dloc + dsz * i, nbp, 1, 0, 0, false, false);
    }
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
    ++launches;
  }
  auto my_diag_slot = [&](int k) {
    size_t idx = 0;
    while (my_diag[idx] != k)
      ++idx;
    return dloc + dsz * idx;
  };

  // ---- workspaces: dbuf = [L_kk | blocks | A_kk (full Hermitian)], two scratch tiles, panel pairs
  T *dbuf = nullptr, *h0 = nullptr, *h1 = nullptr, *pan = nullptr, *panT = nullptr, *rbuf = nullptr;
  dbuf = pool_alloc<T>((2 * tsz + wsz));
  h0 = pool_alloc<T>(tsz);
  h1 = pool_alloc<T>(tsz);
  if (!single) {
    pan = pool_alloc<T>(2 * tsz * (ltr > 0 ? ltr : 1));    // per local row tile: [P(i) | L(i,k)]
    panT = pool_alloc<T>(2 * tsz * (ltc > 0 ? ltc : 1));   // per local column tile: [P(j) | L(j,k)]
  }
  rbuf = pool_alloc<T>(tsz * (ltc > 0 ? ltc : 1));
  BulkUpdate<T> bulk;
  bulk.init(static_cast<long>(ltr) * nbp, static_cast<long>(ltc) * nbp, nbp, 2 * nt, s, 2);
  bulk.init_extra(nbp, nbp);

  // =================================================================================================================
  // phase 1
  for (int k = 0; k < nt; ++k) {
    const int owner_r = k % Pe, owner_c = k % Qe;
    const bool in_col = (ecol == owner_c), in_row = (erow == owner_r);
    const int li1 = cnt(k + 1, erow, Pe), lj1 = cnt(k + 1, ecol, Qe);
    const int nrows = ltr - li1, ncols = ltc - lj1;  // my tiles below / right of k
    const long lck = k / Qe;
    if (in_col) {
      if (in_row) {
        T* akk = sa + static_cast<long>(k / Pe) * nbp + lck * nbp * lds;
        const T* mine = my_diag_slot(k);
        DLAF_CUDA_CHECK(cudaMemcpyAsync(dbuf, mine, sizeof(T) * dsz, cudaMemcpyDeviceToDevice, s));
        herm_fill(akk, lds, h0, nbp);
        launches += solve_rows_against_tile<T>(h0, nbp, nbp, dbuf, nbp, dbuf + tsz, ns, true, s);  // A_kk inv(L_kk)^H
        pack(h0, nbp, h1, nbp, 1, 0, 0, true, is_complex);
        launches += solve_rows_against_tile<T>(h1, nbp, nbp, dbuf, nbp, dbuf + tsz, ns, true, s);  // (inv(L_kk) A_kk inv(L_kk)^H)^H
        herm_fill(h1, nbp, dbuf + dsz, nbp);
        DLAF_CUDA_CHECK(cudaMemcpy2DAsync(akk, sizeof(T) * lds, dbuf + dsz, sizeof(T) * nbp, sizeof(T) * nbp, nbp,
                                          cudaMemcpyDeviceToDevice, s));
      }
      if (k < nt - 1 && Pe > 1)
        DLAF_NCCL_CHECK(ncclBroadcast(dbuf, dbuf, (2 * tsz + wsz) * NT::mult, NT::value, col_rank(owner_r), e_col_comm, s));
    }
    if (k == nt - 1)
      break;
    T* pcol = have ? sa + static_cast<long>(li1) * nbp + lck * nbp * lds : nullptr;        // A(i > k, k), my rows
    const T* lcol = have ? sl + static_cast<long>(li1) * nbp + lck * nbp * lds : nullptr;  // L(i > k, k)
    GemmArgsT<T> hm{};  // P -= 1/2 L(i>k,k) A_kk, on the update engine (the factor 1/2 is an exact rescaling)
    const Operand<T> opLs{lcol, lds, static_cast<long>(nrows) * nbp, 0}, opH{dbuf + dsz, nbp, nbp, 0};
    bulk.begin_step();
    if (in_col && nrows > 0) {
      launches += solve_rows_against_tile<T>(pcol, lds, static_cast<long>(nrows) * nbp, dbuf, nbp, dbuf + tsz, ns, true, s);
      hm.C = pcol;
      hm.ldc = lds;
      hm.M = nrows * nbp;
      hm.N = nbp;
      hm.K = nbp;
      hm.alpha = -0.5;
      hm.mask = kMaskNone;
      hm.nbp = nbp;
      hm.P = hm.Q = 1;
      launches += bulk.split(false, 1, opLs, nbp, s);  // L(i>k,k): also the A-side operand of the trailing update
      launches += bulk.split_extra(opH, nbp, s);
      launches += bulk.gemm_extra(hm, opLs, 1, opH, s);
    }
    // operands of the trailing update
    Operand<T> opP{}, opL{}, opPT{}, opLT{};
    if (single) {
      opP = Operand<T>{pcol, lds, static_cast<long>(nrows) * nbp, 0};
      opL = Operand<T>{lcol, lds, static_cast<long>(nrows) * nbp, 0};
    }
    else {
      if (nrows > 0) {
        if (in_col) {
          pack(pcol, lds, pan, nbp, nrows, nbp, 2 * static_cast<long>(tsz), false, false);
          pack(lcol, lds, pan + tsz, nbp, nrows, nbp, 2 * static_cast<long>(tsz), false, false);
        }
        if (Qe > 1)
          DLAF_NCCL_CHECK(ncclBroadcast(pan, pan, 2 * tsz * nrows * NT::mult, NT::value, row_rank(owner_c), e_row_comm, s));
      }
      if (ncols > 0) {
        // tile pair j from the rank of my process column that sits in process row j % Pe (it holds row j in its panel)
        if (Pe > 1)
          DLAF_NCCL_CHECK(ncclGroupStart());
        for (int lj = lj1; lj < ltc; ++lj) {
          const long j = static_cast<long>(lj) * Qe + ecol;
          const int root_v = static_cast<int>(j % Pe);
          T* recv = panT + 2 * tsz * (lj - lj1);
          const T* mine = (root_v == erow) ? pan + 2 * tsz * (j / Pe - li1) : nullptr;
          if (Pe > 1)
            DLAF_NCCL_CHECK(ncclBroadcast(mine ? mine : recv, recv, 2 * tsz * NT::mult, NT::value, col_rank(root_v), e_col_comm, s));
          else
            DLAF_CUDA_CHECK(cudaMemcpyAsync(recv, mine, sizeof(T) * 2 * tsz, cudaMemcpyDeviceToDevice, s));
        }
        if (Pe > 1)
          DLAF_NCCL_CHECK(ncclGroupEnd());
      }
      opP = Operand<T>{pan, nbp, static_cast<long>(nrows) * nbp, 2 * static_cast<long>(tsz)};
      opL = Operand<T>{pan + tsz, nbp, static_cast<long>(nrows) * nbp, 2 * static_cast<long>(tsz)};
      opPT = Operand<T>{panT, nbp, static_cast<long>(ncols) * nbp, 2 * static_cast<long>(tsz)};
      opLT = Operand<T>{panT + tsz, nbp, static_cast<long>(ncols) * nbp, 2 * static_cast<long>(tsz)};
    }
    // A(i,j) -= P(i) L(j,k)^H + L(i,k) P(j)^H, i >= j > k
    if (nrows > 0 && ncols > 0) {
      GemmArgsT<T> g{};
      g.C = sa + static_cast<long>(li1) * nbp + static_cast<long>(lj1) * nbp * lds;
      g.ldc = lds;
      g.M = nrows * nbp;
      g.N = ncols * nbp;
      g.K = nbp;
      g.alpha = -1.0;
      g.mask = kMaskLower;
      g.nbp = nbp;
      g.P = Pe;
      g.Q = Qe;
      g.prow = erow;
      g.pcol = ecol;
      g.ti0 = li1;
      g.tj0 = lj1;
      launches += bulk.split(false, 0, opP, nbp, s);
      if (!in_col)
        launches += bulk.split(false, 1, opL, nbp, s);  // (ranks of the panel's column: done before the first hemm)
      if (!single) {
        launches += bulk.split(true, 0, opPT, nbp, s);
        launches += bulk.split(true, 1, opLT, nbp, s);
        launches += bulk.gemm(g, opP, 0, opLT, 1, false, s);
        launches += bulk.gemm(g, opL, 1, opPT, 0, false, s);
      }
      else {
        launches += bulk.gemm(g, opP, 0, opL, 1, true, s);
        launches += bulk.gemm(g, opL, 1, opP, 0, true, s);
      }
    }
    // second half of the hemm
    if (in_col && nrows > 0)
      launches += bulk.gemm_extra(hm, opLs, 1, opH, s);
  }

  // =================================================================================================================
  // phase 2: X(j, :j) = inv(L_jj) C(j, :j);  C(t > j, :j) -= L(t,j) X(j, :j)
  for (int j = 1; j < nt; ++j) {
    const int owner_r = j % Pe, owner_c = j % Qe;
    const bool in_col = (ecol == owner_c), in_row = (erow == owner_r);
    const int ncols = cnt(j, ecol, Qe);        // my block columns left of j
    const int li1 = cnt(j + 1, erow, Pe);
    const int nrows = ltr - li1;               // my block rows below j
    if (in_row) {
      if (in_col)
        DLAF_CUDA_CHECK(cudaMemcpyAsync(dbuf, my_diag_slot(j), sizeof(T) * dsz, cudaMemcpyDeviceToDevice, s));
      if (Qe > 1)
        DLAF_NCCL_CHECK(ncclBroadcast(dbuf, dbuf, dsz * NT::mult, NT::value, row_rank(owner_c), e_row_comm, s));
      if (ncols > 0) {
        const long ldr = static_cast<long>(ncols) * nbp;
        T* row = sa + static_cast<long>(j / Pe) * nbp;
        pack(row, lds, rbuf, ldr, ncols, static_cast<long>(nbp) * lds, nbp, true, is_complex);    // R = C(j, :j)^H, plain panel
        launches += solve_rows_against_tile<T>(rbuf, ldr, ldr, dbuf, nbp, dbuf + tsz, ns, true, s);  // R inv(L_jj)^H = X(j, :j)^H
        pack(rbuf, ldr, row, lds, ncols, nbp, static_cast<long>(nbp) * lds, true, is_complex);    // X(j, :j) back into A
      }
    }
    if (j == nt - 1)
      break;
    if (ncols > 0 && Pe > 1)
      DLAF_NCCL_CHECK(ncclBroadcast(rbuf, rbuf, tsz * ncols * NT::mult, NT::value, col_rank(owner_r), e_col_comm, s));
    Operand<T> opL{};
    if (Qe > 1) {
      if (nrows > 0) {
        if (in_col)
          pack(sl + static_cast<long>(li1) * nbp + static_cast<long>(j / Qe) * nbp * lds, lds, pan, nbp, nrows, nbp,
               static_cast<long>(tsz), false, false);
        DLAF_NCCL_CHECK(ncclBroadcast(pan, pan, tsz * nrows * NT::mult, NT::value, row_rank(owner_c), e_row_comm, s));
      }
      opL = Operand<T>{pan, nbp, static_cast<long>(nrows) * nbp, static_cast<long>(tsz)};
    }
    else {
      opL = Operand<T>{have ? sl + static_cast<long>(li1) * nbp + static_cast<long>(j / Qe) * nbp * lds : nullptr, lds,
                       static_cast<long>(nrows) * nbp, 0};
    }
    if (nrows > 0 && ncols > 0) {
      GemmArgsT<T> g{};
      g.C = sa + static_cast<long>(li1) * nbp;
      g.ldc = lds;
      g.M = nrows * nbp;
      g.N = ncols * nbp;
      g.K = nbp;
      g.alpha = -1.0;
      g.mask = kMaskNone;
      g.nbp = nbp;
      g.P = g.Q = 1;
      launches += bulk.run(g, opL, Operand<T>{rbuf, static_cast<long>(ncols) * nbp, static_cast<long>(ncols) * nbp, 0}, false, s);
    }
  }

  if (have) {
    dim3 grid(nbp / 32, nbp / 32, ltr * ltc), block(32, 8);
    inv_convert_kernel<T, false><<<grid, block, 0, s>>>(a_user, lda, sa, lds, p.n, p.nb, nbp, Pe, Qe, erow, ecol, ltr, transposed,
                                                        false, false);
    DLAF_CUDA_CHECK(cudaGetLastError());
    ++launches;
  }
  const int fired = bulk.finish(s);
  if (guard_steps)
    *guard_steps = fired;
  DLAF_CUDA_CHECK(cudaStreamSynchronize(s));
  pool_free(sa);
  pool_free(sl);
  pool_free(dloc);
  pool_free(dbuf);
  pool_free(h0);
  pool_free(h1);
  pool_free(pan);
  pool_free(panT);
  pool_free(rbuf);
  return launches;
}

#define INST(T)                                                                                                            \
  template long generalized_to_standard_device<T>(const HegstProblem&, T*, long, const T*, long, ncclComm_t, ncclComm_t, \
                                                  cudaStream_t, int*);
INST(float)
INST(double)
INST(float2)
INST(double2)

}  // namespace dlaf_b200
