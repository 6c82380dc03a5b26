// See engine.h.
#include "engine.h"

#include <cstdint>
#include <cstdlib>
#include <string>
#include <type_traits>

#include "comm.h"
#include "common.h"

namespace dlaf_b200 {

// sm_partition.cu
cudaStream_t create_bulk_stream_with_reserved_sms(int reserve_sms, int priority);

namespace {
// DLAF_B200_RESERVE_SMS: SMs kept free of bulk-update CTAs for the panel chain (0 = off)
// Default: 8 on grids of 4 or more GPUs, where the run is bound by the panel chain almost from the first step
// (measured N = 32768: 2x2 grid +2.6 %, single GPU -2 %: profiles/r01_sm_reservation.log).
int reserved_sms(int ranks) {
  static const int v = [] {
    const char* e = std::getenv("DLAF_B200_RESERVE_SMS");
    return e ? std::atoi(e) : -1;
  }();
  return v >= 0 ? v : (ranks >= 4 ? 8 : 0);
}
cudaStream_t make_bulk_stream(int priority, int ranks) {
  cudaStream_t st = create_bulk_stream_with_reserved_sms(reserved_sms(ranks), priority);
  if (st == nullptr)
    DLAF_CUDA_CHECK(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, priority));
  return st;
}

// default fp64 bulk-update engine when DLAF_B200_D_BULK is not set
constexpr bool kOzakiDefault = true;

inline int cnt_tiles(long g_end, int r, int grid) {
  // number of global tile indices g in [0, g_end) with g % grid == r
  return g_end > r ? static_cast<int>((g_end - r + grid - 1) / grid) : 0;
}
}  // namespace

template <class T>
PotrfEngine<T>::PotrfEngine(const EngineGeometry& g, ncclComm_t row_comm, ncclComm_t col_comm, ncclComm_t row_comm_h,
                            ncclComm_t col_comm_h)
    : geo_(g), row_comm_(row_comm), col_comm_(col_comm), row_comm_h_(row_comm_h), col_comm_h_(col_comm_h) {
  DLAF_B200_ASSERT(g.n >= 0 && g.nb >= 1, "matrix / block size");
  DLAF_B200_ASSERT(g.P >= 1 && g.Q >= 1 && g.prow >= 0 && g.prow < g.P && g.pcol >= 0 && g.pcol < g.Q,
                   "grid coordinates");
  DLAF_B200_ASSERT(g.P == 1 || col_comm != nullptr, "column communicator required when P > 1");
  DLAF_B200_ASSERT(g.Q == 1 || row_comm != nullptr, "row communicator required when Q > 1");
  nbp_ = static_cast<int>(round_up(g.nb, G));
  nt_ = ceil_div(g.n, g.nb);
  ns_ = nbp_ / G;
  ltr_ = cnt_tiles(nt_, g.prow, g.P);
  ltc_ = cnt_tiles(nt_, g.pcol, g.Q);
  own_ld_ = static_cast<long>(ltr_) * nbp_;
  ld_ = own_ld_;

  int least, greatest;
  DLAF_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  DLAF_CUDA_CHECK(cudaStreamCreateWithPriority(&sH_, cudaStreamNonBlocking, greatest));
  DLAF_CUDA_CHECK(cudaStreamCreateWithPriority(&sM_, cudaStreamNonBlocking, greatest));
  sL_ = make_bulk_stream(least, g.P * g.Q);
  DLAF_CUDA_CHECK(cudaStreamCreateWithPriority(&sR_, cudaStreamNonBlocking, greatest));
  DLAF_CUDA_CHECK(cudaEventCreateWithFlags(&ev_start_, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) {
    DLAF_CUDA_CHECK(cudaEventCreateWithFlags(&evP_[i], cudaEventDisableTiming));
    DLAF_CUDA_CHECK(cudaEventCreateWithFlags(&evB_[i], cudaEventDisableTiming));
    DLAF_CUDA_CHECK(cudaEventCreateWithFlags(&evC_[i], cudaEventDisableTiming));
    DLAF_CUDA_CHECK(cudaEventCreateWithFlags(&evD_[i], cudaEventDisableTiming));
    DLAF_CUDA_CHECK(cudaEventCreateWithFlags(&evF_[i], cudaEventDisableTiming));
    DLAF_CUDA_CHECK(cudaEventCreateWithFlags(&evT1_[i], cudaEventDisableTiming));
    DLAF_CUDA_CHECK(cudaEventCreateWithFlags(&evC1_[i], cudaEventDisableTiming));
    DLAF_CUDA_CHECK(cudaEventCreateWithFlags(&evPc_[i], cudaEventDisableTiming));
  }
  {
    // Two chains on P x Q grids: needs the independent communicator pair for the critical-path stream.
    const char* sc = std::getenv("DLAF_B200_SPLIT_CHAIN");
    dist_split_ = g.P * g.Q > 1 && (g.P == 1 || col_comm_h != nullptr) && (g.Q == 1 || row_comm_h != nullptr) &&
                  (sc == nullptr || std::atoi(sc) != 0);
  }
  const size_t wsz = static_cast<size_t>(ns_) * G * G;
  const size_t tsz = static_cast<size_t>(nbp_) * nbp_;
  for (int i = 0; i < 2; ++i) {
    if (geo_.P > 1) {
      DLAF_CUDA_CHECK(cudaMalloc(&diagbuf_[i], sizeof(T) * (tsz + wsz)));
      DLAF_CUDA_CHECK(cudaMalloc(&panelT_[i], sizeof(T) * tsz * (ltc_ > 0 ? ltc_ : 1)));
    }
    else {
      DLAF_CUDA_CHECK(cudaMalloc(&wbuf_[i], sizeof(T) * wsz));
    }
    if (geo_.P * geo_.Q > 1)
      DLAF_CUDA_CHECK(cudaMalloc(&panel_[i], sizeof(T) * tsz * (ltr_ > 0 ? ltr_ : 1)));
    if (dist_split_ && geo_.Q > 1)
      DLAF_CUDA_CHECK(cudaMalloc(&crit_[i], sizeof(T) * tsz));
  }
  DLAF_CUDA_CHECK(cudaMalloc(&d_info_, sizeof(int)));
  DLAF_CUDA_CHECK(cudaMallocHost(&h_info_, sizeof(int)));
  *h_info_ = 0;
  if constexpr (std::is_same_v<T, double>) {
    // DLAF_B200_D_BULK = ozaki | dmma : fp64 trailing update on tcgen05 (exact int8 slice products) or on DMMA
    const char* e = std::getenv("DLAF_B200_D_BULK");
    const bool want = (e != nullptr) ? (std::string(e) == "ozaki") : kOzakiDefault;
    use_ozaki_ = want && nt_ > 1 && ltr_ > 0 && ltc_ > 0 && nbp_ <= 512;
    const char* sc = std::getenv("DLAF_B200_SPLIT_CHAIN");
    split_chain_ = use_ozaki_ && geo_.P * geo_.Q == 1 && (sc == nullptr || std::atoi(sc) != 0);
    if (use_ozaki_) {
      DLAF_CUDA_CHECK(cudaMalloc(&oz_flag_, sizeof(int) * nt_));
      DLAF_CUDA_CHECK(cudaMallocHost(&h_oz_flags_, sizeof(int) * nt_));
      for (int k = 0; k < nt_; ++k)
        h_oz_flags_[k] = 0;
    }
    if (use_ozaki_) {
      oring_ = (geo_.P * geo_.Q == 1) ? kOzRing : 2;
      if (const char* r = std::getenv("DLAF_B200_OZAKI_RING"))
        oring_ = (geo_.P * geo_.Q == 1) ? std::min(kOzRing, std::max(2, std::atoi(r))) : 2;
      for (int i = 0; i < oring_; ++i)
        osplit_[i].allocate(static_cast<long>(ltr_) * nbp_, nbp_);
      if (geo_.P > 1)
        for (int i = 0; i < 2; ++i)
          osplitT_[i].allocate(static_cast<long>(ltc_) * nbp_, nbp_);
    }
  }
  if constexpr (std::is_same_v<T, float>) {
    // DLAF_B200_S_SIMT=1 keeps the SIMT fp32 kernel everywhere (A/B measurements)
    use_tf32_ = nt_ > 1 && ltr_ > 0 && ltc_ > 0 && std::getenv("DLAF_B200_S_SIMT") == nullptr;
    if (use_tf32_)
      for (int i = 0; i < 2; ++i) {
        split_[i].allocate(static_cast<long>(ltr_) * nbp_, nbp_);
        if (geo_.P > 1)
          splitT_[i].allocate(static_cast<long>(ltc_) * nbp_, nbp_);
      }
  }
}

template <class T>
PotrfEngine<T>::~PotrfEngine() {
  cudaStreamSynchronize(sH_);
  cudaStreamSynchronize(sM_);
  cudaStreamSynchronize(sL_);
  cudaStreamSynchronize(sR_);
  cudaStreamDestroy(sR_);
  for (int i = 0; i < 2; ++i) {
    cudaEventDestroy(evF_[i]);
    cudaEventDestroy(evT1_[i]);
    cudaEventDestroy(evC1_[i]);
    cudaEventDestroy(evPc_[i]);
    cudaFree(crit_[i]);
    cudaEventDestroy(evC_[i]);
    cudaEventDestroy(evD_[i]);
    cudaFree(wbuf_[i]);
    cudaFree(diagbuf_[i]);
    cudaFree(panel_[i]);
    cudaFree(panelT_[i]);
    cudaEventDestroy(evP_[i]);
    cudaEventDestroy(evB_[i]);
  }
  cudaEventDestroy(ev_start_);
  for (auto e : evIn_)
    cudaEventDestroy(e);
  for (auto e : evBc_)
    cudaEventDestroy(e);
  for (size_t c = 1; c < sLc_.size(); ++c)
    cudaStreamDestroy(sLc_[c]);
  if (evOut_)
    cudaEventDestroy(evOut_);
  if (sIn_)
    cudaStreamDestroy(sIn_);
  if (sOut_)
    cudaStreamDestroy(sOut_);
  for (auto e : prof_ev_)
    cudaEventDestroy(e);
  for (auto e : chain_ev_)
    cudaEventDestroy(e);
  for (auto e : diag_ev_)
    cudaEventDestroy(e);
  if constexpr (std::is_same_v<T, float>) {
    for (int i = 0; i < 2; ++i) {
      split_[i].release();
      splitT_[i].release();
    }
  }
  if constexpr (std::is_same_v<T, double>) {
    for (int i = 0; i < kOzRing; ++i)
      osplit_[i].release();
    for (int i = 0; i < 2; ++i)
      osplitT_[i].release();
  }
  cudaFree(oz_flag_);
  cudaFreeHost(h_oz_flags_);
  cudaFree(own_slab_);
  cudaFree(d_info_);
  cudaFreeHost(h_info_);
  cudaStreamDestroy(sH_);
  cudaStreamDestroy(sM_);
  cudaStreamDestroy(sL_);
}

template <class T>
int PotrfEngine<T>::cnt_rows(long g_end) const {
  return cnt_tiles(g_end, geo_.prow, geo_.P);
}
template <class T>
int PotrfEngine<T>::cnt_cols(long g_end) const {
  return cnt_tiles(g_end, geo_.pcol, geo_.Q);
}

template <class T>
LayoutParams PotrfEngine<T>::layout(long ldu, bool transposed) const {
  LayoutParams p;
  p.n = geo_.n;
  p.nb = geo_.nb;
  p.nbp = nbp_;
  p.nt = nt_;
  p.P = geo_.P;
  p.Q = geo_.Q;
  p.prow = geo_.prow;
  p.pcol = geo_.pcol;
  p.ltr = ltr_;
  p.ltc = ltc_;
  p.ld = ld_;
  p.ldu = ldu;
  p.transposed = transposed ? 1 : 0;
  return p;
}

template <class T>
T* PotrfEngine<T>::slab() {
  if (own_slab_ == nullptr && ltr_ > 0 && ltc_ > 0) {
    const size_t bytes = sizeof(T) * static_cast<size_t>(own_ld_) * ltc_ * nbp_;
    DLAF_CUDA_CHECK(cudaMalloc(&own_slab_, bytes));
  }
  if (!external_) {
    data_ = own_slab_;
    ld_ = own_ld_;
  }
  return own_slab_;
}

template <class T>
bool PotrfEngine<T>::can_run_in_place(const EngineGeometry& g, const void* dev, long ld) {
  return g.nb % G == 0 && g.n % g.nb == 0 && ld % 2 == 0 && (reinterpret_cast<uintptr_t>(dev) & 15) == 0;
}

template <class T>
void PotrfEngine<T>::bind_external(T* dev, long ld) {
  DLAF_B200_ASSERT(can_run_in_place(geo_, dev, ld), "in-place operation needs unpadded, aligned tiles");
  external_ = true;
  data_ = dev;
  ld_ = ld;
}

template <class T>
void PotrfEngine<T>::unbind_external() {
  external_ = false;
  data_ = own_slab_;
  ld_ = own_ld_;
}

template <class T>
void PotrfEngine<T>::load(const T* user, long ldu, bool transposed, cudaStream_t s) {
  DLAF_B200_ASSERT(!external_, "load() targets the engine's own slab");
  if (ltr_ == 0 || ltc_ == 0)
    return;
  slab();
  const LayoutParams p = layout(ldu, transposed);
  if (padded())
    launch_pad_identity<T>(data_, p, s);
  launch_to_slab<T>(data_, user, p, s);
}

template <class T>
void PotrfEngine<T>::store(T* user, long ldu, bool transposed, cudaStream_t s) {
  DLAF_B200_ASSERT(!external_, "store() reads the engine's own slab");
  if (ltr_ == 0 || ltc_ == 0)
    return;
  launch_from_slab<T>(data_, user, layout(ldu, transposed), s);
}

template <class T>
void PotrfEngine<T>::gemm(const GemmArgsT<T>& a, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0)
    return;
  launch_gemm_nt<T>(a, st);
  ++launches_;
}

// Diagonal tile (nbp x nbp) = ns diagonal blocks of G: block Cholesky + inverse (potrf_tile.cu), the
// blocks below via GEMM with the inverse, the rest of the tile via a masked SYRK-shaped GEMM.
// (Measured and dropped: a one-block look-ahead inside the tile on a second stream — the critical stream keeps
// two dependent launches per block either way, 349 vs 342 us per tile in isolation, profiles/r01_chain_*.)
template <class T>
void PotrfEngine<T>::factor_diag_tile(T* tile, long ld, T* w, int k, cudaStream_t st) {
  if constexpr (std::is_same_v<T, double>) {
    // fp64, tiles up to 512: the whole tile (factor + inverted 128-blocks) in ONE cluster launch (potrf_tile_cluster.cu)
    if (potrf_tile_cluster_supported(nbp_)) {
      if (profiling_) {
        while (diag_ev_.size() < diag_used_ + 2) {
          cudaEvent_t e;
          DLAF_CUDA_CHECK(cudaEventCreate(&e));
          diag_ev_.push_back(e);
        }
        DLAF_CUDA_CHECK(cudaEventRecord(diag_ev_[diag_used_], st));
      }
      launch_potrf_tile_cluster_f64(tile, ld, w, nbp_, d_info_, k * geo_.nb, st);
      ++launches_;
      if (profiling_) {
        DLAF_CUDA_CHECK(cudaEventRecord(diag_ev_[diag_used_ + 1], st));
        diag_used_ += 2;
      }
      return;
    }
  }
  for (int j = 0; j < ns_; ++j) {
    T* tjj = tile + static_cast<long>(j) * G * (1 + ld);
    T* wj = w + static_cast<long>(j) * G * G;
    if (profiling_) {  // per-launch in-situ time of the diagonal-block kernel (DLAF_B200_CHAIN_DEBUG prints the sum)
      while (diag_ev_.size() < diag_used_ + 2) {
        cudaEvent_t e;
        DLAF_CUDA_CHECK(cudaEventCreate(&e));
        diag_ev_.push_back(e);
      }
      DLAF_CUDA_CHECK(cudaEventRecord(diag_ev_[diag_used_], st));
    }
    launch_potrf_inv<T>(tjj, ld, wj, G, d_info_, k * geo_.nb + j * G, st);
    ++launches_;
    if (profiling_) {
      DLAF_CUDA_CHECK(cudaEventRecord(diag_ev_[diag_used_ + 1], st));
      diag_used_ += 2;
    }
    const int m = (ns_ - 1 - j) * G;
    if (m == 0)
      break;
    T* below = tjj + G;
    GemmArgsT<T> a{};
    a.A = below;
    a.lda = ld;
    a.B = wj;
    a.ldb = G;
    a.C = below;  // in place: below <- below * inv(L_jj)^H
    a.ldc = ld;
    a.M = m;
    a.N = G;
    a.K = G;
    a.alpha = 1.0;
    a.beta = 0.0;
    a.mask = kMaskNone;
    a.nbp = nbp_;
    a.P = a.Q = 1;
    gemm(a, st);
    GemmArgsT<T> u{};
    u.A = below;
    u.lda = ld;
    u.B = below;
    u.ldb = ld;
    u.C = tjj + static_cast<long>(G) * (1 + ld);
    u.ldc = ld;
    u.M = m;
    u.N = m;
    u.K = G;
    u.alpha = -1.0;
    u.beta = 1.0;
    u.mask = kMaskLower;  // relative to the diagonal of this sub-block
    u.nbp = 1 << 30;
    u.P = u.Q = 1;
    gemm(u, st);
  }
}

// Panel TRSM  B <- B * L_kk^-H  (reference: trsm(Right, Lower, ConjTrans, NonUnit), impl.h:62-66) as
// block substitution over the ns diagonal blocks: every step is a tensor-core GEMM.
template <class T>
void PotrfEngine<T>::trsm_panel(T* b, long ldb, int m, const T* tkk, long ldt, const T* w, cudaStream_t st) {
  if constexpr (std::is_same_v<T, double> || std::is_same_v<T, double2>) {
    // one launch: every CTA runs the whole block substitution for its own rows (gemm_dmma.cuh, gemm_zdmma.cu);
    // DLAF_B200_TRSM=steps keeps the 2*ns-1 separate GEMM launches (A/B measurements)
    static const bool fused = [] {
      const char* e = std::getenv("DLAF_B200_TRSM");
      return e == nullptr || std::string(e) != "steps";
    }();
    if (fused && m > 0) {
      if constexpr (std::is_same_v<T, double>) {
        TrsmFusedArgs a{};
        a.B = b;
        a.ldb = ldb;
        a.T = tkk;
        a.ldt = ldt;
        a.W = w;
        a.ns = ns_;
        launch_trsm_fused_f64(a, m, st);
      }
      else {
        launch_trsm_fused_z(b, ldb, m, tkk, ldt, w, ns_, st);
      }
      ++launches_;
      return;
    }
  }
  for (int j = 0; j < ns_; ++j) {
    T* bj = b + static_cast<long>(j) * G * ldb;
    if (j > 0) {
      GemmArgsT<T> a{};
      a.A = b;
      a.lda = ldb;
      a.B = tkk + static_cast<long>(j) * G;  // row block j of L_kk, columns [0, j*G)
      a.ldb = ldt;
      a.C = bj;
      a.ldc = ldb;
      a.M = m;
      a.N = G;
      a.K = j * G;
      a.alpha = -1.0;
      a.beta = 1.0;
      a.mask = kMaskNone;
      a.nbp = nbp_;
      a.P = a.Q = 1;
      gemm(a, st);
    }
    GemmArgsT<T> s{};
    s.A = bj;
    s.lda = ldb;
    s.B = w + static_cast<long>(j) * G * G;
    s.ldb = G;
    s.C = bj;  // in place
    s.ldc = ldb;
    s.M = m;
    s.N = G;
    s.K = G;
    s.alpha = 1.0;
    s.beta = 0.0;
    s.mask = kMaskNone;
    s.nbp = nbp_;
    s.P = s.Q = 1;
    gemm(s, st);
  }
}

// P_k on stream H: diagonal tile, its broadcast down the owning process column, the panel TRSM, and
// the two panel broadcasts (row-wise, then "transposed" column-wise: broadcast_panel.h:107-188).
template <class T>
void PotrfEngine<T>::panel_step(int k, bool wait_column) {
  using NT = NcclType<T>;
  const int P = geo_.P, Q = geo_.Q;
  const int owner_r = k % P, owner_c = k % Q;
  const bool in_row = (geo_.prow == owner_r), in_col = (geo_.pcol == owner_c);
  const int lkr = k / P, lkc = k / Q;
  const int li1 = cnt_rows(k + 1), lj1 = cnt_cols(k + 1);
  const int mt = ltr_ - li1;
  const size_t tsz = static_cast<size_t>(nbp_) * nbp_;
  const size_t wsz = static_cast<size_t>(ns_) * G * G;
  const int slot = k % 2;

  if (k == 0)
    chain_stamp(0, 0);
  chain_stamp(k, 1);
  if (in_col) {
    wait_columns(lkc + 1, sH_);
    const T* tkk = nullptr;
    long ldt = 0;
    const T* w = nullptr;
    if (P > 1) {
      T* dbuf = diagbuf_[slot];
      if (in_row) {
        T* tile = tile_ptr(lkr, lkc);
        factor_diag_tile(tile, ld_, dbuf + tsz, k, sH_);
        DLAF_CUDA_CHECK(cudaMemcpy2DAsync(dbuf, sizeof(T) * nbp_, tile, sizeof(T) * ld_, sizeof(T) * nbp_,
                                          nbp_, cudaMemcpyDeviceToDevice, sH_));
      }
      chain_stamp(k, 2);
      DLAF_NCCL_CHECK(ncclBroadcast(dbuf, dbuf, (tsz + wsz) * NT::mult, NT::value, col_comm_rank(owner_r),
                                    col_comm_, sH_));
      chain_stamp(k, 3);
      tkk = dbuf;
      ldt = nbp_;
      w = dbuf + tsz;
    }
    else {
      T* tile = tile_ptr(lkr, lkc);
      factor_diag_tile(tile, ld_, wbuf_[slot], k, sH_);
      tkk = tile;
      ldt = ld_;
      w = wbuf_[slot];
      chain_stamp(k, 2);
      chain_stamp(k, 3);
    }
    if (mt > 0 && split_chain()) {
      // Two chains instead of one (1 x 1 grid, int8 engine): the critical stream only solves the FIRST tile row of
      // the panel — all the next diagonal tile needs (its update then runs on native fp64, straight from that
      // tile) — while stream R solves the rest, cuts the digit planes and publishes the panel (evP) to the
      // column / bulk updates:      H: potrf(k) | trsm(k+1,k) | update(k+1,k+1) | potrf(k+1) ...
      //                             R:          | trsm(rows k+2..) | split | evP(k)
      DLAF_CUDA_CHECK(cudaEventRecord(evF_[slot], sH_));
      if (wait_column)  // rows of block column k below the diagonal tile: updated on stream M
        DLAF_CUDA_CHECK(cudaStreamWaitEvent(sH_, evC_[(k - 1) % 2], 0));
      trsm_panel(tile_ptr(li1, lkc), ld_, nbp_, tkk, ldt, w, sH_);
      DLAF_CUDA_CHECK(cudaEventRecord(evT1_[slot], sH_));
      wait_columns(lkc + 1, sR_);
      DLAF_CUDA_CHECK(cudaStreamWaitEvent(sR_, evF_[slot], 0));
      if (wait_column)
        DLAF_CUDA_CHECK(cudaStreamWaitEvent(sR_, evC_[(k - 1) % 2], 0));
      if (mt > 1)
        trsm_panel(tile_ptr(li1 + 1, lkc), ld_, (mt - 1) * nbp_, tkk, ldt, w, sR_);
      if (k < nt_ - 1) {
        if (k >= oring_)  // the digit planes of this ring slot were last read by the updates of step k - oring_
          wait_bulk(k - oring_, -1, sR_);
        DLAF_CUDA_CHECK(cudaStreamWaitEvent(sR_, evT1_[slot], 0));
        if constexpr (std::is_same_v<T, double>) {
          osplit_[k % oring_].split(tile_ptr(li1, lkc), ld_, static_cast<long>(mt) * nbp_, sR_, 0, 0, oz_flag_ + k);
          ++launches_;
        }
      }
      chain_stamp(k, 4);
      chain_stamp(k, 5);
      DLAF_CUDA_CHECK(cudaEventRecord(evP_[slot], sR_));
      download_column(lkc, evP_[slot]);
      return;
    }
    if (mt > 0) {
      if (wait_column)  // rows of block column k below the diagonal tile: updated on stream M
        DLAF_CUDA_CHECK(cudaStreamWaitEvent(sH_, evC_[(k - 1) % 2], 0));
      trsm_panel(tile_ptr(li1, lkc), ld_, mt * nbp_, tkk, ldt, w, sH_);
      if constexpr (std::is_same_v<T, float>) {
        if (use_tf32_ && k < nt_ - 1 && P * Q == 1) {
          split_[slot].split(tile_ptr(li1, lkc), ld_, static_cast<long>(mt) * nbp_, sH_);
          ++launches_;
        }
      }
      if constexpr (std::is_same_v<T, double>) {
        if (use_ozaki_ && k < nt_ - 1 && P * Q == 1) {
          if (k >= oring_)  // ring slot last read by the updates of step k - oring_ (all column chunks)
            wait_bulk(k - oring_, -1, sH_);
          osplit_[k % oring_].split(tile_ptr(li1, lkc), ld_, static_cast<long>(mt) * nbp_, sH_, 0, 0, oz_flag_ + k);
          ++launches_;
        }
      }
    }
  }

  if (!in_col) {
    chain_stamp(k, 2);
    chain_stamp(k, 3);
  }
  chain_stamp(k, 4);
  if (k < nt_ - 1 && P * Q > 1) {
    if (mt > 0) {
      if (in_col) {
        launch_pack_panel<T>(tile_ptr(li1, lkc), ld_, panel_[slot], nbp_, mt, sH_);
        ++launches_;
      }
      if (Q > 1)
        DLAF_NCCL_CHECK(ncclBroadcast(panel_[slot], panel_[slot], tsz * mt * NT::mult, NT::value,
                                      row_comm_rank(owner_c), row_comm_, sH_));
    }
    if (P > 1 && ltc_ - lj1 > 0) {
      DLAF_NCCL_CHECK(ncclGroupStart());
      for (int lj = lj1; lj < ltc_; ++lj) {
        const long gj = static_cast<long>(lj) * Q + geo_.pcol;
        const int root_v = static_cast<int>(gj % P);
        T* recv = panelT_[slot] + tsz * (lj - lj1);
        const T* send = recv;
        if (root_v == geo_.prow)
          send = panel_[slot] + tsz * (gj / P - li1);
        DLAF_NCCL_CHECK(ncclBroadcast(send, recv, tsz * NT::mult, NT::value, col_comm_rank(root_v),
                                      col_comm_, sH_));
      }
      DLAF_NCCL_CHECK(ncclGroupEnd());
    }
  }
  if constexpr (std::is_same_v<T, float>) {
    if (use_tf32_ && k < nt_ - 1 && P * Q > 1) {
      // split the (tile-contiguous) panel workspaces every rank now holds
      if (mt > 0) {
        split_[slot].split(panel_[slot], nbp_, static_cast<long>(mt) * nbp_, sH_, nbp_, static_cast<long>(tsz));
        ++launches_;
      }
      if (P > 1 && ltc_ - lj1 > 0) {
        splitT_[slot].split(panelT_[slot], nbp_, static_cast<long>(ltc_ - lj1) * nbp_, sH_, nbp_, static_cast<long>(tsz));
        ++launches_;
      }
    }
  }
  if constexpr (std::is_same_v<T, double>) {
    if (use_ozaki_ && k < nt_ - 1 && P * Q > 1) {
      if (mt > 0) {
        osplit_[slot].split(panel_[slot], nbp_, static_cast<long>(mt) * nbp_, sH_, nbp_, static_cast<long>(tsz), oz_flag_ + k);
        ++launches_;
      }
      if (P > 1 && ltc_ - lj1 > 0) {
        osplitT_[slot].split(panelT_[slot], nbp_, static_cast<long>(ltc_ - lj1) * nbp_, sH_, nbp_, static_cast<long>(tsz),
                             oz_flag_ + k);
        ++launches_;
      }
    }
  }
  chain_stamp(k, 5);
  DLAF_CUDA_CHECK(cudaEventRecord(evP_[slot], sH_));
  if (in_col)
    download_column(lkc, evP_[slot]);  // block column k is final on this rank
}

// P_k on a P x Q grid as TWO chains (round 2; round 1 ran everything below on stream H in program order, so that every
// rank sat inside ncclBroadcast waiting for the owners' potrf + whole-panel TRSM):
//
//   stream H (critical):  potrf(k) on its owner | L_kk + inverted blocks down the owner COLUMN (col_comm_h) |
//                         TRSM of tile (k+1,k) only, on its owner X | that tile along process ROW (k+1) % P (row_comm_h) |
//                         [factorize(): native update of tile (k+1,k+1) from it on its owner N, potrf(k+1) ...]
//   stream R (panel):     rest of the panel TRSM | pack | whole local panel along the process rows (row_comm) |
//                         transposed panel down the process columns (col_comm, grouped) | digit / hi-lo splits | evP(k)
//
// The two streams use DIFFERENT communicators (NCCL orders operations per communicator), the reference's answer to the
// same problem being 3 round-robin communicator clones (src/communication/communicator_grid.cpp:64-75) and per-tile
// MPI_Ibcast pipelines (communication/broadcast_panel.h:156-187). Collective order per communicator is identical on all
// of its ranks: one diagonal broadcast and one critical-tile broadcast per step on the H pair, one panel broadcast and one
// grouped transposed broadcast per step on the R pair.
template <class T>
void PotrfEngine<T>::panel_step_dist(int k, bool wait_column) {
  using NT = NcclType<T>;
  const int P = geo_.P, Q = geo_.Q;
  const int owner_r = k % P, owner_c = k % Q;
  const bool in_row = (geo_.prow == owner_r), in_col = (geo_.pcol == owner_c);
  const int lkr = k / P, lkc = k / Q;
  const int li1 = cnt_rows(k + 1), lj1 = cnt_cols(k + 1);
  const int mt = ltr_ - li1;
  const size_t tsz = static_cast<size_t>(nbp_) * nbp_;
  const size_t wsz = static_cast<size_t>(ns_) * G * G;
  const int slot = k % 2;
  const bool has_next = k < nt_ - 1;
  const bool in_next_row = has_next && geo_.prow == (k + 1) % P;  // takes part in the critical-tile broadcast
  const bool is_x = in_next_row && in_col;                         // owns tile (k+1, k)
  const bool is_n = in_next_row && geo_.pcol == (k + 1) % Q;       // owns tile (k+1, k+1)

  if (k == 0)
    chain_stamp(0, 0);
  chain_stamp(k, 1);
  // ---------------------------------------------------------------- stream H
  const T* tkk = nullptr;
  long ldt = 0;
  const T* w = nullptr;
  if (in_col) {
    wait_columns(lkc + 1, sH_);
    if (k >= 2)  // diagbuf_[slot] was last read by the panel TRSM of step k-2 (stream R, finished before its evP)
      DLAF_CUDA_CHECK(cudaStreamWaitEvent(sH_, evP_[slot], 0));
    if (P > 1) {
      T* dbuf = diagbuf_[slot];
      if (in_row) {
        T* tile = tile_ptr(lkr, lkc);
        factor_diag_tile(tile, ld_, dbuf + tsz, k, sH_);
        DLAF_CUDA_CHECK(cudaMemcpy2DAsync(dbuf, sizeof(T) * nbp_, tile, sizeof(T) * ld_, sizeof(T) * nbp_, nbp_,
                                          cudaMemcpyDeviceToDevice, sH_));
      }
      chain_stamp(k, 2);
      DLAF_NCCL_CHECK(ncclBroadcast(dbuf, dbuf, (tsz + wsz) * NT::mult, NT::value, col_comm_rank(owner_r), col_comm_h_,
                                    sH_));
      chain_stamp(k, 3);
      tkk = dbuf;
      ldt = nbp_;
      w = dbuf + tsz;
    }
    else {
      T* tile = tile_ptr(lkr, lkc);
      factor_diag_tile(tile, ld_, wbuf_[slot], k, sH_);
      tkk = tile;
      ldt = ld_;
      w = wbuf_[slot];
      chain_stamp(k, 2);
      chain_stamp(k, 3);
    }
    DLAF_CUDA_CHECK(cudaEventRecord(evF_[slot], sH_));
  }
  else {
    chain_stamp(k, 2);
    chain_stamp(k, 3);
  }
  crit_next_ = nullptr;
  if (is_x) {
    if (wait_column)  // tile (k+1, k) was updated with panel k-1 on stream M
      DLAF_CUDA_CHECK(cudaStreamWaitEvent(sH_, evC1_[(k - 1) % 2], 0));
    trsm_panel(tile_ptr(li1, lkc), ld_, nbp_, tkk, ldt, w, sH_);
    DLAF_CUDA_CHECK(cudaEventRecord(evT1_[slot], sH_));
  }
  if (in_next_row) {
    if (Q > 1) {
      if (is_x)
        DLAF_CUDA_CHECK(cudaMemcpy2DAsync(crit_[slot], sizeof(T) * nbp_, tile_ptr(li1, lkc), sizeof(T) * ld_,
                                          sizeof(T) * nbp_, nbp_, cudaMemcpyDeviceToDevice, sH_));
      DLAF_NCCL_CHECK(ncclBroadcast(crit_[slot], crit_[slot], tsz * NT::mult, NT::value, row_comm_rank(owner_c),
                                    row_comm_h_, sH_));
      if (is_n) {
        crit_next_ = crit_[slot];
        crit_next_ld_ = nbp_;
      }
    }
    else if (is_n) {  // Q == 1: X and N are the same rank, the tile is used where it lies
      crit_next_ = tile_ptr(li1, lkc);
      crit_next_ld_ = ld_;
    }
  }
  chain_stamp(k, 4);
  // ---------------------------------------------------------------- stream R
  if (in_col) {
    wait_columns(lkc + 1, sR_);
    DLAF_CUDA_CHECK(cudaStreamWaitEvent(sR_, evF_[slot], 0));
    if (wait_column)  // rows of block column k below the diagonal tile: updated on stream M
      DLAF_CUDA_CHECK(cudaStreamWaitEvent(sR_, evC_[(k - 1) % 2], 0));
    const int first = li1 + (is_x ? 1 : 0);
    if (ltr_ - first > 0)
      trsm_panel(tile_ptr(first, lkc), ld_, (ltr_ - first) * nbp_, tkk, ldt, w, sR_);
    if (is_x)
      DLAF_CUDA_CHECK(cudaStreamWaitEvent(sR_, evT1_[slot], 0));
  }
  if (has_next) {
    if (k >= 2) {  // the panel workspaces / splits of this slot were last read by the updates of step k-2
      wait_bulk(k - 2, -1, sR_);
      DLAF_CUDA_CHECK(cudaStreamWaitEvent(sR_, evC_[slot], 0));
    }
    if (mt > 0) {
      if (in_col) {
        launch_pack_panel<T>(tile_ptr(li1, lkc), ld_, panel_[slot], nbp_, mt, sR_);
        ++launches_;
      }
      if (Q > 1)
        DLAF_NCCL_CHECK(ncclBroadcast(panel_[slot], panel_[slot], tsz * mt * NT::mult, NT::value,
                                      row_comm_rank(owner_c), row_comm_, sR_));
    }
    // Transposed panel: the tile of block column k+1 goes FIRST on the ranks that own that column (it is all the update
    // of column k+1 on stream M needs besides the row panel), the rest follows in one group.
    const bool own_next_col = (geo_.pcol == (k + 1) % Q);
    const int nT = ltc_ - lj1;  // transposed tiles I hold (local columns >= k+1)
    auto bcast_T = [&](int lj_begin, int lj_end) {
      if (lj_end - lj_begin > 1)
        DLAF_NCCL_CHECK(ncclGroupStart());
      for (int lj = lj_begin; lj < lj_end; ++lj) {
        const long gj = static_cast<long>(lj) * Q + geo_.pcol;
        const int root_v = static_cast<int>(gj % P);
        T* recv = panelT_[slot] + tsz * (lj - lj1);
        const T* send = recv;
        if (root_v == geo_.prow)
          send = panel_[slot] + tsz * (gj / P - li1);
        DLAF_NCCL_CHECK(ncclBroadcast(send, recv, tsz * NT::mult, NT::value, col_comm_rank(root_v), col_comm_, sR_));
      }
      if (lj_end - lj_begin > 1)
        DLAF_NCCL_CHECK(ncclGroupEnd());
    };
    const bool early = (P > 1 && own_next_col && nT > 0);
    if (early)
      bcast_T(lj1, lj1 + 1);
    if constexpr (std::is_same_v<T, double>) {
      if (use_ozaki_) {
        if (mt > 0) {
          osplit_[slot].split(panel_[slot], nbp_, static_cast<long>(mt) * nbp_, sR_, nbp_, static_cast<long>(tsz),
                              oz_flag_ + k);
          ++launches_;
        }
        if (early) {
          osplitT_[slot].split(panelT_[slot], nbp_, nbp_, sR_, nbp_, static_cast<long>(tsz), oz_flag_ + k);
          ++launches_;
        }
      }
    }
    DLAF_CUDA_CHECK(cudaEventRecord(evPc_[slot], sR_));  // (fp32 / complex: the early tile is there, splits follow below)
    if (P > 1 && nT - (early ? 1 : 0) > 0)
      bcast_T(lj1 + (early ? 1 : 0), ltc_);
    if constexpr (std::is_same_v<T, float>) {
      if (use_tf32_) {
        if (mt > 0) {
          split_[slot].split(panel_[slot], nbp_, static_cast<long>(mt) * nbp_, sR_, nbp_, static_cast<long>(tsz));
          ++launches_;
        }
        if (P > 1 && nT > 0) {
          splitT_[slot].split(panelT_[slot], nbp_, static_cast<long>(nT) * nbp_, sR_, nbp_, static_cast<long>(tsz));
          ++launches_;
        }
      }
    }
    if constexpr (std::is_same_v<T, double>) {
      if (use_ozaki_ && P > 1 && nT - (early ? 1 : 0) > 0) {
        const int t0 = early ? 1 : 0;
        osplitT_[slot].split(panelT_[slot] + tsz * t0, nbp_, static_cast<long>(nT - t0) * nbp_, sR_, nbp_,
                             static_cast<long>(tsz), oz_flag_ + k, static_cast<long>(t0) * nbp_);
        ++launches_;
      }
    }
  }
  else {
    DLAF_CUDA_CHECK(cudaEventRecord(evPc_[slot], sR_));
  }
  chain_stamp(k, 5);
  DLAF_CUDA_CHECK(cudaEventRecord(evP_[slot], sR_));
  if (in_col)
    download_column(lkc, evP_[slot]);  // block column k is final on this rank
}

// Tile (k+1, k+1) -= L(k+1,k) L(k+1,k)^H on its owner, in native arithmetic straight from the critical tile (two-chain
// schedule on grids): all the next potrf waits for.
template <class T>
void PotrfEngine<T>::update_next_diag_from_crit(int k, cudaStream_t st) {
  if (crit_next_ == nullptr)
    return;
  const int li = (k + 1) / geo_.P, lj = (k + 1) / geo_.Q;
  wait_columns(lj + 1, st);
  GemmArgsT<T> a{};
  a.A = crit_next_;
  a.lda = crit_next_ld_;
  a.B = crit_next_;
  a.ldb = crit_next_ld_;
  a.C = tile_ptr(li, lj);
  a.ldc = ld_;
  a.M = a.N = a.K = nbp_;
  a.alpha = -1.0;
  a.beta = 1.0;
  a.mask = kMaskLower;  // relative to the diagonal of this tile
  a.nbp = 1 << 30;
  a.P = a.Q = 1;
  gemm(a, st);
}

// U_k on my local block columns [cj0, cj0 + ncols), rows from local tile row ri0 on (mrows elements):
// ONE masked GEMM launch.
template <class T>
void PotrfEngine<T>::launch_update(int k, int cj0, int ncols, int ri0, int mrows, bool count_flops, cudaStream_t st,
                                   bool native) {
  const int P = geo_.P, Q = geo_.Q;
  const int li1 = cnt_rows(k + 1), lj1 = cnt_cols(k + 1);
  const long gj0 = static_cast<long>(cj0) * Q + geo_.pcol;
  const int slot = k % 2;
  const size_t tsz = static_cast<size_t>(nbp_) * nbp_;
  const int lkc = k / Q;

  GemmArgsT<T> a{};
  a.C = tile_ptr(ri0, cj0);
  a.ldc = ld_;
  a.M = mrows;
  a.N = ncols * nbp_;
  a.K = nbp_;
  a.alpha = -1.0;
  a.beta = 1.0;
  a.mask = kMaskLower;
  a.nbp = nbp_;
  a.P = P;
  a.Q = Q;
  a.prow = geo_.prow;
  a.pcol = geo_.pcol;
  a.ti0 = ri0;
  a.tj0 = cj0;
  if (P * Q == 1) {
    // panel and transposed panel are the block column k of the matrix itself (no copy, like the
    // reference's setTile aliasing, impl.h:261)
    a.A = tile_ptr(ri0, lkc);
    a.lda = ld_;
    a.B = tile_ptr(cj0, lkc);
    a.ldb = ld_;
  }
  else {
    a.A = panel_[slot] + tsz * (ri0 - li1);
    a.lda = nbp_;
    a.a_ts = static_cast<long>(tsz);
    if (P == 1) {
      // every row is local: tile (gj, k) sits in my panel at index gj - (k+1)
      a.B = panel_[slot] + tsz * (gj0 - (k + 1));
      a.b_ts = static_cast<long>(tsz) * Q;
    }
    else {
      a.B = panelT_[slot] + tsz * (cj0 - lj1);
      a.b_ts = static_cast<long>(tsz);
    }
    a.ldb = nbp_;
  }
  if (profiling_ && count_flops) {
    // algorithmic flops of this launch: 2 nbp^3 per off-diagonal tile, nbp^3 per diagonal tile
    // (herk), counted on global tile indices; complex: x4 (6 mul + 2 add per complex mac = 8 flop)
    double tiles = 0;
    for (int lj = cj0; lj < cj0 + ncols; ++lj) {
      const long gj = static_cast<long>(lj) * Q + geo_.pcol;
      for (int li = ri0; li < ltr_; ++li) {
        const long gi = static_cast<long>(li) * P + geo_.prow;
        if (gi > gj)
          tiles += 2.0;
        else if (gi == gj)
          tiles += 1.0;
      }
    }
    const double cplx = (sizeof(T) == 2 * sizeof(base_t<T>)) ? 4.0 : 1.0;
    last_update_flops_ += tiles * cplx * static_cast<double>(nbp_) * nbp_ * nbp_;
  }
  if constexpr (std::is_same_v<T, float>) {
    if (use_tf32_) {
      // A rows: panel row index = (local tile row - li1) * nbp. B rows: 1 x 1 grid and P == 1 alias the column
      // panel (tile gj sits at index gj - (k+1), consecutive local columns Q tiles apart); P > 1 uses the
      // transposed-panel split (tile index = local column - lj1).
      const long a_row = static_cast<long>(ri0 - li1) * nbp_;
      if (P > 1)
        launch_gemm_tf32x3(a, split_[slot], a_row, splitT_[slot], static_cast<long>(cj0 - lj1) * nbp_, st);
      else
        launch_gemm_tf32x3(a, split_[slot], a_row, split_[slot], static_cast<long>(gj0 - (k + 1)) * nbp_, st,
                           static_cast<long>(Q) * nbp_);
      ++launches_;
      return;
    }
  }
  if constexpr (std::is_same_v<T, double>) {
    if (use_ozaki_ && !native) {
      const long a_row = static_cast<long>(ri0 - li1) * nbp_;
      const int* guard = oz_flag_ + k;
      if (P > 1)
        launch_gemm_ozaki_i8(a, osplit_[slot], a_row, osplitT_[slot], static_cast<long>(cj0 - lj1) * nbp_, st, 0, guard);
      else
        launch_gemm_ozaki_i8(a, osplit_[k % oring_], a_row, osplit_[k % oring_], static_cast<long>(gj0 - (k + 1)) * nbp_, st,
                             static_cast<long>(Q) * nbp_, guard);
      // guard raised by the digit split of this step's panel: the same update on the native fp64 kernel (otherwise
      // a handful of CTAs that read the flag and leave)
      launch_gemm_nt_f64_if(a, guard, st);
      launches_ += 2;
      return;
    }
  }
  gemm(a, st);
}

// U_k: trailing update of my local tiles with panel k.
//   kBulk            everything to the right of block column k+1
//   kNextDiag        only the diagonal tile (k+1, k+1) (if mine) — all the next potrf waits for
//   kNextColumnRest  block column k+1 below its diagonal tile (if mine) — what the next TRSM waits for
template <class T>
void PotrfEngine<T>::update(int k, UpdatePart part, cudaStream_t st) {
  const int P = geo_.P, Q = geo_.Q;
  const int li1 = cnt_rows(k + 1), lj1 = cnt_cols(k + 1);
  last_update_flops_ = 0.0;
  if (ltr_ - li1 <= 0 || ltc_ - lj1 <= 0)
    return;
  const bool own_next = ((k + 1) % Q == geo_.pcol);
  int cj0, ncols;
  if (part != kBulk) {
    if (!own_next)
      return;
    cj0 = lj1;
    ncols = 1;
  }
  else {
    cj0 = own_next ? lj1 + 1 : lj1;
    ncols = ltc_ - cj0;
  }
  if (ncols <= 0)
    return;
  auto first_row = [&](int lj) { return cnt_rows(static_cast<long>(lj) * Q + geo_.pcol); };
  int ri0 = first_row(cj0);  // first local row tile on or below the diagonal of column cj0
  int mrows = (ltr_ - ri0) * nbp_;
  if (part != kBulk) {
    const bool own_diag = ((k + 1) % P == geo_.prow);  // then local row ri0 is global row k+1
    if (part == kNextDiag) {
      if (!own_diag)
        return;
      mrows = nbp_;
    }
    else {
      if (own_diag) {
        ri0 += 1;
        mrows -= nbp_;
      }
      // ri0 is now the first local row with global index >= k+2; it IS row k+2 (the tile the next critical TRSM
      // needs) iff this rank's process row owns it
      const bool own_first = (k + 2 < nt_) && ((k + 2) % P == geo_.prow);
      if (part == kNextColumnFirst) {
        if (!own_first)
          return;
        mrows = nbp_;
      }
      else if (part == kNextColumnTail && own_first) {
        ri0 += 1;
        mrows -= nbp_;
      }
    }
  }
  if (mrows <= 0)
    return;
  const int cend = cj0 + ncols;
  wait_columns(cend, st);
  launch_update(k, cj0, ncols, ri0, mrows, part == kBulk, st, part == kNextDiag && split_chain());
}

// ---- host pipelining (factorize_host) -------------------------------------------------------------
template <class T>
int PotrfEngine<T>::chunk_of(int lj) const {
  if (in_end_.empty())
    return 0;
  size_t c = 0;
  while (c + 1 < in_end_.size() && in_end_[c] <= lj)
    ++c;
  return static_cast<int>(c);
}

template <class T>
void PotrfEngine<T>::wait_bulk(int k, int lj, cudaStream_t st) {
  const int nc = nchunks();
  // lj < 0: a workspace of step k is about to be reused -> every chunk must be done with it. lj >= 0: data dependency on
  // one block column -> its chunk is enough, unless the panel workspaces of step k + 2 are about to overwrite what the
  // other chunks still read (grids, fp32 splits: two slots; the int8 digit planes on a 1 x 1 grid live in a ring and
  // are protected separately, see panel_step).
  const bool ring = use_ozaki_ && geo_.P * geo_.Q == 1 && oring_ > 2;
  if (geo_.P * geo_.Q > 1 || lj < 0 || (split_panels() && !ring)) {
    for (int c = 0; c < nc; ++c)
      DLAF_CUDA_CHECK(cudaStreamWaitEvent(st, evBc_[oring_ * c + k % oring_], 0));
  }
  else {
    DLAF_CUDA_CHECK(cudaStreamWaitEvent(st, evBc_[oring_ * chunk_of(lj) + k % oring_], 0));
  }
}

template <class T>
void PotrfEngine<T>::issue_uploads() {
  // one 2D copy per local block column (rows from its diagonal tile down), grouped into chunks of
  // ~1 GB (DLAF_B200_UPLOAD_CHUNK_MB; measured at N = 32768: 128 / 256 / 512 MB chunks 228 / 220 / 228 ms end to end,
  // 1024 MB 197 ms — every chunk costs one bulk launch per step, and narrow launches re-read the panel); an event per chunk. The copy stream is in order, so chunk c resident => chunks < c resident.
  static const size_t chunk_mb = [] {
    const char* e = std::getenv("DLAF_B200_UPLOAD_CHUNK_MB");
    const long v = e ? std::atol(e) : 1024;
    return static_cast<size_t>(v < 16 ? 16 : v);
  }();
  const size_t chunk_bytes = chunk_mb << 20;
  const int nb = geo_.nb;
  in_end_.clear();
  size_t acc = 0, used = 0;
  for (int lj = 0; lj < ltc_; ++lj) {
    const long gj = static_cast<long>(lj) * geo_.Q + geo_.pcol;
    const int ri = cnt_rows(gj);
    const long rows = static_cast<long>(ltr_ - ri) * nb;
    if (rows > 0) {
      const T* h = host_ + static_cast<long>(ri) * nb + static_cast<long>(lj) * nb * ldh_;
      DLAF_CUDA_CHECK(cudaMemcpy2DAsync(tile_ptr(ri, lj), sizeof(T) * ld_, h, sizeof(T) * ldh_, sizeof(T) * rows, nb,
                                        cudaMemcpyHostToDevice, sIn_));
      acc += sizeof(T) * rows * nb;
    }
    if (acc >= chunk_bytes || lj == ltc_ - 1) {
      if (evIn_.size() <= used) {
        cudaEvent_t e;
        DLAF_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        evIn_.push_back(e);
      }
      DLAF_CUDA_CHECK(cudaEventRecord(evIn_[used], sIn_));
      in_end_.push_back(lj + 1);
      ++used;
      acc = 0;
    }
  }
}

template <class T>
void PotrfEngine<T>::wait_columns(int lj_end, cudaStream_t st) {
  if (host_ == nullptr || lj_end <= 0 || in_end_.empty())
    return;
  size_t c = 0;
  while (c + 1 < in_end_.size() && in_end_[c] < lj_end)
    ++c;
  DLAF_CUDA_CHECK(cudaStreamWaitEvent(st, evIn_[c], 0));
}

template <class T>
void PotrfEngine<T>::download_column(int lj, cudaEvent_t final_ev) {
  if (host_ == nullptr)
    return;
  const int nb = geo_.nb;
  const long gj = static_cast<long>(lj) * geo_.Q + geo_.pcol;
  const int ri = cnt_rows(gj);
  const long rows = static_cast<long>(ltr_ - ri) * nb;
  if (rows <= 0)
    return;
  DLAF_CUDA_CHECK(cudaStreamWaitEvent(sOut_, final_ev, 0));
  T* h = host_ + static_cast<long>(ri) * nb + static_cast<long>(lj) * nb * ldh_;
  DLAF_CUDA_CHECK(cudaMemcpy2DAsync(h, sizeof(T) * ldh_, tile_ptr(ri, lj), sizeof(T) * ld_, sizeof(T) * rows, nb,
                                    cudaMemcpyDeviceToHost, sOut_));
}

template <class T>
void PotrfEngine<T>::factorize_host(T* host, long ldh, cudaStream_t s) {
  DLAF_B200_ASSERT(!external_ && !padded(), "pipelined host path needs unpadded tiles in the engine's slab");
  if (sIn_ == nullptr) {
    DLAF_CUDA_CHECK(cudaStreamCreateWithFlags(&sIn_, cudaStreamNonBlocking));
    DLAF_CUDA_CHECK(cudaStreamCreateWithFlags(&sOut_, cudaStreamNonBlocking));
    DLAF_CUDA_CHECK(cudaEventCreateWithFlags(&evOut_, cudaEventDisableTiming));
  }
  host_ = host;
  ldh_ = ldh;
  factorize(s);
  host_ = nullptr;
}

template <class T>
void PotrfEngine<T>::chain_stamp(int k, int which) {
  if (!profiling_)
    return;
  const size_t idx = static_cast<size_t>(k) * 6 + which;
  while (chain_ev_.size() <= idx) {
    cudaEvent_t e;
    DLAF_CUDA_CHECK(cudaEventCreate(&e));
    chain_ev_.push_back(e);
  }
  DLAF_CUDA_CHECK(cudaEventRecord(chain_ev_[idx], sH_));
  if (chain_used_ < idx + 1)
    chain_used_ = idx + 1;
}

template <class T>
void PotrfEngine<T>::read_chain_profile(double out[6]) {
  for (int i = 0; i < 6; ++i)
    out[i] = 0.0;
  const size_t steps = chain_used_ / 6;
  for (size_t k = 0; k < steps; ++k) {
    for (int i = 0; i < 5; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, chain_ev_[k * 6 + i], chain_ev_[k * 6 + i + 1]) == cudaSuccess)
        out[i] += ms;
    }
    out[5] += 1.0;
  }
  if (std::getenv("DLAF_B200_CHAIN_DEBUG")) {
    double pot = 0;
    for (size_t i = 0; i + 1 < diag_used_; i += 2) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, diag_ev_[i], diag_ev_[i + 1]) == cudaSuccess)
        pot += ms;
    }
    std::fprintf(stderr, "[dlaf_b200] chain debug: %zu diagonal-block kernels, in-situ sum %.3f ms (avg %.1f us); diag tiles %.3f ms\n",
                 diag_used_ / 2, pot, diag_used_ ? pot * 2000.0 / diag_used_ : 0.0, out[1]);
  }
}

template <class T>
void PotrfEngine<T>::read_profile(double out[3]) {
  out[0] = out[1] = out[2] = 0.0;
  for (size_t i = 0; i < prof_used_; ++i) {
    if (prof_flops_[i] < 0)
      continue;  // nothing was launched for this step
    float ms = 0.f;
    DLAF_CUDA_CHECK(cudaEventElapsedTime(&ms, prof_ev_[2 * i], prof_ev_[2 * i + 1]));
    out[0] += ms;
    out[1] += prof_flops_[i];
    out[2] += 1.0;
  }
}

template <class T>
void PotrfEngine<T>::factorize(cudaStream_t s) {
  launches_ = 0;
  prof_used_ = 0;
  chain_used_ = 0;
  diag_used_ = 0;
  if (nt_ == 0)
    return;
  if (!external_)
    slab();
  DLAF_CUDA_CHECK(cudaEventRecord(ev_start_, s));
  DLAF_CUDA_CHECK(cudaStreamWaitEvent(sH_, ev_start_, 0));
  DLAF_CUDA_CHECK(cudaStreamWaitEvent(sL_, ev_start_, 0));
  DLAF_CUDA_CHECK(cudaMemsetAsync(d_info_, 0, sizeof(int), sH_));
  if (oz_flag_ != nullptr) {
    DLAF_CUDA_CHECK(cudaMemsetAsync(oz_flag_, 0, sizeof(int) * nt_, sH_));
    DLAF_CUDA_CHECK(cudaEventRecord(evF_[0], sH_));  // the split of step 0 may run on stream R
    DLAF_CUDA_CHECK(cudaStreamWaitEvent(sR_, evF_[0], 0));
  }

  DLAF_CUDA_CHECK(cudaStreamWaitEvent(sM_, ev_start_, 0));
  DLAF_CUDA_CHECK(cudaStreamWaitEvent(sR_, ev_start_, 0));
  if (host_) {
    DLAF_CUDA_CHECK(cudaStreamWaitEvent(sIn_, ev_start_, 0));
    DLAF_CUDA_CHECK(cudaStreamWaitEvent(sOut_, ev_start_, 0));
    issue_uploads();
    // Downloads have slack (a block column is final long before the end), uploads do not (every step needs the whole
    // trailing matrix): keep the link for the upload first. DLAF_B200_HOST_DEFER_D2H=0 lets both directions compete.
    static const bool defer = [] {
      const char* e = std::getenv("DLAF_B200_HOST_DEFER_D2H");
      return e == nullptr || std::atoi(e) != 0;
    }();
    if (defer && !evIn_.empty() && !in_end_.empty())
      DLAF_CUDA_CHECK(cudaStreamWaitEvent(sOut_, evIn_[in_end_.size() - 1], 0));
  }
  else {
    // device-resident: column chunks of equal width (DLAF_B200_BULK_CHUNKS, default 1 = one bulk launch per step)
    static const int want = [] {
      const char* e = std::getenv("DLAF_B200_BULK_CHUNKS");
      return e ? std::atoi(e) : 1;
    }();
    in_end_.clear();
    const int nc_res = want < 1 ? 1 : (want > ltc_ ? (ltc_ > 0 ? ltc_ : 1) : want);
    for (int c = 1; c <= nc_res; ++c)
      in_end_.push_back(static_cast<int>(static_cast<long>(ltc_) * c / nc_res));
  }
  // bulk streams / events, one per column chunk
  const int nc = nchunks();
  if (sLc_.empty())
    sLc_.push_back(sL_);
  while (static_cast<int>(sLc_.size()) < nc) {
    int least, greatest;
    DLAF_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    sLc_.push_back(make_bulk_stream(least, geo_.P * geo_.Q));
  }
  while (static_cast<int>(evBc_.size()) < oring_ * nc) {
    cudaEvent_t e;
    DLAF_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    evBc_.push_back(e);
  }
  for (int c = 1; c < nc; ++c)
    DLAF_CUDA_CHECK(cudaStreamWaitEvent(sLc_[c], ev_start_, 0));

  if (dist_split_)
    panel_step_dist(0, false);
  else
    panel_step(0, false);
  for (int k = 0; k < nt_ - 1; ++k) {
    // bulk of U_k on the low-priority stream(s), one launch per column chunk
    const int lj1 = cnt_cols(k + 1);
    const bool own_next = ((k + 1) % geo_.Q == geo_.pcol);
    const int bulk_c0 = own_next ? lj1 + 1 : lj1;
    for (int c = 0; c < nc; ++c) {
      cudaStream_t st = sLc_[c];
      DLAF_CUDA_CHECK(cudaStreamWaitEvent(st, evP_[k % 2], 0));
      const int cbeg = (nc == 1) ? 0 : (c == 0 ? 0 : in_end_[c - 1]);
      const int cend = (nc == 1) ? ltc_ : in_end_[c];
      const int c0 = cbeg > bulk_c0 ? cbeg : bulk_c0;
      const bool prof = profiling_ && nc == 1;
      if (prof) {
        if (prof_ev_.size() < 2 * (prof_used_ + 1)) {
          cudaEvent_t a, b;
          DLAF_CUDA_CHECK(cudaEventCreate(&a));
          DLAF_CUDA_CHECK(cudaEventCreate(&b));
          prof_ev_.push_back(a);
          prof_ev_.push_back(b);
          prof_flops_.push_back(0.0);
        }
        DLAF_CUDA_CHECK(cudaEventRecord(prof_ev_[2 * prof_used_], st));
      }
      const long before = launches_;
      last_update_flops_ = 0.0;
      if (c0 < cend && ltr_ - cnt_rows(k + 1) > 0) {
        if (host_)
          DLAF_CUDA_CHECK(cudaStreamWaitEvent(st, evIn_[c], 0));
        const int r0 = cnt_rows(static_cast<long>(c0) * geo_.Q + geo_.pcol);
        if (ltr_ - r0 > 0)
          launch_update(k, c0, cend - c0, r0, (ltr_ - r0) * nbp_, true, st);
      }
      if (prof) {
        DLAF_CUDA_CHECK(cudaEventRecord(prof_ev_[2 * prof_used_ + 1], st));
        prof_flops_[prof_used_] = (launches_ > before) ? last_update_flops_ : -1.0;
        ++prof_used_;
      }
      DLAF_CUDA_CHECK(cudaEventRecord(evBc_[oring_ * c + k % oring_], st));
    }
    // stream M: block column k+1 below its diagonal tile (needs panel k and the bulk of step k-1 there)
    DLAF_CUDA_CHECK(cudaStreamWaitEvent(sM_, (dist_split_ && use_ozaki_) ? evPc_[k % 2] : evP_[k % 2], 0));
    if (k >= 1)
      wait_bulk(k - 1, own_next ? lj1 : -1, sM_);
    if (dist_split_) {
      update(k, kNextColumnFirst, sM_);  // tile (k+2, k+1): all the next critical-tile TRSM waits for
      DLAF_CUDA_CHECK(cudaEventRecord(evC1_[k % 2], sM_));
      update(k, kNextColumnTail, sM_);
    }
    else {
      update(k, kNextColumnRest, sM_);
    }
    DLAF_CUDA_CHECK(cudaEventRecord(evC_[k % 2], sM_));
    // critical path on stream H: the diagonal tile (k+1,k+1), then P_{k+1} (its TRSM waits for stream M)
    chain_stamp(k + 1, 0);
    if (k >= 1)
      wait_bulk(k - 1, own_next ? lj1 : -1, sH_);
    if (dist_split_) {
      update_next_diag_from_crit(k, sH_);
      panel_step_dist(k + 1, true);
    }
    else {
      update(k, kNextDiag, sH_);
      panel_step(k + 1, true);
    }
  }
  if (host_) {
    DLAF_CUDA_CHECK(cudaEventRecord(evOut_, sOut_));
    DLAF_CUDA_CHECK(cudaStreamWaitEvent(s, evOut_, 0));
  }
  DLAF_CUDA_CHECK(cudaStreamWaitEvent(s, evP_[(nt_ - 1) % 2], 0));
  if (nt_ >= 2) {
    for (int c = 0; c < nc; ++c)
      DLAF_CUDA_CHECK(cudaStreamWaitEvent(s, evBc_[oring_ * c + (nt_ - 2) % oring_], 0));
    DLAF_CUDA_CHECK(cudaStreamWaitEvent(s, evC_[(nt_ - 2) % 2], 0));
  }
  DLAF_CUDA_CHECK(cudaMemcpyAsync(h_info_, d_info_, sizeof(int), cudaMemcpyDeviceToHost, s));
  if (oz_flag_ != nullptr)
    DLAF_CUDA_CHECK(cudaMemcpyAsync(h_oz_flags_, oz_flag_, sizeof(int) * nt_, cudaMemcpyDeviceToHost, s));
}

template <class T>
int PotrfEngine<T>::guard_fallback_steps() const {
  if (h_oz_flags_ == nullptr)
    return -1;
  int c = 0;
  for (int k = 0; k < nt_; ++k)
    c += h_oz_flags_[k] != 0;
  return c;
}

template <class T>
int PotrfEngine<T>::info(cudaStream_t s) {
  DLAF_CUDA_CHECK(cudaStreamSynchronize(s));
  int v = *h_info_;
  if (v > geo_.n)
    v = 0;  // cannot happen (identity padding never fails); keep the LAPACK range
  return v;
}

template class PotrfEngine<float>;
template class PotrfEngine<double>;
template class PotrfEngine<float2>;
template class PotrfEngine<double2>;

}  // namespace dlaf_b200
