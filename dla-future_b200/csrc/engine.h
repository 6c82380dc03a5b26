// Host scheduler of the B200 POTRF path: issues the hand-written kernels of one rank (= one GPU) in
// look-ahead order over two CUDA streams + events and, on a P x Q grid, the NCCL panel broadcasts.
//
// It replaces Cholesky<Backend::GPU, Device::GPU, T>::call_L (local and distributed,
// include/dlaf/factorization/cholesky/impl.h:150-313) together with the machinery underneath it:
// the pika sender DAG (sender/transform.h:52-110), the per-tile async RW mutexes
// (matrix/internal/tile_pipeline.h:19-81), the round-robin panel workspaces (impl.h:218-221) and the
// per-tile MPI_Ibcast pipeline (communication/broadcast_panel.h:107-188). call_U is served by the same
// code on the conjugate-transposed problem (see layout.cuh).
//
// Dependencies per step k (P_k = diagonal tile + panel of column k, U_k(j) = update of block column j
// with panel k):   P_k <- U_{k-1}(k);   U_k(j) <- P_k, U_{k-1}(j).
// Stream H (highest priority): U_{k-1}(k,k) (diagonal tile only), potrf of tile k, its broadcast, the
//                              panel TRSM and the panel broadcasts     — the critical path
// Stream M (highest priority): U_{k-1}(k) below the diagonal tile      — overlaps the potrf of tile k
// Stream L (lowest priority):  U_k(k+2 ..)                             — the bulk, ~95 % of the flops
// which is the reference's 1-column look-ahead priority rule (impl.h:171-173, :280-281) with the
// diagonal tile split off so that it never waits for the rest of its column.
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>

#include <vector>

#include "gemm_args.h"
#include "gemm_ozaki.h"
#include "gemm_tf32.h"
#include "layout.cuh"
#include "potrf_tile.cuh"
#include "types.h"

namespace dlaf_b200 {

struct EngineGeometry {
  long n = 0;  // global matrix size
  int nb = 1;  // user tile (= distribution block) size
  // ENGINE process grid and my source-rank-adjusted ("virtual") coordinates in it. For uplo == 'U'
  // the caller passes the transposed grid.
  int P = 1, Q = 1, prow = 0, pcol = 0;
  // NCCL rank (inside row_comm / col_comm) of virtual coordinate 0, i.e. the source rank offsets.
  int src_in_col_comm = 0;  // rank in col_comm (size P) of virtual row 0
  int src_in_row_comm = 0;  // rank in row_comm (size Q) of virtual column 0
};

template <class T>
class PotrfEngine {
public:
  static constexpr int G = Gran<T>::value;

  // row_comm_h / col_comm_h: independent clones for the collectives of the critical-path stream (may be null: the
  // distributed two-chain schedule is then off and every collective runs on stream H like in round 1)
  PotrfEngine(const EngineGeometry& g, ncclComm_t row_comm, ncclComm_t col_comm, ncclComm_t row_comm_h = nullptr,
              ncclComm_t col_comm_h = nullptr);
  ~PotrfEngine();
  PotrfEngine(const PotrfEngine&) = delete;
  PotrfEngine& operator=(const PotrfEngine&) = delete;

  // --- geometry
  long n() const { return geo_.n; }
  int nb() const { return geo_.nb; }
  int nbp() const { return nbp_; }
  int nt() const { return nt_; }
  int local_tile_rows() const { return ltr_; }
  int local_tile_cols() const { return ltc_; }
  bool padded() const { return nbp_ != geo_.nb || geo_.n % geo_.nb != 0; }
  LayoutParams layout(long ldu, bool transposed) const;

  // --- storage: either the engine's own padded slab, or (no padding, 16-byte aligned, even ld) the
  // caller's device memory used in place.
  T* slab();
  long slab_ld() const { return ld_; }
  void bind_external(T* dev, long ld);
  void unbind_external();
  static bool can_run_in_place(const EngineGeometry& g, const void* dev, long ld);

  // --- boundary copies (device pointers in the reference's local layout)
  void load(const T* user, long ldu, bool transposed, cudaStream_t s);
  void store(T* user, long ldu, bool transposed, cudaStream_t s);

  // --- the factorization: asynchronous w.r.t. the host; ordered after everything already enqueued on
  // `s` and complete (for stream order purposes) when `s` reaches the point after this call.
  void factorize(cudaStream_t s);
  // Host-resident, pipelined (tiles without padding, lower): the referenced triangle of the caller's
  // HOST local matrix is uploaded in block-column chunks on a copy stream while the factorization is
  // already running (every kernel waits only for the chunks it touches), and every block column is
  // downloaded as soon as its panel is final. Replaces the MatrixMirror bracket of the reference's C
  // API (src/c_api/factorization/cholesky.h:48-53: whole-matrix H2D, factorization, whole-matrix D2H).
  void factorize_host(T* host, long ldh, cudaStream_t s);
  // LAPACK-style info of the last factorize (0 = success, k = leading minor of order k not positive
  // definite). Synchronises `s`.
  int info(cudaStream_t s);

  long launches() const { return launches_; }  // kernels launched by the last factorize()
  // Steps of the last factorize() whose trailing update fell back from the int8-digit engine to native fp64 because the
  // guard fired (valid after info()); -1 when the int8 engine is not in use.
  int guard_fallback_steps() const;

  // Result check of the miniapp on THIS grid (collective; engine_check.cu): max|A - F F^H| / max|A| over the referenced
  // triangle of the global matrix (reference: check_cholesky, miniapp/miniapp_cholesky.cpp:408-446, which rebuilds
  // L L^H tile by tile on the host and reduces the two max norms with MPI). a_user / f_user are DEVICE pointers to this
  // rank's local parts (original matrix / factor) in the user's layout (transposed = user holds the upper triangle).
  // Independent of the factorization's engines: plain native GEMMs (fp64: DMMA) on panel copies broadcast with NCCL.
  double residual(const T* a_user, long lda, const T* f_user, long ldf, bool transposed, ncclComm_t grid_comm,
                  cudaStream_t s);

  // Per-launch timing of the dominant kernel (the bulk trailing update on stream L) with CUDA events on
  // its own stream: enable before factorize(), read after the stream has been synchronised.
  void set_profiling(bool on) { profiling_ = on; }
  // out = {sum of launch durations [ms], algorithmic flops of those launches, number of launches}
  void read_profile(double out[3]);
  // Critical-path breakdown of the last factorize (profiling mode), summed over the steps, in ms:
  // out = {diag-tile update wait+run, diagonal tile factorization, diag broadcast, wait for column + panel TRSM,
  //        panel pack + broadcasts, steps}
  void read_chain_profile(double out[6]);

private:
  void panel_step(int k, bool wait_column);
  void panel_step_dist(int k, bool wait_column);  // P x Q > 1, two chains (engine.cu)
  void update_next_diag_from_crit(int k, cudaStream_t st);
  enum UpdatePart { kBulk = 0, kNextDiag = 1, kNextColumnRest = 2, kNextColumnFirst = 3, kNextColumnTail = 4 };
  void update(int k, UpdatePart part, cudaStream_t st);
  void launch_update(int k, int cj0, int ncols, int ri0, int mrows, bool count_flops, cudaStream_t st,
                     bool native = false);
  void factor_diag_tile(T* tile, long ld, T* w, int k, cudaStream_t st);
  void trsm_panel(T* b, long ldb, int m, const T* tkk, long ldt, const T* w, cudaStream_t st);
  void gemm(const GemmArgsT<T>& a, cudaStream_t st);
  int cnt_rows(long g_end) const;  // local row tiles with global index < g_end
  int cnt_cols(long g_end) const;
  int col_comm_rank(int vrow) const { return (vrow + geo_.src_in_col_comm) % geo_.P; }
  int row_comm_rank(int vcol) const { return (vcol + geo_.src_in_row_comm) % geo_.Q; }
  T* tile_ptr(int li, int lj) { return data_ + static_cast<long>(li) * nbp_ + static_cast<long>(lj) * nbp_ * ld_; }

  EngineGeometry geo_;
  ncclComm_t row_comm_, col_comm_;
  ncclComm_t row_comm_h_ = nullptr, col_comm_h_ = nullptr;
  // distributed two-chain schedule (DLAF_B200_SPLIT_CHAIN, default on when the clones exist)
  bool dist_split_ = false;
  T* crit_[2] = {nullptr, nullptr};      // tile (k+1, k) after its TRSM, broadcast along process row (k+1) % P
  const T* crit_next_ = nullptr;         // operand of the next diagonal-tile update on this rank (null: not the owner)
  long crit_next_ld_ = 0;
  cudaEvent_t evC1_[2] = {nullptr, nullptr};  // tile (k+2, k+1) updated with panel k (stream M)
  cudaEvent_t evPc_[2] = {nullptr, nullptr};  // panel k usable by the update of block column k+1 (stream R, <= evP)
  int nbp_, nt_, ltr_, ltc_, ns_;
  long ld_ = 0;
  T* data_ = nullptr;      // active storage (own slab or external)
  T* own_slab_ = nullptr;
  long own_ld_ = 0;
  bool external_ = false;

  cudaStream_t sH_ = nullptr, sM_ = nullptr, sL_ = nullptr;
  cudaEvent_t ev_start_ = nullptr, evP_[2] = {nullptr, nullptr}, evB_[2] = {nullptr, nullptr},
              evC_[2] = {nullptr, nullptr}, evD_[2] = {nullptr, nullptr};
  T* wbuf_[2] = {nullptr, nullptr};     // inverses of the diagonal blocks (1 x 1 column grid)
  T* diagbuf_[2] = {nullptr, nullptr};  // diagonal tile + inverses, broadcast down the process column
  T* panel_[2] = {nullptr, nullptr};    // column panel, tile-contiguous
  T* panelT_[2] = {nullptr, nullptr};   // transposed panel (tiles (j,k) for my local columns j)
  // host pipelining state (factorize_host)
  void issue_uploads();
  void wait_columns(int lj_end, cudaStream_t st);  // block columns [0, lj_end) resident before `st` goes on
  void download_column(int lj, cudaEvent_t final_ev);
  T* host_ = nullptr;
  long ldh_ = 0;
  cudaStream_t sIn_ = nullptr, sOut_ = nullptr;
  cudaEvent_t evOut_ = nullptr;
  std::vector<cudaEvent_t> evIn_;
  std::vector<int> in_end_;  // chunk c covers local block columns [in_end_[c-1], in_end_[c])
  // Bulk updates run per column chunk, each chunk on its own low-priority stream (chunk 0 = sL_): a chunk
  // only depends on its own history, so updates of early chunks proceed while later ones are still being
  // uploaded. Device-resident runs use a single chunk.
  std::vector<cudaStream_t> sLc_;
  std::vector<cudaEvent_t> evBc_;  // [oring_ * c + k % oring_]
  int chunk_of(int lj) const;
  int nchunks() const { return in_end_.empty() ? 1 : static_cast<int>(in_end_.size()); }
  void wait_bulk(int k, int lj, cudaStream_t st);  // bulk of step k done on the chunk of column lj (or on all)
  // fp32 only: tcgen05 3xTF32 trailing update (gemm_tf32_tcgen05.cu) — the panel of step k is split into
  // K-major hi/lo parts right after its TRSM (two round-robin slots like the panel workspaces)
  bool use_tf32_ = false;
  Tf32Split split_[2];   // column panel (rows = my local panel rows)
  Tf32Split splitT_[2];  // transposed panel (P > 1 only: tiles (j,k) for my local columns j)
  // fp64 only: tcgen05 int8 Ozaki-scheme trailing update (gemm_ozaki_i8.cu) — same life cycle as the TF32 splits
  bool use_ozaki_ = false;
  // 1 x 1 grid: a RING of kOzRing digit-plane slots instead of two, so that the panel chain may run that many steps ahead
  // of the slowest column chunk — in the host path the chunks that are still being uploaded — instead of stalling at
  // step 2 until the whole matrix has arrived (the planes are the only per-step workspace the lagging bulk updates read;
  // the fp64 panel itself is the final block column of the matrix).
  static constexpr int kOzRing = 16;
  int oring_ = 2;  // slots in use: kOzRing on a 1 x 1 grid with the int8 engine, else 2 (like every other workspace)
  OzakiSplit osplit_[kOzRing];
  OzakiSplit osplitT_[2];
  // guard of the int8 engine (gemm_ozaki.h): one device flag per step, raised by the digit split of that step's panel
  // when a row spans too many binades; the updates of a flagged step run on the native fp64 kernel instead
  int* oz_flag_ = nullptr;      // [nt]
  int* h_oz_flags_ = nullptr;   // pinned copy read by guard_fallback_steps()
  bool split_panels() const { return use_tf32_ || use_ozaki_; }
  // 1 x 1 grid, int8 engine: the panel of a step is finished on stream R while stream H already goes on with the next
  // diagonal tile (engine.cu: panel_step). DLAF_B200_SPLIT_CHAIN=0 keeps everything on stream H.
  bool split_chain_ = false;
  bool split_chain() const { return split_chain_; }
  cudaStream_t sR_ = nullptr;
  cudaEvent_t evF_[2] = {nullptr, nullptr}, evT1_[2] = {nullptr, nullptr};
  int* d_info_ = nullptr;
  int* h_info_ = nullptr;
  long launches_ = 0;
  bool profiling_ = false;
  std::vector<cudaEvent_t> prof_ev_;  // pairs (begin, end)
  std::vector<double> prof_flops_;
  size_t prof_used_ = 0;
  double last_update_flops_ = 0.0;
  std::vector<cudaEvent_t> chain_ev_;  // 6 stamps per step on stream H
  size_t chain_used_ = 0;
  std::vector<cudaEvent_t> diag_ev_;  // (begin, end) per diagonal-block kernel (profiling mode)
  size_t diag_used_ = 0;
  void chain_stamp(int k, int which);
};

}  // namespace dlaf_b200
