// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product: nothing under dla-future_b200/ may
// include, link or call this file. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / CPU baseline.
//
// CPU restatement of the reference's Cholesky path (eth-cscs/DLA-Future v0.10.0). The reference's MC
// backend cannot be built here (pika, blaspp, lapackpp, Umpire and MPI are REQUIRED by
// CMakeLists.txt:164-184 and absent), so kind = "port". The arithmetic of the reference lives in the
// vendor BLAS/LAPACK behind blaspp/lapackpp (include/dlaf/lapack/tile.h:453, include/dlaf/blas/tile.h:180,
// :210, :243); here the same routines come from OpenBLAS 0.3.30 shipped in the scipy wheel
// (Fortran symbols with the scipy_ prefix).
//
// Parity pinning: this restatement is checked against the reference's own closed-form golden
// vectors (test/include/dlaf_test/matrix/util_generic_lapack.h:39-68, sizes of
// test/unit/factorization/test_cholesky.cpp:54-58) in tests/test_oracle.py.
//
// What is restated, with the reference lines each function follows:
//   * cholesky_local<T>()      include/dlaf/factorization/cholesky/impl.h:150-189 (call_L, local)
//                              and :316-348 (call_U, local); tile ops :46-94 and :98-146
//   * the tile-op DAG + pool   per-tile FIFO read/readwrite order (matrix/internal/tile_pipeline.h:36-51),
//                              priorities impl.h:171-173, one BLAS thread per tile op
//                              (src/common/single_threaded_blas.cpp:24-36)
//   * set_random_hpd<T>()      include/dlaf/util_matrix.h:161-189, :335-389, :410-453, :529-531
//   * residual<T>()            miniapp/miniapp_cholesky.cpp:408-446 (max|A-LL^H| / max|A|, lower)
//   * triangular_solver_local  include/dlaf/solver/triangular/impl.h:236-480 (the eight local loop nests), pinned to the
//                              closed forms of test/include/dlaf_test/matrix/util_generic_blas.h:259-371
//   * triangular_inverse_local include/dlaf/inverse/triangular/impl.h:183-229 (call_L) and :367-413 (call_U), tile ops
//                              :47-110 / :112-170; pinned to test/include/dlaf_test/matrix/util_generic_lapack.h:261-327
//   * generalized_to_standard_local   include/dlaf/eigensolver/gen_to_std/impl.h:238-281 (call_L) and :507-568 (call_U), tile
//                              ops :43-146 / :148-233; pinned to util_generic_lapack.h:94-149 (itype 1), test table and
//                              tolerance of test/unit/eigensolver/test_gen_to_std.cpp:52-56, :78
//   * assemble_cholesky_inverse_local  include/dlaf/inverse/cholesky/impl.h:180-224 (call_L) and :361-405 (call_U), tile
//                              ops :46-108 / :110-172; pinned to util_generic_lapack.h:165-196 and, chained after the
//                              triangular inverse (include/dlaf/inverse/cholesky.h:38-52), to :212-247
#include <algorithm>
#include <atomic>
#include <cmath>
#include <complex>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <thread>
#include <vector>

extern "C" {
void scipy_openblas_set_num_threads(int);
int scipy_openblas_get_num_threads(void);
char* scipy_openblas_get_config(void);

#define DECL_REAL(p, T)                                                                          \
  void scipy_##p##potrf_(const char*, const int*, T*, const int*, int*);                         \
  void scipy_##p##trsm_(const char*, const char*, const char*, const char*, const int*,          \
                        const int*, const T*, const T*, const int*, T*, const int*);             \
  void scipy_##p##gemm_(const char*, const char*, const int*, const int*, const int*, const T*,  \
                        const T*, const int*, const T*, const int*, const T*, T*, const int*);     \
  void scipy_##p##trmm_(const char*, const char*, const char*, const char*, const int*,          \
                        const int*, const T*, const T*, const int*, T*, const int*);             \
  void scipy_##p##trtri_(const char*, const char*, const int*, T*, const int*, int*);            \
  void scipy_##p##lauum_(const char*, const int*, T*, const int*, int*);
DECL_REAL(s, float)
DECL_REAL(d, double)
DECL_REAL(c, std::complex<float>)
DECL_REAL(z, std::complex<double>)
void scipy_ssygst_(const int*, const char*, const int*, float*, const int*, const float*, const int*, int*);
void scipy_dsygst_(const int*, const char*, const int*, double*, const int*, const double*, const int*, int*);
void scipy_chegst_(const int*, const char*, const int*, std::complex<float>*, const int*, const std::complex<float>*,
                   const int*, int*);
void scipy_zhegst_(const int*, const char*, const int*, std::complex<double>*, const int*, const std::complex<double>*,
                   const int*, int*);
#define DECL_HE(p, hemm, her2k, T, R)                                                                                   \
  void scipy_##p##hemm##_(const char*, const char*, const int*, const int*, const T*, const T*, const int*, const T*,   \
                          const int*, const T*, T*, const int*);                                                        \
  void scipy_##p##her2k##_(const char*, const char*, const int*, const int*, const T*, const T*, const int*, const T*, \
                           const int*, const R*, T*, const int*);
DECL_HE(s, symm, syr2k, float, float)
DECL_HE(d, symm, syr2k, double, double)
DECL_HE(c, hemm, her2k, std::complex<float>, float)
DECL_HE(z, hemm, her2k, std::complex<double>, double)
void scipy_ssyrk_(const char*, const char*, const int*, const int*, const float*, const float*,
                  const int*, const float*, float*, const int*);
void scipy_dsyrk_(const char*, const char*, const int*, const int*, const double*, const double*,
                  const int*, const double*, double*, const int*);
void scipy_cherk_(const char*, const char*, const int*, const int*, const float*,
                  const std::complex<float>*, const int*, const float*, std::complex<float>*,
                  const int*);
void scipy_zherk_(const char*, const char*, const int*, const int*, const double*,
                  const std::complex<double>*, const int*, const double*, std::complex<double>*,
                  const int*);
}

namespace {

template <class T>
struct Base {
  using type = T;
};
template <class T>
struct Base<std::complex<T>> {
  using type = T;
};
template <class T>
using BaseT = typename Base<T>::type;

// ---- BLAS/LAPACK dispatch -------------------------------------------------------------------
inline int potrf(char uplo, int n, float* a, int lda) {
  int info;
  scipy_spotrf_(&uplo, &n, a, &lda, &info);
  return info;
}
inline int potrf(char uplo, int n, double* a, int lda) {
  int info;
  scipy_dpotrf_(&uplo, &n, a, &lda, &info);
  return info;
}
inline int potrf(char uplo, int n, std::complex<float>* a, int lda) {
  int info;
  scipy_cpotrf_(&uplo, &n, a, &lda, &info);
  return info;
}
inline int potrf(char uplo, int n, std::complex<double>* a, int lda) {
  int info;
  scipy_zpotrf_(&uplo, &n, a, &lda, &info);
  return info;
}
#define TRSM_IMPL(p, T)                                                                           \
  inline void trsm(char side, char uplo, char op, char diag, int m, int n, T alpha, const T* a,   \
                   int lda, T* b, int ldb) {                                                      \
    scipy_##p##trsm_(&side, &uplo, &op, &diag, &m, &n, &alpha, a, &lda, b, &ldb);                 \
  }                                                                                               \
  inline void gemm(char opa, char opb, int m, int n, int k, T alpha, const T* a, int lda,         \
                   const T* b, int ldb, T beta, T* c, int ldc) {                                  \
    scipy_##p##gemm_(&opa, &opb, &m, &n, &k, &alpha, a, &lda, b, &ldb, &beta, c, &ldc);           \
  }
TRSM_IMPL(s, float)
TRSM_IMPL(d, double)
TRSM_IMPL(c, std::complex<float>)
TRSM_IMPL(z, std::complex<double>)
inline void herk(char uplo, char op, int n, int k, float alpha, const float* a, int lda, float beta,
                 float* c, int ldc) {
  scipy_ssyrk_(&uplo, &op, &n, &k, &alpha, a, &lda, &beta, c, &ldc);
}
inline void herk(char uplo, char op, int n, int k, double alpha, const double* a, int lda,
                 double beta, double* c, int ldc) {
  scipy_dsyrk_(&uplo, &op, &n, &k, &alpha, a, &lda, &beta, c, &ldc);
}
inline void herk(char uplo, char op, int n, int k, float alpha, const std::complex<float>* a,
                 int lda, float beta, std::complex<float>* c, int ldc) {
  scipy_cherk_(&uplo, &op, &n, &k, &alpha, a, &lda, &beta, c, &ldc);
}
inline void herk(char uplo, char op, int n, int k, double alpha, const std::complex<double>* a,
                 int lda, double beta, std::complex<double>* c, int ldc) {
  scipy_zherk_(&uplo, &op, &n, &k, &alpha, a, &lda, &beta, c, &ldc);
}
template <class T>
constexpr char conj_trans() {
  // For real T, ConjTrans == Trans (include/dlaf/gpu/blas/gpublas.h:107-112)
  return std::is_same_v<T, BaseT<T>> ? 'T' : 'C';
}

// ---- tile task DAG --------------------------------------------------------------------------
// The reference never states dependencies: they follow from the FIFO order in which read() /
// readwrite() senders are requested from each tile (tile_pipeline.h:36-51). The same rule here:
// a task depends on the last writer of every tile it touches, and a writer additionally on every
// reader since that write.
struct Task {
  std::function<void()> fn;
  int priority = 0;  // 1 = high (impl.h:171-173: first trailing column)
  std::vector<int> succ;
  std::atomic<int> deps{0};
};

class TaskGraph {
public:
  explicit TaskGraph(int ntiles) : last_writer_(ntiles, -1), readers_(ntiles) {}

  void add(std::function<void()> fn, int priority, const std::vector<int>& reads,
           const std::vector<int>& writes) {
    const int id = static_cast<int>(tasks_.size());
    tasks_.emplace_back(new Task);
    Task& t = *tasks_.back();
    t.fn = std::move(fn);
    t.priority = priority;
    std::vector<int> deps;
    for (int tile : reads) {
      if (last_writer_[tile] >= 0)
        deps.push_back(last_writer_[tile]);
      readers_[tile].push_back(id);
    }
    for (int tile : writes) {
      if (last_writer_[tile] >= 0)
        deps.push_back(last_writer_[tile]);
      for (int r : readers_[tile])
        if (r != id)
          deps.push_back(r);
      readers_[tile].clear();
      last_writer_[tile] = id;
    }
    std::sort(deps.begin(), deps.end());
    deps.erase(std::unique(deps.begin(), deps.end()), deps.end());
    t.deps = static_cast<int>(deps.size());
    for (int d : deps)
      tasks_[d]->succ.push_back(id);
  }

  // Executes all tasks on nthreads workers: ready tasks by (priority, issue order).
  void run(int nthreads) {
    if (nthreads <= 1) {
      for (auto& t : tasks_)
        t->fn();  // issue order is a valid topological order
      return;
    }
    using Item = std::pair<int, int>;  // (-priority, id) -> min-heap
    std::priority_queue<Item, std::vector<Item>, std::greater<Item>> ready;
    std::mutex m;
    std::condition_variable cv;
    size_t done = 0;
    for (size_t i = 0; i < tasks_.size(); ++i)
      if (tasks_[i]->deps == 0)
        ready.push({-tasks_[i]->priority, static_cast<int>(i)});
    auto worker = [&]() {
      std::unique_lock<std::mutex> lk(m);
      while (true) {
        cv.wait(lk, [&] { return !ready.empty() || done == tasks_.size(); });
        if (ready.empty())
          return;
        const int id = ready.top().second;
        ready.pop();
        lk.unlock();
        tasks_[id]->fn();
        lk.lock();
        ++done;
        for (int s : tasks_[id]->succ)
          if (--tasks_[s]->deps == 0)
            ready.push({-tasks_[s]->priority, s});
        cv.notify_all();
      }
    };
    std::vector<std::thread> pool;
    for (int i = 0; i < nthreads; ++i)
      pool.emplace_back(worker);
    for (auto& th : pool)
      th.join();
  }

private:
  std::vector<std::unique_ptr<Task>> tasks_;
  std::vector<int> last_writer_;
  std::vector<std::vector<int>> readers_;
};

// ---- local Cholesky, tile by tile -----------------------------------------------------------
template <class T>
int cholesky_local(char uplo, long n, long nb, T* a, long lda, int nthreads) {
  if (n == 0)
    return 0;
  const int nt = static_cast<int>((n + nb - 1) / nb);
  auto ts = [&](int i) { return static_cast<int>(std::min<long>(nb, n - i * nb)); };
  auto tile = [&](int i, int j) { return a + i * nb + j * nb * lda; };
  auto id = [&](int i, int j) { return i + j * nt; };
  const int ld = static_cast<int>(lda);
  const char CT = conj_trans<T>();
  std::atomic<int> info{0};
  TaskGraph g(nt * nt);
  const bool lower = (uplo == 'L' || uplo == 'l');

  for (int k = 0; k < nt; ++k) {
    // potrfDiagTile (impl.h:46-53 / :98-105), high priority
    g.add(
        [=, &info] {
          int r = potrf(lower ? 'L' : 'U', ts(k), tile(k, k), ld);
          int expected = 0;
          if (r != 0)
            info.compare_exchange_strong(expected, k * static_cast<int>(nb) + r);
        },
        1, {}, {id(k, k)});
    if (lower) {
      for (int i = k + 1; i < nt; ++i)  // trsmPanelTile (impl.h:55-67): Right, Lower, ConjTrans, NonUnit
        g.add([=] { trsm('R', 'L', CT, 'N', ts(i), ts(k), T(1), tile(k, k), ld, tile(i, k), ld); }, 1,
              {id(k, k)}, {id(i, k)});
      for (int j = k + 1; j < nt; ++j) {
        const int prio = (j == k + 1) ? 1 : 0;  // impl.h:171-173
        // herkTrailingDiagTile (impl.h:69-80): Lower, NoTrans, -1, 1
        g.add([=] { herk('L', 'N', ts(j), ts(k), BaseT<T>(-1), tile(j, k), ld, BaseT<T>(1), tile(j, j), ld); },
              prio, {id(j, k)}, {id(j, j)});
        for (int i = j + 1; i < nt; ++i)  // gemmTrailingMatrixTile (impl.h:82-94): NoTrans, ConjTrans
          g.add([=] { gemm('N', CT, ts(i), ts(j), ts(k), T(-1), tile(i, k), ld, tile(j, k), ld, T(1), tile(i, j), ld); },
                prio, {id(i, k), id(j, k)}, {id(i, j)});
      }
    }
    else {
      for (int j = k + 1; j < nt; ++j)  // trsmPanelTile (impl.h:107-119): Left, Upper, ConjTrans, NonUnit
        g.add([=] { trsm('L', 'U', CT, 'N', ts(k), ts(j), T(1), tile(k, k), ld, tile(k, j), ld); }, 1,
              {id(k, k)}, {id(k, j)});
      for (int i = k + 1; i < nt; ++i) {
        const int prio = (i == k + 1) ? 1 : 0;
        // herkTrailingDiagTile (impl.h:121-132): Upper, ConjTrans
        g.add([=] { herk('U', CT, ts(i), ts(k), BaseT<T>(-1), tile(k, i), ld, BaseT<T>(1), tile(i, i), ld); },
              prio, {id(k, i)}, {id(i, i)});
        for (int j = i + 1; j < nt; ++j)  // gemmTrailingMatrixTile (impl.h:134-146): ConjTrans, NoTrans
          g.add([=] { gemm(CT, 'N', ts(i), ts(j), ts(k), T(-1), tile(k, i), ld, tile(k, j), ld, T(1), tile(i, j), ld); },
                prio, {id(k, i), id(k, j)}, {id(i, j)});
      }
    }
  }
  const int prev = scipy_openblas_get_num_threads();
  scipy_openblas_set_num_threads(1);  // SingleThreadedBlasScope: one BLAS thread per tile op
  g.run(nthreads);
  scipy_openblas_set_num_threads(prev);
  return info.load();
}

// ---- random Hermitian positive definite matrix ---------------------------------------------
template <class T>
struct Rand {  // getter_random (util_matrix.h:161-177)
  explicit Rand(long seed) : eng(static_cast<std::size_t>(seed)) {}
  T operator()() { return dist(eng); }
  std::mt19937_64 eng;
  std::uniform_real_distribution<T> dist{-1, 1};
};
template <class T>
struct Rand<std::complex<T>> : Rand<T> {  // util_matrix.h:180-189
  using Rand<T>::Rand;
  std::complex<T> operator()() {
    // NOTE: evaluation order of the two draws follows the argument evaluation order gcc uses for the
    // reference expression std::polar(abs(r()), pi * r()), i.e. right-to-left: angle first, radius second.
    T angle = static_cast<T>(M_PI) * Rand<T>::operator()();
    T radius = std::abs(Rand<T>::operator()());
    return std::polar<T>(radius, angle);
  }
};
template <class T>
T cj(T v) {
  return v;
}
template <class T>
std::complex<T> cj(std::complex<T> v) {
  return std::conj(v);
}
template <class T>
T from_real(BaseT<T> v) {
  return T(v);
}
template <class T>
BaseT<T> re(T v) {
  return std::real(v);
}

template <class T>
void set_random_hpd(long n, long nb, T* a, long lda) {
  const long nt = (n + nb - 1) / nb;
  const long offset = 2 * n;  // util_matrix.h:529-531
  auto one = [&](long ti, long tj) {
    const long r0 = ti * nb, c0 = tj * nb;
    const long tr = std::min(nb, n - r0), tc = std::min(nb, n - c0);
    T* t = a + r0 + c0 * lda;
    // util_matrix.h:435-439: same seed for a tile and its transposed twin
    const long seed = (ti >= tj) ? c0 + r0 * n : r0 + c0 * n;
    Rand<T> rnd(seed);
    if (ti == tj) {  // util_matrix.h:337-349
      for (long j = 0; j < tc; ++j) {
        for (long i = 0; i < j; ++i) {
          T v = rnd();
          t[i + j * lda] = v;
          t[j + i * lda] = cj(v);
        }
        t[j + j * lda] = from_real<T>(re(rnd()) + BaseT<T>(offset));
      }
    }
    else {  // util_matrix.h:351-389: values drawn over the FULL tile size, column-major
      for (long j = 0; j < nb; ++j)
        for (long i = 0; i < nb; ++i) {
          T v = rnd();
          if (ti > tj) {
            if (i < tr && j < tc)
              t[i + j * lda] = v;
          }
          else {
            if (j < tr && i < tc)
              t[j + i * lda] = cj(v);
          }
        }
    }
  };
  std::vector<std::thread> pool;
  std::atomic<long> next{0};
  const int nthr = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  for (int w = 0; w < nthr; ++w)
    pool.emplace_back([&] {
      for (long idx; (idx = next++) < nt * nt;)
        one(idx % nt, idx / nt);
    });
  for (auto& th : pool)
    th.join();
}

// ---- miniapp residual check ---------------------------------------------------------------
template <class T>
double residual(char uplo, long n, const T* a, long lda, const T* f, long ldf) {
  // max over the uplo triangle of |A - L L^H| (or |A - U^H U|) divided by max |A| over that triangle
  // (miniapp_cholesky.cpp:414-445; the reference only checks 'L', 'U' is the mirrored statement).
  if (n == 0)
    return 0.0;
  const bool lower = (uplo == 'L' || uplo == 'l');
  std::vector<T> fac(static_cast<size_t>(n) * n, T(0));
  for (long j = 0; j < n; ++j)
    for (long i = 0; i < n; ++i)
      if (lower ? i >= j : i <= j)
        fac[i + j * n] = f[i + j * ldf];
  std::vector<T> prod(static_cast<size_t>(n) * n);
  const int ni = static_cast<int>(n);
  if (lower)
    gemm('N', conj_trans<T>(), ni, ni, ni, T(1), fac.data(), ni, fac.data(), ni, T(0), prod.data(), ni);
  else
    gemm(conj_trans<T>(), 'N', ni, ni, ni, T(1), fac.data(), ni, fac.data(), ni, T(0), prod.data(), ni);
  double max_a = 0, max_d = 0;
  for (long j = 0; j < n; ++j)
    for (long i = 0; i < n; ++i)
      if (lower ? i >= j : i <= j) {
        max_a = std::max<double>(max_a, std::abs(a[i + j * lda]));
        max_d = std::max<double>(max_d, std::abs(a[i + j * lda] - prod[i + j * n]));
      }
  return max_d / max_a;
}

// ---- triangular solver --------------------------------------------------------------------------
// Restatement of the reference's LOCAL tile loops Triangular<B,D,T>::call_{LLN,LLT,LUN,LUT,RLN,RLT,RUN,RUT}
// (include/dlaf/solver/triangular/impl.h:236-480): per step one tile TRSM per tile of the k-th row (Left) / column
// (Right) panel of B — with alpha applied there — and GEMM updates of the remaining tiles with beta = -1 / alpha
// (impl.h:257, :287, ...). Tile operations: blas::trsm / blas::gemm (blas/tile.h:175-182, :238-245) = OpenBLAS here.
// B is m x n in tiles mb x nb; A is square of order m (Left, tiles mb) or n (Right, tiles nb).
template <class T>
void triangular_solver_local(char side, char uplo, char op, char diag, T alpha, long m, long n, long mb, long nb,
                             const T* a, long lda, T* b, long ldb) {
  if (m == 0 || n == 0)
    return;
  const bool left = (side == 'L');
  const bool lower = (uplo == 'L');
  const bool notrans = (op == 'N');
  const long mt = (m + mb - 1) / mb, nt = (n + nb - 1) / nb;
  auto rows = [&](long i) { return static_cast<int>(std::min(mb, m - i * mb)); };
  auto cols = [&](long j) { return static_cast<int>(std::min(nb, n - j * nb)); };
  auto B = [&](long i, long j) { return b + i * mb + j * nb * ldb; };
  const long ba = left ? mb : nb;
  auto A = [&](long i, long j) { return a + i * ba + j * ba * lda; };
  const T beta = T(-1) / alpha;
  const int la = static_cast<int>(lda), lb = static_cast<int>(ldb);
  // op(A) lower  <=>  (uplo == Lower) == (op == NoTrans)
  const bool op_a_lower = (lower == notrans);
  if (left) {
    // op(A) X = alpha B: forward over the row panels when op(A) is lower (call_LLN / call_LUT), backward otherwise
    const bool fwd = op_a_lower;
    for (long s = 0; s < mt; ++s) {
      const long k = fwd ? s : mt - 1 - s;
      for (long j = 0; j < nt; ++j) {
        trsm('L', uplo, op, diag, rows(k), cols(j), alpha, A(k, k), la, B(k, j), lb);
        for (long i = fwd ? k + 1 : 0; i < (fwd ? mt : k); ++i) {
          // op(A)(i,k): stored tile (i,k) for NoTrans, (k,i) otherwise
          if (notrans)
            gemm('N', 'N', rows(i), cols(j), rows(k), beta, A(i, k), la, B(k, j), lb, T(1), B(i, j), lb);
          else
            gemm(op, 'N', rows(i), cols(j), rows(k), beta, A(k, i), la, B(k, j), lb, T(1), B(i, j), lb);
        }
      }
    }
  }
  else {
    // X op(A) = alpha B: forward over the column panels when op(A) is upper (call_RUN / call_RLT), backward otherwise
    const bool fwd = !op_a_lower;
    for (long s = 0; s < nt; ++s) {
      const long k = fwd ? s : nt - 1 - s;
      for (long i = 0; i < mt; ++i) {
        trsm('R', uplo, op, diag, rows(i), cols(k), alpha, A(k, k), la, B(i, k), lb);
        for (long j = fwd ? k + 1 : 0; j < (fwd ? nt : k); ++j) {
          // op(A)(k,j): stored tile (k,j) for NoTrans, (j,k) otherwise
          if (notrans)
            gemm('N', 'N', rows(i), cols(j), cols(k), beta, B(i, k), lb, A(k, j), la, T(1), B(i, j), lb);
          else
            gemm('N', op, rows(i), cols(j), cols(k), beta, B(i, k), lb, A(j, k), la, T(1), B(i, j), lb);
        }
      }
    }
  }
}


#define WRAP_INV(p, T)                                                                                        \
  inline void trmm(char side, char uplo, char op, char diag, int m, int n, T alpha, const T* a, int lda, T* b, \
                   int ldb) {                                                                                 \
    scipy_##p##trmm_(&side, &uplo, &op, &diag, &m, &n, &alpha, a, &lda, b, &ldb);                             \
  }                                                                                                           \
  inline int trtri(char uplo, char diag, int n, T* a, int lda) {                                              \
    int info = 0;                                                                                             \
    scipy_##p##trtri_(&uplo, &diag, &n, a, &lda, &info);                                                      \
    return info;                                                                                              \
  }                                                                                                           \
  inline int lauum(char uplo, int n, T* a, int lda) {                                                         \
    int info = 0;                                                                                             \
    scipy_##p##lauum_(&uplo, &n, a, &lda, &info);                                                             \
    return info;                                                                                              \
  }
WRAP_INV(s, float)
WRAP_INV(d, double)
WRAP_INV(c, std::complex<float>)
WRAP_INV(z, std::complex<double>)

// ---- inverse ------------------------------------------------------------------------------------
// Restatement of Triangular<B,D,T>::call_L / call_U (LOCAL), include/dlaf/inverse/triangular/impl.h:183-229, :367-413:
// k from the last tile to the first; column (L) / row (U) panel tile TRSM with alpha = -1 against the still original
// diagonal tile, GEMM of the trailing tiles, row (L) / column (U) panel TRSM with alpha = 1, tile trtri last.
template <class T>
void triangular_inverse_local(char uplo, char diag, long n, long nb, T* a, long lda) {
  if (n == 0)
    return;
  const long nt = (n + nb - 1) / nb;
  auto sz = [&](long i) { return static_cast<int>(std::min(nb, n - i * nb)); };
  auto A = [&](long i, long j) { return a + i * nb + j * nb * lda; };
  const int la = static_cast<int>(lda);
  for (long k = nt - 1; k >= 0; --k) {
    if (uplo == 'L') {
      for (long i = k + 1; i < nt; ++i) {
        trsm('R', 'L', 'N', diag, sz(i), sz(k), T(-1), A(k, k), la, A(i, k), la);  // impl.h:78-90
        for (long j = 0; j < k; ++j)
          gemm('N', 'N', sz(i), sz(j), sz(k), T(1), A(i, k), la, A(k, j), la, T(1), A(i, j), la);  // :92-109
      }
      for (long j = 0; j < k; ++j)
        trsm('L', 'L', 'N', diag, sz(k), sz(j), T(1), A(k, k), la, A(k, j), la);  // :64-76
      trtri('L', diag, sz(k), A(k, k), la);
    }
    else {
      for (long j = k + 1; j < nt; ++j) {
        trsm('L', 'U', 'N', diag, sz(k), sz(j), T(-1), A(k, k), la, A(k, j), la);  // impl.h:143-155
        for (long i = 0; i < k; ++i)
          gemm('N', 'N', sz(i), sz(j), sz(k), T(1), A(i, k), la, A(k, j), la, T(1), A(i, j), la);
      }
      for (long i = 0; i < k; ++i)
        trsm('R', 'U', 'N', diag, sz(i), sz(k), T(1), A(k, k), la, A(i, k), la);  // :129-141
      trtri('U', diag, sz(k), A(k, k), la);
    }
  }
}

// Restatement of AssembleCholeskyInverse<B,D,T>::call_L / call_U (LOCAL), include/dlaf/inverse/cholesky/impl.h:180-224,
// :361-405: k ascending; GEMM / HERK of the leading tiles with row (L) / column (U) k, TRMM of that panel with the
// diagonal tile, tile lauum last.
template <class T>
void assemble_cholesky_inverse_local(char uplo, long n, long nb, T* a, long lda) {
  if (n == 0)
    return;
  using R = BaseT<T>;
  const long nt = (n + nb - 1) / nb;
  auto sz = [&](long i) { return static_cast<int>(std::min(nb, n - i * nb)); };
  auto A = [&](long i, long j) { return a + i * nb + j * nb * lda; };
  const int la = static_cast<int>(lda);
  for (long k = 0; k < nt; ++k) {
    if (uplo == 'L') {
      for (long i = 0; i < k; ++i) {
        for (long j = 0; j < i; ++j)
          gemm('C', 'N', sz(i), sz(j), sz(k), T(1), A(k, i), la, A(k, j), la, T(1), A(i, j), la);  // impl.h:96-108
        herk('L', 'C', sz(i), sz(k), R(1), A(k, i), la, R(1), A(i, i), la);                        // :83-94
      }
      for (long j = 0; j < k; ++j)
        trmm('L', 'L', 'C', 'N', sz(k), sz(j), T(1), A(k, k), la, A(k, j), la);  // :69-81
      lauum('L', sz(k), A(k, k), la);
    }
    else {
      for (long j = 0; j < k; ++j) {
        for (long i = 0; i < j; ++i)
          gemm('N', 'C', sz(i), sz(j), sz(k), T(1), A(i, k), la, A(j, k), la, T(1), A(i, j), la);  // impl.h:158-170
        herk('U', 'N', sz(j), sz(k), R(1), A(j, k), la, R(1), A(j, j), la);                        // :145-156
      }
      for (long i = 0; i < k; ++i)
        trmm('R', 'U', 'C', 'N', sz(i), sz(k), T(1), A(k, k), la, A(i, k), la);  // :131-143
      lauum('U', sz(k), A(k, k), la);
    }
  }
}


inline int hegst(int itype, char uplo, int n, float* a, int lda, const float* b, int ldb) {
  int info = 0;
  scipy_ssygst_(&itype, &uplo, &n, a, &lda, b, &ldb, &info);
  return info;
}
inline int hegst(int itype, char uplo, int n, double* a, int lda, const double* b, int ldb) {
  int info = 0;
  scipy_dsygst_(&itype, &uplo, &n, a, &lda, b, &ldb, &info);
  return info;
}
inline int hegst(int itype, char uplo, int n, std::complex<float>* a, int lda, const std::complex<float>* b, int ldb) {
  int info = 0;
  scipy_chegst_(&itype, &uplo, &n, a, &lda, b, &ldb, &info);
  return info;
}
inline int hegst(int itype, char uplo, int n, std::complex<double>* a, int lda, const std::complex<double>* b, int ldb) {
  int info = 0;
  scipy_zhegst_(&itype, &uplo, &n, a, &lda, b, &ldb, &info);
  return info;
}
#define WRAP_HE(p, hemm_, her2k_, T, R)                                                                              \
  inline void hemm(char side, char uplo, int m, int n, T alpha, const T* a, int lda, const T* b, int ldb, T beta,   \
                   T* c, int ldc) {                                                                                  \
    scipy_##p##hemm_##_(&side, &uplo, &m, &n, &alpha, a, &lda, b, &ldb, &beta, c, &ldc);                             \
  }                                                                                                                  \
  inline void her2k(char uplo, char op, int n, int k, T alpha, const T* a, int lda, const T* b, int ldb, R beta,    \
                    T* c, int ldc) {                                                                                 \
    scipy_##p##her2k_##_(&uplo, &op, &n, &k, &alpha, a, &lda, b, &ldb, &beta, c, &ldc);                              \
  }
WRAP_HE(s, symm, syr2k, float, float)
WRAP_HE(d, symm, syr2k, double, double)
WRAP_HE(c, hemm, her2k, std::complex<float>, float)
WRAP_HE(z, hemm, her2k, std::complex<double>, double)

// ---- generalized to standard ------------------------------------------------------------------
// Restatement of GenToStd<B,D,T>::call_L / call_U (LOCAL), include/dlaf/eigensolver/gen_to_std/impl.h:238-281, :507-568
// (the blocked xHEGST, itype 1): per step the tile hegst, the panel TRSM + first HEMM, the HER2K / two GEMMs of the
// trailing matrix, the second HEMM and the panel solve against the trailing factor.
template <class T>
void generalized_to_standard_local(char uplo, long n, long nb, T* a, long lda, const T* l, long ldl) {
  if (n == 0)
    return;
  using R = BaseT<T>;
  const char CT = std::is_same_v<T, R> ? 'T' : 'C';
  const long nt = (n + nb - 1) / nb;
  auto sz = [&](long i) { return static_cast<int>(std::min(nb, n - i * nb)); };
  auto A = [&](long i, long j) { return a + i * nb + j * nb * lda; };
  auto L = [&](long i, long j) { return l + i * nb + j * nb * ldl; };
  const int la = static_cast<int>(lda), ll = static_cast<int>(ldl);
  for (long k = 0; k < nt; ++k) {
    hegst(1, uplo, sz(k), A(k, k), la, L(k, k), ll);  // impl.h:43-51, :245
    if (k == nt - 1)
      continue;
    if (uplo == 'L') {
      for (long i = k + 1; i < nt; ++i) {
        trsm('R', 'L', CT, 'N', sz(i), sz(k), T(1), L(k, k), ll, A(i, k), la);                              // :53-64
        hemm('R', 'L', sz(i), sz(k), T(-0.5), A(k, k), la, L(i, k), ll, T(1), A(i, k), la);                 // :66-78
      }
      for (long j = k + 1; j < nt; ++j) {
        her2k('L', 'N', sz(j), sz(k), T(-1), A(j, k), la, L(j, k), ll, R(1), A(j, j), la);                  // :80-92
        for (long i = j + 1; i < nt; ++i) {
          gemm('N', CT, sz(i), sz(j), sz(k), T(-1), A(i, k), la, L(j, k), ll, T(1), A(i, j), la);           // :94-106
          gemm('N', CT, sz(i), sz(j), sz(k), T(-1), L(i, k), ll, A(j, k), la, T(1), A(i, j), la);
        }
      }
      for (long i = k + 1; i < nt; ++i)
        hemm('R', 'L', sz(i), sz(k), T(-0.5), A(k, k), la, L(i, k), ll, T(1), A(i, k), la);
      for (long j = k + 1; j < nt; ++j) {
        trsm('L', 'L', 'N', 'N', sz(j), sz(k), T(1), L(j, j), ll, A(j, k), la);                             // :108-119
        for (long i = j + 1; i < nt; ++i)
          gemm('N', 'N', sz(i), sz(k), sz(j), T(-1), L(i, j), ll, A(j, k), la, T(1), A(i, k), la);          // :121-133
      }
    }
    else {
      for (long i = k + 1; i < nt; ++i) {
        trsm('L', 'U', CT, 'N', sz(k), sz(i), T(1), L(k, k), ll, A(k, i), la);                              // :158-169
        hemm('L', 'U', sz(k), sz(i), T(-0.5), A(k, k), la, L(k, i), ll, T(1), A(k, i), la);                 // :171-183
      }
      for (long i = k + 1; i < nt; ++i) {
        her2k('U', CT, sz(i), sz(k), T(-1), A(k, i), la, L(k, i), ll, R(1), A(i, i), la);                   // :185-197
        for (long j = i + 1; j < nt; ++j) {
          gemm(CT, 'N', sz(i), sz(j), sz(k), T(-1), A(k, i), la, L(k, j), ll, T(1), A(i, j), la);           // :199-211
          gemm(CT, 'N', sz(i), sz(j), sz(k), T(-1), L(k, i), ll, A(k, j), la, T(1), A(i, j), la);
        }
      }
      for (long i = k + 1; i < nt; ++i)
        hemm('L', 'U', sz(k), sz(i), T(-0.5), A(k, k), la, L(k, i), ll, T(1), A(k, i), la);
      for (long i = k + 1; i < nt; ++i) {
        trsm('R', 'U', 'N', 'N', sz(k), sz(i), T(1), L(i, i), ll, A(k, i), la);                             // :213-224
        for (long j = i + 1; j < nt; ++j)
          gemm('N', 'N', sz(k), sz(j), sz(i), T(-1), A(k, i), la, L(i, j), ll, T(1), A(k, j), la);          // :226-238
      }
    }
  }
}

}  // namespace

#define EXPORT_TYPE(sfx, T)                                                                      \
  extern "C" int oracle_cholesky_local_##sfx(char uplo, long n, long nb, void* a, long lda,      \
                                             int nthreads) {                                     \
    return cholesky_local<T>(uplo, n, nb, static_cast<T*>(a), lda, nthreads);                    \
  }                                                                                              \
  extern "C" int oracle_lapack_potrf_##sfx(char uplo, long n, void* a, long lda, int nthreads) { \
    const int prev = scipy_openblas_get_num_threads();                                           \
    scipy_openblas_set_num_threads(nthreads);                                                    \
    int info = potrf(uplo, static_cast<int>(n), static_cast<T*>(a), static_cast<int>(lda));      \
    scipy_openblas_set_num_threads(prev);                                                        \
    return info;                                                                                 \
  }                                                                                              \
  extern "C" void oracle_triangular_solver_##sfx(char side, char uplo, char op, char diag,       \
                                                 const void* alpha, long m, long n, long mb,     \
                                                 long nb, const void* a, long lda, void* b,      \
                                                 long ldb) {                                     \
    triangular_solver_local<T>(side, uplo, op, diag, *static_cast<const T*>(alpha), m, n, mb, nb, \
                               static_cast<const T*>(a), lda, static_cast<T*>(b), ldb);          \
  }                                                                                              \
  extern "C" void oracle_generalized_to_standard_##sfx(char uplo, long n, long nb, void* a, long lda, \
                                                       const void* l, long ldl) {                 \
    generalized_to_standard_local<T>(uplo, n, nb, static_cast<T*>(a), lda, static_cast<const T*>(l), ldl); \
  }                                                                                              \
  extern "C" void oracle_triangular_inverse_##sfx(char uplo, char diag, long n, long nb, void* a, \
                                                  long lda) {                                    \
    triangular_inverse_local<T>(uplo, diag, n, nb, static_cast<T*>(a), lda);                     \
  }                                                                                              \
  extern "C" void oracle_assemble_cholesky_inverse_##sfx(char uplo, long n, long nb, void* a,    \
                                                         long lda) {                             \
    assemble_cholesky_inverse_local<T>(uplo, n, nb, static_cast<T*>(a), lda);                    \
  }                                                                                              \
  extern "C" void oracle_inverse_from_cholesky_factor_##sfx(char uplo, long n, long nb, void* a, \
                                                            long lda) {                          \
    triangular_inverse_local<T>(uplo, 'N', n, nb, static_cast<T*>(a), lda);                      \
    assemble_cholesky_inverse_local<T>(uplo, n, nb, static_cast<T*>(a), lda);                    \
  }                                                                                              \
  extern "C" void oracle_set_random_hpd_##sfx(long n, long nb, void* a, long lda) {              \
    set_random_hpd<T>(n, nb, static_cast<T*>(a), lda);                                           \
  }                                                                                              \
  extern "C" double oracle_residual_##sfx(char uplo, long n, const void* a, long lda,            \
                                          const void* f, long ldf) {                             \
    return residual<T>(uplo, n, static_cast<const T*>(a), lda, static_cast<const T*>(f), ldf);   \
  }

// The four tile operations exactly as cholesky_local() calls them (wrappers above), exported so that the tests can
// pin their argument conventions against the reference's own tile known-answer tests
// (test/unit/test_blas_tile/test_{gemm,herk,trsm}.h, test/unit/test_lapack_tile/test_potrf.h).
#define EXPORT_TILE_OPS(sfx, T)                                                                                     \
  extern "C" void oracle_tile_trsm_##sfx(char side, char uplo, char op, char diag, int m, int n, const void* alpha, \
                                         const void* a, int lda, void* b, int ldb) {                                \
    trsm(side, uplo, op, diag, m, n, *static_cast<const T*>(alpha), static_cast<const T*>(a), lda,                  \
         static_cast<T*>(b), ldb);                                                                                  \
  }                                                                                                                 \
  extern "C" void oracle_tile_gemm_##sfx(char opa, char opb, int m, int n, int k, const void* alpha, const void* a, \
                                         int lda, const void* b, int ldb, const void* beta, void* c, int ldc) {     \
    gemm(opa, opb, m, n, k, *static_cast<const T*>(alpha), static_cast<const T*>(a), lda,                           \
         static_cast<const T*>(b), ldb, *static_cast<const T*>(beta), static_cast<T*>(c), ldc);                     \
  }                                                                                                                 \
  extern "C" void oracle_tile_herk_##sfx(char uplo, char op, int n, int k, double alpha, const void* a, int lda,    \
                                         double beta, void* c, int ldc) {                                           \
    herk(uplo, op, n, k, static_cast<BaseT<T>>(alpha), static_cast<const T*>(a), lda, static_cast<BaseT<T>>(beta),  \
         static_cast<T*>(c), ldc);                                                                                  \
  }
EXPORT_TILE_OPS(s, float)
EXPORT_TILE_OPS(d, double)
EXPORT_TILE_OPS(c, std::complex<float>)
EXPORT_TILE_OPS(z, std::complex<double>)

EXPORT_TYPE(s, float)
EXPORT_TYPE(d, double)
EXPORT_TYPE(c, std::complex<float>)
EXPORT_TYPE(z, std::complex<double>)

extern "C" const char* oracle_blas_config() {
  return scipy_openblas_get_config();
}
extern "C" int oracle_hardware_threads() {
  return static_cast<int>(std::thread::hardware_concurrency());
}
// Largest pool size the BLAS underneath tolerates: the scipy wheel's OpenBLAS is built with
// MAX_THREADS=64 and aborts ("too many memory regions") when more threads call it concurrently.
extern "C" int oracle_max_pool_threads() {
  const int hw = static_cast<int>(std::thread::hardware_concurrency());
  return hw < 1 ? 1 : (hw > 56 ? 56 : hw);
}
