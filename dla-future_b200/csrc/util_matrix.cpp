// Host-side input generator of the miniapp: random Hermitian positive definite matrix whose values
// do not depend on the distribution (reference: include/dlaf/util_matrix.h:410-453 and :529-531,
// getter_random :161-189, tile setters :337-389). Product code (miniapp / bench input), written for
// the LOCAL part of a block-cyclic matrix: every local tile is an independent job on a host thread.
#include "util_matrix.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <random>
#include <thread>
#include <vector>

namespace dlaf_b200 {

namespace {

// Stream of values in [-1, 1] (real) or in the unit disc (complex) seeded per tile.
template <class T>
class TileRandom {
public:
  explicit TileRandom(long seed) : engine_(static_cast<std::size_t>(seed)) {}
  T next() { return sampler_(engine_); }

private:
  std::mt19937_64 engine_;
  std::uniform_real_distribution<T> sampler_{-1, 1};
};

template <class R>
class TileRandom<std::complex<R>> {
public:
  explicit TileRandom(long seed) : real_(seed) {}
  std::complex<R> next() {
    // The reference writes polar(abs(r()), pi * r()) (util_matrix.h:185-187); with gcc the second
    // argument is evaluated first, so the FIRST draw is the angle and the second the radius
    // (checked against a g++ build of that expression, see tests/test_generator.py).
    const R angle = static_cast<R>(M_PI) * real_.next();
    const R radius = std::abs(real_.next());
    return std::polar<R>(radius, angle);
  }

private:
  TileRandom<R> real_;
};

template <class T>
inline T conj_of(T v) {
  return v;
}
template <class R>
inline std::complex<R> conj_of(std::complex<R> v) {
  return std::conj(v);
}
template <class T>
struct real_of {
  using type = T;
};
template <class R>
struct real_of<std::complex<R>> {
  using type = R;
};

}  // namespace

template <class T>
void set_random_hermitian_positive_definite_local(const LocalMatrixView<T>& m) {
  using R = typename real_of<T>::type;
  const long n = m.n;
  const int nb = m.nb;
  const long nt = (n + nb - 1) / nb;
  const R offset = static_cast<R>(2 * n);  // util_matrix.h:529-531
  std::vector<std::pair<long, long>> jobs;  // (global tile row, global tile col)
  for (long gj = m.vcol; gj < nt; gj += m.Q)
    for (long gi = m.vrow; gi < nt; gi += m.P)
      jobs.emplace_back(gi, gj);

  auto fill = [&](long gi, long gj) {
    const long r0 = gi * nb, c0 = gj * nb;
    const long rows = std::min<long>(nb, n - r0), cols = std::min<long>(nb, n - c0);
    T* t = m.data + (gi / m.P) * nb + (gj / m.Q) * nb * m.ld;
    // one seed per unordered tile pair: the transposed tile re-draws the same stream (util_matrix.h:435-439)
    const long seed = (gi >= gj) ? c0 + r0 * n : r0 + c0 * n;
    TileRandom<T> rnd(seed);
    if (gi == gj) {
      for (long j = 0; j < cols; ++j) {
        for (long i = 0; i < j; ++i) {
          const T v = rnd.next();
          t[i + j * m.ld] = v;
          t[j + i * m.ld] = conj_of(v);
        }
        t[j + j * m.ld] = T(std::real(rnd.next()) + offset);
      }
      return;
    }
    // off-diagonal: the stream always covers a FULL nb x nb tile in column-major order of the lower
    // twin, also for ragged edge tiles (util_matrix.h:362-389)
    const bool lower = gi > gj;
    for (long j = 0; j < nb; ++j)
      for (long i = 0; i < nb; ++i) {
        const T v = rnd.next();
        if (lower) {
          if (i < rows && j < cols)
            t[i + j * m.ld] = v;
        }
        else if (j < rows && i < cols) {
          t[j + i * m.ld] = conj_of(v);
        }
      }
  };

  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const unsigned nthreads = static_cast<unsigned>(std::min<size_t>(std::min(hw, 64u), std::max<size_t>(1, jobs.size())));
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  for (unsigned w = 0; w < nthreads; ++w)
    pool.emplace_back([&] {
      for (size_t i; (i = next.fetch_add(1)) < jobs.size();)
        fill(jobs[i].first, jobs[i].second);
    });
  for (auto& th : pool)
    th.join();
}

template void set_random_hermitian_positive_definite_local<float>(const LocalMatrixView<float>&);
template void set_random_hermitian_positive_definite_local<double>(const LocalMatrixView<double>&);
template void set_random_hermitian_positive_definite_local<std::complex<float>>(const LocalMatrixView<std::complex<float>>&);
template void set_random_hermitian_positive_definite_local<std::complex<double>>(const LocalMatrixView<std::complex<double>>&);

}  // namespace dlaf_b200
