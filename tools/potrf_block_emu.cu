// Host emulation of the blocked diagonal kernel: runs the __host__ __device__ phase functions of
// potrf_block.cuh for all 256 threads sequentially, phase by phase (= barrier by barrier), and checks
// L and inv(L) against straightforward host loops. Build: nvcc -O2 -o /tmp/emu tools/potrf_block_emu.cu
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

#include "../dla-future_b200/csrc/potrf_block.cuh"

using namespace dlaf_b200;
using namespace dlaf_b200::pblock;

inline float cj(float v) { return v; }
inline double cj(double v) { return v; }
template <class R> inline std::complex<R> cj(std::complex<R> v) { return std::conj(v); }

template <class T> struct H;  // host complex twin
template <> struct H<double> { using type = double; };
template <> struct H<float> { using type = float; };
template <> struct H<double2> { using type = std::complex<double>; };
template <> struct H<float2> { using type = std::complex<float>; };

template <class T, int PBv>
int run(const char* name, bool make_fail) {
  using C = Cfg<T, PBv>;
  using HT = typename H<T>::type;
  using R = base_t<T>;
  constexpr int n = PBv, BS = C::BS;
  const long ld = n + 3;
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> dist(-1, 1);
  std::vector<HT> X(n * n), A(ld * n, HT(-9.9)), Lr(n * n, HT(0));
  for (auto& x : X) {
    if constexpr (sizeof(HT) == 2 * sizeof(R)) x = HT(dist(rng), dist(rng)); else x = HT(dist(rng));
  }
  for (int j = 0; j < n; ++j)
    for (int i = j; i < n; ++i) {
      HT s = (i == j) ? HT(R(n)) : HT(0);
      for (int k = 0; k < n; ++k) s += X[i + k * n] * cj(X[j + k * n]);
      if (i == j) s = HT(std::real(s));
      A[i + j * ld] = s;
    }
  if (make_fail) A[70 % n + (70 % n) * ld] = HT(R(-5));
  // reference Cholesky
  for (int j = 0; j < n; ++j) for (int i = j; i < n; ++i) Lr[i + j * n] = A[i + j * ld];
  int ref_fail = 0;
  for (int j = 0; j < n && !ref_fail; ++j) {
    R ajj = std::real(Lr[j + j * n]);
    if (!(ajj > 0)) { ref_fail = j + 1; break; }
    R d = std::sqrt(ajj); Lr[j + j * n] = d;
    for (int i = j + 1; i < n; ++i) Lr[i + j * n] /= d;
    for (int s = j + 1; s < n; ++s) for (int i = s; i < n; ++i) Lr[i + s * n] -= Lr[i + j * n] * cj(Lr[s + j * n]);
  }
  // emulation
  std::vector<T> Tm(ld * n), W(n * n);
  memcpy(Tm.data(), A.data(), sizeof(T) * ld * n);
  std::vector<T> panel(C::PANEL_ELEMS);
  std::vector<R> dd(n), dinv(n);
  static T reg[256][C::BS][C::BS];
  for (int tid = 0; tid < 256; ++tid) load_block<C, T>(reg[tid], Tm.data(), ld, tid % 16, tid / 16);
  int fail = 0;
  std::vector<T> dfL(BS * BS), msc(BS * BS);
  std::vector<R> dfinv(BS), dfsq(BS);
  {
    int f = factor_pivot_block<C, T>(reg[0], dfL.data(), dfinv.data(), dfsq.data());
    if (f) fail = f;
  }
  for (int J = 0; J < 16 && !fail; ++J) {
    for (int tid = 0; tid < 256; ++tid) if (tid / 16 == J) write_panel<C, T>(reg[tid], panel.data(), tid % 16);
    for (int t = 0; t < n; ++t) solve_panel_row<C, T>(panel.data(), dfL.data(), dfinv.data(), dfsq.data(), msc.data(), dd.data(), dinv.data(), J, t);
    for (int tid = 0; tid < 256; ++tid) update_block<C, T>(reg[tid], panel.data(), dinv.data(), J, tid % 16, tid / 16);
    if (J + 1 < 16) {
      int f = factor_pivot_block<C, T>(reg[(J + 1) * 16 + (J + 1)], dfL.data(), dfinv.data(), dfsq.data());
      if (f) fail = (J + 1) * BS + f;
    }
  }
  if (make_fail || ref_fail) {
    std::printf("%-8s fail test: kernel %d reference %d %s\n", name, fail, ref_fail, fail == ref_fail ? "OK" : "MISMATCH");
    return fail == ref_fail ? 0 : 1;
  }
  for (int tid = 0; tid < 256; ++tid) store_block<C, T>(reg[tid], Tm.data(), ld, W.data(), n, dd.data(), dinv.data(), tid % 16, tid / 16);
  const HT* Lk = reinterpret_cast<const HT*>(Tm.data());
  const HT* Wk = reinterpret_cast<const HT*>(W.data());
  double errL = 0, errI = 0, errU = 0;
  long bad = 0;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      if (i >= j) errL = std::max<double>(errL, std::abs(Lk[i + j * ld] - Lr[i + j * n]));
      else if (Lk[i + j * ld] != HT(-9.9)) bad++;
      if (i < j) errU = std::max<double>(errU, std::abs(Wk[i + j * n]));
      HT s = 0;
      for (int k = 0; k < n; ++k) s += Wk[i + k * n] * ((k >= j) ? Lr[k + j * n] : HT(0));
      errI = std::max<double>(errI, std::abs(s - HT(i == j ? 1 : 0)));
    }
  const double eps = std::numeric_limits<R>::epsilon();
  const bool ok = errL < 200 * eps && errI < 200 * eps && errU == 0 && bad == 0;
  std::printf("%-8s max|L-ref| %.2e  max|W L - I| %.2e  upper(W) %.1e  sentinel writes %ld  %s\n", name, errL, errI, errU, bad, ok ? "OK" : "FAIL");
  return ok ? 0 : 1;
}

int main() {
  int rc = 0;
  rc |= run<double, 128>("double", false);
  rc |= run<float, 128>("float", false);
  rc |= run<double2, 64>("zcomplex", false);
  rc |= run<float2, 64>("ccomplex", false);
  rc |= run<double, 128>("double", true);
  rc |= run<double2, 64>("zcomplex", true);
  return rc;
}
