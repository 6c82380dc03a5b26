// Host emulation of the two-CTA cluster variant of the diagonal-block kernel (potrf_cluster.cuh): the SAME
// orchestration code runs on 2 x 128 host threads; shared memory = one buffer per "CTA", cta_sync / cluster_sync =
// std::barrier, the DSMEM push = a store into the peer's buffer at the same offset. Checks L and inv(L) against host
// loops and the failure path. Build + run (no GPU needed):
//   nvcc -std=c++20 -O2 -o /tmp/potrf_cluster_emu tools/potrf_cluster_emu.cu && /tmp/potrf_cluster_emu
#include <barrier>
#include <chrono>
#include <cstdlib>
#include <complex>
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <thread>
#include <vector>

#include "../dla-future_b200/csrc/potrf_cluster.cuh"

using namespace dlaf_b200;
using namespace dlaf_b200::pblock;

inline float cj(float v) { return v; }
inline double cj(double v) { return v; }
template <class R> inline std::complex<R> cj(std::complex<R> v) { return std::conj(v); }
template <class T> struct H;
template <> struct H<double> { using type = double; };
template <> struct H<float> { using type = float; };
template <> struct H<double2> { using type = std::complex<double>; };
template <> struct H<float2> { using type = std::complex<float>; };

struct Sync {
  std::barrier<> cta[2] = {std::barrier<>(kClusterThreads), std::barrier<>(kClusterThreads)};
  std::barrier<> cluster{kClusterCtas * kClusterThreads};
};

template <class T>
struct HostCtx {
  using R = base_t<T>;
  int cta, tid;
  T* panel[2];
  R *dd, *dinv, *dfinv, *dfsq;
  T *dfL, *msc;
  int* sfail;
  unsigned char *my_base, *peer_base;
  Sync* sync;
  // EMU_JITTER=1: random delays around the synchronisation points and the remote stores, to shake out orderings
  // that only a skewed schedule exposes (thread 0 of a CTA far ahead of / behind its peers)
  unsigned lcg = 12345;
  bool jitter = false;
  void maybe_stall() {
    if (!jitter)
      return;
    lcg = lcg * 1664525u + 1013904223u;
    if ((lcg >> 28) == 0)
      std::this_thread::sleep_for(std::chrono::microseconds(200 + (lcg >> 20) % 800));
    else if ((lcg >> 27) & 1)
      std::this_thread::yield();
  }
  void cta_sync() { maybe_stall(); sync->cta[cta].arrive_and_wait(); maybe_stall(); }
  void cluster_sync() { maybe_stall(); sync->cluster.arrive_and_wait(); maybe_stall(); }
  template <class V>
  void push(V* local, V v) {
    maybe_stall();
    *reinterpret_cast<V*>(peer_base + (reinterpret_cast<unsigned char*>(local) - my_base)) = v;
  }
};

template <class T, int PBv>
int run(const char* name, int fail_at) {
  using C = Cfg<T, PBv>;
  using HT = typename H<T>::type;
  using R = base_t<T>;
  constexpr int n = PBv;
  const long ld = n + 3;
  std::mt19937_64 rng(11);
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
  if (fail_at >= 0) A[fail_at + fail_at * ld] = HT(R(-5));
  for (int j = 0; j < n; ++j) for (int i = j; i < n; ++i) Lr[i + j * n] = A[i + j * ld];
  int ref_fail = 0;
  for (int j = 0; j < n && !ref_fail; ++j) {
    R ajj = std::real(Lr[j + j * n]);
    if (!(ajj > 0)) { ref_fail = j + 1; break; }
    R d = std::sqrt(ajj); Lr[j + j * n] = d;
    for (int i = j + 1; i < n; ++i) Lr[i + j * n] /= d;
    for (int s = j + 1; s < n; ++s) for (int i = s; i < n; ++i) Lr[i + s * n] -= Lr[i + j * n] * cj(Lr[s + j * n]);
  }
  std::vector<T> Tm(ld * n), W(n * n);
  std::memcpy(Tm.data(), A.data(), sizeof(T) * ld * n);
  std::memset(W.data(), 0x7f, sizeof(T) * n * n);
  int info = 0;
  const size_t bytes = ClusterSmem<C, T>::bytes;
  std::vector<unsigned char> smem0(bytes, 0xAB), smem1(bytes, 0xCD);
  unsigned char* base[2] = {smem0.data(), smem1.data()};
  auto sync = std::make_unique<Sync>();
  std::vector<std::thread> th;
  for (int cta = 0; cta < kClusterCtas; ++cta)
    for (int tid = 0; tid < kClusterThreads; ++tid)
      th.emplace_back([&, cta, tid] {
        HostCtx<T> cx;
        cx.cta = cta;
        cx.tid = tid;
        cx.my_base = base[cta];
        cx.peer_base = base[1 - cta];
        cx.sync = sync.get();
        cx.jitter = std::getenv("EMU_JITTER") != nullptr;
        cx.lcg = 977u * (cta * kClusterThreads + tid + 1);
        ClusterSmem<C, T>::carve(cx, base[cta]);
        potrf_inv_cluster2_body<C, T>(cx, Tm.data(), ld, W.data(), n, &info, 1000);
      });
  for (auto& t : th) t.join();
  if (fail_at >= 0 || ref_fail) {
    const int want = ref_fail ? 1000 + ref_fail : 0;
    std::printf("%-8s fail test (pivot %d): kernel info %d reference %d %s\n", name, fail_at, info, want, info == want ? "OK" : "MISMATCH");
    return info == want ? 0 : 1;
  }
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
  const bool ok = errL < 200 * eps && errI < 200 * eps && errU == 0 && bad == 0 && info == 0;
  std::printf("%-8s max|L-ref| %.2e  max|W L - I| %.2e  upper(W) %.1e  sentinel writes %ld  info %d  %s\n", name, errL, errI, errU, bad, info, ok ? "OK" : "FAIL");
  return ok ? 0 : 1;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int rc = 0;
  rc |= run<double, 128>("double", -1);
  rc |= run<float, 128>("float", -1);
  rc |= run<double2, 64>("zcomplex", -1);
  rc |= run<float2, 64>("ccomplex", -1);
  for (int f : {0, 7, 8, 70, 127}) rc |= run<double, 128>("double", f);
  rc |= run<double2, 64>("zcomplex", 37);
  std::printf(rc ? "FAILED\n" : "all OK\n");
  return rc;
}
