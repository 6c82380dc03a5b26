// See potrf_tile.cuh. Algorithm (validated against numpy in tools/proto_potrf_inv*.py):
//
// The G x G block lives in shared memory as S (column-major, odd leading dimension so both
// S(r, c) with r running and S(c, r) with r running are bank-conflict free). The lower triangle
// holds A -> L. The strictly upper triangle holds M^H where M converges to inv(L): applying the
// elementary eliminations L_j^-1 to the identity is, element for element, the SAME rank-1 rule as
// the right-looking Cholesky update:
//
//     for column j:   d = sqrt(S(j,j));  S(:,j) /= d  (all rows but j)
//                     for s > j, for r with (r >= s) or (r < j):   S(r,s) -= S(r,j) * conj(S(s,j))
//                     S(j,s) = -conj(S(s,j)) / d                             (new conj(M(s,j)))
//
// so one sweep with two barriers per column yields both L and inv(L).
#include "potrf_tile.cuh"

#include <cstdlib>
#include <string>

#include "common.h"
#include "potrf_block.cuh"
#include "potrf_cluster.cuh"
#include "types.h"

namespace dlaf_b200 {

namespace {

constexpr int kPotrfThreads = 256;

// ---- element arithmetic -----------------------------------------------------------------------
__device__ __forceinline__ float re_of(float v) { return v; }
__device__ __forceinline__ double re_of(double v) { return v; }
__device__ __forceinline__ float re_of(float2 v) { return v.x; }
__device__ __forceinline__ double re_of(double2 v) { return v.x; }

__device__ __forceinline__ float scale_r(float v, float s) { return v * s; }
__device__ __forceinline__ double scale_r(double v, double s) { return v * s; }
__device__ __forceinline__ float2 scale_r(float2 v, float s) { return make_float2(v.x * s, v.y * s); }
__device__ __forceinline__ double2 scale_r(double2 v, double s) { return make_double2(v.x * s, v.y * s); }

// v - a * conj(b)
__device__ __forceinline__ float sub_mul_conj(float v, float a, float b) { return fmaf(-a, b, v); }
__device__ __forceinline__ double sub_mul_conj(double v, double a, double b) { return fma(-a, b, v); }
__device__ __forceinline__ float2 sub_mul_conj(float2 v, float2 a, float2 b) {
  return make_float2(fmaf(-a.y, b.y, fmaf(-a.x, b.x, v.x)), fmaf(a.x, b.y, fmaf(-a.y, b.x, v.y)));
}
__device__ __forceinline__ double2 sub_mul_conj(double2 v, double2 a, double2 b) {
  return make_double2(fma(-a.y, b.y, fma(-a.x, b.x, v.x)), fma(a.x, b.y, fma(-a.y, b.x, v.y)));
}

__device__ __forceinline__ float neg(float v) { return -v; }
__device__ __forceinline__ double neg(double v) { return -v; }
__device__ __forceinline__ float2 neg(float2 v) { return make_float2(-v.x, -v.y); }
__device__ __forceinline__ double2 neg(double2 v) { return make_double2(-v.x, -v.y); }

// reciprocal square root: hardware estimate + one Newton step (~1 ulp)
__device__ __forceinline__ float fast_rsqrt(float a) {
  float y = rsqrtf(a);
  return fmaf(y * 0.5f, fmaf(-a * y, y, 1.0f), y);
}
__device__ __forceinline__ double fast_rsqrt(double a) {
  double y = rsqrt(a);
  return fma(y * 0.5, fma(-a * y, y, 1.0), y);
}
__device__ __forceinline__ float full_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double full_sqrt(double a) { return sqrt(a); }

template <class T>
__device__ __forceinline__ T zero_of() {
  return make_real<T>(0);
}

// ---- the kernel ---------------------------------------------------------------------------------
template <class T, int PB>
__global__ void __launch_bounds__(kPotrfThreads, 1)
    potrf_inv_kernel(T* __restrict__ Tm, long ldt, T* __restrict__ W, long ldw, int* info, int info_offset) {
  using R = base_t<T>;
  constexpr int PLD = PB + 1;
  constexpr int NG = kPotrfThreads / PB;  // column groups working in parallel
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* S = reinterpret_cast<T*>(smem_raw);
  R* dd = reinterpret_cast<R*>(S + PB * PLD);  // diag(L)
  R* dinv = dd + PB;                           // diag(inv(L))
  const int tid = threadIdx.x;

  for (int idx = tid; idx < PB * PB; idx += kPotrfThreads) {
    const int r = idx % PB, c = idx / PB;
    S[c * PLD + r] = (r >= c) ? Tm[r + c * ldt] : zero_of<T>();
  }
  __syncthreads();

  const int r = tid % PB, grp = tid / PB;
  int fail = 0;
  for (int j = 0; j < PB; ++j) {
    const R ajj = re_of(S[j * PLD + j]);  // nobody writes S(j,j) during or after step j
    if (!(ajj > R(0))) {                  // also catches NaN; uniform across the CTA
      fail = j + 1;
      break;
    }
    // Only the reciprocal is on the critical path; the diagonal entry itself is the correctly rounded
    // sqrt, computed by its owner thread only.
    const R inv_d = fast_rsqrt(ajj);
    if (grp == 0) {
      if (r == j) {
        dd[j] = full_sqrt(ajj);
        dinv[j] = inv_d;
      }
      else {
        S[j * PLD + r] = scale_r(S[j * PLD + r], inv_d);
      }
    }
    __syncthreads();
    const T* colj = S + j * PLD;
    if (r == j) {
      for (int s = j + 1 + grp; s < PB; s += NG)
        S[s * PLD + j] = scale_r(neg(conj_val(colj[s])), inv_d);
    }
    else {
      // rows above the pivot (inverse part) see every remaining column; rows below (Cholesky part)
      // only columns up to their own index.
      const int s_end = (r < j) ? PB : r + 1;
      const T srj = colj[r];
      T* col = S + r;
      int s = j + 1 + grp;
      for (; s + 3 * NG < s_end; s += 4 * NG) {
        const T b0 = colj[s], b1 = colj[s + NG], b2 = colj[s + 2 * NG], b3 = colj[s + 3 * NG];
        T v0 = col[s * PLD], v1 = col[(s + NG) * PLD], v2 = col[(s + 2 * NG) * PLD],
          v3 = col[(s + 3 * NG) * PLD];
        v0 = sub_mul_conj(v0, srj, b0);
        v1 = sub_mul_conj(v1, srj, b1);
        v2 = sub_mul_conj(v2, srj, b2);
        v3 = sub_mul_conj(v3, srj, b3);
        col[s * PLD] = v0;
        col[(s + NG) * PLD] = v1;
        col[(s + 2 * NG) * PLD] = v2;
        col[(s + 3 * NG) * PLD] = v3;
      }
      for (; s < s_end; s += NG)
        col[s * PLD] = sub_mul_conj(col[s * PLD], srj, colj[s]);
    }
    __syncthreads();
  }
  if (fail) {
    if (tid == 0)
      atomicCAS(info, 0, info_offset + fail);
    // Leave T as it is past the failure (a trapped reference run would not have produced a result
    // either); W gets zeros so that downstream GEMMs stay finite.
    for (int idx = tid; idx < PB * PB; idx += kPotrfThreads)
      W[(idx % PB) + (idx / PB) * ldw] = zero_of<T>();
    return;
  }

  for (int idx = tid; idx < PB * PB; idx += kPotrfThreads) {
    const int rr = idx % PB, c = idx / PB;
    if (rr > c) {
      Tm[rr + c * ldt] = S[c * PLD + rr];
      W[rr + c * ldw] = conj_val(S[rr * PLD + c]);  // conj(M(rr,c)) is stored at S(c,rr)
    }
    else if (rr == c) {
      Tm[rr + c * ldt] = make_real<T>(dd[rr]);
      W[rr + c * ldw] = make_real<T>(dinv[rr]);
    }
    else {
      W[rr + c * ldw] = zero_of<T>();
    }
  }
}

// Measurement aid: when non-null, thread 0 and the last thread record clock64() at every phase boundary.
__device__ long long* g_potrf_clock_trace = nullptr;
#define POTRF_TRACE(slot)                                                            \
  do {                                                                               \
    if (trace && (tid == 0 || tid == kPotrfThreads - 1))                              \
      trace[((J) * 8 + (slot)) * 2 + (tid != 0)] = clock64();                        \
  } while (0)

// ---- the blocked, register-resident kernel (potrf_block.cuh) -------------------------------------
template <class T, int PB>
__global__ void __launch_bounds__(kPotrfThreads, 1)
    potrf_inv_blocked_kernel(T* __restrict__ Tm, long ldt, T* __restrict__ W, long ldw, int* info,
                             int info_offset) {
  using C = pblock::Cfg<T, PB>;
  using R = base_t<T>;
  constexpr int BS = C::BS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* panel = reinterpret_cast<T*>(smem_raw);
  R* dd = reinterpret_cast<R*>(panel + C::PANEL_ELEMS);
  R* dinv = dd + PB;
  R* dfinv = dinv + PB;  // pivot-block factor: 1/diag, diag^2, strictly lower part
  R* dfsq = dfinv + BS;
  T* dfL = reinterpret_cast<T*>(dfsq + BS);
  T* msc = dfL + BS * BS;  // scratch: columns of inv(L_D) under construction
  int* sfail = reinterpret_cast<int*>(msc + BS * BS);
  const int tid = threadIdx.x, ti = tid % 16, tj = tid / 16;

  T reg[BS][BS];
  pblock::load_block<C, T>(reg, Tm, ldt, ti, tj);
  if (tid == 0)
    *sfail = 0;
  __syncthreads();
  if (ti == 0 && tj == 0) {
    const int f = pblock::factor_pivot_block<C, T>(reg, dfL, dfinv, dfsq);
    if (f)
      *sfail = f;
  }
  int fail = 0;
  long long* trace = g_potrf_clock_trace;
  for (int J = 0; J < 16; ++J) {
    POTRF_TRACE(0);
    if (tj == J)
      pblock::write_panel<C, T>(reg, panel, ti);
    POTRF_TRACE(1);
    __syncthreads();  // panel J and the factor of its pivot block are published
    POTRF_TRACE(2);
    fail = *sfail;
    if (fail)
      break;
    if (tid < PB)
      pblock::solve_panel_row<C, T>(panel, dfL, dfinv, dfsq, msc, dd, dinv, J, tid);
    POTRF_TRACE(3);
    __syncthreads();
    POTRF_TRACE(4);
    pblock::update_block<C, T>(reg, panel, dinv, J, ti, tj);
    POTRF_TRACE(5);
    if (J + 1 < 16 && ti == J + 1 && tj == J + 1) {
      const int f = pblock::factor_pivot_block<C, T>(reg, dfL, dfinv, dfsq);
      if (f)
        *sfail = (J + 1) * BS + f;
    }
    POTRF_TRACE(6);
    __syncthreads();
    POTRF_TRACE(7);
  }
  if (fail) {
    if (tid == 0)
      atomicCAS(info, 0, info_offset + fail);
    for (int idx = tid; idx < PB * PB; idx += kPotrfThreads)
      W[(idx % PB) + (idx / PB) * ldw] = zero_of<T>();
    return;
  }
  pblock::store_block<C, T>(reg, Tm, ldt, W, ldw, dd, dinv, ti, tj);
}

// ---- the two-CTA cluster variant (potrf_cluster.cuh) --------------------------------------------
template <class T>
struct DevClusterCtx {
  using R = base_t<T>;
  int cta, tid;
  T* panel[2];
  R *dd, *dinv, *dfinv, *dfsq;
  T *dfL, *msc;
  int* sfail;
  unsigned peer;
  __host__ __device__ void cta_sync() {
#ifdef __CUDA_ARCH__
    __syncthreads();
#endif
  }
  // release / acquire at cluster scope: the DSMEM stores issued before the barrier are visible to the peer after it
  __host__ __device__ void cluster_sync() {
#ifdef __CUDA_ARCH__
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
#endif
  }
  template <class V>
  __host__ __device__ void push(V* local, V v) {
#ifdef __CUDA_ARCH__
    static_assert(sizeof(V) == 4 || sizeof(V) == 8 || sizeof(V) == 16, "push: 4, 8 or 16 byte values");
    const unsigned la = static_cast<unsigned>(__cvta_generic_to_shared(local));
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(peer));
    if constexpr (sizeof(V) == 4) {
      asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(*reinterpret_cast<const unsigned*>(&v)) : "memory");
    }
    else if constexpr (sizeof(V) == 8) {
      asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(ra), "l"(*reinterpret_cast<const unsigned long long*>(&v)) : "memory");
    }
    else {
      const unsigned long long* q = reinterpret_cast<const unsigned long long*>(&v);
      asm volatile("st.shared::cluster.v2.u64 [%0], {%1, %2};" ::"r"(ra), "l"(q[0]), "l"(q[1]) : "memory");
    }
#else
    (void) local;
    (void) v;
#endif
  }
};

template <class T, int PB>
__global__ void __cluster_dims__(pblock::kClusterCtas, 1, 1) __launch_bounds__(pblock::kClusterThreads, 1)
    potrf_inv_cluster2_kernel(T* __restrict__ Tm, long ldt, T* __restrict__ W, long ldw, int* info, int info_offset) {
  using C = pblock::Cfg<T, PB>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DevClusterCtx<T> cx;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  cx.cta = static_cast<int>(rank);
  cx.peer = rank ^ 1u;
  cx.tid = threadIdx.x;
  pblock::ClusterSmem<C, T>::carve(cx, smem_raw);
  pblock::potrf_inv_cluster2_body<C, T>(cx, Tm, ldt, W, ldw, info, info_offset);
  // no CTA of the pair may exit while the other one can still store into its shared memory: every push is followed
  // by a cluster barrier inside the body, and the last statement of the body touches only global memory
}

template <class T>
void launch_cluster2(T* t, long ldt, T* w, long ldw, int* info, int info_offset, cudaStream_t stream) {
  constexpr int PB = Gran<T>::value;
  using C = pblock::Cfg<T, PB>;
  constexpr int smem = static_cast<int>(pblock::ClusterSmem<C, T>::bytes);
  static_assert(smem <= 48 * 1024, "cluster kernel shared memory");
  potrf_inv_cluster2_kernel<T, PB><<<pblock::kClusterCtas, pblock::kClusterThreads, smem, stream>>>(t, ldt, w, ldw, info,
                                                                                                   info_offset);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

template <class T>
void launch_blocked(T* t, long ldt, T* w, long ldw, int* info, int info_offset, cudaStream_t stream) {
  constexpr int PB = Gran<T>::value;
  using C = pblock::Cfg<T, PB>;
  constexpr int smem = C::PANEL_ELEMS * sizeof(T) + (2 * PB + 2 * C::BS) * sizeof(base_t<T>) + 2 * C::BS * C::BS * sizeof(T) + 16;
  potrf_inv_blocked_kernel<T, PB><<<1, kPotrfThreads, smem, stream>>>(t, ldt, w, ldw, info, info_offset);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

template <class T>
void launch_impl(T* t, long ldt, T* w, long ldw, int* info, int info_offset, cudaStream_t stream) {
  // DLAF_B200_POTRF_KERNEL=sweep selects the simple shared-memory sweep (kept for A/B measurements),
  // =cluster2 the two-SM cluster variant (potrf_cluster.cuh; validated by host emulation and one GPU run: 326 vs 342 us per 512-tile)
  static const std::string variant = [] {
    const char* e = std::getenv("DLAF_B200_POTRF_KERNEL");
    return std::string(e ? e : "");
  }();
  static const bool use_sweep = (variant == "sweep");
  if (variant == "cluster2") {
    launch_cluster2<T>(t, ldt, w, ldw, info, info_offset, stream);
    return;
  }
  if (!use_sweep) {
    launch_blocked<T>(t, ldt, w, ldw, info, info_offset, stream);
    return;
  }
  constexpr int PB = Gran<T>::value;
  constexpr int smem = (PB * (PB + 1)) * sizeof(T) + 2 * PB * sizeof(base_t<T>);
  static bool configured = false;
  if (!configured) {
    DLAF_CUDA_CHECK(cudaFuncSetAttribute(potrf_inv_kernel<T, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  potrf_inv_kernel<T, PB><<<1, kPotrfThreads, smem, stream>>>(t, ldt, w, ldw, info, info_offset);
  DLAF_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

void potrf_set_clock_trace(long long* dev_buffer) {
  DLAF_CUDA_CHECK(cudaMemcpyToSymbol(g_potrf_clock_trace, &dev_buffer, sizeof(dev_buffer)));
}

void launch_potrf128_inv_f64(double* T, long ldt, double* W, long ldw, int* info, int info_offset,
                             cudaStream_t stream) {
  launch_impl<double>(T, ldt, W, ldw, info, info_offset, stream);
}

template <>
void launch_potrf_inv<double>(double* t, long ldt, double* w, long ldw, int* info, int info_offset, cudaStream_t s) {
  launch_impl<double>(t, ldt, w, ldw, info, info_offset, s);
}
template <>
void launch_potrf_inv<float>(float* t, long ldt, float* w, long ldw, int* info, int info_offset, cudaStream_t s) {
  launch_impl<float>(t, ldt, w, ldw, info, info_offset, s);
}
template <>
void launch_potrf_inv<float2>(float2* t, long ldt, float2* w, long ldw, int* info, int info_offset, cudaStream_t s) {
  launch_impl<float2>(t, ldt, w, ldw, info, info_offset, s);
}
template <>
void launch_potrf_inv<double2>(double2* t, long ldt, double2* w, long ldw, int* info, int info_offset, cudaStream_t s) {
  launch_impl<double2>(t, ldt, w, ldw, info, info_offset, s);
}

}  // namespace dlaf_b200
