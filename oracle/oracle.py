"""ORACLE — TEST INFRASTRUCTURE ONLY (see cholesky_oracle.cpp for the rules).

Python face of the CPU restatement of the reference POTRF path plus the reference's closed-form
golden vectors and comparator. Every function cites the reference file:line it follows
(paths relative to the DLA-Future v0.10.0 tree).

Parity is PINNED: `cholesky_local` is checked against `cholesky_setters` (the reference's own
closed-form test vectors) for every size/uplo/type of test_cholesky.cpp in tests/test_oracle.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DTYPES = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def build() -> str:
    """Compile liboracle.so (g++ + the OpenBLAS inside the scipy wheel). Idempotent."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "cholesky_oracle.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        for t in "sdcz":
            f = getattr(_LIB, f"oracle_cholesky_local_{t}")
            f.argtypes = [ctypes.c_char, ctypes.c_long, ctypes.c_long, ctypes.c_void_p, ctypes.c_long,
                          ctypes.c_int]
            f.restype = ctypes.c_int
            f = getattr(_LIB, f"oracle_lapack_potrf_{t}")
            f.argtypes = [ctypes.c_char, ctypes.c_long, ctypes.c_void_p, ctypes.c_long, ctypes.c_int]
            f.restype = ctypes.c_int
            f = getattr(_LIB, f"oracle_set_random_hpd_{t}")
            f.argtypes = [ctypes.c_long, ctypes.c_long, ctypes.c_void_p, ctypes.c_long]
            f.restype = None
            f = getattr(_LIB, f"oracle_tile_trsm_{t}")
            f.argtypes = [ctypes.c_char] * 4 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                ctypes.c_void_p, ctypes.c_int]
            f.restype = None
            f = getattr(_LIB, f"oracle_tile_gemm_{t}")
            f.argtypes = [ctypes.c_char] * 2 + [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                                       ctypes.c_void_p, ctypes.c_int]
            f.restype = None
            f = getattr(_LIB, f"oracle_tile_herk_{t}")
            f.argtypes = [ctypes.c_char] * 2 + [ctypes.c_int] * 2 + [ctypes.c_double, ctypes.c_void_p, ctypes.c_int,
                                                                       ctypes.c_double, ctypes.c_void_p, ctypes.c_int]
            f.restype = None
            f = getattr(_LIB, f"oracle_residual_{t}")
            f.argtypes = [ctypes.c_char, ctypes.c_long, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
                          ctypes.c_long]
            f.restype = ctypes.c_double
        _LIB.oracle_blas_config.restype = ctypes.c_char_p
        _LIB.oracle_hardware_threads.restype = ctypes.c_int
        _LIB.oracle_max_pool_threads.restype = ctypes.c_int
    return _LIB


def max_pool_threads() -> int:
    """Host threads the tiled CPU run may use (capped by what the wheel's OpenBLAS tolerates)."""
    return lib().oracle_max_pool_threads()


def type_char(dtype) -> str:
    dtype = np.dtype(dtype)
    for k, v in DTYPES.items():
        if np.dtype(v) == dtype:
            return k
    raise TypeError(f"unsupported dtype {dtype}")


def _check_fortran(a: np.ndarray):
    """Column-major storage: unit stride down the columns, any leading dimension >= rows (sub-views of a Fortran
    array included, like a tile inside a slab)."""
    assert a.ndim == 2, "2-D array expected"
    ok = a.flags.f_contiguous or a.size == 0 or (a.strides[0] == a.itemsize and (a.shape[1] <= 1 or a.strides[1] >= a.shape[0] * a.itemsize))
    assert ok, "column-major (Fortran-order) array expected"


def cholesky_local(uplo: str, a: np.ndarray, nb: int, nthreads: int = 1) -> int:
    """In-place tiled Cholesky of the Fortran-ordered square `a` — the reference's local algorithm,
    include/dlaf/factorization/cholesky/impl.h:150-189 (L) / :316-348 (U). Returns LAPACK info."""
    _check_fortran(a)
    n = a.shape[0]
    lda = max(1, a.strides[1] // a.itemsize) if n > 0 else 1
    f = getattr(lib(), f"oracle_cholesky_local_{type_char(a.dtype)}")
    return f(uplo.encode(), n, nb, a.ctypes.data, lda, nthreads)


def _scalar(dtype, v):
    return np.array([v], dtype=dtype)


def tile_trsm(side: str, uplo: str, op: str, diag: str, m: int, n: int, alpha, a: np.ndarray, lda: int, b: np.ndarray,
              ldb: int) -> None:
    """The oracle's tile TRSM wrapper (what trsmPanelTile calls, impl.h:55-67 / :107-119), B (m x n) in place."""
    t = type_char(b.dtype)
    al = _scalar(b.dtype, alpha)
    getattr(lib(), f"oracle_tile_trsm_{t}")(side.encode(), uplo.encode(), op.encode(), diag.encode(), m, n,
                                             al.ctypes.data, a.ctypes.data, lda, b.ctypes.data, ldb)


def tile_gemm(opa: str, opb: str, m: int, n: int, k: int, alpha, a: np.ndarray, lda: int, b: np.ndarray, ldb: int, beta,
              c: np.ndarray, ldc: int) -> None:
    """The oracle's tile GEMM wrapper (gemmTrailingMatrixTile, impl.h:82-94 / :134-146), C in place."""
    t = type_char(c.dtype)
    al, be = _scalar(c.dtype, alpha), _scalar(c.dtype, beta)
    getattr(lib(), f"oracle_tile_gemm_{t}")(opa.encode(), opb.encode(), m, n, k, al.ctypes.data, a.ctypes.data, lda,
                                             b.ctypes.data, ldb, be.ctypes.data, c.ctypes.data, ldc)


def tile_herk(uplo: str, op: str, n: int, k: int, alpha: float, a: np.ndarray, lda: int, beta: float, c: np.ndarray,
              ldc: int) -> None:
    """The oracle's tile HERK/SYRK wrapper (herkTrailingDiagTile, impl.h:69-80 / :121-132), C in place."""
    t = type_char(c.dtype)
    getattr(lib(), f"oracle_tile_herk_{t}")(uplo.encode(), op.encode(), n, k, float(alpha), a.ctypes.data, lda,
                                             float(beta), c.ctypes.data, ldc)


def lapack_potrf(uplo: str, a: np.ndarray, nthreads: int) -> int:
    """Monolithic multithreaded LAPACK ?potrf (second CPU data point of BASELINE.md §3)."""
    _check_fortran(a)
    n = a.shape[0]
    lda = max(1, a.strides[1] // a.itemsize) if n > 0 else 1
    f = getattr(lib(), f"oracle_lapack_potrf_{type_char(a.dtype)}")
    return f(uplo.encode(), n, a.ctypes.data, lda, nthreads)


def set_random_hermitian_positive_definite(n: int, nb: int, dtype) -> np.ndarray:
    """include/dlaf/util_matrix.h:410-453 + :529-531 (per-tile mt19937_64, diagonal + 2N)."""
    a = np.zeros((n, n), dtype=dtype, order="F")
    if n:
        getattr(lib(), f"oracle_set_random_hpd_{type_char(dtype)}")(n, nb, a.ctypes.data, n)
    return a


def residual(uplo: str, a: np.ndarray, fac: np.ndarray) -> float:
    """max|A - L L^H| / max|A| over the `uplo` triangle — miniapp/miniapp_cholesky.cpp:408-446."""
    _check_fortran(a)
    _check_fortran(fac)
    n = a.shape[0]
    if n == 0:
        return 0.0
    f = getattr(lib(), f"oracle_residual_{type_char(a.dtype)}")
    return f(uplo.encode(), n, a.ctypes.data, a.strides[1] // a.itemsize, fac.ctypes.data,
             fac.strides[1] // fac.itemsize)


def residual_gate(dtype, n: int):
    """(clean, error) thresholds of the miniapp: eps*n and 100*eps*n (miniapp_cholesky.cpp:435-445)."""
    eps = np.finfo(np.dtype(dtype).type(0).real.dtype).eps
    return eps * n, 100 * eps * n


# ---------------------------------------------------------------------------------------------
# Golden vectors: closed-form A and its exact factor (getCholeskySetters,
# test/include/dlaf_test/matrix/util_generic_lapack.h:39-68). The unreferenced triangle holds the
# sentinel -9.9 in BOTH, so the comparison also proves the other triangle is never written.
# ---------------------------------------------------------------------------------------------
SENTINEL = -9.9


def cholesky_setters(uplo: str, m: int, dtype):
    dtype = np.dtype(dtype)
    i = np.arange(m, dtype=np.float64)[:, None]
    j = np.arange(m, dtype=np.float64)[None, :]
    if dtype.kind == "c":
        phase = np.exp(1j * (j - i))  # polar(r, -i + j)
    else:
        phase = 1.0
    a = np.exp2(-(i + j)) / 3 * (np.exp2(2 * (np.minimum(i, j) + 1)) - 1) * phase
    t = np.exp2(-np.abs(i - j)) * phase
    unref = (i < j) if uplo.upper() == "L" else (i > j)
    a = np.where(unref, SENTINEL, a)
    t = np.where(unref, SENTINEL, t)
    return np.asfortranarray(a.astype(dtype)), np.asfortranarray(t.astype(dtype))


def type_error(dtype) -> float:
    """TypeUtilities<T>::error: 2 eps (real) / 8 eps (complex), test/include/dlaf_test/util_types.h:40,:62."""
    dtype = np.dtype(dtype)
    eps = np.finfo(dtype.type(0).real.dtype).eps
    return (8 if dtype.kind == "c" else 2) * float(eps)


def cholesky_tolerance(m: int, dtype) -> float:
    """4 (m + 1) error, both relative and absolute (test_cholesky.cpp:76-77)."""
    return 4 * (m + 1) * type_error(dtype)


def check_near(expected: np.ndarray, value: np.ndarray, rel_err: float, abs_err: float):
    """CHECK_MATRIX_NEAR (test/include/dlaf_test/matrix/util_matrix.h:255-282): an element passes if
    |d| < abs_err OR |d| / max(|e|, |v|) < rel_err. Returns (ok, worst_index, message)."""
    assert expected.shape == value.shape
    if expected.size == 0:
        return True, None, ""
    e = expected.astype(np.complex128 if np.iscomplexobj(expected) else np.float64)
    v = value.astype(e.dtype)
    diff = np.abs(e - v)
    amax = np.maximum(np.abs(e), np.abs(v))
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(amax > 0, diff / amax, 0.0)
    ok = (diff < abs_err) | (rel < rel_err)
    ok &= ~np.isnan(diff)
    if ok.all():
        return True, None, ""
    idx = np.unravel_index(np.argmax(np.where(ok, 0.0, np.where(np.isnan(diff), np.inf, diff))), e.shape)
    return False, idx, (f"expected {e[idx]} == {v[idx]} at {idx} (rel {rel[idx]:.3e} > {rel_err:.3e}, "
                        f"abs {diff[idx]:.3e} > {abs_err:.3e}); {np.count_nonzero(~ok)} elements differ")


# Sizes of the reference's algorithm test (test/unit/factorization/test_cholesky.cpp:54-58): (m, mb)
CHOLESKY_TEST_SIZES = [(0, 2), (5, 8), (34, 34), (4, 3), (16, 10), (34, 13), (32, 5)]


# ---------------------------------------------------------------------------------------------
# Block-cyclic index math (include/dlaf/matrix/util_distribution.h:82-196), tiles_per_block = 1,
# tile_offset = 0 (the only case on the POTRF path, factorization/cholesky.h:45,:75).
# ---------------------------------------------------------------------------------------------
def rank_global_tile(g: int, grid: int, src: int) -> int:
    """util_distribution.h:82-92"""
    return (g + src) % grid


def local_tile_from_global_tile(g: int, grid: int, rank: int, src: int) -> int:
    """util_distribution.h:103-126 (-1 when the rank does not own the tile)"""
    return g // grid if rank_global_tile(g, grid, src) == rank else -1


def next_local_tile_from_global_tile(g: int, grid: int, rank: int, src: int) -> int:
    """util_distribution.h:138-166"""
    rank_virt = (rank - src) % grid
    owner_virt = g % grid
    loc = g // grid
    return loc + 1 if rank_virt < owner_virt else loc


def global_tile_from_local_tile(l: int, grid: int, rank: int, src: int) -> int:
    """util_distribution.h:177-196"""
    return grid * l + (rank - src) % grid


def local_nr_tiles(nt: int, grid: int, rank: int, src: int) -> int:
    """src/matrix/distribution.cpp:117-150"""
    return next_local_tile_from_global_tile(nt, grid, rank, src)


def local_size(n: int, nb: int, grid: int, rank: int, src: int) -> int:
    nt = -(-n // nb)
    lt = local_nr_tiles(nt, grid, rank, src)
    sz = lt * nb
    if nt > 0 and rank_global_tile(nt - 1, grid, src) == rank:
        sz -= nt * nb - n
    return sz


def scatter_block_cyclic(a: np.ndarray, nb: int, grid, src=(0, 0)):
    """Split the global matrix into the per-rank local column-major parts of a P x Q block-cyclic
    distribution (what each rank of the reference's distributed tests holds)."""
    P, Q = grid
    n = a.shape[0]
    nt = -(-n // nb)
    out = {}
    for p in range(P):
        for q in range(Q):
            rows = [g for g in range(nt) if rank_global_tile(g, P, src[0]) == p]
            cols = [g for g in range(nt) if rank_global_tile(g, Q, src[1]) == q]
            ridx = np.concatenate([np.arange(g * nb, min(n, (g + 1) * nb)) for g in rows]) if rows else np.zeros(0, int)
            cidx = np.concatenate([np.arange(g * nb, min(n, (g + 1) * nb)) for g in cols]) if cols else np.zeros(0, int)
            out[(p, q)] = np.asfortranarray(a[np.ix_(ridx, cidx)])
    return out


def gather_block_cyclic(parts, n: int, nb: int, grid, dtype, src=(0, 0)):
    P, Q = grid
    nt = -(-n // nb)
    a = np.zeros((n, n), dtype=dtype, order="F")
    for p in range(P):
        for q in range(Q):
            rows = [g for g in range(nt) if rank_global_tile(g, P, src[0]) == p]
            cols = [g for g in range(nt) if rank_global_tile(g, Q, src[1]) == q]
            ridx = np.concatenate([np.arange(g * nb, min(n, (g + 1) * nb)) for g in rows]) if rows else np.zeros(0, int)
            cidx = np.concatenate([np.arange(g * nb, min(n, (g + 1) * nb)) for g in cols]) if cols else np.zeros(0, int)
            a[np.ix_(ridx, cidx)] = parts[(p, q)]
    return a
