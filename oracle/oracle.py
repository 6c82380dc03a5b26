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


# ---------------------------------------------------------------------------------------------
# Triangular solver: the reference's local tile loops (oracle_triangular_solver_*) and its closed-form test systems
# (test/include/dlaf_test/matrix/util_generic_blas.h:259-371), sizes / alpha / tolerance of
# test/unit/solver/test_triangular.cpp:54-66, :100-101, :149.
# ---------------------------------------------------------------------------------------------
TRIANGULAR_TEST_SIZES = [(0, 0, 1, 1), (0, 2, 1, 2), (7, 0, 2, 1), (2, 2, 5, 5), (10, 10, 2, 3), (7, 7, 3, 2), (3, 2, 7, 7),
                         (12, 3, 5, 5), (7, 6, 3, 2), (15, 7, 3, 5), (2, 3, 7, 7), (4, 13, 5, 5), (7, 8, 2, 9), (19, 25, 6, 5)]
TRIANGULAR_TEST_ALPHA = complex(-1.2, 0.7)  # TypeUtilities<T>::element(-1.2, .7): the imaginary part is dropped for real T


def triangular_solver(side: str, uplo: str, op: str, diag: str, alpha, a: np.ndarray, b: np.ndarray, mb: int, nb: int) -> None:
    """In place on the Fortran-ordered b: op(A) X = alpha B (side 'L') or X op(A) = alpha B (side 'R')."""
    _check_fortran(b)
    m, n = b.shape
    if m == 0 or n == 0:
        return
    _check_fortran(a)
    f = getattr(lib(), f"oracle_triangular_solver_{type_char(b.dtype)}")
    cc, cl, vp = ctypes.c_char, ctypes.c_long, ctypes.c_void_p
    f.argtypes = [cc, cc, cc, cc, vp, cl, cl, cl, cl, vp, cl, vp, cl]
    f.restype = None
    if np.dtype(b.dtype).kind != "c" and op.upper() == "C":
        op = "T"  # ConjTrans == Trans for real types (include/dlaf/gpu/blas/gpublas.h:107-112)
    al = np.array([alpha], dtype=b.dtype)
    f(side.upper().encode(), uplo.upper().encode(), op.upper().encode(), diag.upper().encode(), al.ctypes.data, m, n, mb, nb,
      a.ctypes.data, max(1, a.strides[1] // a.itemsize), b.ctypes.data, max(1, b.strides[1] // b.itemsize))


def _polar(r, theta, dtype):
    return (r * np.exp(1j * theta)).astype(dtype) if np.dtype(dtype).kind == "c" else np.asarray(r).astype(dtype)


def triangular_system(side: str, uplo: str, op: str, diag: str, alpha, m: int, n: int, dtype):
    """(A, B, X): A as STORED (so that op(A) has the closed form below; unreferenced entries and a unit diagonal hold the
    sentinel -9.9), B, and the exact solution X — getLeftTriangularSystem / getRightTriangularSystem of the reference.
    `set(mat_a, el_op_a, op)` of the test stores A(i,j) = el_op_a(i,j) (NoTrans), el_op_a(j,i) (Trans) or its conjugate."""
    dtype = np.dtype(dtype)
    left = side.upper() == "L"
    na = m if left else n
    op_a_lower = (uplo.upper() == "L") == (op.upper() == "N")
    i = np.arange(na, dtype=np.float64)[:, None]
    k = np.arange(na, dtype=np.float64)[None, :]
    if left:
        opa = _polar((i + 1) / (k + 0.5), 2 * i - k, dtype)       # op(A)_ik
    else:
        opa = _polar((k + 1) / (i + 0.5), 2 * k - i, dtype)       # op(A)_kj with (row, col) = (i, k) here
    unref = (i < k) if op_a_lower else (i > k)
    if diag.upper() == "U":
        unref = unref | (i == k)
    opa = np.where(unref, np.asarray(SENTINEL).astype(dtype), opa)
    if op.upper() == "N":
        a = opa
    elif op.upper() == "T":
        a = opa.T
    else:
        a = opa.conj().T
    r = np.arange(m, dtype=np.float64)[:, None]
    c = np.arange(n, dtype=np.float64)[None, :]
    al = np.asarray(alpha).astype(dtype)
    if left:
        x = _polar((r + 0.5) / (c + 2), r + c, dtype)
        kk = (r + 1) if op_a_lower else (m - r)
        gamma = _polar((r + 1) / (c + 2), 2 * r + c, dtype)
    else:
        x = _polar((c + 0.5) / (r + 2), r + c, dtype)
        kk = (n - c) if op_a_lower else (c + 1)
        gamma = _polar((c + 1) / (r + 2), r + 2 * c, dtype)
    if diag.upper() == "U":
        bmat = ((kk - 1) * gamma + x) / al
    else:
        bmat = kk * gamma / al
    return np.asfortranarray(a.astype(dtype)), np.asfortranarray(np.broadcast_to(bmat, (m, n)).astype(dtype)), \
        np.asfortranarray(np.broadcast_to(x, (m, n)).astype(dtype))


def triangular_tolerance(m: int, dtype, distributed: bool = False) -> float:
    """test_triangular.cpp:100-101 (local: 40 (m+1) error), :138-139 (distributed: 20 (m+1) error)."""
    dtype = np.dtype(dtype)
    eps = np.finfo(dtype.type(0).real.dtype).eps
    err = (8 if dtype.kind == "c" else 2) * eps  # TypeUtilities<T>::error (util_types.h:40, :62)
    return (20 if distributed else 40) * (m + 1) * err


# ---------------------------------------------------------------------------------------------
# Inverse: the reference's local tile loops (oracle_triangular_inverse_*, oracle_assemble_cholesky_inverse_*,
# oracle_inverse_from_cholesky_factor_*) and the closed forms of its tests
# (test/include/dlaf_test/matrix/util_generic_lapack.h:165-196, :212-247, :261-327), sizes and tolerance of
# test/unit/inverse/test_triangular_inverse.cpp:54-58, :75-76 and test_inverse_from_cholesky_factor.cpp:53-57, :74-75.
# ---------------------------------------------------------------------------------------------
INVERSE_TEST_SIZES = [(0, 2), (5, 8), (34, 34), (4, 3), (16, 10), (34, 13), (32, 5)]  # (m, mb)


def _inverse_call(name: str, a: np.ndarray, nb: int, *chars) -> None:
    _check_fortran(a)
    n = a.shape[0]
    assert a.shape == (n, n)
    if n == 0:
        return
    f = getattr(lib(), f"{name}_{type_char(a.dtype)}")
    cc, cl, vp = ctypes.c_char, ctypes.c_long, ctypes.c_void_p
    f.argtypes = [cc] * len(chars) + [cl, cl, vp, cl]
    f.restype = None
    f(*[c.upper().encode() for c in chars], n, nb, a.ctypes.data, max(1, a.strides[1] // a.itemsize))


def triangular_inverse(uplo: str, diag: str, a: np.ndarray, nb: int) -> None:
    """In place: the `uplo` triangle of the Fortran-ordered a <- inverse of that triangular matrix."""
    _inverse_call("oracle_triangular_inverse", a, nb, uplo, diag)


def assemble_cholesky_inverse(uplo: str, a: np.ndarray, nb: int) -> None:
    """In place: triangular T in the `uplo` triangle <- `uplo` triangle of T^H T ('L') / T T^H ('U')."""
    _inverse_call("oracle_assemble_cholesky_inverse", a, nb, uplo)


def inverse_from_cholesky_factor(uplo: str, a: np.ndarray, nb: int) -> None:
    """In place: Cholesky factor in the `uplo` triangle <- `uplo` triangle of inv(A)."""
    _inverse_call("oracle_inverse_from_cholesky_factor", a, nb, uplo)


def _ij(n: int):
    return np.arange(n, dtype=np.float64)[:, None], np.arange(n, dtype=np.float64)[None, :]


def _unref(uplo: str, i, j, with_diag: bool = False):
    u = (i < j) if uplo.upper() == "L" else (i > j)
    return (u | (i == j)) if with_diag else u


def _finish(x, unref, dtype):
    return np.asfortranarray(np.where(unref, np.asarray(SENTINEL).astype(dtype), x.astype(dtype)))


def triangular_inverse_setters(uplo: str, diag: str, n: int, dtype):
    """(A, inv(A)) of get_triangular_inverse_setters (util_generic_lapack.h:261-327); unreferenced entries (and the
    diagonal for Diag::Unit) hold the sentinel in both."""
    dtype = np.dtype(dtype)
    i, j = _ij(n)
    unit = diag.upper() == "U"
    unref = _unref(uplo, i, j, unit)
    scale, dval, oval = (1.0, 1.0, 0.5) if unit else (4.0, 0.25, 0.125)
    a = _polar(scale * np.exp2(-np.abs(i - j)), -i + j, dtype)
    off = -_polar(np.full((n, n), oval), -i + j, dtype)
    res = np.where(i == j, np.asarray(dval).astype(dtype), np.where(np.abs(i - j) == 1, off, np.asarray(0).astype(dtype)))
    return _finish(a, unref, dtype), _finish(res, unref, dtype)


def assemble_cholesky_inverse_setters(uplo: str, n: int, dtype):
    """(T, A) of get_assemble_cholesky_inverse_setters (util_generic_lapack.h:165-196): A = T^H T (Lower) / T T^H (Upper)."""
    dtype = np.dtype(dtype)
    i, j = _ij(n)
    unref = _unref(uplo, i, j)
    t = _polar(np.exp2(-np.abs(i - j)), -i + j, dtype)
    ri, rj = n - 1 - i, n - 1 - j
    a = _polar(np.exp2(-(ri + rj)) / 3 * (np.exp2(2 * (np.minimum(ri, rj) + 1)) - 1), ri - rj, dtype)
    return _finish(t, unref, dtype), _finish(a, unref, dtype)


def inverse_cholesky_factor_setters(uplo: str, n: int, dtype):
    """(T, A) of get_inverse_cholesky_factor_setters (util_generic_lapack.h:212-247): T the Cholesky factor (bidiagonal),
    A = inv(T^H) inv(T) (Lower) / inv(T) inv(T^H) (Upper)."""
    dtype = np.dtype(dtype)
    i, j = _ij(n)
    unref = _unref(uplo, i, j)
    off = -_polar(np.full((n, n), 0.5), -i + j, dtype)
    t = np.where(i == j, np.asarray(1).astype(dtype), np.where(np.abs(i - j) == 1, off, np.asarray(0).astype(dtype)))
    ri, rj = n - 1 - i, n - 1 - j
    a = _polar(np.exp2(-(ri + rj)) / 3 * (np.exp2(2 * (np.minimum(ri, rj) + 1)) - 1), ri - rj, dtype)
    return _finish(t, unref, dtype), _finish(a, unref, dtype)


def inverse_tolerance(m: int, dtype) -> float:
    """4 (m + 1) error, relative and absolute (test_triangular_inverse.cpp:75-76, test_inverse_from_cholesky_factor.cpp:74-75)."""
    return 4 * (m + 1) * type_error(dtype)


# ---------------------------------------------------------------------------------------------
# Generalized -> standard eigenproblem (HEGST, itype 1): the reference's local tile loops
# (oracle_generalized_to_standard_*) and its closed form (getGenToStdElementSetters,
# test/include/dlaf_test/matrix/util_generic_lapack.h:94-149) with the parameters, sizes and tolerance of
# test/unit/eigensolver/test_gen_to_std.cpp:52-56, :64-66, :78.
# ---------------------------------------------------------------------------------------------
GEN_TO_STD_TEST_SIZES = [(0, 2), (5, 8), (34, 34), (4, 3), (16, 10), (34, 13), (32, 5)]  # (m, mb)
GEN_TO_STD_PARAMS = (-2.0, 1.5, 0.95)  # alpha, beta, gamma


def generalized_to_standard(uplo: str, a: np.ndarray, l: np.ndarray, nb: int) -> None:
    """In place on the Fortran-ordered a: `uplo` triangle of inv(L) A inv(L)^H ('L') / inv(U)^H A inv(U) ('U'); l holds the
    Cholesky factor of B in its `uplo` triangle (read only)."""
    _check_fortran(a)
    _check_fortran(l)
    n = a.shape[0]
    assert a.shape == (n, n) and l.shape == (n, n) and a.dtype == l.dtype
    if n == 0:
        return
    f = getattr(lib(), f"oracle_generalized_to_standard_{type_char(a.dtype)}")
    cc, cl, vp = ctypes.c_char, ctypes.c_long, ctypes.c_void_p
    f.argtypes = [cc, cl, cl, vp, cl, vp, cl]
    f.restype = None
    f(uplo.upper().encode(), n, nb, a.ctypes.data, max(1, a.strides[1] // a.itemsize), l.ctypes.data,
      max(1, l.strides[1] // l.itemsize))


def gen_to_std_setters(uplo: str, n: int, dtype, params=GEN_TO_STD_PARAMS):
    """(T, A, B) of getGenToStdElementSetters(n, itype = 1, uplo, alpha, beta, gamma): B = inv(T) A inv(T)^H (Lower) /
    inv(T)^H A inv(T) (Upper); unreferenced entries hold the sentinel."""
    dtype = np.dtype(dtype)
    alpha, beta, gamma = params
    i, j = _ij(n)
    unref = _unref(uplo, i, j)
    t = _polar(beta / np.exp2(np.abs(i - j)), alpha * (i - j), dtype)
    a = _polar((i + 1) * (j + 1) * (beta * beta * gamma) / np.exp2(i + j), alpha * (i - j), dtype)
    b = _polar(gamma / np.exp2(i + j), alpha * (i - j), dtype)
    return _finish(t, unref, dtype), _finish(a, unref, dtype), _finish(b, unref, dtype)


def gen_to_std_tolerance(m: int, dtype) -> float:
    """absolute 10 (m + 1) error, no relative criterion (test_gen_to_std.cpp:78)."""
    return 10 * (m + 1) * type_error(dtype)


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


def scatter_block_cyclic_rect(a: np.ndarray, mb: int, nb: int, grid, src=(0, 0)):
    """Rectangular flavour (blocks mb x nb) for the right-hand sides of the triangular solver."""
    P, Q = grid
    m, n = a.shape
    mt, nt = -(-m // mb), -(-n // nb)
    out = {}
    for p in range(P):
        for q in range(Q):
            rows = [g for g in range(mt) if rank_global_tile(g, P, src[0]) == p]
            cols = [g for g in range(nt) if rank_global_tile(g, Q, src[1]) == q]
            ridx = np.concatenate([np.arange(g * mb, min(m, (g + 1) * mb)) for g in rows]) if rows else np.zeros(0, int)
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
