"""The oracle's tile operations against the reference's own tile known-answer tests (closed forms, sizes, tolerances):
  GEMM  test/unit/test_blas_tile/test_gemm.h:34-69 + test/include/dlaf_test/matrix/util_generic_blas.h:55-94, sizes test_blas_tile.cpp:50-56
  HERK  test/unit/test_blas_tile/test_herk.h:33-89, sizes test_blas_tile.cpp:136-141
  TRSM  test/unit/test_blas_tile/test_trsm.h:35-62 + util_generic_blas.h:259-371, sizes test_blas_tile.cpp:232-237
  POTRF test/unit/test_lapack_tile/test_potrf.h:33-77 (closed form shared with test_cholesky.cpp; non-SPD -> info 1)
They pin the argument conventions (side / uplo / op / which triangle is referenced / leading dimensions) of the four
wrappers the CPU restatement of call_L / call_U is built from — for every combination the POTRF path uses
 and the remaining ops of the reference's tables."""
import numpy as np
import pytest

TYPES = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
GEMM_SIZES = [(0, 0, 0, 0, 0, 0), (7, 0, 0, 3, 1, 0), (0, 5, 0, 0, 0, 1), (0, 0, 11, 1, 1, 2), (0, 5, 13, 1, 0, 1),
              (7, 0, 4, 1, 2, 0), (3, 11, 0, 0, 1, 0), (1, 1, 1, 0, 3, 0), (1, 12, 1, 1, 0, 7), (17, 12, 16, 1, 3, 0),
              (11, 23, 8, 0, 3, 4), (6, 9, 12, 1, 1, 1), (32, 32, 32, 0, 0, 0), (32, 32, 32, 4, 5, 7), (128, 128, 128, 0, 0, 0)]
HERK_SIZES = [(0, 0, 0, 0), (0, 5, 1, 0), (7, 0, 1, 2), (1, 1, 0, 3), (1, 12, 1, 0), (17, 12, 1, 3), (11, 23, 0, 3),
              (9, 12, 1, 1), (32, 32, 0, 0), (32, 32, 4, 7), (128, 128, 0, 0)]
TRSM_SIZES = [(0, 0, 0, 0), (0, 5, 1, 0), (7, 0, 1, 2), (1, 1, 0, 3), (1, 12, 1, 0), (17, 12, 1, 3), (11, 23, 0, 3),
              (9, 12, 1, 1), (32, 32, 0, 0), (32, 32, 4, 7)]


def is_complex(dt):
    return np.issubdtype(dt, np.complexfloating)


def polar(dt, r, theta):
    """TypeUtilities<T>::polar (test/include/dlaf_test/util_types.h:35,:56): r for real T, r e^{i theta} for complex."""
    r, theta = np.asarray(r, dtype=np.float64), np.asarray(theta, dtype=np.float64)
    return (r * np.exp(1j * theta)).astype(dt) if is_complex(dt) else r.astype(dt)


def element(dt, re, im):
    return dt(complex(re, im)) if is_complex(dt) else dt(re)


def type_error(dt):
    return (8 if is_complex(dt) else 2) * np.finfo(dt).eps


def tile(fn, rows, cols, ld, dt, op="N", fill=None):
    """createTile(el, size, ld, op): storage (ld x cols', Fortran) such that op(tile) = el; padding rows keep `fill`."""
    srows, scols = (rows, cols) if op == "N" else (cols, rows)
    buf = np.full((max(ld, 1), max(scols, 1)), 77.0 if fill is None else fill, dtype=dt, order="F")
    if rows and cols:
        i, j = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
        v = fn(i.astype(np.float64), j.astype(np.float64))
        if op == "N":
            buf[:rows, :cols] = v
        elif op == "T":
            buf[:cols, :rows] = v.T
        else:
            buf[:cols, :rows] = np.conj(v.T)
    return buf


def check_near(expected, value, rel, ab):
    """CHECK_TILE_NEAR (test/include/dlaf_test/matrix/util_tile.h): |d| < abs or |d| / max(|e|, |v|) < rel."""
    d = np.abs(expected - value)
    den = np.maximum(np.abs(expected), np.abs(value))
    ok = (d <= ab) | (d <= rel * den)
    assert ok.all(), f"max diff {d.max()} (rel tol {rel}, abs tol {ab})"


@pytest.mark.parametrize("t", list(TYPES))
@pytest.mark.parametrize("opa,opb", [("N", "C"), ("C", "N"), ("N", "N"), ("T", "T"), ("N", "T")])
def test_tile_gemm_closed_form(oracle, t, opa, opb):
    dt = TYPES[t]
    alpha, beta = element(dt, -1.2, .7), element(dt, 1.1, .4)
    for m, n, k, ea, eb, ec in GEMM_SIZES:
        ra, ca = (m, k) if opa == "N" else (k, m)
        rb, cb = (k, n) if opb == "N" else (n, k)
        lda, ldb, ldc = max(1, ra) + ea, max(1, rb) + eb, max(1, m) + ec
        a = tile(lambda i, kk: polar(dt, .9 * (i + 1) / (kk + .5), 2 * i - kk), m, k, lda, dt, opa)
        b = tile(lambda kk, j: polar(dt, .8 * (kk + .5) / (j + 2), kk + j), k, n, ldb, dt, opb)
        c = tile(lambda i, j: polar(dt, 1.2 * i / (j + 1), -i + j), m, n, ldc, dt)
        c0 = c.copy(order="F")
        oracle.tile_gemm(opa, opb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc)
        if m and n:
            i, j = np.meshgrid(np.arange(m, dtype=np.float64), np.arange(n, dtype=np.float64), indexing="ij")
            gamma = element(dt, .72 * k, 0) * alpha
            res = beta * c0[:m, :n] + gamma * polar(dt, (i + 1) / (j + 2), 2 * i + j)
            tol = 2 * (k + 1) * type_error(dt)
            check_near(res, c[:m, :n], tol, tol)
        assert np.array_equal(c[m:, :], c0[m:, :]), "rows beyond the tile were written"


@pytest.mark.parametrize("t", list(TYPES))
@pytest.mark.parametrize("uplo,op", [("L", "N"), ("U", "C"), ("L", "C"), ("U", "N")])
def test_tile_herk_closed_form(oracle, t, uplo, op):
    dt = TYPES[t]
    alpha, beta = -1.2, 1.1
    for n, k, ea, ec in HERK_SIZES:
        ra = n if op == "N" else k
        lda, ldc = max(1, ra) + ea, max(1, n) + ec
        ela = lambda i, kk: polar(dt, .9 * (i + 1) / (kk + .5), i - kk)  # noqa: E731
        a = tile(ela, n, k, lda, dt, op)

        def elc(i, j):
            v = polar(dt, 1.2 * i / (j + 1), -i + j)
            unref = (i < j) if uplo == "L" else (i > j)
            return np.where(unref, dt(-1), v)

        c = tile(elc, n, n, ldc, dt)
        c0 = c.copy(order="F")
        oracle.tile_herk(uplo, op, n, k, alpha, a, lda, beta, c, ldc)
        if n:
            i, kk = np.meshgrid(np.arange(n, dtype=np.float64), np.arange(max(k, 1), dtype=np.float64), indexing="ij")
            opa = ela(i, kk)[:, :k].astype(np.complex128 if is_complex(dt) else np.float64)
            full = beta * c0[:n, :n] + alpha * (opa @ np.conj(opa.T))
            ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
            ref_mask = (ii >= jj) if uplo == "L" else (ii <= jj)
            res = np.where(ref_mask, full, c0[:n, :n])
            tol = (k + 1) * type_error(dt)
            got = c[:n, :n].copy()
            if is_complex(dt):  # ?herk leaves the imaginary part of the diagonal unspecified-but-zero: compare real parts there
                np.fill_diagonal(got, got.diagonal().real)
                np.fill_diagonal(res, res.diagonal().real)
            check_near(res, got, tol, tol)
            assert np.array_equal(c[:n, :n][~ref_mask], c0[:n, :n][~ref_mask]), "unreferenced triangle written"


@pytest.mark.parametrize("t", list(TYPES))
@pytest.mark.parametrize("side,uplo,op,diag", [("R", "L", "C", "N"), ("L", "U", "C", "N"), ("R", "L", "N", "N"), ("L", "L", "N", "U"),
                                               ("R", "U", "T", "N"), ("L", "U", "N", "N"), ("R", "U", "N", "U")])
def test_tile_trsm_closed_form(oracle, t, side, uplo, op, diag):
    dt = TYPES[t]
    alpha = element(dt, -1.2, .7)
    for m, n, ea, eb in TRSM_SIZES:
        na = m if side == "L" else n
        lda, ldb = max(1, na) + ea, max(1, m) + eb
        op_a_lower = (uplo == "L" and op == "N") or (uplo == "U" and op != "N")

        def el_op_a(i, j):
            unref = (i < j) if op_a_lower else (i > j)
            if diag == "U":
                unref = unref | (i == j)
            v = polar(dt, (i + 1) / (j + .5), 2 * i - j) if side == "L" else polar(dt, (j + 1) / (i + .5), 2 * j - i)
            return np.where(unref, dt(-9.9), v)

        def el_x(i, j):
            return polar(dt, (i + .5) / (j + 2), i + j) if side == "L" else polar(dt, (j + .5) / (i + 2), i + j)

        def el_b(i, j):
            if side == "L":
                kk = (i + 1) if op_a_lower else (m - i)
                gamma = polar(dt, (i + 1) / (j + 2), 2 * i + j)
            else:
                kk = (n - j) if op_a_lower else (j + 1)
                gamma = polar(dt, (j + 1) / (i + 2), i + 2 * j)
            kk = kk.astype(np.float64)
            if diag == "U":
                return (((kk - 1) * gamma + el_x(i, j)) / alpha).astype(dt)
            return (kk * gamma / alpha).astype(dt)

        a = tile(el_op_a, na, na, lda, dt, op)
        b = tile(el_b, m, n, ldb, dt)
        b0 = b.copy(order="F")
        oracle.tile_trsm(side, uplo, op, diag, m, n, alpha, a, lda, b, ldb)
        if m and n:
            i, j = np.meshgrid(np.arange(m, dtype=np.float64), np.arange(n, dtype=np.float64), indexing="ij")
            tol = 10 * (m + 1) * type_error(dt)
            check_near(el_x(i, j), b[:m, :n], tol, tol)
        assert np.array_equal(b[m:, :], b0[m:, :]), "rows beyond the tile were written"


@pytest.mark.parametrize("t", list(TYPES))
@pytest.mark.parametrize("uplo", ["L", "U"])
def test_tile_potrf_closed_form_and_info(oracle, t, uplo):
    dt = TYPES[t]
    for n, extra in [(0, 0), (0, 2), (4, 3), (16, 0), (34, 1), (65, 0)]:  # test_lapack_tile.cpp:144-147 style sizes
        a, res = oracle.cholesky_setters(uplo, n, dt)
        big = np.full((max(1, n) + extra, max(1, n)), 55.0, dtype=dt, order="F")
        big[:n, :n] = a
        view = big[:n, :n]
        assert oracle.lapack_potrf(uplo, view, 1) == 0
        tol = 4 * (n + 1) * type_error(dt)
        if n:
            check_near(res, view, tol, tol)
            assert (big[n:, :] == 55.0).all()
    z = np.zeros((8, 8), dtype=dt, order="F")
    assert oracle.lapack_potrf(uplo, z, 1) == 1  # null matrix -> info 1 (test_potrf.h:59-77)
