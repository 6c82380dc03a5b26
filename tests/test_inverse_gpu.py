"""GPU parity of the inverse algorithms (dlaf_inverse_from_cholesky_factor_*, dlaf_p?potri, dlaf_b200_triangular_inverse_*,
inverse_engine.cu) through the C ABI, mirroring test/unit/inverse/test_triangular_inverse.cpp and
test_inverse_from_cholesky_factor.cpp: the reference's closed forms for every uplo / diag and size of its tables
(element-wise, its tolerance, sentinel triangle untouched), then config-sized random factors against the oracle's
restatement of the reference loops and through A inv(A) = I."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TYPES = ["s", "d", "c", "z"]


@pytest.mark.parametrize("t", TYPES)
def test_triangular_inverse_closed_forms(pkg, oracle, grid11, t):
    dt = pkg.TYPES[t]
    for uplo in "LU":
        for diag in "UN":
            for m, mb in oracle.INVERSE_TEST_SIZES:
                if m == 0:
                    continue
                a, res = oracle.triangular_inverse_setters(uplo, diag, m, dt)
                assert pkg.triangular_inverse(grid11, uplo, diag, a, mb) == 0
                tol = oracle.inverse_tolerance(m, dt)
                ok, _, msg = oracle.check_near(res, a, tol, tol)  # includes the -9.9 sentinels (and the unit diagonal)
                assert ok, f"{t} {uplo}{diag} m={m} mb={mb}: {msg}"
    assert pkg.last_solver_launch_count(grid11) > 0


@pytest.mark.parametrize("t", TYPES)
def test_inverse_from_cholesky_factor_closed_forms(pkg, oracle, grid11, t):
    dt = pkg.TYPES[t]
    for uplo in "LU":
        for m, mb in oracle.INVERSE_TEST_SIZES:
            if m == 0:
                continue
            tol = oracle.inverse_tolerance(m, dt)
            tt, res = oracle.assemble_cholesky_inverse_setters(uplo, m, dt)
            assert pkg.assemble_cholesky_inverse(grid11, uplo, tt, mb) == 0
            ok, _, msg = oracle.check_near(res, tt, tol, tol)
            assert ok, f"assemble {t} {uplo} m={m} mb={mb}: {msg}"
            tt, res = oracle.inverse_cholesky_factor_setters(uplo, m, dt)
            assert pkg.inverse_from_cholesky_factor(grid11, uplo, tt, mb) == 0
            ok, _, msg = oracle.check_near(res, tt, tol, tol)
            assert ok, f"inverse {t} {uplo} m={m} mb={mb}: {msg}"
            tt, res = oracle.inverse_cholesky_factor_setters(uplo, m, dt)
            assert pkg.ppotri(grid11, uplo, tt, mb) == 0  # ScaLAPACK flavour
            ok, _, msg = oracle.check_near(res, tt, tol, tol)
            assert ok, f"p?potri {t} {uplo} m={m} mb={mb}: {msg}"


def test_empty_matrix_is_a_no_op(pkg, grid11):
    a = np.zeros((0, 0), dtype=np.float64, order="F")
    assert pkg.inverse_from_cholesky_factor(grid11, "L", a, 2, n=0) == 0
    assert pkg.triangular_inverse(grid11, "U", "N", a, 2, n=0) == 0


def _sentinel_triangle(n, uplo, dt):
    s = np.full((n, n), -9.9)
    return (np.triu(s, 1) if uplo == "L" else np.tril(s, -1)).astype(dt)


@pytest.mark.parametrize("t,n,nb", [("d", 2048, 512), ("d", 1500, 200), ("d", 1024, 128), ("s", 2048, 1024), ("s", 1024, 256),
                                    ("z", 1024, 256), ("c", 768, 128), ("z", 600, 100)])
@pytest.mark.parametrize("uplo", ["L", "U"])
def test_random_factor_matches_oracle(pkg, oracle, grid11, t, n, nb, uplo):
    """POTRF of the miniapp's matrix (oracle) -> inverse from the factor: product vs the oracle's tile loops element-wise
    (reference tolerance scaled by max|inv(A)| ~ 1/(2n) ... the entries are O(1/n), so the comparison is made relative to
    the largest entry), the untouched triangle, and A inv(A) = I."""
    dt = pkg.TYPES[t]
    a = oracle.set_random_hermitian_positive_definite(n, nb, dt)
    f = a.copy(order="F")
    assert oracle.cholesky_local(uplo, f, nb, nthreads=8) == 0
    tri = np.tril if uplo == "L" else np.triu
    f = np.asfortranarray(tri(f) + _sentinel_triangle(n, uplo, dt))
    ref = f.copy(order="F")
    oracle.inverse_from_cholesky_factor(uplo, ref, nb)
    out = f.copy(order="F")
    assert pkg.inverse_from_cholesky_factor(grid11, uplo, out, nb) == 0
    scale = float(np.abs(tri(ref)).max())
    tol = oracle.inverse_tolerance(n, dt)
    ok, _, msg = oracle.check_near(tri(ref) / scale, tri(out) / scale, tol, tol)
    assert ok, msg
    other = (lambda x: np.triu(x, 1)) if uplo == "L" else (lambda x: np.tril(x, -1))
    assert np.array_equal(other(out), other(f)), "the other triangle must stay untouched"
    wide = np.complex128 if np.dtype(dt).kind == "c" else np.float64
    o = tri(out).astype(wide)
    inv = o + (np.tril(o, -1).conj().T if uplo == "L" else np.triu(o, 1).conj().T)
    eps = float(np.finfo(np.dtype(dt).type(0).real.dtype).eps)
    assert np.abs(inv @ a.astype(wide) - np.eye(n)).max() < 50 * n * eps


@pytest.mark.parametrize("uplo,diag", [("L", "N"), ("L", "U"), ("U", "N"), ("U", "U")])
def test_random_triangular_inverse_matches_oracle(pkg, oracle, grid11, uplo, diag):
    n, nb = 1536, 512
    a = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    f = a.copy(order="F")
    assert oracle.cholesky_local(uplo, f, nb, nthreads=8) == 0
    tri = np.tril if uplo == "L" else np.triu
    if diag == "U":  # scale to a unit diagonal so that the matrix stays well conditioned
        d = np.diag(f).copy()
        f = f / d[None, :] if uplo == "L" else f / d[:, None]
    f = np.asfortranarray(tri(f) + _sentinel_triangle(n, uplo, np.float64))
    ref = f.copy(order="F")
    oracle.triangular_inverse(uplo, diag, ref, nb)
    out = f.copy(order="F")
    assert pkg.triangular_inverse(grid11, uplo, diag, out, nb) == 0
    scale = float(np.abs(tri(ref)).max())
    tol = oracle.inverse_tolerance(n, np.float64)
    ok, _, msg = oracle.check_near(ref / scale, out / scale, tol, tol)  # full matrices: sentinels (and unit diagonal) included
    assert ok, msg


def test_both_fp64_engines_and_guard_report(pkg, oracle, grid11, monkeypatch):
    """int8-digit engine (default) and native fp64 give the same inverse to rounding; the guard counter is readable."""
    n, nb = 1536, 256
    a = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    f = a.copy(order="F")
    assert oracle.cholesky_local("L", f, nb, nthreads=8) == 0
    out = f.copy(order="F")
    assert pkg.inverse_from_cholesky_factor(grid11, "L", out, nb) == 0
    assert pkg.last_inverse_guard_steps(grid11) == 0
    monkeypatch.setenv("DLAF_B200_D_BULK", "dmma")
    out2 = f.copy(order="F")
    assert pkg.inverse_from_cholesky_factor(grid11, "L", out2, nb) == 0
    assert np.abs(np.tril(out) - np.tril(out2)).max() < 1e-13 * np.abs(np.tril(out2)).max() * 16


def test_potrf_then_potri_end_to_end(pkg, oracle, grid11):
    """dlaf_pdpotrf + dlaf_pdpotri = the matrix inverse a ScaLAPACK application computes (config-sized tiles)."""
    n, nb = 4096, 512
    a = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    f = a.copy(order="F")
    assert pkg.ppotrf(grid11, "L", f, nb) == 0
    assert pkg.ppotri(grid11, "L", f, nb) == 0
    inv = np.tril(f) + np.tril(f, -1).T
    assert np.abs(inv @ a - np.eye(n)).max() < 1e-11
    assert pkg.last_solver_device_ms(grid11) > 0
