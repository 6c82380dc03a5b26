"""GPU parity of the triangular solver (dlaf_b200_triangular_solver_*, trsm_engine.cu) through the C ABI, mirroring
test/unit/solver/test_triangular.cpp: the reference's closed-form systems for every side / uplo / op / diag combination and
size of its table (element-wise, its tolerance), and larger random systems against the oracle's tile loops."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TYPES = ["s", "d", "c", "z"]


def _alpha(oracle, dt):
    return oracle.TRIANGULAR_TEST_ALPHA if np.dtype(dt).kind == "c" else oracle.TRIANGULAR_TEST_ALPHA.real


@pytest.mark.parametrize("t", TYPES)
def test_closed_form_systems_all_combinations(pkg, oracle, grid11, t):
    dt = pkg.TYPES[t]
    for side, uplo, op, diag in itertools.product("LR", "LU", "NTC", "NU"):
        for m, n, mb, nb in oracle.TRIANGULAR_TEST_SIZES:
            if m == 0 or n == 0:
                continue
            alpha = _alpha(oracle, dt)
            a, b, x = oracle.triangular_system(side, uplo, op, diag, alpha, m, n, dt)
            a0 = a.copy(order="F")
            pkg.triangular_solver(grid11, side, uplo, op, diag, alpha, a, b, mb, nb)
            tol = oracle.triangular_tolerance(m, dt)
            ok, _, msg = oracle.check_near(x, b, tol, tol)
            assert ok, f"{t} {side}{uplo}{op}{diag} m={m} n={n} mb={mb} nb={nb}: {msg}"
            assert np.array_equal(a, a0), "the triangular matrix is read-only"
    assert pkg.last_solver_launch_count(grid11) > 0


@pytest.mark.parametrize("t,m,n,mb,nb", [("d", 1024, 768, 256, 128), ("d", 700, 300, 100, 64), ("z", 512, 384, 128, 128),
                                         ("s", 640, 640, 128, 128), ("c", 300, 200, 96, 64), ("d", 1536, 512, 512, 512)])
@pytest.mark.parametrize("side,uplo,op", [("L", "L", "N"), ("L", "L", "C"), ("L", "U", "N"), ("L", "U", "C"), ("R", "L", "N"),
                                          ("R", "L", "C"), ("R", "U", "T"), ("R", "U", "N")])
def test_random_systems_match_oracle(pkg, oracle, grid11, t, m, n, mb, nb, side, uplo, op):
    """Well-conditioned random triangular systems (the Cholesky factor of the miniapp's matrix) with config-sized tiles:
    product against the oracle's restatement of the reference loops, element-wise at the reference tolerance scaled by the
    solution size, and through the residual op(A) X - alpha B."""
    dt = pkg.TYPES[t]
    rng = np.random.default_rng(3)
    na, ba = (m, mb) if side == "L" else (n, nb)
    spd = oracle.set_random_hermitian_positive_definite(na, ba, dt)
    assert oracle.cholesky_local(uplo, spd, ba) == 0
    tri = np.tril if uplo == "L" else np.triu
    a = np.asfortranarray(tri(spd) + (np.triu(np.full((na, na), -9.9), 1) if uplo == "L" else np.tril(np.full((na, na), -9.9), -1)).astype(dt))
    b = rng.uniform(-1, 1, (m, n))
    if np.dtype(dt).kind == "c":
        b = b + 1j * rng.uniform(-1, 1, (m, n))
    b = np.asfortranarray(b.astype(dt))
    alpha = _alpha(oracle, dt)
    ref = b.copy(order="F")
    oracle.triangular_solver(side, uplo, op, "N", alpha, a, ref, mb, nb)
    out = b.copy(order="F")
    pkg.triangular_solver(grid11, side, uplo, op, "N", alpha, a, out, mb, nb)
    tol = oracle.triangular_tolerance(max(m, n), dt) * max(1.0, float(np.abs(ref).max()))
    ok, _, msg = oracle.check_near(ref, out, tol, tol)
    assert ok, msg
    opa = {"N": tri(spd), "T": tri(spd).T, "C": tri(spd).conj().T}[op].astype(np.complex128 if np.dtype(dt).kind == "c" else np.float64)
    lhs = opa @ out.astype(opa.dtype) if side == "L" else out.astype(opa.dtype) @ opa
    res = np.abs(lhs - alpha * b).max() / (np.abs(opa).max() * np.abs(out).max() * max(m, n))
    assert res < 10 * np.finfo(np.dtype(dt).type(0).real.dtype).eps, res


def test_solve_with_cholesky_factor_end_to_end(pkg, oracle, grid11):
    """POTRF + two triangular solves = the linear solve A x = b every consumer performs (fp64, config-sized tiles)."""
    n, nb, nrhs = 2048, 512, 256
    a = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    f = a.copy(order="F")
    assert pkg.cholesky_factorization(grid11, "L", f, nb) == 0
    rng = np.random.default_rng(9)
    x_true = np.asfortranarray(rng.uniform(-1, 1, (n, nrhs)))
    b = np.asfortranarray(a @ x_true)
    pkg.triangular_solver(grid11, "L", "L", "N", "N", 1.0, f, b, nb, 128)
    pkg.triangular_solver(grid11, "L", "L", "C", "N", 1.0, f, b, nb, 128)
    assert np.abs(b - x_true).max() < 1e-10
