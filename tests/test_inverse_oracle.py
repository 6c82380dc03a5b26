"""The oracle's restatement of the reference's inverse tile loops — Triangular::call_L / call_U
(inverse/triangular/impl.h:183-229, :367-413) and AssembleCholeskyInverse::call_L / call_U (inverse/cholesky/impl.h:180-224,
:361-405) — pinned to the reference's own closed forms (test/include/dlaf_test/matrix/util_generic_lapack.h:165-196,
:212-247, :261-327) over its test tables (test/unit/inverse/test_triangular_inverse.cpp:54-58,
test_inverse_from_cholesky_factor.cpp:53-57; tolerance 4 (m+1) error), and the numpy model of the product's schedule
(tests/inverse_schedule_model.py, mirrors csrc/inverse_engine.cu) held against the same vectors on simulated grids."""
import numpy as np
import pytest

import inverse_schedule_model as model

TYPES = ["s", "d", "c", "z"]


@pytest.mark.parametrize("t", TYPES)
def test_oracle_triangular_inverse_closed_forms(oracle, t):
    dt = oracle.DTYPES[t]
    for uplo in "LU":
        for diag in "UN":
            for m, mb in oracle.INVERSE_TEST_SIZES:
                a, res = oracle.triangular_inverse_setters(uplo, diag, m, dt)
                oracle.triangular_inverse(uplo, diag, a, mb)
                tol = oracle.inverse_tolerance(m, dt)
                ok, _, msg = oracle.check_near(res, a, tol, tol)
                assert ok, f"{t} {uplo}{diag} m={m} mb={mb}: {msg}"


@pytest.mark.parametrize("t", TYPES)
def test_oracle_assemble_and_inverse_from_factor_closed_forms(oracle, t):
    dt = oracle.DTYPES[t]
    for uplo in "LU":
        for m, mb in oracle.INVERSE_TEST_SIZES:
            tol = oracle.inverse_tolerance(m, dt)
            tt, res = oracle.assemble_cholesky_inverse_setters(uplo, m, dt)
            oracle.assemble_cholesky_inverse(uplo, tt, mb)
            ok, _, msg = oracle.check_near(res, tt, tol, tol)
            assert ok, f"assemble {t} {uplo} m={m} mb={mb}: {msg}"
            tt, res = oracle.inverse_cholesky_factor_setters(uplo, m, dt)
            oracle.inverse_from_cholesky_factor(uplo, tt, mb)
            ok, _, msg = oracle.check_near(res, tt, tol, tol)
            assert ok, f"inverse {t} {uplo} m={m} mb={mb}: {msg}"


def test_closed_forms_in_plain_numpy(oracle):
    """Independent of any BLAS: the generated pairs are inverses / products of each other."""
    dt = np.complex128
    n = 11
    for uplo in "LU":
        tri = np.tril if uplo == "L" else np.triu
        for diag in "UN":
            a, res = oracle.triangular_inverse_setters(uplo, diag, n, dt)
            at, rt = tri(a).copy(), tri(res).copy()
            if diag == "U":
                at[np.arange(n), np.arange(n)] = 1
                rt[np.arange(n), np.arange(n)] = 1
            assert np.allclose(at @ rt, np.eye(n), atol=1e-12)
        t, a = oracle.assemble_cholesky_inverse_setters(uplo, n, dt)
        tt = tri(t)
        full = tt.conj().T @ tt if uplo == "L" else tt @ tt.conj().T
        assert np.allclose(tri(full), tri(a), atol=1e-12)
        t, a = oracle.inverse_cholesky_factor_setters(uplo, n, dt)
        ti = np.linalg.inv(tri(t))
        full = ti.conj().T @ ti if uplo == "L" else ti @ ti.conj().T
        assert np.allclose(tri(full), tri(a), atol=1e-12)


@pytest.mark.parametrize("grid", [(1, 1), (2, 1), (1, 2), (2, 2), (3, 2), (2, 4)])
@pytest.mark.parametrize("t", ["d", "z"])
def test_schedule_model_closed_forms_on_grids(oracle, t, grid):
    """The product's two sweeps (extended column panel, zeroed row, lower-masked assemble update, block-cyclic
    bookkeeping) restated in numpy reproduce the reference's vectors for every grid shape, uplo, diag and size."""
    dt = oracle.DTYPES[t]
    P, Q = grid
    for uplo in "LU":
        for m, mb in oracle.INVERSE_TEST_SIZES:
            tol = oracle.inverse_tolerance(m, dt)
            for diag in "UN":
                a, res = oracle.triangular_inverse_setters(uplo, diag, m, dt)
                out = model.run_user(a, uplo, diag, mb, 4, P, Q, 1)
                ok, _, msg = oracle.check_near(res, out, tol, tol)
                assert ok, f"trtri {uplo}{diag} m={m} mb={mb}: {msg}"
            tt, res = oracle.assemble_cholesky_inverse_setters(uplo, m, dt)
            out = model.run_user(tt, uplo, "N", mb, 4, P, Q, 2)
            ok, _, msg = oracle.check_near(res, out, tol, tol)
            assert ok, f"assemble {uplo} m={m} mb={mb}: {msg}"
            tt, res = oracle.inverse_cholesky_factor_setters(uplo, m, dt)
            out = model.run_user(tt, uplo, "N", mb, 4, P, Q, 3)
            ok, _, msg = oracle.check_near(res, out, tol, tol)
            assert ok, f"inverse {uplo} m={m} mb={mb}: {msg}"


def test_schedule_model_matches_oracle_on_random_factor(oracle):
    n, nb = 96, 16
    a = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    f = a.copy(order="F")
    assert oracle.cholesky_local("L", f, nb) == 0
    ref = f.copy(order="F")
    oracle.inverse_from_cholesky_factor("L", ref, nb)
    out = model.run_user(f, "L", "N", nb, 8, 2, 3, 3)
    assert np.abs(np.tril(out) - np.tril(ref)).max() < 1e-13 * np.abs(ref).max() * n
    assert np.array_equal(np.triu(out, 1), np.triu(f, 1))  # the other triangle is untouched
    inv = np.tril(out) + np.tril(out, -1).T
    assert np.abs(inv @ a - np.eye(n)).max() < 1e-12
