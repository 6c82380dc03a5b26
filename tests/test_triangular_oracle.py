"""The oracle's restatement of the reference's triangular-solver tile loops (solver/triangular/impl.h:236-480) pinned to the
reference's own closed-form systems (test/include/dlaf_test/matrix/util_generic_blas.h:259-371) over the reference's whole
test matrix: sides x uplos x ops x diags x sizes, alpha = (-1.2, 0.7), tolerance 40 (m+1) error
(test/unit/solver/test_triangular.cpp:54-66, :100-101, :143-156)."""
import itertools

import numpy as np
import pytest

TYPES = ["s", "d", "c", "z"]


@pytest.mark.parametrize("t", TYPES)
def test_oracle_reproduces_closed_form_systems(oracle, t):
    dt = oracle.DTYPES[t]
    for side, uplo, op, diag in itertools.product("LR", "LU", "NTC", "NU"):
        for m, n, mb, nb in oracle.TRIANGULAR_TEST_SIZES:
            alpha = oracle.TRIANGULAR_TEST_ALPHA if np.dtype(dt).kind == "c" else oracle.TRIANGULAR_TEST_ALPHA.real
            a, b, x = oracle.triangular_system(side, uplo, op, diag, alpha, m, n, dt)
            oracle.triangular_solver(side, uplo, op, diag, alpha, a, b, mb, nb)
            tol = oracle.triangular_tolerance(m, dt)
            ok, _, msg = oracle.check_near(x, b, tol, tol)
            assert ok, f"{t} {side}{uplo}{op}{diag} m={m} n={n} mb={mb} nb={nb}: {msg}"


def test_closed_form_is_a_solution_in_plain_numpy(oracle):
    """Independent of any BLAS: op(A) X = alpha B / X op(A) = alpha B for the generated triples."""
    dt = np.complex128
    for side, uplo, op, diag in itertools.product("LR", "LU", "NTC", "NU"):
        m, n = 9, 7
        alpha = oracle.TRIANGULAR_TEST_ALPHA
        a, b, x = oracle.triangular_system(side, uplo, op, diag, alpha, m, n, dt)
        na = m if side == "L" else n
        tri = np.tril if uplo == "L" else np.triu
        at = tri(a).copy()
        if diag == "U":
            at[np.arange(na), np.arange(na)] = 1.0
        opa = {"N": at, "T": at.T, "C": at.conj().T}[op]
        lhs = opa @ x if side == "L" else x @ opa
        assert np.allclose(lhs, alpha * b, rtol=1e-12, atol=1e-12), (side, uplo, op, diag)
