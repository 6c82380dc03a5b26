"""The oracle's restatement of the reference's generalized -> standard tile loops (GenToStd::call_L / call_U,
eigensolver/gen_to_std/impl.h:238-281, :507-568) pinned to the reference's closed form (getGenToStdElementSetters, itype 1,
test/include/dlaf_test/matrix/util_generic_lapack.h:94-149) over its test table with its parameters and tolerance
(test/unit/eigensolver/test_gen_to_std.cpp:52-56, :64-66, :78), and the numpy model of the product's schedule
(tests/hegst_schedule_model.py, mirrors csrc/hegst_engine.cu) held against the same vectors on simulated grids."""
import numpy as np
import pytest

import hegst_schedule_model as model

TYPES = ["s", "d", "c", "z"]


@pytest.mark.parametrize("t", TYPES)
def test_oracle_reproduces_closed_form(oracle, t):
    dt = oracle.DTYPES[t]
    for uplo in "LU":
        for m, mb in oracle.GEN_TO_STD_TEST_SIZES:
            tt, a, b = oracle.gen_to_std_setters(uplo, m, dt)
            t0 = tt.copy(order="F")
            oracle.generalized_to_standard(uplo, a, tt, mb)
            ok, _, msg = oracle.check_near(b, a, 0.0, oracle.gen_to_std_tolerance(m, dt))
            assert ok, f"{t} {uplo} m={m} mb={mb}: {msg}"
            assert np.array_equal(tt, t0), "the factor is read-only"


def test_closed_form_in_plain_numpy(oracle):
    n = 10
    for uplo in "LU":
        tt, a, b = oracle.gen_to_std_setters(uplo, n, np.complex128)
        tri = np.tril if uplo == "L" else np.triu
        other = (lambda x: np.tril(x, -1)) if uplo == "L" else (lambda x: np.triu(x, 1))
        full = lambda x: tri(x) + other(x).conj().T
        ti = np.linalg.inv(tri(tt))
        res = ti @ full(a) @ ti.conj().T if uplo == "L" else ti.conj().T @ full(a) @ ti
        assert np.allclose(tri(res), tri(b), atol=1e-13)


@pytest.mark.parametrize("grid", [(1, 1), (2, 1), (1, 2), (2, 2), (3, 2), (2, 4)])
@pytest.mark.parametrize("t", ["d", "z"])
def test_schedule_model_closed_form_on_grids(oracle, t, grid):
    dt = oracle.DTYPES[t]
    P, Q = grid
    for uplo in "LU":
        for m, mb in oracle.GEN_TO_STD_TEST_SIZES:
            tt, a, b = oracle.gen_to_std_setters(uplo, m, dt)
            out = model.run_user(a, tt, uplo, mb, 4, P, Q)
            ok, _, msg = oracle.check_near(b, out, 0.0, oracle.gen_to_std_tolerance(m, dt))
            assert ok, f"{uplo} m={m} mb={mb}: {msg}"


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_schedule_model_matches_oracle_on_random_pencil(oracle, dt):
    n, nb = 96, 16
    a = oracle.set_random_hermitian_positive_definite(n, nb, dt)
    bm = oracle.set_random_hermitian_positive_definite(n, nb, dt)
    bm = np.asfortranarray(bm + bm.conj().T)
    f = bm.copy(order="F")
    assert oracle.cholesky_local("L", f, nb) == 0
    ref = a.copy(order="F")
    oracle.generalized_to_standard("L", ref, f, nb)
    out = model.run_user(a, f, "L", nb, 8, 2, 3)
    assert np.abs(np.tril(out) - np.tril(ref)).max() < 1e-13
    assert np.array_equal(np.triu(out, 1), np.triu(a, 1))
