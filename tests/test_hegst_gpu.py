"""GPU parity of the generalized -> standard reduction (dlaf_b200_generalized_to_standard_*, hegst_engine.cu) through the C ABI,
mirroring test/unit/eigensolver/test_gen_to_std.cpp: the reference's closed form for both uplos and every size of its
table (its absolute tolerance, sentinel triangle and the factor untouched), then config-sized random pencils against the
oracle's restatement of the reference loops and through the eigenvalues of the pencil."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TYPES = ["s", "d", "c", "z"]


@pytest.mark.parametrize("t", TYPES)
def test_closed_form(pkg, oracle, grid11, t):
    dt = pkg.TYPES[t]
    for uplo in "LU":
        for m, mb in oracle.GEN_TO_STD_TEST_SIZES:
            if m == 0:
                continue
            tt, a, b = oracle.gen_to_std_setters(uplo, m, dt)
            t0 = tt.copy(order="F")
            assert pkg.generalized_to_standard(grid11, uplo, a, tt, mb) == 0
            ok, _, msg = oracle.check_near(b, a, 0.0, oracle.gen_to_std_tolerance(m, dt))  # includes the -9.9 sentinels
            assert ok, f"{t} {uplo} m={m} mb={mb}: {msg}"
            assert np.array_equal(tt, t0), "the factor is read-only"
    assert pkg.last_solver_launch_count(grid11) > 0


def test_empty_matrix_is_a_no_op(pkg, grid11):
    a = np.zeros((0, 0), dtype=np.float64, order="F")
    assert pkg.generalized_to_standard(grid11, "L", a, a.copy(order="F"), 2, n=0) == 0


def _sentinel_triangle(n, uplo, dt):
    s = np.full((n, n), -9.9)
    return (np.triu(s, 1) if uplo == "L" else np.tril(s, -1)).astype(dt)


@pytest.mark.parametrize("t,n,nb", [("d", 2048, 512), ("d", 1500, 200), ("d", 1024, 128), ("s", 2048, 1024), ("s", 1024, 256),
                                    ("z", 1024, 256), ("c", 768, 128), ("z", 600, 100)])
@pytest.mark.parametrize("uplo", ["L", "U"])
def test_random_pencil_matches_oracle(pkg, oracle, grid11, t, n, nb, uplo):
    dt = pkg.TYPES[t]
    a = oracle.set_random_hermitian_positive_definite(n, nb, dt)
    bm = oracle.set_random_hermitian_positive_definite(n, nb, dt)
    rng = np.random.default_rng(5)
    pert = rng.uniform(-1, 1, (n, n))
    bm = np.asfortranarray((bm + (pert + pert.T)).astype(dt))  # another HPD matrix (diagonal 2n dominates)
    f = bm.copy(order="F")
    assert oracle.cholesky_local(uplo, f, nb, nthreads=8) == 0
    tri = np.tril if uplo == "L" else np.triu
    f = np.asfortranarray(tri(f) + _sentinel_triangle(n, uplo, dt))
    a_in = np.asfortranarray(tri(a) + _sentinel_triangle(n, uplo, dt))
    ref = a_in.copy(order="F")
    oracle.generalized_to_standard(uplo, ref, f, nb)
    out = a_in.copy(order="F")
    f0 = f.copy(order="F")
    assert pkg.generalized_to_standard(grid11, uplo, out, f, nb) == 0
    assert np.array_equal(f, f0)
    tol = oracle.gen_to_std_tolerance(n, dt) * max(1.0, float(np.abs(tri(ref)).max()))
    ok, _, msg = oracle.check_near(ref, out, 0.0, tol)  # sentinels included: the other triangle must stay untouched
    assert ok, msg


def test_eigenvalues_of_the_pencil(pkg, oracle, grid11):
    """POTRF(B) + generalized_to_standard(A, L) through the C ABI: the standard problem has the eigenvalues of (A, B)."""
    import scipy.linalg as sla

    n, nb = 1024, 256
    a = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    rng = np.random.default_rng(11)
    m = rng.uniform(-1, 1, (n, n))
    a = np.asfortranarray(a + 40 * (m + m.T))  # indefinite symmetric A
    b = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    f = b.copy(order="F")
    assert pkg.cholesky_factorization(grid11, "L", f, nb) == 0
    c = a.copy(order="F")
    assert pkg.generalized_to_standard(grid11, "L", c, f, nb) == 0
    cs = np.tril(c) + np.tril(c, -1).T
    ev = np.linalg.eigvalsh(cs)
    ref = sla.eigh(a, b, eigvals_only=True)
    assert np.abs(ev - ref).max() < 1e-11 * np.abs(ref).max() * n


def test_both_fp64_engines_agree(pkg, oracle, grid11, monkeypatch):
    n, nb = 1536, 256
    a = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    rng = np.random.default_rng(21)
    m = rng.uniform(-1, 1, (n, n))
    a = np.asfortranarray(a + (m + m.T))  # (the generator is deterministic: without this A would be equal to B)
    b = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    f = b.copy(order="F")
    assert oracle.cholesky_local("L", f, nb, nthreads=8) == 0
    out = a.copy(order="F")
    assert pkg.generalized_to_standard(grid11, "L", out, f, nb) == 0
    assert pkg.last_inverse_guard_steps(grid11) == 0
    monkeypatch.setenv("DLAF_B200_D_BULK", "dmma")
    out2 = a.copy(order="F")
    assert pkg.generalized_to_standard(grid11, "L", out2, f, nb) == 0
    assert np.abs(np.tril(out) - np.tril(out2)).max() < 1e-13 * max(1.0, np.abs(np.tril(out2)).max())


def test_identical_pencil_gives_identity_and_trips_the_guard(pkg, oracle, grid11):
    """A == B: the reduced matrix is the identity up to rounding, i.e. the diagonal tiles hold entries ~1e-16 next to 1 — rows
    spanning > 40 binades, exactly what the int8-digit guard is for: those steps must fall back to native fp64 and the
    result must be as accurate as the oracle's."""
    n, nb = 1536, 256
    b = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    f = b.copy(order="F")
    assert oracle.cholesky_local("L", f, nb, nthreads=8) == 0
    out = b.copy(order="F")
    assert pkg.generalized_to_standard(grid11, "L", out, f, nb) == 0
    assert pkg.last_inverse_guard_steps(grid11) > 0
    ref = b.copy(order="F")
    oracle.generalized_to_standard("L", ref, f, nb)
    assert np.abs(np.tril(out) - np.eye(n)).max() < 1e-13
    assert np.abs(np.tril(out) - np.tril(ref)).max() < 1e-13
