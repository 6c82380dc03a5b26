"""GPU parity of the two fp64 trailing-update engines behind dlaf_cholesky_factorization_d: the native DMMA GEMM and
the tcgen05 int8 Ozaki-scheme GEMM (gemm_ozaki_i8.cu, DLAF_B200_D_BULK=ozaki|dmma). Both must meet the reference's
unit-test tolerance against the oracle (test/unit/factorization/test_cholesky.cpp:76-77) and the miniapp residual
gate (miniapp/miniapp_cholesky.cpp:408-446); the emulated path is additionally held to a few ulps of the native one
(it is NOT a reduced-precision path: 8 x 7-bit exact digit products, truncation 2^-55)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _factor(pkg, engine, uplo, a, nb):
    old = os.environ.get("DLAF_B200_D_BULK")
    os.environ["DLAF_B200_D_BULK"] = engine
    try:
        pkg.initialize()
        ctx = pkg.create_grid(None, 1, 1, "R")  # fresh context -> fresh engine that reads the switch
        try:
            out = a.copy(order="F")
            assert pkg.cholesky_factorization(ctx, uplo, out, nb) == 0
            launches = pkg.last_launch_count(ctx)
        finally:
            pkg.free_grid(ctx)
    finally:
        if old is None:
            del os.environ["DLAF_B200_D_BULK"]
        else:
            os.environ["DLAF_B200_D_BULK"] = old
    return out, launches


@pytest.mark.parametrize("n,nb", [(1024, 256), (1536, 512), (2048, 128), (1920, 384), (777, 100)])
@pytest.mark.parametrize("uplo", ["L", "U"])
def test_ozaki_and_dmma_bulk_match_oracle(pkg, oracle, n, nb, uplo):
    dt = np.float64
    a = oracle.set_random_hermitian_positive_definite(n, nb, dt)
    ref = a.copy(order="F")
    assert oracle.cholesky_local(uplo, ref, nb) == 0
    tri = np.tril if uplo == "L" else np.triu
    tol = oracle.cholesky_tolerance(n, dt)
    gate, _ = oracle.residual_gate(dt, n)
    outs = {}
    for engine in ("dmma", "ozaki"):
        out, launches = _factor(pkg, engine, uplo, a, nb)
        assert launches > 0
        ok, _, msg = oracle.check_near(tri(ref), tri(out), tol, tol)
        assert ok, f"{engine}: {msg}"
        res = oracle.residual(uplo, a, out)
        assert res <= gate, f"{engine}: residual {res} > eps*n {gate}"
        other = np.triu(out, 1) == np.triu(a, 1) if uplo == "L" else np.tril(out, -1) == np.tril(a, -1)
        assert other.all(), f"{engine}: unreferenced triangle modified"
        outs[engine] = tri(out)
    # emulated vs native: a few ulps of the largest factor entry
    diff = np.abs(outs["ozaki"] - outs["dmma"]).max()
    assert diff <= 64 * np.finfo(dt).eps * np.abs(outs["dmma"]).max(), diff


def test_ozaki_wide_dynamic_range(pkg, oracle):
    """Badly scaled SPD matrix D A D (D = powers of two over 40 binades): row-wise digit scaling must cope."""
    n, nb = 1024, 256
    rng = np.random.default_rng(5)
    a = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    d = np.ldexp(1.0, rng.integers(-20, 20, n))
    a = np.asfortranarray(a * d[:, None] * d[None, :])
    ref = a.copy(order="F")
    assert oracle.cholesky_local("L", ref, nb) == 0
    out, _ = _factor(pkg, "ozaki", "L", a, nb)
    # row i of L scales with d[i]: compare after undoing the scaling
    tol = oracle.cholesky_tolerance(n, np.float64)
    ok, _, msg = oracle.check_near(np.tril(ref) / d[:, None], np.tril(out) / d[:, None], tol, tol)
    assert ok, msg


@pytest.mark.parametrize("engine", ["ozaki", "dmma"])
def test_nan_in_the_input_is_not_swallowed(pkg, oracle, engine):
    """A NaN below the diagonal must surface (non-zero info, like a non-SPD input), never be rounded to a digit."""
    n, nb = 1024, 256
    a = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    a[n - 3, 5] = np.nan
    old = os.environ.get("DLAF_B200_D_BULK")
    os.environ["DLAF_B200_D_BULK"] = engine
    try:
        pkg.initialize()
        ctx = pkg.create_grid(None, 1, 1, "R")
        try:
            out = a.copy(order="F")
            info = pkg.cholesky_factorization(ctx, "L", out, nb)
        finally:
            pkg.free_grid(ctx)
    finally:
        if old is None:
            del os.environ["DLAF_B200_D_BULK"]
        else:
            os.environ["DLAF_B200_D_BULK"] = old
    assert info != 0 or np.isnan(np.tril(out)).any()
