"""GPU parity tests of the POTRF path, called through the C ABI (dlaf_cholesky_factorization_*,
dlaf_p*potrf) — they mirror the reference's own tests:
  test/unit/factorization/test_cholesky.cpp:54-78   (closed-form golden vectors, sentinel triangle)
  test/unit/c_api/factorization/test_cholesky_c_api.cpp:55-148 (C API, both flavours)
  test/unit/test_lapack_tile/test_potrf.h:59-77     (non-SPD input -> info)
  miniapp/miniapp_cholesky.cpp:408-446              (residual gate on the random HPD input)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TYPES = ["s", "d", "c", "z"]


@pytest.mark.parametrize("uplo", ["L", "U"])
@pytest.mark.parametrize("t", TYPES)
def test_golden_closed_form(pkg, oracle, grid11, t, uplo):
    """Closed-form A -> exact factor, unreferenced triangle = sentinel -9.9 must survive."""
    dt = pkg.TYPES[t]
    for m, mb in oracle.CHOLESKY_TEST_SIZES:
        a, res = oracle.cholesky_setters(uplo, m, dt)
        info = pkg.cholesky_factorization(grid11, uplo, a, mb)
        assert info == 0, (m, mb, info)
        tol = oracle.cholesky_tolerance(m, dt)
        ok, _, msg = oracle.check_near(res, a, tol, tol)
        assert ok, f"type {t} uplo {uplo} m {m} mb {mb}: {msg}"


@pytest.mark.parametrize("uplo", ["L", "U"])
@pytest.mark.parametrize("t", TYPES)
def test_golden_scalapack_api_with_ld(pkg, oracle, grid11, t, uplo):
    """dlaf_p?potrf flavour, local matrix embedded in a larger leading dimension."""
    dt = pkg.TYPES[t]
    for m, mb in oracle.CHOLESKY_TEST_SIZES:
        if m == 0:
            continue
        a, res = oracle.cholesky_setters(uplo, m, dt)
        big = np.full((m + 7, m), 123.0, dtype=dt, order="F")
        big[:m, :] = a
        view = big[:m, :]
        info = pkg.ppotrf(grid11, uplo, view, mb, n=m)
        assert info == 0
        tol = oracle.cholesky_tolerance(m, dt)
        ok, _, msg = oracle.check_near(res, np.asfortranarray(view), tol, tol)
        assert ok, msg
        assert (big[m:, :] == 123.0).all(), "rows beyond the local matrix were written"


@pytest.mark.parametrize("t,n,nb", [("d", 1024, 256), ("d", 1536, 512), ("d", 777, 100), ("s", 1024, 256),
                                    ("z", 512, 128), ("c", 640, 64), ("z", 300, 96)])
@pytest.mark.parametrize("uplo", ["L", "U"])
def test_random_hpd_matches_oracle(pkg, oracle, grid11, t, n, nb, uplo):
    """Same SPD input as the reference miniapp; CUDA result vs the CPU restatement of the reference
    algorithm (element-wise, tolerance of the reference's unit test) and the miniapp residual gate."""
    dt = pkg.TYPES[t]
    a = np.zeros((n, n), dtype=dt, order="F")
    pkg.set_random_hermitian_positive_definite(grid11, a, n, nb)
    a_oracle = oracle.set_random_hermitian_positive_definite(n, nb, dt)
    assert np.array_equal(a, a_oracle), "product and oracle generators disagree"
    ref = a.copy(order="F")
    assert oracle.cholesky_local(uplo, ref, nb) == 0
    out = a.copy(order="F")
    assert pkg.cholesky_factorization(grid11, uplo, out, nb) == 0
    tri = np.tril if uplo == "L" else np.triu
    tol = oracle.cholesky_tolerance(n, dt)
    ok, _, msg = oracle.check_near(tri(ref), tri(out), tol, tol)
    assert ok, msg
    if uplo == "L":
        assert np.array_equal(np.triu(out, 1), np.triu(a, 1)), "upper triangle modified"
    else:
        assert np.array_equal(np.tril(out, -1), np.tril(a, -1)), "lower triangle modified"
    clean, _ = oracle.residual_gate(dt, n)
    res = oracle.residual(uplo, a, out)
    assert res <= clean, f"residual {res} > eps*n {clean}"


@pytest.mark.parametrize("t", TYPES)
def test_not_positive_definite_sets_info(pkg, grid11, t):
    """Zero matrix -> info == 1 (test_potrf.h:59-77); breakdown deeper in the matrix -> its 1-based order."""
    dt = pkg.TYPES[t]
    a = np.zeros((40, 40), dtype=dt, order="F")
    assert pkg.cholesky_factorization(grid11, "L", a, 16) == 1
    n, nb = 300, 64
    b = np.eye(n, dtype=dt, order="F") * 4
    b[200, 200] = -1.0
    assert pkg.cholesky_factorization(grid11, "L", b, nb) == 201
    c = np.eye(n, dtype=dt, order="F") * 4
    c[200, 200] = -1.0
    assert pkg.cholesky_factorization(grid11, "U", c, nb) == 201


def test_device_resident_in_place(pkg, oracle, grid11):
    """dlaf::cholesky_factorization<Backend::GPU, Device::GPU> analogue: device pointer, asynchronous,
    in place when no padding is needed; also the padded / upper route through the internal slab."""
    import torch

    for n, nb, uplo in [(1024, 256, "L"), (1024, 256, "U"), (600, 100, "L")]:
        a = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
        ref = a.copy(order="F")
        assert oracle.cholesky_local(uplo, ref, nb) == 0
        dev = torch.from_numpy(np.ascontiguousarray(a.T)).cuda()  # row-major of A^T == column-major of A
        stream = torch.cuda.current_stream().cuda_stream
        pkg.cholesky_factorization_device(grid11, uplo, dev.data_ptr(), np.float64, n, nb, n, stream)
        assert pkg.wait(grid11, stream) == 0
        out = np.asfortranarray(dev.cpu().numpy().T)
        tri = np.tril if uplo == "L" else np.triu
        tol = oracle.cholesky_tolerance(n, np.float64)
        ok, _, msg = oracle.check_near(tri(ref), tri(out), tol, tol)
        assert ok, msg
        other = np.triu(out, 1) - np.triu(a, 1) if uplo == "L" else np.tril(out, -1) - np.tril(a, -1)
        assert not other.any()
        assert pkg.last_launch_count(grid11) > 0


def test_large_property_residual(pkg, oracle, grid11):
    """Config-sized tiles (nb = 512): parity through the size-independent property ||A - L L^T|| on the
    reference's own input; N chosen so that the oracle-side check stays within seconds."""
    n, nb = 4096, 512
    a = np.zeros((n, n), dtype=np.float64, order="F")
    pkg.set_random_hermitian_positive_definite(grid11, a, n, nb)
    out = a.copy(order="F")
    assert pkg.cholesky_factorization(grid11, "L", out, nb) == 0
    res = oracle.residual("L", a, out)
    assert res <= oracle.residual_gate(np.float64, n)[0]
    ref = a.copy(order="F")
    oracle.cholesky_local("L", ref, nb, nthreads=8)
    tol = oracle.cholesky_tolerance(n, np.float64)
    ok, _, msg = oracle.check_near(np.tril(ref), np.tril(out), tol, tol)
    assert ok, msg


def test_dense_indefinite_reports_first_failing_minor(pkg, oracle, grid11):
    """A dense matrix that stops being positive definite in the middle: info = order of the FIRST non-positive
    leading minor (what LAPACK's potrf reports; the oracle's tile loop returns the same)."""
    n, nb = 700, 64
    a = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    a[325, 325] = -3.0
    want = oracle.cholesky_local("L", a.copy(order="F"), nb)
    assert want == 326
    assert pkg.cholesky_factorization(grid11, "L", a.copy(order="F"), nb) == want
    assert pkg.cholesky_factorization(grid11, "U", a.copy(order="F"), nb) == want


@pytest.mark.parametrize("t,n,nb,uplo", [("d", 1024, 256, "L"), ("d", 777, 100, "U"), ("z", 512, 128, "L"),
                                         ("s", 640, 128, "U"), ("c", 300, 96, "L")])
def test_grid_residual_check_matches_oracle_residual(pkg, oracle, grid11, t, n, nb, uplo):
    """The product's result check (engine_check.cu, the distributed restatement of miniapp_cholesky.cpp:408-446) against
    the oracle's residual of the same factor: same max-norm ratio up to rounding of the two evaluations; a corrupted
    factor is seen."""
    dt = pkg.TYPES[t]
    a = oracle.set_random_hermitian_positive_definite(n, nb, dt)
    f = a.copy(order="F")
    assert pkg.cholesky_factorization(grid11, uplo, f, nb) == 0
    r_gpu = pkg.check_cholesky(grid11, uplo, a, f, nb)
    r_cpu = oracle.residual(uplo, a, f)
    gate = oracle.residual_gate(dt, n)[0]
    assert 0 <= r_gpu <= gate and r_cpu <= gate
    assert abs(r_gpu - r_cpu) <= 0.5 * max(r_gpu, r_cpu) + 1e-3 * gate, (r_gpu, r_cpu)
    g = f.copy(order="F")
    g[n // 2, n // 3 if uplo == "L" else n // 2 + 5] += 0.25
    assert pkg.check_cholesky(grid11, uplo, a, g, nb) > 10 * gate


def test_baseline_config_size_elementwise_vs_oracle(pkg, oracle, grid11):
    """BASELINE config C2 (fp64 N=16384 nb=512, one GPU) compared ELEMENT-WISE with the reference algorithm's factor at the
    reference's unit-test tolerance, not only through the residual (the oracle needs ~10-20 s of host time)."""
    n, nb = 16384, 512
    a = np.zeros((n, n), dtype=np.float64, order="F")
    pkg.set_random_hermitian_positive_definite(grid11, a, n, nb)
    out = a.copy(order="F")
    assert pkg.cholesky_factorization(grid11, "L", out, nb) == 0
    ref = a  # factor the input in place on the CPU (saves 2 GB)
    assert oracle.cholesky_local("L", ref, nb, nthreads=oracle.max_pool_threads()) == 0
    tol = oracle.cholesky_tolerance(n, np.float64)
    bs = 2048
    for j0 in range(0, n, bs):  # block columns of the lower triangle
        e = np.tril(ref[j0:, j0:j0 + bs], 0) if j0 == 0 else ref[j0:, j0:j0 + bs].copy()
        v = out[j0:, j0:j0 + bs].copy()
        if j0:
            iu = np.triu_indices(bs, 1)
            e[iu] = 0
            v[iu] = 0
        else:
            v = np.tril(v, 0)
        ok, _, msg = oracle.check_near(e, v, tol, tol)
        assert ok, f"block column {j0}: {msg}"
