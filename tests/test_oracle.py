"""CPU tests (no GPU): the oracle is pinned against the reference's own golden vectors, and the host
logic of the product (C ABI loading, descriptors, block-cyclic sizes, input generator) is checked."""
import ctypes
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("uplo", ["L", "U"])
@pytest.mark.parametrize("t", ["s", "d", "c", "z"])
def test_oracle_reproduces_reference_golden_vectors(oracle, t, uplo):
    """getCholeskySetters closed form (test/include/dlaf_test/matrix/util_generic_lapack.h:39-68), sizes and
    tolerance of test/unit/factorization/test_cholesky.cpp:54-78; the sentinel triangle must survive."""
    dt = oracle.DTYPES[t]
    for m, mb in oracle.CHOLESKY_TEST_SIZES:
        for nthreads in (1, 3):
            a, res = oracle.cholesky_setters(uplo, m, dt)
            assert oracle.cholesky_local(uplo, a, mb, nthreads) == 0
            tol = oracle.cholesky_tolerance(m, dt)
            ok, _, msg = oracle.check_near(res, a, tol, tol)
            assert ok, msg


def test_oracle_info_on_non_spd(oracle):
    """test/unit/test_lapack_tile/test_potrf.h:59-77: zero matrix -> info == 1."""
    for t, dt in oracle.DTYPES.items():
        a = np.zeros((12, 12), dtype=dt, order="F")
        assert oracle.cholesky_local("L", a, 5) == 1


def test_index_math_against_reference_table(oracle):
    """Golden index-conversion table of the reference's own unit test."""
    with open(os.path.join(HERE, "golden", "util_distribution_cases.json")) as f:
        g = json.load(f)
    for row in g["cases"]:
        c = dict(zip(g["keys"], row))
        assert c["global_element"] // c["tile_size"] == c["global_tile"]
        assert c["global_element"] % c["tile_size"] == c["tile_element"]
        assert oracle.rank_global_tile(c["global_tile"], c["grid_size"], c["src_rank"]) == c["rank_tile"]
        assert oracle.local_tile_from_global_tile(c["global_tile"], c["grid_size"], c["rank"], c["src_rank"]) == c["local_tile"]
        assert oracle.next_local_tile_from_global_tile(c["global_tile"], c["grid_size"], c["rank"], c["src_rank"]) == c["local_tile_next"]
        if c["local_tile"] >= 0:
            assert oracle.global_tile_from_local_tile(c["local_tile"], c["grid_size"], c["rank"], c["src_rank"]) == c["global_tile"]


def test_product_index_math_against_reference_table(pkg):
    """The same golden table against the PRODUCT's index functions (csrc/distribution.h, exported through the C ABI and
    used by the engine / C API for every owner / local-index computation)."""
    L = pkg.lib()
    with open(os.path.join(HERE, "golden", "util_distribution_cases.json")) as f:
        g = json.load(f)
    for row in g["cases"]:
        c = dict(zip(g["keys"], row))
        assert L.dlaf_b200_rank_global_tile(c["global_tile"], c["grid_size"], c["src_rank"]) == c["rank_tile"]
        assert L.dlaf_b200_local_tile_from_global_tile(c["global_tile"], c["grid_size"], c["rank"], c["src_rank"]) == c["local_tile"]
        assert L.dlaf_b200_next_local_tile_from_global_tile(c["global_tile"], c["grid_size"], c["rank"], c["src_rank"]) == c["local_tile_next"]
        if c["local_tile"] >= 0:
            assert L.dlaf_b200_global_tile_from_local_tile(c["local_tile"], c["grid_size"], c["rank"], c["src_rank"]) == c["global_tile"]


def test_random_hpd_properties(oracle):
    """include/dlaf/util_matrix.h:410-453: Hermitian, real diagonal in [2N-1, 2N+1], |offdiag| <= 1,
    values independent of the tile size only through the per-tile seeds (same nb -> same matrix)."""
    for t, dt in oracle.DTYPES.items():
        n, nb = 70, 16
        a = oracle.set_random_hermitian_positive_definite(n, nb, dt)
        assert np.array_equal(a, a.conj().T)
        d = a.diagonal()
        assert np.all(np.imag(d) == 0) and np.all(np.abs(np.real(d) - 2 * n) <= 1)
        off = a - np.diag(d)
        assert np.abs(off).max() <= 1.0 + 1e-6
        assert np.array_equal(a, oracle.set_random_hermitian_positive_definite(n, nb, dt))
        assert np.linalg.eigvalsh(a.astype(np.complex128)).min() > 0


def test_c_abi_library_loads_and_exports_every_symbol(pkg):
    """Every symbol include/dlaf_c/*.h declares is exported (no compute call: works without a GPU)."""
    L = pkg.lib()
    for sym in pkg.C_API_SYMBOLS:
        assert hasattr(L, sym), sym
    # and the headers really declare them
    inc = os.path.join(os.path.dirname(HERE), "include", "dlaf_c")
    text = ""
    for root, _, files in os.walk(inc):
        for f in files:
            text += open(os.path.join(root, f)).read()
    for sym in pkg.C_API_SYMBOLS:
        assert sym in text, f"{sym} not declared in include/dlaf_c"


def test_make_dlaf_descriptor(pkg):
    """src/c_api/utils.cpp:26-34: {dtype, ctxt, m, n, mb, nb, rsrc, csrc, lld} -> DLAF_descriptor."""
    desca = (ctypes.c_int * 9)(1, 77, 100, 100, 16, 16, 1, 2, 64)
    d = pkg.lib().make_dlaf_descriptor(100, 100, 1, 1, desca)
    assert (d.m, d.n, d.mb, d.nb, d.isrc, d.jsrc, d.i, d.j, d.ld) == (100, 100, 16, 16, 1, 2, 0, 0, 64)


def test_grid_contexts_and_local_sizes(pkg, oracle):
    """Contexts count down from INT_MAX (src/c_api/grid.cpp:28-40); local sizes follow
    src/matrix/distribution.cpp:117-150 (1x1 grid here; P x Q is covered by the gloo test)."""
    c1 = pkg.create_grid(None, 1, 1, "R")
    c2 = pkg.create_grid(None, 1, 1, "C")
    assert c1 - c2 == 1 and c1 <= 2**31 - 1
    assert pkg.grid_info(c1) == (1, 1, 0, 0)
    for n, nb in [(0, 2), (5, 8), (34, 13), (4096, 256)]:
        d = pkg.descriptor(n, nb, max(1, n))
        assert pkg.local_shape(c1, d) == (n, n)
        assert oracle.local_size(n, nb, 1, 0, 0) == n
    pkg.free_grid(c1)
    pkg.free_grid(c2)


@pytest.mark.parametrize("t", ["s", "d", "c", "z"])
def test_product_generator_matches_oracle_generator(pkg, oracle, t):
    """Two independent restatements of set_random_hermitian_positive_definite agree bit for bit,
    including ragged edge tiles and a leading dimension larger than the matrix."""
    dt = pkg.TYPES[t]
    ctx = pkg.create_grid(None, 1, 1, "R")
    try:
        for n, nb in [(37, 8), (64, 16), (50, 64)]:
            big = np.zeros((n + 3, n), dtype=dt, order="F")
            view = big[:n, :]
            pkg.set_random_hermitian_positive_definite(ctx, view, n, nb)
            ref = oracle.set_random_hermitian_positive_definite(n, nb, dt)
            assert np.array_equal(np.asfortranarray(view), ref)
            assert not big[n:, :].any()
    finally:
        pkg.free_grid(ctx)


def test_scatter_gather_roundtrip(oracle):
    a = np.arange(35 * 35, dtype=np.float64).reshape(35, 35, order="F")
    for grid, src in [((2, 3), (0, 0)), ((3, 2), (2, 1)), ((1, 4), (0, 3))]:
        parts = oracle.scatter_block_cyclic(a, 4, grid, src)
        for (p, q), loc in parts.items():
            assert loc.shape == (oracle.local_size(35, 4, grid[0], p, src[0]), oracle.local_size(35, 4, grid[1], q, src[1]))
        assert np.array_equal(oracle.gather_block_cyclic(parts, 35, 4, grid, np.float64, src), a)


def test_headers_are_c_and_library_links_with_c_linkage(pkg, tmp_path):
    """gcc (C, not C++) compiles the public headers and links the library — the reference does the same with
    test/unit/c_api/factorization/test_cholesky_c_api_wrapper.c."""
    import subprocess

    root = os.path.dirname(HERE)
    exe = str(tmp_path / "c_api_wrapper")
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(HERE, "c_api_wrapper.c"), "-o", exe, "-L", libdir, "-ldlaf_b200",
                           f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert "ctx_ok 1 desc 300 300 64 64 ld 300 grid 1x1 rank 0,0 local 300x300" in out.stdout
    assert "symbols 1" in out.stdout
