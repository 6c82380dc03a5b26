"""CPU checks of the arithmetic the int8 (Ozaki-scheme) fp64 update relies on — the same formulas as
dla-future_b200/csrc/gemm_ozaki_i8.cu (split_i8_kernel, ozaki_fold_store), restated in numpy (tests/ozaki_model.py):
digit range, slicing error bound, int32 headroom of the group sums, exactness of the two-integer fold and of the
magic-number int64 -> fp64 conversion, the end-to-end product error next to a native fp64 GEMM, and the guard rule."""
import numpy as np

import ozaki_model as P

S = P.S


def test_digits_are_int8_and_slicing_error_is_bounded():
    rng = np.random.default_rng(3)
    for scale in (1.0, 1e-200, 1e200):
        x = rng.uniform(-1, 1, (64, 512)) * scale
        x[5, :] = 0.0  # an all-zero row
        x[7, 3] = x[7].max() * 1024  # one dominant entry
        x[9, :] = -x[9, :].__abs__().max()  # a row sitting exactly at the (negative) maximum
        ds, e, _, _ = P.split_rows(x)
        assert all(d.min() >= -128 and d.max() <= 127 for d in ds), "digits must fit int8"
        assert np.abs(ds[0]).max() <= 65
        recon = sum(np.ldexp(d.astype(np.longdouble), -7 - 8 * t) for t, d in enumerate(ds))
        err = np.abs(np.ldexp(recon, e[:, None]) - x.astype(np.longdouble))
        bound = np.ldexp(np.ones(x.shape[0]), e - 56)[:, None]
        assert (err <= bound).all(), "x = 2^e sum_t d_t 2^(-7-8t) + r with |r| <= 2^(e-56)"
        big = np.abs(x) >= np.ldexp(1.0, e - 3)[:, None]  # within a factor 4 of the row maximum: exact
        assert (err[big] == 0).all()


def test_group_sums_fit_int32_and_fold_is_exact():
    rng = np.random.default_rng(4)
    k = 512
    # worst case: every digit at -128 (top digits at their own bound 65)
    da = [np.full((4, k), 65 if t == 0 else -128, dtype=np.int64) for t in range(S)]
    db = [np.full((4, k), -65 if t == 0 else -128, dtype=np.int64) for t in range(S)]
    G = P.group_sums(da, db)
    assert all(np.abs(g).max() < 2**26 for g in G), "int32 accumulators hold every group sum"
    for trial in range(2):
        if trial == 1:
            G = [rng.integers(-7 * 2**23, 7 * 2**23, (4, 4), dtype=np.int64) for _ in range(S)]
        hi, lo = P.fold(G)
        assert np.abs(hi).max() < 2**51 and np.abs(lo).max() < 2**51
        # magic-number conversion: bits(2^52 + 2^51 + v) - 1.5 * 2^52 == v exactly for |v| < 2^51
        for v in (hi, lo):
            bits = (np.int64(0x4330000000000000) + (v + (np.int64(1) << 51))).view(np.float64)
            assert np.array_equal(bits - 1.5 * 2.0**52, v.astype(np.float64))
            assert np.array_equal(v.astype(np.float64).astype(np.int64), v)
        exact = sum(G[g].astype(object) * (2 ** (8 * (S - 1 - g))) for g in range(S))  # python ints, scaled by 2^48
        folded = hi.astype(object) * 2**24 + lo.astype(object)
        assert (exact == folded).all()


def test_product_error_is_below_a_native_fp64_gemm():
    rng = np.random.default_rng(5)
    a, b = rng.uniform(-1, 1, (96, 512)), rng.uniform(-1, 1, (80, 512))
    ref = P.ref_gemm(a, b)
    scale = np.abs(a).astype(np.longdouble) @ np.abs(b).T.astype(np.longdouble)
    e_native = float((np.abs(a @ b.T - ref) / scale).max())
    e_oz = float((np.abs(P.ozaki_gemm(a, b) - ref) / scale).max())
    assert e_oz < e_native, (e_oz, e_native)
    assert e_oz < 2.0**-52


def test_error_model_row_maxima_bound():
    """|delta| <= K 2^-53.2 max|a_i| max|b_j| (gemm_ozaki.h), also on rows spanning 40 binades."""
    rng = np.random.default_rng(6)
    k = 512
    a = np.ldexp(rng.uniform(-1, 1, (64, k)), rng.integers(-20, 21, (64, k)))
    b = np.ldexp(rng.uniform(-1, 1, (48, k)), rng.integers(-20, 21, (48, k)))
    err = np.abs(P.ozaki_gemm(a, b) - P.ref_gemm(a, b)).astype(np.float64)
    bound = k * 2.0**-53.2 * np.abs(a).max(axis=1)[:, None] * np.abs(b).max(axis=1)[None, :]
    # plus the final rounding of the fp64 result itself
    bound = bound + np.finfo(np.float64).eps * np.abs(P.ref_gemm(a, b)).astype(np.float64)
    assert (err <= bound).all()


def test_guard_rule():
    """starved = rounded AND fewer than 16 significant bits kept (split_i8_kernel): zeros, exact tiny powers of two and
    ordinary data do not trigger; an inexact entry 2^-45 below its row maximum does."""
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, (8, 512))
    x[1, 3] = 0.0
    x[2, 5] = 2.0**-50
    _, _, lossy, M = P.split_rows(x)
    starved = lossy & (np.abs(M) < 2**15)
    assert not starved.any()
    x[3, 7] = np.ldexp(x[3, 7], -45)
    _, _, lossy, M = P.split_rows(x)
    starved = lossy & (np.abs(M) < 2**15)
    assert starved[3, 7] and starved.sum() == 1
