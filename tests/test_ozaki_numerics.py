"""CPU checks of the arithmetic the int8 (Ozaki-scheme) fp64 update relies on — the same formulas as
dla-future_b200/csrc/gemm_ozaki_i8.cu (split_i8_kernel, ozaki_fold_store), restated in numpy (tools/proto_ozaki_i8.py):
digit range, error-free slicing, int32 headroom of the group sums, exactness of the two-integer fold and of the
magic-number int64 -> fp64 conversion, and the end-to-end product error next to a native fp64 GEMM."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import proto_ozaki_i8 as P  # noqa: E402

S = 8


def test_digits_are_int8_and_slicing_is_error_free():
    rng = np.random.default_rng(3)
    for scale in (1.0, 1e-200, 1e200):
        x = rng.uniform(-1, 1, (64, 512)) * scale
        x[5, :] = 0.0  # an all-zero row
        x[7, 3] = x[7].max() * 1024  # one dominant entry
        qs, e = P.split_rows(x, S)
        assert all(np.abs(q).max() <= 64 for q in qs), "digits must fit a signed 7-bit range"
        assert all((q == np.rint(q)).all() for q in qs)
        recon = sum(np.ldexp(q.astype(np.longdouble), -7 * (t + 1)) for t, q in enumerate(qs))
        err = np.abs(np.ldexp(recon, e[:, None]) - x.astype(np.longdouble))
        bound = np.ldexp(np.ones(x.shape[0]), e - 57)[:, None]
        assert (err <= bound).all(), "x = 2^e sum_t q_t 128^-(t+1) + r with |r| <= 2^(e-57)"


def test_group_sums_fit_int32_and_fold_is_exact():
    rng = np.random.default_rng(4)
    k = 512
    # worst case digits +-64 everywhere
    qa = [np.full((4, k), 64.0) for _ in range(S)]
    qb = [np.full((4, k), -64.0) for _ in range(S)]
    acc = []
    for g in range(S):
        a = sum(qa[t] @ qb[g - t].T for t in range(g + 1))
        assert np.abs(a).max() <= 2**24 < 2**31
        acc.append(a.astype(np.int64))
    # the epilogue's fold: hi = a0 2^21 + a1 2^14 + a2 2^7 + a3, lo likewise from a4..a7; value = (hi + lo 2^-28) 2^-21
    for trial in range(2):
        if trial == 1:
            acc = [rng.integers(-2**24, 2**24, (4, 4), dtype=np.int64) for _ in range(S)]
        hi = (acc[0] << 21) + (acc[1] << 14) + (acc[2] << 7) + acc[3]
        lo = (acc[4] << 21) + (acc[5] << 14) + (acc[6] << 7) + acc[7]
        assert np.abs(hi).max() < 2**51 and np.abs(lo).max() < 2**51
        # magic-number conversion: bits(2^52 + 2^51 + v) - 1.5 * 2^52 == v exactly for |v| < 2^51
        for v in (hi, lo):
            bits = (np.int64(0x4330000000000000) + (v + (np.int64(1) << 51))).view(np.float64)
            assert np.array_equal(bits - 1.5 * 2.0**52, v.astype(np.float64))
            assert np.array_equal(v.astype(np.float64).astype(np.int64), v), "46-bit integers are exact in fp64"
        exact = sum(acc[g].astype(object) * (2 ** (7 * (S - 1 - g))) for g in range(S))  # python ints, scaled by 2^49
        folded = hi.astype(object) * 2**28 + lo.astype(object)
        assert (exact == folded).all()


def test_product_error_is_below_a_native_fp64_gemm():
    rng = np.random.default_rng(5)
    a, b = rng.uniform(-1, 1, (96, 512)), rng.uniform(-1, 1, (80, 512))
    ref = P.ref_gemm(a, b)
    scale = np.abs(a).astype(np.longdouble) @ np.abs(b).T.astype(np.longdouble)
    e_native = float((np.abs(a @ b.T - ref) / scale).max())
    e_oz = float((np.abs(P.ozaki_gemm(a, b, S) - ref) / scale).max())
    assert e_oz < e_native, (e_oz, e_native)
    assert e_oz < 2.0**-52
    # 7 digits are NOT enough for that claim (why S = 8)
    assert float((np.abs(P.ozaki_gemm(a, b, 7) - ref) / scale).max()) > e_oz
