"""TEST INFRASTRUCTURE: numpy restatement of the int8-digit (Ozaki-scheme) arithmetic of
dla-future_b200/csrc/gemm_ozaki_i8.cu — split_i8_kernel (balanced radix-256 digits of a 55-bit fixed-point row
mantissa), the 28 exact digit-plane products grouped by g = t + u <= 6, and the epilogue's two-integer fold."""
import numpy as np

S = 7            # digit planes
MANT_BITS = 55   # x * 2^-e is rounded to a multiple of 2^-55 (|x| 2^-e < 1/2  ->  |M| <= 2^54)


def row_exponents(x):
    m = np.abs(x).max(axis=1)
    e = np.zeros(x.shape[0], dtype=np.int64)
    nz = m > 0
    e[nz] = np.frexp(m[nz])[1] + 1  # frexp: m = f 2^p, f in [0.5, 1) -> ilogb(m) = p - 1; e = ilogb + 2
    return e


def split_rows(x):
    """-> (digits[S] int64 arrays in [-128, 127] (top digit in [-65, 65]), e): x = 2^e sum_t d_t 2^(-7-8t) + r,
    |r| <= 2^(e-56)."""
    e = row_exponents(x)
    s = np.ldexp(x, (MANT_BITS - e)[:, None])
    M = np.rint(s).astype(np.int64)
    lossy = (M.astype(np.float64) != s)
    digits = [None] * S
    for t in range(S - 1, 0, -1):
        d = ((M + 128) & 255) - 128
        digits[t] = d
        M = (M - d) >> 8
    digits[0] = M
    return digits, e, lossy, np.rint(s).astype(np.int64)


def group_sums(da, db):
    """G_g = sum_{t+u=g} A_t B_u^T, exact in int64 (the tensor core accumulates the same integers in int32)."""
    return [sum(da[t] @ db[g - t].T for t in range(g + 1)) for g in range(S)]


def fold(G):
    """hi, lo of the epilogue: sum_g G_g 2^(-8g) = (hi + lo 2^-24) 2^-24."""
    hi = (G[0] << 24) + (G[1] << 16) + (G[2] << 8) + G[3]
    lo = (G[4] << 16) + (G[5] << 8) + G[6]
    return hi, lo


def ozaki_gemm(a, b):
    """a (m x k), b (n x k) fp64 -> a b^T through the int8 scheme (fp64 result)."""
    da, ea, _, _ = split_rows(a)
    db, eb, _, _ = split_rows(b)
    hi, lo = fold(group_sums(da, db))
    v = hi.astype(np.float64) + lo.astype(np.float64) * 2.0**-24  # fma(lo, 2^-24, hi): one rounding
    return np.ldexp(v, (ea[:, None] + eb[None, :]) - 14 - 24)


def ref_gemm(a, b):
    return a.astype(np.longdouble) @ b.T.astype(np.longdouble)
