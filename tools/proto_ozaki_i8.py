"""Numerics prototype (CPU, numpy) of the int8 Ozaki-scheme trailing update planned for the fp64 bulk GEMM:
C -= A B^T with A, B split row-wise into S signed 7-bit slices (int8 in [-64, 64]), exact int32 slice products
(what tcgen05.mma kind::i8 computes), anti-diagonal groups g = t + u < S recombined in fp64.
Prints the GEMM error next to a plain fp64 GEMM (both against a long-double-ish reference) and the residual of a
right-looking blocked Cholesky whose bulk updates use the emulation."""
import sys
import numpy as np


def split_rows(x, S):
    """x (m x k) -> (q[S] int8-valued float arrays, e[m]) with x = 2^e * sum_t q_t 128^-(t+1) + O(2^-57)."""
    amax = np.abs(x).max(axis=1)
    e = np.where(amax > 0, np.floor(np.log2(np.where(amax > 0, amax, 1.0))) + 2, 0).astype(np.int64)  # |x| 2^-e <= 0.5
    r = np.ldexp(x, -e[:, None])
    qs = []
    for _ in range(S):
        r = r * 128.0
        q = np.rint(r)
        r = r - q
        qs.append(q)
    return qs, e


def ozaki_gemm(a, b, S):
    qa, ea = split_rows(a, S)
    qb, eb = split_rows(b, S)
    m, n = a.shape[0], b.shape[0]
    out = np.zeros((m, n))
    for g in range(S - 1, -1, -1):  # small terms first
        acc = np.zeros((m, n))
        for t in range(g + 1):
            acc += qa[t] @ qb[g - t].T  # exact: |entries| < 2^24
        assert np.abs(acc).max() < 2**31
        out += np.ldexp(acc, -7 * (g + 2))
    return np.ldexp(out, ea[:, None] + eb[None, :])


def ref_gemm(a, b):
    return (a.astype(np.longdouble) @ b.T.astype(np.longdouble))


def main():
    rng = np.random.default_rng(1)
    m, k = 384, 512
    for name, a, b in [
        ("uniform(-1,1)", rng.uniform(-1, 1, (m, k)), rng.uniform(-1, 1, (m, k))),
        ("wide dynamic range rows", rng.uniform(-1, 1, (m, k)) * 10.0 ** rng.integers(-8, 8, (m, k)),
         rng.uniform(-1, 1, (m, k))),
    ]:
        ref = ref_gemm(a, b)
        scale = np.abs(a).astype(np.longdouble) @ np.abs(b).T.astype(np.longdouble)
        e64 = np.abs(a @ b.T - ref)
        for S in (6, 7, 8, 9):
            eo = np.abs(ozaki_gemm(a, b, S) - ref)
            print(f"{name:26s} S={S}: max err/(|a||b|) ozaki {float((eo / scale).max()):.2e}   fp64 gemm {float((e64 / scale).max()):.2e}")

    # blocked Cholesky with emulated bulk updates
    n, nb = 2048, 256
    a0 = rng.uniform(-1, 1, (n, n))
    a0 = np.tril(a0) + np.tril(a0, -1).T
    a0[np.diag_indices(n)] = rng.uniform(-1, 1, n) + 2 * n
    for S in (0, 7, 8):
        a = a0.copy()
        nt = n // nb
        for kk in range(nt):
            s = slice(kk * nb, (kk + 1) * nb)
            a[s, s] = np.linalg.cholesky(a[s, s])
            if kk + 1 < nt:
                r = slice((kk + 1) * nb, n)
                import scipy.linalg as sl
                a[r, s] = sl.solve_triangular(a[s, s], a[r, s].T, lower=True).T
                p = a[r, s]
                upd = (p @ p.T) if S == 0 else ozaki_gemm(p, p, S)
                a[r, r] -= upd
        L = np.tril(a)
        res = np.abs(np.tril(L @ L.T - a0)).max() / np.abs(a0).max()
        print(f"blocked Cholesky N={n} nb={nb} bulk={'fp64' if S == 0 else f'ozaki S={S}'}: max|A-LL^T|/max|A| = {res:.3e} (gate eps*n = {n * 2.2e-16:.2e})")


if __name__ == "__main__":
    main()
