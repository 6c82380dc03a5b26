import numpy as np
def potrf_inv_c(A):
    n = A.shape[0]
    S = np.zeros((n, n), dtype=complex)
    S[np.tril_indices(n)] = A[np.tril_indices(n)]
    dinv = np.zeros(n)
    dd = np.zeros(n)
    for j in range(n):
        ajj = S[j, j].real
        d = np.sqrt(ajj); dd[j] = d; dinv[j] = 1.0 / d
        for r in range(n):
            if r != j: S[r, j] *= dinv[j]
        colj = S[:, j].copy()
        for s in range(j + 1, n):
            for r in range(n):
                if r >= s or r < j:
                    S[r, s] -= colj[r] * np.conj(colj[s])
            S[j, s] = -np.conj(colj[s]) * dinv[j]
    L = np.tril(S, -1) + np.diag(dd)
    M = np.conj(np.triu(S, 1)).T + np.diag(dinv)
    return L, M
rng = np.random.default_rng(0)
n = 20
X = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)); A = X @ X.conj().T + n * np.eye(n)
L, M = potrf_inv_c(A)
print(np.abs(L - np.linalg.cholesky(A)).max(), np.abs(M @ L - np.eye(n)).max())
