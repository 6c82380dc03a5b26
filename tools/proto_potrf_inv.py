import numpy as np
def potrf_inv(A):
    n = A.shape[0]
    S = np.zeros((n, n)); 
    S[np.tril_indices(n)] = A[np.tril_indices(n)]   # lower = A, upper = 0 (M = I offdiag)
    dinv = np.zeros(n)
    for j in range(n):
        d = np.sqrt(S[j, j]); S[j, j] = d; dinv[j] = 1.0 / d
        # scale whole column j except diagonal (rows<j: M[j,r]; rows>j: L)
        for r in range(n):
            if r != j: S[r, j] /= d
        # rank-1 update: columns s>j, rows r not in [j, s)
        for s in range(j + 1, n):
            for r in range(n):
                if r >= s or r < j:
                    S[r, s] -= S[r, j] * S[s, j]
            S[j, s] = -S[s, j] / d      # new M[s][j], stored transposed
    L = np.tril(S)
    M = np.triu(S, 1).T + np.diag(dinv)
    return L, M
rng = np.random.default_rng(0)
n = 24
X = rng.standard_normal((n, n)); A = X @ X.T + n * np.eye(n)
L, M = potrf_inv(A)
print(np.abs(L - np.linalg.cholesky(A)).max(), np.abs(M @ L - np.eye(n)).max())
