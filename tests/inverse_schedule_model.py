"""numpy model of the two sweeps of dla-future_b200/csrc/inverse_engine.cu on a simulated P x Q grid.

Test infrastructure: it mirrors the schedule of the CUDA engine step by step (same index arithmetic, same packed
operands, same "broadcasts", one GEMM per step) so that the algorithm — the extended column panel, the zeroed row, the
lower-masked assemble update, the block-cyclic bookkeeping — can be checked against the oracle on the CPU. The
arithmetic inside each step is plain numpy.
"""
import numpy as np


def cnt(g_end, v, grid):
    """tiles of virtual rank v with global index < g_end (tri_kernels.cuh: cnt)"""
    return (g_end - v + grid - 1) // grid if g_end > v else 0


class Rank:
    def __init__(self, er, ec, Pe, Qe, nt, nbp, dtype):
        self.er, self.ec = er, ec
        self.ltr, self.ltc = cnt(nt, er, Pe), cnt(nt, ec, Qe)
        self.slab = np.zeros((max(self.ltr, 1) * nbp, max(self.ltc, 1) * nbp), dtype=dtype)


def _load(ranks, l_full, n, nb, nbp, Pe, Qe, unit):
    """engine slab of every rank from the global LOWER matrix (inv_convert_kernel<LOAD>)"""
    nt = -(-n // nb)
    for rk in ranks.values():
        for la in range(rk.ltr):
            for lb in range(rk.ltc):
                ga, gb = la * Pe + rk.er, lb * Qe + rk.ec
                if ga < gb:
                    continue
                rows, cols = min(nb, n - ga * nb), min(nb, n - gb * nb)
                t = np.zeros((nbp, nbp), dtype=l_full.dtype)
                t[:rows, :cols] = l_full[ga * nb:ga * nb + rows, gb * nb:gb * nb + cols]
                if ga == gb:
                    t = np.tril(t)
                    for r in range(nbp):
                        if unit or r >= rows:
                            t[r, r] = 1
                rk.slab[la * nbp:(la + 1) * nbp, lb * nbp:(lb + 1) * nbp] = t
    return nt


def _store(ranks, out, n, nb, nbp, Pe, Qe, unit):
    for rk in ranks.values():
        for la in range(rk.ltr):
            for lb in range(rk.ltc):
                ga, gb = la * Pe + rk.er, lb * Qe + rk.ec
                if ga < gb:
                    continue
                rows, cols = min(nb, n - ga * nb), min(nb, n - gb * nb)
                t = rk.slab[la * nbp:la * nbp + rows, lb * nbp:lb * nbp + cols]
                for r in range(rows):
                    for c in range(cols):
                        if ga > gb or r > c or (r == c and not unit):
                            out[ga * nb + r, gb * nb + c] = t[r, c]


def run(l_full, nb, g, Pe, Qe, phases, unit=False):
    """l_full: global lower-triangular matrix (only its lower triangle is read); g = kernel granularity (padded tile
    edge = round_up(nb, g)); phases: 1 = triangular inverse, 2 = assemble, 3 = both. Returns the global result with the
    untouched entries copied from l_full."""
    n = l_full.shape[0]
    out = l_full.copy()
    if n == 0:
        return out
    dtype = l_full.dtype
    nbp = -(-nb // g) * g
    ranks = {(r, c): Rank(r, c, Pe, Qe, -(-n // nb), nbp, dtype) for r in range(Pe) for c in range(Qe)}
    nt = _load(ranks, l_full, n, nb, nbp, Pe, Qe, unit)
    H = lambda x: x.conj().T
    if phases & 1:
        # all diagonal tiles first: Wh_k = L_kk^-H
        wh = {}
        for k in range(nt):
            rk = ranks[(k % Pe, k % Qe)]
            lkk = rk.slab[(k // Pe) * nbp:(k // Pe + 1) * nbp, (k // Qe) * nbp:(k // Qe + 1) * nbp]
            wh[k] = H(np.linalg.inv(lkk))
        for k in range(nt - 1, -1, -1):
            owner_r, owner_c = k % Pe, k % Qe
            colp = {}  # per process row: the compact extended column panel (after the row broadcast)
            for er in range(Pe):
                rk = ranks[(er, owner_c)]
                li_k, li_k1 = cnt(k, er, Pe), cnt(k + 1, er, Pe)
                mk, vrows = (rk.ltr - li_k) * nbp, (rk.ltr - li_k1) * nbp
                panel = np.zeros((mk, nbp), dtype=dtype)
                lck = k // Qe
                if vrows > 0:
                    a_ik = rk.slab[li_k1 * nbp:, lck * nbp:(lck + 1) * nbp]
                    v = -a_ik @ H(wh[k])  # -A(i,k) L_kk^-1
                    panel[(li_k1 - li_k) * nbp:, :] = v
                    rk.slab[li_k1 * nbp:, lck * nbp:(lck + 1) * nbp] = v
                if er == owner_r:
                    wkk = H(wh[k])
                    panel[:nbp, :] = wkk
                    rk.slab[(k // Pe) * nbp:(k // Pe + 1) * nbp, lck * nbp:(lck + 1) * nbp] = wkk
                colp[er] = panel
            if k == 0:
                break
            panb = {}  # per process column: packed -L(k,j)^H tiles for its local columns j < k
            for ec in range(Qe):
                rk = ranks[(owner_r, ec)]
                ncols = cnt(k, ec, Qe)
                if ncols > 0:
                    row = rk.slab[(k // Pe) * nbp:(k // Pe + 1) * nbp, :ncols * nbp]
                    panb[ec] = [-H(row[:, j * nbp:(j + 1) * nbp]) for j in range(ncols)]
                    row[:, :] = 0
                else:
                    panb[ec] = []
            for (er, ec), rk in ranks.items():
                li_k = cnt(k, er, Pe)
                mk, ncols = (rk.ltr - li_k) * nbp, cnt(k, ec, Qe)
                if mk > 0 and ncols > 0:
                    b = np.vstack(panb[ec])  # (ncols * nbp) x nbp
                    rk.slab[li_k * nbp:li_k * nbp + mk, :ncols * nbp] -= colp[er] @ H(b)
    if phases & 2:
        for k in range(nt):
            owner_r = k % Pe
            panb = {}
            for ec in range(Qe):
                rk = ranks[(owner_r, ec)]
                ncols = cnt(k + 1, ec, Qe)
                if ncols > 0:
                    row = rk.slab[(k // Pe) * nbp:(k // Pe + 1) * nbp, :ncols * nbp]
                    panb[ec] = [H(row[:, j * nbp:(j + 1) * nbp]).copy() for j in range(ncols)]
                    row[:, :] = 0
                else:
                    panb[ec] = []
            for (er, ec), rk in ranks.items():
                nrows, ncols = cnt(k + 1, er, Pe), cnt(k + 1, ec, Qe)
                if nrows == 0 or ncols == 0:
                    continue
                pa = []
                for li in range(nrows):
                    i = li * Pe + er
                    pa.append(panb[i % Qe][i // Qe])  # from the rank of my row that sits in process column i % Qe
                a = np.vstack(pa)
                b = np.vstack(panb[ec])
                upd = a @ H(b)
                # lower mask on global element coordinates
                gi = np.concatenate([np.arange(nbp) + (li * Pe + er) * nbp for li in range(nrows)])[:, None]
                gj = np.concatenate([np.arange(nbp) + (lj * Qe + ec) * nbp for lj in range(ncols)])[None, :]
                rk.slab[:nrows * nbp, :ncols * nbp] += np.where(gi >= gj, upd, 0)
    _store(ranks, out, n, nb, nbp, Pe, Qe, unit)
    return out


def run_user(a_full, uplo, diag, nb, g, P, Q, phases):
    """The user-facing problem: uplo 'U' runs the lower engine on A^H with the grid roles swapped."""
    unit = diag.upper() == "U"
    if uplo.upper() == "L":
        return run(a_full, nb, g, P, Q, phases, unit)
    res = run(np.asfortranarray(a_full.conj().T), nb, g, Q, P, phases, unit)
    return np.asfortranarray(res.conj().T)
