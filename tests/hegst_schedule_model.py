"""numpy model of the two sweeps of dla-future_b200/csrc/hegst_engine.cu (A <- inv(L) A inv(L)^H) on a simulated P x Q grid.

Test infrastructure: mirrors the schedule of the CUDA engine step by step (same index arithmetic, same operands and
"broadcasts", one or two GEMMs per step) so that the algorithm — the deferred panel solve as ONE strictly-lower triangular
sweep, the two lower-masked rank-nb updates, the block-cyclic bookkeeping — can be checked against the oracle on the CPU.
"""
import numpy as np

from inverse_schedule_model import Rank, cnt


def _load(ranks, full, n, nb, nbp, Pe, Qe, pad_identity, key):
    for rk in ranks.values():
        slab = np.zeros_like(rk.slab)
        for la in range(rk.ltr):
            for lb in range(rk.ltc):
                ga, gb = la * Pe + rk.er, lb * Qe + rk.ec
                if ga < gb:
                    continue
                rows, cols = min(nb, n - ga * nb), min(nb, n - gb * nb)
                t = np.zeros((nbp, nbp), dtype=full.dtype)
                t[:rows, :cols] = full[ga * nb:ga * nb + rows, gb * nb:gb * nb + cols]
                if ga == gb:
                    t = np.tril(t)
                    if pad_identity:
                        for r in range(rows, nbp):
                            t[r, r] = 1
                slab[la * nbp:(la + 1) * nbp, lb * nbp:(lb + 1) * nbp] = t
        setattr(rk, key, slab)


def run(a_full, l_full, nb, g, Pe, Qe):
    """a_full: Hermitian matrix given by its LOWER triangle, l_full: lower Cholesky factor of B. Returns the global result
    (lower triangle overwritten, the rest copied from a_full)."""
    n = a_full.shape[0]
    out = a_full.copy()
    if n == 0:
        return out
    dtype = a_full.dtype
    nbp = -(-nb // g) * g
    nt = -(-n // nb)
    ranks = {(r, c): Rank(r, c, Pe, Qe, nt, nbp, dtype) for r in range(Pe) for c in range(Qe)}
    _load(ranks, a_full, n, nb, nbp, Pe, Qe, False, "A")
    _load(ranks, l_full, n, nb, nbp, Pe, Qe, True, "L")
    H = lambda x: x.conj().T

    def herm(t):  # full Hermitian tile from its lower triangle, real diagonal
        full = np.tril(t) + H(np.tril(t, -1))
        full[np.arange(nbp), np.arange(nbp)] = full[np.arange(nbp), np.arange(nbp)].real
        return full

    # ---- phase 1: right-looking two-sided update
    for k in range(nt):
        owner = ranks[(k % Pe, k % Qe)]
        lr, lc = k // Pe, k // Qe
        lkk = owner.L[lr * nbp:(lr + 1) * nbp, lc * nbp:(lc + 1) * nbp]
        linv_h = H(np.linalg.inv(lkk))
        hk = herm(owner.A[lr * nbp:(lr + 1) * nbp, lc * nbp:(lc + 1) * nbp])
        t1 = hk @ linv_h
        t2 = H(t1) @ linv_h  # = (inv(L) H inv(L)^H)^H
        owner.A[lr * nbp:(lr + 1) * nbp, lc * nbp:(lc + 1) * nbp] = np.tril(t2)
        hkk = herm(t2)
        if k == nt - 1:
            break
        # column panels (rows > k) of A and L on the ranks of process column k % Qe
        panP, panL = {}, {}
        for er in range(Pe):
            rk = ranks[(er, k % Qe)]
            li1 = cnt(k + 1, er, Pe)
            p = rk.A[li1 * nbp:, lc * nbp:(lc + 1) * nbp]
            lp = rk.L[li1 * nbp:, lc * nbp:(lc + 1) * nbp]
            p[:, :] = p @ linv_h
            p[:, :] -= 0.5 * lp @ hkk
            panP[er], panL[er] = p.copy(), lp.copy()  # what the row broadcast delivers
        # transposed panels: tile j of both panels to process column j % Qe (from the rank holding the diagonal tile (j, j))
        for (er, ec), rk in ranks.items():
            li1, lj1 = cnt(k + 1, er, Pe), cnt(k + 1, ec, Qe)
            nrows, ncols = rk.ltr - li1, rk.ltc - lj1
            if nrows <= 0 or ncols <= 0:
                continue
            pt, lt = [], []
            for lj in range(lj1, rk.ltc):
                j = lj * Qe + ec
                src_er = j % Pe
                off = (j // Pe - cnt(k + 1, src_er, Pe)) * nbp
                pt.append(panP[src_er][off:off + nbp])
                lt.append(panL[src_er][off:off + nbp])
            pt, lt = np.vstack(pt), np.vstack(lt)
            upd = panP[er] @ H(lt) + panL[er] @ H(pt)
            gi = np.concatenate([np.arange(nbp) + (li * Pe + er) * nbp for li in range(li1, rk.ltr)])[:, None]
            gj = np.concatenate([np.arange(nbp) + (lj * Qe + ec) * nbp for lj in range(lj1, rk.ltc)])[None, :]
            blk = rk.A[li1 * nbp:, lj1 * nbp:]
            blk -= np.where(gi >= gj, upd, 0)
            # the diagonal of a Hermitian update stays real
            dmask = (gi == gj)
            blk[dmask] = blk[dmask].real
        # second half of the hemm
        for er in range(Pe):
            rk = ranks[(er, k % Qe)]
            li1 = cnt(k + 1, er, Pe)
            rk.A[li1 * nbp:, lc * nbp:(lc + 1) * nbp] -= 0.5 * rk.L[li1 * nbp:, lc * nbp:(lc + 1) * nbp] @ hkk
    # ---- phase 2: the deferred panel solves = one strictly-lower sweep  X(j, :j) = inv(L_jj) C(j, :j);  C(t > j, :j) -= L(t, j) X(j, :j)
    for j in range(1, nt):
        owner = ranks[(j % Pe, j % Qe)]
        ljj = owner.L[(j // Pe) * nbp:(j // Pe + 1) * nbp, (j // Qe) * nbp:(j // Qe + 1) * nbp]
        linv_h = H(np.linalg.inv(ljj))
        rpan = {}
        for ec in range(Qe):
            rk = ranks[(j % Pe, ec)]
            ncols = cnt(j, ec, Qe)
            if ncols == 0:
                rpan[ec] = None
                continue
            row = rk.A[(j // Pe) * nbp:(j // Pe + 1) * nbp, :ncols * nbp]
            r = H(row) @ linv_h  # (ncols * nbp) x nbp, plain panel
            row[:, :] = H(r)
            rpan[ec] = r
        lpan = {}
        for er in range(Pe):
            rk = ranks[(er, j % Qe)]
            lj1 = cnt(j + 1, er, Pe)
            lpan[er] = rk.L[lj1 * nbp:, (j // Qe) * nbp:(j // Qe + 1) * nbp].copy()
        for (er, ec), rk in ranks.items():
            lj1, ncols = cnt(j + 1, er, Pe), cnt(j, ec, Qe)
            if rk.ltr - lj1 <= 0 or ncols == 0:
                continue
            rk.A[lj1 * nbp:, :ncols * nbp] -= lpan[er] @ H(rpan[ec])
    for rk in ranks.values():
        for la in range(rk.ltr):
            for lb in range(rk.ltc):
                ga, gb = la * Pe + rk.er, lb * Qe + rk.ec
                if ga < gb:
                    continue
                rows, cols = min(nb, n - ga * nb), min(nb, n - gb * nb)
                t = rk.A[la * nbp:la * nbp + rows, lb * nbp:lb * nbp + cols]
                for r in range(rows):
                    for c in range(cols):
                        if ga > gb or r >= c:
                            out[ga * nb + r, gb * nb + c] = t[r, c]
    return out


def run_user(a_full, l_full, uplo, nb, g, P, Q):
    if uplo.upper() == "L":
        return run(a_full, l_full, nb, g, P, Q)
    res = run(np.asfortranarray(a_full.conj().T), np.asfortranarray(l_full.conj().T), nb, g, Q, P)
    return np.asfortranarray(res.conj().T)
