"""Worker for the multi-process tests (one process per rank).

  mode cpu : gloo, no GPU — host logic of the N>1 path: grid coordinates, local sizes and the
             distribution-independent input generator, assembled across ranks and compared with the oracle.
  mode gpu : nccl, one GPU per rank — distributed POTRF through the C ABI on a P x Q grid (with a non-zero
             source rank like test/unit/factorization/test_cholesky.cpp:85) against the oracle's factor.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def coords(rank, P, Q, order):
    return (rank % P, rank // P) if order == "C" else (rank // Q, rank % Q)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", required=True, choices=["cpu", "gpu"])
    ap.add_argument("--grid", default="2x1")
    ap.add_argument("--order", default="R")
    ap.add_argument("--big", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    P, Q = (int(x) for x in a.grid.split("x"))
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    pkg = ge.load_package()
    O = ge.load_oracle()
    if a.mode == "cpu":
        dist.init_process_group("gloo")
        comm = pkg.comm_create_local(rank, world)
    else:
        lr = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        pkg.initialize()
        comm = pkg.comm_create_from_torch()
    ctx = pkg.create_grid(comm, P, Q, a.order)
    gP, gQ, myrow, mycol = pkg.grid_info(ctx)
    assert (gP, gQ) == (P, Q)
    if rank < P * Q:
        assert (myrow, mycol) == coords(rank, P, Q, a.order), (rank, myrow, mycol)

    src = (max(0, P - 1), min(1, Q - 1))  # test_cholesky.cpp:85
    failures = []
    if a.mode == "cpu":
        cases = [(37, 8, "d"), (64, 16, "z"), (21, 5, "s")]
        for n, nb, t in cases:
            dt = pkg.TYPES[t]
            for s in [(0, 0), src]:
                if rank >= P * Q:
                    continue
                d = pkg.descriptor(n, nb, 1, s[0], s[1])
                lrows, lcols = pkg.local_shape(ctx, d)
                assert lrows == O.local_size(n, nb, P, myrow, s[0]) and lcols == O.local_size(n, nb, Q, mycol, s[1])
                loc = np.zeros((lrows, lcols), dtype=dt, order="F")
                if lrows and lcols:
                    pkg.set_random_hermitian_positive_definite(ctx, loc, n, nb, s[0], s[1])
                ref = O.scatter_block_cyclic(O.set_random_hermitian_positive_definite(n, nb, dt), nb, (P, Q), s)
                if not np.array_equal(loc, ref[(myrow, mycol)]):
                    failures.append(("generator", n, nb, t, s))
        # the 128-byte bootstrap payload travels over torch.distributed unchanged
        obj = [os.urandom(128) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        gathered = [None] * world
        dist.all_gather_object(gathered, obj[0])
        assert all(g == gathered[0] and len(g) == 128 for g in gathered)
    else:
        cases = [(0, 2), (5, 8), (34, 34), (4, 3), (16, 10), (34, 13), (32, 5)]  # test_cholesky.cpp:54-58
        sized = [(m, mb, t) for (m, mb) in cases for t in "sdcz"]
        sized += [(1000, 128, "d"), (1536, 256, "d"), (700, 64, "z"), (1100, 100, "s"), (640, 64, "c")]
        if a.big:
            sized += [(8192, 512, "d"), (6144, 512, "z")]
        for n, nb, t in sized:
            dt = pkg.TYPES[t]
            for uplo in "LU":
                for s in ([(0, 0), src] if n < 2000 else [(0, 0)]):
                    if n <= 64:
                        A, expect = O.cholesky_setters(uplo, n, dt)
                    else:
                        A = O.set_random_hermitian_positive_definite(n, nb, dt)
                        expect = A.copy(order="F")
                        assert O.cholesky_local(uplo, expect, nb, 8) == 0
                    if rank < P * Q:
                        loc = np.asfortranarray(O.scatter_block_cyclic(A, nb, (P, Q), s)[(myrow, mycol)])
                        exp_loc = O.scatter_block_cyclic(expect, nb, (P, Q), s)[(myrow, mycol)]
                        orig = loc.copy(order="F")
                    else:
                        loc = np.zeros((1, 1), dtype=dt, order="F")
                    info = pkg.cholesky_factorization(ctx, uplo, loc, nb, n=n, isrc=s[0], jsrc=s[1])
                    if rank >= P * Q:
                        continue
                    tol = O.cholesky_tolerance(n, dt)
                    if n <= 64:
                        ok, _, msg = O.check_near(exp_loc, loc, tol, tol)
                    else:
                        # compare only the referenced triangle (global indices), the rest must be untouched
                        gi = np.concatenate([np.arange(g * nb, min(n, (g + 1) * nb)) for g in range(-(-n // nb)) if O.rank_global_tile(g, P, s[0]) == myrow] or [np.zeros(0, int)])
                        gj = np.concatenate([np.arange(g * nb, min(n, (g + 1) * nb)) for g in range(-(-n // nb)) if O.rank_global_tile(g, Q, s[1]) == mycol] or [np.zeros(0, int)])
                        mask = (gi[:, None] >= gj[None, :]) if uplo == "L" else (gi[:, None] <= gj[None, :])
                        ok, _, msg = O.check_near(np.where(mask, exp_loc, 0), np.where(mask, loc, 0), tol, tol)
                        if ok and not np.array_equal(np.where(mask, 0, loc), np.where(mask, 0, orig)):
                            ok, msg = False, "unreferenced triangle modified"
                    if info != 0 or not ok:
                        failures.append((n, nb, t, uplo, s, info, msg))
                    if n > 64:
                        # the miniapp's result check on the grid (collective; miniapp_cholesky.cpp:408-446): same value on
                        # every rank, below the eps*n gate, and it must see a corrupted factor
                        res = pkg.check_cholesky(ctx, uplo, orig, loc, nb, n=n, isrc=s[0], jsrc=s[1])
                        gate = O.residual_gate(dt, n)[0]
                        if not (0 <= res <= gate):
                            failures.append(("grid residual", n, nb, t, uplo, s, res, gate))
                        if (n, nb, t) == (1000, 128, "d"):
                            bad = loc.copy(order="F")
                            if (myrow, mycol) == tuple(s):  # the source rank holds global element (0, 0), on the diagonal
                                bad[0, 0] += 0.5
                            res_bad = pkg.check_cholesky(ctx, uplo, orig, bad, nb, n=n, isrc=s[0], jsrc=s[1])
                            if not res_bad > 1e-6:
                                failures.append(("grid residual blind", res_bad))
        # non-SPD: every rank must report the same LAPACK info
        n, nb = 300, 64
        B = np.eye(n, order="F") * 4.0
        B[200, 200] = -1.0
        if rank < P * Q:
            loc = np.asfortranarray(O.scatter_block_cyclic(B, nb, (P, Q), (0, 0))[(myrow, mycol)])
        else:
            loc = np.zeros((1, 1), order="F")
        info = pkg.cholesky_factorization(ctx, "L", loc, nb, n=n)
        if rank < P * Q and info != 201:
            failures.append(("info", info))
        # dense indefinite matrix: after the first failure the trailing updates are poisoned and LATER diagonal tiles
        # (on other ranks) fail too -> the grid must report the FIRST failing leading minor, like LAPACK on one process
        n, nb = 700, 64
        D = O.set_random_hermitian_positive_definite(n, nb, np.float64)
        D[325, 325] = -3.0
        want = O.cholesky_local("L", D.copy(order="F"), nb, 4)
        if rank < P * Q:
            loc = np.asfortranarray(O.scatter_block_cyclic(D, nb, (P, Q), (0, 0))[(myrow, mycol)])
        else:
            loc = np.zeros((1, 1), order="F")
        info = pkg.cholesky_factorization(ctx, "L", loc, nb, n=n)
        if rank < P * Q and (want != 326 or info != want):
            failures.append(("info dense indefinite", info, want))
    if a.mode == "gpu":
        # ---- distributed triangular solver (test/unit/solver/test_triangular.cpp:106-140, :197-213): the reference's
        # closed-form systems on the grid with a non-zero source rank, every side / uplo / op / diag combination
        import itertools

        tcases = [(10, 10, 2, 3), (15, 7, 3, 5), (19, 25, 6, 5), (7, 8, 2, 9)]
        big = [(600, 400, 128, 64, "d"), (384, 512, 64, 128, "z"), (300, 260, 100, 50, "s")]
        combos = list(itertools.product("LR", "LU", "NTC", "NU"))
        for (m, n, mb, nb) in tcases:
            for t in "sdcz":
                dt = pkg.TYPES[t]
                alpha = O.TRIANGULAR_TEST_ALPHA if np.dtype(dt).kind == "c" else O.TRIANGULAR_TEST_ALPHA.real
                for side, uplo, op, diag in combos:
                    A, B, X = O.triangular_system(side, uplo, op, diag, alpha, m, n, dt)
                    ba = mb if side == "L" else nb
                    if rank < P * Q:
                        la = np.asfortranarray(O.scatter_block_cyclic(A, ba, (P, Q), src)[(myrow, mycol)])
                        lb = np.asfortranarray(O.scatter_block_cyclic_rect(B, mb, nb, (P, Q), src)[(myrow, mycol)])
                        lx = O.scatter_block_cyclic_rect(X, mb, nb, (P, Q), src)[(myrow, mycol)]
                    else:
                        la = lb = lx = np.zeros((1, 1), dtype=dt, order="F")
                    if la.size == 0:
                        la = np.zeros((max(1, la.shape[0]), max(1, la.shape[1])), dtype=dt, order="F")
                    lbw = lb if lb.size else np.zeros((max(1, lb.shape[0]), max(1, lb.shape[1])), dtype=dt, order="F")
                    pkg.triangular_solver(ctx, side, uplo, op, diag, alpha, la, lbw, mb, nb, m=m, n=n, isrc=src[0], jsrc=src[1])
                    if rank < P * Q and lb.size:
                        tol = O.triangular_tolerance(m, dt, distributed=True)
                        ok, _, msg = O.check_near(lx, lbw, tol, tol)
                        if not ok:
                            failures.append(("trsm", t, side, uplo, op, diag, m, n, mb, nb, msg))
        for (m, n, mb, nb, t) in big:
            dt = pkg.TYPES[t]
            rng = np.random.default_rng(17)
            for side, uplo, op in [("L", "L", "N"), ("L", "L", "C"), ("R", "U", "N"), ("R", "L", "T"), ("L", "U", "T")]:
                na, ba = (m, mb) if side == "L" else (n, nb)
                spd = O.set_random_hermitian_positive_definite(na, ba, dt)
                assert O.cholesky_local(uplo, spd, ba, 4) == 0
                A = np.asfortranarray(np.tril(spd) if uplo == "L" else np.triu(spd))
                B = rng.uniform(-1, 1, (m, n)).astype(dt)
                B = np.asfortranarray(B)
                ref = B.copy(order="F")
                O.triangular_solver(side, uplo, op, "N", 1.0, A, ref, mb, nb)
                if rank < P * Q:
                    la = np.asfortranarray(O.scatter_block_cyclic(A, ba, (P, Q), src)[(myrow, mycol)])
                    lb = np.asfortranarray(O.scatter_block_cyclic_rect(B, mb, nb, (P, Q), src)[(myrow, mycol)])
                    lx = O.scatter_block_cyclic_rect(ref, mb, nb, (P, Q), src)[(myrow, mycol)]
                else:
                    la = lb = lx = np.zeros((1, 1), dtype=dt, order="F")
                if la.size == 0:
                    la = np.zeros((max(1, la.shape[0]), max(1, la.shape[1])), dtype=dt, order="F")
                lbw = lb if lb.size else np.zeros((max(1, lb.shape[0]), max(1, lb.shape[1])), dtype=dt, order="F")
                pkg.triangular_solver(ctx, side, uplo, op, "N", 1.0, la, lbw, mb, nb, m=m, n=n, isrc=src[0], jsrc=src[1])
                if rank < P * Q and lb.size:
                    tol = O.triangular_tolerance(max(m, n), dt, distributed=True) * max(1.0, float(np.abs(ref).max()))
                    ok, _, msg = O.check_near(lx, lbw, tol, tol)
                    if not ok:
                        failures.append(("trsm big", t, side, uplo, op, m, n, mb, nb, msg))
    if a.mode == "gpu":
        # ---- distributed inverse (test/unit/inverse/test_triangular_inverse.cpp:78-99, test_inverse_from_cholesky_factor.cpp:
        # 77-98): the reference's closed forms on the grid with a non-zero source rank, then a random Cholesky factor vs the oracle
        def local_of(full, nb_, dt_):
            if rank < P * Q:
                loc = np.asfortranarray(O.scatter_block_cyclic(full, nb_, (P, Q), src)[(myrow, mycol)])
            else:
                loc = np.zeros((1, 1), dtype=dt_, order="F")
            work = loc if loc.size else np.zeros((max(1, loc.shape[0]), max(1, loc.shape[1])), dtype=dt_, order="F")
            return loc, work

        for (m, mb) in [(16, 10), (34, 13), (32, 5), (4, 3)]:
            for t in "sdcz":
                dt = pkg.TYPES[t]
                tol = O.inverse_tolerance(m, dt)
                for uplo in "LU":
                    for diag in "UN":
                        A, R = O.triangular_inverse_setters(uplo, diag, m, dt)
                        loc, work = local_of(A, mb, dt)
                        pkg.triangular_inverse(ctx, uplo, diag, work, mb, n=m, isrc=src[0], jsrc=src[1])
                        if rank < P * Q and loc.size:
                            ok, _, msg = O.check_near(O.scatter_block_cyclic(R, mb, (P, Q), src)[(myrow, mycol)], work, tol, tol)
                            if not ok:
                                failures.append(("trtri", t, uplo, diag, m, mb, msg))
                    T0, R = O.inverse_cholesky_factor_setters(uplo, m, dt)
                    loc, work = local_of(T0, mb, dt)
                    pkg.inverse_from_cholesky_factor(ctx, uplo, work, mb, n=m, isrc=src[0], jsrc=src[1])
                    if rank < P * Q and loc.size:
                        ok, _, msg = O.check_near(O.scatter_block_cyclic(R, mb, (P, Q), src)[(myrow, mycol)], work, tol, tol)
                        if not ok:
                            failures.append(("potri", t, uplo, m, mb, msg))
                    T0, R = O.assemble_cholesky_inverse_setters(uplo, m, dt)
                    loc, work = local_of(T0, mb, dt)
                    pkg.assemble_cholesky_inverse(ctx, uplo, work, mb, n=m, isrc=src[0], jsrc=src[1])
                    if rank < P * Q and loc.size:
                        ok, _, msg = O.check_near(O.scatter_block_cyclic(R, mb, (P, Q), src)[(myrow, mycol)], work, tol, tol)
                        if not ok:
                            failures.append(("lauum", t, uplo, m, mb, msg))
        for (m, mb, t) in [(1536, 256, "d"), (1100, 200, "d"), (768, 128, "z"), (1024, 256, "s")]:
            dt = pkg.TYPES[t]
            for uplo in "LU":
                spd = O.set_random_hermitian_positive_definite(m, mb, dt)
                fac = spd.copy(order="F")
                assert O.cholesky_local(uplo, fac, mb, 4) == 0
                tri = np.tril if uplo == "L" else np.triu
                sent = np.full((m, m), -9.9)
                fac = np.asfortranarray(tri(fac) + (np.triu(sent, 1) if uplo == "L" else np.tril(sent, -1)).astype(dt))
                ref = fac.copy(order="F")
                O.inverse_from_cholesky_factor(uplo, ref, mb)
                loc, work = local_of(fac, mb, dt)
                pkg.ppotri(ctx, uplo, work, mb, n=m, isrc=src[0], jsrc=src[1])
                if rank < P * Q and loc.size:
                    scale = float(np.abs(tri(ref)).max())
                    tol = O.inverse_tolerance(m, dt)
                    lref = O.scatter_block_cyclic(ref, mb, (P, Q), src)[(myrow, mycol)]
                    # sentinels are compared unscaled (they must be bit-identical), the inverse relative to its largest entry
                    lsent = O.scatter_block_cyclic(np.asfortranarray(~np.isclose(ref.real, -9.9)), mb, (P, Q), src)[(myrow, mycol)]
                    ok, _, msg = O.check_near(np.where(lsent, lref / scale, lref), np.where(lsent, work / scale, work), tol, tol)
                    if not ok:
                        failures.append(("potri big", t, uplo, m, mb, msg))
    if a.mode == "gpu":
        # ---- distributed generalized -> standard (test/unit/eigensolver/test_gen_to_std.cpp:81-110): closed form on the grid with
        # a non-zero source rank (the factor must come back unchanged), then a random pencil vs the oracle
        for (m, mb) in [(16, 10), (34, 13), (32, 5), (4, 3)]:
            for t in "sdcz":
                dt = pkg.TYPES[t]
                for uplo in "LU":
                    T0, A0, B0 = O.gen_to_std_setters(uplo, m, dt)
                    loc_a, work_a = local_of(A0, mb, dt)
                    loc_t, work_t = local_of(T0, mb, dt)
                    t_before = work_t.copy(order="F")
                    pkg.generalized_to_standard(ctx, uplo, work_a, work_t, mb, n=m, isrc=src[0], jsrc=src[1])
                    if rank < P * Q and loc_a.size:
                        ok, _, msg = O.check_near(O.scatter_block_cyclic(B0, mb, (P, Q), src)[(myrow, mycol)], work_a, 0.0,
                                                  O.gen_to_std_tolerance(m, dt))
                        if not ok:
                            failures.append(("hegst", t, uplo, m, mb, msg))
                        if not np.array_equal(work_t, t_before):
                            failures.append(("hegst factor modified", t, uplo, m, mb))
        for (m, mb, t) in [(1536, 256, "d"), (1100, 200, "d"), (768, 128, "z"), (1024, 256, "s")]:
            dt = pkg.TYPES[t]
            for uplo in "LU":
                A0 = O.set_random_hermitian_positive_definite(m, mb, dt)
                B0 = O.set_random_hermitian_positive_definite(m, mb, dt)
                B0 = np.asfortranarray(B0 + B0.conj().T)
                fac = B0.copy(order="F")
                assert O.cholesky_local(uplo, fac, mb, 4) == 0
                tri = np.tril if uplo == "L" else np.triu
                sent = np.full((m, m), -9.9)
                sent = (np.triu(sent, 1) if uplo == "L" else np.tril(sent, -1)).astype(dt)
                fac = np.asfortranarray(tri(fac) + sent)
                a_in = np.asfortranarray(tri(A0) + sent)
                ref = a_in.copy(order="F")
                O.generalized_to_standard(uplo, ref, fac, mb)
                loc_a, work_a = local_of(a_in, mb, dt)
                loc_t, work_t = local_of(fac, mb, dt)
                pkg.generalized_to_standard(ctx, uplo, work_a, work_t, mb, n=m, isrc=src[0], jsrc=src[1])
                if rank < P * Q and loc_a.size:
                    tol = O.gen_to_std_tolerance(m, dt) * max(1.0, float(np.abs(tri(ref)).max()))
                    ok, _, msg = O.check_near(O.scatter_block_cyclic(ref, mb, (P, Q), src)[(myrow, mycol)], work_a, 0.0, tol)
                    if not ok:
                        failures.append(("hegst big", t, uplo, m, mb, msg))
    flag = torch.tensor([len(failures)], dtype=torch.int64, device="cuda" if a.mode == "gpu" else "cpu")
    dist.all_reduce(flag)
    if failures:
        print(f"rank {rank} FAILURES: {failures[:5]}", flush=True)
    if rank == 0:
        print(f"dist_worker mode={a.mode} grid={P}x{Q} order={a.order}: total failures {int(flag.item())}", flush=True)
    pkg.free_grid(ctx)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
