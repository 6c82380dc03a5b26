#!/usr/bin/env python
"""The consumers of the factor on a GPU grid, fp64: under torchrun every rank generates its local part of
the miniapp's matrix, factorises it (dlaf_cholesky_factorization_d, host buffers), then times
dlaf_inverse_from_cholesky_factor_d, dlaf_b200_generalized_to_standard_d and dlaf_b200_triangular_solver_d through their
host entry points; reported: the library's CUDA-event time of the device-resident part, max over ranks. Rank 0 prints one
JSON line. Correctness on grids is the business of tests/dist_worker.py; here two grid-independent fingerprints tie the
results together: trace(inv(A)) and, for the pencil (A - n I, A), trace(C) = n - n trace(inv(A)).
usage: torchrun --nproc-per-node 4 tools/bench_dist_next.py --grid 2x2 [--matrix-size 32768] [--block-size 512]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="2x2")
    ap.add_argument("--matrix-size", dest="n", type=int, default=32768)
    ap.add_argument("--block-size", dest="nb", type=int, default=512)
    ap.add_argument("--nrhs", type=int, default=0)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    P, Q = (int(x) for x in a.grid.split("x"))
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    pkg = ge.load_package()
    pkg.initialize()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        comm = pkg.comm_create_from_torch()
    else:
        comm = None
    ctx = pkg.create_grid(comm, P, Q, "C")
    n, nb = a.n, a.nb
    nrhs = a.nrhs or n // 2
    d = pkg.descriptor(n, nb, 1)
    lrows, lcols = pkg.local_shape(ctx, d)
    loc = np.zeros((lrows, lcols), order="F")
    pkg.set_random_hermitian_positive_definite(ctx, loc, n, nb)
    fac = loc.copy(order="F")
    assert pkg.cholesky_factorization(ctx, "L", fac, nb, n=n) == 0

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return t.item()

    out = {}
    # inverse from the factor: trace(inv(A)) as a grid-independent fingerprint
    ms = []
    for _ in range(2):
        w = fac.copy(order="F")
        pkg.grid_barrier(ctx)
        pkg.inverse_from_cholesky_factor(ctx, "L", w, nb, n=n)
        ms.append(allmax(pkg.last_solver_device_ms(ctx)))
    _, _, myrow, mycol = pkg.grid_info(ctx)
    tr = 0.0
    for li in range(0, lrows, nb):
        gi = (li // nb) * P + myrow
        for lj in range(0, lcols, nb):
            gj = (lj // nb) * Q + mycol
            if gi == gj:
                tr += float(np.trace(w[li:li + nb, lj:lj + nb]))
    out["inverse_from_cholesky_factor"] = {"ms_device_max_over_ranks": min(ms), "value": 2 * n ** 3 / 3 / (min(ms) * 1e-3) / 1e9,
                                           "unit": "GFLOP/s", "trace_inverse": allsum(tr),
                                           "guard_fallback_steps": pkg.last_inverse_guard_steps(ctx)}
    if rank == 0:
        print("[partial] " + json.dumps(out), file=sys.stderr, flush=True)
    # generalized -> standard with A = B - n I (Hermitian, not a multiple of B): C = I - n inv(L) inv(L)^H, so
    # trace(C) = n - n trace(inv(B)) ties this result to the inverse above
    a_loc = loc.copy(order="F")
    for li in range(0, lrows, nb):
        gi = (li // nb) * P + myrow
        for lj in range(0, lcols, nb):
            if gi == (lj // nb) * Q + mycol:
                blk = a_loc[li:li + nb, lj:lj + nb]
                blk[np.arange(blk.shape[0]), np.arange(blk.shape[0])] -= n
    ms = []
    for _ in range(2):
        w = a_loc.copy(order="F")
        pkg.grid_barrier(ctx)
        pkg.generalized_to_standard(ctx, "L", w, fac, nb, n=n)
        ms.append(allmax(pkg.last_solver_device_ms(ctx)))
    trc = 0.0
    for li in range(0, lrows, nb):
        gi = (li // nb) * P + myrow
        for lj in range(0, lcols, nb):
            if gi == (lj // nb) * Q + mycol:
                trc += float(np.trace(w[li:li + nb, lj:lj + nb]))
    trc = allsum(trc)
    expect = n - n * out["inverse_from_cholesky_factor"]["trace_inverse"]
    out["generalized_to_standard"] = {"ms_device_max_over_ranks": min(ms), "value": float(n) ** 3 / (min(ms) * 1e-3) / 1e9,
                                      "unit": "GFLOP/s", "trace_C": trc, "trace_C_expected_from_the_inverse": expect,
                                      "rel_diff": abs(trc - expect) / abs(expect),
                                      "guard_fallback_steps": pkg.last_inverse_guard_steps(ctx)}
    if rank == 0:
        print("[partial] " + json.dumps(out), file=sys.stderr, flush=True)
    # triangular solver L X = B
    ntc = -(-nrhs // nb)
    lcb = ((ntc - mycol + Q - 1) // Q if ntc > mycol else 0) * nb
    if (ntc - 1) % Q == mycol:
        lcb -= ntc * nb - nrhs
    rng = np.random.default_rng(1 + rank)
    rhs = np.asfortranarray(rng.uniform(-1, 1, (lrows, lcb)))
    ms = []
    for _ in range(2):
        x = rhs.copy(order="F")
        pkg.grid_barrier(ctx)
        pkg.triangular_solver(ctx, "L", "L", "N", "N", 1.0, fac, x, nb, nb, m=n, n=nrhs)
        ms.append(allmax(pkg.last_solver_device_ms(ctx)))
    out["triangular_solver"] = {"ms_device_max_over_ranks": min(ms), "value": float(n) * n * nrhs / (min(ms) * 1e-3) / 1e9,
                                "unit": "GFLOP/s", "nrhs": nrhs}
    if rank == 0:
        print(json.dumps({"metric": f"consumers of the Cholesky factor, fp64 n={n} nb={nb}, grid {P}x{Q}", "n_gpus": world,
                          "timing": "CUDA events inside the library around the device-resident part, max over ranks; host<->device copies excluded",
                          **out}), flush=True)
    pkg.free_grid(ctx)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
