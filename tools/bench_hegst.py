#!/usr/bin/env python
"""Generalized -> standard reduction measurement: POTRF(B) (product) -> dlaf_b200_generalized_to_standard_device_*
on device-resident A and factor (CUDA-event time of the whole call), residual through max|L C L^H - A| / max|A|, next to the
same reduction by two cuBLAS triangular solves on the full matrix (torch) on the same box. One JSON line.
usage: python tools/bench_hegst.py [--n 16384] [--nb 512] [--type d]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--nb", type=int, default=512)
    ap.add_argument("--type", default="d")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import torch

    pkg = ge.load_package()
    pkg.initialize()
    ctx = pkg.create_grid(None, 1, 1, "C")
    n, nb = a.n, a.nb
    dt = pkg.TYPES[a.type]
    cplx = np.dtype(dt).kind == "c"
    bm = np.zeros((n, n), dtype=dt, order="F")
    pkg.set_random_hermitian_positive_definite(ctx, bm, n, nb)
    fac = bm.copy(order="F")
    assert pkg.cholesky_factorization(ctx, "L", fac, nb) == 0
    am = bm.copy(order="F")  # any Hermitian matrix will do: reuse the generator's (A = B gives C = I: use a shifted copy)
    am[np.arange(n), np.arange(n)] -= n  # diagonal n instead of 2n: still Hermitian, not a multiple of B
    d_fac = torch.from_numpy(np.ascontiguousarray(fac.T)).cuda()  # memory = column-major
    d_a = torch.from_numpy(np.ascontiguousarray(am.T)).cuda()
    dev_ms = []
    out = None
    for i in range(1 + a.steps):
        work = d_a.clone()
        torch.cuda.synchronize()
        pkg.generalized_to_standard_device(ctx, "L", work.data_ptr(), d_fac.data_ptr(), dt, n, nb, n)
        if i:
            dev_ms.append(pkg.last_solver_device_ms(ctx))
        out = work
    launches = pkg.last_solver_launch_count(ctx)
    guard = pkg.last_inverse_guard_steps(ctx)
    flops = float(n) ** 3 * (4 if cplx else 1)  # xHEGST itype 1: n^3 (LAPACK working note 41)
    # residual: L C L^H = A (lower parts)
    c_cm = out.T
    low = torch.tril(c_cm)
    cfull = low + torch.tril(low, -1).conj().T
    lmat = torch.tril(d_fac.T)
    back = lmat @ cfull @ lmat.conj().T
    aref = torch.from_numpy(np.ascontiguousarray(am)).cuda()
    res = (torch.tril(back - aref)).abs().max().item() / aref.abs().max().item()
    del back, cfull, low
    # vendor reference: two cuBLAS triangular solves on the full Hermitian matrix (2 n^3 flops for the same result)
    tv = []
    for i in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        y = torch.linalg.solve_triangular(lmat, aref, upper=False, left=True)
        z = torch.linalg.solve_triangular(lmat.conj().T, y, upper=True, left=False)
        torch.cuda.synchronize()
        tv.append(time.perf_counter() - t0)
    line = {"metric": f"generalized_to_standard GFLOP/s ({a.type}, uplo L, n={n}, nb={nb})",
            "value": flops / (min(dev_ms) * 1e-3) / 1e9, "ms_device": min(dev_ms), "unit": "GFLOP/s", "flops_model": flops,
            "launches": launches, "guard_fallback_steps": guard, "residual_max_LCLh_minus_A_over_max_A": res,
            "gpu_library_reference": {"kind": "2 x cuBLAS trsm on the full matrix via torch.linalg.solve_triangular, device-resident",
                                      "ms": min(tv[1:]) * 1e3, "value_same_flop_model": flops / min(tv[1:]) / 1e9},
            "engine": os.environ.get("DLAF_B200_D_BULK", "ozaki")}
    print(json.dumps(line), flush=True)
    pkg.free_grid(ctx)


if __name__ == "__main__":
    main()
