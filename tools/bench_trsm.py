#!/usr/bin/env python
"""Triangular-solver measurement: op(A) X = alpha B through dlaf_b200_triangular_solver_d with HOST
buffers (the C entry's contract) and the GEMM-dominated device time inside, next to cuBLAS Dtrsm (torch) on the same box.
One JSON line. usage: python tools/bench_trsm.py [--n 16384] [--nrhs 16384] [--nb 512] [--side L --uplo L --op N]"""
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
    ap.add_argument("--nrhs", type=int, default=16384)
    ap.add_argument("--nb", type=int, default=512)
    ap.add_argument("--side", default="L")
    ap.add_argument("--uplo", default="L")
    ap.add_argument("--op", default="N")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import torch

    pkg = ge.load_package()
    pkg.initialize()
    ctx = pkg.create_grid(None, 1, 1, "C")
    n, nrhs, nb = a.n, a.nrhs, a.nb
    m, nn = (n, nrhs) if a.side == "L" else (nrhs, n)
    spd = np.zeros((n, n), order="F")
    pkg.set_random_hermitian_positive_definite(ctx, spd, n, nb)
    assert pkg.cholesky_factorization(ctx, a.uplo, spd, nb) == 0
    tri = np.tril if a.uplo == "L" else np.triu
    A = np.asfortranarray(tri(spd))
    rng = np.random.default_rng(1)
    B = np.asfortranarray(rng.uniform(-1, 1, (m, nn)))
    times, dev_ms = [], []
    for i in range(1 + a.steps):
        X = B.copy(order="F")
        t0 = time.perf_counter()
        pkg.triangular_solver(ctx, a.side, a.uplo, a.op, "N", 1.0, A, X, nb, nb)
        dt = time.perf_counter() - t0
        if i:
            times.append(dt)
            dev_ms.append(pkg.last_solver_device_ms(ctx))
    flops = float(n) * n * nrhs
    # residual of the solution
    opa = {"N": A, "T": A.T, "C": A.T}[a.op]
    dA, dX, dB = (torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in (opa, X, B))
    res = ((dA @ dX if a.side == "L" else dX @ dA) - dB).abs().max().item() / (np.abs(X).max() * np.abs(A).max() * n)
    # vendor reference: cuBLAS trsm through torch, device-resident
    upper = (a.uplo == "U") != (a.op != "N")
    tt = []
    for i in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        torch.linalg.solve_triangular(dA, dB, upper=upper, left=(a.side == "L"))
        torch.cuda.synchronize()
        tt.append(time.perf_counter() - t0)
    e2e = min(times)
    line = {"metric": f"triangular solver GFLOP/s (fp64, {a.side}{a.uplo}{a.op}, n={n}, nrhs={nrhs}, nb={nb})",
            "value": flops / (min(dev_ms) * 1e-3) / 1e9, "ms_device": min(dev_ms),
            "value_e2e_host_buffers": flops / e2e / 1e9, "ms_e2e": e2e * 1e3, "unit": "GFLOP/s",
            "h2d_bytes": A.nbytes + B.nbytes, "d2h_bytes": B.nbytes, "launches": pkg.last_solver_launch_count(ctx),
            "residual_max_over_n_maxA_maxX": res, "eps": float(np.finfo(np.float64).eps),
            "gpu_library_reference": {"kind": "cuBLAS Dtrsm via torch.linalg.solve_triangular, device-resident", "ms": min(tt) * 1e3,
                                      "value": flops / min(tt) / 1e9},
            "fp64_tensor_peak_tflops": pkg.measure_fp64_tensor_peak_tflops(),
            "engine": os.environ.get("DLAF_B200_D_BULK", "ozaki")}
    print(json.dumps(line), flush=True)
    pkg.free_grid(ctx)


if __name__ == "__main__":
    main()
