#!/usr/bin/env python
"""Inverse-from-Cholesky-factor measurement: POTRF (product) -> dlaf_b200_inverse_device_* on the
device-resident factor (CUDA-event time of the whole call: layout conversion, diagonal-tile inverses, both sweeps), the
host-buffer call dlaf_inverse_from_cholesky_factor_* (e2e), the residual max|inv(A) A - I|, next to the vendor library
(cuSOLVER potri via torch.cholesky_inverse) on the same box. One JSON line.
usage: python tools/bench_inverse.py [--n 16384] [--nb 512] [--type d] [--phases 3]"""
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
    ap.add_argument("--uplo", default="L")
    ap.add_argument("--phases", type=int, default=3)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    import torch

    pkg = ge.load_package()
    pkg.initialize()
    ctx = pkg.create_grid(None, 1, 1, "C")
    n, nb = a.n, a.nb
    dt = pkg.TYPES[a.type]
    cplx = np.dtype(dt).kind == "c"
    spd = np.zeros((n, n), dtype=dt, order="F")
    pkg.set_random_hermitian_positive_definite(ctx, spd, n, nb)
    fac = spd.copy(order="F")
    assert pkg.cholesky_factorization(ctx, a.uplo, fac, nb) == 0
    tdt = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}[a.type]
    # device-resident timing: torch tensor with the numpy (column-major) layout = transposed row-major tensor
    d_fac = torch.from_numpy(np.ascontiguousarray(fac.T)).cuda()  # memory = column-major fac
    dev_ms = []
    out = None
    for i in range(1 + a.steps):
        work = d_fac.clone()
        torch.cuda.synchronize()
        pkg.inverse_device(ctx, a.phases, a.uplo, "N", work.data_ptr(), dt, n, nb, n)
        if i:
            dev_ms.append(pkg.last_solver_device_ms(ctx))
        out = work
    launches = pkg.last_solver_launch_count(ctx)
    guard = pkg.last_inverse_guard_steps(ctx)
    # flop model (LAPACK working notes): trtri n^3/3, lauum n^3/3 (complex x4)
    flops = (n ** 3 / 3.0) * (bin(a.phases & 3).count("1")) * (4 if cplx else 1)
    res = None
    if a.phases == 3:
        inv_cm = out.T  # column-major view: element (i, j) of the result
        tri = torch.tril if a.uplo == "L" else torch.triu
        low = tri(inv_cm)
        full = low + (torch.tril(low, -1) if a.uplo == "L" else torch.triu(low, 1)).conj().T
        dA = torch.from_numpy(np.ascontiguousarray(spd)).cuda()
        res = (full @ dA - torch.eye(n, dtype=tdt, device="cuda")).abs().max().item()
        del dA, full, low
    e2e = None
    if not a.no_e2e:
        tt = []
        for i in range(2):
            h = fac.copy(order="F")
            t0 = time.perf_counter()
            if a.phases == 3:
                pkg.inverse_from_cholesky_factor(ctx, a.uplo, h, nb)
            elif a.phases == 1:
                pkg.triangular_inverse(ctx, a.uplo, "N", h, nb)
            else:
                pkg.assemble_cholesky_inverse(ctx, a.uplo, h, nb)
            tt.append(time.perf_counter() - t0)
        e2e = min(tt)
    # vendor reference: cuSOLVER potri (torch.cholesky_inverse), device-resident, same factor
    ref = None
    if a.phases == 3:
        try:
            L = torch.from_numpy(np.ascontiguousarray(np.tril(fac) if a.uplo == "L" else np.triu(fac))).cuda()
            tv = []
            for i in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                torch.cholesky_inverse(L, upper=(a.uplo == "U"))
                torch.cuda.synchronize()
                tv.append(time.perf_counter() - t0)
            ref = {"kind": "torch.cholesky_inverse (cuSOLVER potri / potrs path), device-resident", "ms": min(tv[1:]) * 1e3,
                   "value": flops / min(tv[1:]) / 1e9}
        except Exception as e:  # measurement aid only
            ref = {"error": str(e)[:200]}
    line = {"metric": f"inverse phases={a.phases} GFLOP/s ({a.type}, uplo {a.uplo}, n={n}, nb={nb})",
            "value": flops / (min(dev_ms) * 1e-3) / 1e9, "ms_device": min(dev_ms), "unit": "GFLOP/s", "flops_model": flops,
            "value_e2e_host_buffers": (flops / e2e / 1e9) if e2e else None, "ms_e2e": e2e * 1e3 if e2e else None,
            "launches": launches, "guard_fallback_steps": guard, "residual_max_abs_invA_A_minus_I": res,
            "gpu_library_reference": ref, "engine": os.environ.get("DLAF_B200_D_BULK", "ozaki")}
    print(json.dumps(line), flush=True)
    pkg.free_grid(ctx)


if __name__ == "__main__":
    main()
