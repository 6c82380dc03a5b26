"""Measurement aid: time ONE diagonal tile (N = nb) in isolation through the device-resident C ABI, i.e. the
critical-path part of every POTRF step without any bulk update running next to it.
Usage: python tools/chain_probe.py [nb ...]   (env: DLAF_B200_DIAG_LOOKAHEAD=0 for the plain sequence)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
pkg.initialize()
ctx = pkg.create_grid(None, 1, 1, "C")
stream = torch.cuda.current_stream()
for t, dtype in (("d", np.float64), ("z", np.complex128)):
    for nb in [int(a) for a in sys.argv[1:]] or [512]:
        n = nb
        h = np.zeros((n, n), dtype=dtype, order="F")
        pkg.set_random_hermitian_positive_definite(ctx, h, n, nb)
        tdt = torch.float64 if t == "d" else torch.complex128
        d_ref = torch.from_numpy(np.ascontiguousarray(h.T)).cuda()
        d_work = torch.empty_like(d_ref)
        times = []
        for i in range(12):
            d_work.copy_(d_ref)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            pkg.cholesky_factorization_device(ctx, "L", d_work.data_ptr(), dtype, n, nb, n, stream.cuda_stream)
            b.record(stream)
            assert pkg.wait(ctx, stream.cuda_stream) == 0
            times.append(a.elapsed_time(b) * 1e3)
        times = sorted(times[2:])
        print(f"diag tile alone type {t} nb={nb}: median {times[len(times)//2]:.1f} us, min {times[0]:.1f} us, "
              f"{pkg.last_launch_count(ctx)} launches (lookahead={os.environ.get('DLAF_B200_DIAG_LOOKAHEAD', '1')})")
pkg.free_grid(ctx)
