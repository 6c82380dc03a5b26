"""Measurement aid: factor ONE nb x nb tile through the device-resident C ABI with the diagonal-block kernel variant
selected by DLAF_B200_POTRF_KERNEL (unset = blocked single-CTA kernel, cluster2 = two-SM cluster kernel) and check it
against numpy: residual max|A - L L^T| / max|A| and the time of the whole tile (median of 10)."""
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
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
h = np.zeros((nb, nb), dtype=np.float64, order="F")
pkg.set_random_hermitian_positive_definite(ctx, h, nb, nb)
d_ref = torch.from_numpy(np.ascontiguousarray(h.T)).cuda()
d_work = torch.empty_like(d_ref)
times = []
for i in range(12):
    d_work.copy_(d_ref)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    pkg.cholesky_factorization_device(ctx, "L", d_work.data_ptr(), np.float64, nb, nb, nb, stream.cuda_stream)
    b.record(stream)
    info = pkg.wait(ctx, stream.cuda_stream)
    times.append(a.elapsed_time(b) * 1e3)
L = np.tril(d_work.cpu().numpy().T)
res = np.abs(np.tril(L @ L.T - h)).max() / np.abs(h).max()
t = sorted(times[2:])
print(f"variant {os.environ.get('DLAF_B200_POTRF_KERNEL', 'blocked')}: info {info} residual {res:.2e} tile {t[len(t)//2]:.1f} us (min {t[0]:.1f})")
pkg.free_grid(ctx)
