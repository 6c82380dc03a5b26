"""The miniapp driver (C++ surface: dlaf::cholesky_factorization<Backend::GPU, Device::GPU, T>, Matrix,
MatrixMirror, CommunicatorGrid) — command line and output lines of the reference's miniapp_cholesky
(miniapp/miniapp_cholesky.cpp:165-188, :440-445), and the GPU-side residual check."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "miniapp", "miniapp_cholesky")


@pytest.mark.parametrize("t,uplo,n,nb", [("d", "L", 2048, 256), ("d", "U", 1000, 128), ("z", "L", 768, 128),
                                         ("s", "L", 1024, 256), ("c", "U", 640, 64)])
def test_miniapp_output_and_check(t, uplo, n, nb):
    assert os.path.exists(EXE), "run __graft_entry__.build() first"
    r = subprocess.run([EXE, "--matrix-size", str(n), f"--block-size={nb}", "--type", t, "--uplo", uplo, "--nruns", "2",
                        "--nwarmups", "1", "--check-result", "all", "--csv", "--local"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    runs = [l for l in lines if re.match(r"^\[\d+\] [0-9.e+-]+s [0-9.e+-]+GFlop/s", l)]
    assert len(runs) == 2, r.stdout
    assert runs[0].endswith(f"{t}{uplo} ({n}, {n}) ({nb}, {nb}) (1, 1) 1 GPU"), runs[0]
    assert sum(l.startswith("CSVData-2, run, ") for l in lines) == 2
    checks = [l for l in lines if "Max Diff / Max A:" in l]
    assert len(checks) == 3  # warm-up + 2 runs with --check-result all
    assert not any(l.startswith("ERROR") or l.startswith("Warning") for l in checks), checks


def test_gpu_check_matches_oracle_residual(pkg, oracle, grid11):
    n, nb = 1536, 256
    a = oracle.set_random_hermitian_positive_definite(n, nb, np.float64)
    f = a.copy(order="F")
    assert pkg.cholesky_factorization(grid11, "L", f, nb) == 0
    r_gpu = pkg.check_cholesky(grid11, "L", a, f, nb)
    r_cpu = oracle.residual("L", a, f)
    assert 0 <= r_gpu <= oracle.residual_gate(np.float64, n)[0]
    assert abs(r_gpu - r_cpu) <= 0.5 * max(r_gpu, r_cpu) + 1e-18
    g = a.copy(order="F")
    g[700, 100] += 1.0  # a wrong factor must be flagged
    assert pkg.check_cholesky(grid11, "L", a, g, nb) > 1e-6
