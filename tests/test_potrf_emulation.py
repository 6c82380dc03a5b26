"""Host emulation of the diagonal-block kernels (no GPU in the build container): the __host__ __device__ per-thread
phase functions of dla-future_b200/csrc/potrf_block.cuh — the very code the CUDA kernels run — executed on the CPU
and checked against host loops (L, inv(L), untouched upper triangle, LAPACK-style info on non-SPD input):
  tools/potrf_block_emu.cu    single-CTA kernel, 256 "threads" phase by phase
  tools/potrf_cluster_emu.cu  two-CTA cluster variant (potrf_cluster.cuh): the SAME orchestration code on 2 x 128 host
                              threads with std::barrier as cta / cluster barrier and plain stores as DSMEM pushes,
                              also under random schedule jitter
Reference behaviour being restated: lapack::potrf on one tile (include/dlaf/lapack/tile.h:448-464, info semantics
test/unit/test_lapack_tile/test_potrf.h:59-77)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

pytestmark = pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc (host compilation of the .cu emulators) not available")


def _build(src, out, std):
    subprocess.check_call([NVCC, f"-std={std}", "-O2", "-diag-suppress", "20014", "-o", out, os.path.join(ROOT, "tools", src)])


def test_single_cta_kernel_emulation(tmp_path):
    exe = str(tmp_path / "potrf_block_emu")
    _build("potrf_block_emu.cu", exe, "c++17")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FAIL" not in r.stdout and "MISMATCH" not in r.stdout, r.stdout


@pytest.mark.parametrize("jitter", [False, True])
def test_two_cta_cluster_kernel_emulation(tmp_path, jitter):
    exe = str(tmp_path / "potrf_cluster_emu")
    _build("potrf_cluster_emu.cu", exe, "c++20")
    env = dict(os.environ)
    if jitter:
        env["EMU_JITTER"] = "1"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all OK" in r.stdout and "FAIL" not in r.stdout and "MISMATCH" not in r.stdout, r.stdout
