"""Multi-process tests: world_size-2 gloo run of the N>1 host logic on CPU; NCCL runs of the distributed
POTRF when at least 2 GPUs are visible (gpurun --gpus 2/4/8)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = os.path.join(HERE, "dist_worker.py")


def _launch(nproc, args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER] + args
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "total failures 0" in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("grid,order", [("2x1", "R"), ("1x2", "C")])
def test_host_logic_two_ranks_gloo(grid, order):
    _launch(2, ["--mode", "cpu", "--grid", grid, "--order", order], 29611 if grid == "2x1" else 29612)


def _ngpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("grid,order", [("2x1", "R"), ("1x2", "C")])
def test_distributed_potrf_two_gpus(grid, order):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    _launch(2, ["--mode", "gpu", "--grid", grid, "--order", order], 29621)


@pytest.mark.gpu
@pytest.mark.parametrize("grid,order", [("2x2", "C"), ("1x4", "R")])
def test_distributed_potrf_four_gpus(grid, order):
    if _ngpus() < 4:
        pytest.skip("needs 4 GPUs")
    _launch(4, ["--mode", "gpu", "--grid", grid, "--order", order], 29622)


@pytest.mark.gpu
@pytest.mark.parametrize("grid,order", [("2x4", "C"), ("3x2", "R")])
def test_distributed_potrf_eight_gpus(grid, order):
    if _ngpus() < 8:
        pytest.skip("needs 8 GPUs")
    _launch(8, ["--mode", "gpu", "--grid", grid, "--order", order], 29623)
