"""The header-only C++ surface (include/dlaf/**) compiled on the host, no GPU: matrix::Distribution held against the
reference's own table (tests/golden/util_distribution_cases.json <- test/unit/matrix/test_util_distribution.cpp:49-55) by a small
program (tests/cpp_headers_check.cpp) that also instantiates the public templates of every header (Cholesky, triangular
solver, inverse, generalized -> standard) — the signatures a user of the reference would call."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_distribution_header_against_reference_table(tmp_path):
    exe = tmp_path / "cpp_headers_check"
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", f"-I{ROOT}/include", "-I/usr/local/cuda/include", os.path.join(HERE, "cpp_headers_check.cpp"),
           "-o", str(exe), f"-L{ROOT}/dla-future_b200/lib", "-ldlaf_b200", "-L/usr/local/cuda/lib64", "-lcudart",
           f"-Wl,-rpath,{ROOT}/dla-future_b200/lib", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    with open(os.path.join(HERE, "golden", "util_distribution_cases.json")) as f:
        table = json.load(f)
    lines = []
    for (tile_size, rank, grid, src, _ge, gt, rank_tile, lt, lt_next, _te) in table["cases"]:
        lines.append(f"G2L {grid} {src} {rank} {gt} {rank_tile} {lt} {lt_next}")
        if lt >= 0:
            lines.append(f"L2G {grid} {src} {rank} {lt} {gt}")
    r = subprocess.run([str(exe)], input="\n".join(lines) + "\n", capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert f"{len(lines)} table lines, 0 mismatches" in r.stdout
