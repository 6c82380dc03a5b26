"""Writes tests/golden/util_distribution_cases.json.

The rows are the tiles_per_block == 1, tile_offset == 0 cases (the only ones on the POTRF path:
factorization/cholesky.h:45,:75 asserts single_tile_per_block) of the reference's own index-conversion
table, test/unit/matrix/test_util_distribution.cpp:49-55. They are transcribed, not computed: the
reference cannot be built here (no pika/MPI), so its test table is the golden vector.
Columns: tile_size, rank, grid_size, src_rank, global_element, global_tile, rank_tile, local_tile,
local_tile_next, tile_element.
"""
import json
import os

CASES = [
    [10, 0, 1, 0, 31, 3, 0, 3, 3, 1], [10, 0, 5, 0, 102, 10, 0, 2, 2, 2],
    [10, 1, 5, 0, 124, 12, 2, -1, 3, 4], [10, 4, 5, 3, 124, 12, 0, -1, 3, 4],
    [25, 0, 1, 0, 231, 9, 0, 9, 9, 6], [25, 0, 5, 0, 102, 4, 4, -1, 1, 2],
    [25, 3, 5, 4, 102, 4, 3, 0, 0, 2], [25, 4, 5, 3, 0, 0, 3, -1, 0, 0],
    [25, 0, 5, 3, 0, 0, 3, -1, 0, 0], [25, 3, 5, 3, 0, 0, 3, 0, 0, 0],
]
KEYS = ["tile_size", "rank", "grid_size", "src_rank", "global_element", "global_tile", "rank_tile",
        "local_tile", "local_tile_next", "tile_element"]

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "util_distribution_cases.json")
    with open(out, "w") as f:
        json.dump({"source": "test/unit/matrix/test_util_distribution.cpp:49-55 (DLA-Future v0.10.0)",
                   "keys": KEYS, "cases": CASES}, f, indent=1)
    print(out)
