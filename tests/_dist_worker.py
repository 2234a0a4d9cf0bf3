"""Worker of tests/test_dist.py: the per-rank plumbing bench.py uses for N > 1 (gloo, world_size 2, CPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402

dist, rank, world, local_rank = bench.dist_setup(2)
assert dist is not None and world == 2 and rank in (0, 1) and local_rank == rank
bench.barrier(dist)
assert bench.reduce_max(dist, 10.0 + rank) == 11.0
assert bench.reduce_sum(dist, 3.0 + rank) == 7.0
# every rank gets its own raster (seed + rank) and rank 0 reports the aggregate
cells = 1024 * 1024
ms = [4.0, 5.0][rank]
ms_max = bench.reduce_max(dist, ms)
value = cells * world * 3 / (ms_max / 1e3) / 1e6
if rank == 0:
    assert abs(value - cells * 2 * 3 / 5e-3 / 1e6) < 1e-6
    print("DIST_OK", flush=True)
bench.barrier(dist)
dist.destroy_process_group()
