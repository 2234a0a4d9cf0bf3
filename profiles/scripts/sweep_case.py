"""One Strahler / float32 accuflux / HAND call on a synthetic size^2 raster (the target of the ncu captures of the
tile-dataflow sweeps): python profiles/scripts/sweep_case.py [size] [reps] [tile_sweeps]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 2
maxp = int(sys.argv[4]) if len(sys.argv) > 4 else 0
w = bench.Workload(size, 0, 0)
l, L, h, n = w.l, w.L, w.h, w.cells
w.ck(l.pfd_set_option(h, b"tile_sweeps", mode))
w.ck(l.pfd_set_option(h, b"sweep_max_passes", maxp))
w.step_resident()
z_dev = w.dev_alloc(n * 4)
w.ck(l.pfd_synth_elevation(h, size, size, size, bench.octaves_for(size), 0, z_dev))
upa = np.empty(n, np.int32)
w.ck(l.pfd_memcpy(h, L.ptr(upa), w.out_dev[2], n * 4))
drain = np.ascontiguousarray(upa > 1000).view(np.uint8)
drain_dev = w.dev_alloc(n)
w.ck(l.pfd_memcpy(h, drain_dev, L.ptr(drain), n))
so_dev, hand_dev = w.dev_alloc(n), w.dev_alloc(n * 8)
f32 = L.dtype_code(np.float32)
for _ in range(reps):
    t = {}
    info = lambda: (int(l.pfd_get_info(h, b"sweep_passes")), round(int(l.pfd_get_info(h, b"sweep_visits")) / ((size // 64) ** 2), 2))
    t["strahler"] = (round(w.timer(lambda: w.ck(l.pfd_strahler(h, None, so_dev)), 1), 3),) + info()
    t["accuflux_f32"] = (round(w.timer(lambda: w.ck(l.pfd_accuflux(h, z_dev, f32, C.c_double(-9999.0), 0, 0, 0, w.out_dev[1])), 1), 3),) + info()
    t["hand"] = (round(w.timer(lambda: w.ck(l.pfd_hand(h, drain_dev, z_dev, f32, hand_dev)), 1), 3),) + info()
    print("(ms, passes, visits per tile):", t, flush=True)
