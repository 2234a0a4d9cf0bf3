"""fillnodata(direction="up") / basins(custom outlets) at a given size, first call after a parse: path summaries (no ordering) vs
the level replay (which has to order the cells first).   python profiles/scripts/fillup_case.py SIZE"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402

size = int(sys.argv[1])
w = bench.Workload(size, 0, 0)
l, L, h, n = w.l, w.L, w.h, w.cells
f32, i32, u32 = (L.DTYPES[np.dtype(t)] for t in (np.float32, np.int32, np.uint32))
z_dev = w.dev_alloc(n * 4)
w.ck(l.pfd_synth_elevation(h, size, size, size, bench.octaves_for(size), 0, z_dev))
zh = np.empty(n, np.float32)
w.ck(l.pfd_memcpy(h, L.ptr(zh), z_dev, n * 4))
rng = np.random.default_rng(5)
sparse = np.where(rng.random(n) < 0.7, np.float32(-9999.0), zh)
sp_dev, out_dev = w.dev_alloc(n * 4), w.dev_alloc(n * 4)
w.ck(l.pfd_memcpy(h, sp_dev, L.ptr(sparse), n * 4))
outlets = np.ascontiguousarray(rng.choice(n, size=5000, replace=False).astype(np.int32))
ids = np.arange(1, outlets.size + 1, dtype=np.uint32)
reparse = lambda: w.ck(l.pfd_d8_parse(h, w.d8_dev, size, size, 1, None, 0, None, None, None))
calls = {"fillnodata_up_f32": lambda: w.ck(l.pfd_fillnodata(h, sp_dev, f32, C.c_double(-9999.0), 0, 0, 0, 0, out_dev)),
         "basins_5000_outlets": lambda: w.ck(l.pfd_basins(h, L.ptr(outlets), outlets.size, i32, L.ptr(ids), u32, out_dev))}
for name, fn in calls.items():
    res = {}
    for eng, opt in (("paths", 1), ("level_replay_incl_ordering", 0)):
        w.ck(l.pfd_set_option(h, b"hand_pathsum", opt))
        reparse(); fn()
        t = []
        for _ in range(3):
            reparse()
            t.append(w.timer(fn, 1))
        ck = C.c_uint64()
        w.ck(l.pfd_checksum(h, out_dev, 4, n, 0, C.byref(ck)))
        res[eng] = (min(t), ck.value)
    assert res["paths"][1] == res["level_replay_incl_ordering"][1], name
    print(f"{size}^2 {name}: paths {res['paths'][0]:.2f} ms, level replay incl. ordering {res['level_replay_incl_ordering'][0]:.2f} ms, same checksum")
