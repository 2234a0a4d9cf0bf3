"""HAND (BASELINE config 5) at a given size: path sums vs the hop-by-hop engines, device-resident.
    python profiles/scripts/hand_case.py SIZE"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402

size = int(sys.argv[1])
w = bench.Workload(size, 3, 0)
l, L, h, n = w.l, w.L, w.h, w.cells
f32 = L.DTYPES[np.dtype(np.float32)]
z_dev = w.dev_alloc(n * 4)
w.ck(l.pfd_synth_elevation(h, size, size, size, bench.octaves_for(size), 3, z_dev))
w.step_resident()
upa = np.empty(n, np.int32)
w.ck(l.pfd_memcpy(h, L.ptr(upa), w.out_dev[2], n * 4))
drain_h = (upa > 1000).astype(np.uint8)
drain_dev = w.dev_alloc(n)
w.ck(l.pfd_memcpy(h, drain_dev, L.ptr(drain_h), n))
out_dev = w.dev_alloc(n * 8)
res = {}
for name, opts in (("pathsum", dict(hand_pathsum=1)), ("tile_sweep", dict(hand_pathsum=0, tile_sweeps=2))):
    for k, v in opts.items():
        w.ck(l.pfd_set_option(h, k.encode(), v))
    fn = lambda: w.ck(l.pfd_hand(h, drain_dev, z_dev, f32, out_dev))
    fn()
    ms = min(w.timer(fn, 1) for _ in range(3))
    eng = l.pfd_get_info(h, b"hand_engine")
    ck = C.c_uint64()
    w.ck(l.pfd_checksum(h, out_dev, 8, n, 0, C.byref(ck)))
    res[name] = (ms, eng, ck.value)
    print(f"{size}^2 hand {name}: {ms:.2f} ms, engine {eng}, checksum {ck.value:016x}", flush=True)
assert res["pathsum"][2] == res["tile_sweep"][2], "engines disagree"
