"""dem.fill_depressions on the GPU at a given size: time, relaxation passes, tie components; optional oracle comparison.
    python profiles/scripts/fill_case.py SIZE [--check] [--offset X] [--f64] [--tilt METRES]
--tilt adds a regional slope of METRES across the raster diagonal: depressions stay local (the usual case for a real DEM),
without it the fBm terrain has basin-scale sinks that fill into lakes of millions of cells."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import oracle  # noqa: E402
from pyflwdir_b200 import _device  # noqa: E402

size = int(sys.argv[1])
offset = float(sys.argv[sys.argv.index("--offset") + 1]) if "--offset" in sys.argv else 0.0
z = oracle.synth_elevation(size, size, seed=9)
tilt = float(sys.argv[sys.argv.index("--tilt") + 1]) if "--tilt" in sys.argv else 0.0
a = z * np.float32(700.0) + np.float32(offset)
if tilt:
    a = a + (np.arange(size, dtype=np.float32)[:, None] + np.arange(size, dtype=np.float32)[None, :]) * np.float32(tilt / (2 * size))
a = a.astype(np.float64 if "--f64" in sys.argv else np.float32)
g = _device.DeviceGraph(0)
g.fill_depressions(a[:256, :256])  # warm-up
for rep in range(2):
    t0 = time.perf_counter()
    filled, d8 = g.fill_depressions(a)
    dt = time.perf_counter() - t0
    print(f"size {size} {a.dtype} offset {offset}: {dt * 1e3:.1f} ms ({a.size / dt / 1e6:.1f} Mcells/s incl. host copies) stats {g.fill_stats}", flush=True)
if "--check" in sys.argv:
    t0 = time.perf_counter()
    want = oracle.dem.fill_depressions(a)
    dt = time.perf_counter() - t0
    print(f"oracle (C heap, 1 core): {dt * 1e3:.1f} ms ({a.size / dt / 1e6:.2f} Mcells/s); equal: "
          f"{np.array_equal(filled, want[0]) and np.array_equal(d8, want[1])}")
