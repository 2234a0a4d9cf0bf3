import sys, time, numpy as np
sys.path.insert(0, ".")
import oracle, pyflwdir_b200 as pfb
from pyflwdir_b200 import _device
n = 16384
z = oracle.synth_elevation(n, n, seed=4)
a = (z * np.float32(700.0) + np.float32(2000.0) + (np.arange(n, dtype=np.float32)[:, None] + np.arange(n, dtype=np.float32)[None, :]) * np.float32(1.5)).astype(np.float32)
g = _device.DeviceGraph(0)
t0 = time.perf_counter(); filled, d8 = g.fill_depressions(a); dt = time.perf_counter() - t0
print("16384^2 fill incl. host copies", round(dt, 2), "s", g.fill_stats)
assert np.all(filled >= a)
flw = pfb.from_array(d8, ftype="d8", check_ftype=False)
rank = flw.rank
print("all cells reach a pit:", bool(np.all(rank >= 0)), "pits:", flw.idxs_pit.size, "pits on the raster edge only:",
      bool(np.all((flw.idxs_pit // n == 0) | (flw.idxs_pit // n == n - 1) | (flw.idxs_pit % n == 0) | (flw.idxs_pit % n == n - 1))))
