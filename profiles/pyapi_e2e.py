import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import ctypes as C
import pyflwdir_b200 as pfb
from pyflwdir_b200 import _lib, _device
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = _device.DeviceGraph(0)
z = np.empty((size, size), np.float32); d8 = np.empty((size, size), np.uint8)
_lib.check(_lib.lib().pfd_synth_elevation(dev._h, size, size, size, int(np.log2(size)) - 2, 0, _lib.ptr(z)), dev._h)
_lib.check(_lib.lib().pfd_synth_d8(dev._h, _lib.ptr(z), size, size, C.c_float(-np.inf), _lib.ptr(d8)), dev._h)
dev.close()
for rep in range(3):
    t0 = time.perf_counter(); flw = pfb.from_array(d8, ftype="d8"); t1 = time.perf_counter()
    upa = flw.upstream_area(); t2 = time.perf_counter()
    bas = flw.basins(); t3 = time.perf_counter()
    rnk = flw.rank; t4 = time.perf_counter()
    ids = flw.idxs_ds; t5 = time.perf_counter()
    sto = flw.stream_order(); t6 = time.perf_counter()
    print(f"rep{rep}: from_array {1e3*(t1-t0):.1f} ms | upstream_area {1e3*(t2-t1):.1f} | basins {1e3*(t3-t2):.1f} | rank {1e3*(t4-t3):.1f} | idxs_ds {1e3*(t5-t4):.1f} | stream_order {1e3*(t6-t5):.1f} | total {1e3*(t6-t0):.1f} ms -> {d8.size/(t6-t0)/1e6:.0f} Mcells/s")
