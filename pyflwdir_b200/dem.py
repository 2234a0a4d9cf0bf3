"""Elevation-coupled sweeps on the GPU; mirrors /root/reference/pyflwdir/dem.py (height_above_nearest_drain :299-330,
floodplains :333-379). The serial DEM-conditioning algorithms of that module are not provided."""
import numpy as np

from . import _functional


def height_above_nearest_drain(idxs_ds, seq, drain, elevtn, shape=None, ncol=None):
    """Returns the height above the nearest drain (HAND), float64, -9999 outside the sequence."""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "height_above_nearest_drain")
    return g.hand(np.asarray(drain).ravel(), np.asarray(elevtn).ravel())


def floodplains(idxs_ds, seq, elevtn, uparea, upa_min=1000.0, b=0.3, shape=None, ncol=None):
    """Returns floodplain boundaries from a HAND threshold that scales with upstream area, h ~ A**b (dem.py:333-379):
    int8, 1 floodplain, 0 not, -1 outside the sequence. uparea ** b of the drain cells is evaluated on the host with
    numpy scalars, exactly as the reference's Python loop does."""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "floodplains")
    upa = np.asarray(uparea).ravel()
    drain = np.flatnonzero(upa >= upa_min)
    drainh = np.full(g.size, -9999.0, dtype=np.float32)
    if drain.size:
        drainh[drain] = np.array([u ** b for u in upa[drain]], dtype=np.float64).astype(np.float32) \
            if upa.dtype != np.float32 else np.array([u ** b for u in upa[drain]], dtype=np.float32)
    return g.floodplains(drainh, np.asarray(elevtn).ravel())


def _serial(name, where):
    def fn(*args, **kwargs):
        raise NotImplementedError(
            f"dem.{name} ({where}) is a serial algorithm whose result depends on its visiting order cell by cell; it is "
            "outside the D8 hot path that pyflwdir_b200 accelerates (DESIGN.md §1) -- use Deltares/pyflwdir for it")

    fn.__name__ = name
    return fn


fill_depressions = _serial("fill_depressions", "pyflwdir/dem.py:17-143, priority-flood")
adjust_elevation = _serial("adjust_elevation", "pyflwdir/dem.py:146-168")
dig_4connectivity = _serial("dig_4connectivity", "pyflwdir/dem.py:404-439")
slope = _serial("slope", "pyflwdir/dem.py:228-296; its hypot is the C library's, which is not correctly rounded")
