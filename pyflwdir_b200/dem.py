"""Elevation-coupled algorithms on the GPU; mirrors /root/reference/pyflwdir/dem.py (fill_depressions :17-143,
height_above_nearest_drain :299-330, floodplains :333-379). The remaining serial DEM-conditioning algorithms of that module
are not provided."""
import numpy as np

from . import _device, _functional

__all__ = ["fill_depressions", "height_above_nearest_drain", "floodplains"]


def fill_depressions(elevtn, outlets="edge", idxs_pit=None, nodata=-9999.0, max_depth=-1.0, elv_max=None, connectivity=8,
                     device=0):
    """Fill local depressions in elevation data and derive local D8 flow directions (Wang & Liu 2006), the reference's
    priority flood (dem.py:17-143) reproduced bit for bit on the GPU (csrc/pfd_fill.cuh): levels by parallel min/max
    relaxation, the heap order inside flats / filled lakes replayed per tie component.

    Outlets are the edge cells of the valid elevation cells (`outlets='edge'`, optionally only those with elevation
    <= `elv_max`), the lowest edge cell (`outlets='min'`) or the cells `idxs_pit`. `connectivity` is 4 or 8.
    `max_depth >= 0` (pits at depressions deeper than max_depth) is not implemented: NotImplementedError.

    Returns (elevtn_out, d8): the depression-filled elevation (dtype of `elevtn`) and uint8 D8 flow directions."""
    if connectivity not in (4, 8):
        raise ValueError('"connectivity" should either be 4 or 8')
    if max_depth >= 0:
        raise NotImplementedError(
            "dem.fill_depressions(max_depth >= 0) re-opens visited cells in heap order (pyflwdir/dem.py:121-132); only "
            "max_depth < 0 (fill every depression) is implemented on the device -- use Deltares/pyflwdir for it")
    g = _device.DeviceGraph(device)
    try:
        return g.fill_depressions(elevtn, outlets=outlets, idxs_pit=idxs_pit, nodata=nodata, max_depth=max_depth,
                                  elv_max=elv_max, connectivity=connectivity)
    finally:
        g.close()


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


adjust_elevation = _serial("adjust_elevation", "pyflwdir/dem.py:146-168")
dig_4connectivity = _serial("dig_4connectivity", "pyflwdir/dem.py:404-439")
slope = _serial("slope", "pyflwdir/dem.py:228-296; its hypot is the C library's, which is not correctly rounded")
