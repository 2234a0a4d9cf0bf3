"""CaMa-Flood NEXTXY codec on the GPU; mirrors /root/reference/pyflwdir/core_nextxy.py (from_array :24-34, _from_array
:42-67, to_array :37-39, isvalid :86-103, read_nextxy :118-144) with the same signatures. X (column) and Y (row) are
one-based; -9 / -10 are pits, -9999 is nodata. The device graph keeps one byte per cell (slot of the downstream
neighbour), so rasters whose links leave the 8-neighbourhood are refused with a ValueError."""
import numpy as np

from . import _device, _functional, _lib
from . import gis_utils

__all__ = ["read_nextxy"]

_ftype = "nextxy"
_mv = np.int32(-9999)
_pv = np.array([-9, -10], dtype=np.int32)
_us = np.ones((2, 3, 3), dtype=np.int32) * 2
_us[:, 1, 1] = _pv[0]


def isformat(flwdir):
    """True for the two containers the reference accepts: ([:, :], [:, :]) or an array [2, :, :]"""
    return (isinstance(flwdir, tuple) and len(flwdir) == 2) or (
        isinstance(flwdir, np.ndarray) and flwdir.ndim == 3 and flwdir.shape[0] == 2)


def _planes(flwdir):
    if not isformat(flwdir):
        raise TypeError("NEXTXY flwdir data not understood")
    nextx, nexty = flwdir
    return nextx, nexty


def from_array(flwdir, dtype=np.intp, device=0):
    """convert NEXTXY data to 1D next downstream indices -> (idxs_ds, idxs_pit, n)"""
    nextx, nexty = _planes(flwdir)
    dt = np.dtype(dtype)
    fetch_dt = np.dtype(np.int64) if dt == np.uint64 else dt
    g = _device.DeviceGraph(device)
    idxs_ds = g.parse_nextxy(nextx, nexty, idx_dtype=fetch_dt, want_idxs=True, check=False)
    pits = g.fetch(_lib.ARR_PITS, fetch_dt)
    if dt == np.uint64:
        idxs_ds, pits = idxs_ds.astype(np.uint64), pits.astype(np.uint64)
    return idxs_ds, pits, int(g.n_valid)


def to_array(idxs_ds, shape, mv=None, device=0):
    """convert downstream linear indices to a [2, nrow, ncol] NEXTXY raster"""
    g = _functional.graph(idxs_ds, shape=shape, device=device)
    return g.fetch(_lib.ARR_NEXTXY).reshape((2,) + tuple(shape))


def isnodata(dd):
    """True if NEXTXY nodata"""
    return dd == _mv


def ispit(dd, _pv=_pv):
    """True if NEXTXY pit"""
    return np.logical_or(dd == _pv[0], dd == _pv[1])


def isvalid(flwdir):
    """True if NEXTXY raster is valid (core_nextxy.py:86-103; host side: dtype / shape tests and two reductions)"""
    isfmt1 = isinstance(flwdir, tuple) and len(flwdir) == 2
    isfmt2 = isinstance(flwdir, np.ndarray) and flwdir.ndim == 3 and flwdir.shape[0] == 2
    if not (isfmt1 or isfmt2):
        return False
    nextx, nexty = flwdir
    if not (isinstance(nextx, np.ndarray) and isinstance(nexty, np.ndarray)):
        return False
    if not (nexty.dtype == "int32" and nextx.dtype == "int32" and nexty.shape == nextx.shape and nextx.ndim == 2):
        return False
    try:
        _device.DeviceGraph().parse_nextxy(nextx, nexty, check=True)
    except ValueError as err:
        if getattr(err, "status", None) == _lib.ERR_INVALID_D8:
            return False
        if getattr(err, "status", None) == _lib.ERR_UNSUPPORTED:
            return True  # valid NEXTXY data that this library cannot hold
        raise
    return True


def read_nextxy(fn, nrow, ncol, bbox):
    """Read nextxy data from a CaMa-Flood binary file -> (int32 [2, nrow, ncol], Affine)"""
    data = np.fromfile(fn, "i4").reshape(2, nrow, ncol)
    assert len(bbox) == 4, "Bounding box should contain 4 coordinates."
    transform = gis_utils.transform_from_bounds(*bbox, ncol, nrow)
    return data, transform
