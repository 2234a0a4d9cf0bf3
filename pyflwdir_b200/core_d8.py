"""D8 codec on the GPU; mirrors /root/reference/pyflwdir/core_d8.py (from_array :42-67, to_array :86-102,
isvalid :105-112) with the same signatures."""
import numpy as np

from . import _device, _functional, _lib

_ftype = "d8"
_ds = np.array([[32, 64, 128], [16, 0, 1], [8, 4, 2]], dtype=np.uint8)
_us = np.array([[2, 4, 8], [1, 0, 16], [128, 64, 32]], dtype=np.uint8)
_mv = np.uint8(247)
_pv = np.array([0, 255], dtype=np.uint8)
_all = np.array([32, 64, 128, 16, 0, 1, 8, 4, 2, 247, 255], dtype=np.uint8)


def from_array(flwdir, _mv=_mv, dtype=np.intp, device=0):
    """convert 2D D8 data to 1D next downstream indices -> (idxs_ds, idxs_pit, n)"""
    dt = np.dtype(dtype)
    fetch_dt = np.dtype(np.int64) if dt == np.uint64 else dt
    g = _device.DeviceGraph(device)
    idxs_ds = g.parse_d8(flwdir, idx_dtype=fetch_dt, want_idxs=True)
    pits = g.fetch(_lib.ARR_PITS, fetch_dt)
    if dt == np.uint64:  # the reference's uint64 fixtures: same values, mv = 2**64-1
        idxs_ds, pits = idxs_ds.astype(np.uint64), pits.astype(np.uint64)
    return idxs_ds, pits, int(g.n_valid)


def to_array(idxs_ds, shape, mv=None, device=0):
    """convert downstream linear indices to dense D8 raster"""
    g = _functional.graph(idxs_ds, shape=shape, device=device)
    return g.fetch(_lib.ARR_D8).reshape(shape)


def isvalid(flwdir, _all=_all, device=0):
    """True if 2D D8 raster is valid"""
    if not (isinstance(flwdir, np.ndarray) and flwdir.dtype == "uint8" and flwdir.ndim == 2):
        return False
    try:
        _device.DeviceGraph(device).parse_d8(flwdir)
    except ValueError as err:
        if getattr(err, "status", None) == _lib.ERR_INVALID_D8:
            return False
        raise
    return True


# ---- scalar helpers on a handful of cells (host; not part of the hot path) ------------------------------------
def drdc(dd):
    """D8 code -> (row offset, column offset); pits and anything else map to (0, 0) like core_d8.py:22-39 for the legal codes"""
    hit = np.argwhere(_ds == np.uint8(dd))
    if dd in (0, 255) or hit.size == 0:
        return 0, 0
    return int(hit[0, 0]) - 1, int(hit[0, 1]) - 1


def ispit(dd):
    """True if D8 pit (core_d8.py:125-127)"""
    return np.logical_or(np.asarray(dd) == _pv[0], np.asarray(dd) == _pv[1])


def isnodata(dd):
    """True if D8 nodata (core_d8.py:130-132)"""
    return np.asarray(dd) == _mv


def _downstream_idx(idx0, flwdir_flat, shape, mv=np.intp(-1)):
    """linear index of the downstream cell of idx0; mv when it lies outside the raster (core_d8.py:70-83)"""
    nrow, ncol = shape
    r0, c0 = idx0 // ncol, idx0 % ncol
    dr, dc = drdc(flwdir_flat[idx0])
    r, c = r0 + dr, c0 + dc
    if 0 <= r < nrow and 0 <= c < ncol:
        return np.intp(r * ncol + c)
    return mv


def _upstream_idx(idx0, flwdir_flat, shape, dtype=np.intp):
    """linear indices of the cells that drain into idx0 (core_d8.py:137-154)"""
    nrow, ncol = shape
    r0, c0 = idx0 // ncol, idx0 % ncol
    out = []
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            r, c = r0 + dr, c0 + dc
            if (dr or dc) and 0 <= r < nrow and 0 <= c < ncol and flwdir_flat[r * ncol + c] == _us[dr + 1, dc + 1]:
                out.append(r * ncol + c)
    return np.array(out, dtype=dtype)
