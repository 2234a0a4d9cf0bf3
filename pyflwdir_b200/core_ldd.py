"""PCRaster LDD codec on the GPU; mirrors /root/reference/pyflwdir/core_ldd.py (from_array :41-66, to_array :85-102,
isvalid :105-107) with the same signatures. Same parse kernel as D8 with the LDD code table."""
import numpy as np

from . import _device, _functional, _lib

_ftype = "ldd"
_ds = np.array([[7, 8, 9], [4, 5, 6], [1, 2, 3]], dtype=np.uint8)
_us = np.array([[3, 2, 1], [6, 5, 4], [9, 8, 7]], dtype=np.uint8)
_mv = np.uint8(255)
_pv = np.uint8(5)
_all = np.array([7, 8, 9, 4, 5, 6, 1, 2, 3, 255], dtype=np.uint8)


def from_array(flwdir, _mv=_mv, dtype=np.intp, device=0):
    """convert 2D LDD data to 1D next downstream indices -> (idxs_ds, idxs_pit, n)"""
    dt = np.dtype(dtype)
    fetch_dt = np.dtype(np.int64) if dt == np.uint64 else dt
    g = _device.DeviceGraph(device)
    idxs_ds = g.parse_d8(flwdir, idx_dtype=fetch_dt, want_idxs=True, ftype="ldd")
    pits = g.fetch(_lib.ARR_PITS, fetch_dt)
    if dt == np.uint64:
        idxs_ds, pits = idxs_ds.astype(np.uint64), pits.astype(np.uint64)
    return idxs_ds, pits, int(g.n_valid)


def to_array(idxs_ds, shape, mv=None, device=0):
    """convert downstream linear indices to dense LDD raster"""
    g = _functional.graph(idxs_ds, shape=shape, device=device)
    return g.fetch(_lib.ARR_LDD).reshape(shape)


def isvalid(flwdir, _all=_all, device=0):
    """True if 2D LDD raster is valid"""
    if not (isinstance(flwdir, np.ndarray) and flwdir.dtype == "uint8" and flwdir.ndim == 2):
        return False
    try:
        _device.DeviceGraph(device).parse_d8(flwdir, ftype="ldd")
    except ValueError as err:
        if getattr(err, "status", None) == _lib.ERR_INVALID_D8:
            return False
        raise
    return True


# ---- scalar helpers on a handful of cells (host; not part of the hot path) ------------------------------------
def drdc(dd):
    """LDD code -> (row offset, column offset); the pit code 5 maps to (0, 0) (core_ldd.py:20-38)"""
    hit = np.argwhere(_ds == np.uint8(dd))
    if hit.size == 0:
        return 0, 0
    return int(hit[0, 0]) - 1, int(hit[0, 1]) - 1


def ispit(dd):
    """True if LDD pit"""
    return np.asarray(dd) == _pv


def isnodata(dd):
    """True if LDD nodata"""
    return np.asarray(dd) == _mv


def _downstream_idx(idx0, flwdir_flat, shape, mv=np.intp(-1)):
    """linear index of the downstream cell of idx0; mv when it lies outside the raster"""
    nrow, ncol = shape
    dr, dc = drdc(flwdir_flat[idx0])
    r, c = idx0 // ncol + dr, idx0 % ncol + dc
    if 0 <= r < nrow and 0 <= c < ncol:
        return np.intp(r * ncol + c)
    return mv


def _upstream_idx(idx0, flwdir_flat, shape, dtype=np.intp):
    """linear indices of the cells that drain into idx0"""
    nrow, ncol = shape
    r0, c0 = idx0 // ncol, idx0 % ncol
    out = []
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            r, c = r0 + dr, c0 + dc
            if (dr or dc) and 0 <= r < nrow and 0 <= c < ncol and flwdir_flat[r * ncol + c] == _us[dr + 1, dc + 1]:
                out.append(r * ncol + c)
    return np.array(out, dtype=dtype)
