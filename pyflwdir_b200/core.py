"""Format-agnostic graph kernels on the GPU; mirrors /root/reference/pyflwdir/core.py (rank :17-47,
upstream_count :50-61, idxs_seq :87-117, fillnodata_upstream :120-146, pit_indices :225-232, path :400-438,
snap :441-480, inflow_idxs / outflow_idxs :483-514).
`shape=` / `ncol=` are optional extensions (the raster width is inferred from the links otherwise)."""
import numpy as np

from . import _functional, _lib

_mv = np.intp(-1)


def rank(idxs_ds, mv=_mv, shape=None, ncol=None):
    """Returns the rank, i.e. the distance counted in number of cells from the outlet -> (ranks int32, n)"""
    g = _functional.graph(idxs_ds, shape, ncol)
    ranks = g.fetch(_lib.ARR_RANK)
    return ranks, int(np.count_nonzero(ranks >= 0))


def upstream_count(idxs_ds, mv=_mv, mask=None, shape=None, ncol=None):
    """Returns array with number of upstream cells per cell (int8, -9 on nodata); with `mask`, only the upstream
    cells inside the mask count."""
    return _functional.graph(idxs_ds, shape, ncol).upstream_count(mask)


def upstream_matrix(idxs_ds, mv=_mv, shape=None, ncol=None):
    """Returns a 2D array with upstream cell indices for each cell, shape (idxs_ds.size, max number of upstream cells per
    cell), rows in ascending upstream index, padded with mv (core.py:67-84)."""
    idxs_ds = np.asarray(idxs_ds)
    g = _functional.graph(idxs_ds, shape=shape, ncol=ncol)
    dt = idxs_ds.dtype
    out = g.upstream_matrix(np.int64 if dt.itemsize == 8 else np.uint32 if dt == np.uint32 else np.int32)
    return out.astype(dt, copy=False) if out.dtype != dt else out


def main_upstream(idxs_ds, uparea, upa_min=0.0, mv=_mv, shape=None, ncol=None):
    """Returns the index of the upstream cell with the largest uparea, mv (-1) at headwaters (core.py:191-219)."""
    dt = np.asarray(idxs_ds).dtype
    return _functional.graph(idxs_ds, shape, ncol).main_upstream(np.asarray(uparea).ravel(), upa_min, dt)


def idxs_seq(idxs_ds, idxs_pit, mv=_mv, shape=None, ncol=None):
    """Returns indices ordered from down- to upstream ("walk": BFS from the pits, core.py:87-117)."""
    g = _functional.graph(idxs_ds, shape, ncol)
    dt = np.asarray(idxs_ds).dtype
    pits = g.fetch(_lib.ARR_PITS, np.int64)
    if not np.array_equal(np.asarray(idxs_pit).astype(np.int64), pits):
        raise NotImplementedError("idxs_seq from a subset / permutation of the pits is outside the accelerated hot path")
    fetch_dt = np.dtype(np.int64) if dt == np.uint64 else dt
    return g.fetch(_lib.ARR_SEQ, fetch_dt).astype(dt, copy=False)


def pit_indices(idxs_ds, shape=None, ncol=None):
    """Returns pit indices, i.e. cells with no downstream cell"""
    dt = np.asarray(idxs_ds).dtype
    fetch_dt = np.dtype(np.int64) if dt == np.uint64 else dt
    return _functional.graph(idxs_ds, shape, ncol).fetch(_lib.ARR_PITS, fetch_dt).astype(dt, copy=False)


def fillnodata_upstream(idxs_ds, seq, data, nodata, shape=None, ncol=None):
    """Copy of <data> where upstream cells with <nodata> are filled with the first downstream valid value
    (core.py:120-146)."""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "fillnodata_upstream")
    return g.fillnodata(np.asarray(data).ravel(), nodata, "up")


def fillnodata_downstream(idxs_ds, seq, data, nodata, how="max", shape=None, ncol=None):
    """Copy of <data> where downstream cells with <nodata> are filled from their upstream cells, merged at
    confluences with how = "max" | "min" | "sum" (core.py:149-188)."""
    assert how in ["min", "max", "sum"]
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "fillnodata_downstream", order_sensitive=how == "sum" and np.asarray(data).dtype.kind == "f")
    return g.fillnodata(np.asarray(data).ravel(), nodata, "down", how)


def _trace(idxs0, idxs_nxt, ncol, mask, max_length, real_length, latlon, transform, shape, paths):
    from . import gis_utils as gis

    idxs_nxt = np.asarray(idxs_nxt)
    n = idxs_nxt.size
    # idxs_nxt is idxs_ds (downstream trace) or idxs_us_main (upstream trace): the former holds self-links at the pits
    own = np.arange(n, dtype=np.int64)
    nx = idxs_nxt.astype(np.int64)
    down = bool(np.any(nx == own))
    if not down:
        raise NotImplementedError("core.path / core.snap along idxs_us_main need the downstream links as well: use "
                                  "FlwdirRaster.path(direction='up') / snap(direction='up')")
    if shape is None and ncol is not None:
        shape = (n // int(ncol), int(ncol))
    g = _functional.graph(idxs_nxt, shape, None)
    hop = gis.hop_length_table(g.shape[0], transform, latlon, dtype=np.float64) if (real_length and ncol is not None) else None
    return g.trace(np.atleast_1d(idxs0), "down", None, mask, max_length, hop, paths_dtype=idxs_nxt.dtype if paths else None)


def path(idxs0, idxs_nxt, ncol=None, mask=None, max_length=None, real_length=False, latlon=False,
         transform=(1.0, 0.0, 0.0, 0.0, -1.0, 0.0), mv=_mv, shape=None):
    """Traces from every start cell along idxs_nxt -> (list of index arrays, float64 distances); core.py:400-438"""
    paths, _, dist = _trace(idxs0, idxs_nxt, ncol, mask, max_length, real_length, latlon, transform, shape, True)
    return paths, dist


def snap(idxs0, idxs_nxt, ncol=None, mask=None, max_length=None, real_length=False, latlon=False,
         transform=(1.0, 0.0, 0.0, 0.0, -1.0, 0.0), mv=_mv, shape=None):
    """Last cell of every trace -> (indices in the dtype of idxs0, float32 distances); core.py:441-480"""
    idxs0 = np.atleast_1d(idxs0)
    _, ends, dist = _trace(idxs0, idxs_nxt, ncol, mask, max_length, real_length, latlon, transform, shape, False)
    return ends.astype(idxs0.dtype), dist.astype(np.float32)


def inflow_idxs(idxs_ds, seq, region, shape=None, ncol=None):
    """returns linear indices of most upstream cells within region (core.py:483-497)"""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "inflow_idxs", order_sensitive=True)  # the indices come back in sequence order
    return g.inflow_idxs(np.asarray(region).ravel(), np.asarray(idxs_ds).dtype)


def outflow_idxs(idxs_ds, seq, region, shape=None, ncol=None):
    """returns linear indices of most downstream cells within region (core.py:500-514)"""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "outflow_idxs", order_sensitive=True)
    return g.outflow_idxs(np.asarray(region).ravel(), np.asarray(idxs_ds).dtype)


def loop_indices(idxs_ds, mv=_mv, shape=None, ncol=None):
    """Returns indices of loop cells, i.e. cells which do not drain to a pit (core.py:235-243)"""
    ranks = rank(idxs_ds, mv, shape, ncol)[0]
    return np.flatnonzero(ranks == -1).astype(np.asarray(idxs_ds).dtype)


def headwater_indices(idxs_ds, mask=None, mv=_mv, shape=None, ncol=None):
    """Returns indices of headwater cells, i.e. cells with no upstream neighbors (core.py:246-250)"""
    nup = upstream_count(idxs_ds, mv, mask, shape, ncol)
    return np.where(nup == 0)[0].astype(np.asarray(idxs_ds).dtype)


def confluence_indices(idxs_ds, mask=None, mv=_mv, shape=None, ncol=None):
    """Returns indices of confluence cells, i.e. cells with two or more upstream neighbors (core.py:253-257)"""
    nup = upstream_count(idxs_ds, mv, mask, shape, ncol)
    return np.where(nup > 1)[0].astype(np.asarray(idxs_ds).dtype)
