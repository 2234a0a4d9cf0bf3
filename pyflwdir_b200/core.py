"""Format-agnostic graph kernels on the GPU; mirrors /root/reference/pyflwdir/core.py (rank :17-47,
upstream_count :50-61, idxs_seq :87-117, fillnodata_upstream :120-146, pit_indices :225-232).
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
    _functional.check_seq(g, seq, "fillnodata_downstream")
    return g.fillnodata(np.asarray(data).ravel(), nodata, "down", how)
