"""Basin delineation on the GPU; mirrors /root/reference/pyflwdir/basins.py (basins :12-18, interbasin_mask :23-64,
subbasins_streamorder :67-103, subbasins_pfafstetter :106-191, subbasins_area :194-233)."""
import numpy as np

from . import _functional, _lib


def basins(idxs_ds, idxs_pit, seq, ids=None, shape=None, ncol=None):
    """Return basin map"""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "basins")
    idxs_pit = np.asarray(idxs_pit)
    if ids is None:
        if np.array_equal(idxs_pit.astype(np.int64), g.fetch(_lib.ARR_PITS, np.int64)):
            return g.basins()  # all pits, ids 1..npits: the tile solver
        ids = np.arange(1, idxs_pit.size + 1, dtype=np.uint32)
    return g.basins(idxs_pit, np.asarray(ids))


def interbasin_mask(idxs_ds, seq, region, stream=None, shape=None, ncol=None):
    """Returns most downstream contiguous area within region (basins.py:23-64)"""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "interbasin_mask")
    return g.interbasin_mask(np.asarray(region).ravel(), None if stream is None else np.asarray(stream).ravel())


def subbasins_streamorder(idxs_ds, seq, strord, mask=None, min_sto=-2, shape=None, ncol=None):
    """Returns map with basin IDs, with a basin ID for subbasins of each stream order (basins.py:67-103)"""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "subbasins_streamorder", order_sensitive=True)  # labels are numbered in sequence order
    return g.subbasins_streamorder(np.asarray(strord).ravel(), None if mask is None else np.asarray(mask).ravel(), min_sto,
                                   np.asarray(idxs_ds).dtype)


def subbasins_pfafstetter(idxs_pit, idxs_ds, seq, idxs_us_main, uparea, mask=None, depth=1, mv=-1, shape=None, ncol=None):
    """Returns the pfafstetter subbasin map and outlet indices (basins.py:106-191)"""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "subbasins_pfafstetter", order_sensitive=True)
    if not np.array_equal(np.asarray(idxs_pit).astype(np.int64), g.fetch(_lib.ARR_PITS, np.int64)):
        raise NotImplementedError("subbasins_pfafstetter from a subset of the pits is outside the accelerated hot path")
    return g.subbasins_pfafstetter(idxs_us_main, np.asarray(uparea).ravel(), None if mask is None else np.asarray(mask).ravel(),
                                   depth, np.asarray(idxs_ds).dtype)


def subbasins_area(idxs_ds, seq, idxs_us_main, uparea, area_min, shape=None, ncol=None):
    """Returns map with basin IDs, with a minimal area of `area_min` (basins.py:194-233)"""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "subbasins_area", order_sensitive=True)
    return g.subbasins_area(idxs_us_main, np.asarray(uparea).ravel(), area_min, np.asarray(idxs_ds).dtype)
