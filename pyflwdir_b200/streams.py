"""Sweeps over the flow sequence on the GPU; mirrors /root/reference/pyflwdir/streams.py (accuflux :15-41,
accuflux_ds :44-70, stream_order :191-225, strahler_order :228-269). The device always sweeps its own "walk" sequence (the one
`core.idxs_seq` returns): `seq` must cover the same cells, and where the result depends on the order inside `seq` (float
sums, segment numbering) it must BE that sequence (`_functional.check_seq`)."""
import numpy as np

from . import _functional, _lib


def accuflux(idxs_ds, seq, data, nodata, shape=None, ncol=None):
    """Returns maps of accumulate upstream <data>"""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "accuflux", order_sensitive=np.asarray(data).dtype.kind == "f")  # float sums follow seq
    return g.accuflux(np.asarray(data).ravel(), nodata, "up")


def accuflux_ds(idxs_ds, seq, data, nodata, shape=None, ncol=None):
    """Returns maps of accumulate downstream <data>"""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "accuflux_ds")
    return g.accuflux(np.asarray(data).ravel(), nodata, "down")


def stream_order(idxs_ds, seq, idxs_us_main, mask=None, mv=-1, shape=None, ncol=None):
    """Returns the classic or Hack's "bottum up" stream order (uint8; streams.py:191-225)."""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "stream_order")
    return g.stream_order_classic(idxs_us_main, mask)


def strahler_order(idxs_ds, seq, mask=None, shape=None, ncol=None):
    """Returns the strahler "top down" stream order (uint8)."""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "strahler_order")
    return g.strahler(mask)


def streams(idxs_ds, seq, mask=None, max_len=0, mv=-1, shape=None, ncol=None):
    """Returns list of linear indices per stream of equal stream order (streams.py:131-188)."""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "streams", order_sensitive=True)  # segments are listed in sequence order
    return g.streams(mask, max_len, np.asarray(idxs_ds).dtype)


def upstream_area(idxs_ds, seq, ncol, latlon=False, transform=(1.0, 0.0, 0.0, 0.0, -1.0, 0.0), area_factor=1, nodata=-9999.0,
                  dtype=np.float64):
    """Returns the accumulated upstream area without an area grid in memory (streams.py:72-128): the local cell area
    (per row for a geographic CRS) is evaluated on the host, the accumulation runs on the device in `dtype`."""
    from . import gis_utils as gis

    idxs_ds = np.asarray(idxs_ds)
    nrow = idxs_ds.size // int(ncol)
    g = _functional.graph(idxs_ds, (nrow, int(ncol)), None)
    _functional.check_seq(g, seq, "upstream_area", order_sensitive=np.dtype(dtype).kind == "f")
    xres, yres, north = transform[0], transform[4], transform[5]
    uparea = np.full(idxs_ds.size, nodata, dtype=dtype)
    inseq = g.fetch(_lib.ARR_RANK) >= 0
    if latlon:
        lats = [north + (r + 0.5) * yres for r in range(nrow)]
        rows = gis.cellarea(np.asarray(lats, dtype=np.float64), xres, yres) / area_factor
        local = np.repeat(rows, int(ncol))
        uparea[inseq] = local[inseq]
    else:
        uparea[inseq] = abs(xres * yres) / area_factor
    return g.accuflux(uparea, nodata, "up")
