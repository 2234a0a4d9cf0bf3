"""Sweeps over the flow sequence on the GPU; mirrors /root/reference/pyflwdir/streams.py (accuflux :15-41,
accuflux_ds :44-70, stream_order :191-225, strahler_order :228-269). The device always sweeps its own "walk" sequence (the one
`core.idxs_seq` returns); `seq` is only checked for covering the same cells."""
import numpy as np

from . import _functional


def accuflux(idxs_ds, seq, data, nodata, shape=None, ncol=None):
    """Returns maps of accumulate upstream <data>"""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "accuflux")
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
    _functional.check_seq(g, seq, "streams")
    return g.streams(mask, max_len, np.asarray(idxs_ds).dtype)
