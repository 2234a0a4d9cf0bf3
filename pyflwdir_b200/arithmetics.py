"""Window / neighbourhood arithmetics over the flow network on the GPU; mirrors /root/reference/pyflwdir/arithmetics.py
(moving_average :67-103, moving_median :106-147, upstream_sum :150-169). `shape=` / `ncol=` are optional extensions."""
import numpy as np

from . import _functional


def moving_average(data, weights, n, idxs_ds, idxs_us_main, strord=None, nodata=-9999.0, mv=-1, shape=None, ncol=None):
    """Take the moving weighted average over the flow direction network"""
    g = _functional.graph(idxs_ds, shape, ncol)
    return g.moving_average(np.asarray(data).ravel(), weights, n, idxs_us_main, strord=strord, nodata=nodata)


def moving_median(data, n, idxs_ds, idxs_us_main, strord=None, nodata=-9999.0, mv=-1, shape=None, ncol=None):
    """Take the moving median over the flow direction network"""
    g = _functional.graph(idxs_ds, shape, ncol)
    return g.moving_median(np.asarray(data).ravel(), n, idxs_us_main, strord=strord, nodata=nodata)


def upstream_sum(idxs_ds, data, nodata=-9999.0, mv=-1, shape=None, ncol=None):
    """Returns sum of first upstream values"""
    return _functional.graph(idxs_ds, shape, ncol).upstream_sum(np.asarray(data).ravel(), nodata)
