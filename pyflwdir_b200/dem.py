"""Elevation-coupled sweep on the GPU; mirrors /root/reference/pyflwdir/dem.py:299-330."""
import numpy as np

from . import _functional


def height_above_nearest_drain(idxs_ds, seq, drain, elevtn, shape=None, ncol=None):
    """Returns the height above the nearest drain (HAND), float64, -9999 outside the sequence."""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "height_above_nearest_drain")
    return g.hand(np.asarray(drain).ravel(), np.asarray(elevtn).ravel())
