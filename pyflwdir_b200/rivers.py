"""River attributes on the GPU; mirrors /root/reference/pyflwdir/rivers.py (classify_estuary :11-53)."""
import numpy as np

from . import _functional


def classify_estuary(idxs_ds, seq, idxs_pit, rivdst, rivwth, elevtn, max_elevtn=0, min_convergence=1e-2, shape=None, ncol=None):
    """Classifies estuaries based on width convergence -> int8 (>= 1 where estuary, 2 at its upstream end)"""
    g = _functional.graph(idxs_ds, shape, ncol)
    _functional.check_seq(g, seq, "classify_estuary")
    est = np.zeros(g.size, np.int8)
    pits = np.asarray(idxs_pit)
    est[pits[np.asarray(elevtn).ravel()[pits] <= max_elevtn]] = 1
    return g.classify_estuary(est, np.asarray(rivdst).ravel(), np.asarray(rivwth).ravel(), min_convergence)
