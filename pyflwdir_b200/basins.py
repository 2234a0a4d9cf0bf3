"""Basin delineation on the GPU; mirrors /root/reference/pyflwdir/basins.py:12-18."""
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
