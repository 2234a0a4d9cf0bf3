"""d8 <-> ldd code remapping (host side, a 256-entry table); mirrors /root/reference/pyflwdir/core_conversion.py:11-28."""
import numpy as np

from . import core_d8, core_ldd

__all__ = ["d8_to_ldd", "ldd_to_d8"]


def _table(src, dst, extra, default):
    lut = np.full(256, default, dtype=np.uint8)
    lut[src.ravel()] = dst.ravel()
    for k, v in extra.items():
        lut[int(k)] = v
    return lut


def d8_to_ldd(flwdir):
    """Return ldd based on d8 array."""
    lut = _table(core_d8._ds, core_ldd._ds, {core_d8._pv[1]: core_ldd._pv, core_d8._mv: core_ldd._mv}, core_ldd._mv)
    return lut[np.asarray(flwdir, dtype=np.uint8)]


def ldd_to_d8(flwdir):
    """Return d8 based on ldd array."""
    lut = _table(core_ldd._ds, core_d8._ds, {core_ldd._pv: core_d8._pv[0], core_ldd._mv: core_d8._mv}, core_d8._mv)
    return lut[np.asarray(flwdir, dtype=np.uint8)]
