"""Minimal stand-in for the `affine` package (not installed, no network) so that the real reference can be
imported in the build container to generate golden vectors. Tuple subclass because the reference passes
`gis_utils.IDENTITY` into njit functions and indexes it (pyflwdir/gis_utils.py:13, streams.py:79,117).
TEST INFRASTRUCTURE ONLY."""
from collections import namedtuple

import numpy as np

_Base = namedtuple("Affine", "a b c d e f g h i")


class Affine(_Base):
    def __new__(cls, a, b, c, d, e, f, g=0.0, h=0.0, i=1.0):
        return super().__new__(cls, float(a), float(b), float(c), float(d), float(e), float(f), g, h, i)

    @classmethod
    def identity(cls):
        return cls(1, 0, 0, 0, 1, 0)

    @classmethod
    def translation(cls, xoff, yoff):
        return cls(1, 0, xoff, 0, 1, yoff)

    @classmethod
    def scale(cls, *scaling):
        sx, sy = (scaling[0], scaling[0]) if len(scaling) == 1 else scaling
        return cls(sx, 0, 0, 0, sy, 0)

    @property
    def xoff(self):
        return self.c

    @property
    def yoff(self):
        return self.f

    @property
    def determinant(self):
        return self.a * self.e - self.b * self.d

    def __mul__(self, other):
        if isinstance(other, Affine):
            sa, sb, sc, sd, se, sf = self[:6]
            oa, ob, oc, od, oe, of = other[:6]
            return Affine(sa * oa + sb * od, sa * ob + sb * oe, sa * oc + sb * of + sc,
                          sd * oa + se * od, sd * ob + se * oe, sd * oc + se * of + sf)
        x, y = other
        x, y = np.asarray(x) if not np.isscalar(x) else x, np.asarray(y) if not np.isscalar(y) else y
        return (x * self.a + y * self.b + self.c, x * self.d + y * self.e + self.f)

    def __invert__(self):
        det = self.determinant
        ra, rb, rd, re = self.e / det, -self.b / det, -self.d / det, self.a / det
        return Affine(ra, rb, -self.c * ra - self.f * rb, rd, re, -self.c * rd - self.f * re)
