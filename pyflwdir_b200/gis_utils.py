"""Host-side georeferencing helpers needed by the hot path's wrappers.

Only what `FlwdirRaster.upstream_area(unit != "cell")` and `set_transform` touch: an `Affine` value type (`affine.Affine` when that
package is installed, else an equivalent 9-tuple defined here), `IDENTITY`, `AREA_FACTORS`, and the per-row cell-area grid.
Mirrors /root/reference/pyflwdir/gis_utils.py:10-13 (constants), :340-358 (pixel-centre coordinates),
:379-402 (reggrid_area / area_grid), :405-412 (cellarea). The area grid is an O(nrow) host computation that
only produces the `data` vector of an accumulation; it stays on the host.
"""
import math
from collections import namedtuple

import numpy as np

_R = 6371e3  # earth radius [m], gis_utils.py:10
AREA_FACTORS = {"m2": 1.0, "ha": 1e4, "km2": 1e6, "cell": 1}  # gis_utils.py:11

_AffineBase = namedtuple("Affine", "a b c d e f g h i")


class _OwnAffine(_AffineBase):
    """2-D affine transform (x, y) = A * (col, row); a 9-tuple like `affine.Affine` (indexable, iterable)."""

    __slots__ = ()

    def __new__(cls, a, b, c, d, e, f, g=0.0, h=0.0, i=1.0):
        vals = [float(v) for v in (a, b, c, d, e, f)]
        return super().__new__(cls, *vals, float(g), float(h), float(i))

    @classmethod
    def identity(cls):
        return cls(1.0, 0.0, 0.0, 0.0, 1.0, 0.0)

    @classmethod
    def translation(cls, xoff, yoff):
        return cls(1.0, 0.0, xoff, 0.0, 1.0, yoff)

    @classmethod
    def scale(cls, sx, sy=None):
        return cls(sx, 0.0, 0.0, 0.0, sx if sy is None else sy, 0.0)

    @property
    def xoff(self):
        return self.c

    @property
    def yoff(self):
        return self.f

    def __mul__(self, other):
        if isinstance(other, _OwnAffine):
            return _OwnAffine(
                self.a * other.a + self.b * other.d,
                self.a * other.b + self.b * other.e,
                self.a * other.c + self.b * other.f + self.c,
                self.d * other.a + self.e * other.d,
                self.d * other.b + self.e * other.e,
                self.d * other.c + self.e * other.f + self.f,
            )
        x, y = other
        return (x * self.a + y * self.b + self.c, x * self.d + y * self.e + self.f)

    def __invert__(self):
        det = self.a * self.e - self.b * self.d
        if det == 0:
            raise ValueError("Affine transform is not invertible")
        ia, ib, id_, ie = self.e / det, -self.b / det, -self.d / det, self.a / det
        return _OwnAffine(ia, ib, -self.c * ia - self.f * ib, id_, ie, -self.c * id_ - self.f * ie)


try:  # the reference's own dependency, when it is installed: transforms then are the very type callers already hold
    from affine import Affine
except ImportError:  # not a dependency here
    Affine = _OwnAffine

# N->S oriented identity, gis_utils.py:13
IDENTITY = Affine(1.0, 0.0, 0.0, 0.0, -1.0, 0.0)


def transform_from_bounds(west, south, east, north, width, height):
    """Affine transformation of a georeferenced raster given its bounds and its size in pixels; gis_utils.py:162-170."""
    return Affine.translation(west, north) * Affine.scale((east - west) / width, (south - north) / height)


def affine_to_coords(affine, shape):
    """Pixel-centre x (per column) and y (per row) coordinates; gis_utils.py:340-358."""
    height, width = shape
    xs, _ = affine * (np.arange(width) + 0.5, np.zeros(width) + 0.5)
    _, ys = affine * (np.zeros(height) + 0.5, np.arange(height) + 0.5)
    return xs, ys


def cellarea(lat, xres, yres):
    """Area [m2] of a (xres x yres) degree cell centred at latitude `lat`; gis_utils.py:405-412.

    The reference evaluates this inside numba (libm `sin`); `math.sin` is the same libm call, whereas numpy's
    vectorised `np.sin` may differ in the last bit, so rows are evaluated one by one."""
    lat = np.atleast_1d(np.asarray(lat, dtype=np.float64))
    out = np.empty(lat.shape, dtype=np.float64)
    half = abs(yres) / 2.0
    dx = math.radians(abs(xres))
    for k, la in enumerate(lat.tolist()):
        l1 = math.radians(la - half)
        l2 = math.radians(la + half)
        out[k] = _R**2 * dx * (math.sin(l2) - math.sin(l1))
    return out


def reggrid_area(lats, lons):
    """Cell-area grid [m2] of a regular lat/lon grid; gis_utils.py:379-385 (float64 result)."""
    xres = np.abs(np.mean(np.diff(lons)))
    yres = np.abs(np.mean(np.diff(lats)))
    ones = np.ones((lats.size, lons.size), dtype=np.float32)
    return cellarea(lats, xres, yres)[:, None] * ones


def area_grid(transform, shape, latlon=False, unit="m2"):
    """Regular grid of cell areas; gis_utils.py:388-402 (int32 ones for "cell", float64 if latlon, else float32)."""
    unit = str(unit).lower()
    if unit not in AREA_FACTORS:
        fstr = '", "'.join(AREA_FACTORS.keys())
        raise ValueError(f'Unknown unit: {unit}, select from "{fstr}".')
    if unit == "cell":
        return np.ones(shape, dtype=np.int32)
    if latlon:
        lon, lat = affine_to_coords(transform, shape)
        return reggrid_area(lat, lon) / AREA_FACTORS[unit]
    area0 = abs(transform[0] * transform[4]) / AREA_FACTORS[unit]
    return np.full(shape, area0, dtype=np.float32)


def degree_metres_y(lat):
    """Vertical length of a degree in metres at a given latitude; gis_utils.py:415-430 (scalar, libm cos)."""
    radlat = math.radians(lat)
    return 111132.92 + (-559.82 * math.cos(2.0 * radlat)) + (1.175 * math.cos(4.0 * radlat)) + (-0.0023 * math.cos(6.0 * radlat))


def degree_metres_x(lat):
    """Horizontal length of a degree in metres at a given latitude; gis_utils.py:433-447."""
    radlat = math.radians(lat)
    return (111412.84 * math.cos(radlat)) + (-93.5 * math.cos(3.0 * radlat)) + (0.118 * math.cos(5.0 * radlat))


_libm = None


def _libm_hypot(x, y):
    """hypot of the C library: what numba's math.hypot lowers to (numba/cpython/mathimpl.py hypot_float_impl). CPython's
    own math.hypot is a different (correctly rounded) algorithm and differs from glibc's in the last bit for ~0.2 % of
    the arguments."""
    global _libm
    if _libm is None:
        import ctypes
        import ctypes.util

        lib = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        lib.hypot.restype = ctypes.c_double
        lib.hypot.argtypes = [ctypes.c_double, ctypes.c_double]
        _libm = lib
    return _libm.hypot(x, y)


def hop_length_table(nrow, transform, latlon, dtype=np.float32):
    """Table [nrow, 3, 2] of gis_utils.distance(idx0, idx1, ...) (gis_utils.py:451-486) for a hop that starts in
    row r0 with row delta dr in (-1, 0, 1) and |column delta| dc in (0, 1) -- everything the distance depends on.
    float32: for streams.stream_distance, interpreted Python in the reference (CPython's math.hypot, the sum kept in
    float32); float64: for core._trace, compiled by numba (the C library's hypot). Reproduces the reference including
    its quirk for projected rasters (dy = xres, dx = yres)."""
    hyp = math.hypot if np.dtype(dtype) == np.float32 else _libm_hypot
    xres, yres, north = transform[0], transform[4], transform[5]
    tab = np.zeros((nrow, 3, 2), dtype=np.float64)
    if not latlon:
        for j, dr in enumerate((-1, 0, 1)):
            for dc in (0, 1):
                tab[:, j, dc] = hyp(float(xres) * abs(dr), float(yres) * dc)
    else:
        for r0 in range(nrow):
            for j, d in enumerate((-1, 0, 1)):
                lat = north + (r0 + (r0 + d)) / 2.0 * yres
                dr = abs(d)
                dy = 0.0 if dr == 0 else degree_metres_y(lat) * yres
                dx1 = degree_metres_x(lat) * xres
                tab[r0, j, 0] = hyp(dy * dr, 0.0)
                tab[r0, j, 1] = hyp(dy * dr, dx1 * 1)
    return tab.astype(dtype)


def xy(transform, rows, cols, offset="center"):
    """x and y coordinates of pixels at `rows` and `cols`; gis_utils.py:183-223."""
    rows, cols = np.asarray(rows), np.asarray(cols)
    offs = {"center": (0.5, 0.5), "ul": (0, 0), "ur": (1, 0), "ll": (0, 1), "lr": (1, 1)}
    if offset not in offs:
        raise ValueError("Invalid offset")
    coff, roff = offs[offset]
    xs, ys = transform * transform.translation(coff, roff) * (cols, rows)
    return xs, ys


def idxs_to_coords(idxs, transform, shape, offset="center"):
    """Coordinates of linear raster indices; gis_utils.py:264-298."""
    idxs = np.asarray(idxs).astype(int)
    size = np.multiply(*shape)
    if np.any(np.logical_or(idxs < 0, idxs >= size)):
        raise IndexError("idxs coordinates outside domain")
    ncol = shape[1]
    return xy(transform, idxs // ncol, idxs % ncol, offset=offset)


def features(flowpaths, xs=None, ys=None, transform=None, shape=None, **kwargs):
    """LineString geo-feature (dict) per flow path; gis_utils.py:490-549. Host side, like the reference: the output is a
    list of Python dicts."""
    if xs is None or ys is None:
        if transform is None or shape is None:
            raise ValueError("transform and shape should be provided if xs and ys are None")
        _size = shape[0] * shape[1]
    else:
        _size = xs.size
    for key in kwargs:
        if not isinstance(kwargs[key], np.ndarray) or kwargs[key].size != _size:
            raise ValueError(f'Kwargs map "{key}" should be ndarrays of same size as coordinates')
    # coordinates of all paths in one vectorised pass (element-wise identical to the reference's per-path calls)
    paths = [np.asarray(idxs) for idxs in flowpaths if len(idxs) >= 2]
    feats = list()
    if not paths:
        return feats
    flat = np.concatenate(paths)
    if xs is None or ys is None:
        xi, yi = idxs_to_coords(flat, transform, shape)
    else:
        xi, yi = np.asarray(xs).ravel()[flat], np.asarray(ys).ravel()[flat]
    coords = list(zip(xi, yi))  # numpy scalars, like the reference
    o = 0
    for idxs in paths:
        n = len(idxs)
        idx0 = idxs[0]
        pit = idxs[-1] == idxs[-2]
        props = {key: kwargs[key].flat[idx0] for key in kwargs}
        feats.append({
            "type": "Feature",
            "geometry": {"type": "LineString", "coordinates": coords[o:o + n]},
            "properties": {"idx": idx0, "idx_ds": idxs[-1], "pit": pit, **props},
        })
        o += n
    return feats
