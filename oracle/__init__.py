"""CPU oracle for the D8 flow-network hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes wrappers around ``oracle/libpfd_oracle.so`` (built from ``pfd_oracle.c`` by ``oracle/Makefile``)
exposing the reference's L1/L2 function signatures so parity tests read like the reference's own:

    oracle.core_d8.from_array / to_array / isvalid      (pyflwdir/core_d8.py:42-67,86-102,105-122)
    oracle.core.rank / upstream_count / idxs_seq / fillnodata_upstream / pit_indices
                                                          (pyflwdir/core.py:17-47,50-61,87-117,120-146,225-232)
    oracle.streams.accuflux / accuflux_ds / strahler_order (pyflwdir/streams.py:15-41,44-70,228-269)
    oracle.basins.basins                                 (pyflwdir/basins.py:12-18)
    oracle.dem.height_above_nearest_drain                (pyflwdir/dem.py:299-330)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this package. ``pyflwdir_b200`` never does. Parity of this oracle with the real reference is pinned by
``tests/golden`` (see ``tests/golden/make_golden.py``).
"""
import ctypes as C
import os
import subprocess
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpfd_oracle.so")
_lib = None

_SFX = {np.dtype(np.int32): "i32", np.dtype(np.uint32): "u32", np.dtype(np.int64): "i64"}
_DATA_SFX = {
    np.dtype(np.int8): "i8",
    np.dtype(np.uint8): "u8",
    np.dtype(np.int16): "i16",
    np.dtype(np.uint16): "u16",
    np.dtype(np.int32): "i32",
    np.dtype(np.uint32): "u32",
    np.dtype(np.int64): "i64",
    np.dtype(np.float32): "f32",
    np.dtype(np.float64): "f64",
}


def build(force=False):
    """Compile the oracle shared library (gcc). Building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _idx(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.dtype(np.intp) and a.dtype not in _SFX:  # pragma: no cover
        a = a.astype(np.int64)
    if a.dtype == np.uint64:
        # the reference's uint64 fixtures (tests/conftest.py:88-108) -> same values as int64, mv -> -1
        a = a.astype(np.int64)
    if a.dtype not in _SFX:
        raise TypeError(f"unsupported index dtype {a.dtype}")
    return a, _SFX[a.dtype]


def _fn(name, sfx, restype=None):
    f = getattr(lib(), f"{name}_{sfx}")
    f.restype = restype
    return f


# ----------------------------------------------------------------------------- core_d8
def _d8_from_array(flwdir, _mv=np.uint8(247), dtype=np.intp):
    flwdir = np.ascontiguousarray(flwdir, dtype=np.uint8)
    nrow, ncol = flwdir.shape
    dt = np.dtype(dtype)
    if dt == np.uint64:
        dt = np.dtype(np.int64)
    sfx = _SFX[dt]
    idxs_ds = np.empty(flwdir.size, dtype=dt)
    pits = np.empty(flwdir.size, dtype=dt)
    npits = C.c_int64(0)
    n = _fn("orc_d8_from_array", sfx, C.c_int64)(
        _p(flwdir), C.c_int64(nrow), C.c_int64(ncol), _p(idxs_ds), _p(pits), C.byref(npits)
    )
    out_pits = pits[: npits.value].copy()
    if np.dtype(dtype) == np.uint64:
        return idxs_ds.astype(np.uint64), out_pits.astype(np.uint64), int(n)
    return idxs_ds, out_pits, int(n)


def _d8_to_array(idxs_ds, shape, mv=None):
    a, sfx = _idx(idxs_ds)
    out = np.empty(a.size, dtype=np.uint8)
    rc = _fn("orc_d8_to_array", sfx, C.c_int)(_p(a), C.c_int64(a.size), C.c_int64(shape[1]), _p(out))
    if rc != 0:
        raise ValueError("Invalid data downstream index outside 8 neighbors.")
    return out.reshape(shape)


def _d8_isvalid(flwdir):
    if not (isinstance(flwdir, np.ndarray) and flwdir.dtype == np.uint8 and flwdir.ndim == 2):
        return False
    f = lib().orc_d8_check_values
    f.restype = C.c_int
    a = np.ascontiguousarray(flwdir)
    return bool(f(_p(a), C.c_int64(a.size)))


def _drdc_table():
    dr = np.empty(256, np.int8)
    dc = np.empty(256, np.int8)
    lib().orc_drdc_table(_p(dr), _p(dc))
    return dr, dc


core_d8 = types.SimpleNamespace(
    from_array=_d8_from_array, to_array=_d8_to_array, isvalid=_d8_isvalid, drdc_table=_drdc_table
)


def _ldd_from_array(flwdir, _mv=np.uint8(255), dtype=np.intp):
    """pyflwdir/core_ldd.py:41-66"""
    flwdir = np.ascontiguousarray(flwdir, dtype=np.uint8)
    nrow, ncol = flwdir.shape
    dt = np.dtype(dtype)
    sfx = _SFX[dt]
    idxs_ds = np.empty(flwdir.size, dtype=dt)
    pits = np.empty(flwdir.size, dtype=dt)
    npits = C.c_int64(0)
    n = _fn("orc_ldd_from_array", sfx, C.c_int64)(
        _p(flwdir), C.c_int64(nrow), C.c_int64(ncol), _p(idxs_ds), _p(pits), C.byref(npits)
    )
    return idxs_ds, pits[: npits.value].copy(), int(n)


def _ldd_to_array(idxs_ds, shape, mv=None):
    a, sfx = _idx(idxs_ds)
    out = np.empty(a.size, dtype=np.uint8)
    rc = _fn("orc_ldd_to_array", sfx, C.c_int)(_p(a), C.c_int64(a.size), C.c_int64(shape[1]), _p(out))
    if rc != 0:
        raise ValueError("Invalid data downstream index outside 8 neighbors.")
    return out.reshape(shape)


core_ldd = types.SimpleNamespace(from_array=_ldd_from_array, to_array=_ldd_to_array)


# ----------------------------------------------------------------------------- core
def _rank(idxs_ds, mv=None):
    a, sfx = _idx(idxs_ds)
    ranks = np.empty(a.size, dtype=np.int32)
    n = _fn("orc_rank", sfx, C.c_int64)(_p(a), C.c_int64(a.size), _p(ranks))
    return ranks, int(n)


def _upstream_count(idxs_ds, mv=None, mask=None):
    a, sfx = _idx(idxs_ds)
    n_up = np.empty(a.size, dtype=np.int8)
    m = None if mask is None else np.ascontiguousarray(mask).astype(np.uint8)
    _fn("orc_upstream_count", sfx)(_p(a), C.c_int64(a.size), None if m is None else _p(m), _p(n_up))
    return n_up


def _idxs_seq(idxs_ds, idxs_pit, mv=None):
    a, sfx = _idx(idxs_ds)
    pits = np.ascontiguousarray(idxs_pit).astype(a.dtype)
    seq = np.empty(a.size, dtype=a.dtype)
    n = _fn("orc_idxs_seq", sfx, C.c_int64)(_p(a), C.c_int64(a.size), _p(pits), C.c_int64(pits.size), _p(seq))
    out = seq[: int(n)].copy()
    return out.astype(idxs_ds.dtype) if idxs_ds.dtype != out.dtype else out


def _pit_indices(idxs_ds):
    a, sfx = _idx(idxs_ds)
    pits = np.empty(a.size, dtype=a.dtype)
    n = _fn("orc_pit_indices", sfx, C.c_int64)(_p(a), C.c_int64(a.size), _p(pits))
    return pits[: int(n)].copy()


_UINT_OF_SIZE = {1: (np.uint8, "u8", C.c_uint8), 2: (np.uint16, "u16", C.c_uint16),
                 4: (np.uint32, "u32", C.c_uint32), 8: (np.uint64, "u64", C.c_uint64)}


def _fillnodata_upstream(idxs_ds, seq, data, nodata):
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    data = np.ascontiguousarray(data)
    if data.dtype.kind not in "iub":
        raise TypeError("oracle fillnodata_upstream: integer data only")
    udt, tsfx, ctype = _UINT_OF_SIZE[data.dtype.itemsize]
    nd = np.array([nodata]).astype(data.dtype).view(udt)[0]
    out = np.empty(data.size, dtype=data.dtype)
    _fn(f"orc_fillnodata_upstream_{tsfx}", sfx)(
        _p(a), _p(s), C.c_int64(s.size), _p(data), C.c_int64(data.size), ctype(int(nd)), _p(out)
    )
    return out


def _fillnodata_typed(name, idxs_ds, seq, data, nodata, how=None):
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    data = np.ascontiguousarray(data)
    tsfx = _DATA_SFX[data.dtype]
    out = np.empty(data.size, dtype=data.dtype)
    nd_f, nd_i, nd_is_int = _nodata_args(nodata)
    args = [_p(a), _p(s), C.c_int64(s.size), _p(data), C.c_int64(data.size), nd_f, nd_i, nd_is_int]
    if how is not None:
        args.append(C.c_int({"max": 0, "min": 1, "sum": 2}[how]))
    _fn(f"{name}_{tsfx}", sfx)(*args, _p(out))
    return out


def _fillnodata_upstream_any(idxs_ds, seq, data, nodata):
    """pyflwdir/core.py:120-146 (any dtype / nodata)"""
    return _fillnodata_typed("orc_fillnodata_up", idxs_ds, seq, data, nodata)


def _fillnodata_downstream(idxs_ds, seq, data, nodata, how="max"):
    """pyflwdir/core.py:149-188"""
    return _fillnodata_typed("orc_fillnodata_down", idxs_ds, seq, data, nodata, how)


def _main_upstream(idxs_ds, uparea, upa_min=0.0, mv=None):
    """pyflwdir/core.py:191-219"""
    a, sfx = _idx(idxs_ds)
    up = np.ascontiguousarray(uparea)
    tsfx = {np.dtype(np.int32): "i32", np.dtype(np.int64): "i64", np.dtype(np.float32): "f32",
            np.dtype(np.float64): "f64"}[up.dtype]
    out = np.empty(a.size, dtype=a.dtype)
    _fn(f"orc_main_upstream_{tsfx}", sfx)(_p(a), C.c_int64(a.size), _p(up), C.c_double(upa_min), _p(out))
    return out


def _stream_order_classic(idxs_ds, seq, idxs_us_main, mask=None, mv=None):
    """pyflwdir/streams.py:191-225"""
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    um = np.ascontiguousarray(idxs_us_main).astype(a.dtype)
    m = None if mask is None else np.ascontiguousarray(mask).astype(np.uint8)
    out = np.empty(a.size, dtype=np.uint8)
    _fn("orc_stream_order_classic", sfx)(_p(a), _p(s), C.c_int64(s.size), _p(um), None if m is None else _p(m),
                                         C.c_int64(a.size), _p(out))
    return out


def _upstream_sum(idxs_ds, data, nodata=-9999.0, mv=None):
    """pyflwdir/arithmetics.py:150-169"""
    a, sfx = _idx(idxs_ds)
    data = np.ascontiguousarray(data)
    tsfx = _DATA_SFX[data.dtype]
    out = np.empty(data.size, dtype=data.dtype)
    nd_f, nd_i, nd_is_int = _nodata_args(nodata)
    _fn(f"orc_upstream_sum_{tsfx}", sfx)(_p(a), _p(data), C.c_int64(data.size), nd_f, nd_i, nd_is_int, _p(out))
    return out


def _subbasins_streamorder(idxs_ds, seq, strord, mask=None, min_sto=-2):
    """pyflwdir/basins.py:67-103 -> (int32 map, outlet indices)"""
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    so = np.ascontiguousarray(strord, dtype=np.uint8)
    m = None if mask is None else np.ascontiguousarray(mask).astype(np.uint8)
    sub = np.empty(a.size, dtype=np.int32)
    idxs = np.empty(max(s.size, 1), dtype=a.dtype)
    n = _fn("orc_subbasins_streamorder", sfx, C.c_int64)(_p(a), _p(s), C.c_int64(s.size), _p(so), None if m is None else _p(m),
                                                         C.c_int64(int(min_sto)), C.c_int64(a.size), _p(sub), _p(idxs))
    return sub, idxs[: int(n)].copy()


def _subbasins_area(idxs_ds, seq, idxs_us_main, uparea, area_min):
    """pyflwdir/basins.py:194-233 -> (uint32 map, outlet indices)"""
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    um = np.ascontiguousarray(idxs_us_main).astype(a.dtype)
    upa = np.ascontiguousarray(uparea)
    tsfx = {np.dtype(np.int32): "i32", np.dtype(np.int64): "i64", np.dtype(np.float32): "f32",
            np.dtype(np.float64): "f64"}[upa.dtype]
    sub = np.empty(a.size, dtype=np.uint32)
    idxs = np.empty(max(s.size, 1), dtype=a.dtype)
    n = _fn(f"orc_subbasins_area_{tsfx}", sfx, C.c_int64)(_p(a), _p(s), C.c_int64(s.size), _p(um), _p(upa),
                                                          C.c_double(float(area_min)), C.c_int64(a.size), _p(sub), _p(idxs))
    return sub, idxs[: int(n)].copy()


_FSFX = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}


def _moving_average(data, weights, n, idxs_ds, idxs_us_main, strord=None, nodata=-9999.0, mv=None):
    """pyflwdir/arithmetics.py:67-103"""
    a, sfx = _idx(idxs_ds)
    um = np.ascontiguousarray(idxs_us_main).astype(a.dtype)
    data = np.ascontiguousarray(data)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    so = None if strord is None else np.ascontiguousarray(strord, dtype=np.uint8)
    out = np.empty(data.size, dtype=data.dtype)
    _fn(f"orc_moving_average_{_FSFX[data.dtype]}_f64", sfx)(
        _p(data), None if w is None else _p(w), C.c_int(int(n)), _p(a), _p(um), None if so is None else _p(so),
        C.c_double(float(nodata)), C.c_int64(data.size), _p(out))
    return out


def _moving_median(data, n, idxs_ds, idxs_us_main, strord=None, nodata=-9999.0, mv=None):
    """pyflwdir/arithmetics.py:106-147"""
    a, sfx = _idx(idxs_ds)
    um = np.ascontiguousarray(idxs_us_main).astype(a.dtype)
    data = np.ascontiguousarray(data)
    so = None if strord is None else np.ascontiguousarray(strord, dtype=np.uint8)
    out = np.empty(data.size, dtype=data.dtype)
    _fn(f"orc_moving_median_{_FSFX[data.dtype]}", sfx)(
        _p(data), C.c_int(int(n)), _p(a), _p(um), None if so is None else _p(so), C.c_double(float(nodata)),
        C.c_int64(data.size), _p(out))
    return out


arithmetics = types.SimpleNamespace(upstream_sum=_upstream_sum, moving_average=_moving_average, moving_median=_moving_median)

core = types.SimpleNamespace(
    fillnodata_upstream_any=_fillnodata_upstream_any,
    fillnodata_downstream=_fillnodata_downstream,
    main_upstream=_main_upstream,
    rank=_rank,
    upstream_count=_upstream_count,
    idxs_seq=_idxs_seq,
    pit_indices=_pit_indices,
    fillnodata_upstream=_fillnodata_upstream,
)


def _hop_table64(nrow, latlon, transform):
    """float64 table [nrow, 3, 2] of gis_utils.distance (gis_utils.py:452-486) for a D8 hop: (row of idx0, row delta + 1,
    column delta != 0). Includes the reference's swap for projected rasters (dy = xres, dx = yres)."""
    import ctypes
    import ctypes.util
    import math

    libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")  # numba's math.hypot is the C library's hypot
    libm.hypot.restype = C.c_double
    libm.hypot.argtypes = [C.c_double, C.c_double]
    xres, yres, north = transform[0], transform[4], transform[5]
    tab = np.zeros((nrow, 3, 2), dtype=np.float64)
    for r0 in range(nrow):
        for j, d in enumerate((-1, 0, 1)):
            dr = abs(d)
            for dc in (0, 1):
                if latlon:
                    lat = north + (r0 + (r0 + d)) / 2.0 * yres
                    rl = math.radians(lat)
                    dy = 0.0 if dr == 0 else (111132.92 + (-559.82 * math.cos(2.0 * rl)) + (1.175 * math.cos(4.0 * rl))
                                              + (-0.0023 * math.cos(6.0 * rl))) * yres
                    dx = 0.0 if dc == 0 else ((111412.84 * math.cos(rl)) + (-93.5 * math.cos(3.0 * rl))
                                              + (0.118 * math.cos(5.0 * rl))) * xres
                else:
                    dy, dx = xres, yres
                tab[r0, j, dc] = libm.hypot(float(dy * dr), float(dx * dc))
    return tab


def _path_impl(idxs0, idxs_nxt, ncol, mask, max_length, real_length, latlon, transform, want_paths):
    a, sfx = _idx(idxs_nxt)
    starts = np.ascontiguousarray(idxs0).astype(np.int64)
    m = None if mask is None else np.ascontiguousarray(mask).astype(np.uint8)
    hop = None
    if real_length and ncol is not None:
        hop = _hop_table64(a.size // int(ncol), latlon, transform)
    n0 = starts.size
    counts = np.zeros(n0, dtype=np.int64)
    ends = np.zeros(n0, dtype=np.int64)
    dists = np.zeros(n0, dtype=np.float64)
    f = _fn("orc_path", sfx)
    args = [_p(starts), C.c_int64(n0), _p(a), C.c_int64(int(ncol) if ncol is not None else 1), None if m is None else _p(m),
            C.c_int(0 if max_length is None else 1), C.c_double(0.0 if max_length is None else float(max_length)),
            None if hop is None else _p(hop), _p(counts), _p(ends), _p(dists)]
    f(*args, None)
    if not want_paths:
        return ends, dists
    flat = np.empty(max(int(counts.sum()), 1), dtype=a.dtype)
    f(*args, _p(flat))
    offs = np.concatenate([[0], np.cumsum(counts)])
    return [flat[offs[i]:offs[i + 1]].copy() for i in range(n0)], dists


def _path(idxs0, idxs_nxt, ncol=None, mask=None, max_length=None, real_length=False, latlon=False,
          transform=(1.0, 0.0, 0.0, 0.0, -1.0, 0.0), mv=None):
    """pyflwdir/core.py:400-438 -> (list of index arrays, float64 distances)"""
    return _path_impl(idxs0, idxs_nxt, ncol, mask, max_length, real_length, latlon, transform, True)


def _snap(idxs0, idxs_nxt, ncol=None, mask=None, max_length=None, real_length=False, latlon=False,
          transform=(1.0, 0.0, 0.0, 0.0, -1.0, 0.0), mv=None):
    """pyflwdir/core.py:441-480 -> (end indices in idxs0's dtype, float32 distances)"""
    ends, dists = _path_impl(idxs0, idxs_nxt, ncol, mask, max_length, real_length, latlon, transform, False)
    return ends.astype(np.asarray(idxs0).dtype), dists.astype(np.float32)


def _io_idxs(name, idxs_ds, seq, region):
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    r = np.ascontiguousarray(region).astype(np.uint8)
    out = np.empty(max(a.size, 1), dtype=a.dtype)
    n = _fn(name, sfx, C.c_int64)(_p(a), _p(s), C.c_int64(s.size), _p(r), C.c_int64(a.size), _p(out))
    return out[: int(n)].copy()


def _inflow_idxs(idxs_ds, seq, region):
    """pyflwdir/core.py:483-497"""
    return _io_idxs("orc_inflow_idxs", idxs_ds, seq, region)


def _outflow_idxs(idxs_ds, seq, region):
    """pyflwdir/core.py:500-514"""
    return _io_idxs("orc_outflow_idxs", idxs_ds, seq, region)


def _interbasin_mask(idxs_ds, seq, region, stream=None):
    """pyflwdir/basins.py:23-64"""
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    r = np.ascontiguousarray(region).astype(np.uint8)
    st = None if stream is None else np.ascontiguousarray(stream).astype(np.uint8)
    out = np.empty(a.size, dtype=np.uint8)
    _fn("orc_interbasin_mask", sfx)(_p(a), _p(s), C.c_int64(s.size), _p(r), None if st is None else _p(st), C.c_int64(a.size), _p(out))
    return out.astype(np.bool_)


def _region_outlets(regions, idxs_ds, seq):
    """pyflwdir/regions.py:132-163 -> (labels in regions' dtype, outlet indices)"""
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    reg = np.asarray(regions)
    r64 = np.ascontiguousarray(reg.ravel()).astype(np.int64)
    lbs = np.empty(max(s.size, 1), dtype=np.int64)
    idxs = np.empty(max(s.size, 1), dtype=a.dtype)
    n = int(_fn("orc_region_outlets", sfx, C.c_int64)(_p(r64), _p(a), _p(s), C.c_int64(s.size), _p(lbs), _p(idxs)))
    return lbs[:n].astype(reg.dtype), idxs[:n].copy()


core.path = _path
core.snap = _snap
core.inflow_idxs = _inflow_idxs
core.outflow_idxs = _outflow_idxs


def _region_bounds(regions, transform=(1.0, 0.0, 0.0, 0.0, -1.0, 0.0)):
    """pyflwdir/regions.py:58-129 (region_slices + region_bounds); scipy.ndimage.find_objects replaced by its definition:
    per label > 0 the smallest row / column slices that hold every cell with that label."""
    regions = np.asarray(regions)
    nrow, ncol = regions.shape
    lbs = np.unique(regions[regions > 0])
    a, b, c, d, e, f = [float(v) for v in tuple(transform)[:6]]
    xres, yres = a, e
    lons = (np.arange(ncol) + 0.5) * a + (np.zeros(ncol) + 0.5) * b + c  # gis_utils.affine_to_coords, gis_utils.py:340-358
    lats = (np.zeros(nrow) + 0.5) * d + (np.arange(nrow) + 0.5) * e + f
    iy = np.array([0, -1])
    ix = iy.copy()
    if yres < 0:
        iy = iy[::-1]
    if xres < 0:
        ix = ix[::-1]
    dx, dy = np.abs(xres) / 2, np.abs(yres) / 2
    rows, cols = np.nonzero(regions > 0)
    order = np.argsort(regions[rows, cols], kind="stable")
    rows, cols = rows[order], cols[order]
    first = np.flatnonzero(np.diff(regions[rows, cols], prepend=0) != 0)  # start of every label's run
    r0, r1 = np.minimum.reduceat(rows, first), np.maximum.reduceat(rows, first)
    c0, c1 = np.minimum.reduceat(cols, first), np.maximum.reduceat(cols, first)
    bboxs = []
    for k in range(lbs.size):
        xmin, xmax = lons[c0[k]: c1[k] + 1][ix]
        ymin, ymax = lats[r0[k]: r1[k] + 1][iy]
        bboxs.append([xmin - dx, ymin - dy, xmax + dx, ymax + dy])
    bboxs = np.asarray(bboxs)
    total_bbox = np.hstack([bboxs[:, :2].min(axis=0), bboxs[:, 2:].max(axis=0)])
    return lbs, bboxs, total_bbox


regions = types.SimpleNamespace(region_outlets=_region_outlets, region_bounds=_region_bounds)


# ----------------------------------------------------------------------------- core_nextxy
def _nextxy_from_array(flwdir, dtype=np.intp):
    """pyflwdir/core_nextxy.py:24-67"""
    nextx, nexty = flwdir
    nextx = np.ascontiguousarray(nextx, dtype=np.int32)
    nexty = np.ascontiguousarray(nexty, dtype=np.int32)
    dt = np.dtype(dtype)
    if dt == np.dtype(np.intp) and dt not in _SFX:  # pragma: no cover
        dt = np.dtype(np.int64)
    sfx = _SFX[dt]
    nrow, ncol = nextx.shape
    idxs_ds = np.empty(nextx.size, dtype=dt)
    pits = np.empty(nextx.size, dtype=dt)
    npits = C.c_int64()
    n = _fn("orc_nextxy_from_array", sfx, C.c_int64)(_p(nextx), _p(nexty), C.c_int64(nrow), C.c_int64(ncol), _p(idxs_ds), _p(pits),
                                                     C.byref(npits))
    return idxs_ds, pits[: npits.value].copy(), int(n)


def _nextxy_to_array(idxs_ds, shape, mv=None):
    """pyflwdir/core_nextxy.py:37-39,70-83"""
    a, sfx = _idx(idxs_ds)
    out = np.empty((2, a.size), dtype=np.int32)
    _fn("orc_nextxy_to_array", sfx)(_p(a), C.c_int64(a.size), C.c_int64(shape[1]), _p(out[0]), _p(out[1]))
    return out.reshape((2,) + tuple(shape))


core_nextxy = types.SimpleNamespace(from_array=_nextxy_from_array, to_array=_nextxy_to_array)


# ----------------------------------------------------------------------------- rivers
def _classify_estuary(idxs_ds, seq, idxs_pit, rivdst, rivwth, elevtn, max_elevtn=0, min_convergence=1e-2):
    """pyflwdir/rivers.py:11-53"""
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    rivdst, rivwth = np.ascontiguousarray(rivdst), np.ascontiguousarray(rivwth)
    est = np.zeros(a.size, np.int8)
    pits = np.asarray(idxs_pit)
    est[pits[np.asarray(elevtn)[pits] <= max_elevtn]] = 1
    _fn(f"orc_classify_estuary_{_FSFX[rivdst.dtype]}_{_FSFX[rivwth.dtype]}", sfx)(
        _p(a), _p(s), C.c_int64(s.size), _p(rivdst), _p(rivwth), C.c_double(float(min_convergence)), _p(est))
    return est


rivers = types.SimpleNamespace(classify_estuary=_classify_estuary)


# ----------------------------------------------------------------------------- streams
def _nodata_args(nodata):
    is_int = isinstance(nodata, (int, np.integer)) and not isinstance(nodata, (bool, np.bool_))
    nd_f = float(nodata)
    nd_i = int(nodata) if is_int else 0
    return C.c_double(nd_f), C.c_int64(nd_i), C.c_int(1 if is_int else 0)


def _accuflux_impl(idxs_ds, seq, data, nodata, downstream):
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    data = np.ascontiguousarray(data)
    if data.dtype not in _DATA_SFX:
        raise TypeError(f"oracle accuflux: unsupported data dtype {data.dtype}")
    tsfx = _DATA_SFX[data.dtype]
    out = np.empty(data.size, dtype=data.dtype)
    nd_f, nd_i, nd_is_int = _nodata_args(nodata)
    _fn(f"orc_accuflux_{tsfx}", sfx)(
        _p(a), _p(s), C.c_int64(s.size), _p(data), C.c_int64(data.size), nd_f, nd_i, nd_is_int,
        C.c_int(downstream), _p(out)
    )
    return out


def _accuflux(idxs_ds, seq, data, nodata):
    return _accuflux_impl(idxs_ds, seq, data, nodata, 0)


def _accuflux_ds(idxs_ds, seq, data, nodata):
    return _accuflux_impl(idxs_ds, seq, data, nodata, 1)


def _strahler_order(idxs_ds, seq, mask=None):
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    m = None if mask is None else np.ascontiguousarray(mask).astype(np.uint8)
    out = np.empty(a.size, dtype=np.uint8)
    _fn("orc_strahler", sfx)(
        _p(a), _p(s), C.c_int64(s.size), None if m is None else _p(m), C.c_int64(a.size), _p(out)
    )
    return out


def _stream_distance(idxs_ds, seq, ncol, mask=None, real_length=True, latlon=False, transform=(1.0, 0.0, 0.0, 0.0, -1.0, 0.0)):
    """pyflwdir/streams.py:272-315 (plain Python in the reference) with gis_utils.distance (gis_utils.py:451-486)
    restated inline; NumPy-2 scalar arithmetic (float32 array element + Python float -> float32)."""
    import math

    xres, yres, north = transform[0], transform[4], transform[5]

    def dmy(lat):
        rl = math.radians(lat)
        return 111132.92 + (-559.82 * math.cos(2.0 * rl)) + (1.175 * math.cos(4.0 * rl)) + (-0.0023 * math.cos(6.0 * rl))

    def dmx(lat):
        rl = math.radians(lat)
        return (111412.84 * math.cos(rl)) + (-93.5 * math.cos(3.0 * rl)) + (0.118 * math.cos(5.0 * rl))

    def distance(idx0, idx1):
        r0, r1 = int(idx0 // ncol), int(idx1 // ncol)
        dr = abs(r1 - r0)
        dc = abs(int(idx1 % ncol) - int(idx0 % ncol))
        if latlon:
            lat = north + (r0 + r1) / 2.0 * yres
            dy = 0.0 if dr == 0 else dmy(lat) * yres
            dx = 0.0 if dc == 0 else dmx(lat) * xres
        else:
            dy, dx = xres, yres
        return math.hypot(dy * dr, dx * dc)

    dist = np.full(idxs_ds.size, -9999.0, dtype=np.float32 if real_length else np.int32)
    dist[seq] = 0
    d = 1
    for idx0 in seq.tolist():
        idx_ds = int(idxs_ds[idx0])
        if idx0 == idx_ds or (mask is not None and mask[idx0]):
            continue
        if real_length:
            d = distance(idx0, idx_ds)
        dist[idx0] = dist[idx_ds] + d
    return dist



def _streams(idxs_ds, seq, mask=None, max_len=0, mv=None):
    """pyflwdir/streams.py:131-188 -> list of index arrays"""
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    m = None if mask is None else np.ascontiguousarray(mask).astype(np.uint8)
    f = _fn("orc_streams", sfx, C.c_int64)
    ncell = C.c_int64()
    args = [_p(a), _p(s), C.c_int64(s.size), None if m is None else _p(m), C.c_int64(int(max_len)), C.c_int64(a.size)]
    nseg = int(f(*args, None, None, C.byref(ncell)))
    offs = np.zeros(nseg + 1, dtype=np.int64)
    cells = np.empty(max(ncell.value, 1), dtype=a.dtype)
    f(*args, _p(offs), _p(cells), C.byref(ncell))
    return [cells[offs[i]:offs[i + 1]] for i in range(nseg)]

streams = types.SimpleNamespace(streams=_streams, stream_distance=_stream_distance, accuflux=_accuflux, accuflux_ds=_accuflux_ds, strahler_order=_strahler_order,
                                stream_order=lambda *a, **k: _stream_order_classic(*a, **k))


# ----------------------------------------------------------------------------- basins / dem
def _basins(idxs_ds, idxs_pit, seq, ids=None):
    """pyflwdir/basins.py:12-18"""
    if ids is None:
        ids = np.arange(1, idxs_pit.size + 1, dtype=np.uint32)
    b = np.zeros(idxs_ds.size, dtype=ids.dtype)
    b[idxs_pit] = ids
    return _fillnodata_upstream(idxs_ds, seq, b, 0)



def _subbasins_pfafstetter(idxs_pit, idxs_ds, seq, idxs_us_main, uparea, mask=None, depth=1, mv=None):
    """pyflwdir/basins.py:106-191 -> (pfafstetter map, outlet indices)"""
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    pits = np.ascontiguousarray(idxs_pit).astype(a.dtype)
    um = np.ascontiguousarray(idxs_us_main).astype(a.dtype)
    upa = np.ascontiguousarray(uparea, dtype=np.float64)
    m = None if mask is None else np.ascontiguousarray(mask).astype(np.uint8)
    out = np.empty(a.size, dtype=np.int32)
    idxs = np.empty(max(a.size, 1), dtype=a.dtype)
    n = int(_fn("orc_subbasins_pfafstetter", sfx, C.c_int64)(
        _p(pits), C.c_int64(pits.size), _p(a), _p(s), C.c_int64(s.size), _p(um), _p(upa), None if m is None else _p(m),
        C.c_int(int(depth)), C.c_int64(a.size), _p(out), _p(idxs)))
    if n < 0:
        raise IndexError("subbasins_pfafstetter: missing main-upstream index")
    return out.astype(np.int64), idxs[:n].copy()  # numba: int32 % int64 -> int64

basins = types.SimpleNamespace(basins=_basins, subbasins_streamorder=_subbasins_streamorder, subbasins_area=_subbasins_area,
                               interbasin_mask=_interbasin_mask, subbasins_pfafstetter=_subbasins_pfafstetter)


def _hand(idxs_ds, seq, drain, elevtn):
    a, sfx = _idx(idxs_ds)
    s = np.ascontiguousarray(seq).astype(a.dtype)
    d = np.ascontiguousarray(drain).astype(np.uint8)
    e = np.ascontiguousarray(elevtn)
    if e.dtype == np.float32:
        tsfx = "f32"
    elif e.dtype == np.float64:
        tsfx = "f64"
    else:
        raise TypeError("oracle hand: elevtn must be float32/float64")
    out = np.empty(a.size, dtype=np.float64)
    _fn(f"orc_hand_{tsfx}", sfx)(_p(a), _p(s), C.c_int64(s.size), _p(d), _p(e), C.c_int64(a.size), _p(out))
    return out


def _floodplains(idxs_ds, seq, elevtn, uparea, upa_min=1000.0, b=0.3):
    """pyflwdir/dem.py:333-379 (plain Python in the reference; numpy-scalar arithmetic kept as is)."""
    drainh = np.full(uparea.size, -9999.0, dtype=np.float32)
    drainz = np.full(uparea.size, -9999.0, dtype=np.float32)
    fldpln = np.full(uparea.size, -1, dtype=np.int8)
    fldpln[seq] = 0
    for idx0 in seq:
        if uparea[idx0] >= upa_min:
            drainh[idx0] = uparea[idx0] ** b
            drainz[idx0] = elevtn[idx0]
            fldpln[idx0] = 1
        else:
            idx_ds = idxs_ds[idx0]
            if fldpln[idx_ds] == 1:
                z0 = drainz[idx_ds]
                h0 = drainh[idx_ds]
                dh = elevtn[idx0] - z0
                if dh <= h0:
                    fldpln[idx0] = 1
                    drainz[idx0] = z0
                    drainh[idx0] = h0
    return fldpln


def _fill_depressions(elevtn, outlets="edge", idxs_pit=None, nodata=-9999.0, max_depth=-1.0, elv_max=None, connectivity=8):
    """pyflwdir/dem.py:17-143 -> (filled elevation, d8). float32 / float64 / integer elevation rasters."""
    e = np.ascontiguousarray(elevtn)
    if e.ndim != 2:
        raise ValueError("elevtn must be 2D")
    if connectivity not in (4, 8):
        raise ValueError('"connectivity" should either be 4 or 8')
    int_delv = 0
    if e.dtype == np.float32:
        w, sfx = e, "f32"
    elif e.dtype == np.float64:
        w, sfx = e, "f64"
    elif np.issubdtype(e.dtype, np.integer):
        w, sfx, int_delv = e.astype(np.float64), "f64", 1
    else:
        raise TypeError("oracle fill_depressions: float32 / float64 / integer elevation only")
    mode = 2 if idxs_pit is not None else {"edge": 0, "min": 1}.get(outlets, 0)
    pits = np.ascontiguousarray(idxs_pit if idxs_pit is not None else [], dtype=np.int64)
    out = np.empty_like(w)
    d8 = np.empty(w.shape, dtype=np.uint8)
    rc = getattr(lib(), f"orc_fill_depressions_{sfx}")(
        _p(w), C.c_int64(w.shape[0]), C.c_int64(w.shape[1]), C.c_int(mode), _p(pits), C.c_int64(pits.size),
        C.c_double(nodata), C.c_double(max_depth), C.c_int(elv_max is not None and mode != 2),
        C.c_double(0.0 if elv_max is None else elv_max), C.c_int(connectivity), C.c_int(int_delv), _p(out), _p(d8))
    if rc == 1:
        raise ValueError("No initial outlet cells found.")
    if rc != 0:
        raise MemoryError("oracle fill_depressions")
    return out.astype(e.dtype, copy=False), d8


dem = types.SimpleNamespace(height_above_nearest_drain=_hand, floodplains=_floodplains, fill_depressions=_fill_depressions)


# ----------------------------------------------------------------------------- synthetic input (host)
def synth_elevation(nrow, ncol, seed=0, octaves=None, nref=None):
    """Host version of the SURVEY §8(d) generator (bit-identical to the CUDA one)."""
    if nref is None:
        nref = 1 << int(np.ceil(np.log2(max(nrow, ncol, 8))))
    if octaves is None:
        octaves = max(1, int(np.log2(nref)) - 2)
    z = np.empty((nrow, ncol), dtype=np.float32)
    lib().orc_synth_elevation(C.c_int64(nrow), C.c_int64(ncol), C.c_int64(nref), C.c_int(octaves),
                              C.c_uint32(seed), _p(z))
    return z


def synth_d8(z, sea_level=-np.inf):
    z = np.ascontiguousarray(z, dtype=np.float32)
    d8 = np.empty(z.shape, dtype=np.uint8)
    lib().orc_synth_d8(_p(z), C.c_int64(z.shape[0]), C.c_int64(z.shape[1]), C.c_float(sea_level), _p(d8))
    return d8


def get_idxs_dtype(n):
    """pyflwdir/pyflwdir.py:105-127"""
    if n < 2147483647:
        return np.int32
    elif n < 4294967294:
        return np.uint32
    return np.int64
