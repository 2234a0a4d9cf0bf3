"""CPU tests of the host-side logic that mirrors the reference's wrappers (no GPU needed)."""
import numpy as np
import pytest

import _cases as cs
import pyflwdir_b200 as pfb
from pyflwdir_b200 import gis_utils as gis
from pyflwdir_b200._device import nodata_args


def test_get_idxs_dtype():
    """tests/test_pyflwdir.py:42-51 of the reference."""
    assert pfb._get_idxs_dtype(10) == np.int32
    assert pfb._get_idxs_dtype(2147483646) == np.int32
    assert pfb._get_idxs_dtype(2147483647) == np.uint32
    assert pfb._get_idxs_dtype(4294967293) == np.uint32
    assert pfb._get_idxs_dtype(4294967294) == np.int64
    assert pfb._get_idxs_dtype(2**40) == np.int64


def test_from_array_input_errors():
    """Error messages of pyflwdir.from_array raised before anything touches the device
    (reference tests/test_pyflwdir.py:29-39)."""
    with pytest.raises(ValueError, match="could not be inferred"):
        pfb.from_array(np.zeros((3, 3), dtype=np.float32))
    with pytest.raises(ValueError, match="should be 2 dimensional"):
        pfb.from_array(np.zeros(9, dtype=np.uint8), ftype="d8")
    with pytest.raises(ValueError, match='type "d8" is invalid'):
        pfb.from_array(np.zeros((3, 3), dtype=np.int16), ftype="d8")
    with pytest.raises(ValueError, match="Unknown flow direction type"):
        pfb.from_array(np.zeros((3, 3), dtype=np.uint8), ftype="d16")
    with pytest.raises(ValueError, match='"mask" shape does not match'):
        pfb.from_array(np.zeros((3, 3), dtype=np.uint8), ftype="d8", mask=np.ones((2, 2)))
    with pytest.raises(ValueError, match='type "nextxy" is invalid'):
        pfb.from_array((np.ones((3, 3), dtype=np.int64), np.ones((3, 3), dtype=np.int32)), ftype="nextxy")
    with pytest.raises(ValueError, match="should be 2 dimensional"):  # data[0] of a 2-D raster is a row (pyflwdir.py:168-170)
        pfb.from_array(np.zeros((3, 3), dtype=np.uint8), ftype="nextxy")
    with pytest.raises(ValueError, match='type "ldd" is invalid'):
        pfb.from_array(np.zeros((3, 3), dtype=np.int16), ftype="ldd")


def test_affine():
    a = gis.Affine(0.5, 0.0, 10.0, 0.0, -0.5, 50.0)
    assert tuple(a)[:6] == (0.5, 0.0, 10.0, 0.0, -0.5, 50.0) and a[0] == 0.5 and a[4] == -0.5
    x, y = a * (np.array([0.0, 2.0]), np.array([0.0, 4.0]))
    assert np.allclose(x, [10.0, 11.0]) and np.allclose(y, [50.0, 48.0])
    inv = ~a
    c, r = inv * (x, y)
    assert np.allclose(c, [0.0, 2.0]) and np.allclose(r, [0.0, 4.0])
    assert (a * gis.Affine.identity()) == a
    assert gis.IDENTITY == gis.Affine(1.0, 0.0, 0.0, 0.0, -1.0, 0.0)


def test_area_grid_matches_reference_rows():
    """gis_utils.area_grid against the reference's own area grid of the rhine raster (golden, float64)."""
    t = gis.Affine(*cs.RHINE_TRANSFORM)
    area = gis.area_grid(t, (682, 997), latlon=True)
    assert area.dtype == np.float64 and area.shape == (682, 997)
    assert np.array_equal(area[:, 0], cs.small()["out/rhine/area_col0"])
    assert np.array_equal(area[:, 0], area[:, -1])
    flat = gis.area_grid(gis.Affine(100.0, 0, 0, 0, -100.0, 0), (4, 5), latlon=False, unit="ha")
    assert flat.dtype == np.float32 and np.all(flat == np.float32(1.0))
    assert gis.area_grid(t, (2, 2), unit="cell").dtype == np.int32
    with pytest.raises(ValueError, match="Unknown unit"):
        gis.area_grid(t, (2, 2), unit="acre")


def test_nodata_args():
    f, i, is_int = nodata_args(-9999)
    assert (f.value, i.value, is_int.value) == (-9999.0, -9999, 1)
    f, i, is_int = nodata_args(-9999.0)
    assert (f.value, is_int.value) == (-9999.0, 0)
    f, i, is_int = nodata_args(np.float32(1.5))
    assert (f.value, is_int.value) == (1.5, 0)
    f, i, is_int = nodata_args(np.int16(7))
    assert (i.value, is_int.value) == (7, 1)


def test_ring_slot_numbering():
    """The tile solver's ring-slot numbering (pfd_tiles.cuh tl_ring_pos) is a bijection onto 0..251."""
    TL = 64
    seen = set()
    for ly in range(TL):
        for lx in range(TL):
            if ly in (0, TL - 1) or lx in (0, TL - 1):
                if ly == 0:
                    p = lx
                elif ly == TL - 1:
                    p = TL + lx
                elif lx == 0:
                    p = 2 * TL + ly - 1
                else:
                    p = 2 * TL + (TL - 2) + ly - 1
                assert p not in seen
                seen.add(p)
    assert seen == set(range(4 * TL - 4))


def test_split_rows():
    """Row blocks of the multi-GPU decomposition: multiples of 64 rows, contiguous, cover the raster."""
    from pyflwdir_b200 import tiled

    for nrow, R in [(65536, 8), (682, 3), (100, 4), (64, 2), (8192, 5), (1, 1)]:
        blocks = tiled.split_rows(nrow, R)
        assert blocks[0][0] == 0 and blocks[-1][1] == nrow
        for (a0, a1), (b0, b1) in zip(blocks[:-1], blocks[1:]):
            assert a1 == b0 and a0 % 64 == 0 and (a1 - a0) % 64 == 0 and a1 > a0
        assert len(blocks) <= R
    assert tiled.split_rows(65536, 8) == [(i * 8192, (i + 1) * 8192) for i in range(8)]
    d8 = np.arange(12, dtype=np.uint8).reshape(4, 3)
    blk, ht, hb = tiled.block_with_halo(d8, 0, 4)
    assert (ht, hb) == (0, 0) and blk.shape == (4, 3)


def test_infer_ncol():
    """Raster width inferred from a flat idxs_ds (the reference's free functions do not carry the shape)."""
    import oracle
    from pyflwdir_b200 import _functional as F

    for name in cs.SMALL_CASES:
        d8 = cs.case_d8(name)
        for dt in (np.int32, np.uint32, np.int64):
            ids, _, _ = oracle.core_d8.from_array(d8, dtype=dt)
            assert F.infer_ncol(ids) == d8.shape[1]
            assert F.resolve_shape(ids) == d8.shape
    assert F.infer_ncol(np.array([0, 1, 2, 3], dtype=np.int32)) == 4  # only pits: read as a single row
    with pytest.raises(ValueError, match="cannot infer"):
        F.infer_ncol(np.array([3, 1, 2, 3, 4, 5], dtype=np.int32))  # 0 -> 3 is "S" on 2x3 and "SE" on 3x2
    assert F.resolve_shape(np.arange(12, dtype=np.int32), shape=(3, 4)) == (3, 4)
    with pytest.raises(ValueError, match="does not match"):
        F.resolve_shape(np.arange(12, dtype=np.int32), shape=(5, 4))


def test_pinned_pool_recycles(monkeypatch):
    """Result arrays come from a capped pool of page-locked blocks that are recycled when the array and all its
    views are gone (host logic only: the allocator is faked)."""
    import ctypes as C
    import gc

    from pyflwdir_b200 import _lib

    class FakeLib:
        def __init__(self):
            self.bufs = {}

        def pfd_host_alloc(self, n, pp):
            b = (C.c_uint8 * n)()
            self.bufs[C.addressof(b)] = b
            pp._obj.value = C.addressof(b)
            return 0

        def pfd_host_free(self, p):
            self.bufs.pop(p.value, None)
            return 0

    fake = FakeLib()
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setenv("PFD_PINNED_POOL_MB", "16")
    pool = _lib.PinnedPool()
    a = pool.empty(1 << 20, np.int32)
    a[:] = 7
    v = a.reshape(1024, 1024)
    assert pool.total == 4 << 20 and not pool.free
    del a
    gc.collect()
    assert not pool.free and v.sum() == 7 * (1 << 20)   # the view keeps the block leased
    del v
    gc.collect()
    assert len(pool.free) == 1
    b = pool.empty(1 << 20, np.uint32)
    assert not pool.free and pool.total == 4 << 20        # same block handed out again
    c = pool.empty(3 << 20, np.int32)                      # 12 MiB more: 16 MiB cap reached exactly
    d = pool.empty(1 << 20, np.int32)                      # over the cap -> ordinary pageable array
    assert pool.total == 16 << 20 and d.flags.owndata
    assert pool.empty(10, np.int32).flags.owndata           # tiny results are never pinned
    del b, c, d


def test_geo_helpers_and_features():
    """Host-side geometry helpers behind path(unit="m") / streams / basin_bounds: transform_from_bounds, xy,
    idxs_to_coords, the two hop-length tables, gis_utils.features (vectorised, same dicts as the reference's loop)."""
    import ctypes
    import ctypes.util
    import math

    t = gis.transform_from_bounds(0.0, -10.0, 20.0, 6.0, 200, 160)
    assert tuple(t)[:6] == (0.1, 0.0, 0.0, 0.0, -0.1, 6.0)
    x, y = gis.xy(t, np.array([0, 3]), np.array([0, 7]))
    assert np.array_equal(x, 0.1 * np.array([0, 7]) + 0.0 * np.array([0, 3]) + (0.1 * 0.5 + 0.0 * 0.5 + 0.0))
    assert np.array_equal(y, 0.0 * np.array([0, 7]) + -0.1 * np.array([0, 3]) + (0.0 * 0.5 + -0.1 * 0.5 + 6.0))
    with pytest.raises(ValueError, match="Invalid offset"):
        gis.xy(t, 0, 0, offset="middle")
    xs, ys = gis.idxs_to_coords(np.array([0, 201]), t, (160, 200))
    assert np.array_equal(xs, gis.xy(t, [0, 1], [0, 1])[0]) and np.array_equal(ys, gis.xy(t, [0, 1], [0, 1])[1])
    with pytest.raises(IndexError, match="outside domain"):
        gis.idxs_to_coords(np.array([160 * 200]), t, (160, 200))

    # hop lengths: float32 table = CPython's math.hypot (interpreted stream_distance), float64 = the C library's (numba)
    tr = gis.Affine(*cs.RHINE_TRANSFORM)
    t32 = gis.hop_length_table(50, tr, True)
    t64 = gis.hop_length_table(50, tr, True, dtype=np.float64)
    assert t32.dtype == np.float32 and t64.dtype == np.float64 and t32.shape == t64.shape == (50, 3, 2)
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.hypot.restype = ctypes.c_double
    libm.hypot.argtypes = [ctypes.c_double, ctypes.c_double]
    r0, d = 17, 1
    lat = tr[5] + (r0 + (r0 + d)) / 2.0 * tr[4]
    dy, dx = gis.degree_metres_y(lat) * tr[4], gis.degree_metres_x(lat) * tr[0]
    assert t64[r0, 2, 1] == libm.hypot(dy, dx) and t32[r0, 2, 1] == np.float32(math.hypot(dy, dx))
    lat0 = tr[5] + (r0 + r0) / 2.0 * tr[4]  # a hop inside the row
    assert t64[r0, 1, 0] == 0.0 and t64[r0, 1, 1] == libm.hypot(0.0, gis.degree_metres_x(lat0) * tr[0])
    proj = gis.hop_length_table(3, gis.Affine(30.0, 0, 0, 0, -20.0, 0), False, dtype=np.float64)
    assert proj[0, 0, 0] == 30.0 and proj[0, 1, 1] == 20.0  # the reference's swap: dy = xres, dx = yres

    # features: one dict per path with >= 2 cells
    paths = [np.array([0, 1, 201], dtype=np.int32), np.array([5], dtype=np.int32), np.array([402, 402], dtype=np.int32)]
    upa = np.arange(160 * 200, dtype=np.float64).reshape(160, 200)
    feats = gis.features(paths, transform=t, shape=(160, 200), uparea=upa)
    assert len(feats) == 2
    f0, f1 = feats
    assert f0["type"] == "Feature" and f0["geometry"]["type"] == "LineString"
    want = list(zip(*gis.idxs_to_coords(paths[0], t, (160, 200))))
    assert f0["geometry"]["coordinates"] == want and isinstance(f0["geometry"]["coordinates"][0][0], np.float64)
    assert f0["properties"] == {"idx": 0, "idx_ds": 201, "pit": False, "uparea": 0.0} and f0["properties"]["idx"].dtype == np.int32
    assert f1["properties"]["pit"] and f1["properties"]["uparea"] == 402.0
    xs2, ys2 = np.arange(160 * 200) * 2.0, np.arange(160 * 200) * -1.0
    assert gis.features(paths[:1], xs=xs2, ys=ys2)[0]["geometry"]["coordinates"] == [(0.0, 0.0), (2.0, -1.0), (402.0, -201.0)]
    with pytest.raises(ValueError, match="transform and shape should be provided"):
        gis.features(paths)
    with pytest.raises(ValueError, match='Kwargs map "uparea"'):
        gis.features(paths, transform=t, shape=(160, 200), uparea=np.ones((2, 2)))
    assert gis.features([np.array([3])], transform=t, shape=(160, 200)) == []


def test_out_of_scope_entry_points_raise():
    """What stays on the reference side of the boundary says so instead of computing something else."""
    from pyflwdir_b200 import dem

    for fn, args in ((dem.adjust_elevation, (None, None, None)), (dem.dig_4connectivity, (None, None, None, (1, 1))),
                     (dem.slope, (np.zeros((4, 4), np.float32),))):
        with pytest.raises(NotImplementedError, match="outside the D8 hot path"):
            fn(*args)
    # the one branch of the priority flood that is not restated on the device says so before touching the GPU
    with pytest.raises(NotImplementedError, match="max_depth"):
        dem.fill_depressions(np.zeros((4, 4), np.float32), max_depth=2.0)
    with pytest.raises(ValueError, match="connectivity"):
        dem.fill_depressions(np.zeros((4, 4), np.float32), connectivity=6)


def test_region_sum_and_area_match_scipy():
    """regions.region_sum / region_area (regions.py:16-55): same labels and bit-identical float64 sums as the
    reference's scipy.ndimage.sum call."""
    from scipy import ndimage

    from pyflwdir_b200 import gis_utils as gis
    from pyflwdir_b200 import regions

    rng = np.random.default_rng(3)
    reg = rng.integers(0, 40, size=(120, 77)).astype(np.int32) * (rng.random((120, 77)) > 0.2)
    for data in (rng.random((120, 77)), rng.random((120, 77)).astype(np.float32) * 1e3, rng.integers(-5, 99, (120, 77))):
        lbs, sums = regions.region_sum(data, reg)
        want_lbs = np.unique(reg[reg > 0])
        assert np.array_equal(lbs, want_lbs) and np.array_equal(sums, ndimage.sum(data, reg, index=want_lbs))
    t = gis.Affine(0.01, 0.0, 4.0, 0.0, -0.01, 52.0)
    lbs, area = regions.region_area(reg, transform=t, latlon=True)
    assert np.array_equal(area, ndimage.sum(gis.area_grid(t, reg.shape, latlon=True), reg, index=lbs))
    with pytest.raises(NotImplementedError):
        regions.region_dissolve(reg, labels=[1])
