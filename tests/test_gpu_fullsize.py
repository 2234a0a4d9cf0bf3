"""GPU tests at BASELINE.json sizes. The CPU oracle takes ~6 s at 8192^2 and minutes beyond, so parity at full
size is established through
  * the oracle itself at 2048^2 (seconds),
  * size-independent properties of the reference's definitions checked with numpy on the full outputs:
        rank[i]   == rank[ds[i]] + 1           (core.py:17-47; pits 0)
        basins[i] == basins[ds[i]]             (core.py:120-146)
        uparea[i] == 1 + sum(uparea[upstream]) (streams.py:15-41)  <=>  scatter-add of uparea over ds
        sum(uparea[pits]) == number of cells that drain to a pit
  * agreement of the two independent device implementations (tile solver vs level-synchronous BFS + sweeps),
  * agreement of the row-tiled multi-rank solve with the single-GPU solve."""
import ctypes as C

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _synth_on_device(size, seed, sea_quantile=None):
    from pyflwdir_b200 import _device, _lib

    dev = _device.DeviceGraph(0)
    l = _lib.lib()
    z = np.empty((size, size), np.float32)
    d8 = np.empty((size, size), np.uint8)
    octaves = max(1, int(np.log2(size)) - 2)
    _lib.check(l.pfd_synth_elevation(dev._h, size, size, size, octaves, seed, _lib.ptr(z)), dev._h)
    sea = float(np.quantile(z[::16, ::16], sea_quantile)) if sea_quantile else -np.inf
    _lib.check(l.pfd_synth_d8(dev._h, _lib.ptr(z), size, size, C.c_float(sea), _lib.ptr(d8)), dev._h)
    return d8, z


def _check_properties(d8, ids, rank, upa, bas, n_pits):
    n = d8.size
    ids = ids.astype(np.int64)
    idx = np.arange(n, dtype=np.int64)
    valid = ids >= 0
    pit = valid & (ids == idx)
    link = valid & ~pit
    assert int(pit.sum()) == n_pits
    assert np.all(rank[~valid] == -9999) and np.all(rank[pit] == 0) and np.all(rank[valid] >= 0)  # acyclic by construction
    assert np.all(rank[link] == rank[ids[link]] + 1)
    assert np.all(bas[link] == bas[ids[link]]) and np.all(bas[~valid] == 0)
    assert np.array_equal(bas[pit], np.arange(1, n_pits + 1, dtype=np.uint32))  # pit order = ascending index
    # uparea: every cell = 1 + sum over cells draining into it
    inflow = np.zeros(n, dtype=np.int64)
    np.add.at(inflow, ids[link], upa[link].astype(np.int64))
    assert np.array_equal(upa[valid].astype(np.int64), 1 + inflow[valid])
    assert np.all(upa[~valid] == -9999)
    assert int(upa[pit].astype(np.int64).sum()) == int(valid.sum())


def test_oracle_parity_2048():
    import pyflwdir_b200 as pfb

    d8, _ = _synth_on_device(2048, 5, sea_quantile=0.04)
    flw = pfb.from_array(d8, ftype="d8")
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    seq = oracle.core.idxs_seq(ids, pits)
    assert np.array_equal(flw.idxs_ds, ids) and np.array_equal(flw.idxs_pit, pits)
    assert np.array_equal(flw.idxs_seq, seq)
    assert np.array_equal(flw.rank.ravel(), oracle.core.rank(ids)[0])
    upa = oracle.streams.accuflux(ids, seq, np.ones(d8.size, np.int32), -9999)
    upa[ids == -1] = -9999
    assert np.array_equal(flw.upstream_area().ravel(), upa)
    assert np.array_equal(flw.basins().ravel(), oracle.basins.basins(ids, pits, seq))
    assert np.array_equal(flw.stream_order().ravel(), oracle.streams.strahler_order(ids, seq))


@pytest.mark.parametrize("size", [8192])
def test_properties_full_size(size):
    """BASELINE configs[1] size through pfd_d8_flow_all (the benchmarked call), both solvers."""
    from pyflwdir_b200 import _device, _lib

    d8, _ = _synth_on_device(size, 0)
    res = {}
    for tiles in (1, 0):
        dev = _device.DeviceGraph(0)
        dev.set_option("tiles", tiles)
        n = d8.size
        ids = np.empty(n, np.int32)
        rank = np.empty(n, np.int32)
        upa = np.empty(n, np.int32)
        bas = np.empty(n, np.uint32)
        nv, npit, nn = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().pfd_d8_flow_all(dev._h, _lib.ptr(d8), size, size, _lib.ptr(ids), _lib.dtype_code(np.int32),
                                             _lib.ptr(rank), _lib.ptr(upa), _lib.ptr(bas), C.byref(nv), C.byref(npit),
                                             C.byref(nn)), dev._h)
        res[tiles] = (ids, rank, upa, bas, nv.value, npit.value, nn.value)
        dev.close()
    ids, rank, upa, bas, nv, npit, nn = res[1]
    assert nv == n and nn == n  # no nodata, acyclic
    _check_properties(d8, ids, rank, upa, bas, npit)
    for a, b in zip(res[1], res[0]):
        assert np.array_equal(a, b), "tile solver and BFS + sweeps disagree"


def test_tiled_equals_single_gpu_4096():
    """Row-tiled solve (5 emulated ranks) == single-GPU solve on a 4096 x 2048 raster with a sea."""
    import pyflwdir_b200 as pfb
    from pyflwdir_b200 import tiled

    d8, _ = _synth_on_device(4096, 3, sea_quantile=0.03)
    d8 = np.ascontiguousarray(d8[:, :2048])
    flw = pfb.from_array(d8, ftype="d8")
    got = tiled.solve_emulated(d8, 5)
    assert np.array_equal(got["idxs_ds"], flw.idxs_ds)
    assert np.array_equal(got["rank"], flw.rank)
    assert np.array_equal(got["uparea"], flw.upstream_area())
    assert np.array_equal(got["basins"], flw.basins())


def test_later_rows_oracle_parity_3072():
    """The later rows with the most intricate ordering rules (stream segments, Pfafstetter coding, region outlets with
    ties, inflow / outflow cells, moving median, estuaries) against the oracle on a 3072^2 raster (9.4 M cells)."""
    import pyflwdir_b200 as pfb

    size = 3072
    d8, z = _synth_on_device(size, 9, sea_quantile=0.03)
    flw = pfb.from_array(d8, ftype="d8")
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    seq = oracle.core.idxs_seq(ids, pits)
    assert np.array_equal(flw.idxs_seq, seq)
    upa = flw.upstream_area()
    um = flw.idxs_us_main
    assert np.array_equal(um, oracle.core.main_upstream(ids, upa.ravel()))
    # stream segments of the channel network, split at 25 cells
    mask = upa > 200
    got = flw._dev.streams(mask, 25, np.int32)
    want = oracle.streams.streams(ids, seq, mask.ravel(), 25)
    assert len(got) == len(want) and np.array_equal(np.concatenate(got), np.concatenate(want))
    assert np.array_equal(np.array([g.size for g in got]), np.array([w.size for w in want]))
    # Pfafstetter, two levels, small tributaries masked out
    sub, sidx = flw.subbasins_pfafstetter(depth=2, upa_min=50)
    wsub, widx = oracle.basins.subbasins_pfafstetter(pits, ids, seq, um, upa.ravel(), mask=(upa >= 50).ravel(), depth=2)
    assert np.array_equal(sub.ravel(), wsub) and np.array_equal(sidx, widx)
    # region outlets on a label raster with many ties, inflow / outflow of a window, interbasin mask
    rr, cc = np.arange(size)[:, None], np.arange(size)[None, :]
    blocks = (((rr // 97) % 5) * 5 + ((cc // 131) % 5)).astype(np.int32)
    lbs, oidx = flw.basin_outlets(blocks)
    wl, wi = oracle.regions.region_outlets(blocks, ids, seq)
    assert np.array_equal(lbs, wl) and np.array_equal(oidx, wi)
    region = np.zeros((size, size), np.bool_)
    region[size // 5: 3 * size // 5, size // 4: 3 * size // 4] = True
    assert np.array_equal(flw.inflow_idxs(region), oracle.core.inflow_idxs(ids, seq, region.ravel()))
    assert np.array_equal(flw.outflow_idxs(region), oracle.core.outflow_idxs(ids, seq, region.ravel()))
    assert np.array_equal(flw.interbasin_mask(region, stream=mask).ravel(),
                          oracle.basins.interbasin_mask(ids, seq, region.ravel(), mask.ravel()))
    # bounding boxes of all basins against numpy
    bas = flw.basins()
    lbs, boxes, total = flw.basin_bounds(basins=bas)
    assert np.array_equal(lbs, np.arange(1, pits.size + 1, dtype=np.uint32))
    rows, cols = np.nonzero(bas)
    lab = bas[rows, cols].astype(np.int64) - 1
    rmin = np.full(pits.size, size, np.int64); np.minimum.at(rmin, lab, rows)
    rmax = np.full(pits.size, -1, np.int64); np.maximum.at(rmax, lab, rows)
    cmin = np.full(pits.size, size, np.int64); np.minimum.at(cmin, lab, cols)
    cmax = np.full(pits.size, -1, np.int64); np.maximum.at(cmax, lab, cols)
    assert np.array_equal(boxes, np.stack([cmin, -(rmax + 1.0), cmax + 1.0, -rmin.astype(np.float64)], axis=1))
    # moving median of the elevation along the main stem, estuary classification
    got = flw.moving_median(z, 5, nodata=-9999.0)
    assert np.array_equal(got.ravel(), oracle.arithmetics.moving_median(z.ravel(), 5, ids, um, None, -9999.0), equal_nan=True)
    rivwth = np.sqrt(np.abs(upa).astype(np.float32))
    dist = flw.stream_distance(unit="cell").astype(np.float32)
    est = flw.classify_estuaries(z - np.float32(np.median(z.ravel()[pits])), rivwth, rivdst=dist)
    assert np.array_equal(est, oracle.rivers.classify_estuary(ids, seq, pits, dist.ravel(), rivwth.ravel(),
                                                              z.ravel() - np.float32(np.median(z.ravel()[pits])), 0, 1e-2))
