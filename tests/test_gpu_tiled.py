"""GPU tests of the row-tiled (multi-GPU) solve: R row blocks solved as R ranks on ONE GPU with the two exchanges
(pit counts, boundary-table all-reduce) emulated on the host -- the same kernels and step functions the NCCL path
uses. Compared bit for bit with the CPU oracle on the whole raster."""
import numpy as np
import pytest

import _cases as cs
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["fused", "separate-parse"], autouse=True)
def parse_mode(request, monkeypatch):
    """Both row-block paths: the block parsed inside phase A (default, the same program as the single-GPU step) and
    the separate parse pass."""
    monkeypatch.setenv("PFD_FUSE_PARSE", "1" if request.param == "fused" else "0")
    return request.param


def _oracle_whole(d8):
    dtype = oracle.get_idxs_dtype(d8.size)
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=dtype)
    seq = oracle.core.idxs_seq(ids, pits)
    rank, _ = oracle.core.rank(ids)
    upa = oracle.streams.accuflux(ids, seq, np.ones(d8.size, np.int32), -9999)
    upa[ids == ids.dtype.type(-1)] = -9999
    bas = oracle.basins.basins(ids, pits, seq)
    return ids, rank.reshape(d8.shape), upa.reshape(d8.shape), bas.reshape(d8.shape), pits.size


def _check(d8, nranks):
    from pyflwdir_b200 import tiled

    got = tiled.solve_emulated(d8, nranks)
    ids, rank, upa, bas, npits = _oracle_whole(d8)
    assert got["n_pits"] == npits
    assert np.array_equal(got["idxs_ds"], ids), "idxs_ds"
    assert np.array_equal(got["rank"], rank), "rank"
    assert np.array_equal(got["uparea"], upa), "uparea"
    assert np.array_equal(got["basins"], bas), "basins"
    return got


@pytest.mark.parametrize("nranks", [1, 2, 3, 5])
def test_tiled_synthetic(nranks):
    z = oracle.synth_elevation(448, 333, seed=41)
    d8 = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.05)))
    got = _check(d8, nranks)
    assert len(got["blocks"]) == min(nranks, 7)


@pytest.mark.parametrize("nranks", [2, 4])
def test_tiled_random_codes_with_loops(nranks):
    """Random legal codes: loops and chains that cross the block boundaries in both directions."""
    rng = np.random.default_rng(9)
    legal = np.array([32, 64, 128, 16, 0, 1, 8, 4, 2, 247, 255], dtype=np.uint8)
    p = np.array([1, 1, 1, 1, 0.02, 1, 1, 1, 1, 0.05, 0.02])
    d8 = legal[rng.choice(legal.size, size=(256, 200), p=p / p.sum())]
    _check(d8, nranks)


def test_tiled_zigzag_across_boundary():
    """One long river that crosses the boundary between two blocks many times (row 63 <-> row 64)."""
    nrow, ncol = 128, 96
    d8 = np.full((nrow, ncol), 4, dtype=np.uint8)   # everything flows south ...
    d8[64:, :] = 64                                  # ... or north, towards the boundary
    # the river alternates between row 63 and row 64 while heading east
    for c in range(ncol - 1):
        if c % 2 == 0:
            d8[63, c] = 2    # SE: (63,c) -> (64,c+1)
            d8[64, c] = 1    # E
        else:
            d8[64, c] = 128  # NE: (64,c) -> (63,c+1)
            d8[63, c] = 1    # E
    d8[63, ncol - 1] = 0
    d8[64, ncol - 1] = 64
    got = _check(d8, 2)
    assert got["uparea"].max() > nrow * ncol // 2


def test_tiled_rhine():
    d8 = cs.case_d8("rhine")
    got = _check(d8, 3)
    h = cs.hashes()["rhine"]
    assert cs.sha(got["rank"]) == h["rank"] and cs.sha(got["uparea"]) == h["uparea_cell"]
    assert cs.sha(got["basins"]) == h["basins"] and cs.sha(got["idxs_ds"]) == h["idxs_ds"]


@pytest.mark.parametrize("nranks", [1, 2, 3, 5])
def test_tiled_sweeps_strahler_accuflux_hand(nranks):
    """The order-sensitive outputs across row blocks (halo rounds emulated on the host): bit-identical to the oracle on
    the whole raster -- Strahler order, float32 / float64 / int32 accuflux with nodata, HAND with float32 / float64 elevation."""
    from pyflwdir_b200 import tiled

    z = oracle.synth_elevation(448, 333, seed=43)
    d8 = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.04)))
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    seq = oracle.core.idxs_seq(ids, pits)
    got, rounds, resolved = tiled.sweep_emulated(d8, nranks, "strahler")
    want = oracle.streams.strahler_order(ids, seq)
    bad = np.flatnonzero(got.ravel() != want)
    assert bad.size == 0, (bad.size, bad[:8] // d8.shape[1], bad[:8] % d8.shape[1], got.ravel()[bad[:8]], want[bad[:8]], rounds)
    assert resolved == seq.size
    assert rounds >= (2 if nranks > 1 else 1)
    rng = np.random.default_rng(5)
    for data, nodata in ((np.abs(z) + np.float32(0.25), -9999.0), (z.astype(np.float64) * 1e-3, -9999.0),
                         (rng.integers(0, 50, z.shape).astype(np.int32), -9999)):
        data = np.ascontiguousarray(data).copy()
        data.ravel()[rng.integers(0, data.size, 40)] = nodata
        got, _, _ = tiled.sweep_emulated(d8, nranks, "accuflux", data=data, nodata=nodata)
        assert got.dtype == data.dtype and np.array_equal(got.ravel(), oracle.streams.accuflux(ids, seq, data.ravel(), nodata))
    upa = oracle.streams.accuflux(ids, seq, np.ones(d8.size, np.int32), -9999)
    for drain, elev in ((upa > 60, z), (np.zeros(d8.size, np.bool_), z.astype(np.float64))):
        got, _, _ = tiled.sweep_emulated(d8, nranks, "hand", data=elev, drain=drain)
        want = oracle.dem.height_above_nearest_drain(ids, seq, np.asarray(drain).ravel(), np.asarray(elev).ravel())
        assert got.dtype == np.float64 and np.array_equal(got.ravel(), want)


def test_tiled_sweeps_zigzag_and_wide():
    """A river that crosses a block boundary many times (many rounds) and a raster wider than tall with an unaligned width."""
    from pyflwdir_b200 import tiled

    d8 = np.full((130, 70), 247, np.uint8)
    # a zig-zag channel around the block boundary (rows 127 | 128 for two blocks) flowing west -> east, pit at the east end
    r = 127
    for c in range(0, 68):
        d8[r, c] = 2 if r == 127 else 128  # SE from the upper row, NE from the lower row
        r = 128 if r == 127 else 127
    d8[r, 68] = 0
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    seq = oracle.core.idxs_seq(ids, pits)
    got, rounds, _ = tiled.sweep_emulated(d8, 2, "strahler")
    assert np.array_equal(got.ravel(), oracle.streams.strahler_order(ids, seq)) and rounds > 30
    area = np.linspace(0.5, 3.0, d8.size, dtype=np.float32).reshape(d8.shape)
    got, _, _ = tiled.sweep_emulated(d8, 2, "accuflux", data=area, nodata=-9999.0)
    assert np.array_equal(got.ravel(), oracle.streams.accuflux(ids, seq, area.ravel(), -9999.0))
    elev = np.linspace(9.0, 1.0, d8.size, dtype=np.float32).reshape(d8.shape)
    got, _, _ = tiled.sweep_emulated(d8, 2, "hand", data=elev, drain=np.zeros(d8.shape, np.bool_))
    assert np.array_equal(got.ravel(), oracle.dem.height_above_nearest_drain(ids, seq, np.zeros(d8.size, np.bool_), elev.ravel()))
    z = oracle.synth_elevation(200, 1001, seed=44)
    d8 = oracle.synth_d8(z)
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    seq = oracle.core.idxs_seq(ids, pits)
    got, _, _ = tiled.sweep_emulated(d8, 3, "strahler")
    assert np.array_equal(got.ravel(), oracle.streams.strahler_order(ids, seq))
