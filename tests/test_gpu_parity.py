"""GPU parity tests (run on the B200 box with -m gpu): pyflwdir_b200 (CUDA through the C ABI) against
 (1) the golden vectors generated from the real reference, bit for bit, and
 (2) the CPU oracle on fresh seeded inputs.
Nothing here reads /root/reference."""
import numpy as np
import pytest

import _cases as cs
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pfb():
    import pyflwdir_b200
    from pyflwdir_b200 import _lib

    assert _lib.device_count() > 0, "no CUDA device: the GPU tests must run on the B200 box"
    return pyflwdir_b200


@pytest.fixture(params=["tiles", "bfs"])
def solver(request, monkeypatch):
    """Both implementations of rank / basins() / upstream_area("cell") -- the tile-hierarchical solver (default) and the
    level-synchronous BFS + sweeps -- and of accuflux / Strahler / HAND: tile-dataflow sweeps (default) and level
    replays over the BFS order."""
    monkeypatch.setenv("PFD_TILES", "1" if request.param == "tiles" else "0")
    monkeypatch.setenv("PFD_TILE_SWEEPS", "2" if request.param == "tiles" else "0")
    return request.param


@pytest.mark.parametrize("name", cs.SMALL_CASES + ["synth512x768"])
def test_golden_cases(name, pfb, solver):
    d8 = cs.case_d8(name)
    aux = cs.case_inputs(name, d8, cs.case_seed(name))
    out = cs.run_api_case(pfb, d8, aux)
    for key, val in out.items():
        cs.check(name, key, val)


def test_golden_rhine(pfb, solver):
    d8 = cs.case_d8("rhine")
    aux = cs.case_inputs("rhine", d8, cs.case_seed("rhine"))
    out = cs.run_api_case(pfb, d8, aux, transform=cs.RHINE_TRANSFORM, latlon=True)
    for key, val in out.items():
        cs.check("rhine", key, val)
    assert int(out["rank"].max()) == 1674


@pytest.mark.parametrize("shape,seed,sea", [((257, 1023), 31, 0.1), ((1024, 1024), 32, 0.03), ((1500, 700), 33, 0.0),
                                             ((64, 4096), 34, 0.2), ((3000, 5), 35, 0.0), ((1, 977), 36, 0.0)])
def test_vs_oracle_synthetic(shape, seed, sea, pfb, solver):
    """Fresh synthetic terrain (aligned and unaligned widths, degenerate shapes) against the CPU oracle."""
    z = oracle.synth_elevation(shape[0], shape[1], seed=seed)
    d8 = oracle.synth_d8(z, sea_level=float(np.quantile(z, sea)) if sea > 0 else -np.inf)
    aux = cs.case_inputs("x", d8, seed)
    got = cs.run_api_case(pfb, d8, aux)
    want = cs.run_oracle_case(d8, aux, area=np.ones(d8.size, dtype=np.float32))
    for key in want:
        assert np.asarray(got[key]).dtype == np.asarray(want[key]).dtype, key
        assert np.array_equal(got[key], want[key], equal_nan=True), f"{key} differs from the oracle"


def test_vs_oracle_random_codes(pfb, solver):
    """Random legal codes: loops, forced pits at borders and next to nodata."""
    rng = np.random.default_rng(77)
    legal = np.array([32, 64, 128, 16, 0, 1, 8, 4, 2, 247, 255], dtype=np.uint8)
    for shape in [(97, 131), (128, 256), (301, 7)]:
        d8 = legal[rng.integers(0, legal.size, size=shape)]
        aux = cs.case_inputs("x", d8, 5)
        got = cs.run_api_case(pfb, d8, aux)
        want = cs.run_oracle_case(d8, aux, area=np.ones(d8.size, dtype=np.float32))
        for key in want:
            assert np.array_equal(got[key], want[key], equal_nan=True), f"{shape} {key} differs from the oracle"


def test_synth_generator_host_equals_device(pfb):
    """The CUDA input generator used by bench.py is bit-identical to the host one used by the tests."""
    import ctypes as C
    from pyflwdir_b200 import _lib, _device

    dev = _device.DeviceGraph(0)
    nrow, ncol = 300, 517
    z = np.empty((nrow, ncol), np.float32)
    d8 = np.empty((nrow, ncol), np.uint8)
    _lib.check(_lib.lib().pfd_synth_elevation(dev._h, nrow, ncol, 1024, 8, 9, _lib.ptr(z)), dev._h)
    zh = oracle.synth_elevation(nrow, ncol, seed=9, octaves=8, nref=1024)
    assert np.array_equal(z, zh)
    sea = float(np.quantile(zh, 0.1))
    _lib.check(_lib.lib().pfd_synth_d8(dev._h, _lib.ptr(z), nrow, ncol, C.c_float(sea), _lib.ptr(d8)), dev._h)
    assert np.array_equal(d8, oracle.synth_d8(zh, sea_level=sea))


def test_errors(pfb):
    with pytest.raises(ValueError, match="could not be inferred"):
        pfb.from_array(np.full((4, 4), 10, dtype=np.uint8))  # neither a D8 nor an LDD code
    with pytest.raises(ValueError, match="should be 2 dimensional"):
        pfb.from_array(np.zeros(16, dtype=np.uint8), ftype="d8")
    with pytest.raises(ValueError, match='type "d8" is invalid'):
        pfb.from_array(np.full((4, 4), 3, dtype=np.uint8), ftype="d8")
    with pytest.raises(ValueError, match="no pits found"):
        pfb.from_array(np.array([[1, 16], [1, 16]], dtype=np.uint8), ftype="d8")
    with pytest.raises(ValueError, match="mask"):
        pfb.from_array(np.zeros((4, 4), dtype=np.uint8), ftype="d8", mask=np.ones((3, 3)))
    flw = pfb.from_array(np.zeros((4, 4), dtype=np.uint8), ftype="d8")
    with pytest.raises(ValueError, match="Unknown unit"):
        flw.upstream_area("furlong")
    with pytest.raises(ValueError, match="IDs size does not match"):
        flw.basins(ids=np.arange(3))
    with pytest.raises(ValueError, match="cannot contain a value zero"):
        flw.basins(ids=np.zeros(16, dtype=np.uint32))
    with pytest.raises(ValueError, match="size does not match"):
        flw.accuflux(np.ones(5))
    with pytest.raises(ValueError, match="Invalid method"):
        flw.order_cells("bogus")


def test_constructor_from_idxs_ds(pfb):
    """FlwdirRaster(idxs_ds, shape, 'd8') round trip + mask + dump/load."""
    d8 = cs.case_d8("flwdir1_asc")
    flw = pfb.from_array(d8, ftype="d8")
    flw2 = pfb.FlwdirRaster(flw.idxs_ds.copy(), d8.shape, "d8")
    assert np.array_equal(flw2.idxs_pit, flw.idxs_pit)
    assert np.array_equal(flw2.idxs_seq, flw.idxs_seq)
    assert np.array_equal(flw2.upstream_area(), flw.upstream_area())
    assert np.array_equal(flw2.to_array() == 247, d8 == 247)
    m = np.zeros(d8.shape, bool)
    m[40:120, 30:150] = True
    flw3 = pfb.from_array(d8, ftype="d8", mask=m)
    ids, pits, _ = oracle.core_d8.from_array(np.where(m, d8, 247).astype(np.uint8), dtype=np.int32)
    assert np.array_equal(flw3.idxs_ds, ids) and np.array_equal(flw3.idxs_pit, pits)


def test_module_level_functions(pfb):
    """The L1/L2 free functions with the reference's signatures (what the reference's own tests call):
    core_d8.from_array / to_array / isvalid, core.rank / idxs_seq / upstream_count / pit_indices,
    streams.accuflux / strahler_order, basins.basins, dem.height_above_nearest_drain."""
    from pyflwdir_b200 import basins, core, core_d8, dem, streams

    d8 = cs.case_d8("flwdir_asc")
    # tests/conftest.py:23-26 of the reference parses the fixture with uint32 indices
    ids, pits, n = core_d8.from_array(d8, dtype=np.uint32)
    assert np.array_equal(ids, cs.small()["out/flwdir_asc/idxs_ds_uint32"]) and ids.dtype == np.uint32
    assert np.array_equal(pits, cs.small()["out/flwdir_asc/idxs_pit_uint32"]) and n == 407
    assert core_d8.isvalid(d8) and not core_d8.isvalid(np.full((3, 3), 3, np.uint8)) and not core_d8.isvalid(d8.astype(np.int32))
    # tests/test_core_xx.py:54-62: to_array -> from_array is the identity
    d8b = core_d8.to_array(ids, d8.shape)
    ids2, pits2, _ = core_d8.from_array(d8b, dtype=np.uint32)
    assert np.array_equal(ids2, ids) and np.array_equal(pits2, pits)

    for name in ["flwdir1_asc", "random48x61"]:
        d8 = cs.case_d8(name)
        ids, pits, _ = core_d8.from_array(d8, dtype=np.int32)
        rnk, n = core.rank(ids)
        assert np.array_equal(rnk.reshape(d8.shape), cs.golden(name, "rank")) and n == int(cs.golden(name, "nnodes"))
        # tests/test_core.py:18-22: rank[i] == rank[ds[i]] + 1 away from pits, #rank0 == #pits
        ok = rnk > 0
        assert np.all(rnk[ok] == rnk[ids[ok]] + 1) and np.count_nonzero(rnk == 0) == pits.size
        seq = core.idxs_seq(ids, pits)
        assert np.array_equal(seq, cs.golden(name, "idxs_seq")) and np.all(np.diff(rnk[seq]) >= 0)
        assert np.array_equal(core.upstream_count(ids).reshape(d8.shape), cs.golden(name, "n_upstream"))
        assert np.array_equal(core.pit_indices(ids), pits)
        aux = cs.case_inputs(name, d8, cs.case_seed(name))
        acc = streams.accuflux(ids, seq, aux["data_f64"].ravel(), -9999.0)
        assert np.array_equal(acc.reshape(d8.shape), cs.golden(name, "accu_f64"))
        accd = streams.accuflux_ds(ids, seq, aux["data_i64"].ravel(), -9999)
        assert np.array_equal(accd.reshape(d8.shape), cs.golden(name, "accu_ds_i64"))
        assert np.array_equal(streams.strahler_order(ids, seq).reshape(d8.shape), cs.golden(name, "strord"))
        assert np.array_equal(streams.strahler_order(ids, seq, mask=aux["smask"].ravel()).reshape(d8.shape),
                              cs.golden(name, "strord_mask"))
        assert np.array_equal(basins.basins(ids, pits, seq).reshape(d8.shape), cs.golden(name, "basins"))
        sub = basins.basins(ids, cs.golden(name, "sub_idxs"), seq, ids=cs.golden(name, "sub_ids"))
        assert np.array_equal(sub.reshape(d8.shape), cs.golden(name, "basins_sub")) and sub.dtype == np.int32
        drain = cs.golden(name, "uparea_cell") > max(4, int(0.002 * d8.size))
        hand = dem.height_above_nearest_drain(ids, seq, drain.ravel(), aux["elevtn"].ravel())
        assert np.array_equal(hand.reshape(d8.shape), cs.golden(name, "hand_f32"))
    with pytest.raises(ValueError, match="outside 8 neighbors"):
        core_d8.to_array(np.array([5, 1, 2, 3, 4, 5], dtype=np.int32), (2, 3))


def test_ldd(pfb):
    """PCRaster LDD rasters (core_ldd.py): inferred ftype, same graph as the D8 original, to_array both ways."""
    from pyflwdir_b200 import core_ldd

    s = cs.small()
    ldd = s["in/ldd_flwdir1/ldd"]
    flw = pfb.from_array(ldd)  # d8 is tried first and refused, then ldd (pyflwdir.py:39-48)
    assert flw.ftype == "ldd"
    assert np.array_equal(flw.idxs_ds, s["out/ldd_flwdir1/idxs_ds"])
    assert np.array_equal(flw.idxs_pit, s["out/ldd_flwdir1/idxs_pit"])
    assert np.array_equal(flw.idxs_outlet, s["out/ldd_flwdir1/idxs_outlet"])
    assert np.array_equal(flw.to_array(), s["out/ldd_flwdir1/to_array"])
    assert np.array_equal(flw.to_array("d8"), s["out/ldd_flwdir1/to_array_d8"])
    assert np.array_equal(flw.upstream_area(), s["out/ldd_flwdir1/uparea_cell"])
    d8 = cs.case_d8("flwdir1_asc")
    assert np.array_equal(pfb.d8_to_ldd(d8), s["out/ldd_flwdir1/d8_to_ldd"])
    assert np.array_equal(pfb.ldd_to_d8(ldd), s["out/ldd_flwdir1/ldd_to_d8"])
    ids, pits, n = core_ldd.from_array(ldd, dtype=np.uint32)
    assert np.array_equal(ids.astype(np.int64), s["out/ldd_flwdir1/idxs_ds"].astype(np.uint32).astype(np.int64))
    assert core_ldd.isvalid(ldd) and not core_ldd.isvalid(d8)
    assert np.array_equal(core_ldd.to_array(s["out/ldd_flwdir1/idxs_ds"], ldd.shape), s["out/ldd_flwdir1/to_array"])
    with pytest.raises(ValueError, match='type "ldd" is invalid'):
        pfb.from_array(d8, ftype="ldd")
    # unaligned width + mask through the LDD table
    lddu = np.ascontiguousarray(ldd[:, :197])
    ids_o, pits_o, _ = oracle.core_ldd.from_array(lddu, dtype=np.int32)
    flwu = pfb.from_array(lddu, ftype="ldd")
    assert np.array_equal(flwu.idxs_ds, ids_o) and np.array_equal(flwu.idxs_pit, pits_o)


def test_next_rows_module_level(pfb):
    """main_upstream / classic stream order / fillnodata through the module-level mirrors."""
    from pyflwdir_b200 import core, streams

    name = "flwdir1_asc"
    d8 = cs.case_d8(name)
    aux = cs.case_inputs(name, d8, cs.case_seed(name))
    ids, seq = cs.golden(name, "idxs_ds"), cs.golden(name, "idxs_seq")
    um = core.main_upstream(ids, cs.golden(name, "uparea_cell").ravel())
    assert np.array_equal(um, cs.golden(name, "us_main")) and um.dtype == ids.dtype
    so = streams.stream_order(ids, seq, um, mask=aux["smask"].ravel())
    assert np.array_equal(so.reshape(d8.shape), cs.golden(name, "strord_classic_mask"))
    assert np.array_equal(core.upstream_count(ids, mask=aux["smask"].ravel()),
                          oracle.core.upstream_count(ids, mask=aux["smask"].ravel()))
    f = core.fillnodata_downstream(ids, seq, aux["fill_f32"].ravel(), -1.5, how="sum")
    assert np.array_equal(f.reshape(d8.shape), cs.golden(name, "fill_down_sum_f32"))
    f = core.fillnodata_upstream(ids, seq, aux["fill_f64"].ravel(), np.nan)
    assert np.array_equal(f.reshape(d8.shape), cs.golden(name, "fill_up_f64nan"), equal_nan=True)


def test_local_traces_errors_and_snapped_pits(pfb):
    """pfd_trace refuses start cells outside the raster and traces that never end; add_pits(streams=...) snaps the new
    pits to the stream mask first (flwdir.py:805-811), like the reference."""
    z = oracle.synth_elevation(200, 260, seed=12)
    d8 = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.05)))
    flw = pfb.from_array(d8, ftype="d8")
    with pytest.raises(ValueError, match="outside the raster"):
        flw.path(idxs=np.array([d8.size + 3]))
    with pytest.raises(ValueError, match="outside the raster"):
        flw.snap(idxs=np.array([-5]))
    with pytest.raises(ValueError, match="Either idxs or xy"):
        flw.path()
    with pytest.raises(ValueError, match="Unknown unit"):
        flw.snap(idxs=np.array([1]), unit="km")
    # xy start points go through FlwdirRaster.index
    xs, ys = flw.xy(np.array([5 * 260 + 7, 100 * 260 + 31]))
    p_xy, d_xy = flw.path(xy=(xs, ys))
    p_ix, d_ix = flw.path(idxs=np.array([5 * 260 + 7, 100 * 260 + 31]))
    assert all(np.array_equal(a, b) for a, b in zip(p_xy, p_ix)) and np.array_equal(d_xy, d_ix)
    # a 2-cell loop without a stop condition does not end
    loop = np.array([[1, 16, 0]], dtype=np.uint8)
    flw_loop = pfb.from_array(loop, ftype="d8")
    with pytest.raises(ValueError, match="does not end"):
        flw_loop.path(idxs=np.array([0]))
    paths, dist = flw_loop.path(idxs=np.array([0]), max_length=4)
    assert paths[0].tolist() == [0, 1, 0, 1, 0] and dist[0] == 4.0
    # add_pits with a stream mask
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    seq = oracle.core.idxs_seq(ids, pits)
    upa = oracle.streams.accuflux(ids, seq, np.ones(d8.size, np.int32), -9999)
    streams = (upa > 40).reshape(d8.shape)
    new = seq[:: seq.size // 9][1:6].astype(np.int64)
    snapped, _ = oracle.core.snap(new, ids, mask=streams.ravel())
    flw.add_pits(idxs=new, streams=streams)
    want = np.unique(np.concatenate([pits, snapped.astype(pits.dtype)]))
    assert np.array_equal(flw.idxs_pit, want)
    ids2 = ids.copy()
    ids2[snapped] = snapped
    assert np.array_equal(flw.idxs_seq, oracle.core.idxs_seq(ids2, want))


def test_nextxy(pfb):
    """CaMa-Flood NEXTXY codec (core_nextxy.py) against the reference's golden outputs: inferred ftype, -10 pits, a pit in
    the nexty plane only, mask=, to_array in all three formats, and the refusal of links longer than one cell."""
    from pyflwdir_b200 import core_nextxy

    s = cs.small()
    nxy = s["in/nextxy_flwdir1/nextxy"]
    shape = nxy.shape[1:]
    for data in (nxy, (nxy[0], nxy[1])):
        flw = pfb.from_array(data)  # ftype inferred
        assert flw.ftype == "nextxy" and flw.shape == shape
        # idxs_seq of a nextxy raster = np.argsort(rank) in the reference: ties follow numpy's unstable sort
        seq, want = flw.idxs_seq, s["out/nextxy_flwdir1/idxs_seq"]
        rank = flw.rank.ravel()
        assert seq.dtype == want.dtype and np.array_equal(np.sort(seq), np.sort(want)) and np.all(np.diff(rank[seq]) >= 0)
        for key, got in dict(idxs_ds=flw.idxs_ds, idxs_pit=flw.idxs_pit, idxs_outlet=flw.idxs_outlet,
                             to_array=flw.to_array(), to_array_d8=flw.to_array("d8"), uparea_cell=flw.upstream_area(),
                             basins=flw.basins()).items():
            want = s[f"out/nextxy_flwdir1/{key}"]
            assert np.asarray(got).dtype == want.dtype, key
            assert np.array_equal(got, want), key
    mask = np.ones(nxy.shape, dtype=np.uint8)
    mask[:, 100:, 150:] = 0
    flw_m = pfb.from_array(nxy, ftype="nextxy", mask=mask)
    assert np.array_equal(flw_m.idxs_ds, s["out/nextxy_flwdir1/masked_idxs_ds"])
    assert np.array_equal(flw_m.idxs_pit, s["out/nextxy_flwdir1/masked_idxs_pit"])
    d8 = cs.case_d8("flwdir1_asc")
    assert np.array_equal(pfb.from_array(d8, ftype="d8").to_array("nextxy"), s["out/nextxy_flwdir1/d8_to_nextxy"])
    # module-level mirror
    ids, pits, n = core_nextxy.from_array(nxy, dtype=np.uint32)
    assert ids.dtype == np.uint32 and np.array_equal(ids.astype(np.int64), s["out/nextxy_flwdir1/idxs_ds"].astype(np.uint32).astype(np.int64))
    assert np.array_equal(core_nextxy.to_array(s["out/nextxy_flwdir1/idxs_ds"], shape), s["out/nextxy_flwdir1/to_array"])
    assert core_nextxy.isvalid(nxy) and not core_nextxy.isvalid(nxy.astype(np.int64))
    bad = nxy.copy()
    bad[0, 50, 50] = -3
    assert not core_nextxy.isvalid(bad)
    with pytest.raises(ValueError, match='type "nextxy" is invalid'):
        pfb.from_array(bad, ftype="nextxy")
    far = nxy.copy()
    far[:, 60, 60] = (150, 120)  # a link across the raster: valid NEXTXY, but not representable here
    with pytest.raises(ValueError, match="leaves the 8 neighbours"):
        pfb.from_array(far, ftype="nextxy")
    # read_nextxy: CaMa-Flood binary layout
    import os
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        fn = os.path.join(tmp, "nextxy.bin")
        nxy.tofile(fn)
        data, transform = pfb.read_nextxy(fn, shape[0], shape[1], [0.0, -10.0, 20.0, 6.0])
        assert np.array_equal(data, nxy) and tuple(transform)[:6] == (0.1, 0.0, 0.0, 0.0, -0.1, 6.0)


def test_streams_raw_segments(pfb):
    """streams.streams on a 1024 x 1536 raster against the oracle, as raw index lists (no geo-features): no mask, a
    contiguous stream mask, a mask with gaps; unsplit and split segments."""
    from pyflwdir_b200 import streams as gstreams

    z = oracle.synth_elevation(1024, 1536, seed=44)
    d8 = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.04)))
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    seq = oracle.core.idxs_seq(ids, pits)
    upa = oracle.streams.accuflux(ids, seq, np.ones(d8.size, np.int32), -9999)
    gaps = np.random.default_rng(9).random(d8.size) < 0.5
    for mask, max_len in ((None, 0), (None, 5), (upa > 25, 0), (upa > 25, 12), (gaps, 2)):
        got = gstreams.streams(ids, seq, mask, max_len, shape=d8.shape)
        want = oracle.streams.streams(ids, seq, mask, max_len)
        assert len(got) == len(want)
        assert np.array_equal(np.array([g.size for g in got]), np.array([w.size for w in want]))
        assert np.array_equal(np.concatenate(got), np.concatenate(want)) and got[0].dtype == want[0].dtype


def test_later_rows_module_level(pfb):
    """The module-level mirrors of the later rows (arithmetics / basins / core / regions / rivers / streams) against the
    golden outputs of the reference on tests/data/flwdir1.asc."""
    from pyflwdir_b200 import arithmetics, basins, core, regions, rivers, streams

    name = "flwdir1_asc"
    d8 = cs.case_d8(name)
    aux = cs.case_inputs(name, d8, cs.case_seed(name))
    g = lambda key: cs.golden(name, key)
    ids, seq, pits, um = g("idxs_ds"), g("idxs_seq"), g("idxs_pit"), g("us_main")
    shape = d8.shape
    so, upa = g("strord").ravel(), g("uparea_cell").ravel()
    assert np.array_equal(arithmetics.moving_average(aux["data_f32_nd"].ravel(), None, 3, ids, um).reshape(shape), g("movavg_f32"),
                          equal_nan=True)
    assert np.array_equal(arithmetics.moving_median(aux["data_f64"].ravel(), 4, ids, um, strord=so).reshape(shape),
                          g("movmed_f64_so"), equal_nan=True)
    assert np.array_equal(arithmetics.upstream_sum(ids, aux["data_f64"].ravel(), -9999.0).reshape(shape), g("upsum_f64"))
    sub, idxs = basins.subbasins_pfafstetter(pits, ids, seq, um, upa, mask=upa >= 0.0, depth=2)
    assert np.array_equal(sub.reshape(shape), g("pfaf_d2")) and np.array_equal(idxs, g("pfaf_d2_idxs")) and idxs.dtype == ids.dtype
    sub, idxs = basins.subbasins_streamorder(ids, seq, so, None, -2)
    assert np.array_equal(sub.reshape(shape), g("subbas_so")) and np.array_equal(idxs, g("subbas_so_idxs"))
    sub, idxs = basins.subbasins_area(ids, seq, um, upa, max(3, d8.size // 400))
    assert np.array_equal(sub.reshape(shape), g("subbas_area_cell")) and np.array_equal(idxs, g("subbas_area_cell_idxs"))
    starts, region, blocks, labels = cs.local_inputs(d8, seq, aux)
    stream = upa > max(4, int(0.002 * d8.size))
    assert np.array_equal(basins.interbasin_mask(ids, seq, region.ravel(), stream).reshape(shape), g("interbasin_stream"))
    assert np.array_equal(core.inflow_idxs(ids, seq, region.ravel()), g("inflow_idxs"))
    assert np.array_equal(core.outflow_idxs(ids, seq, region.ravel()), g("outflow_idxs"))
    paths, dist = core.path(starts, ids, mask=stream)
    assert np.array_equal(np.concatenate(paths), g("path_down_mask")) and np.array_equal(dist, g("path_down_mask_dist"))
    ends, dist = core.snap(starts, ids, max_length=7.5)
    assert np.array_equal(ends, g("snap_down_max")) and np.array_equal(dist, g("snap_down_max_dist")) and dist.dtype == np.float32
    lbs, idxs = regions.region_outlets(blocks, ids, seq)
    assert np.array_equal(lbs, g("outlets_blocks_lbs")) and np.array_equal(idxs, g("outlets_blocks_idxs")) and lbs.dtype == blocks.dtype
    lbs, boxes, total = regions.region_bounds(labels)
    assert np.array_equal(lbs, g("bounds_labels_lbs")) and np.array_equal(boxes, g("bounds_labels_boxes"))
    assert np.array_equal(total, g("bounds_labels_total"))
    lbs2, slices = regions.region_slices(labels)
    assert len(slices) == lbs.size and all(np.all(labels[s] == lb) or (labels[s] == lb).any() for lb, s in zip(lbs2, slices))
    segs = streams.streams(ids, seq, so >= 2, 0)
    segs = [s for s in segs if s.size >= 2]
    assert np.array_equal(np.array([s[0] for s in segs]), g("streams_so2_idx")) and np.array_equal(
        np.array([s.size for s in segs]), g("streams_so2_n"))
    distnc = g("sdist_m").ravel()
    rivwth32 = np.sqrt(np.abs(upa).astype(np.float32)) + aux["data_f32"].ravel()
    elev0 = aux["elevtn"].ravel() - np.float32(np.median(aux["elevtn"].ravel()[pits]))
    assert np.array_equal(rivers.classify_estuary(ids, seq, pits, distnc, rivwth32, elev0, 0, 1e-2), g("estuary_f32"))
    from pyflwdir_b200 import dem

    fp = dem.floodplains(ids, seq, (aux["elevtn"].astype(np.float64) * 1.1).ravel(), upa, upa_min=max(4, int(0.002 * d8.size)), b=0.5)
    assert np.array_equal(fp.reshape(shape), g("fldpln_f64")) and fp.dtype == np.int8


def test_small_mirrors(pfb):
    """streams.upstream_area (no area grid) against the reference's outputs; loop / headwater / confluence indices
    against their definitions."""
    from pyflwdir_b200 import core, streams

    s = cs.small()
    d8 = cs.case_d8("flwdir1_asc")
    ids, seq = cs.golden("flwdir1_asc", "idxs_ds"), cs.golden("flwdir1_asc", "idxs_seq")
    got = streams.upstream_area(ids, seq, d8.shape[1], dtype=np.int32)
    assert got.dtype == np.int32 and np.array_equal(got, s["out/flwdir1_asc/streams_uparea_i32"])
    got = streams.upstream_area(ids, seq, d8.shape[1], transform=(30.0, 0.0, 0.0, 0.0, -20.0, 0.0), area_factor=1e4, dtype=np.float32)
    assert got.dtype == np.float32 and np.array_equal(got, s["out/flwdir1_asc/streams_uparea_f32"])
    rhine = cs.case_d8("rhine")
    flw = pfb.from_array(rhine, ftype="d8", transform=cs.RHINE_TRANSFORM, latlon=True)
    got = streams.upstream_area(flw.idxs_ds, flw.idxs_seq, rhine.shape[1], latlon=True, transform=flw.transform, area_factor=1e6)
    assert got.dtype == np.float64 and cs.sha(got) == cs.hashes()["rhine"]["streams_uparea_km2"]
    # index lists
    rnd = cs.case_d8("random48x61")  # has loops
    rids = cs.golden("random48x61", "idxs_ds")
    rank = cs.golden("random48x61", "rank").ravel()
    nup = cs.golden("random48x61", "n_upstream").ravel()
    assert np.array_equal(core.loop_indices(rids, shape=rnd.shape), np.flatnonzero(rank == -1)) and (rank == -1).any()
    assert np.array_equal(core.headwater_indices(rids, shape=rnd.shape), np.flatnonzero(nup == 0))
    assert np.array_equal(core.confluence_indices(rids, shape=rnd.shape), np.flatnonzero(nup > 1))
    assert core.loop_indices(rids, shape=rnd.shape).dtype == rids.dtype


def test_upstream_matrix_and_dump_load(pfb, tmp_path):
    """core.upstream_matrix (core.py:67-84) against its definition on a raster with loops, and the pickle round trip
    FlwdirRaster.dump / load (flwdir.py:290-306, pyflwdir.py:362-373)."""
    from pyflwdir_b200 import core

    d8 = cs.case_d8("random48x61")
    ids = cs.golden("random48x61", "idxs_ds")
    got = core.upstream_matrix(ids, shape=d8.shape)
    nup = cs.golden("random48x61", "n_upstream").ravel()
    d = int(nup.max())
    want = np.full((ids.size, d), -1, dtype=ids.dtype)
    fill = np.zeros(ids.size, np.int64)
    for i0 in range(ids.size):  # the reference's loop: upstream cells land in ascending index
        ds = ids[i0]
        if ds != i0 and ds != -1:
            want[ds, fill[ds]] = i0
            fill[ds] += 1
    assert got.dtype == ids.dtype and got.shape == want.shape and np.array_equal(got, want)
    rhine = cs.case_d8("rhine")
    flw = pfb.from_array(rhine, ftype="d8", transform=cs.RHINE_TRANSFORM, latlon=True)
    seq = flw.idxs_seq
    fn = str(tmp_path / "flw.pkl")
    flw.dump(fn)
    flw2 = pfb.FlwdirRaster.load(fn)
    assert flw2.shape == flw.shape and flw2.ftype == "d8" and flw2.latlon and tuple(flw2.transform) == tuple(flw.transform)
    assert np.array_equal(flw2.idxs_ds, flw.idxs_ds) and np.array_equal(flw2.idxs_pit, flw.idxs_pit)
    assert np.array_equal(flw2.idxs_seq, seq) and flw2.nnodes == flw.nnodes
    assert np.array_equal(flw2.upstream_area(), flw.upstream_area()) and cs.sha(flw2.basins()) == cs.hashes()["rhine"]["basins"]


def test_module_level_seq_order_is_honoured_or_refused(pfb):
    """The module-level mirrors sweep the "walk" order. A caller-supplied `seq` in another valid order (the reference's
    fixtures use np.argsort(rank)) is accepted where the result cannot depend on it (integer accumulation, Strahler, HAND)
    and refused where it would (float sums, label numbering) -- never silently answered for a different order."""
    from pyflwdir_b200 import basins, dem, streams

    d8 = cs.case_d8("flwdir1_asc")
    ids, walk = cs.golden("flwdir1_asc", "idxs_ds"), cs.golden("flwdir1_asc", "idxs_seq")
    rank = cs.golden("flwdir1_asc", "rank").ravel()
    sort_seq = np.argsort(rank, kind="stable")[-walk.size:].astype(walk.dtype)
    assert not np.array_equal(sort_seq, walk) and np.all(np.diff(rank[sort_seq]) >= 0)
    ones = np.ones(ids.size, np.int32)
    upa = streams.accuflux(ids, sort_seq, ones, -9999, shape=d8.shape)
    assert np.array_equal(upa, oracle.streams.accuflux(ids, sort_seq, ones, -9999))  # integers: any valid order
    assert np.array_equal(streams.strahler_order(ids, sort_seq, shape=d8.shape), oracle.streams.strahler_order(ids, sort_seq))
    elev = np.random.default_rng(1).random(ids.size, dtype=np.float32)
    assert np.array_equal(dem.height_above_nearest_drain(ids, sort_seq, upa > 20, elev, shape=d8.shape),
                          oracle.dem.height_above_nearest_drain(ids, sort_seq, upa > 20, elev))
    area = np.random.default_rng(2).random(ids.size)
    assert np.array_equal(streams.accuflux(ids, walk, area, -9999, shape=d8.shape), oracle.streams.accuflux(ids, walk, area, -9999))
    with pytest.raises(NotImplementedError, match="order of the cells"):
        streams.accuflux(ids, sort_seq, area, -9999, shape=d8.shape)
    with pytest.raises(NotImplementedError, match="order of the cells"):
        basins.subbasins_streamorder(ids, sort_seq, streams.strahler_order(ids, walk, shape=d8.shape), shape=d8.shape)
