"""CPU tests: the C oracle (oracle/pfd_oracle.c) against the golden vectors that tests/golden/make_golden.py
generated from the REAL reference (numba) -- this is what pins the oracle. When /root/reference is mounted
(build container) the oracle is additionally checked against the live reference on fresh random inputs."""
import numpy as np
import pytest

import _cases as cs
import oracle
from oracle import reference


@pytest.mark.parametrize("name", cs.SMALL_CASES + ["synth512x768"])
def test_oracle_matches_reference_golden(name, oracle_lib):
    d8 = cs.case_d8(name)
    aux = cs.case_inputs(name, d8, cs.case_seed(name))
    out = cs.run_oracle_case(d8, aux, area=np.ones(d8.size, dtype=np.float32))
    for key, val in out.items():
        cs.check(name, key, val)


def test_oracle_rhine_golden(oracle_lib):
    from pyflwdir_b200 import gis_utils

    d8 = cs.case_d8("rhine")
    aux = cs.case_inputs("rhine", d8, cs.case_seed("rhine"))
    for k, v in aux.items():
        assert cs.sha(v) == cs.hashes()["rhine"]["_aux"][k], f"aux input {k} drifted"
    area = gis_utils.area_grid(gis_utils.Affine(*cs.RHINE_TRANSFORM), d8.shape, latlon=True)
    assert np.array_equal(area[:, 0], cs.small()["out/rhine/area_col0"])
    out = cs.run_oracle_case(d8, aux, area=area.ravel(), transform=cs.RHINE_TRANSFORM, latlon=True)
    for key, val in out.items():
        cs.check("rhine", key, val)
    assert int(out["rank"].max()) == cs.hashes()["rhine"]["_max_rank"] == 1674


def test_oracle_drdc_table(oracle_lib):
    dr, dc = oracle.core_d8.drdc_table()
    g = cs.small()["out/drdc_table"]
    assert np.array_equal(dr, g[:, 0]) and np.array_equal(dc, g[:, 1])


@pytest.mark.parametrize("dt", [np.uint32, np.int64])
def test_oracle_index_dtypes(dt, oracle_lib):
    d8 = cs.case_d8("flwdir_asc")
    ids, pits, n = oracle.core_d8.from_array(d8, dtype=dt)
    name = np.dtype(dt).name
    assert np.array_equal(ids, cs.small()[f"out/flwdir_asc/idxs_ds_{name}"]) and ids.dtype == dt
    assert np.array_equal(pits, cs.small()[f"out/flwdir_asc/idxs_pit_{name}"])
    assert n == 407


def test_loop_kat(oracle_lib):
    """SURVEY.md §8c hand-made 3x3 raster with a 2-cell loop."""
    d8 = np.array([[1, 16, 4], [1, 4, 4], [247, 0, 16]], dtype=np.uint8)
    ids, pits, n = oracle.core_d8.from_array(d8, dtype=np.int32)
    assert ids.tolist() == [1, 0, 5, 4, 7, 8, -1, 7, 7] and pits.tolist() == [7] and n == 8
    rank, nn = oracle.core.rank(ids)
    assert rank.tolist() == [-1, -1, 3, 2, 1, 2, -9999, 0, 1] and nn == 6
    seq = oracle.core.idxs_seq(ids, pits)
    assert seq.tolist() == [7, 4, 8, 3, 5, 2]
    upa = oracle.streams.accuflux(ids, seq, np.ones(9, np.int32), -9999)
    assert upa.tolist() == [1, 1, 1, 1, 2, 2, 1, 6, 3]
    assert oracle.basins.basins(ids, pits, seq).tolist() == [0, 0, 1, 1, 1, 1, 0, 1, 1]
    assert oracle.streams.strahler_order(ids, seq).tolist() == [0, 0, 1, 1, 1, 1, 0, 2, 1]
    assert oracle.core.upstream_count(ids).tolist() == [1, 1, 0, 0, 1, 1, -9, 2, 1]
    hand = oracle.dem.height_above_nearest_drain(ids, seq, np.zeros(9, bool), np.arange(9, dtype=np.float32))
    assert hand.tolist() == [-9999, -9999, -5, -4, -3, -2, -9999, 0, 1]


@pytest.mark.skipif(not reference.available(), reason="reference not mounted (GPU box)")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_vs_live_reference(seed, oracle_lib):
    """Fresh random rasters (legal codes, loops, nodata) through the real numba reference and the oracle."""
    pf = reference.load()
    rng = np.random.default_rng(100 + seed)
    if seed == 0:
        legal = np.array([32, 64, 128, 16, 0, 1, 8, 4, 2, 247, 255], dtype=np.uint8)
        d8 = legal[rng.integers(0, legal.size, size=(37, 53))]
    else:
        z = oracle.synth_elevation(150 + 13 * seed, 211, seed=50 + seed)
        d8 = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.1)))
    aux = cs.case_inputs("live", d8, 77 + seed)
    ref = cs.run_api_case(pf, d8, aux)  # every entry point of the object API, on the real reference
    out = cs.run_oracle_case(d8, aux, area=np.ones(d8.size, dtype=np.float32))
    assert len(out) > 150
    for key, val in out.items():
        want = np.asarray(ref[key])
        assert np.asarray(val).dtype == want.dtype, f"{key}: dtype {np.asarray(val).dtype} != {want.dtype}"
        assert np.array_equal(val, want, equal_nan=True), f"{key} differs from the live reference"


def test_oracle_ldd_golden(oracle_lib):
    """core_ldd.from_array / to_array restatement against the reference on the LDD version of flwdir1.asc."""
    ldd = cs.small()["in/ldd_flwdir1/ldd"]
    ids, pits, n = oracle.core_ldd.from_array(ldd, dtype=np.int32)
    assert np.array_equal(ids, cs.small()["out/ldd_flwdir1/idxs_ds"]) and ids.dtype == np.int32
    assert np.array_equal(pits, cs.small()["out/ldd_flwdir1/idxs_pit"])
    assert np.array_equal(oracle.core_ldd.to_array(ids, ldd.shape), cs.small()["out/ldd_flwdir1/to_array"])
    assert np.array_equal(oracle.core_d8.to_array(ids, ldd.shape), cs.small()["out/ldd_flwdir1/to_array_d8"])


def test_oracle_nextxy_golden():
    """core_nextxy.from_array / to_array against the reference's outputs (incl. -10 pits and a nexty-only pit)."""
    s = cs.small()
    nxy = s["in/nextxy_flwdir1/nextxy"]
    ids, pits, n = oracle.core_nextxy.from_array(nxy, dtype=np.int32)
    assert np.array_equal(ids, s["out/nextxy_flwdir1/idxs_ds"]) and ids.dtype == s["out/nextxy_flwdir1/idxs_ds"].dtype
    assert np.array_equal(pits, s["out/nextxy_flwdir1/idxs_pit"])
    assert np.array_equal(oracle.core_nextxy.to_array(ids, nxy.shape[1:]), s["out/nextxy_flwdir1/to_array"])
    assert np.array_equal(oracle.core_d8.to_array(ids, nxy.shape[1:]), s["out/nextxy_flwdir1/to_array_d8"])
    # the reference orders nextxy rasters with np.argsort(rank) (pyflwdir.py:296): same cells, rank-sorted, but the
    # order inside a rank level is numpy's unstable sort, not core.idxs_seq
    seq, want = oracle.core.idxs_seq(ids, pits), s["out/nextxy_flwdir1/idxs_seq"]
    rank = oracle.core.rank(ids)[0]
    assert np.array_equal(np.sort(seq), np.sort(want)) and np.all(np.diff(rank[want]) >= 0) and np.all(np.diff(rank[seq]) >= 0)
    masked = nxy.copy()
    masked[:, 100:, 150:] = -9999
    ids_m, pits_m, _ = oracle.core_nextxy.from_array(masked, dtype=np.int32)
    assert np.array_equal(ids_m, s["out/nextxy_flwdir1/masked_idxs_ds"]) and np.array_equal(pits_m, s["out/nextxy_flwdir1/masked_idxs_pit"])


def _fill_golden():
    import os

    return dict(np.load(os.path.join(cs.GOLDEN, "fill_cases.npz")))


@pytest.mark.parametrize("name", sorted(cs.fill_cases()))
def test_oracle_fill_depressions_golden(name, oracle_lib):
    """oracle/pfd_oracle_fill.inc against the reference's dem.fill_depressions outputs (tests/golden/make_golden_fill.py)."""
    g = _fill_golden()
    a, kw = cs.fill_cases()[name]
    filled, d8 = oracle.dem.fill_depressions(a.copy(), **kw)
    assert filled.dtype == g[f"{name}/filled"].dtype
    assert np.array_equal(filled, g[f"{name}/filled"], equal_nan=True), f"{name}: filled elevation differs from the reference"
    assert np.array_equal(d8, g[f"{name}/d8"]), f"{name}: d8 differs from the reference"


def test_oracle_fill_depressions_mid_hash(oracle_lib):
    import json
    import os

    a, kw = cs.fill_mid_case()
    want = json.load(open(os.path.join(cs.GOLDEN, "fill_hashes.json")))["synth512x768"]
    assert cs.sha(a) == want["input"], "synthetic input drifted"
    filled, d8 = oracle.dem.fill_depressions(a, **kw)
    assert cs.sha(filled) == want["filled"] and cs.sha(d8) == want["d8"]


@pytest.mark.skipif(not reference.available(), reason="/root/reference not mounted")
def test_oracle_fill_depressions_vs_live_reference(oracle_lib):
    """fresh random rasters (ties, holes, every outlet mode, both connectivities, max_depth) against the numba reference"""
    reference.load()
    from pyflwdir import dem as rdem

    rng = np.random.default_rng(123)
    for trial in range(24):
        nr, nc = int(rng.integers(3, 30)), int(rng.integers(3, 30))
        kind = trial % 4
        a = [rng.random((nr, nc)), rng.random((nr, nc)).astype(np.float32) * 9, rng.integers(0, 5, (nr, nc)).astype(np.float32),
             rng.integers(0, 20, (nr, nc)).astype(np.int32)][kind]
        if kind != 3 and trial % 3 == 0:
            a[rng.random((nr, nc)) < 0.15] = -9999
        for kw in (dict(), dict(outlets="min"), dict(connectivity=4), dict(idxs_pit=np.array([0, a.size // 2]))):
            r = rdem.fill_depressions(a.copy(), **kw)
            o = oracle.dem.fill_depressions(a.copy(), **kw)
            assert np.array_equal(r[0], o[0]) and np.array_equal(r[1], o[1]) and r[0].dtype == o[0].dtype, (trial, kw)
    a = rng.random((20, 24)).astype(np.float32) * 10
    a[0, :] += 100; a[-1, :] += 100; a[:, 0] += 100; a[:, -1] += 100  # the revisit of max_depth stays inside the raster
    for md in (0.5, 2.0):
        r = rdem.fill_depressions(a.copy(), max_depth=md)
        o = oracle.dem.fill_depressions(a.copy(), max_depth=md)
        assert np.array_equal(r[0], o[0]) and np.array_equal(r[1], o[1])
