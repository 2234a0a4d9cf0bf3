"""GPU parity of dem.fill_depressions / from_dem (csrc/pfd_fill.cuh) -- SURVEY.md §8f-3: bit-exact filled elevation AND D8
codes against the goldens generated from the reference (tests/golden/make_golden_fill.py), against the oracle on fresh
rasters at sizes it finishes in seconds, and the properties that hold at any size."""
import json
import os

import numpy as np
import pytest

import _cases as cs
import oracle
import pyflwdir_b200 as pfb
from pyflwdir_b200 import dem

pytestmark = pytest.mark.gpu


def _golden():
    return dict(np.load(os.path.join(cs.GOLDEN, "fill_cases.npz")))


@pytest.mark.parametrize("name", sorted(cs.fill_cases()))
def test_fill_depressions_golden(name):
    g = _golden()
    a, kw = cs.fill_cases()[name]
    filled, d8 = dem.fill_depressions(a.copy(), **kw)
    assert filled.dtype == g[f"{name}/filled"].dtype and filled.shape == a.shape and d8.dtype == np.uint8
    assert np.array_equal(filled, g[f"{name}/filled"], equal_nan=True), f"{name}: filled elevation differs from the reference"
    assert np.array_equal(d8, g[f"{name}/d8"]), f"{name}: d8 differs from the reference"


def test_fill_depressions_mid_hash():
    a, kw = cs.fill_mid_case()
    want = json.load(open(os.path.join(cs.GOLDEN, "fill_hashes.json")))["synth512x768"]
    filled, d8 = dem.fill_depressions(a, **kw)
    assert cs.sha(filled) == want["filled"] and cs.sha(d8) == want["d8"]


def test_fill_depressions_random_vs_oracle():
    """many small rasters: ties (flats), holes, outlet modes, connectivities, float32 / float64 / integers"""
    rng = np.random.default_rng(2024)
    for trial in range(60):
        nr, nc = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        kind = trial % 5
        a = [rng.random((nr, nc)), rng.random((nr, nc)).astype(np.float32) * 9 + 1, rng.integers(0, 4, (nr, nc)).astype(np.float32),
             rng.integers(0, 20, (nr, nc)).astype(np.int32), np.round(rng.random((nr, nc)) * 5, 1)][kind]
        if kind != 3 and trial % 3 == 0:
            a[rng.random((nr, nc)) < 0.2] = -9999
        for kw in (dict(), dict(outlets="min"), dict(connectivity=4), dict(idxs_pit=np.array([0, a.size // 2, a.size - 1]))):
            if "idxs_pit" in kw and np.any(a.ravel()[kw["idxs_pit"]] == -9999):
                continue
            want = oracle.dem.fill_depressions(a.copy(), **kw)
            got = dem.fill_depressions(a.copy(), **kw)
            assert got[0].dtype == want[0].dtype
            assert np.array_equal(got[0], want[0]), (trial, a.shape, a.dtype, kw, "filled")
            assert np.array_equal(got[1], want[1]), (trial, a.shape, a.dtype, kw, "d8")


def test_fill_depressions_terrain_vs_oracle_and_properties():
    """2048^2 synthetic terrain with a sea (tiles + many passes): oracle comparison, then the size-independent properties:
    filling is idempotent (float64), never lowers a cell, every valid cell drains to an outlet (no pits besides the outlets)."""
    z = oracle.synth_elevation(2048, 2048, seed=9)
    a = np.where(z < np.quantile(z, 0.05), np.float32(-9999.0), z * np.float32(700.0)).astype(np.float32)
    want = oracle.dem.fill_depressions(a)
    filled, d8 = dem.fill_depressions(a)
    assert np.array_equal(filled, want[0]) and np.array_equal(d8, want[1])
    valid = a != -9999
    assert np.all(filled[valid] >= a[valid]) and np.array_equal(filled[~valid], a[~valid])
    f64 = dem.fill_depressions(a.astype(np.float64))[0]
    assert np.array_equal(dem.fill_depressions(f64)[0], f64)  # (float32 raises drift by ulps, so only float64 is idempotent)
    flw = pfb.from_array(d8, ftype="d8", check_ftype=False)
    assert flw.isvalid
    assert np.all(flw.rank.ravel()[valid.ravel()] >= 0)  # every valid cell reaches a pit
    pits = flw.idxs_pit
    edge = np.zeros_like(valid)
    edge[0, :] = edge[-1, :] = edge[:, 0] = edge[:, -1] = True
    inv = ~valid
    near_nodata = np.zeros_like(valid)
    near_nodata[1:-1, 1:-1] = (inv[:-2, :-2] | inv[:-2, 1:-1] | inv[:-2, 2:] | inv[1:-1, :-2] | inv[1:-1, 2:] | inv[2:, :-2]
                               | inv[2:, 1:-1] | inv[2:, 2:])
    assert np.all((edge | near_nodata).ravel()[pits]), "a pit away from the edge of the valid cells survived the fill"


def test_from_dem_matches_reference_fixture():
    """the reference's own fixture (tests/conftest.py:57-60): from_dem(np.random.rand(15, 10)) with seed 2345"""
    np.random.seed(2345)
    flw = pfb.from_dem(np.random.rand(15, 10))
    assert np.array_equal(flw.to_array("d8"), _golden()["rand15x10/d8"])
    np.random.seed(2345)
    flw = pfb.from_dem(np.random.rand(15, 10), outlets="min")
    assert np.array_equal(flw.to_array("d8"), _golden()["rand15x10_min/d8"])
    assert flw.idxs_pit.size == 1


def test_fill_depressions_errors():
    a = np.random.default_rng(1).random((12, 13)).astype(np.float32)
    with pytest.raises(ValueError, match="No initial outlet cells found"):
        dem.fill_depressions(a, elv_max=-5.0)
    with pytest.raises(ValueError, match="idxs_pit"):
        dem.fill_depressions(a, idxs_pit=np.array([a.size + 3]))


def test_fill_depressions_float32_key_drift():
    """float32 raises that miss the pour level (mixed signs / magnitudes): the reference's keys drift inside the lakes and the
    replay has to follow them -- band widening included (the second raster drifts by thousands of float32 steps)."""
    b = np.full((5, 5), 10.0, dtype=np.float32)
    b[2, 2] = np.float32(-0.3)
    b[0, 0] = np.float32(0.1)
    b[1, 1] = np.float32(0.1)
    got, want = dem.fill_depressions(b), oracle.dem.fill_depressions(b)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    rng = np.random.default_rng(5)
    for trial in range(6):
        z = oracle.synth_elevation(200, 240, seed=40 + trial) * np.float32(700.0)
        z = (z - np.float32(np.median(z))).astype(np.float32)  # lakes around zero with deep negative bottoms
        if trial % 2:
            z[rng.random(z.shape) < 0.05] = -9999.0
        want = oracle.dem.fill_depressions(z)
        g = __import__("pyflwdir_b200")._device.DeviceGraph(0)
        got = g.fill_depressions(z)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (trial, g.fill_stats)
        g.close()


def test_fill_depressions_degenerate_rasters():
    """all nodata, one cell, one row / column, one flat (a single tie component, replayed by a warp), a flat with holes"""
    cases = [np.full((7, 9), -9999.0, dtype=np.float32), np.array([[3.5]]), np.arange(17, dtype=np.float64)[None, :] % 5,
             (np.arange(23, dtype=np.float32)[:, None] * np.float32(0.5)) % np.float32(3.0), np.zeros((90, 110), dtype=np.float32),
             np.ones((64, 64), dtype=np.int32)]
    holes = np.zeros((120, 97), dtype=np.float64)
    holes[np.random.default_rng(8).random(holes.shape) < 0.3] = -9999.0
    cases.append(holes)
    for a in cases:
        for kw in (dict(), dict(connectivity=4), dict(outlets="min")):
            if kw.get("outlets") == "min" and not np.any(a != -9999.0):
                continue  # (the reference pops an empty heap there)
            want = oracle.dem.fill_depressions(a.copy(), **kw)
            got = dem.fill_depressions(a.copy(), **kw)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (a.shape, a.dtype, kw)


def test_fill_depressions_sloped_terrain_4096_vs_oracle():
    """4096^2 terrain on a regional slope (local depressions: many small tie components, the level wavefront crosses 64 tiles),
    float32 and float64, against the oracle"""
    n = 4096
    z = oracle.synth_elevation(n, n, seed=6)
    a = (z * np.float32(700.0) + np.float32(2000.0)
         + (np.arange(n, dtype=np.float32)[:, None] + np.arange(n, dtype=np.float32)[None, :]) * np.float32(1.5)).astype(np.float32)
    for arr in (a, a.astype(np.float64)):
        want = oracle.dem.fill_depressions(arr)
        got = dem.fill_depressions(arr)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), arr.dtype
