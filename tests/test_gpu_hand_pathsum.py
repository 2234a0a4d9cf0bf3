"""HAND as re-associated path sums (csrc/pfd_hand.cuh): accepted only when every cell satisfies the reference's statement bit
for bit, otherwise the hop-by-hop sweeps take over. Both outcomes, all engines, against the oracle."""
import numpy as np
import pytest

import _cases as cs
import oracle
import pyflwdir_b200 as pfb

pytestmark = pytest.mark.gpu


def _oracle_hand(d8, drain, elev):
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    seq = oracle.core.idxs_seq(ids, pits)
    return oracle.dem.height_above_nearest_drain(ids, seq, drain.ravel(), elev.ravel()).reshape(d8.shape)


@pytest.mark.parametrize("shape,seed", [((300, 420), 3), ((64, 64), 4), ((65, 129), 5), ((1, 37), 6), ((700, 5), 7), ((1030, 1100), 8)])
def test_pathsum_engine_matches_oracle(shape, seed):
    z = oracle.synth_elevation(shape[0], shape[1], seed=seed)
    d8 = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.05)))
    flw = pfb.from_array(d8, ftype="d8")
    upa = flw.upstream_area()
    for thr in (3, 40, 10**9):
        drain = upa > thr
        for elev in (z, (z * 1000).astype(np.float64), (z * np.float32(700.0) + np.float32(50.0)).astype(np.float32)):
            want = _oracle_hand(d8, drain, elev)
            got = flw.hand(drain, elev)
            assert np.array_equal(got, want), (shape, thr, elev.dtype)
            assert flw._dev.info("hand_engine") == 1, "the path sums should have been exact (and accepted) on this terrain"
    flw._dev.set_option("hand_pathsum", 0)
    drain = upa > 40
    assert np.array_equal(flw.hand(drain, z), _oracle_hand(d8, drain, z))
    assert flw._dev.info("hand_engine") in (2, 3)


def test_pathsum_with_loops_and_random_codes():
    """random legal codes: loops inside tiles and across tiles, forced pits, nodata -- cells that reach no pit stay -9999"""
    d8 = cs.case_d8("random48x61")
    rng = np.random.default_rng(12)
    legal = np.array([32, 64, 128, 16, 0, 1, 8, 4, 2, 247, 255], dtype=np.uint8)
    p = np.array([1, 1, 1, 1, 0.02, 1, 1, 1, 1, 0.1, 0.02])
    big = legal[rng.choice(legal.size, size=(200, 333), p=p / p.sum())]
    for raster in (d8, big):
        flw = pfb.from_array(raster, ftype="d8")
        elev = rng.random(raster.shape, dtype=np.float32) * np.float32(100.0)
        drain = rng.random(raster.shape) < 0.02
        want = _oracle_hand(raster, drain, elev)
        got = flw.hand(drain, elev)
        assert np.array_equal(got, want)
        assert flw._dev.info("hand_engine") == 1


def test_inexact_sums_fall_back_to_the_sweep():
    """elevations spanning 60 orders of magnitude: the float64 additions round, the re-associated sums differ from the
    reference's left fold somewhere, the check refuses them and the hop-by-hop sweep answers"""
    z = oracle.synth_elevation(260, 300, seed=21)
    d8 = oracle.synth_d8(z)
    rng = np.random.default_rng(3)
    elev = (z.astype(np.float64) + 2.0) * 10.0 ** rng.integers(-30, 30, size=z.shape)
    flw = pfb.from_array(d8, ftype="d8")
    drain = flw.upstream_area() > 200
    want = _oracle_hand(d8, drain, elev)
    got = flw.hand(drain, elev)
    assert np.array_equal(got, want)
    assert flw._dev.info("hand_engine") in (2, 3), "inexact float64 sums must not be accepted"


def test_drain_cells_above_loops_all_engines():
    """a drain cell that drains to no pit is outside the reference's sequence (-9999 there and upstream of it): path sums, the
    tile-dataflow fallback (after a rejected attempt and with the attempt switched off) and the level replay agree with the oracle"""
    rng = np.random.default_rng(77)
    legal = np.array([32, 64, 128, 16, 0, 1, 8, 4, 2, 247, 255], dtype=np.uint8)
    p = np.array([1, 1, 1, 1, 0.03, 1, 1, 1, 1, 0.1, 0.03])
    d8 = legal[rng.choice(legal.size, size=(411, 305), p=p / p.sum())]
    drain = rng.random(d8.shape) < 0.04
    exact = rng.random(d8.shape, dtype=np.float32) * np.float32(100.0)
    inexact = (rng.random(d8.shape) + 2.0) * 10.0 ** rng.integers(-25, 25, size=d8.shape)
    for elev, engines in ((exact, (1,)), (inexact, (2, 3))):
        want = _oracle_hand(d8, drain, elev)
        assert np.any((want == -9999.0) & (d8 != 247)), "the case must hold cells outside the sequence"
        flw = pfb.from_array(d8, ftype="d8", check_ftype=False)
        assert np.array_equal(flw.hand(drain, elev), want) and flw._dev.info("hand_engine") in engines
        flw._dev.set_option("hand_pathsum", 0)
        flw._dev.set_option("tile_sweeps", 2)
        assert np.array_equal(flw.hand(drain, elev), want) and flw._dev.info("hand_engine") == 2
        flw._dev.set_option("tile_sweeps", 0)
        assert np.array_equal(flw.hand(drain, elev), want) and flw._dev.info("hand_engine") == 3
