"""Fused-parse headline path (pfd_d8_flow_all on device-resident buffers: phase A parses the raw D8 codes, phase C
writes idxs_ds) against the oracle, against the separate-parse path, and followed by the entry points that need the
lazily derived upstream mask."""
import numpy as np
import pytest

import oracle
import _cases

pytestmark = pytest.mark.gpu


def _oracle_all(d8, idx_dtype=np.int32):
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=idx_dtype)
    seq = oracle.core.idxs_seq(ids, pits)
    rank = oracle.core.rank(ids)[0]
    upa = oracle.streams.accuflux(ids, seq, np.ones(d8.size, np.int32), -9999)
    upa[ids == ids.dtype.type(-1)] = -9999
    bas = oracle.basins.basins(ids, pits, seq)
    return ids, pits, seq, rank, upa, bas


def _rasters():
    rng = np.random.default_rng(11)
    out = {}
    sm = _cases.small()
    for name in _cases.SMALL_CASES:
        out[name] = sm[f"in/{name}/d8"]
    z = oracle.synth_elevation(300, 420, seed=5)
    out["synth300x420_sea"] = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.05)))
    z = oracle.synth_elevation(257, 131, seed=6)  # unaligned width, partial tiles on both axes
    out["synth257x131"] = oracle.synth_d8(z, sea_level=-np.inf)
    codes = np.array([0, 1, 2, 4, 8, 16, 32, 64, 128, 247, 255], np.uint8)
    out["random130x67"] = codes[rng.integers(0, codes.size, size=(130, 67))]  # loops, forced pits, nodata
    out["random64x64"] = codes[rng.integers(0, codes.size, size=(64, 64))]
    out["random200x192"] = codes[rng.integers(0, codes.size, size=(200, 192))]  # aligned width, partial tile rows
    z = oracle.synth_elevation(640, 1024, seed=8)  # more tiles than one wave of a small grid
    out["synth640x1024"] = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.02)))
    out["row1x200"] = codes[rng.integers(0, codes.size, size=(1, 200))]
    out["col200x1"] = codes[rng.integers(0, codes.size, size=(200, 1))]
    return out


@pytest.mark.parametrize("name", list(_rasters().keys()))
def test_fused_flow_all_matches_oracle(name):
    from pyflwdir_b200 import _device

    d8 = np.ascontiguousarray(_rasters()[name])
    ids, pits, seq, rank, upa, bas = _oracle_all(d8)
    if pits.size == 0:
        pytest.skip("no pits")
    dev = _device.DeviceGraph(0)
    try:
        assert dev.info("fuse_parse") == 1
        got = dev.flow_all(d8, np.int32, resident=True)
        assert dev.info("have_upmask") == 0, "the fused path should not have run the separate parse pass"
        assert np.array_equal(got[0], ids), "idxs_ds"
        assert np.array_equal(got[1], rank), "rank"
        assert np.array_equal(got[2], upa), "upstream_area"
        assert np.array_equal(got[3], bas), "basins"
        assert dev.n_pits == pits.size and dev.n_valid == int((d8 != 247).sum())
        # the handle is fully usable afterwards: pits, exact ordering and a sweep that needs the upstream mask
        from pyflwdir_b200 import _lib
        assert np.array_equal(dev.fetch(_lib.ARR_PITS, np.int32), pits)
        assert np.array_equal(dev.fetch(_lib.ARR_SEQ, np.int32), seq)
        assert dev.info("have_upmask") == 1
        assert np.array_equal(dev.strahler(), oracle.streams.strahler_order(ids, seq))
        nup = oracle.core.upstream_count(ids)
        assert np.array_equal(dev.fetch(_lib.ARR_N_UPSTREAM), nup)
        # and identical to the separate-parse path on the same handle
        dev.set_option("fuse_parse", 0)
        ref = dev.flow_all(d8, np.int32, resident=True)
        assert dev.info("have_upmask") == 1
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)
    finally:
        dev.close()


@pytest.mark.parametrize("idx_dtype", [np.uint32, np.int64])
def test_fused_idx_dtypes(idx_dtype):
    from pyflwdir_b200 import _device

    z = oracle.synth_elevation(200, 264, seed=9)
    d8 = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.1)))
    ids = oracle.core_d8.from_array(d8, dtype=idx_dtype)[0]
    dev = _device.DeviceGraph(0)
    try:
        got = dev.flow_all(d8, idx_dtype, resident=True)
        assert got[0].dtype == np.dtype(idx_dtype) and np.array_equal(got[0], ids)
    finally:
        dev.close()


def test_fused_rejects_illegal_codes_and_pitless_rasters():
    from pyflwdir_b200 import _device

    dev = _device.DeviceGraph(0)
    try:
        d8 = np.full((70, 90), 1, np.uint8)
        d8[5, 5] = 3  # not a D8 code
        with pytest.raises(ValueError, match="D8 code set"):
            dev.flow_all(d8, np.int32, resident=True)
        loop = np.array([[1, 16]], np.uint8)  # two cells pointing at each other: no pit
        with pytest.raises(ValueError, match="no pits found"):
            dev.flow_all(loop, np.int32, resident=True)
        # the handle recovers
        ok = np.zeros((3, 3), np.uint8)
        got = dev.flow_all(ok, np.int32, resident=True)
        assert np.array_equal(got[0], np.arange(9))
    finally:
        dev.close()
