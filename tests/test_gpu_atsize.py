"""GPU parity at the sizes BASELINE.json names (8192^2, 16384^2, 32768^2, 65536^2).

Three layers:
  * the oracle itself, directly, at 8192^2 (the largest size it finishes in seconds) and 4096^2 for Strahler / HAND /
    float accuflux;
  * pfd_verify_* (csrc/pfd_verify.cuh): an independent element-wise pass that re-evaluates the reference's per-cell
    statement from the finished values of the cell's graph neighbours and compares bit for bit. Every output of the
    path is a deterministic function of its neighbours' final values (rank[i] = rank[ds] + 1, hand[i] = hand[ds] + dz,
    accu[i] = running sum over the upstream cells in descending index, ...), so zero violations proves the whole array
    at any size. test_verify_* pins the verifier itself against the oracle: it accepts the oracle's arrays and flags
    single corrupted cells in every category;
  * agreement of the two independent device implementations (tile solver vs level-synchronous BFS + sweeps) by checksum."""
import ctypes as C

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _lib():
    from pyflwdir_b200 import _lib as L

    return L, L.lib()


def _verify_flow(L, l, h, ids, idt, rk, upa, bas):
    bad = (C.c_int64 * 5)()
    L.check(l.pfd_verify_flow(h, ids, idt, rk, upa, bas, bad), h)
    return list(bad)


def test_verify_accepts_oracle_and_flags_corruption():
    """The verifier against the oracle: 0 violations on the oracle's own arrays (sea, forced pits, loops from random
    codes), > 0 in exactly the corrupted category when one cell of one array is changed."""
    L, l = _lib()
    from pyflwdir_b200 import _device

    rng = np.random.default_rng(4)
    z = oracle.synth_elevation(384, 520, seed=8)
    d8a = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.06)))
    codes = np.array([0, 1, 2, 4, 8, 16, 32, 64, 128, 247, 255], np.uint8)
    d8b = codes[rng.integers(0, 9, size=(200, 300))]  # random legal codes: loops and trees hanging on loops
    for d8 in (d8a, d8b):
        dev = _device.DeviceGraph(0)
        dev.parse_d8(d8)
        h = dev._h
        ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
        seq = oracle.core.idxs_seq(ids, pits)
        rk = oracle.core.rank(ids)[0]
        upa = oracle.streams.accuflux(ids, seq, np.ones(d8.size, np.int32), -9999)
        upa[ids == -1] = -9999
        bas = oracle.basins.basins(ids, pits, seq)
        i32 = L.dtype_code(np.int32)
        assert _verify_flow(L, l, h, L.ptr(ids), i32, L.ptr(rk), L.ptr(upa), L.ptr(bas)) == [0, 0, 0, 0, 0]
        cell = int(seq[seq.size // 2])
        for k, arr in enumerate((ids, rk, bas, upa)):
            keep = arr[cell]
            arr[cell] = keep + 1
            bad = _verify_flow(L, l, h, L.ptr(ids), i32, L.ptr(rk), L.ptr(upa), L.ptr(bas))
            assert bad[k] > 0, (k, bad)
            arr[cell] = keep
        # pit numbering
        bas2 = bas.copy()
        if pits.size > 1:
            a, b = bas2 == 1, bas2 == 2
            bas2[a], bas2[b] = 2, 1
            assert _verify_flow(L, l, h, None, i32, None, None, L.ptr(bas2))[4] > 0
        # Strahler, HAND, float32 / float64 / int64 accuflux
        nb = C.c_int64()
        so = oracle.streams.strahler_order(ids, seq)
        L.check(l.pfd_verify_strahler(h, None, L.ptr(so), C.byref(nb)), h)
        assert nb.value == 0
        so[cell] += 1
        L.check(l.pfd_verify_strahler(h, None, L.ptr(so), C.byref(nb)), h)
        assert nb.value > 0
        mask = np.ascontiguousarray(upa > 3).view(np.uint8)
        som = oracle.streams.strahler_order(ids, seq, mask=mask.view(np.bool_))
        L.check(l.pfd_verify_strahler(h, L.ptr(mask), L.ptr(som), C.byref(nb)), h)
        assert nb.value == 0
        elev = np.ascontiguousarray(rng.random(d8.size, dtype=np.float32) * 100)
        drain = np.ascontiguousarray(upa > 30).view(np.uint8)
        for ev in (elev, elev.astype(np.float64)):
            hand = oracle.dem.height_above_nearest_drain(ids, seq, drain.view(np.bool_), ev)
            L.check(l.pfd_verify_hand(h, L.ptr(drain), L.ptr(ev), L.dtype_code(ev.dtype), L.ptr(hand), C.byref(nb)), h)
            assert nb.value == 0
            hand[cell] = np.nextafter(hand[cell], np.inf)
            L.check(l.pfd_verify_hand(h, L.ptr(drain), L.ptr(ev), L.dtype_code(ev.dtype), L.ptr(hand), C.byref(nb)), h)
            assert nb.value > 0
        for data, nodata in ((elev, -9999.0), (elev.astype(np.float64), -9999.0), ((elev * 10).astype(np.int64), -9999)):
            data = data.copy()
            data[rng.integers(0, d8.size, 50)] = nodata
            accu = oracle.streams.accuflux(ids, seq, data, nodata)
            is_int = int(np.issubdtype(data.dtype, np.integer))
            args = (L.ptr(data), L.dtype_code(data.dtype), C.c_double(float(nodata)), int(nodata), is_int)
            L.check(l.pfd_verify_accuflux(h, *args, L.ptr(accu), C.byref(nb)), h)
            assert nb.value == 0, data.dtype
            accu[cell] = accu[cell] + 1
            L.check(l.pfd_verify_accuflux(h, *args, L.ptr(accu), C.byref(nb)), h)
            assert nb.value > 0
        dev.close()


def test_oracle_parity_8192():
    """BASELINE configs[1] size, directly against the oracle (about 10 s of CPU): idxs_ds, pits, rank, upstream area,
    basins of the benchmarked call pfd_d8_flow_all, plus Strahler order."""
    import bench

    L, l = _lib()
    w = bench.Workload(8192, 0, 0)
    w.step_resident()
    n = w.cells
    d8 = np.empty((8192, 8192), np.uint8)
    w.ck(l.pfd_memcpy(w.h, L.ptr(d8), w.d8_dev, n))
    got = []
    for p, dt in zip(w.out_dev, (np.int32, np.int32, np.int32, np.uint32)):
        a = np.empty(n, dt)
        w.ck(l.pfd_memcpy(w.h, L.ptr(a), p, n * 4))
        got.append(a)
    so_dev = w.dev_alloc(n)
    w.ck(l.pfd_strahler(w.h, None, so_dev))
    so = np.empty(n, np.uint8)
    w.ck(l.pfd_memcpy(w.h, L.ptr(so), so_dev, n))
    w.ck(l.pfd_dev_free(w.h, so_dev))
    w.free()
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    assert np.array_equal(got[0], ids)
    assert np.array_equal(got[1], oracle.core.rank(ids)[0])
    seq = oracle.core.idxs_seq(ids, pits)
    upa = oracle.streams.accuflux(ids, seq, np.ones(n, np.int32), -9999)
    assert np.array_equal(got[2], upa)
    assert np.array_equal(got[3], oracle.basins.basins(ids, pits, seq))
    assert np.array_equal(so, oracle.streams.strahler_order(ids, seq))


@pytest.mark.parametrize("size", [32768, 65536])
def test_flow_all_verified_at_size(size):
    """The benchmarked call at the north-star sizes (65536^2 = 2^32 cells: int64 idxs_ds, cell_t at its limit): every
    output satisfies its defining recurrence in every cell; at 32768^2 the BFS + sweeps solver agrees by checksum."""
    import bench

    L, l = _lib()
    w = bench.Workload(size, 1 if size == 32768 else 2, 0)
    assert np.dtype(w.idx_dtype) == (np.int32 if size == 32768 else np.int64)
    w.step_resident()
    o = w.out_dev
    bad = _verify_flow(L, l, w.h, o[0], L.dtype_code(w.idx_dtype), o[1], o[2], o[3])
    assert bad == [0, 0, 0, 0, 0], bad
    nv, npit = int(l.pfd_get_info(w.h, b"n_valid")), int(l.pfd_get_info(w.h, b"n_pits"))
    assert nv == w.cells and npit > 0
    if size == 32768:
        tiles = w.checksums()
        w.ck(l.pfd_set_option(w.h, b"tiles", 0))
        w.step_resident()
        assert w.checksums() == tiles, "tile solver and BFS + sweeps disagree"
    w.free()


def _synth(size, seed, sea_q=None):
    from pyflwdir_b200 import _device

    L, l = _lib()
    dev = _device.DeviceGraph(0)
    z = np.empty((size, size), np.float32)
    d8 = np.empty((size, size), np.uint8)
    L.check(l.pfd_synth_elevation(dev._h, size, size, size, max(1, int(np.log2(size)) - 2), seed, L.ptr(z)), dev._h)
    sea = float(np.quantile(z[::16, ::16], sea_q)) if sea_q else -np.inf
    L.check(l.pfd_synth_d8(dev._h, L.ptr(z), size, size, C.c_float(sea), L.ptr(d8)), dev._h)
    dev.close()
    return d8, z


def test_strahler_hand_accuflux_oracle_4096():
    """BASELINE configs 3 / 5 outputs against the oracle at 4096^2 (with a sea => nodata and forced pits)."""
    import pyflwdir_b200 as pfb

    d8, z = _synth(4096, 3, sea_q=0.03)
    flw = pfb.from_array(d8, ftype="d8")
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    seq = oracle.core.idxs_seq(ids, pits)
    upa = flw.upstream_area()
    assert np.array_equal(flw.stream_order().ravel(), oracle.streams.strahler_order(ids, seq))
    mask = upa > 10
    flw_nc = pfb.from_array(d8, ftype="d8", cache=False)  # (the reference caches strord irrespective of the mask)
    assert np.array_equal(flw_nc.stream_order(mask=mask).ravel(), oracle.streams.strahler_order(ids, seq, mask=mask.ravel()))
    del flw_nc
    drain = upa > 1000
    assert np.array_equal(flw.hand(drain, z).ravel(), oracle.dem.height_above_nearest_drain(ids, seq, drain.ravel(), z.ravel()))
    assert np.array_equal(flw.hand(np.zeros_like(drain), z.astype(np.float64)).ravel(),
                          oracle.dem.height_above_nearest_drain(ids, seq, np.zeros(d8.size, np.bool_), z.ravel().astype(np.float64)))
    area = (np.abs(z) + np.float32(0.5)).astype(np.float32)
    assert np.array_equal(flw.accuflux(area).ravel(), oracle.streams.accuflux(ids, seq, area.ravel(), -9999))
    a64 = area.astype(np.float64) * 1e-3
    assert np.array_equal(flw.accuflux(a64).ravel(), oracle.streams.accuflux(ids, seq, a64.ravel(), -9999))


@pytest.mark.parametrize("size", [16384, 32768])
def test_sweeps_verified_at_size(size):
    """BASELINE config 3 (32768^2: Strahler) and config 5 (16384^2: accuflux -> stream mask -> HAND), device-resident,
    every cell checked against its defining recurrence; float32 accuflux too."""
    import bench

    L, l = _lib()
    w = bench.Workload(size, 1 if size == 32768 else 3, 0)
    h, n = w.h, w.cells
    w.ck(l.pfd_set_option(h, b"tile_sweeps", 2))  # the tile-dataflow sweeps (the level replays are covered at 4096^2)
    w.step_resident()
    nb = C.c_int64()
    so_dev = w.dev_alloc(n)
    w.ck(l.pfd_strahler(h, None, so_dev))
    w.ck(l.pfd_verify_strahler(h, None, so_dev, C.byref(nb)))
    assert nb.value == 0
    w.ck(l.pfd_dev_free(h, so_dev))
    if size == 16384:
        z_dev = w.dev_alloc(n * 4)
        w.ck(l.pfd_synth_elevation(h, size, size, size, bench.octaves_for(size), 3, z_dev))
        upa = np.empty(n, np.int32)
        w.ck(l.pfd_memcpy(h, L.ptr(upa), w.out_dev[2], n * 4))
        drain = np.ascontiguousarray(upa > 1000).view(np.uint8)
        drain_dev = w.dev_alloc(n)
        w.ck(l.pfd_memcpy(h, drain_dev, L.ptr(drain), n))
        hand_dev = w.dev_alloc(n * 8)
        f32 = L.dtype_code(np.float32)
        w.ck(l.pfd_hand(h, drain_dev, z_dev, f32, hand_dev))
        w.ck(l.pfd_verify_hand(h, drain_dev, z_dev, f32, hand_dev, C.byref(nb)))
        assert nb.value == 0
        acc_dev = w.out_dev[1]
        w.ck(l.pfd_accuflux(h, z_dev, f32, C.c_double(-9999.0), 0, 0, 0, acc_dev))
        w.ck(l.pfd_verify_accuflux(h, z_dev, f32, C.c_double(-9999.0), 0, 0, acc_dev, C.byref(nb)))
        assert nb.value == 0
    w.free()
