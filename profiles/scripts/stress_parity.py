"""Randomised parity sweep of the two round-2 solvers against the oracle: dem.fill_depressions (drift / band logic, ties, holes,
outlet modes) and the path-sum HAND (acceptance or fallback, loops, nodata).   python profiles/scripts/stress_parity.py [SEED] [N]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import oracle  # noqa: E402
import pyflwdir_b200 as pfb  # noqa: E402
from pyflwdir_b200 import _device  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 120
rng = np.random.default_rng(seed)
g = _device.DeviceGraph(0)
t0 = time.time()
nfill = nretry = 0
for k in range(ncase):
    nr, nc = int(rng.integers(2, 700)), int(rng.integers(2, 700))
    kind = k % 6
    z = oracle.synth_elevation(nr, nc, seed=int(rng.integers(1 << 30)))
    if kind == 0:
        a = (z * np.float32(700.0)).astype(np.float32)                                   # zero-crossing float32: drift, band retries
    elif kind == 1:
        a = (z * np.float32(700.0) + np.float32(rng.uniform(100, 5000))).astype(np.float32)
    elif kind == 2:
        a = np.round(z.astype(np.float64) * 50.0)                                          # integer-valued: flats everywhere
    elif kind == 3:
        a = np.round(z * 30).astype(np.int32)
    elif kind == 4:
        a = (z.astype(np.float64) * 1e-3 + rng.random(z.shape) * 1e-9)
    else:
        a = (np.abs(z) * np.float32(1e6) * rng.random(z.shape, dtype=np.float32)).astype(np.float32)
    if a.dtype.kind == "f" and k % 2:
        a[rng.random(a.shape) < 0.1] = -9999.0
    kw = [dict(), dict(outlets="min"), dict(connectivity=4), dict(elv_max=float(np.median(a)))][k % 4]
    try:
        want = oracle.dem.fill_depressions(a.copy(), **kw)
    except ValueError:
        continue
    got = g.fill_depressions(a.copy(), **kw)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), ("fill", k, a.shape, a.dtype, kw, g.fill_stats)
    nfill += 1
    nretry += g.fill_stats["tries"] > 1
print(f"fill_depressions: {nfill} rasters equal to the oracle ({nretry} needed a wider band), {time.time() - t0:.0f} s", flush=True)
nh = nacc = 0
legal = np.array([32, 64, 128, 16, 0, 1, 8, 4, 2, 247, 255], dtype=np.uint8)
for k in range(ncase):
    nr, nc = int(rng.integers(1, 600)), int(rng.integers(1, 600))
    z = oracle.synth_elevation(nr, nc, seed=int(rng.integers(1 << 30)))
    if k % 3 == 2:
        p = np.array([1, 1, 1, 1, 0.03, 1, 1, 1, 1, 0.1, 0.03])
        d8 = legal[rng.choice(legal.size, size=(nr, nc), p=p / p.sum())]                   # loops, forced pits, nodata
    else:
        d8 = oracle.synth_d8(z, sea_level=float(np.quantile(z, rng.uniform(0, 0.2))))
    elev = [z, z.astype(np.float64) * 1e3, (z * np.float32(1e4)).astype(np.float32),
            (z.astype(np.float64) + 2.0) * 10.0 ** rng.integers(-25, 25, size=z.shape)][k % 4]
    drain = rng.random(z.shape) < rng.uniform(0, 0.05)
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    if pits.size == 0:
        continue
    seq = oracle.core.idxs_seq(ids, pits)
    want = oracle.dem.height_above_nearest_drain(ids, seq, drain.ravel(), elev.ravel()).reshape(d8.shape)
    flw = pfb.from_array(d8, ftype="d8", check_ftype=False)
    got = flw.hand(drain, elev)
    if not np.array_equal(got, want):
        bad = np.flatnonzero(got.ravel() != want.ravel())
        eng = flw._dev.info("hand_engine")
        flw._dev.set_option("hand_pathsum", 0)
        got2 = flw.hand(drain, elev)
        print("HAND MISMATCH", k, d8.shape, elev.dtype, "engine", eng, "n_bad", bad.size, "first", bad[:5], got.ravel()[bad[:5]], want.ravel()[bad[:5]],
              "sweep-only equal:", np.array_equal(got2, want), "rank<0 there:", (flw.rank.ravel()[bad[:5]]), flush=True)
        np.savez_compressed(f"gpurun_out/hand_mismatch_{seed}_{k}.npz", d8=d8, elev=elev, drain=drain)
        raise SystemExit(1)
    nh += 1
    nacc += flw._dev.info("hand_engine") == 1
print(f"hand: {nh} rasters equal to the oracle ({nacc} by path sums, {nh - nacc} by the hop-by-hop fallback), {time.time() - t0:.0f} s")
