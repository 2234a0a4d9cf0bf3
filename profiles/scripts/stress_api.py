"""Randomised sweep of the WHOLE object API (tests/_cases.py:run_api_case, every output the parity tests know) against the oracle on
random rasters: random legal codes (loops, forced pits, nodata), synthetic terrain with a sea, shapes around the tile edges, and
every engine combination (tile solver on / off, tile-dataflow sweeps / level replays, HAND path sums on / off).
    python profiles/scripts/stress_api.py [SEED] [N]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import _cases as cs  # noqa: E402
import oracle  # noqa: E402
import pyflwdir_b200 as pfb  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rng = np.random.default_rng(seed)
legal = np.array([32, 64, 128, 16, 0, 1, 8, 4, 2, 247, 255], dtype=np.uint8)
engines = [dict(PFD_TILES="1", PFD_TILE_SWEEPS="1", PFD_HAND_PATHSUM="1"), dict(PFD_TILES="1", PFD_TILE_SWEEPS="2", PFD_HAND_PATHSUM="0"),
           dict(PFD_TILES="0", PFD_TILE_SWEEPS="0", PFD_HAND_PATHSUM="0")]
sizes = [2, 3, 5, 63, 64, 65, 127, 128, 129, 200, 257, 333, 512]
t0 = time.time()
ncmp = ndone = 0
for k in range(ncase):
    nr, nc = int(rng.choice(sizes)), int(rng.choice(sizes))
    if k % 2:
        p = np.array([1, 1, 1, 1, rng.uniform(0.01, 0.3), 1, 1, 1, 1, rng.uniform(0, 0.5), rng.uniform(0, 0.1)])
        d8 = legal[rng.choice(legal.size, size=(nr, nc), p=p / p.sum())]
    else:
        z = oracle.synth_elevation(nr, nc, seed=int(rng.integers(1 << 30)))
        d8 = oracle.synth_d8(z, sea_level=float(np.quantile(z, rng.uniform(0, 0.3))) if k % 4 == 0 else -np.inf)
    ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    if pits.size == 0:
        continue
    aux = cs.case_inputs("x", d8, int(rng.integers(1 << 20)))
    try:
        want = cs.run_oracle_case(d8, aux, area=np.ones(d8.size, dtype=np.float32))
    except (IndexError, ValueError):  # the case runner needs at least one labelled region / one ranked cell (tiny rasters)
        continue
    ndone += 1
    for env in engines:
        os.environ.update(env)
        got = cs.run_api_case(pfb, d8, aux)
        for key in want:
            if not np.array_equal(got[key], want[key], equal_nan=True):
                np.savez_compressed(f"gpurun_out/api_mismatch_{seed}_{k}.npz", d8=d8)
                print("MISMATCH", k, d8.shape, env, key, flush=True)
                raise SystemExit(1)
            ncmp += 1
print(f"{ndone} rasters x {len(engines)} engine settings: {ncmp} outputs equal to the oracle, {time.time() - t0:.0f} s")
