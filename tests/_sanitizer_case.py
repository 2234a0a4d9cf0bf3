import sys, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import _cases as cs, oracle
import pyflwdir_b200 as pfb
from pyflwdir_b200 import tiled
for name in ["flwdir1_asc", "random48x61", "synth96x130"]:
    d8 = cs.case_d8(name)
    aux = cs.case_inputs(name, d8, cs.case_seed(name))
    out = cs.run_api_case(pfb, d8, aux)
    for k, v in out.items():
        cs.check(name, k, v)
nxy = cs.small()["in/nextxy_flwdir1/nextxy"]
flw = pfb.from_array(nxy)
assert np.array_equal(flw.to_array(), cs.small()["out/nextxy_flwdir1/to_array"])
z = oracle.synth_elevation(200, 150, seed=41)
d8 = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.05)))
got = tiled.solve_emulated(d8, 3)
# tile-dataflow sweeps, forced (single GPU) and across three row blocks with the halo rounds emulated
import os
os.environ["PFD_TILE_SWEEPS"] = "2"
flw = pfb.from_array(d8, ftype="d8")
ids, pits, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
seq = oracle.core.idxs_seq(ids, pits)
assert np.array_equal(flw.stream_order().ravel(), oracle.streams.strahler_order(ids, seq))
area = (np.abs(z) + np.float32(0.5)).astype(np.float32)
assert np.array_equal(flw.accuflux(area).ravel(), oracle.streams.accuflux(ids, seq, area.ravel(), -9999))
drain = flw.upstream_area() > 40
assert np.array_equal(flw.hand(drain, z).ravel(), oracle.dem.height_above_nearest_drain(ids, seq, drain.ravel(), z.ravel()))
for kind, kw in (("strahler", {}), ("accuflux", dict(data=area, nodata=-9999.0)), ("hand", dict(data=z, drain=drain))):
    tiled.sweep_emulated(d8, 3, kind, **kw)
print("sanitizer case ok", got["n_pits"])
