import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
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
print("sanitizer case ok", got["n_pits"])
