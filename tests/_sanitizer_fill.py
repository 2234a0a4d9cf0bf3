"""compute-sanitizer case for dem.fill_depressions (csrc/pfd_fill.cuh): ties, drift with a band retry, holes, both replay
kernels (the 130 x 140 flat forms one component of more than 2048 cells -> warp replay), every outlet mode."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _cases as cs  # noqa: E402
import oracle  # noqa: E402
from pyflwdir_b200 import dem  # noqa: E402

for name, (a, kw) in cs.fill_cases().items():
    got, want = dem.fill_depressions(a.copy(), **kw), oracle.dem.fill_depressions(a.copy(), **kw)
    assert np.array_equal(got[0], want[0], equal_nan=True) and np.array_equal(got[1], want[1]), name
flat = np.full((130, 140), 5.0, dtype=np.float32)
flat[0, :] = flat[-1, :] = flat[:, 0] = flat[:, -1] = 9.0
flat[0, 70] = 1.0
got, want = dem.fill_depressions(flat), oracle.dem.fill_depressions(flat)
assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
z = oracle.synth_elevation(150, 170, seed=44) * np.float32(700.0)
z = (z - np.float32(np.median(z))).astype(np.float32)
got, want = dem.fill_depressions(z), oracle.dem.fill_depressions(z)
assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
print("sanitizer fill case ok")
