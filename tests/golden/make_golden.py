"""Generate the golden vectors under tests/golden/ by running the REAL reference (Deltares/pyflwdir v0.5.12,
numba) in the build container, where it is mounted read-only at /root/reference.

    python tests/golden/make_golden.py

Outputs (committed):
    tests/golden/small_cases.npz   inputs + full reference outputs for the small rasters
    tests/golden/rhine_d8.npz      the 682x997 D8 raster of examples/rhine_d8.tif (input data only)
    tests/golden/hashes.json       SHA-256 of every reference output (all cases), plus scalars

The reference cannot travel to the GPU box; these files can. Nothing here is product code.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference  # noqa: E402
import oracle  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from _cases import RHINE_TRANSFORM, case_inputs, run_api_case, sha  # noqa: E402


def run_case(pf, name, d8, seed, transform=None, latlon=False):
    """Reference outputs for one D8 raster: the very same API calls the GPU tests make (tests/_cases.py)."""
    aux = case_inputs(name, d8, seed)
    return run_api_case(pf, d8, aux, transform=transform, latlon=latlon), aux


def main():
    pf = reference.load()
    import pyflwdir.core as rcore  # noqa: F401

    small_inputs = {}
    small_outputs = {}
    hashes = {"reference_version": pf.__version__, "cases": {}}

    # --- the reference's own test rasters (tests/data/*.asc, tests/conftest.py:18-20,116-124)
    d8_small = np.loadtxt(os.path.join(reference.REFERENCE_ROOT, "tests/data/flwdir.asc"), dtype=np.uint8)
    d8_large = np.loadtxt(os.path.join(reference.REFERENCE_ROOT, "tests/data/flwdir1.asc"), dtype=np.uint8)
    # --- hand-made loop KAT
    d8_loop = np.array([[1, 16, 4], [1, 4, 4], [247, 0, 16]], dtype=np.uint8)
    # --- random legal codes: many loops, forced pits at the border and next to nodata
    rng = np.random.default_rng(4242)
    legal = np.array([32, 64, 128, 16, 0, 1, 8, 4, 2, 247, 255], dtype=np.uint8)
    p = np.array([1, 1, 1, 1, 0.1, 1, 1, 1, 1, 0.4, 0.1])
    d8_rand = legal[rng.choice(legal.size, size=(48, 61), p=p / p.sum())]
    # --- synthetic terrain with a "sea" (nodata) from the repo's own generator
    z = oracle.synth_elevation(96, 130, seed=11)
    d8_syn = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.08)))

    cases = {
        "flwdir_asc": (d8_small, 1),
        "flwdir1_asc": (d8_large, 2),
        "loop3x3": (d8_loop, 3),
        "random48x61": (d8_rand, 4),
        "synth96x130": (d8_syn, 5),
    }
    for name, (d8, seed) in cases.items():
        out, _ = run_case(pf, name, d8, seed)
        small_inputs[f"{name}/d8"] = d8
        for k, v in out.items():
            small_outputs[f"{name}/{k}"] = np.asarray(v)
        hashes["cases"][name] = {k: sha(np.asarray(v)) for k, v in out.items()}
        hashes["cases"][name]["_shape"] = list(d8.shape)
        hashes["cases"][name]["_seed"] = seed

    # --- index dtypes: the reference's fixtures parse with uint32 / uint64 (tests/conftest.py:23-26,88-108)
    import pyflwdir.core_d8 as rd8
    for dt in (np.uint32, np.int64):
        ids, pits, n = rd8.from_array(d8_small, dtype=dt)
        small_outputs[f"flwdir_asc/idxs_ds_{np.dtype(dt).name}"] = ids
        small_outputs[f"flwdir_asc/idxs_pit_{np.dtype(dt).name}"] = pits

    # --- PCRaster LDD (core_ldd.py): the 160x200 fixture converted with the reference's own d8_to_ldd
    ldd = pf.d8_to_ldd(d8_large).astype(np.uint8)
    flw_ldd = pf.from_array(ldd)  # ftype inferred
    assert flw_ldd.ftype == "ldd"
    small_inputs["ldd_flwdir1/ldd"] = ldd
    for k, v in dict(idxs_ds=flw_ldd.idxs_ds, idxs_pit=flw_ldd.idxs_pit, idxs_outlet=flw_ldd.idxs_outlet,
                     to_array=flw_ldd.to_array(), to_array_d8=flw_ldd.to_array("d8"), uparea_cell=flw_ldd.upstream_area(),
                     d8_to_ldd=ldd, ldd_to_d8=pf.ldd_to_d8(ldd).astype(np.uint8)).items():
        small_outputs[f"ldd_flwdir1/{k}"] = np.asarray(v)

    # --- CaMa-Flood NEXTXY (core_nextxy.py): the 160x200 fixture written with the reference's own to_array, two river
    #     mouths turned into inland pits (-10), one cell with a pit in the nexty plane only, a masked-out corner
    flw_d8 = pf.from_array(d8_large, ftype="d8")
    nxy = flw_d8.to_array("nextxy").astype(np.int32)
    pit_cells = flw_d8.idxs_pit
    for idx in pit_cells[:2]:
        nxy[:, idx // d8_large.shape[1], idx % d8_large.shape[1]] = -10
    free = np.flatnonzero((nxy[0].ravel() > 0))[37]
    nxy[1, free // d8_large.shape[1], free % d8_large.shape[1]] = -9
    nxy[:, :9, :13] = -9999
    flw_xy = pf.from_array(nxy)  # ftype inferred
    assert flw_xy.ftype == "nextxy"
    small_inputs["nextxy_flwdir1/nextxy"] = nxy
    mask_xy = np.ones(nxy.shape, dtype=np.uint8)
    mask_xy[:, 100:, 150:] = 0
    flw_xy_m = pf.from_array(nxy, ftype="nextxy", mask=mask_xy)
    for k, v in dict(idxs_ds=flw_xy.idxs_ds, idxs_pit=flw_xy.idxs_pit, idxs_outlet=flw_xy.idxs_outlet, idxs_seq=flw_xy.idxs_seq,
                     to_array=flw_xy.to_array(), to_array_d8=flw_xy.to_array("d8"), uparea_cell=flw_xy.upstream_area(),
                     basins=flw_xy.basins(), masked_idxs_ds=flw_xy_m.idxs_ds, masked_idxs_pit=flw_xy_m.idxs_pit,
                     d8_to_nextxy=flw_d8.to_array("nextxy")).items():
        small_outputs[f"nextxy_flwdir1/{k}"] = np.asarray(v)

    # --- drdc over all 256 codes (core_d8.py:22-39), including the illegal ones
    drdc = np.array([rd8.drdc(np.uint8(i)) for i in range(256)], dtype=np.int8)
    small_outputs["drdc_table"] = drdc

    # --- rhine (config 1 of BASELINE.json): hashes only + the input raster
    from PIL import Image

    rhine = np.array(Image.open(os.path.join(reference.REFERENCE_ROOT, "examples/rhine_d8.tif")))
    out, aux = run_case(pf, "rhine", rhine, 6, transform=RHINE_TRANSFORM, latlon=True)
    hashes["cases"]["rhine"] = {k: sha(np.asarray(v)) for k, v in out.items()}
    hashes["cases"]["rhine"]["_shape"] = list(rhine.shape)
    hashes["cases"]["rhine"]["_seed"] = 6
    hashes["cases"]["rhine"]["_transform"] = list(RHINE_TRANSFORM)
    hashes["cases"]["rhine"]["_max_rank"] = int(out["rank"].max())
    hashes["cases"]["rhine"]["_max_uparea_km2"] = float(out["uparea_km2"].max())
    hashes["cases"]["rhine"]["_strord_hist"] = np.bincount(out["strord"].ravel(), minlength=10).tolist()
    hashes["cases"]["rhine"]["_aux"] = {k: sha(v) for k, v in aux.items()}
    # area grid rows (input of upstream_area("km2")), to pin the host-side cellarea restatement
    small_outputs["rhine/area_col0"] = np.asarray(
        pf.from_array(rhine, ftype="d8", transform=RHINE_TRANSFORM, latlon=True).area[:, 0]
    )

    # --- streams.upstream_area (no area grid): rhine in m2 / km2 (geographic), the 160x200 fixture in cells (int32)
    import pyflwdir.streams as rstreams
    flw_r = pf.from_array(rhine, ftype="d8", transform=RHINE_TRANSFORM, latlon=True)
    ua = rstreams.upstream_area(flw_r.idxs_ds, flw_r.idxs_seq, rhine.shape[1], latlon=True, transform=flw_r.transform, area_factor=1e6)
    hashes["cases"]["rhine"]["streams_uparea_km2"] = sha(ua)
    flw_l = pf.from_array(d8_large, ftype="d8")
    small_outputs["flwdir1_asc/streams_uparea_i32"] = rstreams.upstream_area(flw_l.idxs_ds, flw_l.idxs_seq, d8_large.shape[1], dtype=np.int32)
    small_outputs["flwdir1_asc/streams_uparea_f32"] = rstreams.upstream_area(
        flw_l.idxs_ds, flw_l.idxs_seq, d8_large.shape[1], transform=(30.0, 0.0, 0.0, 0.0, -20.0, 0.0), area_factor=1e4, dtype=np.float32)

    # --- a mid-size synthetic (512 x 768, with sea): hashes only, input regenerated in the tests
    z = oracle.synth_elevation(512, 768, seed=21)
    sea = float(np.quantile(z, 0.05))
    d8_mid = oracle.synth_d8(z, sea_level=sea)
    out, aux = run_case(pf, "synth512x768", d8_mid, 8)
    hashes["cases"]["synth512x768"] = {k: sha(np.asarray(v)) for k, v in out.items()}
    hashes["cases"]["synth512x768"].update(
        {"_shape": [512, 768], "_seed": 8, "_synth_seed": 21, "_sea_level": sea, "_d8": sha(d8_mid),
         "_max_rank": int(out["rank"].max()), "_npits": int(out["idxs_pit"].size),
         "_aux": {k: sha(v) for k, v in aux.items()}}
    )

    np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **{f"in/{k}": v for k, v in small_inputs.items()},
                        **{f"out/{k}": v for k, v in small_outputs.items()})
    np.savez_compressed(os.path.join(HERE, "rhine_d8.npz"), d8=rhine)
    with open(os.path.join(HERE, "hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
