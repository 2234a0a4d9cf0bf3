"""Shared helpers for the parity tests: golden fixtures (generated from the real reference by
tests/golden/make_golden.py) and the deterministic auxiliary inputs of every case."""
import hashlib
import json
import os

import numpy as np

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

RHINE_TRANSFORM = (0.008333333333325754, 0.0, 3.5666666664997138, 0.0, -0.008333333333339965, 52.00833333330708)

SMALL_CASES = ["flwdir_asc", "flwdir1_asc", "loop3x3", "random48x61", "synth96x130"]
HASH_CASES = ["rhine", "synth512x768"]
FEATS_MAX_CELLS = 40000  # geo-feature cases (Python dict per segment) run on rasters up to this size


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def case_inputs(name, d8, seed):
    """Deterministic auxiliary inputs for a case (identical in make_golden.py and in the tests)."""
    rng = np.random.default_rng(seed)
    shape = d8.shape
    data_f32 = rng.random(shape, dtype=np.float32) * np.float32(10.0)
    data_f64 = rng.random(shape) * 1e3
    holes = rng.random(shape) < 0.01
    data_f32_nd = np.where(holes, np.float32(-9999.0), data_f32)
    data_i64 = rng.integers(-50, 1000, size=shape, dtype=np.int64)
    elevtn = (oracle.synth_elevation(shape[0], shape[1], seed=seed + 7) * np.float32(1000.0)).astype(np.float32)
    smask = rng.random(shape) < 0.6
    gaps = rng.random(shape) < 0.7  # sparse data for fillnodata
    fill_i64 = np.where(gaps, np.int64(-9999), data_i64)
    fill_f32 = np.where(gaps, np.float32(-1.5), data_f32)
    fill_f64 = np.where(gaps, np.nan, data_f64)
    return dict(data_f32=data_f32, data_f64=data_f64, data_f32_nd=data_f32_nd, data_i64=data_i64,
                elevtn=elevtn, smask=smask, fill_i64=fill_i64, fill_f32=fill_f32, fill_f64=fill_f64)


_small = None
_hashes = None


def small():
    global _small
    if _small is None:
        _small = dict(np.load(os.path.join(GOLDEN, "small_cases.npz")))
    return _small


def hashes():
    global _hashes
    if _hashes is None:
        with open(os.path.join(GOLDEN, "hashes.json")) as f:
            _hashes = json.load(f)
    return _hashes["cases"]


def case_d8(name):
    if name in SMALL_CASES:
        return small()[f"in/{name}/d8"]
    if name == "rhine":
        return np.load(os.path.join(GOLDEN, "rhine_d8.npz"))["d8"]
    if name == "synth512x768":
        h = hashes()[name]
        z = oracle.synth_elevation(512, 768, seed=h["_synth_seed"])
        d8 = oracle.synth_d8(z, sea_level=h["_sea_level"])
        assert sha(d8) == h["_d8"], "synthetic generator drifted from the golden input"
        return d8
    raise KeyError(name)


def case_seed(name):
    return hashes()[name]["_seed"]


def golden(name, key):
    """Full golden array for small cases, else None (hash-only)."""
    return small().get(f"out/{name}/{key}")


def check(name, key, arr):
    """Assert `arr` equals the reference output `key` of case `name` bit for bit."""
    arr = np.asarray(arr)
    g = golden(name, key)
    if g is not None:
        assert arr.dtype == g.dtype, f"{name}/{key}: dtype {arr.dtype} != {g.dtype}"
        assert arr.shape == g.shape, f"{name}/{key}: shape {arr.shape} != {g.shape}"
        if not np.array_equal(arr, g, equal_nan=True):
            bad = np.flatnonzero(arr.ravel() != g.ravel())
            raise AssertionError(f"{name}/{key}: {bad.size} mismatches, first at {bad[:5]}: "
                                 f"{arr.ravel()[bad[:5]]} != {g.ravel()[bad[:5]]}")
    assert sha(arr) == hashes()[name][key], f"{name}/{key}: SHA-256 differs from the reference's"


def local_inputs(d8, seq, aux):
    """Deterministic start cells / region masks / label rasters for the local-trace and region cases."""
    nrow, ncol = d8.shape
    starts = np.asarray(seq[:: max(1, seq.size // 37)][:60]).astype(np.int64)
    nod = np.flatnonzero(d8.ravel() == 247)[:2].astype(np.int64)
    starts = np.concatenate([starts, nod])
    region = np.zeros(d8.shape, dtype=np.bool_)
    region[nrow // 4: max(nrow // 4 + 1, 3 * nrow // 4), ncol // 5: max(ncol // 5 + 1, 4 * ncol // 5)] = True
    rr, cc = np.arange(nrow)[:, None], np.arange(ncol)[None, :]
    blocks = (((rr // 7) % 3) * 3 + ((cc // 9) % 3)).astype(np.int32)  # labels 0..8 repeated over the raster: ties
    labels = ((rr // max(1, nrow // 5)) * 7 + (cc // max(1, ncol // 4)) * 3 + 2).astype(np.int64)  # gaps in the numbering
    labels[aux["smask"] & (rr % 3 == 0)] = 0
    return starts, region, blocks, labels


def row_starts(d8, seq):
    """One start cell per raster row (the first cell of the row that drains to a pit): metric traces then touch the
    hop-length table of every row."""
    seq = np.sort(np.asarray(seq).astype(np.int64))
    rows = seq // d8.shape[1]
    first = np.flatnonzero(np.diff(rows, prepend=-1) != 0)
    return seq[first]


def _flat_paths(paths):
    counts = np.array([p.size for p in paths], dtype=np.int64)
    flat = np.concatenate(paths) if len(paths) else np.zeros(0, dtype=np.int64)
    return flat, counts


def _feats_arrays(feats, prefix, out, props=()):
    """Geo-features (list of dicts) -> comparable arrays."""
    out[f"{prefix}_idx"] = np.array([f["properties"]["idx"] for f in feats], dtype=np.int64)
    out[f"{prefix}_idx_ds"] = np.array([f["properties"]["idx_ds"] for f in feats], dtype=np.int64)
    out[f"{prefix}_pit"] = np.array([bool(f["properties"]["pit"]) for f in feats], dtype=np.bool_)
    out[f"{prefix}_n"] = np.array([len(f["geometry"]["coordinates"]) for f in feats], dtype=np.int64)
    xy = [c for f in feats for c in f["geometry"]["coordinates"]]
    out[f"{prefix}_xy"] = np.array(xy, dtype=np.float64).reshape(-1, 2)
    for key in props:
        out[f"{prefix}_{key}"] = np.array([f["properties"][key] for f in feats])


def _oracle_feats(paths, prefix, out, transform, ncol, props=None):
    """gis_utils.features (gis_utils.py:490-549) over index lists, with Affine * translation(0.5, 0.5) * (cols, rows)
    spelled out (gis_utils.py:222)."""
    a, b, c, d, e, f = [float(v) for v in tuple(transform)[:6]]
    cx, cy = a * 0.5 + b * 0.5 + c, d * 0.5 + e * 0.5 + f
    paths = [p for p in paths if len(p) >= 2]
    out[f"{prefix}_idx"] = np.array([p[0] for p in paths], dtype=np.int64)
    out[f"{prefix}_idx_ds"] = np.array([p[-1] for p in paths], dtype=np.int64)
    out[f"{prefix}_pit"] = np.array([p[-1] == p[-2] for p in paths], dtype=np.bool_)
    out[f"{prefix}_n"] = np.array([len(p) for p in paths], dtype=np.int64)
    flat = np.concatenate(paths).astype(np.int64) if paths else np.zeros(0, np.int64)
    rows, cols = flat // ncol, flat % ncol
    out[f"{prefix}_xy"] = np.stack([cols * a + rows * b + cx, cols * d + rows * e + cy], axis=1).reshape(-1, 2)
    for key, arr in (props or {}).items():
        out[f"{prefix}_{key}"] = np.asarray(arr).ravel()[out[f"{prefix}_idx"]]


def run_oracle_case(d8, aux, area=None, transform=(1.0, 0.0, 0.0, 0.0, -1.0, 0.0), latlon=False):
    """The whole hot path on the CPU oracle, mirroring make_golden.run_case (which runs the real reference).
    `area` = flat cell-area vector [m2] for upstream_area("km2") (host-side input, see gis_utils.area_grid)."""
    o = oracle
    shape = d8.shape
    dtype = o.get_idxs_dtype(d8.size)
    idxs_ds, idxs_pit, n = o.core_d8.from_array(d8, dtype=dtype)
    out = {}
    out["idxs_ds"] = idxs_ds
    out["idxs_pit"] = idxs_pit
    out["idxs_outlet"] = idxs_pit[np.isin(d8.flat[idxs_pit], [0, 255])]
    rank, nn = o.core.rank(idxs_ds)
    out["rank"] = rank.reshape(shape)
    seq = o.core.idxs_seq(idxs_ds, idxs_pit)
    out["idxs_seq"] = seq
    out["nnodes"] = np.int64(seq.size)
    out["isvalid"] = np.bool_(np.all(rank != -1))
    out["n_upstream"] = o.core.upstream_count(idxs_ds).reshape(shape)
    mask = idxs_ds != idxs_ds.dtype.type(-1)

    def uparea(a):
        u = o.streams.accuflux(idxs_ds, seq, a, -9999)
        u[~mask] = -9999
        return u.reshape(shape)

    out["uparea_cell"] = uparea(np.ones(d8.size, dtype=np.int32))
    if area is not None:
        out["uparea_km2"] = uparea(area / 1e6)
    out["basins"] = o.basins.basins(idxs_ds, idxs_pit, seq).reshape(shape)
    out["strord"] = o.streams.strahler_order(idxs_ds, seq).reshape(shape)
    out["strord_mask"] = o.streams.strahler_order(idxs_ds, seq, mask=aux["smask"].ravel()).reshape(shape)
    out["accu_f32"] = o.streams.accuflux(idxs_ds, seq, aux["data_f32"].ravel(), -9999).reshape(shape)
    out["accu_f64"] = o.streams.accuflux(idxs_ds, seq, aux["data_f64"].ravel(), -9999.0).reshape(shape)
    out["accu_f32_nd"] = o.streams.accuflux(idxs_ds, seq, aux["data_f32_nd"].ravel(), -9999).reshape(shape)
    out["accu_i64"] = o.streams.accuflux(idxs_ds, seq, aux["data_i64"].ravel(), -9999).reshape(shape)
    out["accu_ds_f64"] = o.streams.accuflux_ds(idxs_ds, seq, aux["data_f64"].ravel(), -9999.0).reshape(shape)
    out["accu_ds_i64"] = o.streams.accuflux_ds(idxs_ds, seq, aux["data_i64"].ravel(), -9999).reshape(shape)
    drain = out["uparea_cell"] > max(4, int(0.002 * d8.size))
    out["hand_f32"] = o.dem.height_above_nearest_drain(idxs_ds, seq, drain.ravel(), aux["elevtn"].ravel()).reshape(shape)
    out["hand_f64"] = o.dem.height_above_nearest_drain(
        idxs_ds, seq, drain.ravel(), (aux["elevtn"].astype(np.float64) * 1.1).ravel()).reshape(shape)
    sub_idxs = seq[:: max(1, seq.size // 23)][:40]
    sub_ids = (np.arange(sub_idxs.size, dtype=np.int64) * 3 + 5).astype(np.int32)
    out["sub_idxs"] = sub_idxs
    out["sub_ids"] = sub_ids
    out["basins_sub"] = o.basins.basins(idxs_ds, sub_idxs, seq, sub_ids).reshape(shape)
    out["to_array"] = o.core_d8.to_array(idxs_ds, shape)
    out["us_main"] = o.core.main_upstream(idxs_ds, out["uparea_cell"].ravel())
    if area is not None:
        out["us_main_km2"] = o.core.main_upstream(idxs_ds, out["uparea_km2"].ravel())
    out["strord_classic"] = o.streams.stream_order(idxs_ds, seq, out["us_main"]).reshape(shape)
    out["strord_classic_mask"] = o.streams.stream_order(idxs_ds, seq, out["us_main"], mask=aux["smask"].ravel()).reshape(shape)
    out["fill_up_i64"] = o.core.fillnodata_upstream_any(idxs_ds, seq, aux["fill_i64"].ravel(), -9999).reshape(shape)
    out["fill_up_f64nan"] = o.core.fillnodata_upstream_any(idxs_ds, seq, aux["fill_f64"].ravel(), np.nan).reshape(shape)
    out["fill_down_max_f32"] = o.core.fillnodata_downstream(idxs_ds, seq, aux["fill_f32"].ravel(), -1.5, how="max").reshape(shape)
    out["fill_down_min_i64"] = o.core.fillnodata_downstream(idxs_ds, seq, aux["fill_i64"].ravel(), -9999, how="min").reshape(shape)
    out["fill_down_sum_f32"] = o.core.fillnodata_downstream(idxs_ds, seq, aux["fill_f32"].ravel(), -1.5, how="sum").reshape(shape)
    out["to_array_ldd"] = o.core_ldd.to_array(idxs_ds, shape)
    sm = aux["smask"].ravel()
    sd = lambda mask, real: o.streams.stream_distance(idxs_ds, seq, shape[1], mask=mask, real_length=real, latlon=latlon,
                                                      transform=tuple(transform)[:6]).reshape(shape)
    out["sdist_cell"] = sd(None, False)
    out["sdist_cell_mask"] = sd(sm, False)
    out["sdist_m"] = sd(None, True)
    out["sdist_m_mask"] = sd(sm, True)
    if area is not None:
        ukm = out["uparea_km2"]
        upa_min = float(np.quantile(ukm[ukm > 0], 0.97)) if (ukm > 0).any() else 1.0
        out["fldpln_f32"] = o.dem.floodplains(idxs_ds, seq, aux["elevtn"].ravel(), ukm.ravel(), upa_min=upa_min, b=0.3).reshape(shape)
    out["fldpln_f64"] = o.dem.floodplains(idxs_ds, seq, (aux["elevtn"].astype(np.float64) * 1.1).ravel(),
                                          out["uparea_cell"].ravel(), upa_min=max(4, int(0.002 * d8.size)), b=0.5).reshape(shape)
    out["upsum_f32"] = o.arithmetics.upstream_sum(idxs_ds, aux["data_f32_nd"].ravel(), -9999).reshape(shape)
    out["upsum_i64"] = o.arithmetics.upstream_sum(idxs_ds, aux["fill_i64"].ravel(), -9999).reshape(shape)
    out["upsum_f64"] = o.arithmetics.upstream_sum(idxs_ds, aux["data_f64"].ravel(), -9999.0).reshape(shape)
    sub, sidx = o.basins.subbasins_streamorder(idxs_ds, seq, out["strord"].ravel(), None, -2)
    out["subbas_so"], out["subbas_so_idxs"] = sub.reshape(shape), sidx
    sub, sidx = o.basins.subbasins_streamorder(idxs_ds, seq, out["strord"].ravel(), sm, 2)
    out["subbas_so_mask"], out["subbas_so_mask_idxs"] = sub.reshape(shape), sidx
    if area is not None:
        ukm = out["uparea_km2"]
        amin = float(np.quantile(ukm[ukm > 0], 0.9)) if (ukm > 0).any() else 1.0
        sub, sidx = o.basins.subbasins_area(idxs_ds, seq, out["us_main"], ukm.ravel(), amin)
        out["subbas_area_km2"], out["subbas_area_km2_idxs"] = sub.reshape(shape), sidx
    sub, sidx = o.basins.subbasins_area(idxs_ds, seq, out["us_main"], out["uparea_cell"].ravel(), max(3, d8.size // 400))
    out["subbas_area_cell"], out["subbas_area_cell_idxs"] = sub.reshape(shape), sidx
    um, so = out["us_main"], out["strord"].ravel()
    ar = o.arithmetics
    out["movavg_f32"] = ar.moving_average(aux["data_f32_nd"].ravel(), None, 3, idxs_ds, um, None, -9999.0).reshape(shape)
    out["movavg_f64_w"] = ar.moving_average(aux["data_f64"].ravel(), aux["data_f64"].ravel(), 5, idxs_ds, um, so, -9999.0).reshape(shape)
    out["movavg_f32_w"] = ar.moving_average(aux["fill_f32"].ravel(), aux["data_f64"].ravel(), 2, idxs_ds, um, None, -1.5).reshape(shape)
    out["movavg_f64_nan"] = ar.moving_average(aux["fill_f64"].ravel(), None, 4, idxs_ds, um, None, np.nan).reshape(shape)
    out["movmed_f32"] = ar.moving_median(aux["data_f32_nd"].ravel(), 3, idxs_ds, um, None, -9999.0).reshape(shape)
    out["movmed_f64_so"] = ar.moving_median(aux["data_f64"].ravel(), 4, idxs_ds, um, so, -9999.0).reshape(shape)
    out["movmed_f32_sparse"] = ar.moving_median(aux["fill_f32"].ravel(), 6, idxs_ds, um, None, -1.5).reshape(shape)
    # local traces and region post-processing
    starts, region, blocks, labels = local_inputs(d8, seq, aux)
    tr6 = tuple(transform)[:6]
    upc = out["uparea_cell"].ravel()
    stream = upc > max(4, int(0.002 * d8.size))
    for tag, nxt in (("down", idxs_ds), ("up", um)):
        for key, kw in (("", {}), ("_mask", dict(mask=stream)), ("_max", dict(max_length=7.5)),
                        ("_m", dict(mask=stream, max_length=3000.0 * abs(tr6[0]) * (111e3 if latlon else 1.0) / 1e3,
                                    real_length=True, ncol=shape[1], latlon=latlon, transform=tr6))):
            paths, dist = o.core.path(starts, nxt, **kw)
            out[f"path_{tag}{key}"], out[f"path_{tag}{key}_n"] = _flat_paths(paths)
            out[f"path_{tag}{key}_dist"] = dist
            out[f"snap_{tag}{key}"], out[f"snap_{tag}{key}_dist"] = o.core.snap(starts, nxt, **kw)
    rs = row_starts(d8, seq)
    for tag, nxt in (("down", idxs_ds), ("up", um)):
        kw = dict(max_length=40.0 * abs(tr6[0]) * (111e3 if latlon else 1.0), real_length=True, ncol=shape[1], latlon=latlon, transform=tr6)
        paths, dist = o.core.path(rs, nxt, **kw)
        out[f"path_{tag}_rows"], out[f"path_{tag}_rows_n"] = _flat_paths(paths)
        out[f"path_{tag}_rows_dist"] = dist
        out[f"snap_{tag}_rows"], out[f"snap_{tag}_rows_dist"] = o.core.snap(rs, nxt, **kw)
    out["downstream_f32"] = np.where(mask, aux["data_f32"].ravel()[np.where(mask, idxs_ds, 0).astype(np.int64)],
                                     aux["data_f32"].ravel()).reshape(shape)
    out["downstream_i64"] = np.where(mask, aux["data_i64"].ravel()[np.where(mask, idxs_ds, 0).astype(np.int64)],
                                     aux["data_i64"].ravel()).reshape(shape)
    out["inflow_idxs"] = o.core.inflow_idxs(idxs_ds, seq, region.ravel())
    out["outflow_idxs"] = o.core.outflow_idxs(idxs_ds, seq, region.ravel())
    out["interbasin"] = o.basins.interbasin_mask(idxs_ds, seq, region.ravel()).reshape(shape)
    out["interbasin_stream"] = o.basins.interbasin_mask(idxs_ds, seq, region.ravel(), stream).reshape(shape)
    out["outlets_basins_lbs"], out["outlets_basins_idxs"] = o.regions.region_outlets(out["basins"], idxs_ds, seq)
    out["outlets_blocks_lbs"], out["outlets_blocks_idxs"] = o.regions.region_outlets(blocks, idxs_ds, seq)
    out["bounds_basins_lbs"], out["bounds_basins_boxes"], out["bounds_basins_total"] = o.regions.region_bounds(out["basins"], tr6)
    out["bounds_labels_lbs"], out["bounds_labels_boxes"], out["bounds_labels_total"] = o.regions.region_bounds(labels, tr6)
    # stream segments and flow-direction vectors as geo-features
    so2 = out["strord"].ravel() >= 2
    _oracle_feats(o.streams.streams(idxs_ds, seq, so2, 0), "streams_so2", out, tr6, shape[1],
                  props=dict(strord=out["strord"], uparea=out["uparea_cell"]))
    if d8.size <= FEATS_MAX_CELLS:
        _oracle_feats(o.streams.streams(idxs_ds, seq, None, 7), "streams_len7", out, tr6, shape[1])
        _oracle_feats(o.streams.streams(idxs_ds, seq, aux["smask"].ravel(), 3), "streams_gaps", out, tr6, shape[1])
        vmask = aux["smask"].ravel()
        pairs = [np.array([i, idxs_ds[i]]) for i in np.flatnonzero(mask & vmask)]
        _oracle_feats(pairs, "vector", out, tr6, shape[1])
    # Pfafstetter subbasins
    upc_flat = out["uparea_cell"].ravel()
    for key, depth, upa_min in (("pfaf_d1", 1, 0.0), ("pfaf_d2", 2, 0.0), ("pfaf_d3_min", 3, 5.0)):
        sub, sidx = o.basins.subbasins_pfafstetter(idxs_pit, idxs_ds, seq, um, upc_flat, mask=upc_flat >= upa_min, depth=depth)
        out[key], out[key + "_idxs"] = sub.reshape(shape), sidx
    # rivers: estuaries and manning river depth
    distnc = out["sdist_m"].ravel()
    rivwth32 = np.sqrt(np.abs(upc_flat).astype(np.float32)) + aux["data_f32"].ravel()
    elev0 = aux["elevtn"].ravel() - np.float32(np.median(aux["elevtn"].ravel()[idxs_pit]))
    out["estuary_f32"] = o.rivers.classify_estuary(idxs_ds, seq, idxs_pit, distnc, rivwth32, elev0, 0, 1e-2)
    out["estuary_f64"] = o.rivers.classify_estuary(idxs_ds, seq, idxs_pit, distnc.astype(np.float64) * 0.5,
                                                   rivwth32.astype(np.float64), elev0, 1.5, 1e-3)
    return out


def run_api_case(pf, d8, aux, transform=None, latlon=False):
    """The hot path through the object API of `pf` (the reference `pyflwdir` in make_golden.py, `pyflwdir_b200`
    in the GPU tests) -- same calls, same order."""
    kw = {}
    if transform is not None:
        kw = dict(transform=transform, latlon=latlon)
    flw = pf.from_array(d8, ftype="d8", cache=False, **kw)
    out = {}
    out["idxs_ds"] = flw.idxs_ds
    out["idxs_pit"] = flw.idxs_pit
    out["idxs_outlet"] = flw.idxs_outlet
    out["rank"] = flw.rank
    out["idxs_seq"] = flw.idxs_seq
    out["nnodes"] = np.int64(flw.nnodes)
    out["isvalid"] = np.bool_(flw.isvalid)
    out["n_upstream"] = flw.n_upstream
    out["uparea_cell"] = flw.upstream_area()
    out["uparea_km2"] = flw.upstream_area("km2")
    out["basins"] = flw.basins()
    out["strord"] = flw.stream_order()
    out["strord_mask"] = flw.stream_order(mask=aux["smask"])
    out["accu_f32"] = flw.accuflux(aux["data_f32"], nodata=-9999)
    out["accu_f64"] = flw.accuflux(aux["data_f64"], nodata=-9999.0)
    out["accu_f32_nd"] = flw.accuflux(aux["data_f32_nd"], nodata=-9999)
    out["accu_i64"] = flw.accuflux(aux["data_i64"], nodata=-9999)
    out["accu_ds_f64"] = flw.accuflux(aux["data_f64"], nodata=-9999.0, direction="down")
    out["accu_ds_i64"] = flw.accuflux(aux["data_i64"], nodata=-9999, direction="down")
    drain = out["uparea_cell"] > max(4, int(0.002 * d8.size))
    out["hand_f32"] = flw.hand(drain, aux["elevtn"])
    out["hand_f64"] = flw.hand(drain, aux["elevtn"].astype(np.float64) * 1.1)
    seq = flw.idxs_seq
    sub_idxs = seq[:: max(1, seq.size // 23)][:40]
    sub_ids = (np.arange(sub_idxs.size, dtype=np.int64) * 3 + 5).astype(np.int32)
    out["sub_idxs"] = sub_idxs
    out["sub_ids"] = sub_ids
    out["basins_sub"] = flw.basins(idxs=sub_idxs, ids=sub_ids)
    out["to_array"] = flw.to_array()
    # SURVEY.md §8f "next" rows
    out["us_main"] = flw.main_upstream()
    out["us_main_km2"] = flw.main_upstream(uparea=out["uparea_km2"])
    out["strord_classic"] = flw.stream_order(type="classic")
    out["strord_classic_mask"] = flw.stream_order(type="classic", mask=aux["smask"])
    out["fill_up_i64"] = flw.fillnodata(aux["fill_i64"], -9999, direction="up")
    out["fill_up_f64nan"] = flw.fillnodata(aux["fill_f64"], np.nan, direction="up")
    out["fill_down_max_f32"] = flw.fillnodata(aux["fill_f32"], -1.5, direction="down", how="max")
    out["fill_down_min_i64"] = flw.fillnodata(aux["fill_i64"], -9999, direction="down", how="min")
    out["fill_down_sum_f32"] = flw.fillnodata(aux["fill_f32"], -1.5, direction="down", how="sum")
    out["to_array_ldd"] = flw.to_array("ldd")
    out["sdist_cell"] = flw.stream_distance()
    out["sdist_cell_mask"] = flw.stream_distance(mask=aux["smask"])
    out["sdist_m"] = flw.stream_distance(unit="m")
    out["sdist_m_mask"] = flw.stream_distance(mask=aux["smask"], unit="m")
    upa_min = float(np.quantile(out["uparea_km2"][out["uparea_km2"] > 0], 0.97)) if (out["uparea_km2"] > 0).any() else 1.0
    out["fldpln_f32"] = flw.floodplains(aux["elevtn"], upa_min=upa_min, b=0.3)
    out["fldpln_f64"] = flw.floodplains(aux["elevtn"].astype(np.float64) * 1.1, uparea=out["uparea_cell"],
                                        upa_min=max(4, int(0.002 * d8.size)), b=0.5)
    out["upsum_f32"] = flw.upstream_sum(aux["data_f32_nd"], mv=-9999)
    out["upsum_i64"] = flw.upstream_sum(aux["fill_i64"], mv=-9999)
    out["upsum_f64"] = flw.upstream_sum(aux["data_f64"], mv=-9999.0)
    out["subbas_so"], out["subbas_so_idxs"] = flw.subbasins_streamorder(min_sto=-2)
    out["subbas_so_mask"], out["subbas_so_mask_idxs"] = flw.subbasins_streamorder(min_sto=2, mask=aux["smask"])
    ukm = out["uparea_km2"]
    amin = float(np.quantile(ukm[ukm > 0], 0.9)) if (ukm > 0).any() else 1.0
    out["subbas_area_km2"], out["subbas_area_km2_idxs"] = flw.subbasins_area(amin)
    out["subbas_area_cell"], out["subbas_area_cell_idxs"] = flw.subbasins_area(max(3, d8.size // 400), uparea=out["uparea_cell"])
    out["movavg_f32"] = flw.moving_average(aux["data_f32_nd"], 3, nodata=-9999.0)
    out["movavg_f64_w"] = flw.moving_average(aux["data_f64"], 5, weights=aux["data_f64"], restrict_strord=True, strord=out["strord"], nodata=-9999.0)
    out["movavg_f32_w"] = flw.moving_average(aux["fill_f32"], 2, weights=aux["data_f64"], nodata=-1.5)
    out["movavg_f64_nan"] = flw.moving_average(aux["fill_f64"], 4, nodata=np.nan)
    out["movmed_f32"] = flw.moving_median(aux["data_f32_nd"], 3, nodata=-9999.0)
    out["movmed_f64_so"] = flw.moving_median(aux["data_f64"], 4, restrict_strord=True, strord=out["strord"], nodata=-9999.0)
    out["movmed_f32_sparse"] = flw.moving_median(aux["fill_f32"], 6, nodata=-1.5)
    # local traces and region post-processing
    starts, region, blocks, labels = local_inputs(d8, seq, aux)
    stream = out["uparea_cell"] > max(4, int(0.002 * d8.size))
    xres = abs(flw.transform[0])
    for tag in ("down", "up"):
        for key, kw in (("", {}), ("_mask", dict(mask=stream)), ("_max", dict(max_length=7.5)),
                        ("_m", dict(mask=stream, max_length=3000.0 * xres * (111e3 if latlon else 1.0) / 1e3, unit="m"))):
            paths, dist = flw.path(idxs=starts, direction=tag, **kw)
            out[f"path_{tag}{key}"], out[f"path_{tag}{key}_n"] = _flat_paths(list(paths))
            out[f"path_{tag}{key}_dist"] = dist
            out[f"snap_{tag}{key}"], out[f"snap_{tag}{key}_dist"] = flw.snap(idxs=starts, direction=tag, **kw)
    rs = row_starts(d8, seq)
    for tag in ("down", "up"):
        kw = dict(max_length=40.0 * xres * (111e3 if latlon else 1.0), unit="m", direction=tag)
        paths, dist = flw.path(idxs=rs, **kw)
        out[f"path_{tag}_rows"], out[f"path_{tag}_rows_n"] = _flat_paths(list(paths))
        out[f"path_{tag}_rows_dist"] = dist
        out[f"snap_{tag}_rows"], out[f"snap_{tag}_rows_dist"] = flw.snap(idxs=rs, **kw)
    out["downstream_f32"] = flw.downstream(aux["data_f32"])
    out["downstream_i64"] = flw.downstream(aux["data_i64"])
    out["inflow_idxs"] = flw.inflow_idxs(region)
    out["outflow_idxs"] = flw.outflow_idxs(region)
    out["interbasin"] = flw.interbasin_mask(region)
    out["interbasin_stream"] = flw.interbasin_mask(region, stream=stream)
    out["outlets_basins_lbs"], out["outlets_basins_idxs"] = flw.basin_outlets(out["basins"])
    out["outlets_blocks_lbs"], out["outlets_blocks_idxs"] = flw.basin_outlets(blocks)
    out["bounds_basins_lbs"], out["bounds_basins_boxes"], out["bounds_basins_total"] = flw.basin_bounds()
    out["bounds_labels_lbs"], out["bounds_labels_boxes"], out["bounds_labels_total"] = flw.basin_bounds(basins=labels)
    # stream segments and flow-direction vectors as geo-features
    _feats_arrays(flw.streams(min_sto=2, strord=out["strord"], uparea=out["uparea_cell"]), "streams_so2", out, props=("strord", "uparea"))
    if d8.size <= FEATS_MAX_CELLS:  # one Python dict per segment / cell: small rasters only
        _feats_arrays(flw.streams(max_len=7), "streams_len7", out)
        _feats_arrays(flw.streams(mask=aux["smask"], max_len=3), "streams_gaps", out)
        _feats_arrays(flw.vectorize(mask=aux["smask"]), "vector", out)
    # Pfafstetter subbasins
    for key, depth, upa_min in (("pfaf_d1", 1, 0.0), ("pfaf_d2", 2, 0.0), ("pfaf_d3_min", 3, 5.0)):
        out[key], out[key + "_idxs"] = flw.subbasins_pfafstetter(depth=depth, upa_min=upa_min)
    # rivers: estuaries and manning river depth
    idxs_pit = flw.idxs_pit
    rivwth32 = np.sqrt(np.abs(out["uparea_cell"]).astype(np.float32)) + aux["data_f32"]
    elev0 = aux["elevtn"] - np.float32(np.median(aux["elevtn"].ravel()[idxs_pit]))
    out["estuary_f32"] = flw.classify_estuaries(elev0, rivwth32)
    out["estuary_f64"] = flw.classify_estuaries(elev0, rivwth32.astype(np.float64), rivdst=flw.distnc.astype(np.float64) * 0.5,
                                                min_convergence=1e-3, max_elevtn=1.5)
    return out


# ----------------------------------------------------------------------------- dem.fill_depressions cases
def fill_cases():
    """name -> (elevation raster, kwargs) of the dem.fill_depressions parity cases: the reference's own test rasters
    (tests/conftest.py:57-60 rand(15, 10) seed 2345; tests/test_dem.py:13-22 Wang & Liu's example), ties (integer-valued
    and integer-typed rasters: every level is a flat), nodata holes, both connectivities and all outlet modes."""
    wl = np.array([[15, 15, 14, 15, 12, 6, 12], [14, 13, 10, 12, 15, 17, 15], [15, 15, 9, 11, 8, 15, 15],
                   [16, 17, 8, 16, 15, 7, 5], [19, 18, 19, 18, 17, 15, 14]])
    np.random.seed(2345)
    rand1510 = np.random.rand(15, 10)
    rng = np.random.default_rng(77)
    f32 = rng.random((40, 50), dtype=np.float32) * np.float32(100.0)
    f32_holes = f32.copy()
    f32_holes[rng.random(f32.shape) < 0.12] = -9999.0
    ties = rng.integers(0, 6, size=(30, 33)).astype(np.float32)
    ties_holes = ties.copy()
    ties_holes[rng.random(ties.shape) < 0.1] = -9999.0
    i32 = rng.integers(0, 25, size=(37, 29)).astype(np.int32)
    f64r = np.round(rng.random((33, 41)) * 20.0, 1)  # float64 with many float32-key ties
    nanr = rng.random((25, 31))
    nanr[rng.random(nanr.shape) < 0.15] = np.nan
    z = oracle.synth_elevation(96, 130, seed=5)
    zsea = np.where(z < np.quantile(z, 0.1), np.float32(-9999.0), z * np.float32(500.0)).astype(np.float32)
    cases = {
        "rand15x10": (rand1510, {}),
        "rand15x10_min": (rand1510, dict(outlets="min")),
        "wangliu_f32": (wl.astype(np.float32), {}),
        "wangliu_i32": (wl.astype(np.int32), {}),
        "wangliu_min": (wl.astype(np.float32), dict(outlets="min")),
        "wangliu_c4": (wl.astype(np.float32), dict(connectivity=4)),
        "f32_40x50": (f32, {}),
        "f32_holes": (f32_holes, {}),
        "f32_holes_c4": (f32_holes, dict(connectivity=4)),
        "f32_elvmax": (f32, dict(elv_max=30.0)),
        "f32_pits": (f32, dict(idxs_pit=np.array([7, 1033, 1999]))),
        "ties_f32": (ties, {}),
        "ties_f32_c4": (ties, dict(connectivity=4)),
        "ties_holes_min": (ties_holes, dict(outlets="min")),
        "ties_i32": (i32, {}),
        "round_f64": (f64r, {}),
        "nan_nodata": (nanr, dict(nodata=np.nan)),
        "synth_sea": (zsea, {}),
        "synth_f64": (z.astype(np.float64), {}),
    }
    return cases


def fill_mid_case():
    """512 x 768 synthetic terrain (hash-only golden)."""
    z = oracle.synth_elevation(512, 768, seed=31)
    return (z * np.float32(800.0)).astype(np.float32), {}
